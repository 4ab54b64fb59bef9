// kernels_tc.cuh -- tensor-core (tcgen05 + TMA + TMEM) GEMM path for the dense 1x1 convolutions.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

enum TcMode { TC_BF16 = 0, TC_TF32X1 = 1, TC_TF32X3 = 2 };

struct TcLayerWeights {
  void* w = nullptr;     // [n_pad][K] K-major (bf16, or fp32 "hi" part), device
  void* w_lo = nullptr;  // fp32 "lo" part (TF32X3)
  int n_pad = 0, k = 0, block_n = 0;
  bool ready = false;
  alignas(64) unsigned char tmap_b[128];     // CUtensorMap of w
  alignas(64) unsigned char tmap_b_lo[128];  // CUtensorMap of w_lo
};

struct TcWeights {
  int mode = TC_BF16;
  std::vector<TcLayerWeights> layers;  // indexed by layer number (unset entries for non-GEMM layers)
};

bool tc_layer_supported(const wb_layer& L);
int tc_prepare_weights(const std::vector<wb_layer>& layers, const std::vector<wb_tensor_entry>& tensors,
                       const float* host_data, int mode, TcWeights* out, std::string* err);
void tc_free_weights(TcWeights* w);
int tc_launch_gemm(const LaunchCtx& lc, const TcWeights& tw, int layer_index, int n, const wb_layer& L, const void* in,
                   const float* scale, const float* offset, void* out, float* enc, float* logits, int num_anchors,
                   int num_classes_p1, float* partial, size_t partial_floats, int* tile_counters, const void* residual,
                   std::string* err);

// generic tiled tensor-map encoder (rank <= 5); `map` points to 128 bytes aligned to 64
bool tc_encode_map(void* map, const void* base, int elem_bytes, int rank, const unsigned long long* dims,
                   const unsigned long long* strides_bytes, const unsigned* box, bool swizzle128, std::string* err,
                   const unsigned* elem_strides = nullptr);

// fused depthwise 3x3 (+BN+ReLU6) -> 1x1 conv (+BN+ReLU6) on tensor cores (kernels_fused.cu); TF32X3 only
bool fused_dwpw_supported(const TcWeights& tw, int pw_layer_index, const wb_layer& dw, const wb_layer& pw, int n);
int fused_launch_dwpw(const LaunchCtx& lc, const TcWeights& tw, int pw_layer_index, int n, const wb_layer& dw,
                      const wb_layer& pw, const void* in, const float* dw_w, const float* dw_scale, const float* dw_offset,
                      const float* scale, const float* offset, void* out, std::string* err);

// MobileNet-v2 inverted residual block (1x1 expand -> depthwise 3x3 -> linear 1x1 projection [-> Add]) as one kernel
// (kernels_fused.cu: k_irb_x3); TF32X3 only.  `add` may be NULL.
bool fused_irb_supported(const TcWeights& tw, int pw_layer_index, const wb_layer& ex, const wb_layer& dw, const wb_layer& pw,
                         const wb_layer* add, int n);
int fused_launch_irb(const LaunchCtx& lc, const TcWeights& tw, int pw_layer_index, int n, const wb_layer& ex,
                     const wb_layer& dw, const wb_layer& pw, bool with_add, const void* in, const float* ex_w,
                     const float* ex_scale, const float* ex_offset, const float* dw_w, const float* dw_scale,
                     const float* dw_offset, const float* scale, const float* offset, void* out, std::string* err);
