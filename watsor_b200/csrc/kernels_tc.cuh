// kernels_tc.cuh -- bf16 tensor-core (tcgen05 + TMA + TMEM) GEMM path for the dense convolutions.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

struct TcLayerWeights {
  __nv_bfloat16* w = nullptr;  // [n_pad][K] K-major bf16 (B operand of the UMMA), device
  int n_pad = 0, k = 0;
  void* tmap_b = nullptr;      // host copy of the CUtensorMap for the weights
};

struct TcWeights {
  std::vector<TcLayerWeights> layers;  // indexed by layer number (empty entries for non-GEMM layers)
};

int tc_prepare_weights(const std::vector<wb_layer>& layers, const std::vector<wb_tensor_entry>& tensors,
                       const float* host_data, TcWeights* out, std::string* err);
void tc_free_weights(TcWeights* w);
int tc_launch_gemm(const LaunchCtx& lc, const TcWeights& tw, int layer_index, int n, const wb_layer& L,
                   const __nv_bfloat16* in, const float* scale, const float* offset, __nv_bfloat16* out, float* enc,
                   float* logits, int num_anchors, int num_classes_p1, std::string* err);
