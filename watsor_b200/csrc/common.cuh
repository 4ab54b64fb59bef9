// common.cuh -- shared device-side types of libwatsor_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/watsor_b200.h"
#include "model_format.h"

// One frame of a batch: where its RGB24 pixels are and which camera it belongs to.
// Replaces the (image_shape, image_np) pair of ObjectDetector.detect (tensorflow_cpu.py:74).
struct FrameDesc {
  const uint8_t* ptr;  // device pointer, H*W*3 bytes, row-major RGB24 (share.py:68-73)
  int32_t w, h;
  int32_t cam;
  int32_t _pad;
};

// Per-camera filter state resident in HBM (ConfidenceFilter / AreaFilter / MaskFilter __init__).
#define WB_MAX_LABELS 128
struct CameraCfg {
  int32_t width, height;
  int32_t n_zones;
  int32_t has_mask;
  int32_t check_label;      // require label > 0 (track.py:26)
  int32_t default_present;  // entry used for labels without one of their own
  int32_t default_has_zone_list;
  uint32_t default_zone_bits;
  double default_conf, default_area;
  const int32_t* sat;  // [n_zones][(height+1)*(width+1)] inclusive-exclusive summed-area tables
  double conf[WB_MAX_LABELS];
  double area[WB_MAX_LABELS];
  uint32_t zone_bits[WB_MAX_LABELS];
  uint8_t present[WB_MAX_LABELS];
  uint8_t has_zone_list[WB_MAX_LABELS];
};

struct PostParams {
  int32_t num_anchors, num_classes;  // classes without background
  float scale_y, scale_x, scale_h, scale_w, logit_scale;
  float iou_thr, score_thr;
  int32_t max_per_class, max_total;
  float class_offset;
};

// activation element type helpers (fp32 parity path / bf16 tensor-core path)
template <typename T> struct ActIO;
template <> struct ActIO<float> {
  static __device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct ActIO<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ float relu6f(float v) { return fminf(fmaxf(v, 0.0f), 6.0f); }

// y = acc*scale + offset exactly as two roundings (the graph runs Conv2D and
// FusedBatchNormV3 / BiasAdd as separate fp32 ops)
__device__ __forceinline__ float affine_rn(float acc, float s, float o) {
  return __fadd_rn(__fmul_rn(acc, s), o);
}

// ---- launchers implemented in the .cu files -------------------------------------------------------
// cudaFuncSetAttribute applies to the current device only: "already done" is remembered per device, so that
// several contexts on different GPUs can live in one process (the reference runs one process per detector,
// detector.py:12-55, but nothing in the C-ABI forbids the other arrangement).
struct PerDeviceFlag {
  bool done[64] = {};
  static int dev() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 63;
  }
  bool get() const { return done[dev()]; }
  void set() { done[dev()] = true; }
};

struct LaunchCtx {
  cudaStream_t stream;
  int* launch_counter;  // host-side counter of kernel launches (bench: gpu_launches)
};

void launch_preprocess_f32(const LaunchCtx& lc, const FrameDesc* frames, int n, float* out, int oh, int ow,
                           float mul, float sub, int max_src_w);
// max_src_w: widest source frame the launch may see (sizes the shared-memory staging of source rows; frames that do
// not fit take the direct path inside the kernel); 0 = never stage
template <typename T>
void launch_stem(const LaunchCtx& lc, const FrameDesc* frames, const float* pre, int n, const wb_layer& L,
                 int in_h, int in_w, float mul, float sub, const float* w, const float* scale,
                 const float* offset, T* out, int max_src_w);
template <typename T>
void launch_dw(const LaunchCtx& lc, int n, const wb_layer& L, const T* in, const float* w, const float* scale,
               const float* offset, T* out);
template <typename T>
void launch_add(const LaunchCtx& lc, size_t elems, const T* a, const T* b, T* out);
template <typename T>
void launch_pool(const LaunchCtx& lc, int n, const wb_layer& L, const T* in, T* out);
template <typename T>
void launch_copy_channels(const LaunchCtx& lc, int n, const wb_layer& L, const T* in, T* out);
struct SplitKReduceArgs {
  const float* partial;  // [splits][M][ld] raw accumulators
  const float* scale;
  const float* offset;
  void* out;  // fp32 or bf16 [M][N]
  int out_is_bf16;
  float* enc;
  float* logits;
  int M, N, ld, splits, act;
  int is_head, anchors_per_loc, row_off, n_box, num_anchors, ncp1, hw;
};
void launch_splitk_reduce(const LaunchCtx& lc, const SplitKReduceArgs& r);
template <typename T>
void launch_gemm_cc(const LaunchCtx& lc, int n, const wb_layer& L, const T* in, const float* w,
                    const float* scale, const float* offset, T* out, float* enc, float* logits,
                    int num_anchors, int num_classes_p1, float* partial, size_t partial_floats);
void launch_post(const LaunchCtx& lc, int n, const PostParams& pp, const float* enc, const float* logits,
                 const float* anchors, const FrameDesc* frames, const CameraCfg* cams, uint32_t flags,
                 float* dec_boxes, int* cand_count, unsigned long long* cand, int* sel_count,
                 unsigned long long* sel, wb_detection* out, uint32_t* verdicts, float* raw_boxes,
                 float* raw_scores, float* raw_classes, int* raw_num, int* kept_hist);
void launch_filter_rows(const LaunchCtx& lc, const CameraCfg* cam, int n_rows, wb_detection* rows,
                        uint32_t* verdicts);
void launch_build_sat(const LaunchCtx& lc, const uint8_t* raster, int n_zones, int h, int w, int32_t* sat);
