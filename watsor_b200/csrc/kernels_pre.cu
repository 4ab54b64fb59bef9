// kernels_pre.cu -- K1 (resize + normalise) and the fused K1+K2 stem convolution.
//
// Restates graph nodes `Cast`, `Preprocessor/map/while/ResizeImage/resize/ResizeBilinear`
// (align_corners=false, half_pixel_centers=false), `Preprocessor/mul`, `Preprocessor/sub` and
// `FeatureExtractor/.../Conv2d_0/{Conv2D,BatchNorm,Relu6}` of the frozen graph that
// watsor/detection/tensorflow_cpu.py:114 runs.  Compiled with -fmad=false: the bilinear lerp is
// a chain of separately rounded fp32 ops, exactly like TF's CPU kernel, so K1 is bit-exact
// against the oracle.
#include "common.cuh"

struct AxisTap {
  int lo, hi;
  float lerp;
};

// TF legacy sampling: scale = in/(float)out; pos = dst*scale; lo = floor(pos);
// hi = min(ceil(pos), in-1); lerp = pos - floor(pos).
__device__ __forceinline__ float axis_scale(int in_size, int out_size) {
  return __fdiv_rn((float)in_size, (float)out_size);
}
__device__ __forceinline__ AxisTap axis_tap(int dst, int in_size, float scale) {
  float pos = __fmul_rn((float)dst, scale);
  float fl = floorf(pos);
  AxisTap t;
  t.lo = max((int)fl, 0);
  t.hi = min((int)ceilf(pos), in_size - 1);
  t.lerp = __fsub_rn(pos, fl);
  return t;
}

__device__ __forceinline__ float lerp_px(float tl, float tr, float bl, float br, float lx, float ly) {
  float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
  float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
  return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
}

// one resized + normalised pixel (3 channels)
__device__ __forceinline__ void resized_pixel(const uint8_t* __restrict__ img, int w, const AxisTap& ty,
                                              const AxisTap& tx, float mul, float sub, float* out3) {
  const uint8_t* r0 = img + (size_t)ty.lo * w * 3;
  const uint8_t* r1 = img + (size_t)ty.hi * w * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float tl = (float)__ldg(r0 + tx.lo * 3 + c), tr = (float)__ldg(r0 + tx.hi * 3 + c);
    float bl = (float)__ldg(r1 + tx.lo * 3 + c), br = (float)__ldg(r1 + tx.hi * 3 + c);
    float v = lerp_px(tl, tr, bl, br, tx.lerp, ty.lerp);
    out3[c] = __fsub_rn(__fmul_rn(mul, v), sub);
  }
}

// resized_pixel() in two halves, so that a thread can have the 12 byte loads of a second pixel in flight while it
// lerps the first one (same operations in the same order: bit-identical)
__device__ __forceinline__ void resized_pixel_load(const uint8_t* __restrict__ img, int w, const AxisTap& ty,
                                                   const AxisTap& tx, uint32_t (&raw)[12]) {
  const uint8_t* r0 = img + (size_t)ty.lo * w * 3;
  const uint8_t* r1 = img + (size_t)ty.hi * w * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    raw[c * 4 + 0] = __ldg(r0 + tx.lo * 3 + c);
    raw[c * 4 + 1] = __ldg(r0 + tx.hi * 3 + c);
    raw[c * 4 + 2] = __ldg(r1 + tx.lo * 3 + c);
    raw[c * 4 + 3] = __ldg(r1 + tx.hi * 3 + c);
  }
}
__device__ __forceinline__ void resized_pixel_lerp(const uint32_t (&raw)[12], float lx, float ly, float mul, float sub,
                                                   float* out3) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = lerp_px((float)raw[c * 4 + 0], (float)raw[c * 4 + 1], (float)raw[c * 4 + 2], (float)raw[c * 4 + 3], lx, ly);
    out3[c] = __fsub_rn(__fmul_rn(mul, v), sub);
  }
}

// ---------------------------------------------------------------------------------------------------
// Source staging.  `resized_pixel` above issues 12 dependent single-byte global loads per resized pixel; a CTA that
// builds a tile of resized pixels instead copies the source rows the tile samples into shared memory with aligned
// 16-byte loads (all in flight at once, every sector fetched once) and lerps out of shared memory.  The bytes are the
// same bytes and the lerp is the same code, so the result is bit-identical (tests/test_gpu_stages.py).
//
// Resized rows [ry_a, ry_b) x columns [rx_a, rx_b) of frame `fd` sample source columns [x_first, x_last] (lo/hi are
// monotone in the destination index) of either a contiguous range of source rows (small scale factors) or of the
// {lo, hi} row pair of every resized row (scale > 2: rows in between are never touched and are not fetched).
struct StagePlan {
  int x_first, seg;     // first source column, bytes per row segment
  int y_first;          // contiguous mode: first staged source row
  int ry_a;             // pair mode: staged row 2*(ry - ry_a) + {0, 1}
  int rows, cpr;        // staged rows, 16-byte chunks per row (upper bound)
  bool contiguous, on;
};

__device__ __forceinline__ StagePlan stage_plan(const FrameDesc& fd, int ry_a, int ry_b, int rx_a, int rx_b, float sy,
                                                float sx, int pitch, int max_rows) {
  StagePlan sp;
  sp.on = false;
  if (pitch <= 0 || ry_a >= ry_b || rx_a >= rx_b) return sp;
  sp.x_first = axis_tap(rx_a, fd.w, sx).lo;
  const int x_last = axis_tap(rx_b - 1, fd.w, sx).hi;
  sp.seg = (x_last - sp.x_first + 1) * 3;
  sp.y_first = axis_tap(ry_a, fd.h, sy).lo;
  const int y_last = axis_tap(ry_b - 1, fd.h, sy).hi;
  sp.ry_a = ry_a;
  sp.contiguous = y_last - sp.y_first + 1 <= max_rows;
  sp.rows = sp.contiguous ? y_last - sp.y_first + 1 : 2 * (ry_b - ry_a);
  sp.cpr = (sp.seg + 30) >> 4;
  sp.on = sp.seg + 30 <= pitch && sp.rows <= max_rows;
  return sp;
}

__device__ __forceinline__ const uint8_t* stage_src(const FrameDesc& fd, const StagePlan& sp, int r, float sy) {
  int y;
  if (sp.contiguous) {
    y = sp.y_first + r;
  } else {
    AxisTap t = axis_tap(sp.ry_a + (r >> 1), fd.h, sy);
    y = (r & 1) ? t.hi : t.lo;
  }
  return fd.ptr + ((size_t)y * fd.w + sp.x_first) * 3;
}

// all threads of the CTA; caller synchronises afterwards
__device__ __forceinline__ void stage_rows(const FrameDesc& fd, const StagePlan& sp, float sy, uint8_t* s_stage,
                                           int pitch, int tid, int nthreads) {
  const uint8_t* f_begin = fd.ptr;
  const uint8_t* f_end = fd.ptr + (size_t)fd.h * fd.w * 3;
  for (int i = tid; i < sp.rows * sp.cpr; i += nthreads) {
    const int r = i / sp.cpr, k = i - r * sp.cpr;
    const uint8_t* src = stage_src(fd, sp, r, sy);
    const uint8_t* g = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15) + 16 * k;
    if (g >= src + sp.seg) continue;
    uint4 v;
    if (g >= f_begin && g + 16 <= f_end) {
      v = __ldg(reinterpret_cast<const uint4*>(g));
    } else {  // first / last chunk of the frame: never touch bytes outside the buffer
      uint32_t wds[4] = {0u, 0u, 0u, 0u};
      for (int b = 0; b < 16; ++b)
        if (g + b >= f_begin && g + b < f_end) wds[b >> 2] |= (uint32_t)__ldg(g + b) << (8 * (b & 3));
      v = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
    *reinterpret_cast<uint4*>(s_stage + (size_t)r * pitch + 16 * k) = v;
  }
}

// resized_pixel() reading the staged rows
__device__ __forceinline__ void resized_pixel_staged(const FrameDesc& fd, const StagePlan& sp, const uint8_t* s_stage,
                                                     int pitch, int ry, const AxisTap& ty, const AxisTap& tx, float mul,
                                                     float sub, float* out3) {
  const int r0 = sp.contiguous ? ty.lo - sp.y_first : 2 * (ry - sp.ry_a);
  const int r1 = sp.contiguous ? ty.hi - sp.y_first : 2 * (ry - sp.ry_a) + 1;
  const uintptr_t base = reinterpret_cast<uintptr_t>(fd.ptr) + (size_t)sp.x_first * 3;
  const size_t row_bytes = (size_t)fd.w * 3;
  const int sh0 = (int)((base + (size_t)ty.lo * row_bytes) & 15), sh1 = (int)((base + (size_t)ty.hi * row_bytes) & 15);
  const uint8_t* p0 = s_stage + (size_t)r0 * pitch + sh0;
  const uint8_t* p1 = s_stage + (size_t)r1 * pitch + sh1;
  const int xl = (tx.lo - sp.x_first) * 3, xh = (tx.hi - sp.x_first) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float tl = (float)p0[xl + c], tr = (float)p0[xh + c];
    float bl = (float)p1[xl + c], br = (float)p1[xh + c];
    float v = lerp_px(tl, tr, bl, br, tx.lerp, ty.lerp);
    out3[c] = __fsub_rn(__fmul_rn(mul, v), sub);
  }
}

// bytes per staged row for frames up to `max_src_w` wide resized to `in_w`, tile of `tile_w` resized columns; 0 = do
// not stage (unknown width, or the rows would not fit)
static int stage_pitch_for(int max_src_w, int in_w, int tile_w, int rows) {
  // opt-in (WB_STAGE=1): measured slower than the direct path on B200 (profiles/r02_stem.md) -- the byte loads hit L1
  // and the kernel is issue bound, not latency bound
  const char* on = getenv("WB_STAGE");
  if (max_src_w <= 0 || on == nullptr || on[0] != '1' || getenv("WB_NO_STAGE")) return 0;
  const double scale = (double)max_src_w / in_w;
  int px = (int)(tile_w * (scale > 1.0 ? scale : 1.0)) + 4;
  if (px > max_src_w) px = max_src_w;
  int pitch = (px * 3 + 30 + 15) & ~15;
  return (size_t)pitch * rows <= 96 * 1024 ? pitch : 0;
}

// ---------------------------------------------------------------------------------------------------
// K1 stand-alone: u8 HWC (any size) -> f32 [n][oh][ow][3].  Used by wb_preprocess (parity tests)
// and by wb_backbone-less debugging; the production path is the fused stem below.
constexpr int PP_TY = 8, PP_TX = 32;  // resized pixels per CTA of the stand-alone kernel

__global__ void __launch_bounds__(PP_TY* PP_TX) k_preprocess_f32(const FrameDesc* __restrict__ frames,
                                                                  float* __restrict__ out, int oh, int ow, float mul,
                                                                  float sub, int pitch) {
  extern __shared__ uint4 s_dyn[];
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_dyn);
  const FrameDesc fd = frames[blockIdx.z];
  const float sy = axis_scale(fd.h, oh), sx = axis_scale(fd.w, ow);
  const int oy0 = blockIdx.y * PP_TY, ox0 = blockIdx.x * PP_TX;
  const StagePlan sp = stage_plan(fd, oy0, min(oy0 + PP_TY, oh), ox0, min(ox0 + PP_TX, ow), sy, sx, pitch, 2 * PP_TY);
  if (sp.on) {
    stage_rows(fd, sp, sy, s_stage, pitch, threadIdx.x, PP_TY * PP_TX);
    __syncthreads();
  }
  const int oy = oy0 + (int)threadIdx.x / PP_TX, ox = ox0 + (int)threadIdx.x % PP_TX;
  if (oy >= oh || ox >= ow) return;
  AxisTap ty = axis_tap(oy, fd.h, sy), tx = axis_tap(ox, fd.w, sx);
  float v[3];
  if (sp.on)
    resized_pixel_staged(fd, sp, s_stage, pitch, oy, ty, tx, mul, sub, v);
  else
    resized_pixel(fd.ptr, fd.w, ty, tx, mul, sub, v);
  float* o = out + (((size_t)blockIdx.z * oh + oy) * ow + ox) * 3;
  o[0] = v[0];
  o[1] = v[1];
  o[2] = v[2];
}

void launch_preprocess_f32(const LaunchCtx& lc, const FrameDesc* frames, int n, float* out, int oh, int ow,
                           float mul, float sub, int max_src_w) {
  dim3 grid((ow + PP_TX - 1) / PP_TX, (oh + PP_TY - 1) / PP_TY, n);
  const int pitch = stage_pitch_for(max_src_w, ow, PP_TX, 2 * PP_TY);
  const size_t smem = (size_t)pitch * 2 * PP_TY;
  static PerDeviceFlag attr_done;
  if (!attr_done.get()) {
    cudaFuncSetAttribute(k_preprocess_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_done.set();
  }
  k_preprocess_f32<<<grid, PP_TY * PP_TX, smem, lc.stream>>>(frames, out, oh, ow, mul, sub, pitch);
  ++*lc.launch_counter;
}

// ---------------------------------------------------------------------------------------------------
// Fused stem: resized tile built in shared memory (the 300x300x3 tensor never reaches HBM), then the
// first KxK stride-S convolution (C_in = 3) + folded BatchNorm + ReLU6.
// Tile = 8 x 32 output pixels, one thread per output pixel, output channels in chunks of 16.
constexpr int ST_TY = 8, ST_TX = 32;

template <typename T>
__global__ void __launch_bounds__(ST_TY* ST_TX)
    k_stem(const FrameDesc* __restrict__ frames, const float* __restrict__ pre, wb_layer L, int in_h, int in_w,
           float mul, float sub, const float* __restrict__ w, const float* __restrict__ scale,
           const float* __restrict__ offset, T* __restrict__ out, int pitch) {
  extern __shared__ float smem[];
  const int K = L.kh, S = L.stride;
  const int tile_h = (ST_TY - 1) * S + K, tile_w = (ST_TX - 1) * S + K;
  float* s_in = smem;                              // [tile_h][tile_w][3]
  float* s_w = smem + ((tile_h * tile_w * 3 + 3) & ~3);  // [K*K*3][n_pad], 16-byte aligned
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_w + ((K * K * 3 * (int)L.n_pad + 3) & ~3));  // [2*tile_h][pitch]
  const int f = blockIdx.z;
  const int oy0 = blockIdx.y * ST_TY, ox0 = blockIdx.x * ST_TX;
  const int tid = threadIdx.x;

  for (int i = tid; i < K * K * 3 * (int)L.n_pad; i += blockDim.x) s_w[i] = w[i];

  // resized tile; rows/cols outside the 300x300 image are the SAME-padding zeros
  const int ry0 = oy0 * S - (int)L.pad_t, rx0 = ox0 * S - (int)L.pad_l;
  FrameDesc fd;
  float sy = 1.f, sx = 1.f;
  StagePlan sp;
  sp.on = false;
  if (pre == nullptr) {
    fd = frames[f];
    sy = axis_scale(fd.h, in_h);  // hoisted: one division per thread, not per sampled pixel
    sx = axis_scale(fd.w, in_w);
    sp = stage_plan(fd, max(ry0, 0), min(ry0 + tile_h, in_h), max(rx0, 0), min(rx0 + tile_w, in_w), sy, sx, pitch,
                    2 * tile_h);
    if (sp.on) {
      stage_rows(fd, sp, sy, s_stage, pitch, tid, (int)blockDim.x);
      __syncthreads();
    }
  }
  for (int i = tid; i < tile_h * tile_w; i += blockDim.x) {
    int ly = i / tile_w, lx = i - ly * tile_w;
    int ry = ry0 + ly, rx = rx0 + lx;
    float v[3] = {0.f, 0.f, 0.f};
    if (ry >= 0 && ry < in_h && rx >= 0 && rx < in_w) {
      if (pre != nullptr) {
        const float* p = pre + (((size_t)f * in_h + ry) * in_w + rx) * 3;
        v[0] = p[0];
        v[1] = p[1];
        v[2] = p[2];
      } else {
        AxisTap ty = axis_tap(ry, fd.h, sy), tx = axis_tap(rx, fd.w, sx);
        if (sp.on)
          resized_pixel_staged(fd, sp, s_stage, pitch, ry, ty, tx, mul, sub, v);
        else
          resized_pixel(fd.ptr, fd.w, ty, tx, mul, sub, v);
      }
    }
    s_in[i * 3 + 0] = v[0];
    s_in[i * 3 + 1] = v[1];
    s_in[i * 3 + 2] = v[2];
  }
  __syncthreads();

  const int ty = tid / ST_TX, tx = tid - ty * ST_TX;
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy >= (int)L.out_h || ox >= (int)L.out_w) return;
  T* o = out + (((size_t)f * L.out_h + oy) * L.out_w + ox) * L.out_c;
  const int taps = K * K * 3;
  for (int oc0 = 0; oc0 < (int)L.out_c; oc0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < K; ++ky)
      for (int kx = 0; kx < K; ++kx) {
        const float* ip = s_in + ((ty * S + ky) * tile_w + tx * S + kx) * 3;
        const float* wp = s_w + (size_t)((ky * K + kx) * 3) * L.n_pad + oc0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float x = ip[c];
          const float4* w4 = reinterpret_cast<const float4*>(wp + c * L.n_pad);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 ww = w4[q];
            acc[q * 4 + 0] = fmaf(x, ww.x, acc[q * 4 + 0]);
            acc[q * 4 + 1] = fmaf(x, ww.y, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(x, ww.z, acc[q * 4 + 2]);
            acc[q * 4 + 3] = fmaf(x, ww.w, acc[q * 4 + 3]);
          }
        }
      }
    (void)taps;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int oc = oc0 + q * 4;
      if (oc >= (int)L.out_c) break;
      float4 v;
      v.x = affine_rn(acc[q * 4 + 0], scale[oc + 0], offset[oc + 0]);
      v.y = affine_rn(acc[q * 4 + 1], scale[oc + 1], offset[oc + 1]);
      v.z = affine_rn(acc[q * 4 + 2], scale[oc + 2], offset[oc + 2]);
      v.w = affine_rn(acc[q * 4 + 3], scale[oc + 3], offset[oc + 3]);
      if (L.act == WB_ACT_RELU6) {
        v.x = relu6f(v.x);
        v.y = relu6f(v.y);
        v.z = relu6f(v.z);
        v.w = relu6f(v.w);
      }
      ActIO<T>::st4(o + oc, v);
    }
  }
}

// Register-tiled variant for the common 3x3 / stride-2 / 32-channel stem: one thread = 4 consecutive output
// pixels x 8 output channels, so every weight vector read from shared memory feeds 4 pixels (the generic
// kernel re-reads all 27x32 weights per pixel: `LDS.128` broadcasts cost 4 wavefronts each and dominated).
// Same tile (8 x 32 pixels), same resize code, same (ky, kx, c) accumulation order => identical results.
template <typename T>
__global__ void __launch_bounds__(256)
    k_stem_3x3s2_c32(const FrameDesc* __restrict__ frames, const float* __restrict__ pre, wb_layer L, int in_h, int in_w,
                     float mul, float sub, const float* __restrict__ w, const float* __restrict__ scale,
                     const float* __restrict__ offset, T* __restrict__ out, int pitch) {
  extern __shared__ uint4 s_dyn[];  // [2*tile_h][pitch] staged source rows
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_dyn);
  constexpr int K = 3, S = 2, OC = 32;
  constexpr int tile_h = (ST_TY - 1) * S + K, tile_w = (ST_TX - 1) * S + K;  // 17 x 65
  __shared__ __align__(16) float s_in[tile_h * tile_w * 3];
  __shared__ __align__(16) float s_w[K * K * 3 * OC];
  const int f = blockIdx.z;
  const int oy0 = blockIdx.y * ST_TY, ox0 = blockIdx.x * ST_TX;
  const int tid = threadIdx.x;
  // the 864 weights go through registers: the loads are issued here and waited for after the tile has been built
  // (13 % of the kernel's stall samples sat on this copy when it stored right away)
  constexpr int WREGS = (K * K * 3 * OC + 255) / 256;
  float wreg[WREGS];
#pragma unroll
  for (int j = 0; j < WREGS; ++j) {
    const int i = tid + j * 256;
    wreg[j] = i < K * K * 3 * OC ? __ldg(w + (i / OC) * L.n_pad + (i % OC)) : 0.f;
  }
  const int ry0 = oy0 * S - (int)L.pad_t, rx0 = ox0 * S - (int)L.pad_l;
  FrameDesc fd;
  float sy = 1.f, sx = 1.f;
  StagePlan sp;
  sp.on = false;
  if (pre == nullptr) {
    fd = frames[f];
    sy = axis_scale(fd.h, in_h);
    sx = axis_scale(fd.w, in_w);
    sp = stage_plan(fd, max(ry0, 0), min(ry0 + tile_h, in_h), max(rx0, 0), min(rx0 + tile_w, in_w), sy, sx, pitch,
                    2 * tile_h);
    if (sp.on) {
      stage_rows(fd, sp, sy, s_stage, pitch, tid, 256);
      __syncthreads();
    }
  }
  if (pre == nullptr && !sp.on) {
    // two pixels per round: the byte loads of both are in flight before either is interpolated
    for (int i0 = tid; i0 < tile_h * tile_w; i0 += 512) {
      uint32_t raw[2][12];
      AxisTap tys[2], txs[2];
      bool inside[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = i0 + u * 256;
        const int ly = i / tile_w, lx = i - ly * tile_w;
        const int ry = ry0 + ly, rx = rx0 + lx;
        inside[u] = i < tile_h * tile_w && ry >= 0 && ry < in_h && rx >= 0 && rx < in_w;
        if (inside[u]) {
          tys[u] = axis_tap(ry, fd.h, sy);
          txs[u] = axis_tap(rx, fd.w, sx);
          resized_pixel_load(fd.ptr, fd.w, tys[u], txs[u], raw[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = i0 + u * 256;
        if (i < tile_h * tile_w) {
          float v[3] = {0.f, 0.f, 0.f};
          if (inside[u]) resized_pixel_lerp(raw[u], txs[u].lerp, tys[u].lerp, mul, sub, v);
          s_in[i * 3 + 0] = v[0];
          s_in[i * 3 + 1] = v[1];
          s_in[i * 3 + 2] = v[2];
        }
      }
    }
  } else
  for (int i = tid; i < tile_h * tile_w; i += 256) {
    int ly = i / tile_w, lx = i - ly * tile_w;
    int ry = ry0 + ly, rx = rx0 + lx;
    float v[3] = {0.f, 0.f, 0.f};
    if (ry >= 0 && ry < in_h && rx >= 0 && rx < in_w) {
      if (pre != nullptr) {
        const float* p = pre + (((size_t)f * in_h + ry) * in_w + rx) * 3;
        v[0] = p[0];
        v[1] = p[1];
        v[2] = p[2];
      } else {
        AxisTap ty = axis_tap(ry, fd.h, sy), tx = axis_tap(rx, fd.w, sx);
        if (sp.on)
          resized_pixel_staged(fd, sp, s_stage, pitch, ry, ty, tx, mul, sub, v);
        else
          resized_pixel(fd.ptr, fd.w, ty, tx, mul, sub, v);
      }
    }
    s_in[i * 3 + 0] = v[0];
    s_in[i * 3 + 1] = v[1];
    s_in[i * 3 + 2] = v[2];
  }
#pragma unroll
  for (int j = 0; j < WREGS; ++j) {
    const int i = tid + j * 256;
    if (i < K * K * 3 * OC) s_w[i] = wreg[j];
  }
  __syncthreads();

  const int ocg = tid & 3, quad = tid >> 2;  // 4 channel groups of 8, 64 pixel quads (8 rows x 8 quads)
  const int ty = quad >> 3, tx0 = (quad & 7) * 4;
  float acc[4][8];
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
#pragma unroll
  for (int ky = 0; ky < K; ++ky)
#pragma unroll
    for (int kx = 0; kx < K; ++kx)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[((ky * K + kx) * 3 + c) * OC + ocg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[((ky * K + kx) * 3 + c) * OC + ocg * 8 + 4]);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const float x = s_in[((ty * S + ky) * tile_w + (tx0 + o) * S + kx) * 3 + c];
          acc[o][0] = fmaf(x, w0.x, acc[o][0]);
          acc[o][1] = fmaf(x, w0.y, acc[o][1]);
          acc[o][2] = fmaf(x, w0.z, acc[o][2]);
          acc[o][3] = fmaf(x, w0.w, acc[o][3]);
          acc[o][4] = fmaf(x, w1.x, acc[o][4]);
          acc[o][5] = fmaf(x, w1.y, acc[o][5]);
          acc[o][6] = fmaf(x, w1.z, acc[o][6]);
          acc[o][7] = fmaf(x, w1.w, acc[o][7]);
        }
      }
  const int oy = oy0 + ty;
  if (oy >= (int)L.out_h) return;
  float sc[8], of[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(scale + ocg * 8 + j);
    of[j] = __ldg(offset + ocg * 8 + j);
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int ox = ox0 + tx0 + o;
    if (ox >= (int)L.out_w) break;
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = affine_rn(acc[o][j], sc[j], of[j]);
      y[j] = L.act == WB_ACT_RELU6 ? relu6f(v) : v;
    }
    T* dst = out + (((size_t)f * L.out_h + oy) * L.out_w + ox) * OC + ocg * 8;
    ActIO<T>::st4(dst, make_float4(y[0], y[1], y[2], y[3]));
    ActIO<T>::st4(dst + 4, make_float4(y[4], y[5], y[6], y[7]));
  }
}

template <typename T>
void launch_stem(const LaunchCtx& lc, const FrameDesc* frames, const float* pre, int n, const wb_layer& L,
                 int in_h, int in_w, float mul, float sub, const float* w, const float* scale,
                 const float* offset, T* out, int max_src_w) {
  const int tile_h = (ST_TY - 1) * L.stride + L.kh, tile_w = (ST_TX - 1) * L.stride + L.kw;
  const int pitch = pre == nullptr ? stage_pitch_for(max_src_w, in_w, tile_w, 2 * tile_h) : 0;
  const size_t stage = (size_t)pitch * 2 * tile_h;
  dim3 grid((L.out_w + ST_TX - 1) / ST_TX, (L.out_h + ST_TY - 1) / ST_TY, n);
  if (L.kh == 3 && L.kw == 3 && L.stride == 2 && L.out_c == 32) {
    static PerDeviceFlag attr3;
    if (!attr3.get()) {
      cudaFuncSetAttribute(k_stem_3x3s2_c32<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr3.set();
    }
    k_stem_3x3s2_c32<T><<<grid, 256, stage, lc.stream>>>(frames, pre, L, in_h, in_w, mul, sub, w, scale, offset, out,
                                                        pitch);
    ++*lc.launch_counter;
    return;
  }
  const size_t smem = ((((size_t)tile_h * tile_w * 3 + 3) & ~(size_t)3) +
                       (((size_t)L.kh * L.kw * 3 * L.n_pad + 3) & ~(size_t)3)) * sizeof(float) + stage;
  static PerDeviceFlag attr_done;
  if (!attr_done.get()) {
    cudaFuncSetAttribute(k_stem<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done.set();
  }
  k_stem<T><<<grid, ST_TY * ST_TX, smem, lc.stream>>>(frames, pre, L, in_h, in_w, mul, sub, w, scale, offset, out,
                                                     pitch);
  ++*lc.launch_counter;
}

template void launch_stem<float>(const LaunchCtx&, const FrameDesc*, const float*, int, const wb_layer&, int,
                                 int, float, float, const float*, const float*, const float*, float*, int);
template void launch_stem<__nv_bfloat16>(const LaunchCtx&, const FrameDesc*, const float*, int,
                                         const wb_layer&, int, int, float, float, const float*, const float*,
                                         const float*, __nv_bfloat16*, int);
