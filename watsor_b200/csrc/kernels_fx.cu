// kernels_fx.cu -- the reference's output-stage visual effects as one CUDA pass per batch of frames
// (SURVEY.md section 8 (f)4), behind the wb_fx_* entry points of include/watsor_b200.h:
//
//   CopyImageEffect            watsor/output/copy.py:14-18     image_out = image_in
//   BlendEffect                watsor/output/blend.py:8-32     image_out = u8(f32(image_in) * a/255 + 255 * (1 - a/255))
//   DrawEffect                 watsor/output/draw.py:9-88      per detection with label > 0, in row order:
//                                                              cv2.rectangle, the alpha-blended label box
//                                                              (cv2.addWeighted) and the label text (cv2.putText)
//   DrawEffectWithContours     watsor/output/draw.py:91-103    + the zone contours of the detections' zones
//
// Every byte equals what the reference computes with numpy / OpenCV on the CPU (tests/test_gpu_effects.py).  Where
// the arithmetic is OpenCV's, it is either restated after being pinned on all inputs (addWeighted: one fused
// multiply-add in float, round half to even) or taken over by construction as tables the host builds with the
// installed OpenCV (glyph tables: watsor_b200/output/font.py; contour pixels: cv2.drawContours on an empty raster).
// Compiled with the default -fmad=true: the float arithmetic that must not be contracted uses the _rn intrinsics.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/watsor_b200.h"

namespace {

thread_local std::string g_fx_err;
int fx_fail(const std::string& m) {
  g_fx_err = m;
  return 1;
}

constexpr int FX_MAX_GLYPHS = 48;  // "traffic light: 100%" is 19; prefix <= 40 + up to 7 digits + '%'

struct FxDet {                 // one drawable detection, computed by k_fx_prepare
  int32_t x0, y0, x1, y1;      // rectangle corners, min/max ordered (cv2.rectangle draws the same pixels either way)
  int32_t bx0, by0, bx1, by1;  // label box [bx0, bx1) x [by0, by1), clamped to the image; empty: no box and no text
  int32_t ox, oy;              // text origin (cv2.putText org)
  int32_t style;               // label style index
  int32_t n_glyphs;
  int32_t text_x1;             // one past the last column any glyph can touch
  uint8_t glyph[FX_MAX_GLYPHS];
  uint16_t pen2[FX_MAX_GLYPHS];  // pen offset from ox in half pixels
};

struct FxFrame {           // per frame, after k_fx_prepare
  int32_t n_active;        // detections with label > 0, in row order
  uint32_t zone_sel;       // bit z-1: some active detection lists zone z (draw.py:100-103)
  uint8_t order[WB_MAX_DETECTIONS];
  FxDet det[WB_MAX_DETECTIONS];
};

struct FxFont {  // device copies
  int32_t n_glyphs, rows, cols, y0, text_height, baseline, margin;
  const int32_t* advance;
  const uint8_t* lut;  // [n_glyphs][2][cols + 1][rows][cols][256]
};

struct FxLabel {
  uint8_t box_color[3];
  uint8_t n_prefix;
  uint8_t prefix[60];
};

struct FxCamera {
  int32_t w, h;
  const uint8_t* alpha;      // [h][w] or nullptr
  const uint32_t* contours;  // [h][w] or nullptr
};

struct FxFrameDesc {
  const uint8_t* in;
  uint8_t* out;
  FxCamera cam;
};

// ---------------------------------------------------------------------------------------------------
// one block per frame, thread t = detection row t.  Restates the geometry of DrawEffect._draw (draw.py:51-88).
__global__ void __launch_bounds__(128)
    k_fx_prepare(const wb_detection* __restrict__ rows, const FxFrameDesc* __restrict__ frames, FxFont font,
                 const FxLabel* __restrict__ labels, int n_labels, const uint8_t* __restrict__ digit_glyphs,
                 FxFrame* __restrict__ out) {
  const int f = blockIdx.x, t = threadIdx.x;
  const int W = frames[f].cam.w, H = frames[f].cam.h;
  FxFrame& fr = out[f];
  __shared__ uint32_t s_zone;
  __shared__ uint32_t s_active[4];
  if (t == 0) s_zone = 0u;
  __syncthreads();
  bool active = false;
  if (t < WB_MAX_DETECTIONS) {
    const wb_detection d = rows[(size_t)f * WB_MAX_DETECTIONS + t];
    active = d.label > 0;  // draw.py:13 `filter(lambda d: d.label > 0, ...)`
    if (active) {
      FxDet& r = fr.det[t];
      const int left = d.bounding_box.x_min, top = d.bounding_box.y_min;
      const int right = d.bounding_box.x_max, bottom = d.bounding_box.y_max;
      r.x0 = min(left, right);
      r.x1 = max(left, right);
      r.y0 = min(top, bottom);
      r.y1 = max(top, bottom);
      const int style = d.label < n_labels ? d.label : 0;  // coco.py:124-131: unknown index -> 'unlabeled'
      r.style = style;
      // display_str = "{}: {}".format(label, "{0:.0%}".format(confidence))      draw.py:15
      const FxLabel lb = labels[style];
      int ng = 0, pen = 0;
      for (int i = 0; i < lb.n_prefix && ng < FX_MAX_GLYPHS; ++i) {
        r.glyph[ng] = lb.prefix[i];
        r.pen2[ng] = (uint16_t)pen;
        pen += font.advance[lb.prefix[i]];
        ++ng;
      }
      // '.0%': confidence * 100 rounded to the nearest integer, ties to even, on the double product (Python float
      // formatting is correctly rounded); confidences outside [0, 9999.99] have no counterpart in the detector's output
      const double pct = rint(d.confidence * 100.0);
      long long n = pct >= 0.0 && pct < 1.0e6 ? (long long)pct : 0;
      uint8_t digits[8];
      int nd = 0;
      do {
        digits[nd++] = (uint8_t)(n % 10);
        n /= 10;
      } while (n > 0 && nd < 8);
      for (int i = nd - 1; i >= 0 && ng < FX_MAX_GLYPHS; --i) {
        const uint8_t g = digit_glyphs[digits[i]];
        r.glyph[ng] = g;
        r.pen2[ng] = (uint16_t)pen;
        pen += font.advance[g];
        ++ng;
      }
      if (ng < FX_MAX_GLYPHS) {
        const uint8_t g = digit_glyphs[10];  // '%'
        r.glyph[ng] = g;
        r.pen2[ng] = (uint16_t)pen;
        pen += font.advance[g];
        ++ng;
      }
      r.n_glyphs = ng;
      // cv2.getTextSize: width = cvRound(sum(advance) * 0.5 + thickness), half to even      draw.py:55-59
      const int text_width = (int)rint((double)pen * 0.5 + 1.0);
      const int text_height = font.text_height, baseline = font.baseline, margin = font.margin;
      const int total = text_height + 2 * margin;  // draw.py:67
      int text_bottom;
      if (top - baseline > total)
        text_bottom = top;
      else if (bottom + total + baseline < H)
        text_bottom = bottom + total + baseline;
      else
        text_bottom = top + total + baseline;
      const int p1x = left, p1y = text_bottom - baseline - text_height - 2 * margin;  // draw.py:76-77
      const int p2x = left + text_width + 2 * margin, p2y = text_bottom;
      // image[p1y:p2y, p1x:p2x]: non-negative indices clamp to the image.  (Negative ones would wrap around in numpy;
      // the detector never produces them and such rows get no label here.)
      const bool valid = p1x >= 0 && p1y >= 0;
      r.bx0 = min(p1x, W);
      r.bx1 = min(p2x, W);
      r.by0 = min(p1y, H);
      r.by1 = min(p2y, H);
      if (!valid || r.by0 >= r.by1) {  // draw.py:80 `if len(cropped_image) == 0: return`
        r.bx0 = r.bx1 = r.by0 = r.by1 = 0;
        r.n_glyphs = 0;
      }
      r.ox = left + margin;  // draw.py:86-87
      r.oy = text_bottom - baseline - margin;
      r.text_x1 = r.ox + (pen >> 1) + font.cols;
      // draw.py:100-103: zones of the active detections select the contours to outline
      uint32_t z = 0;
      for (int i = 0; i < WB_MAX_ZONES; ++i)
        if (d.zones[i] > 0 && d.zones[i] <= 32) z |= 1u << (d.zones[i] - 1);
      if (z) atomicOr(&s_zone, z);
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, active);
  if ((t & 31) == 0) s_active[t >> 5] = m;
  __syncthreads();
  if (active) {
    int pos = __popc(m & ((1u << (t & 31)) - 1u));
    for (int w = 0; w < (t >> 5); ++w) pos += __popc(s_active[w]);
    fr.order[pos] = (uint8_t)t;
  }
  if (t == 0) {
    fr.n_active = __popc(s_active[0]) + __popc(s_active[1]) + __popc(s_active[2]) + __popc(s_active[3]);
    fr.zone_sel = s_zone;
  }
}

// ---------------------------------------------------------------------------------------------------
// grid (ceil(W / 128), ceil(H / 8), frames), 256 threads; thread = 4 consecutive pixels of one row.  (32-row tiles with
// four rows per thread and a quarter of the blocks measured 45 % slower: the pass is bound by the latency of dependent
// loads, and more resident threads hide it better than fewer culling preambles.)  For the same reason a thread issues
// the loads of its pixels, alpha values and outline bits first and only then takes part in building the tile's list of
// detections, so that the two chains of global round trips overlap.
constexpr int FX_TW = 128, FX_TH = 8;

// (forcing 32 registers for 8 blocks per SM instead of 5 spills and measured 8 % slower)
__global__ void __launch_bounds__(256, 5)
    k_fx_render(const FxFrameDesc* __restrict__ frames, const FxFrame* __restrict__ prep, FxFont font,
                const FxLabel* __restrict__ labels, const uint8_t* __restrict__ aw_lut, uint32_t flags) {
  __shared__ FxDet s_det[WB_MAX_DETECTIONS];
  __shared__ int s_n;
  __shared__ uint32_t s_hit[4];
  const FxFrameDesc fd = frames[blockIdx.z];
  const FxFrame& fr = prep[blockIdx.z];
  const int W = fd.cam.w, H = fd.cam.h;
  const int tx0 = blockIdx.x * FX_TW, ty0 = blockIdx.y * FX_TH;
  if (tx0 >= W || ty0 >= H) return;
  const int tx1 = min(tx0 + FX_TW, W), ty1 = min(ty0 + FX_TH, H);
  const int t = threadIdx.x;
  const int y = ty0 + t / (FX_TW / 4);
  const int xb = tx0 + (t % (FX_TW / 4)) * 4;
  const bool live = y < H && xb < W;  // threads outside the frame still take part in the barriers below
  const size_t row = (size_t)(live ? y : 0) * W;
  const int npx = live ? min(4, W - xb) : 0;
  const size_t px0 = row + (live ? xb : 0);
  const uint8_t* src = fd.in + px0 * 3;
  uint8_t* dst = fd.out + px0 * 3;
  const bool blend = (flags & WB_FX_BLEND) && fd.cam.alpha != nullptr;
  const bool outline = (flags & WB_FX_CONTOURS) && fd.cam.contours != nullptr;
  const bool vec = npx == 4 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0;
  // ---- loads first
  uint32_t ws[3] = {0u, 0u, 0u};
  uint8_t v[4][3];
  uint8_t al[4] = {255, 255, 255, 255};
  uint32_t cb[4] = {0u, 0u, 0u, 0u};
  if (vec) {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
    ws[0] = __ldg(s32);
    ws[1] = __ldg(s32 + 1);
    ws[2] = __ldg(s32 + 2);
  } else {
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if (p < npx) {
#pragma unroll
        for (int c = 0; c < 3; ++c) v[p][c] = __ldg(src + p * 3 + c);
      }
  }
  if (blend) {
    const uint8_t* ap = fd.cam.alpha + px0;
    if (npx == 4 && (reinterpret_cast<uintptr_t>(ap) & 3) == 0) {
      const uint32_t a4 = __ldg(reinterpret_cast<const uint32_t*>(ap));
#pragma unroll
      for (int p = 0; p < 4; ++p) al[p] = (uint8_t)(a4 >> (8 * p));
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (p < npx) al[p] = __ldg(ap + p);
    }
  }
  if (outline) {
    const uint32_t* cp = fd.cam.contours + px0;
    if (npx == 4 && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
      const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(cp));
      cb[0] = c4.x;
      cb[1] = c4.y;
      cb[2] = c4.z;
      cb[3] = c4.w;
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (p < npx) cb[p] = __ldg(cp + p);
    }
  }
  // ---- detections that can touch this tile, in row order
  int n_here = 0;
  uint32_t zone_sel = 0u;
  if (flags & WB_FX_DRAW) {
    const int na = fr.n_active;
    zone_sel = fr.zone_sel;
    if (na > 0) {  // uniform
      bool hit = false;
      int idx = 0;
      if (t < na) {
        idx = fr.order[t];
        const FxDet& d = fr.det[idx];
        // the tile holds outline pixels unless it misses the box or lies strictly inside it
        const bool rect = d.x0 < tx1 && d.x1 >= tx0 && d.y0 < ty1 && d.y1 >= ty0 &&
                          !(tx0 > d.x0 && tx1 - 1 < d.x1 && ty0 > d.y0 && ty1 - 1 < d.y1);
        const bool box = d.bx0 < tx1 && max(d.bx1, d.text_x1) > tx0 && d.by0 < ty1 && d.by1 > ty0 && d.by1 > d.by0;
        hit = rect || box;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (t < 128 && (t & 31) == 0) s_hit[t >> 5] = m;
      __syncthreads();
      if (hit) {
        int pos = __popc(m & ((1u << (t & 31)) - 1u));
        for (int w = 0; w < (t >> 5); ++w) pos += __popc(s_hit[w]);
        s_det[pos] = fr.det[idx];
      }
      if (t == 0) s_n = __popc(s_hit[0]) + __popc(s_hit[1]) + __popc(s_hit[2]) + __popc(s_hit[3]);
      __syncthreads();
      n_here = s_n;
    }
  }
  if (!live) return;
  if (vec) {
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i / 3][i % 3] = (uint8_t)(ws[i >> 2] >> (8 * (i & 3)));
  }
  // CopyImageEffect / BlendEffect
  if (blend) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (p >= npx) continue;
      const float af = __fdiv_rn((float)al[p], 255.f);         // blend.py:15
      const float wi = __fmul_rn(255.f, __fsub_rn(1.f, af));   // blend.py:21-22
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[p][c] = (uint8_t)__float2int_rz(__fadd_rn(__fmul_rn((float)v[p][c], af), wi));  // blend.py:28-32
    }
  }
  // DrawEffect: detections in row order; within one detection rectangle, label box, text (draw.py:51-88)
  for (int i = 0; i < n_here; ++i) {
    const FxDet& d = s_det[i];
    const uint8_t* col = labels[d.style].box_color;
    const bool on_h = (y == d.y0 || y == d.y1);
    const bool in_y = y >= d.y0 && y <= d.y1;
    const bool box_y = y >= d.by0 && y < d.by1;
    const int gy = y - (d.oy + font.y0);
    const bool text_y = d.n_glyphs > 0 && gy >= 0 && gy < font.rows;
    if (!in_y && !box_y && !text_y) continue;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (p >= npx) continue;
      const int x = xb + p;
      if ((on_h && x >= d.x0 && x <= d.x1) || (in_y && (x == d.x0 || x == d.x1))) {  // cv2.rectangle, thickness 1
        v[p][0] = col[0];
        v[p][1] = col[1];
        v[p][2] = col[2];
      }
      if (box_y && x >= d.bx0 && x < d.bx1) {  // cv2.addWeighted(cropped, alpha, solid, 1 - alpha, 0)   draw.py:81-85
        const uint8_t* lut = aw_lut + (size_t)d.style * 768;
        v[p][0] = lut[v[p][0]];
        v[p][1] = lut[256 + v[p][1]];
        v[p][2] = lut[512 + v[p][2]];
      }
      if (text_y && x >= d.ox && x < d.text_x1) {  // cv2.putText, glyph after glyph                       draw.py:86-88
        for (int g = 0; g < d.n_glyphs; ++g) {
          const int pen2 = d.pen2[g];
          const int px = d.ox + (pen2 >> 1);
          const int gc = x - px;
          if (gc < 0) break;  // pens only move right
          if (gc >= font.cols) continue;
          const int k = W - px;  // distance from the pen to the right border: OpenCV clips the strokes there
          const int clip = min(k, font.cols + 1) - 1;
          const uint8_t* tab =
              font.lut + ((((size_t)(d.glyph[g] * 2 + (pen2 & 1)) * (font.cols + 1) + clip) * font.rows + gy) * font.cols + gc) * 256;
          v[p][0] = __ldg(tab + v[p][0]);
          v[p][1] = __ldg(tab + v[p][1]);
          v[p][2] = __ldg(tab + v[p][2]);
        }
      }
    }
  }
  // DrawEffectWithContours: outline every zone some drawn detection lies in (draw.py:100-103), colour (255, 255, 0)
  if (outline && zone_sel != 0u) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if (p < npx && (cb[p] & zone_sel)) {
        v[p][0] = 255;
        v[p][1] = 255;
        v[p][2] = 0;
      }
  }
  if (vec) {
    uint32_t wo[3] = {0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < 12; ++i) wo[i >> 2] |= (uint32_t)v[i / 3][i % 3] << (8 * (i & 3));
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    d32[0] = wo[0];
    d32[1] = wo[1];
    d32[2] = wo[2];
  } else {
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if (p < npx) {
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[p * 3 + c] = v[p][c];
      }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
struct wb_fx {
  int device = 0;
  std::mutex mu;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  FxFont font{};
  int32_t* d_advance = nullptr;
  uint8_t* d_lut = nullptr;
  FxLabel* d_labels = nullptr;
  int n_labels = 0;
  uint8_t* d_digits = nullptr;
  uint8_t* d_aw = nullptr;
  std::map<int, FxCamera> cams;
  std::vector<void*> cam_allocs;
  // staging
  int cap_n = 0;
  size_t cap_bytes = 0;
  uint8_t *d_in = nullptr, *d_out = nullptr;
  wb_detection* d_rows = nullptr;
  wb_detection* h_rows = nullptr;
  FxFrameDesc* d_desc = nullptr;
  FxFrameDesc* h_desc = nullptr;
  FxFrame* d_prep = nullptr;
};

#define FXCK(call)                                                                                       \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fx_fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +        \
                     std::to_string(__LINE__) + ")");                                                    \
  } while (0)
#define FXREQ(cond, msg) \
  do {                   \
    if (!(cond)) return fx_fail(msg); \
  } while (0)

const char* wb_fx_last_error(void) { return g_fx_err.c_str(); }

int wb_fx_create(int device, const wb_fx_font* font, int n_labels, const wb_fx_label* labels, const uint8_t* digit_glyphs,
                 double alpha, wb_fx** out) {
  FXREQ(font && labels && digit_glyphs && out, "NULL argument");
  FXREQ(font->n_glyphs > 0 && font->n_glyphs <= 128 && font->rows > 0 && font->cols > 0 && font->cols <= 64,
        "bad font geometry");
  FXREQ(n_labels > 0 && n_labels <= 256, "n_labels must be in 1..256");
  for (int i = 0; i < n_labels; ++i) {
    FXREQ(labels[i].n_prefix <= sizeof(labels[i].prefix), "label prefix too long");
    for (int j = 0; j < labels[i].n_prefix; ++j) FXREQ(labels[i].prefix[j] < font->n_glyphs, "glyph index out of range");
  }
  for (int i = 0; i < 11; ++i) FXREQ(digit_glyphs[i] < font->n_glyphs, "digit glyph index out of range");
  int count = 0;
  FXCK(cudaGetDeviceCount(&count));
  FXREQ(device >= 0 && device < count, "no such CUDA device");
  FXCK(cudaSetDevice(device));
  cudaDeviceProp prop;
  FXCK(cudaGetDeviceProperties(&prop, device));
  FXREQ(prop.major == 10, "libwatsor_b200 is built for sm_100a (B200) only");
  std::unique_ptr<wb_fx> fx(new wb_fx());
  fx->device = device;
  FXCK(cudaStreamCreateWithFlags(&fx->stream, cudaStreamNonBlocking));
  FXCK(cudaEventCreate(&fx->ev0));
  FXCK(cudaEventCreate(&fx->ev1));
  const size_t lut_bytes = (size_t)font->n_glyphs * 2 * (font->cols + 1) * font->rows * font->cols * 256;
  FXCK(cudaMalloc(&fx->d_advance, sizeof(int32_t) * font->n_glyphs));
  FXCK(cudaMalloc(&fx->d_lut, lut_bytes));
  FXCK(cudaMemcpy(fx->d_advance, font->advance, sizeof(int32_t) * font->n_glyphs, cudaMemcpyHostToDevice));
  FXCK(cudaMemcpy(fx->d_lut, font->lut, lut_bytes, cudaMemcpyHostToDevice));
  fx->font.n_glyphs = font->n_glyphs;
  fx->font.rows = font->rows;
  fx->font.cols = font->cols;
  fx->font.y0 = font->y0;
  fx->font.text_height = font->text_height;
  fx->font.baseline = font->baseline;
  fx->font.margin = font->margin;
  fx->font.advance = fx->d_advance;
  fx->font.lut = fx->d_lut;
  static_assert(sizeof(FxLabel) == sizeof(wb_fx_label), "label style layout");
  fx->n_labels = n_labels;
  FXCK(cudaMalloc(&fx->d_labels, sizeof(FxLabel) * n_labels));
  FXCK(cudaMemcpy(fx->d_labels, labels, sizeof(FxLabel) * n_labels, cudaMemcpyHostToDevice));
  FXCK(cudaMalloc(&fx->d_digits, 16));
  FXCK(cudaMemcpy(fx->d_digits, digit_glyphs, 11, cudaMemcpyHostToDevice));
  // cv2.addWeighted(src1, alpha, src2, beta = 1 - alpha, 0) on 8-bit images:  saturate(rint(fmaf(a, alpha, b * beta)))
  // in float -- pinned against OpenCV on all 256 x 256 inputs (tests/test_effects_host.py)
  std::vector<uint8_t> aw((size_t)n_labels * 768);
  const float fa = (float)alpha, fb = (float)(1.0 - alpha);
  for (int l = 0; l < n_labels; ++l)
    for (int c = 0; c < 3; ++c)
      for (int a = 0; a < 256; ++a) {
        const float t = (float)labels[l].box_color[c] * fb;
        const float r = nearbyintf(fmaf((float)a, fa, t));
        aw[(size_t)l * 768 + c * 256 + a] = (uint8_t)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
      }
  FXCK(cudaMalloc(&fx->d_aw, aw.size()));
  FXCK(cudaMemcpy(fx->d_aw, aw.data(), aw.size(), cudaMemcpyHostToDevice));
  *out = fx.release();
  return 0;
}

int wb_fx_set_camera(wb_fx* fx, int cam_id, int width, int height, const uint8_t* alpha, const uint32_t* contour_bits) {
  FXREQ(fx, "NULL fx");
  FXREQ(width > 0 && height > 0, "bad frame size");
  std::lock_guard<std::mutex> lock(fx->mu);
  FXCK(cudaSetDevice(fx->device));
  FXCK(cudaStreamSynchronize(fx->stream));
  FxCamera cam{width, height, nullptr, nullptr};
  const size_t px = (size_t)width * height;
  if (alpha) {
    uint8_t* d = nullptr;
    FXCK(cudaMalloc(&d, px));
    FXCK(cudaMemcpy(d, alpha, px, cudaMemcpyHostToDevice));
    fx->cam_allocs.push_back(d);
    cam.alpha = d;
  }
  if (contour_bits) {
    uint32_t* d = nullptr;
    FXCK(cudaMalloc(&d, px * 4));
    FXCK(cudaMemcpy(d, contour_bits, px * 4, cudaMemcpyHostToDevice));
    fx->cam_allocs.push_back(d);
    cam.contours = d;
  }
  fx->cams[cam_id] = cam;
  return 0;
}

int wb_fx_render(wb_fx* fx, int n, const uint8_t* const* images_in, uint8_t* const* images_out, const int32_t* cam_ids,
                 const wb_detection* const* rows, uint32_t flags, float* gpu_ms) {
  FXREQ(fx, "NULL fx");
  FXREQ(n > 0 && n <= 4096, "n out of range");
  FXREQ(images_in && images_out && cam_ids && rows, "NULL argument");
  std::lock_guard<std::mutex> lock(fx->mu);
  FXCK(cudaSetDevice(fx->device));
  const bool on_device = (flags & WB_FX_ON_DEVICE) != 0;
  size_t total = 0;
  int max_w = 0, max_h = 0;
  for (int i = 0; i < n; ++i) {
    auto it = fx->cams.find(cam_ids[i]);
    FXREQ(it != fx->cams.end(), "cam_id " + std::to_string(cam_ids[i]) + " has not been configured with wb_fx_set_camera");
    FXREQ(images_in[i] && images_out[i] && rows[i], "NULL frame / rows pointer");
    // labels are placed inside the frame only if it is high enough for one above/below/inside a box (draw.py:68-73);
    // lower frames would need OpenCV's re-capping of strokes cut by the bottom border, which the tables do not hold
    const int min_h = 2 * (fx->font.text_height + 2 * fx->font.margin + fx->font.baseline) + 1;
    FXREQ(!(flags & WB_FX_DRAW) || it->second.h >= min_h,
          "the draw effect needs frames of at least " + std::to_string(min_h) + " rows");
    total += ((size_t)it->second.w * it->second.h * 3 + 255) / 256 * 256;
    max_w = std::max(max_w, it->second.w);
    max_h = std::max(max_h, it->second.h);
  }
  if (n > fx->cap_n) {
    FXCK(cudaStreamSynchronize(fx->stream));
    cudaFree(fx->d_rows);
    cudaFree(fx->d_desc);
    cudaFree(fx->d_prep);
    cudaFreeHost(fx->h_rows);
    cudaFreeHost(fx->h_desc);
    fx->cap_n = 0;
    FXCK(cudaMalloc(&fx->d_rows, sizeof(wb_detection) * WB_MAX_DETECTIONS * n));
    FXCK(cudaMalloc(&fx->d_desc, sizeof(FxFrameDesc) * n));
    FXCK(cudaMalloc(&fx->d_prep, sizeof(FxFrame) * n));
    FXCK(cudaMallocHost(&fx->h_rows, sizeof(wb_detection) * WB_MAX_DETECTIONS * n));
    FXCK(cudaMallocHost(&fx->h_desc, sizeof(FxFrameDesc) * n));
    fx->cap_n = n;
  }
  if (!on_device && total > fx->cap_bytes) {
    FXCK(cudaStreamSynchronize(fx->stream));
    cudaFree(fx->d_in);
    cudaFree(fx->d_out);
    fx->cap_bytes = 0;
    FXCK(cudaMalloc(&fx->d_in, total));
    FXCK(cudaMalloc(&fx->d_out, total));
    fx->cap_bytes = total;
  }
  cudaStream_t st = fx->stream;
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const FxCamera& cam = fx->cams[cam_ids[i]];
    const size_t bytes = (size_t)cam.w * cam.h * 3;
    memcpy(fx->h_rows + (size_t)i * WB_MAX_DETECTIONS, rows[i], sizeof(wb_detection) * WB_MAX_DETECTIONS);
    FxFrameDesc d;
    d.cam = cam;
    if (on_device) {
      d.in = images_in[i];
      d.out = images_out[i];
    } else {
      FXCK(cudaMemcpyAsync(fx->d_in + off, images_in[i], bytes, cudaMemcpyHostToDevice, st));
      d.in = fx->d_in + off;
      d.out = fx->d_out + off;
    }
    fx->h_desc[i] = d;
    off += (bytes + 255) / 256 * 256;
  }
  FXCK(cudaMemcpyAsync(fx->d_rows, fx->h_rows, sizeof(wb_detection) * WB_MAX_DETECTIONS * n, cudaMemcpyHostToDevice, st));
  FXCK(cudaMemcpyAsync(fx->d_desc, fx->h_desc, sizeof(FxFrameDesc) * n, cudaMemcpyHostToDevice, st));
  FXCK(cudaEventRecord(fx->ev0, st));
  if (flags & WB_FX_DRAW)
    k_fx_prepare<<<n, 128, 0, st>>>(fx->d_rows, fx->d_desc, fx->font, fx->d_labels, fx->n_labels, fx->d_digits, fx->d_prep);
  dim3 grid((max_w + FX_TW - 1) / FX_TW, (max_h + FX_TH - 1) / FX_TH, n);
  k_fx_render<<<grid, 256, 0, st>>>(fx->d_desc, fx->d_prep, fx->font, fx->d_labels, fx->d_aw, flags);
  FXCK(cudaGetLastError());
  FXCK(cudaEventRecord(fx->ev1, st));
  if (!on_device) {
    off = 0;
    for (int i = 0; i < n; ++i) {
      const FxCamera& cam = fx->cams[cam_ids[i]];
      const size_t bytes = (size_t)cam.w * cam.h * 3;
      FXCK(cudaMemcpyAsync(images_out[i], fx->d_out + off, bytes, cudaMemcpyDeviceToHost, st));
      off += (bytes + 255) / 256 * 256;
    }
  }
  FXCK(cudaStreamSynchronize(st));
  if (gpu_ms) FXCK(cudaEventElapsedTime(gpu_ms, fx->ev0, fx->ev1));
  return 0;
}

int wb_fx_destroy(wb_fx* fx) {
  if (!fx) return 0;
  cudaSetDevice(fx->device);
  if (fx->stream) cudaStreamSynchronize(fx->stream);
  for (void* p : fx->cam_allocs) cudaFree(p);
  cudaFree(fx->d_advance);
  cudaFree(fx->d_lut);
  cudaFree(fx->d_labels);
  cudaFree(fx->d_digits);
  cudaFree(fx->d_aw);
  cudaFree(fx->d_in);
  cudaFree(fx->d_out);
  cudaFree(fx->d_rows);
  cudaFree(fx->d_desc);
  cudaFree(fx->d_prep);
  cudaFreeHost(fx->h_rows);
  cudaFreeHost(fx->h_desc);
  if (fx->ev0) cudaEventDestroy(fx->ev0);
  if (fx->ev1) cudaEventDestroy(fx->ev1);
  if (fx->stream) cudaStreamDestroy(fx->stream);
  delete fx;
  return 0;
}
