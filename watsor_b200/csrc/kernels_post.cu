// kernels_post.cu -- K7 (anchor decode + sigmoid), K8 (per-class NMS, global top-100),
// K9 (integer conversion + confidence / area / mask-zone predicates -> Detection[100]).
//
// Restates `Postprocessor/Decode/*`, `Postprocessor/convert_scores`, `Postprocessor/Slice`,
// `Postprocessor/BatchMultiClassNonMaxSuppression/*` and the final `add` of the frozen graph
// (watsor/detection/tensorflow_cpu.py:114), the python write loop tensorflow_cpu.py:79-90, and
// watsor/filter/{confidence,area,mask}.py as applied by watsor/filter/track.py:26.
//
// Every float op is an explicit round-to-nearest intrinsic (no FMA contraction) in the graph's op
// order, so that given identical head outputs the result is bit-identical to the oracle
// (up to the last-ulp difference between CUDA's expf and the host libm's).
#include <algorithm>

#include "common.cuh"

// ------------------------------------------------------------------------------------------- K7
__device__ __forceinline__ float4 decode_box(float4 e, float4 a, const PostParams& pp) {
  // anchors are corner boxes [ymin,xmin,ymax,xmax]  (get_center_coordinates_and_sizes)
  float wa = __fsub_rn(a.w, a.y), ha = __fsub_rn(a.z, a.x);
  float ycenter_a = __fadd_rn(a.x, __fdiv_rn(ha, 2.0f));
  float xcenter_a = __fadd_rn(a.y, __fdiv_rn(wa, 2.0f));
  float ty = __fdiv_rn(e.x, pp.scale_y), tx = __fdiv_rn(e.y, pp.scale_x);
  float th = __fdiv_rn(e.z, pp.scale_h), tw = __fdiv_rn(e.w, pp.scale_w);
  float w = __fmul_rn(expf(tw), wa), h = __fmul_rn(expf(th), ha);
  float ycenter = __fadd_rn(__fmul_rn(ty, ha), ycenter_a);
  float xcenter = __fadd_rn(__fmul_rn(tx, wa), xcenter_a);
  float hh = __fdiv_rn(h, 2.0f), hw = __fdiv_rn(w, 2.0f);
  return make_float4(__fsub_rn(ycenter, hh), __fsub_rn(xcenter, hw), __fadd_rn(ycenter, hh),
                     __fadd_rn(xcenter, hw));
}

__device__ __forceinline__ float sigmoid_score(float logit, float logit_scale) {
  float z = __fdiv_rn(logit, logit_scale);
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-z)));
}

constexpr int DEC_TILE = 64;

// grid (ceil(N/DEC_TILE), frames).  Decodes DEC_TILE boxes, then sweeps the [DEC_TILE][C+1] logit tile
// with coalesced loads; every (anchor, class) with score > threshold becomes a candidate key
// score_bits << 32 | ~anchor  (sorting keys descending gives "score descending, lower anchor index first",
// the pop order of TF's NonMaxSuppressionV5).  Appends are aggregated per block: positions inside the
// block come from shared-memory counters, one global atomicAdd per (block, class) reserves the range
// (90 classes x 1917 anchors would otherwise be 172 k global atomics per frame on 90 addresses).
__global__ void __launch_bounds__(256)
    k_decode_scores(PostParams pp, const float* __restrict__ enc, const float* __restrict__ logits,
                    const float* __restrict__ anchors, float4* __restrict__ dec, int* __restrict__ cand_count,
                    unsigned long long* __restrict__ cand) {
  extern __shared__ unsigned char s_dec_raw[];
  const int f = blockIdx.y, a0 = blockIdx.x * DEC_TILE;
  const int N = pp.num_anchors, C = pp.num_classes, C1 = C + 1;
  const int na = min(DEC_TILE, N - a0);
  const int total = na * C1;
  float* s_score = reinterpret_cast<float*>(s_dec_raw);                         // [DEC_TILE*C1]
  int* s_cnt = reinterpret_cast<int*>(s_score + DEC_TILE * C1);                 // [C] then base [C]
  int* s_base = s_cnt + C;
  short* s_pos = reinterpret_cast<short*>(s_base + C);                          // [DEC_TILE*C1]
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_cnt[c] = 0;
  if ((int)threadIdx.x < na) {
    int i = a0 + threadIdx.x;
    float4 e = reinterpret_cast<const float4*>(enc)[(size_t)f * N + i];
    float4 a = reinterpret_cast<const float4*>(anchors)[i];
    dec[(size_t)f * N + i] = decode_box(e, a, pp);
  }
  __syncthreads();
  const float* lt = logits + ((size_t)f * N + a0) * C1;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int il = e / C1, c = e - il * C1;
    short pos = -1;
    if (c != 0) {  // `Postprocessor/Slice`: background column dropped after the sigmoid
      float sc = sigmoid_score(lt[e], pp.logit_scale);
      s_score[e] = sc;
      if (sc > pp.score_thr) pos = (short)atomicAdd(&s_cnt[c - 1], 1);
    }
    s_pos[e] = pos;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    s_base[c] = s_cnt[c] > 0 ? atomicAdd(&cand_count[f * C + c], s_cnt[c]) : 0;
  __syncthreads();
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const short pos = s_pos[e];
    if (pos < 0) continue;
    int il = e / C1, c = e - il * C1;
    cand[((size_t)f * C + (c - 1)) * N + s_base[c - 1] + pos] =
        ((unsigned long long)__float_as_uint(s_score[e]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(a0 + il));
  }
}

// ------------------------------------------------------------------------------------------- K8
// TF non_max_suppression_op.cc IOU<float>() -- same op order, separately rounded
__device__ __forceinline__ float iou_tf(float4 bi, float4 bj) {
  float ymin_i = fminf(bi.x, bi.z), xmin_i = fminf(bi.y, bi.w);
  float ymax_i = fmaxf(bi.x, bi.z), xmax_i = fmaxf(bi.y, bi.w);
  float ymin_j = fminf(bj.x, bj.z), xmin_j = fminf(bj.y, bj.w);
  float ymax_j = fmaxf(bj.x, bj.z), xmax_j = fmaxf(bj.y, bj.w);
  float area_i = __fmul_rn(__fsub_rn(ymax_i, ymin_i), __fsub_rn(xmax_i, xmin_i));
  float area_j = __fmul_rn(__fsub_rn(ymax_j, ymin_j), __fsub_rn(xmax_j, xmin_j));
  if (area_i <= 0.f || area_j <= 0.f) return 0.f;
  float iy0 = fmaxf(ymin_i, ymin_j), ix0 = fmaxf(xmin_i, xmin_j);
  float iy1 = fminf(ymax_i, ymax_j), ix1 = fminf(xmax_i, xmax_j);
  float inter = __fmul_rn(fmaxf(__fsub_rn(iy1, iy0), 0.f), fmaxf(__fsub_rn(ix1, ix0), 0.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_i, area_j), inter));
}

// Same predicate as `iou_tf(a, b) > thr` for boxes whose corners were normalised (min/max) and whose areas
// were computed once: disjoint boxes are rejected after 4 min/max + 2 subtractions, without the IEEE
// division.  inter = max(dy,0)*max(dx,0) is 0 when dy <= 0 or dx <= 0, so IoU is 0 <= thr there.
struct NBox {
  float4 c;  // ymin, xmin, ymax, xmax (normalised)
  float area;
};
__device__ __forceinline__ NBox normalise_box(float4 b) {
  NBox r;
  r.c = make_float4(fminf(b.x, b.z), fminf(b.y, b.w), fmaxf(b.x, b.z), fmaxf(b.y, b.w));
  r.area = __fmul_rn(__fsub_rn(r.c.z, r.c.x), __fsub_rn(r.c.w, r.c.y));
  return r;
}
__device__ __forceinline__ bool suppresses(const float4& a, float area_a, const float4& b, float area_b, float thr) {
  const float iy0 = fmaxf(a.x, b.x), ix0 = fmaxf(a.y, b.y), iy1 = fminf(a.z, b.z), ix1 = fminf(a.w, b.w);
  const float dy = __fsub_rn(iy1, iy0), dx = __fsub_rn(ix1, ix0);
  if (!(dy > 0.f && dx > 0.f)) return false;
  if (area_a <= 0.f || area_b <= 0.f) return false;
  const float inter = __fmul_rn(dy, dx);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter)) > thr;
}

__device__ __forceinline__ float4 clip_unit(float4 b) {  // ClipToWindow [0,0,1,1]
  return make_float4(fmaxf(fminf(b.x, 1.f), 0.f), fmaxf(fminf(b.y, 1.f), 0.f), fmaxf(fminf(b.z, 1.f), 0.f),
                     fmaxf(fminf(b.w, 1.f), 0.f));
}

// descending bitonic sort of P (power of two) keys in shared memory, whole block
__device__ __forceinline__ void bitonic_desc(unsigned long long* a, int P) {
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = a[i], y = a[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? x < y : x > y) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
}

// grid (classes, frames), 256 threads.  Per (frame, class):
//  1. candidate keys -> shared memory, in CHUNKS of descending score.  With many candidates (score threshold 1e-8
//     makes every anchor one) only the head of the order is ever visited before max_per_class boxes are kept or the
//     early exit fires (most classes stop after two rounds), so the keys are never sorted as a whole: a score
//     histogram (sign, exponent and 5 mantissa bits of the float: 32 bins per octave) is built once, and each chunk
//     is "the highest remaining bins that hold >= want keys" (96 for the first chunk, 384 after), gathered from the
//     L2-resident key list and sorted (bitonic).  A chunk is a whole number of bins and every key of a higher bin is
//     larger than every key of a lower one, so the visiting order is exactly the descending key order.
//  2. greedy suppression with exact sequential semantics, 32 candidates per round:
//     phase 1 (all 8 warps): candidate i vs every box kept in EARLIER rounds (warp w takes candidates
//     4w..4w+3, lanes stride over the kept list, a ballot decides);
//     phase 2 (warp 0): walk the 32 candidates in order; a surviving candidate is kept and immediately
//     suppresses the later candidates of the same round that it overlaps.
// Output: merge keys  score_bits << 32 | (0xFFFF - class) << 16 | (0xFFFF - rank)  for the kept boxes whose
// window-clipped area is positive (0 otherwise), plus the anchor index of every kept box.
constexpr int KEPT_BINS = 1024;
// monotone non-decreasing in the score (scores are sigmoid outputs in [0, 1]); resolution 1/1024
__device__ __forceinline__ int score_bin(unsigned score_bits) {
  return min(KEPT_BINS - 1, max(0, (int)(__uint_as_float(score_bits) * (float)KEPT_BINS)));
}
constexpr int NMS_CHUNK0 = 96;   // keys wanted in the first chunk (three rounds of 32)
constexpr int NMS_CHUNK = 384;   // ... in every later chunk
constexpr int NMS_PREFILTER_MIN = 640;
constexpr int NMS_BINS = 1024;
// bin of a candidate key, monotone non-decreasing in the key: sign + exponent + 5 mantissa bits of the score, counted
// from 2^-27 (scores are in (0, 1]: exponents 100..127; anything smaller lands in bin 0)
__device__ __forceinline__ int nms_bin(unsigned long long key) {
  return min(NMS_BINS - 1, max(0, (int)(key >> 50) - (100 << 5)));
}

__global__ void __launch_bounds__(256)
    k_nms(PostParams pp, const float4* __restrict__ dec, const int* __restrict__ cand_count,
          const unsigned long long* __restrict__ cand, int sort_cap, int* __restrict__ sel_count,
          unsigned long long* __restrict__ sel_key, int* __restrict__ sel_idx, int* __restrict__ kept_hist) {
  extern __shared__ unsigned long long s_a[];  // [sort_cap]: the current chunk, sorted
  __shared__ float4 s_kept[128];  // normalised corners of the kept boxes
  __shared__ float s_karea[128];
  __shared__ float4 s_cbox[32];  // normalised corners of the round's candidates
  __shared__ float s_carea[32];
  __shared__ float4 s_craw[32];  // as decoded (for the clipped-area test)
  __shared__ unsigned s_alive;   // bit i: candidate i of the round survived phase 1
  __shared__ int s_nkept_sh, s_na, s_cut;
  __shared__ int s_hist[NMS_BINS];
  __shared__ int s_above[8];
  int* fhist = kept_hist + (size_t)blockIdx.y * KEPT_BINS;  // this frame's histogram of kept, selectable scores
  const int c = blockIdx.x, f = blockIdx.y, C = pp.num_classes, N = pp.num_anchors;
  const int n = min(cand_count[f * C + c], N);
  const int max_out = min(pp.max_per_class, N);
  unsigned long long* out_key = sel_key + ((size_t)f * C + c) * pp.max_per_class;
  int* out_idx = sel_idx + ((size_t)f * C + c) * pp.max_per_class;
  for (int i = threadIdx.x; i < pp.max_per_class; i += blockDim.x) out_key[i] = 0ull;
  if (n == 0) {
    if (threadIdx.x == 0) sel_count[f * C + c] = 0;
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long* ck = cand + ((size_t)f * C + c) * N;
  const bool chunked = n > NMS_PREFILTER_MIN;
  if (chunked) {
    for (int i = threadIdx.x; i < NMS_BINS; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
      // scores of one class crowd into few bins: aggregate equal bins inside the warp (__match_any_sync) so that one
      // lane per distinct bin issues the shared-memory atomic
      const int i = i0 + threadIdx.x;
      const int bin = i < n ? nms_bin(ck[i]) : -1;
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      if (bin >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s_hist[bin], __popc(peers));
    }
  }
  const float4* fdec = dec + (size_t)f * N;
  if (threadIdx.x == 0) s_nkept_sh = 0;
  __syncthreads();

  bool stop = false;
  int hi_bin = NMS_BINS;  // bins >= hi_bin have been visited
  for (int chunk = 0; !stop; ++chunk) {
    int count;
    if (!chunked) {
      if (chunk > 0) break;
      int P = 32;
      while (P < n) P <<= 1;
      for (int i = threadIdx.x; i < P; i += blockDim.x) s_a[i] = i < n ? ck[i] : 0ull;
      __syncthreads();
      bitonic_desc(s_a, P);
      count = n;
    } else {
      if (hi_bin <= 0 || s_nkept_sh >= max_out) break;  // uniform: both were published before a barrier
      const int want = chunk == 0 ? NMS_CHUNK0 : NMS_CHUNK;
      if (warp == 0) {  // highest remaining bins first until `want` keys are covered (or nothing is left: cut = 0)
        int acc = 0, cut = 0;
        for (int top = hi_bin - 1; top >= 0 && acc < want; top -= 32) {
          const int b = top - lane;
          const int v = b >= 0 ? s_hist[b] : 0;
          int incl = v;  // inclusive prefix over lanes (lane 0 = highest bin)
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          const unsigned reach = __ballot_sync(0xffffffffu, acc + incl >= want);
          if (reach) {
            const int l = __ffs(reach) - 1;
            cut = top - l;
            acc += __shfl_sync(0xffffffffu, incl, l);
            break;
          }
          acc += __shfl_sync(0xffffffffu, incl, 31);
          cut = max(top - 31, 0);
        }
        if (lane == 0) {
          s_cut = cut;
          s_na = 0;
        }
      }
      __syncthreads();
      const int cut = s_cut;
      for (int i0 = 0; i0 < n; i0 += blockDim.x) {  // pass over the L2-resident keys: bins [cut, hi_bin)
        const int i = i0 + threadIdx.x;
        const unsigned long long k = i < n ? ck[i] : 0ull;
        const int bin = nms_bin(k);
        const bool in = i < n && bin >= cut && bin < hi_bin;
        // warp-aggregated append: one shared-memory atomic per warp instead of one per key
        const unsigned m = __ballot_sync(0xffffffffu, in);
        int base_pos = 0;
        if (lane == 0 && m) base_pos = atomicAdd(&s_na, __popc(m));
        base_pos = __shfl_sync(0xffffffffu, base_pos, 0);
        if (in) s_a[base_pos + __popc(m & ((1u << lane) - 1u))] = k;
      }
      __syncthreads();
      count = s_na;
      hi_bin = cut;
      if (count == 0) continue;  // uniform; an empty range can only be the last one (cut == 0)
      int P = 32;
      while (P < count) P <<= 1;
      for (int i = count + threadIdx.x; i < P; i += blockDim.x) s_a[i] = 0ull;
      __syncthreads();
      bitonic_desc(s_a, P);
    }
    const unsigned long long* sorted = s_a;
    for (int base = 0; base < count; base += 32) {
      const int nkept = s_nkept_sh;
      if (nkept >= max_out) break;
      if (base > 0 || chunk > 0) {
        // Early exit (exact): the frame-wide histogram counts boxes that some class has already KEPT with a positive
        // clipped area -- final facts.  If max_total of them sit in score bins strictly above this class's best
        // remaining candidate, neither it nor anything after it can reach the frame's top max_total, and whatever
        // this class would still select is never read.  Stale (smaller) counts only delay the exit.
        const int b = score_bin((unsigned)(sorted[base] >> 32));
        int above = 0;
        for (int i = threadIdx.x * 4; i < KEPT_BINS; i += blockDim.x * 4) {
          const int4 h = __ldcg(reinterpret_cast<const int4*>(fhist + i));
          above += (i > b ? h.x : 0) + (i + 1 > b ? h.y : 0) + (i + 2 > b ? h.z : 0) + (i + 3 > b ? h.w : 0);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
        if (lane == 0) s_above[warp] = above;
        __syncthreads();
        above = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) above += s_above[w];
        __syncthreads();
        if (above >= pp.max_total) {
          stop = true;
          break;
        }
      }
      const int cnt = min(32, count - base);
      if (threadIdx.x < 32) {
        if (lane < cnt) {
          const unsigned idx = 0xFFFFFFFFu - (unsigned)(sorted[base + lane] & 0xFFFFFFFFull);
          const float4 raw = fdec[idx];
          const NBox nb = normalise_box(raw);
          s_craw[lane] = raw;
          s_cbox[lane] = nb.c;
          s_carea[lane] = nb.area;
        }
        if (lane == 0) s_alive = 0u;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = warp * 4 + k;
        if (i < cnt) {
          const float4 bi = s_cbox[i];
          const float ai = s_carea[i];
          bool sup = false;
          for (int j = lane; j < nkept; j += 32) sup |= suppresses(bi, ai, s_kept[j], s_karea[j], pp.iou_thr);
          if (!__any_sync(0xffffffffu, sup) && lane == 0) atomicOr(&s_alive, 1u << i);
        }
      }
      __syncthreads();
      if (warp == 0) {
        unsigned alive = s_alive;
        const float4 box = lane < cnt ? s_cbox[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float barea = lane < cnt ? s_carea[lane] : 0.f;
        int nk = nkept;
        for (int t = 0; t < cnt && nk < max_out; ++t) {
          if (!((alive >> t) & 1u)) continue;  // warp-uniform
          float4 bt;
          bt.x = __shfl_sync(0xffffffffu, box.x, t);
          bt.y = __shfl_sync(0xffffffffu, box.y, t);
          bt.z = __shfl_sync(0xffffffffu, box.z, t);
          bt.w = __shfl_sync(0xffffffffu, box.w, t);
          const float at = __shfl_sync(0xffffffffu, barea, t);
          if (lane == t) {
            const unsigned long long key = sorted[base + t];
            s_kept[nk] = box;
            s_karea[nk] = barea;
            float4 cb = clip_unit(s_craw[t]);
            float area = __fmul_rn(__fsub_rn(cb.z, cb.x), __fsub_rn(cb.w, cb.y));
            out_idx[nk] = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
            out_key[nk] = area > 0.f ? ((key & 0xFFFFFFFF00000000ull) | ((unsigned long long)(0xFFFFu - (unsigned)c) << 16) |
                                        (unsigned long long)(0xFFFFu - (unsigned)nk))
                                     : 0ull;
            if (area > 0.f) atomicAdd(&fhist[score_bin((unsigned)(key >> 32))], 1);
          }
          ++nk;
          const bool hit = lane > t && lane < cnt && suppresses(box, barea, bt, at, pp.iou_thr);
          alive &= ~__ballot_sync(0xffffffffu, hit);
        }
        if (lane == 0) s_nkept_sh = nk;
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) sel_count[f * C + c] = s_nkept_sh;
}

// ------------------------------------------------------------------------------------------- K9
__device__ __forceinline__ int sat_count(const int32_t* s, int W1, int x0, int y0, int x1, int y1) {
  return s[(size_t)(y1 + 1) * W1 + (x1 + 1)] - s[(size_t)y0 * W1 + (x1 + 1)] - s[(size_t)(y1 + 1) * W1 + x0] +
         s[(size_t)y0 * W1 + x0];
}

// The lazily evaluated predicate chain of watsor/filter/track.py:26.
__device__ uint32_t apply_filters(const CameraCfg* __restrict__ cam, wb_detection* d, bool write_zones) {
  uint32_t v = 0;
  const int lab = d->label;
  if (cam == nullptr) return lab > 0 ? WB_V_LABEL : 0u;
  if (cam->check_label) {
    if (!(lab > 0)) return v;
    v |= WB_V_LABEL;
  }
  bool present = lab >= 0 && lab < WB_MAX_LABELS && cam->present[lab];
  double conf_thr, area_thr;
  uint32_t allowed;
  if (present) {
    conf_thr = cam->conf[lab];
    area_thr = cam->area[lab];
    allowed = cam->has_zone_list[lab] ? cam->zone_bits[lab] : 0xFFFFFFFFu;
  } else if (cam->default_present) {
    present = true;
    conf_thr = cam->default_conf;
    area_thr = cam->default_area;
    allowed = cam->default_has_zone_list ? cam->default_zone_bits : 0xFFFFFFFFu;
  } else {
    return v;  // confidence.py:18 / area.py:21: `... is not None and ...`
  }
  // confidence.py:17-19
  if (!(d->confidence >= conf_thr)) return v;
  v |= WB_V_CONFIDENCE;
  // area.py:20-26  abs((x_max - x_min + 1) * (y_max - y_min + 1)) >= pct/100 * W*H
  const wb_bounding_box bb = d->bounding_box;
  long long a = (long long)(bb.x_max - bb.x_min + 1) * (long long)(bb.y_max - bb.y_min + 1);
  if (a < 0) a = -a;
  if (!((double)a >= area_thr)) return v;
  v |= WB_V_AREA;
  if (cam->has_mask) {
    // mask.py:44-59: closed bbox rectangle intersects zone polygon  <=>  it covers >= 1 pixel of the
    // filled-contour raster (summed-area table, 4 loads per zone)
    int xa = min(bb.x_min, bb.x_max), xb = max(bb.x_min, bb.x_max);
    int ya = min(bb.y_min, bb.y_max), yb = max(bb.y_min, bb.y_max);
    xa = max(xa, 0);
    ya = max(ya, 0);
    xb = min(xb, cam->width - 1);
    yb = min(yb, cam->height - 1);
    bool hit = false;
    int z = 0;
    if (xa <= xb && ya <= yb) {
      const int W1 = cam->width + 1;
      const size_t plane = (size_t)(cam->height + 1) * W1;
      for (int p = 0; p < cam->n_zones && z < WB_MAX_ZONES; ++p) {
        if (!((allowed >> p) & 1u)) continue;
        if (sat_count(cam->sat + p * plane, W1, xa, ya, xb, yb) > 0) {
          if (write_zones) d->zones[z] = p + 1;
          ++z;
          hit = true;
        }
      }
    }
    if (!hit) return v;
    v |= WB_V_MASK;
  }
  return v | WB_V_PASS;
}

// One block per frame.  (1) the merge keys of all classes (C x max_per_class, zero = not selectable) are streamed
// once: a 2048-bin histogram of their top bits finds the bin that holds the max_total-th largest key; (2) the keys at or
// above that bin (normally ~max_total of them) are gathered into shared memory and sorted descending (bitonic) =
// `SortByField` (TopKV2, ties -> lower concat index: the keys carry (class, rank) below the score, so they are distinct
// and the descending key order IS the C-way merge order of the per-class lists) restricted to boxes with positive
// clipped area, top max_total; (3) one thread per output row: clip, `add` +1, tensorflow_cpu.py:79-90 integer
// conversion, predicates, Detection write.
constexpr int MERGE_THREADS = 512;
constexpr int MERGE_BINS = 2048;   // sign + exponent + 2 mantissa bits of the score
constexpr int MERGE_CAP = 4096;    // gathered keys that fit the fast path (ties in the cut bin can exceed max_total)

__global__ void __launch_bounds__(MERGE_THREADS)
    k_merge_filter(PostParams pp, const float4* __restrict__ dec, const int* __restrict__ sel_count,
                   const unsigned long long* __restrict__ sel_key, const int* __restrict__ sel_idx,
                   const FrameDesc* __restrict__ frames, const CameraCfg* __restrict__ cams, uint32_t flags,
                   wb_detection* __restrict__ out, uint32_t* __restrict__ verdicts, float* __restrict__ raw_boxes,
                   float* __restrict__ raw_scores, float* __restrict__ raw_classes, int* __restrict__ raw_num) {
  extern __shared__ unsigned long long s_keys[];  // [cap]: gathered keys (cap = MERGE_CAP, or all keys in the slow path)
  __shared__ int s_hist[MERGE_BINS];
  __shared__ unsigned long long s_win[128];
  __shared__ int s_nvalid, s_cut, s_n;
  const int f = blockIdx.x, C = pp.num_classes, MP = pp.max_per_class, N = pp.num_anchors;
  const unsigned long long* keys = sel_key + (size_t)f * C * MP;
  const int total = C * MP;
  const int want = min(pp.max_total, 128);
  for (int i = threadIdx.x; i < MERGE_BINS; i += blockDim.x) s_hist[i] = 0;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  // keys beyond a class's selection count were zeroed by k_nms, so the whole [C][MP] block can be read blindly
  for (int i0 = 0; i0 < total; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const unsigned long long k = i < total ? __ldg(keys + i) : 0ull;
    const int bin = k != 0ull ? (int)(k >> 53) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, bin);  // equal bins inside the warp: one atomic
    if (bin >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[bin], __popc(peers));
  }
  __syncthreads();
  if (threadIdx.x < 32) {  // highest bins first until `want` keys are covered
    const int lane = threadIdx.x;
    int acc = 0, cut = 0;
    for (int top = MERGE_BINS - 1; top >= 0 && acc < want; top -= 32) {
      const int b = top - lane;
      const int v = b >= 0 ? s_hist[b] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned reach = __ballot_sync(0xffffffffu, acc + incl >= want);
      if (reach) {
        cut = top - (__ffs(reach) - 1);
        acc = want;
        break;
      }
      acc += __shfl_sync(0xffffffffu, incl, 31);
      cut = max(top - 31, 0);
    }
    if (lane == 0) s_cut = cut;
  }
  __syncthreads();
  const int cut = s_cut;
  int cap = MERGE_CAP;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const unsigned long long k = __ldg(keys + i);  // second pass, L2 / L1 resident
    if (k != 0ull && (int)(k >> 53) >= cut) {
      const int p = atomicAdd(&s_n, 1);
      if (p < cap) s_keys[p] = k;
    }
  }
  __syncthreads();
  int n_g = s_n;
  if (n_g > cap) {
    // pathological: thousands of keys share the cut bin (e.g. every score identical).  Sort everything; the
    // dynamic shared memory was sized for the whole [C][MP] block by the launcher.
    __syncthreads();
    int P = 32;
    while (P < total) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_keys[i] = i < total ? __ldg(keys + i) : 0ull;
    n_g = total;
    __syncthreads();
    bitonic_desc(s_keys, P);
  } else {
    int P = 32;
    while (P < n_g) P <<= 1;
    for (int i = n_g + threadIdx.x; i < P; i += blockDim.x) s_keys[i] = 0ull;
    __syncthreads();
    bitonic_desc(s_keys, P);
  }
  if (threadIdx.x < 128) {
    const unsigned long long k = threadIdx.x < min(n_g, want) ? s_keys[threadIdx.x] : 0ull;
    s_win[threadIdx.x] = k;
  }
  if (threadIdx.x == 0) {
    // number of valid rows = non-zero keys among the first `want` (zeros sort last)
    int lo = 0, hi = min(n_g, want);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_keys[mid] != 0ull) lo = mid + 1; else hi = mid;
    }
    s_nvalid = lo;
  }
  __syncthreads();
  const int nvalid = s_nvalid;
  const int r = threadIdx.x;
  if (r == 0 && raw_num) raw_num[f] = nvalid;
  if (r >= WB_MAX_DETECTIONS) return;
  float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
  float score = 0.f, cls = 0.f;
  if (r < nvalid && r < pp.max_total) {
    unsigned long long k = s_win[r];
    int c = 0xFFFF - (int)((k >> 16) & 0xFFFFull), rank = 0xFFFF - (int)(k & 0xFFFFull);
    score = __uint_as_float((unsigned)(k >> 32));
    cls = (float)c;
    int idx = sel_idx[((size_t)f * C + c) * MP + rank];
    box = clip_unit(dec[(size_t)f * N + idx]);
  }
  cls = __fadd_rn(cls, pp.class_offset);  // graph node `add` (+1 also on the zero padding)
  if (raw_boxes) {
    reinterpret_cast<float4*>(raw_boxes)[(size_t)f * WB_MAX_DETECTIONS + r] = box;
    raw_scores[(size_t)f * WB_MAX_DETECTIONS + r] = score;
    raw_classes[(size_t)f * WB_MAX_DETECTIONS + r] = cls;
  }
  const FrameDesc fd = frames[f];
  wb_detection d;
  d.label = (int)cls;
  for (int z = 0; z < WB_MAX_ZONES; ++z) d.zones[z] = 0;
  d.confidence = (double)score;
  // int(np.float32 * int): exact product (float64), truncation toward zero
  const double mh = (double)(fd.h - 1), mw = (double)(fd.w - 1);
  d.bounding_box.y_min = (int)((double)box.x * mh);
  d.bounding_box.x_min = (int)((double)box.y * mw);
  d.bounding_box.y_max = (int)((double)box.z * mh);
  d.bounding_box.x_max = (int)((double)box.w * mw);
  const CameraCfg* cam = (cams != nullptr && fd.cam >= 0) ? cams + fd.cam : nullptr;
  uint32_t v = apply_filters(cam, &d, (flags & WB_F_FUSE_FILTERS) != 0);
  out[(size_t)f * WB_MAX_DETECTIONS + r] = d;
  if (verdicts) verdicts[(size_t)f * WB_MAX_DETECTIONS + r] = v;
}

void launch_post(const LaunchCtx& lc, int n, const PostParams& pp, const float* enc, const float* logits,
                 const float* anchors, const FrameDesc* frames, const CameraCfg* cams, uint32_t flags,
                 float* dec_boxes, int* cand_count, unsigned long long* cand, int* sel_count,
                 unsigned long long* sel, wb_detection* out, uint32_t* verdicts, float* raw_boxes,
                 float* raw_scores, float* raw_classes, int* raw_num, int* kept_hist) {
  const int C = pp.num_classes, N = pp.num_anchors;
  cudaMemsetAsync(cand_count, 0, sizeof(int) * (size_t)n * C, lc.stream);
  cudaMemsetAsync(kept_hist, 0, sizeof(int) * (size_t)n * KEPT_BINS, lc.stream);
  dim3 g1((N + DEC_TILE - 1) / DEC_TILE, n);
  const size_t dec_smem = (size_t)DEC_TILE * (C + 1) * (sizeof(float) + sizeof(short)) + 2 * sizeof(int) * C + 16;
  k_decode_scores<<<g1, 256, dec_smem, lc.stream>>>(pp, enc, logits, anchors, reinterpret_cast<float4*>(dec_boxes),
                                                    cand_count, cand);
  ++*lc.launch_counter;
  int sort_cap = 32;
  while (sort_cap < N) sort_cap <<= 1;
  // sel holds keys [n][C][max_per_class] followed by anchor indices (int) of the same shape
  unsigned long long* sel_key = sel;
  int* sel_idx = reinterpret_cast<int*>(sel + (size_t)n * C * pp.max_per_class);
  static PerDeviceFlag attr_done;
  if (!attr_done.get()) {
    cudaFuncSetAttribute(k_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024);
    cudaFuncSetAttribute(k_decode_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_merge_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done.set();
  }
  k_nms<<<dim3(C, n), 256, sizeof(unsigned long long) * sort_cap, lc.stream>>>(
      pp, reinterpret_cast<const float4*>(dec_boxes), cand_count, cand, sort_cap, sel_count, sel_key, sel_idx, kept_hist);
  ++*lc.launch_counter;
  {
    int P = 32;
    while (P < C * pp.max_per_class) P <<= 1;
    const size_t merge_smem = sizeof(unsigned long long) * (size_t)std::max(P, MERGE_CAP);
    k_merge_filter<<<n, MERGE_THREADS, merge_smem, lc.stream>>>(
        pp, reinterpret_cast<const float4*>(dec_boxes), sel_count, sel_key, sel_idx, frames, cams, flags, out,
        verdicts, raw_boxes, raw_scores, raw_classes, raw_num);
  }
  ++*lc.launch_counter;
}

// ---------------------------------------------------------------------------------------------------
// stand-alone predicate chain on caller rows (ConfidenceFilter / AreaFilter / MaskFilter __call__)
__global__ void k_filter_rows(const CameraCfg* __restrict__ cam, int n_rows, wb_detection* __restrict__ rows,
                              uint32_t* __restrict__ verdicts) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  wb_detection d = rows[r];
  uint32_t v = apply_filters(cam, &d, true);
  rows[r] = d;
  verdicts[r] = v;
}
void launch_filter_rows(const LaunchCtx& lc, const CameraCfg* cam, int n_rows, wb_detection* rows,
                        uint32_t* verdicts) {
  k_filter_rows<<<(n_rows + 127) / 128, 128, 0, lc.stream>>>(cam, n_rows, rows, verdicts);
  ++*lc.launch_counter;
}

// ---------------------------------------------------------------------------------------------------
// summed-area tables of the zone rasters (MaskFilter.__init__): sat[z][y][x] = #zone pixels in
// rows < y, cols < x.  Two passes (row scan, column scan); set-up time only.
__global__ void k_sat_rows(const uint8_t* __restrict__ raster, int n_zones, int h, int w, int32_t* __restrict__ sat) {
  int y = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
  if (y > h) return;
  int32_t* row = sat + ((size_t)z * (h + 1) + y) * (w + 1);
  row[0] = 0;
  if (y == 0) {
    for (int x = 1; x <= w; ++x) row[x] = 0;
    return;
  }
  const uint8_t* src = raster + ((size_t)z * h + (y - 1)) * w;
  int run = 0;
  for (int x = 0; x < w; ++x) {
    run += src[x] ? 1 : 0;
    row[x + 1] = run;
  }
}
__global__ void k_sat_cols(int n_zones, int h, int w, int32_t* __restrict__ sat) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
  if (x > w) return;
  int32_t* base = sat + (size_t)z * (h + 1) * (w + 1) + x;
  int run = 0;
  for (int y = 0; y <= h; ++y) {
    run += base[(size_t)y * (w + 1)];
    base[(size_t)y * (w + 1)] = run;
  }
}
void launch_build_sat(const LaunchCtx& lc, const uint8_t* raster, int n_zones, int h, int w, int32_t* sat) {
  k_sat_rows<<<dim3((h + 1 + 127) / 128, n_zones), 128, 0, lc.stream>>>(raster, n_zones, h, w, sat);
  k_sat_cols<<<dim3((w + 1 + 127) / 128, n_zones), 128, 0, lc.stream>>>(n_zones, h, w, sat);
  *lc.launch_counter += 2;
}
