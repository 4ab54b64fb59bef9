// wb_api.cu -- the C-ABI of libwatsor_b200.so (include/watsor_b200.h): context, model upload,
// per-camera filter state, the layer-program executor, two-slot asynchronous pipeline and the
// stage-level entry points the parity tests use.
#include <dlfcn.h>
#include <nccl.h>  // types only: the functions are bound with dlopen in wb_comm_*

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_tc.cuh"

#define WB_MAX_CAMERAS 256
#define WB_SLOTS 6

static thread_local std::string g_err;
static int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                  std::to_string(__LINE__) + ")");                                            \
  } while (0)
#define REQUIRE(cond, msg) \
  do {                     \
    if (!(cond)) return fail(msg); \
  } while (0)

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  FrameDesc* h_desc = nullptr;  // pinned
  FrameDesc* d_desc = nullptr;
  uint8_t* d_frames = nullptr;
  size_t d_frames_cap = 0;
  void* arena = nullptr;
  float* d_pre = nullptr;
  float *d_enc = nullptr, *d_logits = nullptr, *d_dec = nullptr;
  float* d_partial = nullptr;  // split-K scratch
  size_t partial_floats = 0;
  int* d_tile_counters = nullptr;  // split-K arrival tickets (zero between launches)
  int *d_cand_count = nullptr, *d_sel_count = nullptr;
  int* d_kept_hist = nullptr;  // [B][1024] per-frame histogram of kept scores (exact NMS early exit)
  unsigned long long *d_cand = nullptr, *d_sel = nullptr;
  wb_detection *d_out = nullptr, *h_out = nullptr;
  uint32_t *d_verdicts = nullptr, *h_verdicts = nullptr;
  float *d_raw = nullptr, *h_raw = nullptr;  // boxes[n][100][4] scores[n][100] classes[n][100]
  int *d_raw_num = nullptr, *h_raw_num = nullptr;
  int n = 0;
  uint32_t flags = 0;
  bool busy = false;
  int launches = 0;
  // CUDA graph of the kernel sequence, keyed by (n, flags)
  cudaGraphExec_t graph_exec = nullptr;
  int graph_n = -1;
  uint32_t graph_flags = 0;
  int graph_src_w = 0;  // widest camera the captured stem kernel sized its staging for
};

struct wb_ctx {
  int device = 0;
  int max_batch = 0;
  int precision = 0;
  bool use_graph = true;
  int max_src_w = 0;  // widest configured camera: sizes the stem's shared-memory staging of source rows
  // frame scatter (wb_comm_*): NCCL communicator bound at run time, its stream and the event the slots wait on
  void* comm = nullptr;
  int comm_rank = -1, comm_world = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t comm_ev = nullptr;
  cudaDeviceProp prop;
  wb_model_header hdr;
  std::vector<wb_layer> layers;
  std::vector<wb_tensor_entry> tensors;
  float* d_weights = nullptr;
  TcWeights tc;  // bf16 copies of the GEMM weights (precision 1)
  PostParams pp;
  CameraCfg* d_cams = nullptr;
  std::vector<CameraCfg> h_cams;
  std::vector<int32_t*> cam_sat;
  Slot slots[WB_SLOTS];
  cudaStream_t user_stream = nullptr;
  bool has_user_stream = false;
  int last_launches = 0;
  std::vector<void*> registered;
  // One lock for everything that touches shared per-context state: the camera table (h_cams / d_cams / SATs), slot 0's
  // staging buffers used by the stage-level calls, the filter staging buffers and the registration list.  The
  // reference runs one DetectionSieve thread per camera in one process (ref: watsor/main.py:378-384) and ctypes drops
  // the GIL, so wb_filter_rows / wb_set_camera DO get called concurrently.  Held for the host-side enqueue only,
  // except in the synchronous calls (filter_rows, set_camera, stage-level test hooks), which hold it until their
  // results are back.
  std::mutex mu;
  // wb_filter_rows: its own stream and staging buffers (never slot 0's: a detector batch may be in flight there)
  cudaStream_t fstream = nullptr;
  wb_detection* d_frows = nullptr;
  uint32_t* d_fverd = nullptr;
  int frows_cap = 0;

  const float* tensor(int idx) const { return d_weights + tensors[idx].offset; }
  size_t elem_size() const { return precision == 1 ? 2 : 4; }
  cudaStream_t stream_of(int s) { return has_user_stream ? user_stream : slots[s].stream; }
};

extern "C" {

int wb_abi_version(void) { return WB_ABI_VERSION; }
const char* wb_last_error(void) { return g_err.c_str(); }

int wb_device_count(int* count) {
  REQUIRE(count != nullptr, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  return 0;
}

static int alloc_slot(wb_ctx* c, Slot& s) {
  const int B = c->max_batch, N = c->hdr.num_anchors, C = c->hdr.num_classes, MP = c->hdr.max_per_class;
  CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&s.ev0));
  CK(cudaEventCreate(&s.ev1));
  CK(cudaMallocHost(&s.h_desc, sizeof(FrameDesc) * B));
  CK(cudaMalloc(&s.d_desc, sizeof(FrameDesc) * B));
  CK(cudaMalloc(&s.arena, (size_t)B * c->hdr.arena_elems * c->elem_size() + 1024));
  CK(cudaMalloc(&s.d_pre, sizeof(float) * (size_t)B * c->hdr.input_h * c->hdr.input_w * 3));
  CK(cudaMalloc(&s.d_enc, sizeof(float) * (size_t)B * N * 4));
  CK(cudaMalloc(&s.d_logits, sizeof(float) * (size_t)B * N * (C + 1)));
  CK(cudaMalloc(&s.d_dec, sizeof(float) * (size_t)B * N * 4));
  s.partial_floats = (size_t)4 * 1024 * 1024 + (size_t)B * 512 * 1024;
  CK(cudaMalloc(&s.d_partial, sizeof(float) * s.partial_floats));
  CK(cudaMalloc(&s.d_tile_counters, sizeof(int) * 4096));
  CK(cudaMemset(s.d_tile_counters, 0, sizeof(int) * 4096));
  CK(cudaMalloc(&s.d_cand_count, sizeof(int) * (size_t)B * C));
  CK(cudaMalloc(&s.d_sel_count, sizeof(int) * (size_t)B * C));
  CK(cudaMalloc(&s.d_kept_hist, sizeof(int) * (size_t)B * 1024));
  CK(cudaMalloc(&s.d_cand, sizeof(unsigned long long) * (size_t)B * C * N));
  CK(cudaMalloc(&s.d_sel, (sizeof(unsigned long long) + sizeof(int)) * (size_t)B * C * MP));
  CK(cudaMalloc(&s.d_out, sizeof(wb_detection) * (size_t)B * WB_MAX_DETECTIONS));
  CK(cudaMallocHost(&s.h_out, sizeof(wb_detection) * (size_t)B * WB_MAX_DETECTIONS));
  CK(cudaMalloc(&s.d_verdicts, sizeof(uint32_t) * (size_t)B * WB_MAX_DETECTIONS));
  CK(cudaMallocHost(&s.h_verdicts, sizeof(uint32_t) * (size_t)B * WB_MAX_DETECTIONS));
  CK(cudaMalloc(&s.d_raw, sizeof(float) * (size_t)B * WB_MAX_DETECTIONS * 6));
  CK(cudaMallocHost(&s.h_raw, sizeof(float) * (size_t)B * WB_MAX_DETECTIONS * 6));
  CK(cudaMalloc(&s.d_raw_num, sizeof(int) * B));
  CK(cudaMallocHost(&s.h_raw_num, sizeof(int) * B));
  return 0;
}

int wb_create(int device, const void* model_blob, size_t blob_bytes, int max_batch, int precision,
              wb_ctx** out) {
  REQUIRE(out != nullptr && model_blob != nullptr, "NULL argument");
  REQUIRE(blob_bytes >= sizeof(wb_model_header), "model blob too small");
  REQUIRE(max_batch >= 1 && max_batch <= 4096, "max_batch out of range");
  REQUIRE(precision >= 0 && precision <= 3,
          "precision must be 0 (fp32 CUDA cores), 1 (bf16 tcgen05), 2 (fp32 via 3xTF32 tcgen05) or 3 (1xTF32, diagnostic)");
  CK(cudaSetDevice(device));
  struct CtxFree {
    void operator()(wb_ctx* p) const { wb_destroy(p); }
  };
  std::unique_ptr<wb_ctx, CtxFree> guard(new wb_ctx());  // every REQUIRE / CK early return below frees it
  wb_ctx* c = guard.get();
  c->device = device;
  c->max_batch = max_batch;
  c->precision = precision;
  if (const char* g = getenv("WB_NO_GRAPH")) c->use_graph = !(g[0] == '1');
  CK(cudaGetDeviceProperties(&c->prop, device));
  REQUIRE(c->prop.major == 10, std::string("libwatsor_b200 is built for sm_100a only; device is ") +
                                   c->prop.name + " (sm_" + std::to_string(c->prop.major) +
                                   std::to_string(c->prop.minor) + ")");
  memcpy(&c->hdr, model_blob, sizeof(wb_model_header));
  REQUIRE(memcmp(c->hdr.magic, WB_MODEL_MAGIC, 8) == 0, "bad model blob magic");
  const uint8_t* p = static_cast<const uint8_t*>(model_blob) + sizeof(wb_model_header);
  size_t need = sizeof(wb_model_header) + (size_t)c->hdr.n_layers * sizeof(wb_layer) +
                (size_t)c->hdr.n_tensors * sizeof(wb_tensor_entry);
  REQUIRE(blob_bytes >= need, "model blob truncated (tables)");
  c->layers.resize(c->hdr.n_layers);
  memcpy(c->layers.data(), p, c->layers.size() * sizeof(wb_layer));
  p += c->layers.size() * sizeof(wb_layer);
  c->tensors.resize(c->hdr.n_tensors);
  memcpy(c->tensors.data(), p, c->tensors.size() * sizeof(wb_tensor_entry));
  p += c->tensors.size() * sizeof(wb_tensor_entry);
  size_t floats = 0;
  for (auto& t : c->tensors) floats = std::max<size_t>(floats, t.offset + ((t.count + 63) / 64) * 64);
  REQUIRE(blob_bytes >= need + floats * sizeof(float), "model blob truncated (data)");
  REQUIRE(c->hdr.num_classes >= 1 && c->hdr.num_classes <= WB_MAX_LABELS, "num_classes must be in 1..128");
  REQUIRE(c->hdr.max_per_class >= 1 && c->hdr.max_per_class <= 128, "max_per_class must be in 1..128");
  REQUIRE(c->hdr.max_total >= 1 && c->hdr.max_total <= 128, "max_total must be in 1..128");
  REQUIRE(c->hdr.score_thr >= 0.f, "negative score threshold is not supported");
  REQUIRE(c->hdr.iou_thr >= 0.f, "negative IoU threshold is not supported");
  REQUIRE(c->hdr.num_anchors >= 1 && c->hdr.num_anchors <= 8192, "num_anchors must be in 1..8192 (NMS sort buffers live in shared memory)");
  for (auto& L : c->layers) {
    if (L.op == WB_OP_PW || L.op == WB_OP_HEAD)
      REQUIRE(L.in_c % 4 == 0, std::string("layer ") + L.name + ": in_c must be a multiple of 4");
    if (L.op == WB_OP_CONV)
      REQUIRE(L.in_c % 16 == 0, std::string("layer ") + L.name + ": KxK convs need in_c to be a multiple of 16");
    if (L.op == WB_OP_DW) REQUIRE(L.out_c % 4 == 0 && L.kh == 3 && L.kw == 3, "depthwise must be 3x3, C%4==0");
    if (L.op == WB_OP_PW || L.op == WB_OP_CONV) REQUIRE(L.out_c % 4 == 0, "out_c must be a multiple of 4");
    if (L.op == WB_OP_MAXPOOL || L.op == WB_OP_AVGPOOL)
      REQUIRE(L.out_c % 4 == 0 && L.in_c == L.out_c && L.kh >= 1 && L.kw >= 1, "pooling needs C % 4 == 0");
    if (L.op == WB_OP_COPY)
      REQUIRE(L.in_c % 4 == 0 && L.out_c % 4 == 0 && L.row_off % 4 == 0 && L.row_off + L.in_c <= L.out_c,
              "channel copy: slice must be 4-aligned and inside the destination");
    REQUIRE(L.op >= WB_OP_STEM && L.op <= WB_OP_COPY, std::string("layer ") + L.name + ": unknown op");
  }
  CK(cudaMalloc(&c->d_weights, floats * sizeof(float)));
  CK(cudaMemcpy(c->d_weights, p, floats * sizeof(float), cudaMemcpyHostToDevice));
  if (precision != 0) {
    std::string err;
    const int mode = precision == 1 ? TC_BF16 : (precision == 2 ? TC_TF32X3 : TC_TF32X1);
    if (tc_prepare_weights(c->layers, c->tensors, reinterpret_cast<const float*>(p), mode, &c->tc, &err))
      return fail("tensor-core weight preparation: " + err);
  }
  c->pp.num_anchors = c->hdr.num_anchors;
  c->pp.num_classes = c->hdr.num_classes;
  c->pp.scale_y = c->hdr.scale_y;
  c->pp.scale_x = c->hdr.scale_x;
  c->pp.scale_h = c->hdr.scale_h;
  c->pp.scale_w = c->hdr.scale_w;
  c->pp.logit_scale = c->hdr.logit_scale;
  c->pp.iou_thr = c->hdr.iou_thr;
  c->pp.score_thr = c->hdr.score_thr;
  c->pp.max_per_class = c->hdr.max_per_class;
  c->pp.max_total = c->hdr.max_total;
  c->pp.class_offset = c->hdr.class_offset;
  c->h_cams.assign(WB_MAX_CAMERAS, CameraCfg{});
  c->cam_sat.assign(WB_MAX_CAMERAS, nullptr);
  CK(cudaMalloc(&c->d_cams, sizeof(CameraCfg) * WB_MAX_CAMERAS));
  CK(cudaMemset(c->d_cams, 0, sizeof(CameraCfg) * WB_MAX_CAMERAS));
  for (int s = 0; s < WB_SLOTS; ++s)
    if (alloc_slot(c, c->slots[s])) return 1;
  c->frows_cap = WB_MAX_DETECTIONS * max_batch;
  CK(cudaStreamCreateWithFlags(&c->fstream, cudaStreamNonBlocking));
  CK(cudaMalloc(&c->d_frows, sizeof(wb_detection) * (size_t)c->frows_cap));
  CK(cudaMalloc(&c->d_fverd, sizeof(uint32_t) * (size_t)c->frows_cap));
  CK(cudaDeviceSynchronize());
  *out = guard.release();
  return 0;
}

int wb_destroy(wb_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  wb_comm_destroy(c);
  for (void* r : c->registered) cudaHostUnregister(r);
  for (auto& s : c->slots) {
    if (s.graph_exec) cudaGraphExecDestroy(s.graph_exec);
    cudaFree(s.d_desc);
    cudaFree(s.d_frames);
    cudaFree(s.arena);
    cudaFree(s.d_pre);
    cudaFree(s.d_enc);
    cudaFree(s.d_logits);
    cudaFree(s.d_dec);
    cudaFree(s.d_partial);
    cudaFree(s.d_tile_counters);
    cudaFree(s.d_cand_count);
    cudaFree(s.d_sel_count);
    cudaFree(s.d_kept_hist);
    cudaFree(s.d_cand);
    cudaFree(s.d_sel);
    cudaFree(s.d_out);
    cudaFree(s.d_verdicts);
    cudaFree(s.d_raw);
    cudaFree(s.d_raw_num);
    cudaFreeHost(s.h_desc);
    cudaFreeHost(s.h_out);
    cudaFreeHost(s.h_verdicts);
    cudaFreeHost(s.h_raw);
    cudaFreeHost(s.h_raw_num);
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  for (auto* p : c->cam_sat) cudaFree(p);
  cudaFree(c->d_cams);
  cudaFree(c->d_frows);
  cudaFree(c->d_fverd);
  if (c->fstream) cudaStreamDestroy(c->fstream);
  cudaFree(c->d_weights);
  tc_free_weights(&c->tc);
  delete c;
  return 0;
}

int wb_device_name(wb_ctx* c, char* buf, size_t n) {
  REQUIRE(c && buf && n > 0, "NULL argument");
  snprintf(buf, n, "%s (cuda:%d, sm_%d%d, %s)", c->prop.name, c->device, c->prop.major, c->prop.minor,
           c->precision == 1 ? "bf16 tcgen05" : (c->precision == 2 ? "fp32 3xTF32 tcgen05" : (c->precision == 3 ? "tf32 tcgen05" : "fp32")));
  return 0;
}

int wb_set_stream(wb_ctx* c, uint64_t stream) {
  REQUIRE(c, "NULL ctx");
  c->user_stream = reinterpret_cast<cudaStream_t>(stream);
  c->has_user_stream = stream != 0;
  for (auto& s : c->slots) s.graph_n = -1;  // graphs are stream-agnostic, but keep it simple
  return 0;
}

int wb_model_info(wb_ctx* c, int32_t* ih, int32_t* iw, int32_t* nc, int32_t* na, int32_t* nl) {
  REQUIRE(c, "NULL ctx");
  if (ih) *ih = c->hdr.input_h;
  if (iw) *iw = c->hdr.input_w;
  if (nc) *nc = c->hdr.num_classes;
  if (na) *na = c->hdr.num_anchors;
  if (nl) *nl = c->hdr.n_layers;
  return 0;
}

int wb_anchors(wb_ctx* c, float* out) {
  REQUIRE(c && out, "NULL argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpy(out, c->tensor(c->hdr.anchors_tensor), sizeof(float) * 4 * c->hdr.num_anchors,
                cudaMemcpyDeviceToHost));
  return 0;
}

int wb_set_camera(wb_ctx* c, int cam, int width, int height, int n_zones, const uint8_t* raster, int n_filters,
                  const wb_class_filter* filters, uint32_t cam_flags) {
  REQUIRE(c, "NULL ctx");
  REQUIRE(cam >= 0 && cam < WB_MAX_CAMERAS, "cam_id out of range (0..255)");
  REQUIRE(width > 0 && height > 0, "bad frame size");
  REQUIRE(n_zones >= 0 && n_zones <= WB_MAX_CAMERA_ZONES, "a mask may hold at most 32 zones");
  REQUIRE(n_zones == 0 || raster != nullptr, "zone_raster is NULL");
  REQUIRE(n_filters == 0 || filters != nullptr, "filters is NULL");
  std::lock_guard<std::mutex> lock(c->mu);
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());  // no batch may be reading the table while it changes
  CameraCfg cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.width = width;
  cfg.height = height;
  c->max_src_w = std::max(c->max_src_w, width);
  cfg.n_zones = n_zones;
  cfg.has_mask = raster != nullptr ? 1 : 0;
  cfg.check_label = (cam_flags & WB_CAM_NO_LABEL_CHECK) ? 0 : 1;
  for (int i = 0; i < n_filters; ++i) {
    const wb_class_filter& f = filters[i];
    if (f.label == -1) {
      cfg.default_present = 1;
      cfg.default_conf = f.confidence;
      cfg.default_area = f.area;
      cfg.default_has_zone_list = f.has_zone_list ? 1 : 0;
      cfg.default_zone_bits = f.zone_bits;
      continue;
    }
    REQUIRE(f.label >= 0 && f.label < WB_MAX_LABELS, "filter label out of range (0..127)");
    cfg.present[f.label] = 1;
    cfg.conf[f.label] = f.confidence;
    cfg.area[f.label] = f.area;
    cfg.has_zone_list[f.label] = f.has_zone_list ? 1 : 0;
    cfg.zone_bits[f.label] = f.zone_bits;
  }
  if (c->cam_sat[cam]) {
    CK(cudaFree(c->cam_sat[cam]));
    c->cam_sat[cam] = nullptr;
  }
  if (n_zones > 0) {
    size_t sat_elems = (size_t)n_zones * (height + 1) * (width + 1);
    CK(cudaMalloc(&c->cam_sat[cam], sat_elems * sizeof(int32_t)));
    uint8_t* d_r = nullptr;
    size_t rb = (size_t)n_zones * height * width;
    CK(cudaMalloc(&d_r, rb));
    CK(cudaMemcpy(d_r, raster, rb, cudaMemcpyHostToDevice));
    int lcnt = 0;
    LaunchCtx lc{c->slots[0].stream, &lcnt};
    launch_build_sat(lc, d_r, n_zones, height, width, c->cam_sat[cam]);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->slots[0].stream));
    CK(cudaFree(d_r));
    cfg.sat = c->cam_sat[cam];
  }
  c->h_cams[cam] = cfg;
  CK(cudaMemcpy(c->d_cams + cam, &cfg, sizeof(cfg), cudaMemcpyHostToDevice));
  return 0;
}

int wb_register_host(wb_ctx* c, void* ptr, size_t bytes) {
  REQUIRE(c && ptr && bytes, "NULL argument");
  std::lock_guard<std::mutex> lock(c->mu);
  CK(cudaSetDevice(c->device));
  CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  c->registered.push_back(ptr);
  return 0;
}
int wb_unregister_host(wb_ctx* c, void* ptr) {
  REQUIRE(c && ptr, "NULL argument");
  std::lock_guard<std::mutex> lock(c->mu);
  for (size_t i = 0; i < c->registered.size(); ++i)
    if (c->registered[i] == ptr) {
      CK(cudaHostUnregister(ptr));
      c->registered.erase(c->registered.begin() + i);
      return 0;
    }
  return fail("pointer was not registered");
}

}  // extern "C"

// Number of layers starting at `li` that run as one fused inverted-residual-block kernel (0 = none, 3 = expand +
// depthwise + projection, 4 = ... + Add), all inside [li, end).
static int irb_span(wb_ctx* c, int li, int end, int n) {
  if (c->precision != 2 || li + 2 >= end) return 0;
  const wb_layer& E = c->layers[li];
  const wb_layer& D = c->layers[li + 1];
  const wb_layer& P = c->layers[li + 2];
  if (E.op != WB_OP_PW || E.act != WB_ACT_RELU6 || D.op != WB_OP_DW || P.op != WB_OP_PW || P.act != WB_ACT_NONE) return 0;
  if (li + 3 < end && c->layers[li + 3].op == WB_OP_ADD &&
      (c->layers[li + 3].in_off == P.out_off || c->layers[li + 3].in2_off == P.out_off) &&
      fused_irb_supported(c->tc, li + 2, E, D, P, &c->layers[li + 3], n))
    return 4;
  // without the Add the projection output must not feed a later Add as a fused pair elsewhere: plain 3-layer block
  if (li + 3 < (int)c->layers.size() && c->layers[li + 3].op == WB_OP_ADD &&
      (c->layers[li + 3].in_off == P.out_off || c->layers[li + 3].in2_off == P.out_off))
    return 0;  // a residual block whose Add is outside the requested range: run it unfused
  return fused_irb_supported(c->tc, li + 2, E, D, P, nullptr, n) ? 3 : 0;
}

// ---------------------------------------------------------------------------------------------------
// the layer program.  `pre` != NULL feeds an already pre-processed input (wb_backbone); otherwise the
// fused stem samples the frames directly.  When `times` is given every launch is bracketed by events.
template <typename T>
static int run_layers(wb_ctx* c, Slot& s, cudaStream_t st, int n, const float* pre, int first_layer,
                      int last_layer) {
  LaunchCtx lc{st, &s.launches};
  T* arena = static_cast<T*>(s.arena);
  const int NA = c->hdr.num_anchors, C1 = c->hdr.num_classes + 1;
  const size_t end = last_layer < 0 ? c->layers.size() : (size_t)last_layer + 1;
  for (size_t li = (size_t)first_layer; li < end; ++li) {
    const wb_layer& L = c->layers[li];
    const T* in = arena + (size_t)L.in_off * n;
    const T* in2 = arena + (size_t)L.in2_off * n;
    T* outp = arena + (size_t)L.out_off * n;
    const float* w = L.w_tensor >= 0 ? c->tensor(L.w_tensor) : nullptr;
    const float* sc = L.scale_tensor >= 0 ? c->tensor(L.scale_tensor) : nullptr;
    const float* of = L.offset_tensor >= 0 ? c->tensor(L.offset_tensor) : nullptr;
    switch (L.op) {
      case WB_OP_STEM:
        launch_stem<T>(lc, s.d_desc, pre, n, L, c->hdr.input_h, c->hdr.input_w, c->hdr.pre_mul, c->hdr.pre_sub, w,
                       sc, of, outp, c->max_src_w);
        break;
      case WB_OP_DW: {
        // depthwise -> 1x1 pairs run as one tensor-core kernel when both layers are in the requested range
        if (c->precision == 2 && li + 1 < end) {
          const wb_layer& P = c->layers[li + 1];
          if (P.in_off == L.out_off && fused_dwpw_supported(c->tc, (int)li + 1, L, P, n)) {
            std::string err;
            if (fused_launch_dwpw(lc, c->tc, (int)li + 1, n, L, P, static_cast<const void*>(in), w, sc, of,
                                  c->tensor(P.scale_tensor), c->tensor(P.offset_tensor),
                                  static_cast<void*>(arena + (size_t)P.out_off * n), &err))
              return fail("layers " + std::string(L.name) + " + " + P.name + ": " + err);
            ++li;  // the 1x1 layer is done
            break;
          }
        }
        launch_dw<T>(lc, n, L, in, w, sc, of, outp);
        break;
      }
      case WB_OP_ADD:
        launch_add<T>(lc, (size_t)n * L.out_h * L.out_w * L.out_c, in, in2, outp);
        break;
      case WB_OP_MAXPOOL:
      case WB_OP_AVGPOOL:
        launch_pool<T>(lc, n, L, in, outp);
        break;
      case WB_OP_COPY:
        launch_copy_channels<T>(lc, n, L, in, outp);
        break;
      case WB_OP_PW:
      case WB_OP_CONV:
      case WB_OP_HEAD:
        // MobileNet-v2 inverted residual block (expand -> depthwise -> projection [-> Add]) as one kernel
        if (int span = irb_span(c, (int)li, (int)end, n)) {
          const wb_layer& D = c->layers[li + 1];
          const wb_layer& P = c->layers[li + 2];
          const wb_layer& last = c->layers[li + span - 1];
          std::string err;
          if (fused_launch_irb(lc, c->tc, (int)li + 2, n, L, D, P, span == 4, static_cast<const void*>(in), w, sc, of,
                               c->tensor(D.w_tensor), c->tensor(D.scale_tensor), c->tensor(D.offset_tensor),
                               c->tensor(P.scale_tensor), c->tensor(P.offset_tensor),
                               static_cast<void*>(arena + (size_t)last.out_off * n), &err))
            return fail("block " + std::string(L.name) + ": " + err);
          li += span - 1;
          break;
        }
        if (c->precision != 0 && tc_layer_supported(L)) {
          // MobileNet-v2 bottleneck: a linear projection followed by `Add(shortcut, projection)` runs as one kernel,
          // the shortcut is added in the GEMM epilogue (fp32 modes) and the Add layer is skipped
          const void* residual = nullptr;
          void* dst = static_cast<void*>(outp);
          bool fuse_add = false;
          if (c->precision != 1 && L.op == WB_OP_PW && L.act == WB_ACT_NONE && li + 1 < end && getenv("WB_NO_FUSE_ADD") == nullptr) {
            const wb_layer& A = c->layers[li + 1];
            // the fused kernel reads this layer's input while it writes the Add's output: they must not overlap
            // (model.py plan_arena keeps the input alive through the Add; older blobs may not)
            const unsigned long long a0 = L.in_off, a1 = a0 + (unsigned long long)L.in_h * L.in_w * L.in_c;
            const unsigned long long b0 = A.out_off, b1 = b0 + (unsigned long long)A.out_h * A.out_w * A.out_c;
            if (A.op == WB_OP_ADD && (A.in_off == L.out_off || A.in2_off == L.out_off) && A.in_off != A.in2_off &&
                !(a0 < b1 && b0 < a1)) {
              const uint32_t other = A.in_off == L.out_off ? A.in2_off : A.in_off;
              residual = static_cast<const void*>(arena + (size_t)other * n);
              dst = static_cast<void*>(arena + (size_t)A.out_off * n);
              fuse_add = true;
            }
          }
          std::string err;
          if (tc_launch_gemm(lc, c->tc, (int)li, n, L, static_cast<const void*>(in), sc, of, dst, s.d_enc, s.d_logits, NA,
                             C1, s.d_partial, s.partial_floats, s.d_tile_counters, residual, &err))
            return fail("layer " + std::string(L.name) + ": " + err);
          if (fuse_add) ++li;  // the Add layer is done
        } else {
          launch_gemm_cc<T>(lc, n, L, in, w, sc, of, outp, s.d_enc, s.d_logits, NA, C1, s.d_partial, s.partial_floats);
        }
        break;
      default:
        return fail("unknown layer op " + std::to_string(L.op));
    }
  }
  CK(cudaGetLastError());
  return 0;
}

static int run_post(wb_ctx* c, Slot& s, cudaStream_t st, int n, uint32_t flags, bool want_raw = false) {
  LaunchCtx lc{st, &s.launches};
  // the float boxes / scores / classes of `sess.run` are a test hook (wb_postprocess); the product path writes
  // Detection rows only
  float* rb = want_raw ? s.d_raw : nullptr;
  float* rs = want_raw ? rb + (size_t)c->max_batch * WB_MAX_DETECTIONS * 4 : nullptr;
  float* rc = want_raw ? rs + (size_t)c->max_batch * WB_MAX_DETECTIONS : nullptr;
  launch_post(lc, n, c->pp, s.d_enc, s.d_logits, c->tensor(c->hdr.anchors_tensor), s.d_desc, c->d_cams, flags,
              s.d_dec, s.d_cand_count, s.d_cand, s.d_sel_count, s.d_sel, s.d_out, s.d_verdicts, rb, rs, rc,
              want_raw ? s.d_raw_num : nullptr, s.d_kept_hist);
  CK(cudaGetLastError());
  return 0;
}

static int run_all(wb_ctx* c, Slot& s, cudaStream_t st, int n, uint32_t flags) {
  int rc = c->precision == 1 ? run_layers<__nv_bfloat16>(c, s, st, n, nullptr, 0, -1)
                             : run_layers<float>(c, s, st, n, nullptr, 0, -1);
  if (rc) return rc;
  return run_post(c, s, st, n, flags);
}

// kernels of one batch, through a CUDA graph when possible (launch-bound at small batch)
static int enqueue_kernels(wb_ctx* c, Slot& s, cudaStream_t st, int n, uint32_t flags) {
  const uint32_t gflags = flags & WB_F_FUSE_FILTERS;
  if (!c->use_graph) {
    s.launches = 0;
    return run_all(c, s, st, n, gflags);
  }
  if (s.graph_exec == nullptr || s.graph_n != n || s.graph_flags != gflags || s.graph_src_w != c->max_src_w) {
    if (s.graph_exec) {
      cudaGraphExecDestroy(s.graph_exec);
      s.graph_exec = nullptr;
    }
    // warm-up run outside capture (sets function attributes, validates launches)
    s.launches = 0;
    if (int rc = run_all(c, s, st, n, gflags)) return rc;
    CK(cudaStreamSynchronize(st));
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    s.launches = 0;
    int rc = run_all(c, s, st, n, gflags);
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    CK(cudaGraphInstantiate(&s.graph_exec, graph, 0));
    CK(cudaGraphDestroy(graph));
    s.graph_n = n;
    s.graph_flags = gflags;
    s.graph_src_w = c->max_src_w;
  }
  CK(cudaGraphLaunch(s.graph_exec, st));
  return 0;
}

static int fill_desc(wb_ctx* c, Slot& s, int n, const uint8_t* const* frames, const int32_t* cam_ids,
                     bool on_device, cudaStream_t st) {
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    int cam = cam_ids[i];
    REQUIRE(cam >= 0 && cam < WB_MAX_CAMERAS && c->h_cams[cam].width > 0,
            "cam_id " + std::to_string(cam) + " has not been configured with wb_set_camera");
    REQUIRE(frames == nullptr || frames[i] != nullptr, "NULL frame pointer");
    total += ((size_t)c->h_cams[cam].width * c->h_cams[cam].height * 3 + 255) / 256 * 256;
  }
  if (!on_device && frames != nullptr && total > s.d_frames_cap) {
    CK(cudaStreamSynchronize(st));
    if (s.d_frames) CK(cudaFree(s.d_frames));
    s.d_frames_cap = total + total / 4;
    CK(cudaMalloc(&s.d_frames, s.d_frames_cap));
  }
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const CameraCfg& cc = c->h_cams[cam_ids[i]];
    size_t bytes = (size_t)cc.width * cc.height * 3;
    FrameDesc d;
    d.w = cc.width;
    d.h = cc.height;
    d.cam = cam_ids[i];
    d._pad = 0;
    if (frames == nullptr) {
      d.ptr = nullptr;
    } else if (on_device) {
      d.ptr = frames[i];
    } else {
      d.ptr = s.d_frames + off;
      CK(cudaMemcpyAsync(s.d_frames + off, frames[i], bytes, cudaMemcpyHostToDevice, st));
      off += (bytes + 255) / 256 * 256;
    }
    s.h_desc[i] = d;
  }
  CK(cudaMemcpyAsync(s.d_desc, s.h_desc, sizeof(FrameDesc) * n, cudaMemcpyHostToDevice, st));
  return 0;
}

extern "C" {

int wb_submit(wb_ctx* c, int slot, int n, const uint8_t* const* frames, const int32_t* cam_ids, uint32_t flags) {
  REQUIRE(c && frames && cam_ids, "NULL argument");
  REQUIRE(slot >= 0 && slot < WB_SLOTS, "slot out of range");
  REQUIRE(n >= 1 && n <= c->max_batch, "batch size out of range (1..max_batch)");
  std::lock_guard<std::mutex> lock(c->mu);
  Slot& s = c->slots[slot];
  REQUIRE(!s.busy, "slot is busy: collect it first");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream_of(slot);
  CK(cudaEventRecord(s.ev0, st));
  if (int rc = fill_desc(c, s, n, frames, cam_ids, (flags & WB_F_FRAMES_ON_DEVICE) != 0, st)) return rc;
  if (int rc = enqueue_kernels(c, s, st, n, flags)) return rc;
  if (!(flags & WB_F_OUT_ON_DEVICE)) {
    CK(cudaMemcpyAsync(s.h_out, s.d_out, sizeof(wb_detection) * (size_t)n * WB_MAX_DETECTIONS,
                       cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(s.h_verdicts, s.d_verdicts, sizeof(uint32_t) * (size_t)n * WB_MAX_DETECTIONS,
                       cudaMemcpyDeviceToHost, st));
  }
  CK(cudaEventRecord(s.ev1, st));
  s.n = n;
  s.flags = flags;
  s.busy = true;
  c->last_launches = s.launches;
  return 0;
}

int wb_collect(wb_ctx* c, int slot, wb_detection* const* out, uint32_t* const* verdicts, float* gpu_ms) {
  REQUIRE(c, "NULL ctx");
  REQUIRE(slot >= 0 && slot < WB_SLOTS, "slot out of range");
  Slot& s = c->slots[slot];
  {
    std::lock_guard<std::mutex> lock(c->mu);
    REQUIRE(s.busy, "slot has no batch in flight");
    s.busy = false;
  }
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(s.ev1));  // not under the lock: other threads keep submitting to other slots
  if (gpu_ms) CK(cudaEventElapsedTime(gpu_ms, s.ev0, s.ev1));
  if (s.flags & WB_F_OUT_ON_DEVICE) {
    cudaStream_t st = c->stream_of(slot);
    for (int i = 0; i < s.n; ++i) {
      if (out && out[i])
        CK(cudaMemcpyAsync(out[i], s.d_out + (size_t)i * WB_MAX_DETECTIONS, sizeof(wb_detection) * WB_MAX_DETECTIONS,
                           cudaMemcpyDeviceToDevice, st));
      if (verdicts && verdicts[i])
        CK(cudaMemcpyAsync(verdicts[i], s.d_verdicts + (size_t)i * WB_MAX_DETECTIONS,
                           sizeof(uint32_t) * WB_MAX_DETECTIONS, cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaStreamSynchronize(st));
    return 0;
  }
  for (int i = 0; i < s.n; ++i) {
    if (out && out[i])
      memcpy(out[i], s.h_out + (size_t)i * WB_MAX_DETECTIONS, sizeof(wb_detection) * WB_MAX_DETECTIONS);
    if (verdicts && verdicts[i])
      memcpy(verdicts[i], s.h_verdicts + (size_t)i * WB_MAX_DETECTIONS, sizeof(uint32_t) * WB_MAX_DETECTIONS);
  }
  return 0;
}

int wb_stream_fence(wb_ctx* c, uint64_t stream, int direction) {
  REQUIRE(c, "NULL ctx");
  REQUIRE(direction == 0 || direction == 1, "direction must be 0 or 1");
  CK(cudaSetDevice(c->device));
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
  cudaEvent_t ev;
  CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  if (direction == 0) {
    CK(cudaEventRecord(ev, user));
    for (auto& s : c->slots) CK(cudaStreamWaitEvent(s.stream, ev, 0));
  } else {
    for (auto& s : c->slots) {
      CK(cudaEventRecord(ev, s.stream));
      CK(cudaStreamWaitEvent(user, ev, 0));
    }
  }
  CK(cudaEventDestroy(ev));
  return 0;
}

int wb_detect(wb_ctx* c, int n, const uint8_t* const* frames, const int32_t* cam_ids, uint32_t flags,
              wb_detection* const* out, uint32_t* const* verdicts, float* gpu_ms) {
  if (int rc = wb_submit(c, 0, n, frames, cam_ids, flags)) return rc;
  return wb_collect(c, 0, out, verdicts, gpu_ms);
}

// ---------------------------------------------------------------------------------------------------
int wb_preprocess(wb_ctx* c, int n, const uint8_t* const* frames, const int32_t* widths, const int32_t* heights,
                  float* out) {
  REQUIRE(c && frames && widths && heights && out, "NULL argument");
  REQUIRE(n >= 1 && n <= c->max_batch, "batch size out of range");
  std::lock_guard<std::mutex> lock(c->mu);
  CK(cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  REQUIRE(!s.busy, "slot 0 is busy");
  cudaStream_t st = c->stream_of(0);
  size_t total = 0;
  for (int i = 0; i < n; ++i) total += (size_t)widths[i] * heights[i] * 3;
  if (total > s.d_frames_cap) {
    if (s.d_frames) CK(cudaFree(s.d_frames));
    s.d_frames_cap = total;
    CK(cudaMalloc(&s.d_frames, total));
  }
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    size_t bytes = (size_t)widths[i] * heights[i] * 3;
    CK(cudaMemcpyAsync(s.d_frames + off, frames[i], bytes, cudaMemcpyHostToDevice, st));
    s.h_desc[i] = FrameDesc{s.d_frames + off, widths[i], heights[i], -1, 0};
    off += bytes;
  }
  CK(cudaMemcpyAsync(s.d_desc, s.h_desc, sizeof(FrameDesc) * n, cudaMemcpyHostToDevice, st));
  LaunchCtx lc{st, &s.launches};
  int max_w = 0;
  for (int i = 0; i < n; ++i) max_w = std::max(max_w, (int)widths[i]);
  launch_preprocess_f32(lc, s.d_desc, n, s.d_pre, c->hdr.input_h, c->hdr.input_w, c->hdr.pre_mul, c->hdr.pre_sub,
                        max_w);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, s.d_pre, sizeof(float) * (size_t)n * c->hdr.input_h * c->hdr.input_w * 3,
                     cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int wb_backbone(wb_ctx* c, int n, const float* pre, float* enc, float* logits, int stop_layer, float* layer_out,
                size_t layer_out_floats) {
  REQUIRE(c && pre, "NULL argument");
  REQUIRE(n >= 1 && n <= c->max_batch, "batch size out of range");
  REQUIRE(stop_layer < (int)c->layers.size(), "stop_layer out of range");
  std::lock_guard<std::mutex> lock(c->mu);
  CK(cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  REQUIRE(!s.busy, "slot 0 is busy");
  cudaStream_t st = c->stream_of(0);
  const size_t pre_floats = (size_t)n * c->hdr.input_h * c->hdr.input_w * 3;
  CK(cudaMemcpyAsync(s.d_pre, pre, sizeof(float) * pre_floats, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(s.d_enc, 0, sizeof(float) * (size_t)n * c->hdr.num_anchors * 4, st));
  CK(cudaMemsetAsync(s.d_logits, 0, sizeof(float) * (size_t)n * c->hdr.num_anchors * (c->hdr.num_classes + 1), st));
  s.launches = 0;
  int rc = c->precision == 1 ? run_layers<__nv_bfloat16>(c, s, st, n, s.d_pre, 0, stop_layer)
                             : run_layers<float>(c, s, st, n, s.d_pre, 0, stop_layer);
  if (rc) return rc;
  if (enc) CK(cudaMemcpyAsync(enc, s.d_enc, sizeof(float) * (size_t)n * c->hdr.num_anchors * 4, cudaMemcpyDeviceToHost, st));
  if (logits)
    CK(cudaMemcpyAsync(logits, s.d_logits, sizeof(float) * (size_t)n * c->hdr.num_anchors * (c->hdr.num_classes + 1),
                       cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (stop_layer >= 0 && layer_out) {
    const wb_layer& L = c->layers[stop_layer];
    REQUIRE(L.op != WB_OP_HEAD, "head layers have no activation output");
    size_t elems = (size_t)n * L.out_h * L.out_w * L.out_c;
    REQUIRE(layer_out_floats >= elems, "layer_out too small");
    if (c->elem_size() == 4) {
      CK(cudaMemcpy(layer_out, static_cast<float*>(s.arena) + (size_t)L.out_off * n, elems * 4, cudaMemcpyDeviceToHost));
    } else {
      std::vector<uint16_t> tmp(elems);
      CK(cudaMemcpy(tmp.data(), static_cast<uint16_t*>(s.arena) + (size_t)L.out_off * n, elems * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < elems; ++i) {
        uint32_t u = (uint32_t)tmp[i] << 16;
        memcpy(&layer_out[i], &u, 4);
      }
    }
  }
  c->last_launches = s.launches;
  return 0;
}

int wb_postprocess(wb_ctx* c, int n, const float* enc, const float* logits, const int32_t* cam_ids, uint32_t flags,
                   wb_detection* const* out, uint32_t* const* verdicts, float* boxes, float* scores, float* classes,
                   int32_t* num) {
  REQUIRE(c && enc && logits && cam_ids, "NULL argument");
  REQUIRE(n >= 1 && n <= c->max_batch, "batch size out of range");
  std::lock_guard<std::mutex> lock(c->mu);
  CK(cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  REQUIRE(!s.busy, "slot 0 is busy");
  cudaStream_t st = c->stream_of(0);
  const int NA = c->hdr.num_anchors, C1 = c->hdr.num_classes + 1;
  CK(cudaMemcpyAsync(s.d_enc, enc, sizeof(float) * (size_t)n * NA * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.d_logits, logits, sizeof(float) * (size_t)n * NA * C1, cudaMemcpyHostToDevice, st));
  if (int rc = fill_desc(c, s, n, nullptr, cam_ids, false, st)) return rc;
  s.launches = 0;
  if (int rc = run_post(c, s, st, n, flags, true)) return rc;
  const size_t B = c->max_batch;
  CK(cudaMemcpyAsync(s.h_out, s.d_out, sizeof(wb_detection) * (size_t)n * WB_MAX_DETECTIONS, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(s.h_verdicts, s.d_verdicts, sizeof(uint32_t) * (size_t)n * WB_MAX_DETECTIONS, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(s.h_raw, s.d_raw, sizeof(float) * B * WB_MAX_DETECTIONS * 6, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(s.h_raw_num, s.d_raw_num, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int i = 0; i < n; ++i) {
    if (out && out[i]) memcpy(out[i], s.h_out + (size_t)i * WB_MAX_DETECTIONS, sizeof(wb_detection) * WB_MAX_DETECTIONS);
    if (verdicts && verdicts[i])
      memcpy(verdicts[i], s.h_verdicts + (size_t)i * WB_MAX_DETECTIONS, sizeof(uint32_t) * WB_MAX_DETECTIONS);
  }
  if (boxes) memcpy(boxes, s.h_raw, sizeof(float) * (size_t)n * WB_MAX_DETECTIONS * 4);
  if (scores) memcpy(scores, s.h_raw + B * WB_MAX_DETECTIONS * 4, sizeof(float) * (size_t)n * WB_MAX_DETECTIONS);
  if (classes) memcpy(classes, s.h_raw + B * WB_MAX_DETECTIONS * 5, sizeof(float) * (size_t)n * WB_MAX_DETECTIONS);
  if (num) memcpy(num, s.h_raw_num, sizeof(int) * n);
  c->last_launches = s.launches;
  return 0;
}

int wb_filter_rows(wb_ctx* c, int cam, int n_rows, wb_detection* rows, uint32_t* verdicts) {
  REQUIRE(c && rows && verdicts, "NULL argument");
  REQUIRE(n_rows >= 1 && n_rows <= c->frows_cap, "n_rows out of range");
  // thread-safe: one DetectionSieve thread per camera calls this concurrently (ref: watsor/main.py:378-384)
  std::lock_guard<std::mutex> lock(c->mu);
  REQUIRE(cam >= 0 && cam < WB_MAX_CAMERAS && c->h_cams[cam].width > 0, "camera has not been configured");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->fstream;
  CK(cudaMemcpyAsync(c->d_frows, rows, sizeof(wb_detection) * n_rows, cudaMemcpyHostToDevice, st));
  int launches = 0;
  LaunchCtx lc{st, &launches};
  launch_filter_rows(lc, c->d_cams + cam, n_rows, c->d_frows, c->d_fverd);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(rows, c->d_frows, sizeof(wb_detection) * n_rows, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(verdicts, c->d_fverd, sizeof(uint32_t) * n_rows, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int wb_last_launch_count(wb_ctx* c, int* launches) {
  REQUIRE(c && launches, "NULL argument");
  *launches = c->last_launches;
  return 0;
}

// runs the program once, un-graphed, with an event pair around every layer and one around the post
// stage: per-layer device times for bench.py's roofline (kinds[i] = layer op, 100 = post stage)
int wb_profile_layers(wb_ctx* c, int n, const uint8_t* const* device_frames, const int32_t* cam_ids,
                      float* ms, int32_t* kinds, int max_launches, int* n_out) {
  REQUIRE(c && device_frames && cam_ids && ms && kinds && n_out, "NULL argument");
  REQUIRE(n >= 1 && n <= c->max_batch, "batch size out of range");
  std::lock_guard<std::mutex> lock(c->mu);
  CK(cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  REQUIRE(!s.busy, "slot 0 is busy");
  cudaStream_t st = c->stream_of(0);
  if (int rc = fill_desc(c, s, n, device_frames, cam_ids, true, st)) return rc;
  const int nl = (int)c->layers.size();
  REQUIRE(max_launches >= nl + 1, "max_launches too small");
  std::vector<cudaEvent_t> ev(nl + 2);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  // every layer is launched REPS times back to back between two events (layers are idempotent: they
  // never write their own input), so host launch gaps do not leak into the per-kernel time
  const int REPS = 10;
  s.launches = 0;
  if (int rc = run_all(c, s, st, n, 0)) return rc;  // warm-up, also fills every activation buffer
  s.launches = 0;
  CK(cudaEventRecord(ev[0], st));
  for (int li = 0; li < nl; ++li) {
    // a depthwise layer that the executor fuses into the following 1x1 conv is timed together with it:
    // the pair's time is reported on the 1x1 layer, the depthwise entry reads 0
    int last = li;
    if (int span = irb_span(c, li, nl, n)) {
      // a fused inverted residual block: the kernel's time is reported on the expand entry, the others read 0
      for (int r = 0; r < REPS; ++r) {
        int rc = run_layers<float>(c, s, st, n, nullptr, li, li + span - 1);
        if (rc) return rc;
      }
      CK(cudaEventRecord(ev[li + 1], st));
      kinds[li] = (int)c->layers[li].op;
      for (int k = 1; k < span; ++k) {
        CK(cudaEventRecord(ev[li + k + 1], st));
        kinds[li + k] = (int)c->layers[li + k].op;
      }
      li += span - 1;
      continue;
    }
    if (c->precision == 2 && li + 1 < nl && c->layers[li].op == WB_OP_DW &&
        c->layers[li + 1].in_off == c->layers[li].out_off &&
        fused_dwpw_supported(c->tc, li + 1, c->layers[li], c->layers[li + 1], n))
      last = li + 1;
    // a linear projection + residual Add pair runs as one kernel too (same conditions as run_layers); the
    // pair's time is reported on the Add entry, the projection entry reads 0
    if (last == li && c->precision != 0 && c->precision != 1 && li + 1 < nl && c->layers[li].op == WB_OP_PW &&
        c->layers[li].act == WB_ACT_NONE && c->layers[li + 1].op == WB_OP_ADD && tc_layer_supported(c->layers[li]) &&
        (c->layers[li + 1].in_off == c->layers[li].out_off || c->layers[li + 1].in2_off == c->layers[li].out_off) &&
        getenv("WB_NO_FUSE_ADD") == nullptr)
      last = li + 1;
    const int first = li;
    const bool time_on_first = last != li && c->layers[li].op == WB_OP_PW;  // projection + Add: time on the GEMM
    if (last != li && !time_on_first) {
      CK(cudaEventRecord(ev[li + 1], st));  // zero-length interval for the depthwise entry
      kinds[li] = (int)c->layers[li].op;
      ++li;
    }
    for (int r = 0; r < REPS; ++r) {
      int rc = c->precision == 1 ? run_layers<__nv_bfloat16>(c, s, st, n, nullptr, first, last)
                                 : run_layers<float>(c, s, st, n, nullptr, first, last);
      if (rc) return rc;
    }
    CK(cudaEventRecord(ev[li + 1], st));
    kinds[li] = (int)c->layers[li].op;
    if (time_on_first) {
      ++li;
      CK(cudaEventRecord(ev[li + 1], st));  // zero-length interval for the Add entry
      kinds[li] = (int)c->layers[li].op;
    }
  }
  for (int r = 0; r < REPS; ++r)
    if (int rc = run_post(c, s, st, n, 0)) return rc;
  CK(cudaEventRecord(ev[nl + 1], st));
  CK(cudaEventSynchronize(ev[nl + 1]));
  for (int li = 0; li <= nl; ++li) {
    CK(cudaEventElapsedTime(&ms[li], ev[li], ev[li + 1]));
    ms[li] /= REPS;
  }
  s.launches /= REPS;
  kinds[nl] = 100;
  for (auto& e : ev) cudaEventDestroy(e);
  *n_out = nl + 1;
  c->last_launches = s.launches;
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Engine frame scatter over NCCL, bound with dlopen so that the library itself does not depend on libnccl.
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // one NCCL per process: the copy that is already loaded wins (torch brings its own libnccl.so.2 and cannot be
    // imported after a different one -- same SONAME); then WB_NCCL_LIB (the Python shim points it at the copy bundled
    // with torch, so a later `import torch` still works); then the system library
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char* override_path = getenv("WB_NCCL_LIB");
    if (!api.handle && override_path && override_path[0]) api.handle = dlopen(override_path, RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) {
      const char* e = dlerror();
      api.error = std::string("cannot load NCCL (libnccl.so.2): ") + (e ? e : "not found");
      return;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(api.handle, name);
      if (!p) {
        ok = false;
        api.error = std::string("NCCL symbol missing: ") + name;
      }
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) api.handle = nullptr;
  });
  return &api;
}
}  // namespace

#define NCCL_CK(api, call)                                                                      \
  do {                                                                                          \
    ncclResult_t _r = (call);                                                                   \
    if (_r != ncclSuccess) return fail(std::string(#call) + ": " + (api)->GetErrorString(_r));  \
  } while (0)

static_assert(sizeof(ncclUniqueId) == WB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");

int wb_comm_unique_id(uint8_t* id_out) {
  REQUIRE(id_out, "NULL id_out");
  NcclApi* api = nccl_api();
  if (!api->handle) return fail(api->error);
  ncclUniqueId id;
  NCCL_CK(api, api->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

int wb_comm_init(wb_ctx* c, int rank, int world, const uint8_t* id_bytes) {
  REQUIRE(c, "NULL ctx");
  REQUIRE(id_bytes, "NULL id");
  REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank must be in [0, world)");
  NcclApi* api = nccl_api();
  if (!api->handle) return fail(api->error);
  std::lock_guard<std::mutex> lock(c->mu);
  REQUIRE(c->comm == nullptr, "the context already has a communicator (wb_comm_destroy first)");
  CK(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm = nullptr;
  NCCL_CK(api, api->CommInitRank(&comm, world, id, rank));
  c->comm = comm;
  c->comm_rank = rank;
  c->comm_world = world;
  CK(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->comm_ev, cudaEventDisableTiming));
  return 0;
}

int wb_scatter_frames(wb_ctx* c, int root, const uint8_t* const* send_per_rank, uint8_t* recv, size_t bytes_per_rank,
                      uint64_t cuda_stream) {
  REQUIRE(c, "NULL ctx");
  REQUIRE(c->comm != nullptr, "wb_comm_init has not been called on this context");
  REQUIRE(root >= 0 && root < c->comm_world, "root out of range");
  REQUIRE(recv != nullptr && bytes_per_rank > 0, "NULL receive buffer / empty slab");
  const bool is_root = c->comm_rank == root;
  REQUIRE(!is_root || send_per_rank != nullptr, "the root rank must pass send_per_rank");
  if (is_root)
    for (int r = 0; r < c->comm_world; ++r) REQUIRE(send_per_rank[r] != nullptr, "NULL slab pointer");
  NcclApi* api = nccl_api();
  ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
  std::lock_guard<std::mutex> lock(c->mu);
  CK(cudaSetDevice(c->device));
  const bool own = cuda_stream == 0;
  cudaStream_t st = own ? c->comm_stream : reinterpret_cast<cudaStream_t>(cuda_stream);
  if (is_root) {
    NCCL_CK(api, api->GroupStart());
    for (int r = 0; r < c->comm_world; ++r)
      if (r != root) NCCL_CK(api, api->Send(send_per_rank[r], bytes_per_rank, ncclUint8, r, comm, st));
    NCCL_CK(api, api->GroupEnd());
    if (send_per_rank[root] != recv)
      CK(cudaMemcpyAsync(recv, send_per_rank[root], bytes_per_rank, cudaMemcpyDeviceToDevice, st));
  } else {
    NCCL_CK(api, api->Recv(recv, bytes_per_rank, ncclUint8, root, comm, st));
  }
  if (own) {
    CK(cudaEventRecord(c->comm_ev, st));
    for (auto& s : c->slots) CK(cudaStreamWaitEvent(s.stream, c->comm_ev, 0));
  }
  return 0;
}

int wb_comm_destroy(wb_ctx* c) {
  if (!c || !c->comm) return 0;
  NcclApi* api = nccl_api();
  cudaSetDevice(c->device);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  if (api->handle) api->CommDestroy(static_cast<ncclComm_t>(c->comm));
  c->comm = nullptr;
  c->comm_rank = -1;
  c->comm_world = 0;
  if (c->comm_ev) cudaEventDestroy(c->comm_ev);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  c->comm_ev = nullptr;
  c->comm_stream = nullptr;
  return 0;
}
