// C ABI (include/yolact_b200.h).  Every entry point converts C++ exceptions into a status code
// and a thread-local message; nothing here computes on the CPU beyond packing weights.
#include <stdlib.h>
#include <string.h>

#include "engine.cuh"

namespace yb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace yb

using namespace yb;

#define YB_API_BEGIN try {
#define YB_API_END                                   \
  }                                                  \
  catch (const yb::Error& e) {                       \
    yb::set_last_error(e.what());                    \
    return e.code;                                   \
  }                                                  \
  catch (const std::exception& e) {                  \
    yb::set_last_error(std::string("internal: ") + e.what()); \
    return YB_ERR_INVALID;                           \
  }                                                  \
  return YB_OK;

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) YB_CHECK_CUDA(cudaSetDevice(dev));
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

// OIHW fp32 (device) -> [K][Cout] (T), K = (r*KW+s)*Cin + c
template <typename T>
__global__ void pack_w_simt_kernel(const float* __restrict__ w, T* __restrict__ out, int Co, int Ci, int taps) {
  const int64_t total = (int64_t)Co * Ci * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % taps);
    int64_t r = i / taps;
    int c = (int)(r % Ci);
    int o = (int)(r / Ci);
    out[((int64_t)t * Ci + c) * Co + o] = from_f32<T>(w[i]);
  }
}
// OIHW fp32 (device) -> [Cout][tap*Cin + c] half (DCN contraction as a 1x1 conv over gathered columns)
__global__ void pack_w_dcn_tc_kernel(const float* __restrict__ w, __half* __restrict__ out, int Co, int Ci, int taps) {
  const int64_t total = (int64_t)Co * Ci * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % taps);
    int64_t r = i / taps;
    int c = (int)(r % Ci);
    int o = (int)(r / Ci);
    out[(int64_t)o * taps * Ci + (int64_t)t * Ci + c] = from_f32<__half>(w[i]);
  }
}
// split-precision variant: [Cout][hi(9*Cin) | lo(9*Cin)] of w * up (up = a power of two)
__global__ void pack_w_dcn_tc_split_kernel(const float* __restrict__ w, __half* __restrict__ out, int Co, int Ci, int taps,
                                           float up) {
  const int64_t total = (int64_t)Co * Ci * taps;
  const int64_t K = (int64_t)taps * Ci;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % taps);
    int64_t r = i / taps;
    int c = (int)(r % Ci);
    int o = (int)(r / Ci);
    __half hi, lo;
    split_f32(w[i] * up, hi, lo);   // same pair format as the activations
    out[(int64_t)o * 2 * K + (int64_t)t * Ci + c] = hi;
    out[(int64_t)o * 2 * K + K + (int64_t)t * Ci + c] = lo;
  }
}
// offset [B,18,HW] + mask [B,9,HW] (NCHW fp32) -> om [B,HW,27]
__global__ void pack_om_kernel(const float* __restrict__ off, const float* __restrict__ msk, float* __restrict__ om,
                               int B, int HW) {
  const int64_t total = (int64_t)B * HW * 27;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % 27);
    int64_t r = i / 27;
    int p = (int)(r % HW);
    int b = (int)(r / HW);
    om[i] = ch < 18 ? off[((int64_t)b * 18 + ch) * HW + p] : msk[((int64_t)b * 9 + (ch - 18)) * HW + p];
  }
}

// Every entry point that takes a handle holds its mutex while it enqueues work (the handle's plans, launch counter
// and lazily grown workspaces are not thread-safe by themselves).  Entry points that run on the handle's SHARED device
// buffers (executor activations, Detect workspace, scratch) additionally order themselves on the device behind the
// previous such call when it was issued on a different stream -- eval.py's ThreadPool callers (eval.py:793-827) may
// drive one net from several threads / streams; results are then serialised, never corrupted.
struct CallGuard {
  yb_handle* h;
  DeviceGuard dg;
  std::unique_lock<std::recursive_mutex> lk;
  cudaStream_t stream = nullptr;
  bool chain = false;
  explicit CallGuard(yb_handle* h_) : h(h_), dg(h_->device), lk(h_->mu) {}
  static bool capturing(cudaStream_t s) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) {
      cudaGetLastError();
      return true;
    }
    return st != cudaStreamCaptureStatusNone;
  }
  CallGuard(yb_handle* h_, cudaStream_t s) : h(h_), dg(h_->device), lk(h_->mu), stream(s), chain(true) {
    // (a stream the CALLER is capturing neither waits on nor records the handle's event: ordering is the caller's graph)
    if (h->ev_last && h->has_last && h->last_stream != s && !capturing(s)) cudaStreamWaitEvent(s, h->ev_last, 0);
  }
  ~CallGuard() {
    if (!chain || capturing(stream)) return;
    if (!h->ev_last && cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      h->ev_last = nullptr;
      return;
    }
    if (cudaEventRecord(h->ev_last, stream) == cudaSuccess) {
      h->last_stream = stream;
      h->has_last = true;
    } else {
      cudaGetLastError();
    }
  }
};

struct TempPool {  // RAII device temporaries for the op-level hooks
  std::vector<void*> v;
  void* get(size_t bytes) {
    void* p = nullptr;
    YB_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 256)));
    v.push_back(p);
    return p;
  }
  ~TempPool() {
    for (void* p : v) cudaFree(p);
  }
};

inline int grid1d(int64_t n) { return (int)std::min<int64_t>(148 * 16, (n + 255) / 256); }

}  // namespace

extern "C" {

int yb_abi_version(void) { return YB_ABI_VERSION; }
const char* yb_last_error(void) { return yb::g_last_error.c_str(); }

int yb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    yb::set_last_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return YB_ERR_NO_DEVICE;
  }
  return n;
}

int yb_create(const yb_config* cfg, int device, yb_handle** out) {
  YB_API_BEGIN
  YB_REQUIRE(cfg && out, "yb_create: null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    throw Error(YB_ERR_NO_DEVICE, "yb_create: no CUDA device is visible (this library has no CPU fallback)");
  }
  YB_REQUIRE(device >= 0 && device < n, "yb_create: bad device index");
  cudaDeviceProp prop;
  YB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  YB_REQUIRE(prop.major == 10, "yb_create: this library is built for sm_100a (Blackwell B200) only");
  std::unique_ptr<yb_handle> h(new yb_handle());
  h->cfg = *cfg;
  h->device = device;
  h->ops_only = (cfg->backbone == YB_BACKBONE_NONE);
  // measured defaults (profiles/r2_*): programmatic dependent launch gains 3.6 % of the conv stack in the single-pass
  // fp16 mode and loses 1.5 % in the split mode (whose plans fill the SM, nothing can become resident early); stream-K
  // gains 1.6 % in the split mode and nothing in the fp16 mode
  h->pdl = (cfg->precision == YB_PREC_F16TC);
  h->sk_candidates = (cfg->precision == YB_PREC_F16X3);
  if (const char* at = getenv("YB_AUTOTUNE")) h->autotune = (atoi(at) != 0);
  if (const char* pc = getenv("YB_PAIR")) h->pair_candidates = (atoi(pc) != 0);
  if (const char* ec = getenv("YB_EPI2")) h->epi2_candidates = (atoi(ec) != 0);
  if (const char* sk = getenv("YB_SK")) h->sk_candidates = (atoi(sk) != 0);
  if (const char* ch = getenv("YB_CHAIN")) h->chain_mode = std::min(2, std::max(0, atoi(ch)));
  if (const char* sw = getenv("YB_STEM_WG")) h->stem_wg = (atoi(sw) == 2 || atoi(sw) == 4) ? atoi(sw) : 1;   // default 0: per precision mode
  if (const char* st = getenv("YB_STEM_TC")) h->stem_on_tc = (atoi(st) != 0);
  if (const char* df = getenv("YB_DCN_FUSED")) h->dcn_fused = (atoi(df) != 0);
  if (const char* pd = getenv("YB_PDL")) h->pdl = (atoi(pd) != 0);
  if (const char* fh = getenv("YB_FUSE_HEADS")) h->fuse_heads = (atoi(fh) != 0);
  if (const char* br = getenv("YB_BRANCHES")) h->multi_stream = (atoi(br) != 0);
  if (!h->ops_only) {
    YB_REQUIRE(cfg->backbone == YB_BACKBONE_RESNET || cfg->backbone == YB_BACKBONE_DARKNET, "unknown backbone");
    YB_REQUIRE(cfg->num_stages >= 4 && cfg->num_stages <= 5, "num_stages must be 4 or 5");
    YB_REQUIRE(cfg->fpn_features == 256 || cfg->fpn_features % 64 == 0, "fpn_features must be a multiple of 64");
    YB_REQUIRE(cfg->num_scales >= 1 && cfg->num_scales <= 4 && cfg->num_ars >= 1 && cfg->num_ars <= 4, "bad anchors");
    for (int i = 0; i < 3; ++i)
      YB_REQUIRE(cfg->selected_layers[i] >= 0 && cfg->selected_layers[i] < cfg->num_stages, "bad selected_layers");
  }
  YB_REQUIRE(cfg->precision == YB_PREC_F32 || cfg->precision == YB_PREC_F16TC || cfg->precision == YB_PREC_F16X3,
             "unknown precision");
  YB_REQUIRE(cfg->mask_dim % 4 == 0 && cfg->mask_dim > 0, "mask_dim must be a positive multiple of 4");
  DeviceGuard g(device);
  YB_CHECK_CUDA(cudaFree(0));
  *out = h.release();
  YB_API_END
}

int yb_destroy(yb_handle* h) {
  YB_API_BEGIN
  if (h) {
    DeviceGuard dg(h->device);
    cudaDeviceSynchronize();
    delete h;
  }
  YB_API_END
}

int yb_load_weight(yb_handle* h, const char* name, const float* h_data, const int64_t* shape, int ndim) {
  YB_API_BEGIN
  YB_REQUIRE(h && name && h_data && (shape || ndim == 0), "yb_load_weight: null argument");
  YB_REQUIRE(ndim >= 0 && ndim <= 8, "yb_load_weight: bad ndim");
  HostTensor t;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    YB_REQUIRE(shape[i] >= 0, "yb_load_weight: negative dim");
    t.shape.push_back(shape[i]);
    n *= shape[i];
  }
  t.data.assign(h_data, h_data + n);
  h->host[name] = std::move(t);
  h->finalized = false;
  YB_API_END
}

int yb_finalize_weights(yb_handle* h) {
  YB_API_BEGIN
  YB_REQUIRE(h, "null handle");
  CallGuard g(h);
  h->finalize();
  YB_API_END
}

int yb_num_priors(yb_handle* h, int img_h, int img_w, int64_t* num_priors, int32_t* level_hw) {
  YB_API_BEGIN
  YB_REQUIRE(h && !h->ops_only, "yb_num_priors: handle has no network");
  int lhw[5][2];
  compute_level_sizes(h->cfg, img_h, img_w, lhw, nullptr, nullptr);
  int64_t P = 0;
  for (int l = 0; l < 5; ++l) {
    P += (int64_t)lhw[l][0] * lhw[l][1] * h->cfg.num_scales * h->cfg.num_ars;
    if (level_hw) {
      level_hw[2 * l] = lhw[l][0];
      level_hw[2 * l + 1] = lhw[l][1];
    }
  }
  if (num_priors) *num_priors = P;
  YB_API_END
}

int yb_priors(yb_handle* h, int img_h, int img_w, float* d_priors, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && !h->ops_only && d_priors, "yb_priors: bad argument");
  CallGuard g(h);
  int lhw[5][2];
  compute_level_sizes(h->cfg, img_h, img_w, lhw, nullptr, nullptr);
  std::vector<float> pri = make_priors_host(h->cfg, lhw);
  YB_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  YB_CHECK_CUDA(cudaMemcpy(d_priors, pri.data(), pri.size() * 4, cudaMemcpyHostToDevice));
  YB_API_END
}

int yb_proto_size(yb_handle* h, int img_h, int img_w, int32_t* ph, int32_t* pw) {
  YB_API_BEGIN
  YB_REQUIRE(h && !h->ops_only, "yb_proto_size: handle has no network");
  int lhw[5][2], a, b;
  compute_level_sizes(h->cfg, img_h, img_w, lhw, &a, &b);
  if (ph) *ph = a;
  if (pw) *pw = b;
  YB_API_END
}

int yb_forward(yb_handle* h, const float* d_x, int B, int H, int W, float* d_loc, float* d_conf, float* d_coef,
               float* d_proto, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_x && B > 0 && H > 0 && W > 0, "yb_forward: bad argument");
  YB_REQUIRE(!h->ops_only, "yb_forward: handle has no network");
  CallGuard g(h, (cudaStream_t)stream);   // shared workspaces: ordered behind the previous call
  h->forward(d_x, B, H, W, d_loc, d_conf, d_coef, d_proto, (cudaStream_t)stream);
  YB_API_END
}

int yb_infer(yb_handle* h, const float* d_x, int B, int H, int W, int cross_class, int max_out, float* d_box,
             float* d_coef_out, int64_t* d_cls, float* d_score, int32_t* d_count, float* d_proto, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_x && B > 0 && H > 0 && W > 0, "yb_infer: bad argument");
  YB_REQUIRE(!h->ops_only, "yb_infer: handle has no network");
  YB_REQUIRE(d_box && d_coef_out && d_cls && d_score && d_count, "yb_infer: null output");
  CallGuard g(h, (cudaStream_t)stream);   // shared workspaces: ordered behind the previous call
  h->infer(d_x, B, H, W, cross_class, max_out, d_box, d_coef_out, d_cls, d_score, d_count, d_proto,
           (cudaStream_t)stream);
  YB_API_END
}

int yb_debug_feature(yb_handle* h, int which, float* d_out, int32_t* chw, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && h->last_exec && which >= 0 && which < 9, "yb_debug_feature: no forward has run / bad index");
  CallGuard g(h, (cudaStream_t)stream);   // shared workspaces: ordered behind the previous call
  const Act& a = h->last_exec->feats[which];
  YB_REQUIRE(a.ptr != nullptr, "yb_debug_feature: feature not available for this backbone");
  if (chw) {
    chw[0] = a.C;
    chw[1] = a.H;
    chw[2] = a.W;
  }
  if (d_out) {
    if (a.f32)
      launch_nhwc_to_nchw_f32<float>((const float*)a.ptr, d_out, a.B, a.H, a.W, a.C, (cudaStream_t)stream, &h->lc);
    else
      launch_nhwc_to_nchw_f32<__half>((const __half*)a.ptr, d_out, a.B, a.H, a.W, a.C, (cudaStream_t)stream, &h->lc,
                                      a.split ? 1 : 0);
  }
  YB_API_END
}

int yb_softmax(yb_handle* h, const float* d_in, float* d_out, int64_t rows, int cols, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_in && d_out && rows >= 0 && cols > 0, "yb_softmax: bad argument");
  CallGuard g(h);
  launch_softmax_rows(d_in, d_out, rows, cols, (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_set_detect_params(yb_handle* h, int top_k, float conf_thresh, float nms_thresh, int max_num_detections) {
  YB_API_BEGIN
  YB_REQUIRE(h, "yb_set_detect_params: null handle");
  YB_REQUIRE(top_k >= 1 && top_k <= 256, "yb_set_detect_params: top_k must be in [1, 256] (one CTA sorts a class)");
  YB_REQUIRE(max_num_detections >= 1 && max_num_detections <= 256, "yb_set_detect_params: max_num_detections must be in [1, 256]");
  YB_REQUIRE(nms_thresh > 0.f, "nms_threshold must be non negative.");   // detection.py:25-26
  CallGuard g(h);
  yb_config& c = h->cfg;
  if (c.nms_top_k == top_k && c.nms_conf_thresh == conf_thresh && c.nms_thresh == nms_thresh &&
      c.max_num_detections == max_num_detections)
    return YB_OK;
  YB_CHECK_CUDA(cudaDeviceSynchronize());
  c.nms_top_k = top_k;
  c.nms_conf_thresh = conf_thresh;
  c.nms_thresh = nms_thresh;
  c.max_num_detections = max_num_detections;
  for (auto& kv : h->execs) kv.second->drop_detect_state();   // workspace sizes and captured graphs depend on them
  YB_API_END
}

int yb_detect(yb_handle* h, const float* d_loc, const float* d_conf, const float* d_coef, const float* d_priors,
              int B, int64_t P, int conf_is_logits, int cross_class, int max_out, float* d_box, float* d_coef_out,
              int64_t* d_cls, float* d_score, int32_t* d_count, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_loc && d_conf && d_coef && d_priors && d_box && d_coef_out && d_cls && d_score && d_count,
             "yb_detect: null argument");
  YB_REQUIRE(B > 0 && P > 0, "yb_detect: empty input");
  CallGuard g(h, (cudaStream_t)stream);   // shared workspaces: ordered behind the previous call
  DetectParams dp;
  dp.B = B;
  dp.P = P;
  dp.num_classes = h->cfg.num_classes;
  dp.mask_dim = h->cfg.mask_dim;
  dp.top_k = h->cfg.nms_top_k;
  dp.conf_thresh = h->cfg.nms_conf_thresh;
  dp.nms_thresh = h->cfg.nms_thresh;
  dp.max_dets = h->cfg.max_num_detections;
  dp.conf_is_logits = conf_is_logits;
  dp.cross_class = cross_class & 0xFF;
  dp.second_threshold = (cross_class & YB_NMS_FLAG_SECOND_THRESHOLD) ? 1 : 0;
  dp.max_size = (float)h->cfg.max_size;
  dp.max_out = max_out;
  YB_REQUIRE(dp.nms_thresh > 0.f, "nms_threshold must be non negative.");  // detection.py:25-26
  void* ws = h->get_detect_ws(detect_workspace_bytes(B, P, dp.num_classes, dp.top_k));
  DetectWorkspace dws;
  detect_workspace_bind(&dws, ws, B, P, dp.num_classes, dp.top_k);
  launch_detect(dp, d_loc, d_conf, d_coef, d_priors, dws, d_box, d_coef_out, d_cls, d_score, d_count,
                (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_postprocess(yb_handle* h, const float* d_proto, int ph, int pw, int k, const float* d_coef, const float* d_box,
                   int n, int out_h, int out_w, int crop_masks, int mask_format, void* d_masks, int64_t* d_boxes_px,
                   float* d_proto_masks, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_proto && d_coef && d_box, "yb_postprocess: null argument");
  YB_REQUIRE(n >= 0, "yb_postprocess: negative detection count");
  CallGuard g(h);
  launch_mask_assembly(d_proto, ph, pw, k, d_coef, d_box, n, out_h, out_w, crop_masks, mask_format, d_masks,
                       d_boxes_px, d_proto_masks, (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_postprocess_batch(yb_handle* h, const float* d_proto, int ph, int pw, int k, const float* d_coef,
                         const float* d_box, int n, int batch, int out_h, int out_w, int crop_masks, int mask_format,
                         void* d_masks, int64_t* d_boxes_px, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_proto && d_coef && d_box, "yb_postprocess_batch: null argument");
  YB_REQUIRE(n >= 0 && batch >= 0, "yb_postprocess_batch: negative count");
  CallGuard g(h);
  launch_mask_assembly(d_proto, ph, pw, k, d_coef, d_box, n, out_h, out_w, crop_masks, mask_format, d_masks,
                       d_boxes_px, nullptr, (cudaStream_t)stream, &h->lc, batch);
  YB_API_END
}

int yb_maskiou(yb_handle* h, const float* d_proto_masks, int n, int ph, int pw, const int64_t* d_cls,
               float* d_maskiou, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_proto_masks && d_maskiou, "yb_maskiou: null argument");
  YB_REQUIRE(h->cfg.use_maskiou && h->finalized, "yb_maskiou: network has no maskiou_net / weights not finalized");
  if (n <= 0) return YB_OK;
  CallGuard g(h, (cudaStream_t)stream);   // shared workspaces: ordered behind the previous call
  cudaStream_t s = (cudaStream_t)stream;
  // FastMaskIoUNet (yolact.py:363-375, config.py:785-789): 5x (3x3 s2 p0 + ReLU), 1x1 + ReLU, global max
  const char* idx[6] = {"0", "2", "4", "6", "8", "10"};
  int H = ph, W = pw, C = 1;
  size_t need = 0;
  {
    int hh = ph, ww = pw;
    for (int i = 0; i < 6; ++i) {
      ConvW& cw = h->convs[std::string("maskiou_net.maskiou_net.") + idx[i]];
      int k = cw.KH, st = (i < 5) ? 2 : 1;
      hh = (hh - k) / st + 1;
      ww = (ww - k) / st + 1;
      need = std::max(need, (size_t)n * hh * ww * cw.Cout * 4);
    }
  }
  const size_t half = (need + 255) / 256 * 256;
  char* scratch = (char*)h->get_scratch(2 * half);
  const float* cur = d_proto_masks;
  for (int i = 0; i < 6; ++i) {
    ConvW& cw = h->convs[std::string("maskiou_net.maskiou_net.") + idx[i]];
    YB_REQUIRE(cw.w_f32 && cw.Cin == C, "yb_maskiou: weights missing");
    ConvProblem p;
    p.B = n;
    p.H = H;
    p.W = W;
    p.Cin = C;
    p.KH = cw.KH;
    p.KW = cw.KW;
    p.stride = (i < 5) ? 2 : 1;
    p.pad = 0;
    p.Ho = (H - cw.KH) / p.stride + 1;
    p.Wo = (W - cw.KW) / p.stride + 1;
    YB_REQUIRE(p.Ho >= 1 && p.Wo >= 1, "yb_maskiou: mask too small for the network");
    p.Cout = cw.Cout;
    p.act = ACT_RELU;
    p.x = cur;
    float* out = (float*)(scratch + (i & 1) * half);
    p.y = out;
    p.y_f32 = 1;
    p.y_batch_stride = (int64_t)p.Ho * p.Wo * cw.Cout;
    p.y_pix_stride = cw.Cout;
    p.bias = cw.bias;
    launch_simt_conv(p, cw.w_f32, SIMT_F32, s, &h->lc);
    cur = out;
    H = p.Ho;
    W = p.Wo;
    C = cw.Cout;
  }
  launch_maxpool_gather(cur, n, H, W, C, d_cls, d_maskiou, s, &h->lc);
  YB_API_END
}

// ---- frame preparation / eval.py consumers -----------------------------------------------------------
int yb_fast_base_transform(yb_handle* h, const void* d_img, int img_is_u8, int B, int H, int W, int out_h, int out_w,
                           int mode, const float* h_mean_bgr, const float* h_std_bgr, float* d_out, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_img && d_out, "yb_fast_base_transform: null argument");
  YB_REQUIRE(mode >= YB_XFORM_NORMALIZE && mode <= YB_XFORM_NONE, "yb_fast_base_transform: unknown transform mode");
  // data/config.py:28-29 (BGR order)
  static const float kMeans[3] = {103.94f, 116.78f, 123.68f};
  static const float kStd[3] = {57.38f, 57.12f, 58.40f};
  CallGuard g(h);
  launch_fast_base_transform(d_img, img_is_u8, B, H, W, out_h, out_w, mode, h_mean_bgr ? h_mean_bgr : kMeans,
                             h_std_bgr ? h_std_bgr : kStd, d_out, (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_pack_mask_bits(yb_handle* h, const void* d_in, int in_format, int64_t rows, int w, uint32_t* d_bits,
                      void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && rows >= 0 && w > 0 && (rows == 0 || (d_in && d_bits)), "yb_pack_mask_bits: bad argument");
  CallGuard g(h);
  launch_pack_mask_bits(d_in, in_format, rows, w, d_bits, (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_mask_iou(yb_handle* h, const uint32_t* d_a, int n, const uint32_t* d_b, int m, int64_t words, int iscrowd,
                float* d_iou, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && n >= 0 && m >= 0 && words >= 0, "yb_mask_iou: bad argument");
  YB_REQUIRE(n == 0 || m == 0 || (d_a && d_b && d_iou), "yb_mask_iou: null argument");
  CallGuard g(h);
  launch_mask_iou_bits(d_a, n, d_b, m, words, iscrowd, d_iou, (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_box_iou(yb_handle* h, const float* d_a, int n, const float* d_b, int m, int iscrowd, float* d_iou,
               void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && n >= 0 && m >= 0, "yb_box_iou: bad argument");
  YB_REQUIRE(n == 0 || m == 0 || (d_a && d_b && d_iou), "yb_box_iou: null argument");
  CallGuard g(h);
  launch_box_iou(d_a, n, d_b, m, iscrowd, d_iou, (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_mask_rle(yb_handle* h, const void* d_masks, int mask_format, int n, int mask_h, int mask_w, uint32_t* d_counts,
                int64_t cap, int32_t* d_nruns, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && n >= 0, "yb_mask_rle: bad argument");
  YB_REQUIRE(n == 0 || (d_masks && d_counts && d_nruns), "yb_mask_rle: null argument");
  CallGuard g(h);
  launch_mask_rle(d_masks, mask_format, n, mask_h, mask_w, d_counts, cap, d_nruns, (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_pack_detections(yb_handle* h, const float* d_box, const float* d_coef, const int64_t* d_cls, const float* d_score,
                       const int32_t* d_count, int B, int M, int k, float* d_rec, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_box && d_coef && d_cls && d_score && d_count && d_rec, "yb_pack_detections: null argument");
  YB_REQUIRE(B >= 0 && M >= 1 && k >= 1, "yb_pack_detections: bad sizes");
  CallGuard g(h);
  launch_pack_detections(d_box, d_coef, d_cls, d_score, d_count, B, M, k, d_rec, (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_display_blend(yb_handle* h, const float* d_img, int img_is_255, const void* d_masks, int mask_format, int n,
                     int img_h, int img_w, const float* d_colors, float alpha, uint8_t* d_out, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_img && d_out, "yb_display_blend: null argument");
  YB_REQUIRE(n == 0 || (d_masks && d_colors), "yb_display_blend: null masks / colors");
  CallGuard g(h);
  launch_display_blend(d_img, img_is_255, d_masks, mask_format, n, img_h, img_w, d_colors, alpha, d_out,
                       (cudaStream_t)stream, &h->lc);
  YB_API_END
}

int yb_dcn_forward(yb_handle* h, const float* d_input, const float* d_weight, const float* d_bias,
                   const float* d_offset, const float* d_mask, float* d_output, int B, int C, int H, int W, int Co,
                   int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w, int dilation_h,
                   int dilation_w, int deformable_group, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_input && d_weight && d_offset && d_mask && d_output, "yb_dcn_forward: null argument");
  YB_REQUIRE(kernel_h == 3 && kernel_w == 3, "yb_dcn_forward: only 3x3 kernels (all YOLACT++ DCN layers)");
  YB_REQUIRE(stride_h == stride_w && pad_h == pad_w && dilation_h == dilation_w, "yb_dcn_forward: anisotropic params");
  YB_REQUIRE(deformable_group == 1, "yb_dcn_forward: deformable_group must be 1");
  YB_REQUIRE(C % 16 == 0, "yb_dcn_forward: C must be a multiple of 16");
  CallGuard g(h);
  cudaStream_t s = (cudaStream_t)stream;
  const int Ho = (H + 2 * pad_h - (dilation_h * 2 + 1)) / stride_h + 1;
  const int Wo = (W + 2 * pad_w - (dilation_w * 2 + 1)) / stride_w + 1;
  TempPool tp;
  float* om = (float*)tp.get((size_t)B * Ho * Wo * 27 * 4);
  pack_om_kernel<<<grid1d((int64_t)B * Ho * Wo * 27), 256, 0, s>>>(d_offset, d_mask, om, B, Ho * Wo);
  YB_CHECK_LAUNCH();
  const int64_t wn = (int64_t)Co * C * 9;
  const bool f16 = (h->cfg.precision != YB_PREC_F32);
  const int sp = (h->cfg.precision == YB_PREC_F16X3) ? 1 : 0;
  YB_REQUIRE(!sp || C % 64 == 0, "yb_dcn_forward: the split-precision mode needs C % 64 == 0");
  const int npl = sp ? 2 : 1;
  if (!f16) {
    float* x = (float*)tp.get((size_t)B * H * W * C * 4);
    float* y = (float*)tp.get((size_t)B * Ho * Wo * Co * 4);
    float* wk = (float*)tp.get((size_t)wn * 4);
    launch_nchw_f32_to_nhwc<float>(d_input, x, B, C, H, W, s, &h->lc);
    pack_w_simt_kernel<float><<<grid1d(wn), 256, 0, s>>>(d_weight, wk, Co, C, 9);
    YB_CHECK_LAUNCH();
    launch_dcn_simt<float>(x, om, wk, d_bias, y, B, H, W, C, Ho, Wo, Co, stride_h, pad_h, dilation_h, ACT_NONE, 0, s,
                           &h->lc);
    launch_nhwc_to_nchw_f32<float>(y, d_output, B, Ho, Wo, Co, s, &h->lc);
  } else {
    __half* x = (__half*)tp.get((size_t)B * H * W * C * 2 * npl);
    __half* y = (__half*)tp.get((size_t)B * Ho * Wo * Co * 2 * npl);
    launch_nchw_f32_to_nhwc<__half>(d_input, x, B, C, H, W, s, &h->lc, sp);
    if (C % 64 == 0) {
      __half* cols = (__half*)tp.get((size_t)B * Ho * Wo * 9 * C * 2 * npl);
      __half* wk = (__half*)tp.get((size_t)wn * 2 * npl);
      float out_scale = 1.f;
      if (sp) {
        // one power of two for the whole weight tensor: max |w| is read back (op-level hook, not the hot path)
        std::vector<float> hw((size_t)wn);
        YB_CHECK_CUDA(cudaMemcpyAsync(hw.data(), d_weight, (size_t)wn * 4, cudaMemcpyDeviceToHost, s));
        YB_CHECK_CUDA(cudaStreamSynchronize(s));
        float mx = 0.f;
        for (float v : hw) mx = std::max(mx, fabsf(v));
        int ex = 0;
        if (mx > 0.f) frexpf(mx, &ex);
        const int e = mx > 0.f ? std::max(-24, std::min(40, 14 - ex)) : 0;
        out_scale = ldexpf(1.f, -e);
        pack_w_dcn_tc_split_kernel<<<grid1d(wn), 256, 0, s>>>(d_weight, wk, Co, C, 9, ldexpf(1.f, e));
      } else {
        pack_w_dcn_tc_kernel<<<grid1d(wn), 256, 0, s>>>(d_weight, wk, Co, C, 9);
      }
      YB_CHECK_LAUNCH();
      if (h->dcn_fused && dcn_tc_supported(C, Co)) {
        DcnTcPlan* dp = dcn_tc_plan_create(x, om, wk, d_bias, y, B, H, W, C, Ho, Wo, Co, stride_h, pad_h, dilation_h,
                                           ACT_NONE, 0, sp, out_scale);
        try {
          launch_dcn_tc(dp, s, &h->lc);
        } catch (...) {
          dcn_tc_plan_destroy(dp);
          throw;
        }
        dcn_tc_plan_destroy(dp);
        launch_nhwc_to_nchw_f32<__half>(y, d_output, B, Ho, Wo, Co, s, &h->lc, sp);
        YB_CHECK_CUDA(cudaStreamSynchronize(s));
        return YB_OK;
      }
      launch_dcn_gather_f16(x, om, cols, B, H, W, C, Ho, Wo, stride_h, pad_h, dilation_h, 0, s, &h->lc, sp);
      ConvProblem p;
      p.B = B;
      p.H = Ho;
      p.W = Wo;
      p.Cin = 9 * C;
      p.Ho = Ho;
      p.Wo = Wo;
      p.Cout = Co;
      p.x = cols;
      p.y = y;
      p.split = sp;
      p.out_scale = out_scale;
      p.y_pix_stride = Co * npl;
      p.y_batch_stride = (int64_t)Ho * Wo * p.y_pix_stride;
      p.bias = d_bias;
      TcConvPlan* plan = tc_conv_plan_create(p, wk);
      try {
        launch_tc_conv(plan, s, &h->lc);
      } catch (...) {
        tc_conv_plan_destroy(plan);
        throw;
      }
      tc_conv_plan_destroy(plan);
    } else {
      __half* wk = (__half*)tp.get((size_t)wn * 2);
      pack_w_simt_kernel<__half><<<grid1d(wn), 256, 0, s>>>(d_weight, wk, Co, C, 9);
      YB_CHECK_LAUNCH();
      launch_dcn_simt<__half>(x, om, wk, d_bias, y, B, H, W, C, Ho, Wo, Co, stride_h, pad_h, dilation_h, ACT_NONE, 0,
                              s, &h->lc);
    }
    launch_nhwc_to_nchw_f32<__half>(y, d_output, B, Ho, Wo, Co, s, &h->lc, sp);
  }
  YB_CHECK_CUDA(cudaStreamSynchronize(s));  // temporaries are freed on return
  YB_API_END
}

int yb_conv2d(yb_handle* h, const float* d_x, const float* h_w, const float* h_bias, const float* d_residual,
              float* d_y, int B, int Ci, int H, int W, int Co, int kh, int kw, int stride, int pad, int act,
              int precision, int iters, float* ms, void* stream) {
  YB_API_BEGIN
  YB_REQUIRE(h && d_x && h_w && d_y, "yb_conv2d: null argument");
  YB_REQUIRE(precision >= 0 && precision <= 3,
             "yb_conv2d: precision must be 0 (f32 simt), 1 (f16 tcgen05), 2 (f16 simt), 3 (split-precision tcgen05)");
  const int sp = (precision == 3) ? 1 : 0;
  CallGuard g(h);
  cudaStream_t s = (cudaStream_t)stream;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  YB_REQUIRE(Ho >= 1 && Wo >= 1, "yb_conv2d: empty output");
  TempPool tp;
  const int taps = kh * kw;
  const size_t K = (size_t)taps * Ci;
  const bool f16 = precision != 0;
  const size_t es = (f16 && !sp) ? 2 : 4;   // split: two halfs per element
  void* x = tp.get((size_t)B * H * W * Ci * es);
  void* y = tp.get((size_t)B * Ho * Wo * Co * es);
  void* res = d_residual ? tp.get((size_t)B * Ho * Wo * Co * es) : nullptr;
  float* bias = nullptr;
  if (h_bias) {
    bias = (float*)tp.get((size_t)Co * 4);
    YB_CHECK_CUDA(cudaMemcpy(bias, h_bias, (size_t)Co * 4, cudaMemcpyHostToDevice));
  }
  ConvProblem p;
  p.B = B;
  p.H = H;
  p.W = W;
  p.Cin = Ci;
  p.Ho = Ho;
  p.Wo = Wo;
  p.Cout = Co;
  p.KH = kh;
  p.KW = kw;
  p.stride = stride;
  p.pad = pad;
  p.act = act;
  p.x = x;
  p.y = y;
  p.split = sp;
  p.y_pix_stride = sp ? 2 * Co : Co;
  p.y_batch_stride = (int64_t)Ho * Wo * p.y_pix_stride;
  p.bias = bias;
  p.residual = res;
  if (!f16) {
    launch_nchw_f32_to_nhwc<float>(d_x, (float*)x, B, Ci, H, W, s, &h->lc);
    if (res) launch_nchw_f32_to_nhwc<float>(d_residual, (float*)res, B, Co, Ho, Wo, s, &h->lc);
  } else {
    launch_nchw_f32_to_nhwc<__half>(d_x, (__half*)x, B, Ci, H, W, s, &h->lc, sp);
    if (res) launch_nchw_f32_to_nhwc<__half>(d_residual, (__half*)res, B, Co, Ho, Wo, s, &h->lc, sp);
  }
  std::function<void()> run;
  TcConvPlan* plan = nullptr;
  if (precision == 1 || precision == 3) {
    YB_REQUIRE(tc_conv_supported(p), "yb_conv2d: shape not supported by the tcgen05 kernel (Cin % 64, taps <= 9)");
    std::vector<__half> pk(K * Co * (sp ? 2 : 1));
    if (sp) {
      float mx = 0.f;
      for (size_t i = 0; i < K * Co; ++i) mx = std::max(mx, fabsf(h_w[i]));
      int ex = 0;
      if (mx > 0.f) frexpf(mx, &ex);
      const int e = mx > 0.f ? std::max(-24, std::min(40, 14 - ex)) : 0;
      const float up = ldexpf(1.f, e);
      p.out_scale = ldexpf(1.f, -e);
      for (int o = 0; o < Co; ++o)
        for (int c = 0; c < Ci; ++c)
          for (int t = 0; t < taps; ++t) {
            const float vs = h_w[((size_t)o * Ci + c) * taps + t] * up;
            const __half hi = __float2half_rn(vs);
            const size_t idx = ((size_t)t * Co + o) * 2 * Ci + c;
            pk[idx] = hi;
            pk[idx + Ci] = __float2half_rn((vs - __half2float(hi)) * 2048.f);   // lo' = residual * 2^11 (common.cuh)
          }
    } else
    for (int o = 0; o < Co; ++o)
      for (int c = 0; c < Ci; ++c)
        for (int t = 0; t < taps; ++t)
          pk[((size_t)t * Co + o) * Ci + c] = __float2half_rn(h_w[((size_t)o * Ci + c) * taps + t]);
    __half* wd = (__half*)tp.get(pk.size() * 2);
    YB_CHECK_CUDA(cudaMemcpy(wd, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    {
      const char* be = getenv("YB_CONV2D_BN");
      const char* ge = getenv("YB_CONV2D_GRID");
      const char* pe = getenv("YB_CONV2D_PAIR");
      const char* ee = getenv("YB_CONV2D_EPI");
      const char* de = getenv("YB_CONV2D_PDL");   // PDL-friendly plan + programmatic dependent launch
      const char* ke = getenv("YB_CONV2D_SK");    // stream-K
      plan = tc_conv_plan_create(p, wd, be ? atoi(be) : 0, 0, ge ? atoi(ge) : 0, pe ? atoi(pe) : 0, ee ? atoi(ee) : 0,
                                 de ? atoi(de) : 0, ke ? atoi(ke) : 0);
      if (de && atoi(de)) tc_conv_plan_set_pdl(plan, 1);
      if (tc_conv_plan_sk(plan)) {
        void* ws = tp.get(tc_conv_sk_workspace_bytes());
        YB_CHECK_CUDA(cudaMemsetAsync(ws, 0, tc_conv_sk_workspace_bytes(), s));
        tc_conv_plan_set_sk_workspace(plan, ws);
      }
    }
    run = [&]() { launch_tc_conv(plan, s, &h->lc); };
  } else if (precision == 2) {
    std::vector<__half> pk(K * Co);
    for (int o = 0; o < Co; ++o)
      for (int c = 0; c < Ci; ++c)
        for (int t = 0; t < taps; ++t)
          pk[((size_t)t * Ci + c) * Co + o] = __float2half_rn(h_w[((size_t)o * Ci + c) * taps + t]);
    __half* wd = (__half*)tp.get(pk.size() * 2);
    YB_CHECK_CUDA(cudaMemcpy(wd, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    run = [&, wd]() { launch_simt_conv(p, wd, SIMT_F16, s, &h->lc); };
  } else {
    std::vector<float> pk(K * Co);
    for (int o = 0; o < Co; ++o)
      for (int c = 0; c < Ci; ++c)
        for (int t = 0; t < taps; ++t) pk[((size_t)t * Ci + c) * Co + o] = h_w[((size_t)o * Ci + c) * taps + t];
    float* wd = (float*)tp.get(pk.size() * 4);
    YB_CHECK_CUDA(cudaMemcpy(wd, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice));
    run = [&, wd]() { launch_simt_conv(p, wd, SIMT_F32, s, &h->lc); };
  }
  try {
    run();  // warm-up + result
    if (iters > 1 || ms) {
      const int n = std::max(1, iters);
      cudaEvent_t e0, e1;
      YB_CHECK_CUDA(cudaEventCreate(&e0));
      YB_CHECK_CUDA(cudaEventCreate(&e1));
      YB_CHECK_CUDA(cudaEventRecord(e0, s));
      for (int i = 0; i < n; ++i) run();
      YB_CHECK_CUDA(cudaEventRecord(e1, s));
      YB_CHECK_CUDA(cudaEventSynchronize(e1));
      float t = 0.f;
      YB_CHECK_CUDA(cudaEventElapsedTime(&t, e0, e1));
      if (ms) *ms = t / n;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
  } catch (...) {
    if (plan) tc_conv_plan_destroy(plan);
    throw;
  }
  if (plan) tc_conv_plan_destroy(plan);
  if (!f16)
    launch_nhwc_to_nchw_f32<float>((const float*)y, d_y, B, Ho, Wo, Co, s, &h->lc);
  else
    launch_nhwc_to_nchw_f32<__half>((const __half*)y, d_y, B, Ho, Wo, Co, s, &h->lc, sp);
  YB_CHECK_CUDA(cudaStreamSynchronize(s));
  YB_API_END
}

int64_t yb_launch_count(yb_handle* h) { return h ? h->lc.n : 0; }

int yb_set_profiling(yb_handle* h, int enable) {
  YB_API_BEGIN
  YB_REQUIRE(h, "null handle");
  h->profiling = enable != 0;
  YB_API_END
}

int yb_last_forward_ms(yb_handle* h, float* total_ms, float* conv_ms) {
  YB_API_BEGIN
  YB_REQUIRE(h, "null handle");
  if (total_ms) *total_ms = h->last_total_ms;
  if (conv_ms) *conv_ms = h->last_conv_ms;
  YB_API_END
}

int yb_last_forward_profile(yb_handle* h, char* buf, int64_t cap) {
  YB_API_BEGIN
  YB_REQUIRE(h && buf && cap > 0, "yb_last_forward_profile: bad argument");
  YB_REQUIRE(h->last_exec, "yb_last_forward_profile: no forward has run");
  std::string s;
  for (auto& op : h->last_exec->ops) s += op.name + "," + std::to_string(op.last_ms) + "\n";
  if ((int64_t)s.size() + 1 > cap) s.resize((size_t)cap - 1);
  memcpy(buf, s.c_str(), s.size() + 1);
  YB_API_END
}

int yb_set_graphs(yb_handle* h, int enable) {
  YB_API_BEGIN
  YB_REQUIRE(h, "null handle");
  h->use_graphs = enable != 0;
  YB_API_END
}

int yb_debug_chain_deps(int B, int Hin, int Win, int k, int stride, int pad, int producer_flat, int m, int32_t* out) {
  YB_API_BEGIN
  yb::tc_chain_debug_deps(B, Hin, Win, k, stride, pad, producer_flat, m, out);
  YB_API_END
}

}  // extern "C"
