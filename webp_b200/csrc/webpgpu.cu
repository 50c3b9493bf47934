// libwebpgpu: C ABI (include/webpgpu.h) over the sm_100a kernels.  One wgpu_ctx == one GPU + one stream.
// There is no CPU fallback anywhere in this file: every pixel stage is a kernel launch, and a missing /
// failing CUDA device is an error (WGPU_ERR_CUDA), never a silent host path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/webpgpu.h"
#include "misc_kernels.cuh"
#include "token_kernels.cuh"
#include "boolcode_par.cuh"
#include "part0_kernels.cuh"
#include "enc_phased.cuh"
#include "dec_parse.cuh"
#include "sharp_kernels.cuh"
#include "ssim_sep.cuh"
#include "host_dec.h"

namespace {

std::string g_create_error;

// Device image-batch allocator (the GPU analogue of internal/pool/pool.go:14-71).  Every device / pinned buffer of a context
// is grow-only and registers itself with its context; capacities come in size classes -- the reference's seven buckets up to
// 1 MiB, above that sixteenths of the next power of two (at most 12.5 % slack) -- so a batch a little larger than the last
// one reuses the buffer instead of reallocating.  wgpu_ctx_mem_info reports what a context holds, wgpu_ctx_trim gives the
// working buffers back (constant tables stay) and resets every flag that described their contents.
struct DevBuf;
struct PinBuf;
struct BufRegistry { std::vector<DevBuf*> dev; std::vector<PinBuf*> pin; };
thread_local BufRegistry* g_reg = nullptr;  // set while a wgpu_ctx's members are being constructed

inline size_t pool_bucket(size_t n) {
  static const size_t classes[7] = {256, 1024, 4096, 16384, 65536, 262144, 1048576};  // pool.go:5-13
  for (size_t c : classes) if (n <= c) return c;
  size_t p = 1;
  while (p < n) p <<= 1;
  const size_t step = p >> 4;
  return (n + step - 1) / step * step;
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool table = false;  // constant table uploaded once (upload_table): survives wgpu_ctx_trim
  DevBuf() { if (g_reg) g_reg->dev.push_back(this); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  bool reserve(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = pool_bucket(n);
    if (cudaMalloc(&p, want) != cudaSuccess) {  // the bucket's slack does not fit: take the exact size
      cudaGetLastError();
      want = n;
      if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return false; }
    }
    cap = want;
    return true;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  PinBuf() { if (g_reg) g_reg->pin.push_back(this); }
  PinBuf(const PinBuf&) = delete;
  PinBuf& operator=(const PinBuf&) = delete;
  bool reserve(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = pool_bucket(n);
    if (cudaMallocHost(&p, want) != cudaSuccess) {
      cudaGetLastError();
      want = n;
      if (cudaMallocHost(&p, want) != cudaSuccess) { p = nullptr; return false; }
    }
    cap = want;
    return true;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
bool trace_on() { static int v = -1; if (v < 0) v = getenv("WGPU_TRACE") ? 1 : 0; return v == 1; }

template <class F>
void parallel_for(int n, int threads, F f) {
  if (threads <= 1 || n <= 1) { for (int i = 0; i < n; ++i) f(i); return; }
  if (threads > n) threads = n;
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&]() { for (int i; (i = next.fetch_add(1)) < n;) f(i); });
  for (auto& th : pool) th.join();
}

}  // namespace

struct wgpu_ctx {
  BufRegistry reg;  // every DevBuf / PinBuf member below registers here (declaration order: reg_open first, reg_close last)
  struct RegOpen { explicit RegOpen(BufRegistry* r) { g_reg = r; } } reg_open{&reg};
  int dev = 0;
  int sm_count = 148;  // of this context's device
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: the chroma chains of the plane-split serial RD path
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_hdr = nullptr, ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  uint64_t launches = 0;
  uint64_t xfer_h2d = 0, xfer_d2h = 0;  // bytes copied by this context since the last wgpu_transfer_bytes(reset)
  int host_threads = 0;
  std::mutex mu;
  // constant tables
  DevBuf t_lc, t_eob, t_lfc, t_i4cost, t_g2l, t_l2g, t_proba0, t_upd, t_ecost;
  // encoder state (device)
  DevBuf rgba, sy, su, sv, ry, ru, rv, alpha, uv_alpha, segment, img_params, ctxw, ctxw2, ctxw_uv, derr, dither_y, dither_uv, hdr, coeffs, stats, proba, mb_tokens, mb_offset, img_total, img_base, tokens, coded, coded_size, coded_packed, pack_offsets, bcp_work, p0_plan, p0_info, p0_mb_tokens, p0_mb_offset, p0_total, t_i4paths, lc_img, eob_img, stats_cuts, hdr_prev, coeffs_prev;
  DevBuf sharp_best_y, sharp_target_y, sharp_best_uv, sharp_target_uv, t_sharp;  // SharpYUV import working planes + gamma tables
  PinBuf h_stats_cuts, h_lc_img, h_coded_size, h_alpha, h_uv_alpha, h_segment, h_params, h_hdr, h_coeffs, h_stats, h_proba, h_totals, h_bases, h_tokens, h_bcp, h_p0, h_packed;
  int e_n = 0, e_w = 0, e_h = 0, e_mbw = 0, e_mbh = 0, e_rgba_stride = 0;
  std::vector<long long> e_alpha_sum;  // per image: sum of the analysis alphas (low = busy picture), taken by enc_analyze_locked; e_order_valid until the next upload
  bool e_order_valid = false;
  DevBuf img_order; PinBuf h_img_order;
  bool e_uploaded = false, e_analyzed = false, e_done = false, e_keep_derr = false, e_keep_stats = false;
  int dither_w = 0, dither_h = 0, dither_amp_cached = 0;  // what the device dither tables currently hold
  wgpu_enc_options e_opt;
  std::vector<wgh::FramePlan> plans;
  // serial RD path with mid-stream probability refreshes: per-image probability state (carried across refreshes and
  // rate-control passes, like enc.proba), its history over the last pass (what each macroblock's tokens were recorded under)
  std::vector<uint8_t> sp_proba;
  std::vector<std::vector<uint8_t>> sp_hist;  // [image] -> tables [k][1056] in force from macroblock sp_starts[k] on
  std::vector<int> sp_starts;
  bool e_refresh_route = false, e_token_route = false;
  // decoder state
  DevBuf d_coeffs, d_meta, d_ftype, dy, du, dv, d_nrgba, d_alpha, d_streams, d_hdrs, d_perr, t_bmodes;
  PinBuf hd_streams, hd_hdrs, hd_perr;
  bool d_dev_parsed = false;
  // CUDA graph of the decoder's device step (444 wave launches of a few microseconds each are launch-bound otherwise)
  cudaGraphExec_t d_graph = nullptr;
  uint64_t d_graph_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t d_graph_launches = 0;
  PinBuf hd_coeffs, hd_meta, hd_ftype, hd_planes, hd_nrgba;
  int d_n = 0, d_w = 0, d_h = 0, d_mbw = 0, d_mbh = 0;
  bool d_ready = false, d_any_filter = false, d_has_nrgba = false;
  // metrics
  DevBuf m_a, m_b, m_sse_part, m_ssim_part, m_sse, m_ssim;
  struct RegClose { RegClose() { g_reg = nullptr; } } reg_close;
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
      return e_ == cudaErrorMemoryAllocation ? WGPU_ERR_NOMEM : WGPU_ERR_CUDA;                     \
    }                                                                                              \
  } while (0)
#define RESERVE(buf, bytes)                                                                        \
  do {                                                                                             \
    if (!(buf).reserve(bytes)) { ctx->err = "out of memory reserving " #buf; cudaGetLastError(); return WGPU_ERR_NOMEM; } \
  } while (0)
#define FAIL(code, msg) do { ctx->err = (msg); return (code); } while (0)

static int threads_of(const wgpu_ctx* ctx);
static int getenv_int(const char* name, int dflt) { const char* e = getenv(name); return (e && *e) ? atoi(e) : dflt; }
// Where the token partition is boolean-coded.  On the GPU the coder is chunk-parallel (boolcode_par.cuh: ~4.5 ms for the 480 M
// tokens of a 256-image batch, 0.3 ms for one 1536x1024 frame), so it is the route for every batch size; the host coder (a
// thread codes ~4 ns per token) stays for multi-partition frames and as the cross-check WGPU_DEVICE_CODER=0 selects.
static bool device_coder_wanted(const wgpu_ctx* ctx, size_t n_images) {
  const char* e = getenv("WGPU_DEVICE_CODER");  // read per call: tests flip it
  if (e && *e) return atoi(e) != 0;
  (void)ctx; (void)n_images;
  return true;
}
// Where the macroblock data of the decoder is parsed (intra modes + coefficient tokens; the frame headers always on the
// host).  On the GPU: one warp per image (dec_parse_kernel), 46 MB of compressed bytes up instead of 1.26 GB of
// coefficients per 256-image batch, no host cores.  WGPU_DEVICE_PARSER=0/1 forces either.
static bool device_parser_wanted(const wgpu_ctx* ctx, size_t n_images) {
  (void)ctx;
  const char* e = getenv("WGPU_DEVICE_PARSER");  // read per call: tests flip it
  if (e && *e) return atoi(e) != 0;
  return n_images >= 32;
}
static int threads_of(const wgpu_ctx* ctx) {
  int t = ctx->host_threads;
  if (t <= 0) t = (int)std::thread::hardware_concurrency();
  return t > 0 ? t : 1;
}

extern "C" {

void wgpu_enc_options_default(wgpu_enc_options* o, int quality) {
  // DefaultOptions (encode.go:196-214) mapped onto lossy.EncodeConfig (internal/lossy/encode.go:66-86)
  o->quality = quality; o->method = 4; o->sns_strength = 50; o->filter_strength = 60; o->filter_sharpness = 0;
  o->filter_type = 1; o->partitions = 0; o->segments = 4; o->preprocessing = 0; o->has_alpha = 0; o->passes = 1; o->dither_amp = 0;
  o->target_size = 0; o->target_psnr = 0.f; o->qmin = 0; o->qmax = 100; o->use_sharp_yuv = 0;
}

static int upload_table(wgpu_ctx* ctx, DevBuf& b, const void* src, size_t bytes) {
  RESERVE(b, bytes);
  b.table = true;
  CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)(bytes);
  return 0;
}

int wgpu_ctx_create(int device_ordinal, wgpu_ctx** out) {
  if (!out) return WGPU_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                     " (libwebpgpu has no CPU fallback)";
    cudaGetLastError();
    return WGPU_ERR_CUDA;
  }
  if (device_ordinal < 0 || device_ordinal >= count) { g_create_error = "device ordinal out of range"; return WGPU_ERR_INVALID; }
  wgpu_ctx* ctx = new wgpu_ctx();
  ctx->dev = device_ordinal;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal);
  auto bail = [&](const char* what, cudaError_t ce) { g_create_error = std::string(what) + ": " + cudaGetErrorString(ce); delete ctx; return WGPU_ERR_CUDA; };
  if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_hdr, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  // tables
  static uint16_t i4costs[1000];
  wgh::compute_i4_costs(i4costs);
  static uint16_t g2l[256], l2g[34];
  for (int i = 0; i < 256; ++i) {  // yuv.go:193-215 (kGamma = 0.80, 12-bit linear, 32-entry interpolation table)
    const double v = (double)i / 255.0;
    g2l[i] = (uint16_t)((v <= 0 ? 0.0 : pow(v, 0.80)) * 4095.0 + 0.5);
  }
  for (int i = 0; i <= 32; ++i) {
    const double v = (128.0 / 4095.0) * (double)i;
    l2g[i] = (uint16_t)((v <= 0 ? 0.0 : pow(v, 1.0 / 0.80)) * 255.0 + 0.5);
  }
  l2g[33] = 255;
  int rc = 0;
  // folded per-(type, band, ctx, level) token costs under the default probabilities (vp8_dev.cuh CostTabs)
  static_assert(wg::LC_LEVELS == 68, "host_enc.h build_cost_tables folds 68 levels");
  static uint16_t lc[wg::LC_SIZE], eobc[wg::EOB_SIZE];
  wgh::build_cost_tables(wgh::kCoeffsProba0, lc, eobc);
  rc |= upload_table(ctx, ctx->t_lc, lc, sizeof(lc));
  rc |= upload_table(ctx, ctx->t_eob, eobc, sizeof(eobc));
  rc |= upload_table(ctx, ctx->t_lfc, wgh::kLevelFixedCosts, sizeof(wgh::kLevelFixedCosts));
  rc |= upload_table(ctx, ctx->t_proba0, wgh::kCoeffsProba0, sizeof(wgh::kCoeffsProba0));
  rc |= upload_table(ctx, ctx->t_bmodes, wgh::kBModesProba, sizeof(wgh::kBModesProba));
  static_assert(sizeof(wg::I4PathDev) == sizeof(wgh::I4Path), "I4Path layout");
  rc |= upload_table(ctx, ctx->t_i4paths, wgh::i4_paths(), 10 * sizeof(wgh::I4Path));
  rc |= upload_table(ctx, ctx->t_upd, wgh::kCoeffsUpdateProba, sizeof(wgh::kCoeffsUpdateProba));
  rc |= upload_table(ctx, ctx->t_ecost, wgh::kEntropyCost, sizeof(wgh::kEntropyCost));
  rc |= upload_table(ctx, ctx->t_i4cost, i4costs, sizeof(i4costs));
  rc |= upload_table(ctx, ctx->t_g2l, g2l, sizeof(g2l));
  rc |= upload_table(ctx, ctx->t_l2g, l2g, sizeof(l2g));
  if (rc || cudaStreamSynchronize(ctx->stream) != cudaSuccess) { g_create_error = "table upload failed: " + ctx->err; delete ctx; return WGPU_ERR_CUDA; }
  *out = ctx;
  return WGPU_OK;
}

void wgpu_ctx_destroy(wgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->dev);
  cudaStreamSynchronize(ctx->stream);
  for (DevBuf* b : ctx->reg.dev) b->release();
  for (PinBuf* b : ctx->reg.pin) b->release();
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_hdr) cudaEventDestroy(ctx->ev_hdr);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->d_graph) cudaGraphExecDestroy(ctx->d_graph);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* wgpu_last_error(const wgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int wgpu_sync(wgpu_ctx* ctx) {
  if (!ctx) return WGPU_ERR_INVALID;
  CK(cudaSetDevice(ctx->dev));
  CK(cudaStreamSynchronize(ctx->stream));
  return WGPU_OK;
}
size_t wgpu_pool_bucket(size_t bytes) { return pool_bucket(bytes); }
int wgpu_ctx_mem_info(wgpu_ctx* ctx, size_t* device_bytes, size_t* pinned_bytes, int* buffers) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  size_t d = 0, h = 0;
  int k = 0;
  for (const DevBuf* b : ctx->reg.dev) { d += b->cap; k += b->p != nullptr; }
  for (const PinBuf* b : ctx->reg.pin) { h += b->cap; k += b->p != nullptr; }
  if (device_bytes) *device_bytes = d;
  if (pinned_bytes) *pinned_bytes = h;
  if (buffers) *buffers = k;
  return WGPU_OK;
}
int wgpu_ctx_trim(wgpu_ctx* ctx) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->dev));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->stream2) CK(cudaStreamSynchronize(ctx->stream2));
  if (ctx->d_graph) { cudaGraphExecDestroy(ctx->d_graph); ctx->d_graph = nullptr; }  // it holds the buffers' addresses
  memset(ctx->d_graph_key, 0, sizeof(ctx->d_graph_key));
  for (DevBuf* b : ctx->reg.dev) if (!b->table) b->release();
  for (PinBuf* b : ctx->reg.pin) b->release();
  // nothing on the device describes a picture any more
  ctx->e_uploaded = ctx->e_analyzed = ctx->e_done = ctx->e_keep_derr = ctx->e_keep_stats = ctx->e_order_valid = false;
  ctx->e_n = 0;
  ctx->dither_w = ctx->dither_h = ctx->dither_amp_cached = 0;
  ctx->d_ready = ctx->d_has_nrgba = ctx->d_dev_parsed = false;
  ctx->d_n = 0;
  return WGPU_OK;
}
int wgpu_set_host_threads(wgpu_ctx* ctx, int n) {
  if (!ctx || n < 0) return WGPU_ERR_INVALID;
  ctx->host_threads = n;
  return WGPU_OK;
}
void* wgpu_host_alloc(wgpu_ctx* ctx, size_t bytes) {
  if (!ctx) return nullptr;
  cudaSetDevice(ctx->dev);
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { ctx->err = "cudaMallocHost failed"; cudaGetLastError(); return nullptr; }
  return p;
}
void wgpu_host_free(wgpu_ctx* ctx, void* p) {
  if (ctx) cudaSetDevice(ctx->dev);
  if (p) cudaFreeHost(p);
}

// ======================================================================================== encoder
static int validate_enc_options(wgpu_ctx* ctx, const wgpu_enc_options* o, int width, int height) {
  // validateConfig (encode.go:259-334)
  if (!o) FAIL(WGPU_ERR_INVALID, "webp: nil options");
  if (width <= 0 || height <= 0 || width > 16383 || height > 16383) FAIL(WGPU_ERR_INVALID, "webp: invalid image dimensions");
  if (o->quality < 0 || o->quality > 100) FAIL(WGPU_ERR_INVALID, "webp: quality out of range [0, 100]");
  if (o->method < 0 || o->method > 6) FAIL(WGPU_ERR_INVALID, "webp: method out of range [0, 6]");
  if (o->sns_strength < 0 || o->sns_strength > 100) FAIL(WGPU_ERR_INVALID, "webp: sns strength out of range [0, 100]");
  if (o->filter_strength < 0 || o->filter_strength > 100) FAIL(WGPU_ERR_INVALID, "webp: filter strength out of range [0, 100]");
  if (o->filter_sharpness < 0 || o->filter_sharpness > 7) FAIL(WGPU_ERR_INVALID, "webp: filter sharpness out of range [0, 7]");
  if (o->filter_type < 0 || o->filter_type > 1) FAIL(WGPU_ERR_INVALID, "webp: filter type out of range [0, 1]");
  if (o->partitions < 0 || o->partitions > 3) FAIL(WGPU_ERR_INVALID, "webp: partitions out of range [0, 3]");
  if (o->segments < 1 || o->segments > 4) FAIL(WGPU_ERR_INVALID, "webp: segments out of range [1, 4]");
  // The GPU path restates the reference's row-parallel encoder (internal/lossy/encode.go:1356):
  // Method >= 3, mbH >= 4, single pass.  The serial-path configurations are a later row of the scope table.
  if (o->passes < 0 || o->passes > 10) FAIL(WGPU_ERR_INVALID, "webp: invalid Pass (must be 1-10 or 0 for default)");
  if (o->dither_amp < 0 || o->dither_amp > 256) FAIL(WGPU_ERR_INVALID, "webp: dithering amplitude out of range [0, 256]");
  // Method < 3: statLoop + serial encodeFrame semantics (non-RD decisions) -- built.  Method >= 3 on frames of fewer than
  // 4 macroblock rows takes the reference's serial RD path (probability refreshes feed the RD costs) -- not built yet.
  if (o->target_size < 0) FAIL(WGPU_ERR_INVALID, "webp: invalid TargetSize (must be >= 0)");
  if (!(o->target_psnr >= 0.f) || o->target_psnr > 1e30f) FAIL(WGPU_ERR_INVALID, "webp: invalid TargetPSNR (must be >= 0, finite)");
  if (o->qmin < 0 || o->qmax > 100 || (o->qmax > 0 && o->qmin > o->qmax)) FAIL(WGPU_ERR_INVALID, "webp: invalid QMin/QMax (must be 0-100, QMin <= QMax)");
  // The serial RD path (Method >= 3 on fewer than 4 macroblock rows, or any size under TargetSize / TargetPSNR) is built only
  // where no mid-stream probability refresh can occur (<= 96 macroblocks: encode_frame.go:24-40).
  const bool do_search = o->target_size > 0 || o->target_psnr > 0.f;
  // With several token partitions the reference emits [mbStart[i], mbStart[i+1]) per macroblock, and mbStart keeps entries of
  // EARLIER passes for macroblocks skipped in the last one: rate control x partitions would need every pass's token counts.
  if (do_search && o->partitions > 0) FAIL(WGPU_ERR_UNSUPPORTED, "TargetSize/TargetPSNR with Partitions > 0 is not built (the reference's per-macroblock token start table carries stale entries across passes)");
  if (o->method >= 3 && (((height + 15) >> 4) < 4 || do_search) && ((height + 15) >> 4) * ((width + 15) >> 4) > 96 && o->partitions > 0)
    FAIL(WGPU_ERR_UNSUPPORTED, "the serial RD path with mid-stream probability refreshes (more than 96 macroblocks) is built for one token partition only");
  return WGPU_OK;
}

int wgpu_enc_upload(wgpu_ctx* ctx, const uint8_t* rgba, int n, int width, int height, int stride, size_t image_stride) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!rgba || n <= 0 || width <= 0 || height <= 0 || stride < 4 * width) FAIL(WGPU_ERR_INVALID, "webp: invalid input image");
  if (width > 16383 || height > 16383) FAIL(WGPU_ERR_INVALID, "webp: invalid image dimensions");
  CK(cudaSetDevice(ctx->dev));
  const int dstride = (4 * width + 15) & ~15;
  RESERVE(ctx->rgba, (size_t)n * height * dstride);
  if (image_stride == (size_t)stride * height) {
    CK(cudaMemcpy2DAsync(ctx->rgba.p, dstride, rgba, stride, (size_t)4 * width, (size_t)n * height, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)4 * width * height * n;
  } else {
    for (int i = 0; i < n; ++i)
      CK(cudaMemcpy2DAsync(ctx->rgba.as<uint8_t>() + (size_t)i * height * dstride, dstride, rgba + (size_t)i * image_stride, stride,
                           (size_t)4 * width, height, cudaMemcpyHostToDevice, ctx->stream));
    if (image_stride != (size_t)stride * height) ctx->xfer_h2d += (uint64_t)4 * width * height * n;
  }
  ctx->e_n = n; ctx->e_w = width; ctx->e_h = height; ctx->e_mbw = (width + 15) >> 4; ctx->e_mbh = (height + 15) >> 4;
  ctx->e_rgba_stride = dstride;
  ctx->e_uploaded = true; ctx->e_order_valid = false;
  ctx->e_analyzed = false;
  ctx->e_done = false;
  return WGPU_OK;
}

extern "C++" {
namespace {
int wave_rows(int wave, int mb_w, int mb_h) {
  const int y_lo = std::max(0, (wave - (mb_w - 1) + 1) >> 1), y_hi = std::min(mb_h - 1, wave >> 1);
  return y_hi - y_lo + 1;
}
template <int G, int WARPS, int MINB>
int launch_enc_fast_waves(wgpu_ctx* ctx, const wg::EncKernelParams& P) {  // Method < 3: non-RD body, same wave schedule
  constexpr int per_cta = WARPS * (32 / G);
  constexpr size_t smem = sizeof(wg::MBShared) * per_cta;
  {  // the attribute is per device (contexts may sit on different GPUs of one process): set it on every call
    cudaError_t e = cudaFuncSetAttribute(wg::encode_fast_wave_kernel<G, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ctx->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return WGPU_ERR_CUDA; }
  }
  const int waves = P.mb_w + 2 * (P.mb_h - 1);
  for (int w = 0; w < waves; ++w) {
    const long long tasks = (long long)wave_rows(w, P.mb_w, P.mb_h) * P.n_images;
    if (tasks <= 0) continue;  // one macroblock column: odd waves hold no macroblock (x = w - 2y)
    const unsigned grid = (unsigned)((tasks + per_cta - 1) / per_cta);
    wg::encode_fast_wave_kernel<G, WARPS, MINB><<<grid, WARPS * 32, smem, ctx->stream>>>(P, w);
    ctx->launches++;
  }
  return WGPU_OK;
}
// Serial RD path (Method >= 3, fewer than 4 macroblock rows): raster order, one launch per macroblock index over the batch.
template <int G, int WARPS, int MINB>
int launch_enc_serial(wgpu_ctx* ctx, const wg::EncKernelParams& P, int mb_begin, int mb_end) {
  constexpr int per_cta = WARPS * (32 / G);
  constexpr size_t smem = sizeof(wg::MBShared) * per_cta;
  {
    cudaError_t e = cudaFuncSetAttribute(wg::encode_serial_kernel<G, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ctx->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return WGPU_ERR_CUDA; }
  }
  const unsigned grid = (unsigned)((P.n_images + per_cta - 1) / per_cta);
  for (int i = mb_begin; i < mb_end; ++i) {
    wg::encode_serial_kernel<G, WARPS, MINB><<<grid, WARPS * 32, smem, ctx->stream>>>(P, i);
    ctx->launches++;
  }
  return WGPU_OK;
}
// Serial RD path with refreshes, split by plane (enc_kernels.cuh): luma waves over the macroblocks [mb_begin, mb_end) on the
// context's stream, the chroma chains of the same segment on its second stream, then the merge of the two halves.
int launch_enc_serial_split(wgpu_ctx* ctx, wg::EncKernelParams& P, int mb_begin, int mb_end) {
  constexpr int G = 8;
  const int mbw = P.mb_w, mbh = P.mb_h, nmb = mbw * mbh, n = P.n_images;
  const size_t max_smem = sizeof(wg::MBShared) * 4 + (size_t)4 * (wg::LC_SIZE + wg::EOB_SIZE) * 2;
  cudaError_t e = cudaFuncSetAttribute(wg::encode_serial_luma_wave_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(wg::encode_serial_chroma_chain_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
  if (e != cudaSuccess) { ctx->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return WGPU_ERR_CUDA; }
  P.mb_begin = mb_begin; P.mb_end = mb_end;
  // chroma: everything queued on the main stream so far (the cost tables of this segment) comes first
  CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
  {
    wg::EncKernelParams C = P;
    C.serial_wave = 0;
    C.serial_gpw = n <= 8 * ctx->sm_count ? 1 : (n <= 24 * ctx->sm_count ? 2 : 4);
    const size_t smem = sizeof(wg::MBShared) * (32 / G) + (size_t)C.serial_gpw * (wg::LC_SIZE + wg::EOB_SIZE) * 2;
    wg::encode_serial_chroma_chain_kernel<G><<<(unsigned)((n + C.serial_gpw - 1) / C.serial_gpw), 32, smem, ctx->stream2>>>(C);
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
  }
  // luma: the waves that hold a macroblock of the segment
  const int y_first = mb_begin / mbw, y_last = (mb_end - 1) / mbw;
  for (int w = 0; w < mbw + 2 * (mbh - 1); ++w) {
    const int rows = wave_rows(w, mbw, mbh);
    if (rows <= 0) continue;
    const int y_lo = std::max(0, (w - (mbw - 1) + 1) >> 1), y_hi = std::min(mbh - 1, w >> 1);
    bool any = false;
    for (int y = std::max(y_lo, y_first); y <= std::min(y_hi, y_last) && !any; ++y) {
      const int idx = y * mbw + (w - 2 * y);
      any = idx >= mb_begin && idx < mb_end;
    }
    if (!any) continue;
    wg::EncKernelParams L = P;
    L.serial_wave = 1;
    const long long tasks = (long long)rows * n;
    L.serial_gpw = tasks <= 8LL * ctx->sm_count ? 1 : (tasks <= 24LL * ctx->sm_count ? 2 : 4);
    const size_t smem = sizeof(wg::MBShared) * (32 / G) + (size_t)L.serial_gpw * (wg::LC_SIZE + wg::EOB_SIZE) * 2;
    wg::encode_serial_luma_wave_kernel<G><<<(unsigned)((tasks + L.serial_gpw - 1) / L.serial_gpw), 32, smem, ctx->stream>>>(L, w);
    ctx->launches++;
  }
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  wg::SerialMergeParams M;
  M.hdr = P.out_hdr; M.ctx = P.ctx; M.ctx_uv = P.ctx_uv; M.img = P.img; M.n_images = n; M.nmb = nmb; M.mb_begin = mb_begin; M.mb_end = mb_end;
  wg::serial_merge_kernel<<<(unsigned)(((long long)(mb_end - mb_begin) * n + 255) / 256), 256, 0, ctx->stream>>>(M);
  ctx->launches++;
  CK(cudaGetLastError());
  return WGPU_OK;
}
// Row-parallel RD path: the phase-synchronous kernel (enc_phased.cuh).  M macroblocks per CTA, 16 threads per macroblock;
// narrow waves take smaller CTAs so that they still spread over the SMs.  Token statistics follow in one pass (mb_stats_kernel).
template <int M, int NT, int MINB>
int launch_phased_wave(wgpu_ctx* ctx, const wg::EncKernelParams& P, int w, long long tasks) {
  constexpr size_t smem = sizeof(wg::PhMB) * M;
  cudaError_t e = cudaFuncSetAttribute(wg::encode_phased_kernel<M, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { ctx->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return WGPU_ERR_CUDA; }
  wg::encode_phased_kernel<M, NT, MINB><<<(unsigned)((tasks + M - 1) / M), NT, smem, ctx->stream>>>(P, w);
  ctx->launches++;
  return WGPU_OK;
}
int launch_enc_phased(wgpu_ctx* ctx, const wg::EncKernelParams& P) {
  const int sm_count = ctx->sm_count;
  const int waves = P.mb_w + 2 * (P.mb_h - 1);
#ifdef WG_PHASE_CLOCK
  static unsigned long long* clk = nullptr;
  if (!clk) cudaMalloc(&clk, 256 * 8);
  cudaMemsetAsync(clk, 0, 256 * 8, ctx->stream);
  const int clk_wave = getenv_int("WGPU_CLOCK_WAVE", 110);
  const wg::EncKernelParams P0 = P;
#endif
  for (int w = 0; w < waves; ++w) {
    const long long tasks = (long long)wave_rows(w, P.mb_w, P.mb_h) * P.n_images;
    if (tasks <= 0) continue;  // one macroblock column: odd waves hold no macroblock (x = w - 2y)
    int rc;
#ifdef WG_PHASE_CLOCK
    wg::EncKernelParams& Pm = const_cast<wg::EncKernelParams&>(P);
    Pm = P0;
    if (w == clk_wave) { Pm.phase_clock = clk; Pm.clock_cta = getenv_int("WGPU_CLOCK_CTA", 300); }
#endif
    if (tasks <= (long long)sm_count * 8 * 4) rc = launch_phased_wave<8, 128, 4>(ctx, P, w, tasks);  // narrow wave: smaller CTAs reach more SMs
    else rc = launch_phased_wave<16, 256, 3>(ctx, P, w, tasks);
    if (rc) return rc;
  }
#ifdef WG_PHASE_CLOCK
  {
    unsigned long long h[256];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[phase clock] wave %d:", clk_wave);
    for (int i = 1; i < 256 && h[i]; ++i) fprintf(stderr, " %llu", h[i] - h[i - 1]);
    fprintf(stderr, "\n");
  }
#endif
  if (P.stats) {
    wg::MBStatsParams S;
    S.hdr = P.out_hdr; S.coeffs = P.out_coeffs; S.ctxw = P.ctx; S.stats = P.stats; S.mb_w = P.mb_w; S.mb_h = P.mb_h;
    S.mbs_per_cta = 128;
    const int nmb = P.mb_w * P.mb_h;
    wg::mb_stats_kernel<<<dim3((unsigned)((nmb + S.mbs_per_cta - 1) / S.mbs_per_cta), (unsigned)P.n_images), 256, 0, ctx->stream>>>(S);
    ctx->launches++;
  }
  return WGPU_OK;
}
}  // namespace
}  // extern "C++"

extern "C++" {
template <int ITEMS>
static void launch_sharp_ring(const wg::SharpParams& P, int n, cudaStream_t st) {
  cudaFuncSetAttribute(wg::sharp_refine_ring_kernel<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
  wg::sharp_refine_ring_kernel<ITEMS><<<(unsigned)n, 256, (size_t)24 * P.uv_w, st>>>(P);
}
}  // extern "C++"
// EncoderOptions.UseSharpYUV (encode.go:531-535): sharpyuv.Convert with the WebP matrix and the sRGB transfer, then importYCbCr.
static int enc_launch_import_sharp(wgpu_ctx* ctx) {
  const int n = ctx->e_n, pad_w = ctx->e_mbw * 16, pad_h = ctx->e_mbh * 16;
  if (!ctx->t_sharp.p) {  // gamma tables (sharpyuv/gamma.go:47-91): doubles, as the reference builds them
    std::vector<uint32_t> tab(wg::kSharpG2L + wg::kSharpL2G);
    wg::sharp_build_tables(tab.data());
    int rc = upload_table(ctx, ctx->t_sharp, tab.data(), tab.size() * 4);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));  // the host vector goes out of scope
  }
  wg::SharpParams P;
  P.rgba = ctx->rgba.as<uint8_t>(); P.image_stride = (size_t)ctx->e_h * ctx->e_rgba_stride; P.stride = ctx->e_rgba_stride;
  P.n = n; P.width = ctx->e_w; P.height = ctx->e_h;
  P.w = (ctx->e_w + 1) & ~1; P.h = (ctx->e_h + 1) & ~1; P.uv_w = P.w >> 1; P.uv_h = P.h >> 1;
  const size_t luma = (size_t)n * P.w * P.h, chroma = (size_t)n * 3 * P.uv_w * P.uv_h;
  RESERVE(ctx->sharp_best_y, luma * 2); RESERVE(ctx->sharp_target_y, luma * 2);
  RESERVE(ctx->sharp_best_uv, chroma * 2); RESERVE(ctx->sharp_target_uv, chroma * 2);
  P.best_y = ctx->sharp_best_y.as<uint16_t>(); P.target_y = ctx->sharp_target_y.as<uint16_t>();
  P.best_uv = ctx->sharp_best_uv.as<int16_t>(); P.target_uv = ctx->sharp_target_uv.as<int16_t>();
  P.g2l = ctx->t_sharp.as<uint32_t>(); P.l2g = P.g2l + wg::kSharpG2L;
  P.y = ctx->sy.as<uint8_t>(); P.u = ctx->su.as<uint8_t>(); P.v = ctx->sv.as<uint8_t>();
  P.y_plane = (size_t)pad_w * pad_h; P.uv_plane = P.y_plane / 4; P.pad_w = pad_w; P.pad_h = pad_h;
  P.iterations = nullptr;
  const long long blocks2 = (long long)P.uv_w * P.uv_h * n;
  wg::sharp_init_kernel<<<(unsigned)std::min<long long>((blocks2 + 255) / 256, 148LL * 64), 256, 0, ctx->stream>>>(P);
  // refinement: operands fetched a row pair ahead, residual rows in a shared-memory ring (11.2 ms per 256 x 1536x1024 against 17.6 ms
  // for the plain row-pair sweep, 3.5 vs 5.3 ms for one image); rows wider than the ring take the plain sweep with 1024 threads
  if (P.uv_w <= 256 * wg::SHARP_ITEMS) {
    const int items = (P.uv_w + 255) / 256;
    if (items <= 1) launch_sharp_ring<1>(P, n, ctx->stream);
    else if (items == 2) launch_sharp_ring<2>(P, n, ctx->stream);
    else if (items == 3) launch_sharp_ring<3>(P, n, ctx->stream);
    else if (items == 4) launch_sharp_ring<4>(P, n, ctx->stream);
    else if (items <= 6) launch_sharp_ring<6>(P, n, ctx->stream);
    else launch_sharp_ring<8>(P, n, ctx->stream);
  }
  else wg::sharp_refine_kernel<1024><<<(unsigned)n, 1024, 0, ctx->stream>>>(P);  // up to 8192 chroma samples per row: any WebP width
  const long long blocks3 = (long long)(pad_w / 2) * (pad_h / 2) * n;
  wg::sharp_finish_kernel<<<(unsigned)std::min<long long>((blocks3 + 255) / 256, 148LL * 64), 256, 0, ctx->stream>>>(P);
  ctx->launches += 3;
  CK(cudaGetLastError());
  return WGPU_OK;
}
static int enc_launch_import(wgpu_ctx* ctx) {
  if (ctx->e_opt.use_sharp_yuv) return enc_launch_import_sharp(ctx);
  const int n = ctx->e_n, pad_w = ctx->e_mbw * 16, pad_h = ctx->e_mbh * 16;
  wg::ImportParams ip;
  ip.rgba = ctx->rgba.as<uint8_t>(); ip.image_stride = (size_t)ctx->e_h * ctx->e_rgba_stride; ip.stride = ctx->e_rgba_stride;
  ip.n = n; ip.width = ctx->e_w; ip.height = ctx->e_h; ip.pad_w = pad_w; ip.pad_h = pad_h; ip.has_alpha = ctx->e_opt.has_alpha;
  ip.y = ctx->sy.as<uint8_t>(); ip.u = ctx->su.as<uint8_t>(); ip.v = ctx->sv.as<uint8_t>();
  ip.y_plane = (size_t)pad_w * pad_h; ip.uv_plane = ip.y_plane / 4;
  ip.gamma_to_linear = ctx->t_g2l.as<uint16_t>(); ip.linear_to_gamma = ctx->t_l2g.as<uint16_t>();
  ip.dither_y = nullptr; ip.dither_uv = nullptr;
  if (ctx->e_opt.dither_amp > 0) {
    if (ctx->dither_w != pad_w || ctx->dither_h != pad_h || ctx->dither_amp_cached != ctx->e_opt.dither_amp) {
      // VP8Random (internal/dsp/random.go:17-79) run once on the host in the reference's draw order (encode.go:793-809, 925-936)
      std::vector<uint16_t> dy((size_t)pad_w * pad_h);
      std::vector<uint32_t> duv((size_t)(pad_w / 2) * (pad_h / 2) * 2);
      uint32_t tab[55];
      memcpy(tab, wgh::kRandomTable, sizeof(tab));
      int i1 = 0, i2 = 31;
      const int amp = ctx->e_opt.dither_amp;
      auto bits = [&](int num_bits) {
        long long diff = (long long)tab[i1] - (long long)tab[i2];
        if (diff < 0) diff += 1ll << 31;
        tab[i1] = (uint32_t)diff;
        if (++i1 == 55) i1 = 0;
        if (++i2 == 55) i2 = 0;
        int d = (int)((int32_t)((uint32_t)diff << 1) >> (32 - num_bits));
        d = (d * amp) >> 8;
        return d + (1 << (num_bits - 1));
      };
      for (size_t i = 0; i < dy.size(); ++i) dy[i] = (uint16_t)bits(16);
      for (size_t i = 0; i < duv.size(); ++i) duv[i] = (uint32_t)bits(18);
      RESERVE(ctx->dither_y, dy.size() * 2);
      RESERVE(ctx->dither_uv, duv.size() * 4);
      CK(cudaMemcpyAsync(ctx->dither_y.p, dy.data(), dy.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
      ctx->xfer_h2d += (uint64_t)(dy.size() * 2);
      CK(cudaMemcpyAsync(ctx->dither_uv.p, duv.data(), duv.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
      ctx->xfer_h2d += (uint64_t)(duv.size() * 4);
      CK(cudaStreamSynchronize(ctx->stream));  // the host vectors go out of scope
      ctx->dither_w = pad_w; ctx->dither_h = pad_h; ctx->dither_amp_cached = amp;
    }
    ip.dither_y = ctx->dither_y.as<uint16_t>(); ip.dither_uv = ctx->dither_uv.as<uint32_t>();
  }
  const dim3 block(64, 4), grid((unsigned)((pad_w / 4 + 63) / 64), (unsigned)((pad_h / 2 + 3) / 4), (unsigned)n);
  const bool dither = ip.dither_y != nullptr;
  if (ip.has_alpha) {
    if (dither) wg::import_rgba_kernel<true, true><<<grid, block, 0, ctx->stream>>>(ip);
    else wg::import_rgba_kernel<true, false><<<grid, block, 0, ctx->stream>>>(ip);
  } else {
    if (dither) wg::import_rgba_kernel<false, true><<<grid, block, 0, ctx->stream>>>(ip);
    else wg::import_rgba_kernel<false, false><<<grid, block, 0, ctx->stream>>>(ip);
  }
  ctx->launches++;
  CK(cudaGetLastError());
  return WGPU_OK;
}
static int enc_launch_analysis(wgpu_ctx* ctx) {
  const int n = ctx->e_n, nmb = ctx->e_mbw * ctx->e_mbh;
  wg::AnalysisParams ap;
  ap.y = ctx->sy.as<uint8_t>(); ap.u = ctx->su.as<uint8_t>(); ap.v = ctx->sv.as<uint8_t>();
  ap.y_plane = (size_t)nmb * 256; ap.uv_plane = (size_t)nmb * 64;
  ap.n = n; ap.mb_w = ctx->e_mbw; ap.mb_h = ctx->e_mbh; ap.width = ctx->e_w; ap.height = ctx->e_h;
  ap.alpha = ctx->alpha.as<uint8_t>(); ap.uv_alpha = ctx->uv_alpha.as<uint8_t>();
  (void)nmb;
  wg::analysis_kernel<<<dim3((unsigned)((ctx->e_mbw + 7) / 8), (unsigned)ctx->e_mbh, (unsigned)n), 128, 0, ctx->stream>>>(ap);
  ctx->launches++;
  CK(cudaGetLastError());
  return WGPU_OK;
}
static int enc_launch_waves(wgpu_ctx* ctx) {
  ctx->e_refresh_route = false;
  const int n = ctx->e_n, mbw = ctx->e_mbw, mbh = ctx->e_mbh, nmb = mbw * mbh;
  wg::EncKernelParams P;
#ifdef WG_PHASE_CLOCK
  P.phase_clock = nullptr; P.clock_cta = 0;
#endif
  P.src_y = ctx->sy.as<uint8_t>(); P.src_u = ctx->su.as<uint8_t>(); P.src_v = ctx->sv.as<uint8_t>();
  P.rec_y = ctx->ry.as<uint8_t>(); P.rec_u = ctx->ru.as<uint8_t>(); P.rec_v = ctx->rv.as<uint8_t>();
  P.segment = ctx->segment.as<uint8_t>(); P.img = ctx->img_params.as<wg::ImageParams>();
  P.stats = ctx->stats.as<unsigned int>();
  if (!ctx->e_keep_stats) CK(cudaMemsetAsync(ctx->stats.p, 0, (size_t)n * wg::STATS_SIZE * 4, ctx->stream));  // rate-control passes zero per image
  P.ctx2 = ctx->ctxw2.as<uint32_t>();
  P.ctx = ctx->ctxw.as<uint32_t>(); P.out_hdr = ctx->hdr.as<uint8_t>(); P.out_coeffs = ctx->coeffs.as<int16_t>();
  P.i4_costs = ctx->t_i4cost.as<uint16_t>(); P.lc = ctx->t_lc.as<uint16_t>(); P.eob = ctx->t_eob.as<uint16_t>(); P.lfc = ctx->t_lfc.as<uint16_t>();
  P.n_images = n; P.width = ctx->e_w; P.height = ctx->e_h; P.mb_w = mbw; P.mb_h = mbh;
  P.method = ctx->e_opt.method;
  P.max_i4_modes = ctx->e_opt.quality < 50 ? 2 : 3;  // getMaxI4RDModes (encode_parallel.go:931)
  P.y_plane = (size_t)nmb * 256; P.uv_plane = (size_t)nmb * 64;
  P.top_derr = nullptr; P.left_derr = nullptr; P.lc_img = nullptr; P.eob_img = nullptr; P.serial_gpw = 0;
  P.serial_wave = 0; P.mb_begin = 0; P.mb_end = nmb; P.ctx_uv = nullptr; P.img_order = nullptr;
  int rc;
  const bool do_search = ctx->e_opt.target_size > 0 || ctx->e_opt.target_psnr > 0.f;
  if (ctx->e_opt.method >= 3 && (mbh < 4 || do_search)) {  // useParallel == false (encode.go:1356)
    RESERVE(ctx->derr, (size_t)n * (mbw + 1) * 4);
    // topDerr / leftDerr are zeroed when the encoder is created, not between the passes of the rate-control loop
    if (!ctx->e_keep_derr) CK(cudaMemsetAsync(ctx->derr.p, 0, (size_t)n * (mbw + 1) * 4, ctx->stream));
    P.top_derr = ctx->derr.as<int8_t>();
    P.left_derr = ctx->derr.as<int8_t>() + (size_t)n * mbw * 4;
    // refreshProbas every max(total >> 3, 96) macroblocks (encode_frame.go:24-47): before macroblock k*M + (k-1), k = 1, 2, ..
    const int max_count = std::max(nmb >> 3, 96);
    ctx->e_refresh_route = nmb > max_count;
    if (!ctx->e_refresh_route) {
      rc = launch_enc_serial<8, 4, 3>(ctx, P, 0, nmb);
      if (rc) return rc;
      CK(cudaGetLastError());
      return WGPU_OK;
    }
    // ---- with refreshes: the RD costs follow the per-image probability state, rebuilt on the host at every refresh from
    // statistics over the whole per-macroblock array (this pass above the refresh point, the previous pass or the zero
    // state below it)
    if (!ctx->e_keep_derr) {  // first pass of this encode: default probabilities, zero-state macroblock array
      ctx->sp_proba.resize((size_t)n * 1056);
      for (int i = 0; i < n; ++i) memcpy(&ctx->sp_proba[(size_t)i * 1056], wgh::kCoeffsProba0, 1056);
      CK(cudaMemsetAsync(ctx->hdr.p, 0, (size_t)n * nmb * 48, ctx->stream));
    }
    P.stats = nullptr;  // statistics are taken by collect_all_stats_kernel, over the whole array
    RESERVE(ctx->ctxw_uv, (size_t)n * nmb * 4);
    P.ctx_uv = ctx->ctxw_uv.as<uint32_t>();
    RESERVE(ctx->lc_img, (size_t)n * wg::LC_SIZE * 2); RESERVE(ctx->eob_img, (size_t)n * wg::EOB_SIZE * 2);
    RESERVE(ctx->h_lc_img, (size_t)n * (wg::LC_SIZE + wg::EOB_SIZE) * 2);
    RESERVE(ctx->h_stats, (size_t)n * wg::STATS_SIZE * 4);
    P.lc_img = ctx->lc_img.as<uint16_t>(); P.eob_img = ctx->eob_img.as<uint16_t>();
    const wg::ImageParams* hp = ctx->h_params.as<wg::ImageParams>();  // seg[0].flags bit 8: parked by the rate-control loop
    ctx->sp_hist.resize(n); ctx->sp_starts.clear();
    for (int i = 0; i < n; ++i) if (!(hp[i].seg[0].flags & 0x100)) ctx->sp_hist[i].clear();
    auto upload_tables = [&](int first_mb) -> int {
      uint16_t* hl = ctx->h_lc_img.as<uint16_t>();
      uint16_t* he = hl + (size_t)n * wg::LC_SIZE;
      parallel_for(n, threads_of(ctx), [&](int i) { wgh::build_cost_tables(&ctx->sp_proba[(size_t)i * 1056], hl + (size_t)i * wg::LC_SIZE, he + (size_t)i * wg::EOB_SIZE); });
      CK(cudaMemcpyAsync(ctx->lc_img.p, hl, (size_t)n * wg::LC_SIZE * 2, cudaMemcpyHostToDevice, ctx->stream));
      ctx->xfer_h2d += (uint64_t)((size_t)n * wg::LC_SIZE * 2);
      CK(cudaMemcpyAsync(ctx->eob_img.p, he, (size_t)n * wg::EOB_SIZE * 2, cudaMemcpyHostToDevice, ctx->stream));
      ctx->xfer_h2d += (uint64_t)((size_t)n * wg::EOB_SIZE * 2);
      for (int i = 0; i < n; ++i)
        if (!(hp[i].seg[0].flags & 0x100)) ctx->sp_hist[i].insert(ctx->sp_hist[i].end(), &ctx->sp_proba[(size_t)i * 1056], &ctx->sp_proba[(size_t)i * 1056] + 1056);
      ctx->sp_starts.push_back(first_mb);
      return WGPU_OK;
    };
    auto all_stats = [&]() -> int {
      wg::AllStatsParams A;
      A.hdr = ctx->hdr.as<uint8_t>(); A.coeffs = ctx->coeffs.as<int16_t>(); A.stats = ctx->stats.as<unsigned int>();
      A.n_images = n; A.mb_w = mbw; A.mb_h = mbh; A.cut = nmb; A.hdr_prev = nullptr; A.coeffs_prev = nullptr;
      CK(cudaMemsetAsync(ctx->stats.p, 0, (size_t)n * wg::STATS_SIZE * 4, ctx->stream));
      wg::collect_all_stats_kernel<<<(unsigned)(((long long)n * nmb + 127) / 128), 128, 0, ctx->stream>>>(A);
      ctx->launches++;
      CK(cudaGetLastError());
      return WGPU_OK;
    };
    if ((rc = upload_tables(0))) return rc;
    int start = 0;
    for (int k = 1;; ++k) {
      const int end = std::min(k * max_count + (k - 1), nmb);
      if ((rc = launch_enc_serial_split(ctx, P, start, end))) return rc;
      if (end >= nmb) break;
      if ((rc = all_stats())) return rc;
      CK(cudaMemcpyAsync(ctx->h_stats.p, ctx->stats.p, (size_t)n * wg::STATS_SIZE * 4, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->xfer_d2h += (uint64_t)((size_t)n * wg::STATS_SIZE * 4);
      CK(cudaStreamSynchronize(ctx->stream));  // also: the table upload of the previous refresh has been consumed
      // images parked by the rate-control loop keep their state (their macroblocks are not re-encoded either)
      parallel_for(n, threads_of(ctx), [&](int i) {
        if (hp[i].seg[0].flags & 0x100) return;
        wgh::optimize_proba_host(*reinterpret_cast<const wgh::Stats*>(ctx->h_stats.as<uint32_t>() + (size_t)i * wg::STATS_SIZE),
                                 *reinterpret_cast<uint8_t (*)[4][8][3][11]>(&ctx->sp_proba[(size_t)i * 1056]));
      });
      if ((rc = upload_tables(end))) return rc;
      start = end;
    }
    if ((rc = all_stats())) return rc;  // collectAllStats for the final optimizeProba (encode.go:1378)
    CK(cudaGetLastError());
    return WGPU_OK;
  }
  if (ctx->e_opt.method < 3) {
    rc = launch_enc_fast_waves<8, 4, 3>(ctx, P);
    if (rc) return rc;
    CK(cudaGetLastError());
    return WGPU_OK;
  }
  if (ctx->e_order_valid && (int)ctx->e_alpha_sum.size() == n && n > 1 && getenv_int("WGPU_WAVE_ORDER", 1) != 0) {
    // A wave launch ends with its slowest CTAs, and those belong to the busiest pictures: give them the first slots of every
    // wave's task list (longest first), the easy pictures fill in behind.  The analysis alpha is high for smooth macroblocks
    // (finalAlphaValue, encode_analysis.go:237), so the lowest sum is the busiest picture.  Results do not depend on the order.
    RESERVE(ctx->img_order, (size_t)n * 4); RESERVE(ctx->h_img_order, (size_t)n * 4);
    int* ord = ctx->h_img_order.as<int>();
    for (int i = 0; i < n; ++i) ord[i] = i;
    const int sign = getenv_int("WGPU_WAVE_ORDER", 1) < 0 ? -1 : 1;
    std::stable_sort(ord, ord + n, [&](int a, int b) { return sign * ctx->e_alpha_sum[a] < sign * ctx->e_alpha_sum[b]; });
    CK(cudaMemcpyAsync(ctx->img_order.p, ord, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n * 4);
    P.img_order = ctx->img_order.as<int>();
  }
  rc = launch_enc_phased(ctx, P);
  if (rc) return rc;
  CK(cudaGetLastError());
  return WGPU_OK;
}

static int enc_reserve(wgpu_ctx* ctx) {
  const size_t n = ctx->e_n, nmb = (size_t)ctx->e_mbw * ctx->e_mbh;
  RESERVE(ctx->sy, n * nmb * 256); RESERVE(ctx->su, n * nmb * 64); RESERVE(ctx->sv, n * nmb * 64);
  RESERVE(ctx->ry, n * nmb * 256); RESERVE(ctx->ru, n * nmb * 64); RESERVE(ctx->rv, n * nmb * 64);
  RESERVE(ctx->alpha, n * nmb); RESERVE(ctx->uv_alpha, n * nmb); RESERVE(ctx->segment, n * nmb);
  RESERVE(ctx->img_params, n * sizeof(wg::ImageParams));
  RESERVE(ctx->ctxw, n * nmb * 4); RESERVE(ctx->ctxw2, n * nmb * 4); RESERVE(ctx->hdr, n * nmb * 48); RESERVE(ctx->coeffs, n * nmb * 800);
  RESERVE(ctx->stats, n * wg::STATS_SIZE * 4); RESERVE(ctx->h_stats, n * wg::STATS_SIZE * 4);
  RESERVE(ctx->proba, n * 1056); RESERVE(ctx->h_proba, n * 1056);
  RESERVE(ctx->mb_tokens, n * nmb * 4); RESERVE(ctx->mb_offset, n * nmb * 8);
  RESERVE(ctx->img_total, n * 8); RESERVE(ctx->img_base, 6 * n * 8); RESERVE(ctx->h_totals, n * 8); RESERVE(ctx->h_bases, 6 * n * 8);
  RESERVE(ctx->coded_size, n * 4); RESERVE(ctx->h_coded_size, n * 4);
  RESERVE(ctx->h_alpha, n * nmb); RESERVE(ctx->h_uv_alpha, n * nmb); RESERVE(ctx->h_segment, n * nmb);
  RESERVE(ctx->h_params, n * sizeof(wg::ImageParams));
  return WGPU_OK;
}

// Chunk-parallel boolean coder (boolcode_par.cuh): plan on the host from the per-partition token counts, then the usual number
// of relaxation rounds, the prefix sums, the byte pass and the joins, queued without a host round trip.  The per-round change
// counters come back with the coded sizes; finish_boolcode_par checks the last one.
constexpr int kBcpRounds = 8;
static int launch_bcp_tail(wgpu_ctx* ctx, const wg::BcpParams& BP) {
  const unsigned cb = (BP.n_chunks + 127) / 128;
  wg::bcp_scan_kernel<<<(unsigned)((BP.n_images * 32 + 127) / 128), 128, 0, ctx->stream>>>(BP);
  wg::bcp_bytes_kernel<<<cb, 128, 0, ctx->stream>>>(BP);
  wg::bcp_join_kernel<<<cb, 128, 0, ctx->stream>>>(BP);
  wg::bcp_join_fix_kernel<<<(unsigned)((BP.n_images + 63) / 64), 64, 0, ctx->stream>>>(BP);
  ctx->launches += 4;
  CK(cudaGetLastError());
  return WGPU_OK;
}
static int launch_boolcode_par(wgpu_ctx* ctx, const wg::BoolCodeParams& B, const unsigned long long* totals, size_t n, wg::BcpParams* out) {
  RESERVE(ctx->h_bcp, (n + 1) * 4 + 64 * 4);
  uint32_t* first = ctx->h_bcp.as<uint32_t>();
  uint32_t nchunks = 0;
  for (size_t i = 0; i < n; ++i) { first[i] = nchunks; nchunks += wg::bcp_chunks_of(totals[i]); }
  first[n] = nchunks;
  // device work area: chunk_first [n + 1] | changed [64] | any_pending [n] | shift, bit, head, head_carry, pending [chunks] x 4 B |
  // tail x 2 B | entry, walked x 1 B
  const size_t head_words = n + 1 + 64 + n;
  RESERVE(ctx->bcp_work, head_words * 4 + (size_t)nchunks * (20 + 2 + 2) + 64);
  uint32_t* w = ctx->bcp_work.as<uint32_t>();
  CK(cudaMemcpyAsync(w, first, (n + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)((n + 1) * 4);
  CK(cudaMemsetAsync(w + n + 1, 0, (64 + n) * 4, ctx->stream));
  wg::BcpParams BP;
  BP.tokens = B.tokens; BP.img_base = B.img_base; BP.img_total = B.img_total; BP.chunk_first = w; BP.n_images = (int)n; BP.n_chunks = nchunks;
  BP.changed = w + n + 1; BP.any_pending = w + n + 1 + 64;
  BP.shift_total = w + head_words; BP.chunk_bit = BP.shift_total + nchunks; BP.head = BP.chunk_bit + nchunks; BP.head_carry = BP.head + nchunks;
  BP.pending = BP.head_carry + nchunks;
  BP.tail = reinterpret_cast<uint16_t*>(BP.pending + nchunks);
  BP.entry = reinterpret_cast<uint8_t*>(BP.tail + nchunks); BP.walked = BP.entry + nchunks;
  BP.out = B.out; BP.out_base = B.out_base; BP.out_size = B.out_size;
  const unsigned cb = (nchunks + 127) / 128;
  for (int r = 0; r < kBcpRounds; ++r) {
    BP.round = r;
    wg::bcp_state_kernel<<<cb, 128, 0, ctx->stream>>>(BP);
    ctx->launches++;
  }
  int rc = launch_bcp_tail(ctx, BP);
  if (rc) return rc;
  CK(cudaMemcpyAsync(first + n + 1, BP.changed, 64 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += 64 * 4;
  *out = BP;
  return WGPU_OK;
}
// After the stream has drained: if the last queued round still changed an entry state (token streams whose range states
// merge slowly), keep relaxing -- a round that changes nothing is the exact fixed point -- and redo the passes behind it.
static int finish_boolcode_par(wgpu_ctx* ctx, wg::BcpParams& BP, bool* redone) {
  *redone = false;
  uint32_t* changed = ctx->h_bcp.as<uint32_t>() + BP.n_images + 1;
  int last = kBcpRounds - 1;
  if (changed[last] == 0) return WGPU_OK;
  const unsigned cb = (BP.n_chunks + 127) / 128;
  for (int guard = 0; guard < (1 << 20); ++guard) {
    CK(cudaMemsetAsync(BP.changed, 0, 64 * 4, ctx->stream));
    for (int r = 1; r <= 32; ++r) {  // round numbers > 0: walk only what changed
      BP.round = r;
      wg::bcp_state_kernel<<<cb, 128, 0, ctx->stream>>>(BP);
      ctx->launches++;
    }
    CK(cudaMemcpyAsync(changed, BP.changed, 64 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (changed[32] == 0) break;
  }
  *redone = true;  // the caller fetches out_size again
  return launch_bcp_tail(ctx, BP);
}

static wg::TokenParams token_params(wgpu_ctx* ctx) {
  wg::TokenParams T;
  T.hdr = ctx->hdr.as<uint8_t>(); T.coeffs = ctx->coeffs.as<int16_t>(); T.ctxw = ctx->ctxw.as<uint32_t>();
  T.proba = ctx->proba.as<uint8_t>(); T.mb_tokens = ctx->mb_tokens.as<uint32_t>();
  T.mb_offset = ctx->mb_offset.as<unsigned long long>(); T.img_total = ctx->img_total.as<unsigned long long>();
  T.img_base = ctx->img_base.as<unsigned long long>(); T.tokens = ctx->tokens.as<uint16_t>();
  T.n_images = ctx->e_n; T.mb_w = ctx->e_mbw; T.mb_h = ctx->e_mbh;
  return T;
}
static wg::P0Params p0_params(wgpu_ctx* ctx) {
  wg::P0Params Q;
  Q.hdr = ctx->hdr.as<uint8_t>(); Q.segment = ctx->segment.as<uint8_t>(); Q.proba = ctx->proba.as<uint8_t>();
  Q.proba0 = ctx->t_proba0.as<uint8_t>(); Q.update = ctx->t_upd.as<uint8_t>(); Q.bmodes = ctx->t_bmodes.as<uint8_t>();
  Q.i4paths = ctx->t_i4paths.as<wg::I4PathDev>(); Q.plan = ctx->p0_plan.as<wg::P0Plan>(); Q.info = ctx->p0_info.as<uint32_t>();
  Q.mb_tokens = ctx->p0_mb_tokens.as<uint32_t>(); Q.mb_offset = ctx->p0_mb_offset.as<unsigned long long>();
  Q.img_base = ctx->img_base.as<unsigned long long>() + ctx->e_n; Q.tokens = ctx->tokens.as<uint16_t>();
  Q.n_images = ctx->e_n; Q.mb_w = ctx->e_mbw; Q.mb_h = ctx->e_mbh;
  return Q;
}
// Single-partition route, device part 1 (queued right behind the waves): final probabilities from the token statistics,
// tokens per macroblock, per-image prefix sums and totals.
static int enc_launch_token_prepass(wgpu_ctx* ctx, bool run_optimize = true) {
  const int n = ctx->e_n, nmb = ctx->e_mbw * ctx->e_mbh;
  if (run_optimize) {
    wg::ProbaParams pp;
    pp.stats = ctx->stats.as<unsigned int>(); pp.proba0 = ctx->t_proba0.as<uint8_t>(); pp.update = ctx->t_upd.as<uint8_t>();
    pp.ecost = ctx->t_ecost.as<uint16_t>(); pp.proba = ctx->proba.as<uint8_t>(); pp.n_images = n;
    wg::optimize_proba_kernel<<<n, 352, 0, ctx->stream>>>(pp);
    ctx->launches++;
  }
  const wg::TokenParams T = token_params(ctx);
  const long long total = (long long)nmb * n;
  wg::token_kernel<false><<<(unsigned)((total + 15) / 16), 128, 0, ctx->stream>>>(T);
  wg::token_scan_kernel<<<n, 256, 0, ctx->stream>>>(T);
  ctx->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->h_totals.p, ctx->img_total.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += (uint64_t)((size_t)n * 8);
  // partition 0 as tokens (part0_kernels.cuh): counts now, so that their totals reach the host with the others
  if (ctx->plans.size() == (size_t)n) {
    RESERVE(ctx->p0_plan, (size_t)n * sizeof(wg::P0Plan)); RESERVE(ctx->p0_info, (size_t)n * 16);
    RESERVE(ctx->p0_mb_tokens, (size_t)n * nmb * 4); RESERVE(ctx->p0_mb_offset, (size_t)n * nmb * 8); RESERVE(ctx->p0_total, (size_t)n * 8);
    RESERVE(ctx->h_p0, (size_t)n * (sizeof(wg::P0Plan) + 16 + 8));
    wg::P0Plan* hp = ctx->h_p0.as<wg::P0Plan>();
    for (int i = 0; i < n; ++i) {
      const wgh::FramePlan& fp = ctx->plans[i];
      wg::P0Plan& d = hp[i];
      memset(&d, 0, sizeof(d));
      d.seg_use = fp.seg_use; d.seg_update_map = fp.seg_update_map; d.f_simple = fp.f_simple; d.f_level = (uint8_t)fp.f_level;
      d.f_sharpness = (uint8_t)fp.f_sharpness; d.parts_code = (uint8_t)(fp.num_parts == 8 ? 3 : fp.num_parts == 4 ? 2 : fp.num_parts == 2 ? 1 : 0);
      d.base_quant = (uint8_t)fp.seg[0].quant;
      for (int k = 0; k < 4; ++k) { d.seg_quantizer[k] = fp.seg_quantizer[k]; d.seg_fstrength[k] = fp.seg_fstrength[k]; }
      for (int k = 0; k < 3; ++k) d.seg_proba[k] = fp.seg_proba[k];
      d.dq_uv_dc = (int16_t)fp.dq_uv_dc; d.dq_uv_ac = (int16_t)fp.dq_uv_ac;
    }
    CK(cudaMemcpyAsync(ctx->p0_plan.p, hp, (size_t)n * sizeof(wg::P0Plan), cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n * sizeof(wg::P0Plan));
    if (ctx->e_opt.target_size > 0 || ctx->e_opt.target_psnr > 0.f) {
      // the rate-control loop re-derives the segment parameters (and with them the map the header codes) after its last pass:
      // partition 0 is written from the host's final map, as emit_partition0 would
      CK(cudaMemcpyAsync(ctx->segment.p, ctx->h_segment.p, (size_t)n * nmb, cudaMemcpyHostToDevice, ctx->stream));
      ctx->xfer_h2d += (uint64_t)((size_t)n * nmb);
    }
    const wg::P0Params Q = p0_params(ctx);
    wg::p0_info_kernel<<<n, 256, 0, ctx->stream>>>(Q);
    wg::p0_mb_kernel<false><<<dim3((unsigned)((nmb + 127) / 128), (unsigned)n), 128, 0, ctx->stream>>>(Q);
    wg::TokenParams S = T;  // the same scan over the partition-0 counts
    S.mb_tokens = ctx->p0_mb_tokens.as<uint32_t>(); S.mb_offset = ctx->p0_mb_offset.as<unsigned long long>();
    S.img_total = ctx->p0_total.as<unsigned long long>();
    wg::token_scan_kernel<<<n, 256, 0, ctx->stream>>>(S);
    ctx->launches += 3;
    CK(cudaGetLastError());
    uint8_t* hb = ctx->h_p0.as<uint8_t>() + (size_t)n * sizeof(wg::P0Plan);
    CK(cudaMemcpyAsync(hb, ctx->p0_info.p, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hb + (size_t)n * 16, ctx->p0_total.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)((size_t)n * 24);
  }
  return WGPU_OK;
}

// import + analysis, alphas back on the host (ctx->mu held by the caller)
static int enc_analyze_locked(wgpu_ctx* ctx, const wgpu_enc_options* opt) {
  if (!ctx->e_uploaded) FAIL(WGPU_ERR_INVALID, "encoder stage called before wgpu_enc_upload");
  int rc = validate_enc_options(ctx, opt, ctx->e_w, ctx->e_h);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->dev));
  ctx->e_opt = *opt;
  ctx->e_done = false;
  if ((rc = enc_reserve(ctx))) return rc;
  const int n = ctx->e_n, nmb = ctx->e_mbw * ctx->e_mbh;
  if ((rc = enc_launch_import(ctx))) return rc;
  if ((rc = enc_launch_analysis(ctx))) return rc;
  CK(cudaMemcpyAsync(ctx->h_alpha.p, ctx->alpha.p, (size_t)n * nmb, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += (uint64_t)((size_t)n * nmb);
  CK(cudaMemcpyAsync(ctx->h_uv_alpha.p, ctx->uv_alpha.p, (size_t)n * nmb, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += (uint64_t)((size_t)n * nmb);
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->e_analyzed = true;
  // per picture: the sum of its analysis alphas (low = busy), what the mode search orders its wave task lists by
  ctx->e_alpha_sum.assign(n, 0);
  parallel_for(n, threads_of(ctx), [&](int i) {
    const uint8_t* al = ctx->h_alpha.as<uint8_t>() + (size_t)i * nmb;
    long long a_sum = 0;
    for (int k = 0; k < nmb; ++k) a_sum += al[k];
    ctx->e_alpha_sum[i] = a_sum;
  });
  ctx->e_order_valid = true;  // until the next upload: the sums describe the pictures on the device
  return WGPU_OK;
}
// segment map + per-image parameters (already in the pinned staging buffers) -> device, then all waves
// Method < 3, one token partition: the final probabilities of the serial path without bringing the levels back.  The mode
// decisions of this path do not read the probabilities, so every pass of statLoop and the main pass encode the same
// macroblocks; what evolves is the probability state: refreshProbas at macroblocks k*M + (k-1) of the FIRST pass sees this
// pass's macroblocks above the refresh point and a fresh (zero-state) array below it, every later optimisation sees the whole
// frame and is idempotent (host_enc.h::serialize_frame_serial).  The statistics of each refresh point are taken on the GPU
// (collect_all_stats_kernel with a cut), the stateful optimizeProba chain is folded on the host (8 KB per image and point),
// and the result goes back up for token generation.
static int enc_fold_serial_probas(wgpu_ctx* ctx) {
  const int n = ctx->e_n, mbw = ctx->e_mbw, mbh = ctx->e_mbh, nmb = mbw * mbh;
  const int max_count = std::max(nmb >> 3, 96);
  std::vector<int> cuts;
  for (int k = 1; k * max_count + (k - 1) < nmb; ++k) cuts.push_back(k * max_count + (k - 1));
  const size_t K = cuts.size(), one = (size_t)n * wg::STATS_SIZE * 4;
  RESERVE(ctx->stats_cuts, std::max<size_t>(K, 1) * one);
  RESERVE(ctx->h_stats_cuts, std::max<size_t>(K, 1) * one);
  RESERVE(ctx->h_stats, one);
  for (size_t k = 0; k < K; ++k) {
    wg::AllStatsParams A;
    A.hdr = ctx->hdr.as<uint8_t>(); A.coeffs = ctx->coeffs.as<int16_t>();
    A.stats = ctx->stats_cuts.as<unsigned int>() + k * (size_t)n * wg::STATS_SIZE;
    A.n_images = n; A.mb_w = mbw; A.mb_h = mbh; A.cut = cuts[k]; A.hdr_prev = nullptr; A.coeffs_prev = nullptr;
    CK(cudaMemsetAsync(A.stats, 0, one, ctx->stream));
    wg::collect_all_stats_kernel<<<(unsigned)(((long long)n * nmb + 127) / 128), 128, 0, ctx->stream>>>(A);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  if (K) { CK(cudaMemcpyAsync(ctx->h_stats_cuts.p, ctx->stats_cuts.p, K * one, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)(K * one); }
  CK(cudaMemcpyAsync(ctx->h_stats.p, ctx->stats.p, one, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += (uint64_t)one;
  CK(cudaStreamSynchronize(ctx->stream));
  parallel_for(n, threads_of(ctx), [&](int i) {
    uint8_t (*pr)[8][3][11] = reinterpret_cast<uint8_t (*)[8][3][11]>(ctx->h_proba.as<uint8_t>() + (size_t)i * 1056);
    memcpy(pr, wgh::kCoeffsProba0, 1056);
    for (size_t k = 0; k < K; ++k)
      wgh::optimize_proba_host(*reinterpret_cast<const wgh::Stats*>(ctx->h_stats_cuts.as<uint32_t>() + (k * n + i) * wg::STATS_SIZE), pr);
    wgh::optimize_proba_host(*reinterpret_cast<const wgh::Stats*>(ctx->h_stats.as<uint32_t>() + (size_t)i * wg::STATS_SIZE), pr);
  });
  CK(cudaMemcpyAsync(ctx->proba.p, ctx->h_proba.p, (size_t)n * 1056, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)((size_t)n * 1056);
  return WGPU_OK;
}

// Rate control (doSearch: TargetSize / TargetPSNR), internal/lossy/encode.go:1338-1374 + adjustQuantForTarget / computeNextQ
// (:1505-1590): at least three serial encodeFrame passes; after each one the quality moves by the secant rule on the size of
// a trial frame (TargetSize) or on the reference's PSNR reading, which is 99.0 dB every time because MBEncInfo.Disto is
// never written (passes at Q, Q-10, Q-10).  Per image: the images of a batch converge independently, a
// converged image is parked (seg[0].flags bit 8) while the others go on.  setSegmentParams / buildSegmentHeader are re-run
// with the new quality -- also after the last pass, whose plan then no longer matches its levels, exactly as in the
// reference.  Error-diffusion state carries from pass to pass.  Requires the segment map and per-image parameters of the
// plan (wgpu_enc_device route).
static int enc_search_rate_control(wgpu_ctx* ctx) {
  const int n = ctx->e_n, nmb = ctx->e_mbw * ctx->e_mbh;
  if (ctx->plans.size() != (size_t)n) FAIL(WGPU_ERR_UNSUPPORTED, "TargetSize/TargetPSNR need the library's own segment plan (wgpu_enc_device), not wgpu_enc_search");
  struct RC { bool is_first, active; double dq, q, last_q, qmin, qmax, value, last_value, target; int quality; };
  const wgpu_enc_options base = ctx->e_opt;
  const bool do_size = base.target_size > 0;
  std::vector<RC> rcs(n);
  for (auto& r : rcs) {  // initPassStats (encode.go:1455)
    r.is_first = true; r.active = true; r.dq = 10.0;
    r.qmin = base.qmin; r.qmax = base.qmax <= 0 ? 100.0 : (double)base.qmax;
    r.q = std::min(std::max((double)base.quality, r.qmin), r.qmax);
    r.last_q = r.q; r.value = r.last_value = 0;
    r.target = do_size ? (double)base.target_size : (double)base.target_psnr;
    r.quality = base.quality;
  }
  int max_passes = std::max(base.passes, 1);
  if (max_passes < 3) max_passes = 3;
  if (do_size) { RESERVE(ctx->h_hdr, (size_t)n * nmb * 48); RESERVE(ctx->h_coeffs, (size_t)n * nmb * 800); }
  std::vector<uint32_t> zero_stats(wg::STATS_SIZE, 0);
  int rc = 0;
  // Method < 3: the decisions do not read the probabilities, so a pass is one run of the wavefront kernel and the refresh
  // schedule is replayed afterwards: the statistics of refresh point k are those of this pass above the point and of the
  // PREVIOUS pass below it (collect_all_stats_kernel with a cut and the previous arrays), folded statefully on the host.
  // statLoop (encode.go:1405) runs first at the starting quality; its first pass sees a zero-state array below the points.
  const bool fast = base.method < 3;
  const int mbw = ctx->e_mbw, mbh = ctx->e_mbh;
  std::vector<int> cuts;
  const size_t one = (size_t)n * wg::STATS_SIZE * 4;
  auto cut_stats = [&](bool with_prev) -> int {  // statistics of every refresh point -> h_stats_cuts (synchronises)
    for (size_t k = 0; k < cuts.size(); ++k) {
      wg::AllStatsParams A;
      A.hdr = ctx->hdr.as<uint8_t>(); A.coeffs = ctx->coeffs.as<int16_t>();
      A.stats = ctx->stats_cuts.as<unsigned int>() + k * (size_t)n * wg::STATS_SIZE;
      A.n_images = n; A.mb_w = mbw; A.mb_h = mbh; A.cut = cuts[k]; A.hdr_prev = nullptr; A.coeffs_prev = nullptr;
      A.hdr_prev = with_prev ? ctx->hdr_prev.as<uint8_t>() : nullptr; A.coeffs_prev = with_prev ? ctx->coeffs_prev.as<int16_t>() : nullptr;
      CK(cudaMemsetAsync(A.stats, 0, one, ctx->stream));
      wg::collect_all_stats_kernel<<<(unsigned)(((long long)n * nmb + 127) / 128), 128, 0, ctx->stream>>>(A);
      ctx->launches++;
    }
    CK(cudaGetLastError());
    if (!cuts.empty()) { CK(cudaMemcpyAsync(ctx->h_stats_cuts.p, ctx->stats_cuts.p, cuts.size() * one, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)(cuts.size() * one); }
    CK(cudaStreamSynchronize(ctx->stream));
    return WGPU_OK;
  };
  auto upload_plans = [&]() -> int {
    wg::ImageParams* hp = ctx->h_params.as<wg::ImageParams>();
    for (int i = 0; i < n; ++i) {
      memcpy(&hp[i], ctx->plans[i].dev, sizeof(wg::ImageParams));
      hp[i].seg[0].flags = (rcs[i].quality < 50 ? 2 : 3) | (rcs[i].active ? 0 : 0x100);  // getMaxI4RDModes follows the quality
      if (rcs[i].active) CK(cudaMemsetAsync(ctx->stats.as<uint32_t>() + (size_t)i * wg::STATS_SIZE, 0, wg::STATS_SIZE * 4, ctx->stream));
    }
    CK(cudaMemcpyAsync(ctx->segment.p, ctx->h_segment.p, (size_t)n * nmb, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n * nmb);
    CK(cudaMemcpyAsync(ctx->img_params.p, ctx->h_params.p, (size_t)n * sizeof(wg::ImageParams), cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n * sizeof(wg::ImageParams));
    return WGPU_OK;
  };
  if (fast) {
    const int max_count = std::max(nmb >> 3, 96);
    for (int k = 1; k * max_count + (k - 1) < nmb; ++k) cuts.push_back(k * max_count + (k - 1));
    RESERVE(ctx->stats_cuts, std::max<size_t>(cuts.size(), 1) * one); RESERVE(ctx->h_stats_cuts, std::max<size_t>(cuts.size(), 1) * one);
    RESERVE(ctx->h_stats, one);
    RESERVE(ctx->hdr_prev, (size_t)n * nmb * 48); RESERVE(ctx->coeffs_prev, (size_t)n * nmb * 800);
    RESERVE(ctx->h_hdr, (size_t)n * nmb * 48); RESERVE(ctx->h_coeffs, (size_t)n * nmb * 800);
    ctx->sp_proba.resize((size_t)n * 1056);
    ctx->sp_hist.assign(n, std::vector<uint8_t>());
    ctx->sp_starts.assign(1, 0);
    for (int c : cuts) ctx->sp_starts.push_back(c);
    // statLoop at the starting quality
    if ((rc = upload_plans())) return rc;
    ctx->e_keep_stats = true;
    rc = enc_launch_waves(ctx);
    ctx->e_keep_stats = false;
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_stats.p, ctx->stats.p, one, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)one;
    if ((rc = cut_stats(false))) return rc;
    parallel_for(n, threads_of(ctx), [&](int i) {
      uint8_t (*pr)[8][3][11] = reinterpret_cast<uint8_t (*)[8][3][11]>(&ctx->sp_proba[(size_t)i * 1056]);
      memcpy(pr, wgh::kCoeffsProba0, 1056);
      for (size_t k = 0; k < cuts.size(); ++k)
        wgh::optimize_proba_host(*reinterpret_cast<const wgh::Stats*>(ctx->h_stats_cuts.as<uint32_t>() + (k * n + i) * wg::STATS_SIZE), pr);
      wgh::optimize_proba_host(*reinterpret_cast<const wgh::Stats*>(ctx->h_stats.as<uint32_t>() + (size_t)i * wg::STATS_SIZE), pr);
    });
  }
  for (int pass = 0; pass < max_passes; ++pass) {
    if (fast) {
      if (pass > 0) {  // pass 0 repeats statLoop's arrays (same quality, same decisions): nothing to run, its refreshes are idempotent
        CK(cudaMemcpyAsync(ctx->hdr_prev.p, ctx->hdr.p, (size_t)n * nmb * 48, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->coeffs_prev.p, ctx->coeffs.p, (size_t)n * nmb * 800, cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc = upload_plans())) return rc;
        ctx->e_keep_stats = true;
        rc = enc_launch_waves(ctx);
        ctx->e_keep_stats = false;
        if (rc) return rc;
        if ((rc = cut_stats(true))) return rc;
      }
      // probability state through this pass's refreshes, and the table each stretch of macroblocks was recorded under
      parallel_for(n, threads_of(ctx), [&](int i) {
        if (!rcs[i].active) return;
        uint8_t (*pr)[8][3][11] = reinterpret_cast<uint8_t (*)[8][3][11]>(&ctx->sp_proba[(size_t)i * 1056]);
        std::vector<uint8_t>& hist = ctx->sp_hist[i];
        hist.assign(&ctx->sp_proba[(size_t)i * 1056], &ctx->sp_proba[(size_t)i * 1056] + 1056);
        for (size_t k = 0; k < cuts.size(); ++k) {
          if (pass > 0)
            wgh::optimize_proba_host(*reinterpret_cast<const wgh::Stats*>(ctx->h_stats_cuts.as<uint32_t>() + (k * n + i) * wg::STATS_SIZE), pr);
          hist.insert(hist.end(), &ctx->sp_proba[(size_t)i * 1056], &ctx->sp_proba[(size_t)i * 1056] + 1056);
        }
      });
    } else {
      if ((rc = upload_plans())) return rc;
      ctx->e_keep_derr = pass > 0; ctx->e_keep_stats = true;
      rc = enc_launch_waves(ctx);
      ctx->e_keep_derr = ctx->e_keep_stats = false;
      if (rc) return rc;
    }
    const bool tables_route = fast || ctx->e_refresh_route;
    if (do_size) {  // trial frame of every image still searching: emitFrame with the probabilities as they stand (defaults)
      CK(cudaMemcpyAsync(ctx->h_hdr.p, ctx->hdr.p, (size_t)n * nmb * 48, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->xfer_d2h += (uint64_t)((size_t)n * nmb * 48);
      CK(cudaMemcpyAsync(ctx->h_coeffs.p, ctx->coeffs.p, (size_t)n * nmb * 800, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->xfer_d2h += (uint64_t)((size_t)n * nmb * 800);
    }
    CK(cudaStreamSynchronize(ctx->stream));  // the plan below is read by the kernels of this pass until here
    parallel_for(n, threads_of(ctx), [&](int i) {
      RC& r = rcs[i];
      if (!r.active) return;
      wgh::FramePlan& fp = ctx->plans[i];
      if (do_size) {
        std::vector<uint8_t> riff;
        if (tables_route) {  // tokens as recorded during the pass: each stretch under the table then in force
          const int nt = (int)ctx->sp_starts.size();
          std::vector<const uint8_t*> tabs(nt);
          for (int k = 0; k < nt; ++k) tabs[k] = &ctx->sp_hist[i][(size_t)k * 1056];
          wgh::serialize_frame_tables(fp, ctx->h_hdr.as<uint8_t>() + (size_t)i * nmb * 48, ctx->h_coeffs.as<int16_t>() + (size_t)i * nmb * 400,
                                      ctx->h_segment.as<uint8_t>() + (size_t)i * nmb, &ctx->sp_proba[(size_t)i * 1056], nt, ctx->sp_starts.data(),
                                      tabs.data(), &riff);
        } else {
          wgh::serialize_frame(fp, ctx->h_hdr.as<uint8_t>() + (size_t)i * nmb * 48, ctx->h_coeffs.as<int16_t>() + (size_t)i * nmb * 400,
                               ctx->h_segment.as<uint8_t>() + (size_t)i * nmb, zero_stats.data(), &riff);
        }
        r.value = (double)((uint32_t)riff[16] | ((uint32_t)riff[17] << 8) | ((uint32_t)riff[18] << 16) | ((uint32_t)riff[19] << 24));  // VP8 payload
      } else {
        r.value = 99.0;
      }
      if (std::fabs(r.dq) <= 0.4 && !r.is_first) { r.active = false; return; }  // converged (DQ_LIMIT)
      double dq;  // computeNextQ (encode.go:1505)
      if (r.is_first) { dq = r.value > r.target ? -r.dq : r.dq; r.is_first = false; }
      else if (r.value != r.last_value) dq = (r.target - r.value) / (r.last_value - r.value) * (r.last_q - r.q);
      else dq = 0;
      dq = std::min(30.0, std::max(-30.0, dq));
      r.dq = dq; r.last_q = r.q; r.last_value = r.value;
      r.q = std::min(r.qmax, std::max(r.qmin, r.q + dq));
      r.quality = (int)(r.q + 0.5);
      wgpu_enc_options o = base;
      o.quality = r.quality;
      wgh::set_segment_params(&fp, o, fp.num_segments, ctx->h_segment.as<uint8_t>() + (size_t)i * nmb);
      wgh::build_segment_header(&fp, o, fp.num_segments);
    });
    bool any = false;
    for (const auto& r : rcs) any |= r.active;
    if (!any) break;
  }
  if (fast) ctx->e_refresh_route = true;  // wgpu_enc_finish: final optimizeProba on this state, tokens per table or re-recorded
  return WGPU_OK;
}
static int enc_search_locked(wgpu_ctx* ctx) {
  const int n = ctx->e_n, nmb = ctx->e_mbw * ctx->e_mbh;
  CK(cudaMemcpyAsync(ctx->segment.p, ctx->h_segment.p, (size_t)n * nmb, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)((size_t)n * nmb);
  CK(cudaMemcpyAsync(ctx->img_params.p, ctx->h_params.p, (size_t)n * sizeof(wg::ImageParams), cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)((size_t)n * sizeof(wg::ImageParams));
  int rc;
  if (ctx->e_opt.target_size > 0 || ctx->e_opt.target_psnr > 0.f) {
    if ((rc = enc_search_rate_control(ctx))) return rc;
  } else if ((rc = enc_launch_waves(ctx))) {
    return rc;
  }
  ctx->e_token_route = false;
  if (ctx->e_opt.partitions == 0 && !ctx->e_refresh_route) {
    if (ctx->e_opt.method >= 3) {
      if ((rc = enc_launch_token_prepass(ctx))) return rc;
      ctx->e_token_route = true;
    } else if (ctx->plans.size() == (size_t)ctx->e_n && getenv_int("WGPU_FAST_TOKEN_ROUTE", 1)) {
      if ((rc = enc_fold_serial_probas(ctx))) return rc;
      if ((rc = enc_launch_token_prepass(ctx, false))) return rc;
      ctx->e_token_route = true;
    }
  }
  ctx->e_done = true;
  return WGPU_OK;
}

int wgpu_setup_segment(int quant_index, int dq_uv_dc, int dq_uv_ac, int method, int sns_strength, wgpu_segment* out) {
  if (!out || quant_index < 0 || quant_index > 127) return WGPU_ERR_INVALID;
  wgh::FramePlan fp;
  memset(&fp, 0, sizeof(fp));
  fp.seg[0].quant = quant_index;
  fp.dq_uv_dc = dq_uv_dc; fp.dq_uv_ac = dq_uv_ac;
  wgpu_enc_options o;
  wgpu_enc_options_default(&o, 75);
  o.method = method; o.sns_strength = sns_strength;
  wgh::setup_segment(&fp, o, 0);
  memcpy(out, &fp.dev[0], sizeof(*out));
  return WGPU_OK;
}

int wgpu_enc_analyze(wgpu_ctx* ctx, const wgpu_enc_options* opt, uint8_t* alphas, int64_t* uv_alpha_sum) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  int rc = enc_analyze_locked(ctx, opt);
  if (rc) return rc;
  const int n = ctx->e_n, nmb = ctx->e_mbw * ctx->e_mbh;
  if (alphas) memcpy(alphas, ctx->h_alpha.p, (size_t)n * nmb);
  if (uv_alpha_sum)
    for (int i = 0; i < n; ++i) {
      int64_t s = 0;
      const uint8_t* ua = ctx->h_uv_alpha.as<uint8_t>() + (size_t)i * nmb;
      for (int k = 0; k < nmb; ++k) s += ua[k];
      uv_alpha_sum[i] = s;
    }
  return WGPU_OK;
}

int wgpu_enc_search(wgpu_ctx* ctx, const wgpu_segment* segments, const uint8_t* segment_map) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->e_analyzed) FAIL(WGPU_ERR_INVALID, "wgpu_enc_search called before wgpu_enc_analyze");
  if (!segments || !segment_map) FAIL(WGPU_ERR_INVALID, "wgpu_enc_search: nil segment parameters");
  static_assert(sizeof(wgpu_segment) == sizeof(wg::SegParams), "wgpu_segment layout");
  const size_t n = ctx->e_n, nmb = (size_t)ctx->e_mbw * ctx->e_mbh;
  for (size_t i = 0; i < n * nmb; ++i)
    if (segment_map[i] > 3) FAIL(WGPU_ERR_INVALID, "wgpu_enc_search: segment id out of range");
  for (size_t i = 0; i < n * 4; ++i)
    if (segments[i].y1.quant <= 0 || segments[i].y1.dc_quant <= 0 || segments[i].y2.quant <= 0 || segments[i].y2.dc_quant <= 0 ||
        segments[i].uv.quant <= 0 || segments[i].uv.dc_quant <= 0)
      FAIL(WGPU_ERR_INVALID, "wgpu_enc_search: non-positive quantiser");
  CK(cudaSetDevice(ctx->dev));
  memcpy(ctx->h_segment.p, segment_map, n * nmb);
  memcpy(ctx->h_params.p, segments, n * sizeof(wg::ImageParams));
  ctx->plans.clear();  // the host owns segmentation on this route: wgpu_enc_finish is not available, use wgpu_enc_fetch
  return enc_search_locked(ctx);
}

int wgpu_enc_device(wgpu_ctx* ctx, const wgpu_enc_options* opt) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  const double t0 = now_ms();
  int rc = enc_analyze_locked(ctx, opt);
  if (rc) return rc;
  const double t1 = now_ms();
  const int n = ctx->e_n, nmb = ctx->e_mbw * ctx->e_mbh;
  // host: segment clustering + quantiser / lambda setup (microseconds per image; float64 pow as in the reference)
  ctx->plans.resize(n);
  static_assert(sizeof(wgh::SegParams) == sizeof(wg::SegParams), "SegParams layout");
  parallel_for(n, threads_of(ctx), [&](int i) {
    const uint8_t* ua = ctx->h_uv_alpha.as<uint8_t>() + (size_t)i * nmb;
    long long uv_sum = 0;
    for (int k = 0; k < nmb; ++k) uv_sum += ua[k];
    wgh::FramePlan& fp = ctx->plans[i];
    wgh::plan_frame(&fp, ctx->e_opt, ctx->e_w, ctx->e_h, ctx->h_alpha.as<uint8_t>() + (size_t)i * nmb, uv_sum,
                    ctx->h_segment.as<uint8_t>() + (size_t)i * nmb);
    fp.num_parts = 1 << ctx->e_opt.partitions;
    memcpy(ctx->h_params.as<uint8_t>() + (size_t)i * sizeof(wg::ImageParams), fp.dev, sizeof(wg::ImageParams));
  });
  const double t2 = now_ms();
  if ((rc = enc_search_locked(ctx))) return rc;
  if (trace_on()) {
    const double t3 = now_ms();
    cudaStreamSynchronize(ctx->stream);
    fprintf(stderr, "[wgpu] enc_device: import+analysis+D2H %.2f ms, host plan %.2f ms, wave launch %.2f ms, waves done after %.2f ms\n", t1 - t0,
            t2 - t1, t3 - t2, now_ms() - t2);
  }
  return WGPU_OK;
}

int wgpu_enc_finish(wgpu_ctx* ctx, uint8_t* out, size_t out_stride, size_t* out_sizes) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->e_done || ctx->plans.size() != (size_t)ctx->e_n) FAIL(WGPU_ERR_INVALID, "wgpu_enc_finish called before wgpu_enc_device");
  if (!out || !out_sizes) FAIL(WGPU_ERR_INVALID, "webp: nil writer");
  CK(cudaSetDevice(ctx->dev));
  const size_t n = ctx->e_n, nmb = (size_t)ctx->e_mbw * ctx->e_mbh;
  RESERVE(ctx->h_hdr, n * nmb * 48);
  std::atomic<int> too_small(0);
  const double t0 = now_ms();
  if (ctx->e_refresh_route) {
    // ---- serial RD path with probability refreshes: final optimizeProba on top of the state the refreshes left (a slot that
    // does not beat the default keeps what an earlier refresh wrote), then all tokens under the final table -- or, if
    // nothing changed, as recorded during the pass (encode.go:1376-1383); levels come back, the host serialises
    RESERVE(ctx->h_coeffs, n * nmb * 800);
    CK(cudaMemcpyAsync(ctx->h_hdr.p, ctx->hdr.p, n * nmb * 48, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)(n * nmb * 48);
    CK(cudaMemcpyAsync(ctx->h_coeffs.p, ctx->coeffs.p, n * nmb * 800, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)(n * nmb * 800);
    CK(cudaMemcpyAsync(ctx->h_stats.p, ctx->stats.p, n * wg::STATS_SIZE * 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)(n * wg::STATS_SIZE * 4);
    CK(cudaStreamSynchronize(ctx->stream));
    parallel_for((int)n, threads_of(ctx), [&](int i) {
      uint8_t final_proba[1056];
      memcpy(final_proba, &ctx->sp_proba[(size_t)i * 1056], 1056);
      const int updates = wgh::optimize_proba_host(*reinterpret_cast<const wgh::Stats*>(ctx->h_stats.as<uint32_t>() + (size_t)i * wg::STATS_SIZE),
                                                   *reinterpret_cast<uint8_t (*)[4][8][3][11]>(final_proba));
      const int nt = updates > 0 ? 1 : (int)ctx->sp_starts.size();
      std::vector<const uint8_t*> tabs(nt);
      const int zero = 0;
      if (updates > 0) tabs[0] = final_proba;
      else for (int k = 0; k < nt; ++k) tabs[k] = &ctx->sp_hist[i][(size_t)k * 1056];
      std::vector<uint8_t> riff;
      riff.reserve(nmb * 64 + 4096);
      wgh::serialize_frame_tables(ctx->plans[i], ctx->h_hdr.as<uint8_t>() + (size_t)i * nmb * 48, ctx->h_coeffs.as<int16_t>() + (size_t)i * nmb * 400,
                                  ctx->h_segment.as<uint8_t>() + (size_t)i * nmb, final_proba, nt, updates > 0 ? &zero : ctx->sp_starts.data(),
                                  tabs.data(), &riff);
      out_sizes[i] = riff.size();
      if (riff.size() > out_stride) { too_small.store(1); return; }
      memcpy(out + (size_t)i * out_stride, riff.data(), riff.size());
    });
  } else if (ctx->e_token_route && device_coder_wanted(ctx, n)) {
    // ---- single token partition: the tokens of BOTH partitions of a frame are generated and boolean-coded on the GPU
    // (token_kernel, part0_kernels.cuh, boolcode_par.cuh); the host only lays the frames out around the coded partitions
    CK(cudaStreamSynchronize(ctx->stream));  // waves + token pre-pass done, per-image totals on the host
    const double t1 = now_ms();
    const unsigned long long* main_totals = ctx->h_totals.as<unsigned long long>();
    const uint32_t* p0_info = reinterpret_cast<const uint32_t*>(ctx->h_p0.as<uint8_t>() + n * sizeof(wg::P0Plan));
    const unsigned long long* p0_mb_total = reinterpret_cast<const unsigned long long*>(ctx->h_p0.as<uint8_t>() + n * (sizeof(wg::P0Plan) + 16));
    // coder partitions: [0, n) the token partition of image i, [n, 2n) partition 0 of image i - n
    unsigned long long* bases = ctx->h_bases.as<unsigned long long>();  // [0, 2n) token offsets, [2n, 4n) coded byte offsets, [4n, 6n) totals
    unsigned long long* totals = bases + 4 * n;
    unsigned long long all = 0, oall = 0;
    for (size_t k = 0; k < 2 * n; ++k) {
      totals[k] = k < n ? main_totals[k] : (unsigned long long)p0_info[(k - n) * 4 + 2] + p0_mb_total[k - n];
      bases[k] = all; all += (totals[k] + 7) & ~7ull;                        // 128-bit aligned token runs
      bases[2 * n + k] = oall; oall += (totals[k] + 16 + 15) & ~15ull;        // <= 7 bits out per token + the closing flush
    }
    RESERVE(ctx->tokens, (size_t)(all + 512) * 2);
    RESERVE(ctx->coded, (size_t)oall);
    RESERVE(ctx->coded_size, 2 * n * 4); RESERVE(ctx->h_coded_size, 2 * n * 4);
    CK(cudaMemcpyAsync(ctx->img_base.p, bases, 6 * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)(6 * n * 8);
    const wg::TokenParams T = token_params(ctx);
    wg::token_kernel<true><<<(unsigned)((n * nmb + 15) / 16), 128, 0, ctx->stream>>>(T);
    const wg::P0Params Q = p0_params(ctx);
    wg::p0_mb_kernel<true><<<dim3((unsigned)((nmb + 127) / 128), (unsigned)n), 128, 0, ctx->stream>>>(Q);
    ctx->launches += 2;
    wg::BoolCodeParams B;
    B.tokens = ctx->tokens.as<uint16_t>(); B.img_base = ctx->img_base.as<unsigned long long>(); B.img_total = ctx->img_base.as<unsigned long long>() + 4 * n;
    B.out = ctx->coded.as<uint8_t>(); B.out_base = ctx->img_base.as<unsigned long long>() + 2 * n; B.out_size = ctx->coded_size.as<unsigned int>();
    B.n_images = (int)(2 * n);
    wg::BcpParams BP;
    {
      const int rc_par = launch_boolcode_par(ctx, B, totals, 2 * n, &BP);
      if (rc_par) return rc_par;
    }
    CK(cudaGetLastError());
    // frames packed back to back on the device ([partition 0][token partition] per image): one copy brings the batch over
    RESERVE(ctx->coded_packed, (size_t)oall); RESERVE(ctx->pack_offsets, n * 8);
    wg::FramePackParams FP;
    FP.coded = ctx->coded.as<uint8_t>(); FP.in_base = ctx->img_base.as<unsigned long long>() + 2 * n; FP.sizes = ctx->coded_size.as<unsigned int>();
    FP.packed = ctx->coded_packed.as<uint8_t>(); FP.offsets = ctx->pack_offsets.as<unsigned long long>(); FP.n = (int)n;
    unsigned long long max_cap = 0;
    for (size_t i = 0; i < n; ++i) max_cap = std::max(max_cap, totals[i] + totals[n + i] + 64);
    auto launch_pack = [&]() -> int {
      wg::frame_pack_scan_kernel<<<1, 1024, 0, ctx->stream>>>(FP);
      wg::frame_pack_copy_kernel<<<dim3((unsigned)n, (unsigned)((max_cap + 16383) / 16384)), 256, 0, ctx->stream>>>(FP);
      ctx->launches += 2;
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(ctx->h_coded_size.p, ctx->coded_size.p, 2 * n * 4, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->xfer_d2h += (uint64_t)(2 * n * 4);
      return WGPU_OK;
    };
    if (int rc_pack = launch_pack()) return rc_pack;
    CK(cudaStreamSynchronize(ctx->stream));
    {  // more relaxation rounds if the usual ones were not enough
      bool redone = false;
      const int rc_par = finish_boolcode_par(ctx, BP, &redone);
      if (rc_par) return rc_par;
      if (redone) {
        if (int rc_pack = launch_pack()) return rc_pack;
        CK(cudaStreamSynchronize(ctx->stream));
      }
    }
    const double t4 = now_ms();
    const unsigned int* csz = ctx->h_coded_size.as<unsigned int>();
    std::vector<unsigned long long> offs(n + 1);
    offs[0] = 0;
    for (size_t i = 0; i < n; ++i) offs[i + 1] = offs[i] + csz[i] + csz[n + i];  // the same prefix frame_pack_scan_kernel took
    RESERVE(ctx->h_packed, (size_t)offs[n] + 16);
    CK(cudaMemcpyAsync(ctx->h_packed.p, ctx->coded_packed.p, (size_t)offs[n], cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)offs[n];
    CK(cudaStreamSynchronize(ctx->stream));
    parallel_for((int)n, threads_of(ctx), [&](int i) {
      out_sizes[i] = wgh::frame_file_size(csz[n + i], csz[i]);
      if (out_sizes[i] > out_stride) { too_small.store(1); return; }
      uint8_t* dst = out + (size_t)i * out_stride;
      memcpy(dst + 30, ctx->h_packed.as<uint8_t>() + offs[i], (size_t)(offs[i + 1] - offs[i]));
      wgh::write_frame_headers(ctx->plans[i], dst, csz[n + i], csz[i]);
    });
    if (trace_on())
      fprintf(stderr, "[wgpu] enc_finish: wait device %.2f ms, tokens + coder + pack %.2f ms (%.1f M tokens, longest partition %.2f M), frames D2H + layout %.2f ms\n",
              t1 - t0, t4 - t1, (double)all / 1e6, (double)*std::max_element(totals, totals + 2 * n) / 1e6, now_ms() - t4);
  } else if (ctx->e_token_route) {
    // ---- single partition: tokens are generated on the GPU, the host only boolean-codes flat arrays
    CK(cudaStreamSynchronize(ctx->stream));  // waves + token pre-pass done, per-image totals on the host
    const double t1 = now_ms();
    const unsigned long long* totals = ctx->h_totals.as<unsigned long long>();
    unsigned long long* bases = ctx->h_bases.as<unsigned long long>();
    unsigned long long all = 0;
    for (size_t i = 0; i < n; ++i) { bases[i] = all; all += totals[i]; }
    RESERVE(ctx->tokens, (size_t)(all + 8) * 2);
    RESERVE(ctx->h_tokens, (size_t)(all + 8) * 2);
    CK(cudaMemcpyAsync(ctx->img_base.p, bases, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)(n * 8);
    const wg::TokenParams T = token_params(ctx);
    wg::token_kernel<true><<<(unsigned)((n * nmb + 15) / 16), 128, 0, ctx->stream>>>(T);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_hdr.p, ctx->hdr.p, n * nmb * 48, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)(n * nmb * 48);
    CK(cudaMemcpyAsync(ctx->h_proba.p, ctx->proba.p, n * 1056, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)(n * 1056);
    if (all) { CK(cudaMemcpyAsync(ctx->h_tokens.p, ctx->tokens.p, (size_t)all * 2, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)((size_t)all * 2); }
    CK(cudaStreamSynchronize(ctx->stream));
    const double t2 = now_ms();
    // pair images of similar length (longest first) so two coders run interleaved in each task
    std::vector<int> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return totals[a] > totals[b] || (totals[a] == totals[b] && a < b); });
    const int pairs = (int)((n + 1) / 2);
    std::vector<std::vector<uint8_t>> coded(n);
    parallel_for(pairs, threads_of(ctx), [&](int p) {
      const int ia = order[2 * p], ib = (size_t)(2 * p + 1) < n ? order[2 * p + 1] : -1;
      coded[ia].reserve((size_t)totals[ia] / 4 + 4096);
      if (ib >= 0) coded[ib].reserve((size_t)totals[ib] / 4 + 4096);
      wgh::code_token_streams(ctx->h_tokens.as<uint16_t>() + bases[ia], (size_t)totals[ia], &coded[ia],
                              ib >= 0 ? ctx->h_tokens.as<uint16_t>() + bases[ib] : nullptr, ib >= 0 ? (size_t)totals[ib] : 0,
                              ib >= 0 ? &coded[ib] : nullptr);
      for (int k = 0; k < 2; ++k) {
        const int i = k == 0 ? ia : ib;
        if (i < 0) continue;
        std::vector<uint8_t> riff;
        riff.reserve(coded[i].size() + nmb * 4 + 4096);
        wgh::assemble_frame_tokens(ctx->plans[i], ctx->h_hdr.as<uint8_t>() + (size_t)i * nmb * 48, ctx->h_segment.as<uint8_t>() + (size_t)i * nmb,
                                   ctx->h_proba.as<uint8_t>() + (size_t)i * 1056, coded[i], &riff);
        out_sizes[i] = riff.size();
        if (riff.size() > out_stride) { too_small.store(1); continue; }
        memcpy(out + (size_t)i * out_stride, riff.data(), riff.size());
        std::vector<uint8_t>().swap(coded[i]);
      }
    });
    if (trace_on())
      fprintf(stderr, "[wgpu] enc_finish: wait device %.2f ms, token emit + D2H %.2f ms (%.1f MB tokens), host code %.2f ms (%d threads)\n", t1 - t0,
              t2 - t1, (double)all * 2 / 1e6, now_ms() - t2, threads_of(ctx));
  } else {
    // ---- multi-partition and Method < 3: levels + statistics come back, the host walks them
    RESERVE(ctx->h_coeffs, n * nmb * 800);
    CK(cudaMemcpyAsync(ctx->h_hdr.p, ctx->hdr.p, n * nmb * 48, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)(n * nmb * 48);
    CK(cudaMemcpyAsync(ctx->h_coeffs.p, ctx->coeffs.p, n * nmb * 800, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)(n * nmb * 800);
    CK(cudaMemcpyAsync(ctx->h_stats.p, ctx->stats.p, n * wg::STATS_SIZE * 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)(n * wg::STATS_SIZE * 4);
    CK(cudaStreamSynchronize(ctx->stream));
    const double t1 = now_ms();
    parallel_for((int)n, threads_of(ctx), [&](int i) {
      std::vector<uint8_t> riff;
      riff.reserve(nmb * 64 + 4096);
      if (ctx->e_opt.method < 3)  // serial-path semantics: probability refreshes + inline token recording restated on the host
        wgh::serialize_frame_serial(ctx->plans[i], ctx->h_hdr.as<uint8_t>() + (size_t)i * nmb * 48, ctx->h_coeffs.as<int16_t>() + (size_t)i * nmb * 400,
                                    ctx->h_segment.as<uint8_t>() + (size_t)i * nmb, ctx->h_stats.as<uint32_t>() + (size_t)i * wg::STATS_SIZE,
                                    ctx->e_opt.passes, &riff);
      else
        wgh::serialize_frame(ctx->plans[i], ctx->h_hdr.as<uint8_t>() + (size_t)i * nmb * 48, ctx->h_coeffs.as<int16_t>() + (size_t)i * nmb * 400,
                             ctx->h_segment.as<uint8_t>() + (size_t)i * nmb, ctx->h_stats.as<uint32_t>() + (size_t)i * wg::STATS_SIZE, &riff);
      out_sizes[i] = riff.size();
      if (riff.size() > out_stride) { too_small.store(1); return; }
      memcpy(out + (size_t)i * out_stride, riff.data(), riff.size());
    });
    if (trace_on())
      fprintf(stderr, "[wgpu] enc_finish: D2H %.2f ms (%.1f MB), host serialise %.2f ms (%d threads)\n", t1 - t0, (double)(n * nmb * 848) / 1e6,
              now_ms() - t1, threads_of(ctx));
  }
  if (too_small.load()) FAIL(WGPU_ERR_TOO_SMALL, "output buffer too small (out_sizes holds the required sizes)");
  return WGPU_OK;
}

int wgpu_encode_batch(wgpu_ctx* ctx, const uint8_t* rgba, int n, int width, int height, int stride, size_t image_stride,
                      const wgpu_enc_options* opt, uint8_t* out, size_t out_stride, size_t* out_sizes) {
  if (!ctx) return WGPU_ERR_INVALID;
  int rc = validate_enc_options(ctx, opt, width, height);
  if (rc) return rc;
  if ((rc = wgpu_enc_upload(ctx, rgba, n, width, height, stride, image_stride))) return rc;
  if ((rc = wgpu_enc_device(ctx, opt))) return rc;
  return wgpu_enc_finish(ctx, out, out_stride, out_sizes);
}

int wgpu_enc_stats(wgpu_ctx* ctx, int upto_mb, uint32_t* stats) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->e_done) FAIL(WGPU_ERR_INVALID, "wgpu_enc_stats needs a completed wgpu_enc_device / wgpu_enc_search");
  const int n = ctx->e_n, nmb = ctx->e_mbw * ctx->e_mbh;
  if (!stats || upto_mb < 0 || upto_mb > nmb) FAIL(WGPU_ERR_INVALID, "wgpu_enc_stats: bad arguments");
  CK(cudaSetDevice(ctx->dev));
  const size_t bytes = (size_t)n * wg::STATS_SIZE * 4;
  RESERVE(ctx->stats_cuts, bytes);
  wg::AllStatsParams A;
  A.hdr = ctx->hdr.as<uint8_t>(); A.coeffs = ctx->coeffs.as<int16_t>(); A.stats = ctx->stats_cuts.as<unsigned int>();
  A.n_images = n; A.mb_w = ctx->e_mbw; A.mb_h = ctx->e_mbh; A.cut = upto_mb; A.hdr_prev = nullptr; A.coeffs_prev = nullptr;
  CK(cudaMemsetAsync(A.stats, 0, bytes, ctx->stream));
  wg::collect_all_stats_kernel<<<(unsigned)(((long long)n * nmb + 127) / 128), 128, 0, ctx->stream>>>(A);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(stats, A.stats, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += (uint64_t)bytes;
  CK(cudaStreamSynchronize(ctx->stream));
  return WGPU_OK;
}

int wgpu_transfer_bytes(wgpu_ctx* ctx, uint64_t* h2d, uint64_t* d2h, int reset) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (h2d) *h2d = ctx->xfer_h2d;
  if (d2h) *d2h = ctx->xfer_d2h;
  if (reset) ctx->xfer_h2d = ctx->xfer_d2h = 0;
  return WGPU_OK;
}

int wgpu_enc_fetch(wgpu_ctx* ctx, int image, uint8_t* mb_hdr, uint8_t* mb_modes, uint8_t* mb_nz, int16_t* mb_coeffs, uint8_t* recon_y,
                   uint8_t* recon_u, uint8_t* recon_v, uint8_t* src_y, uint8_t* src_u, uint8_t* src_v, uint8_t* alphas) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->e_done || image < 0 || image >= ctx->e_n) FAIL(WGPU_ERR_INVALID, "wgpu_enc_fetch: no encoded batch / bad image index");
  CK(cudaSetDevice(ctx->dev));
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t nmb = (size_t)ctx->e_mbw * ctx->e_mbh, i = image;
  if (mb_hdr || mb_modes || mb_nz) {
    std::vector<uint8_t> h(nmb * 48);
    CK(cudaMemcpy(h.data(), ctx->hdr.as<uint8_t>() + i * nmb * 48, nmb * 48, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < nmb; ++k) {
      if (mb_hdr) memcpy(mb_hdr + 8 * k, &h[48 * k], 8);
      if (mb_modes) memcpy(mb_modes + 16 * k, &h[48 * k + 8], 16);
      if (mb_nz) memcpy(mb_nz + 24 * k, &h[48 * k + 24], 24);
    }
  }
  if (mb_coeffs) CK(cudaMemcpy(mb_coeffs, ctx->coeffs.as<int16_t>() + i * nmb * 400, nmb * 800, cudaMemcpyDeviceToHost));
  if (recon_y) CK(cudaMemcpy(recon_y, ctx->ry.as<uint8_t>() + i * nmb * 256, nmb * 256, cudaMemcpyDeviceToHost));
  if (recon_u) CK(cudaMemcpy(recon_u, ctx->ru.as<uint8_t>() + i * nmb * 64, nmb * 64, cudaMemcpyDeviceToHost));
  if (recon_v) CK(cudaMemcpy(recon_v, ctx->rv.as<uint8_t>() + i * nmb * 64, nmb * 64, cudaMemcpyDeviceToHost));
  if (src_y) CK(cudaMemcpy(src_y, ctx->sy.as<uint8_t>() + i * nmb * 256, nmb * 256, cudaMemcpyDeviceToHost));
  if (src_u) CK(cudaMemcpy(src_u, ctx->su.as<uint8_t>() + i * nmb * 64, nmb * 64, cudaMemcpyDeviceToHost));
  if (src_v) CK(cudaMemcpy(src_v, ctx->sv.as<uint8_t>() + i * nmb * 64, nmb * 64, cudaMemcpyDeviceToHost));
  if (alphas) CK(cudaMemcpy(alphas, ctx->alpha.as<uint8_t>() + i * nmb, nmb, cudaMemcpyDeviceToHost));
  return WGPU_OK;
}

static int launch_metrics(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int width, int height, int stride, size_t plane_stride, bool want_ssim);
static int launch_upsample(wgpu_ctx* ctx, int n, int width, int height, const uint8_t* y, int ys, const uint8_t* u, const uint8_t* v,
                           int uvs, size_t y_plane, size_t uv_plane, const uint8_t* alpha, uint8_t* out);
int wgpu_enc_stage_time(wgpu_ctx* ctx, const wgpu_enc_options* opt, int stage, int reps, float* ms_per_rep) {
  if (!ctx || !ms_per_rep || reps <= 0) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if ((stage <= 3 || stage == 5) && !ctx->e_done) FAIL(WGPU_ERR_INVALID, "wgpu_enc_stage_time needs a completed wgpu_enc_device");
  if (stage == 4 && !ctx->d_ready) FAIL(WGPU_ERR_INVALID, "stage 4 (upsampling) needs a decoded batch (wgpu_dec_parse + wgpu_dec_device)");
  (void)opt;
  CK(cudaSetDevice(ctx->dev));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = 0;
  for (int r = 0; r < reps && !rc; ++r) {
    if (stage == 0) rc = enc_launch_import(ctx);
    else if (stage == 1) rc = enc_launch_analysis(ctx);
    else if (stage == 2) rc = enc_launch_waves(ctx);
    else if (stage == 3)  // SSE + SSIM of the source luma planes against the reconstruction, both resident from the encode
      rc = launch_metrics(ctx, ctx->e_n, ctx->sy.as<uint8_t>(), ctx->ry.as<uint8_t>(), ctx->e_mbw * 16, ctx->e_mbh * 16, ctx->e_mbw * 16, (size_t)ctx->e_mbw * ctx->e_mbh * 256, true);
    else if (stage == 5)  // SSE alone (PSNR): the streaming form
      rc = launch_metrics(ctx, ctx->e_n, ctx->sy.as<uint8_t>(), ctx->ry.as<uint8_t>(), ctx->e_mbw * 16, ctx->e_mbh * 16, ctx->e_mbw * 16, (size_t)ctx->e_mbw * ctx->e_mbh * 256, false);
    else if (stage == 4) {
      RESERVE(ctx->d_nrgba, (size_t)ctx->d_n * ctx->d_w * ctx->d_h * 4);
      rc = launch_upsample(ctx, ctx->d_n, ctx->d_w, ctx->d_h, ctx->dy.as<uint8_t>(), ctx->d_mbw * 16, ctx->du.as<uint8_t>(), ctx->dv.as<uint8_t>(), ctx->d_mbw * 8,
                           (size_t)ctx->d_mbw * ctx->d_mbh * 256, (size_t)ctx->d_mbw * ctx->d_mbh * 64, nullptr, ctx->d_nrgba.as<uint8_t>());
    } else FAIL(WGPU_ERR_INVALID, "unknown stage");
  }
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_per_rep = ms / reps;
  return WGPU_OK;
}

// ======================================================================================== decoder
int wgpu_decode_info(const uint8_t* data, size_t len, int* width, int* height) {
  if (!data || !width || !height) return WGPU_ERR_INVALID;
  const uint8_t* vp8; size_t vlen; const char* err = nullptr;
  const int fr = wgh::find_vp8_ex(data, len, &vp8, &vlen);
  if (fr < 0) return WGPU_ERR_UNSUPPORTED;
  if (!fr) return WGPU_ERR_BITSTREAM;
  if (!wgh::peek_dims(vp8, vlen, width, height, &err)) return WGPU_ERR_BITSTREAM;
  return WGPU_OK;
}

namespace {
constexpr int kDecG = 8, kDecWarps = 4, kFiltWarps = 4;
}

static int dec_launch_recon(wgpu_ctx* ctx, const wg::DecKernelParams& P) {
  const int waves = P.mb_w + 2 * (P.mb_h - 1);
  const int per_cta = kDecWarps * (32 / kDecG);
  for (int w = 0; w < waves; ++w) {
    const long long tasks = (long long)wave_rows(w, P.mb_w, P.mb_h) * P.n_images;
    if (tasks <= 0) continue;  // one macroblock column: odd waves hold no macroblock (x = w - 2y)
    wg::recon_wave_kernel<kDecG, kDecWarps><<<(unsigned)((tasks + per_cta - 1) / per_cta), kDecWarps * 32, 0, ctx->stream>>>(P, w);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  return WGPU_OK;
}
static int dec_launch_filter(wgpu_ctx* ctx, const wg::DecKernelParams& P) {
  const int waves = P.mb_w + 2 * (P.mb_h - 1);
  const int per_cta = kFiltWarps * 2;
  for (int w = 0; w < waves; ++w) {
    const long long tasks = (long long)wave_rows(w, P.mb_w, P.mb_h) * P.n_images;
    if (tasks <= 0) continue;  // one macroblock column: odd waves hold no macroblock (x = w - 2y)
    wg::filter_wave_kernel<kFiltWarps><<<(unsigned)((tasks + per_cta - 1) / per_cta), kFiltWarps * 32, 0, ctx->stream>>>(P, w);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  return WGPU_OK;
}
static int launch_upsample(wgpu_ctx* ctx, int n, int width, int height, const uint8_t* y, int ys, const uint8_t* u, const uint8_t* v,
                           int uvs, size_t y_plane, size_t uv_plane, const uint8_t* alpha, uint8_t* out) {
  wg::UpsampleParams up;
  up.y = y; up.u = u; up.v = v; up.alpha = alpha;
  up.y_plane = y_plane; up.uv_plane = uv_plane; up.alpha_plane = (size_t)width * height; up.out_image = (size_t)width * height * 4;
  up.y_stride = ys; up.uv_stride = uvs; up.width = width; up.height = height; up.n = n; up.out = out;
  // 16 pixels of a line pair per thread
  const int gw = (width + 15) / 16, bx = gw <= 8 ? 8 : (gw <= 16 ? 16 : 32), by = 256 / bx;
  const dim3 block(bx, by), grid((unsigned)((gw + bx - 1) / bx), (unsigned)((height / 2 + 1 + by - 1) / by), (unsigned)n);
  if (alpha) wg::upsample_nrgba_kernel<true><<<grid, block, 0, ctx->stream>>>(up);
  else wg::upsample_nrgba_kernel<false><<<grid, block, 0, ctx->stream>>>(up);
  ctx->launches++;
  CK(cudaGetLastError());
  return WGPU_OK;
}

int wgpu_dec_parse(wgpu_ctx* ctx, const uint8_t* const* streams, const size_t* lens, int n, int* width_out, int* height_out) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!streams || !lens || n <= 0) FAIL(WGPU_ERR_INVALID, "webp: nil reader");
  CK(cudaSetDevice(ctx->dev));
  ctx->d_ready = false;
  // dimensions from the first stream; every stream of the batch must match
  const uint8_t* vp8; size_t vlen; const char* perr = nullptr;
  int width = 0, height = 0;
  for (int i = 0; i < n; ++i) {  // a VP8X + ALPH (or animated) file is rejected, never decoded as if it were opaque
    const uint8_t* q; size_t ql;
    if (wgh::find_vp8_ex(streams[i], lens[i], &q, &ql) < 0) FAIL(WGPU_ERR_UNSUPPORTED, std::string("image ") + std::to_string(i) + ": " + wgh::kErrOutsideLossyPath);
  }
  if (!wgh::find_vp8(streams[0], lens[0], &vp8, &vlen) || !wgh::peek_dims(vp8, vlen, &width, &height, &perr))
    FAIL(WGPU_ERR_BITSTREAM, perr ? perr : "webp: no VP8 chunk");
  const int mbw = (width + 15) >> 4, mbh = (height + 15) >> 4;
  const size_t nmb = (size_t)mbw * mbh;
  RESERVE(ctx->hd_coeffs, (size_t)n * nmb * 768);
  RESERVE(ctx->hd_meta, (size_t)n * nmb * sizeof(wgh::MBMetaH));
  RESERVE(ctx->hd_ftype, (size_t)n);
  RESERVE(ctx->d_coeffs, (size_t)n * nmb * 768);
  RESERVE(ctx->d_meta, (size_t)n * nmb * sizeof(wg::MBMeta));
  RESERVE(ctx->d_ftype, (size_t)n);
  RESERVE(ctx->dy, (size_t)n * nmb * 256); RESERVE(ctx->du, (size_t)n * nmb * 64); RESERVE(ctx->dv, (size_t)n * nmb * 64);
  std::vector<wgh::DecFrame> frames(n);
  std::atomic<int> bad(-1);
  ctx->d_dev_parsed = device_parser_wanted(ctx, (size_t)n);
  std::atomic<int> irregular(0);
  size_t stream_bytes = 0;
  if (ctx->d_dev_parsed) {
    // ---- headers on the host, macroblocks on the GPU
    static_assert(sizeof(wgh::DecHeaderH) == sizeof(wg::DecHeader), "DecHeader layout");
    std::vector<size_t> off(n + 1, 0);
    std::vector<const uint8_t*> vp(n, nullptr);
    std::vector<size_t> vl(n, 0);
    for (int i = 0; i < n; ++i) {
      if (!wgh::find_vp8(streams[i], lens[i], &vp[i], &vl[i])) FAIL(WGPU_ERR_BITSTREAM, std::string("image ") + std::to_string(i) + ": webp: no VP8 chunk");
      off[i + 1] = off[i] + ((vl[i] + 15) & ~(size_t)15);
    }
    RESERVE(ctx->hd_streams, off[n] + 16); RESERVE(ctx->d_streams, off[n] + 16);
    RESERVE(ctx->hd_hdrs, (size_t)n * sizeof(wgh::DecHeaderH)); RESERVE(ctx->d_hdrs, (size_t)n * sizeof(wgh::DecHeaderH));
    RESERVE(ctx->hd_perr, (size_t)n * 4); RESERVE(ctx->d_perr, (size_t)n * 4);
    parallel_for(n, threads_of(ctx), [&](int i) {
      wgh::DecFrame& F = frames[i];
      wgh::DecHeaderH* D = ctx->hd_hdrs.as<wgh::DecHeaderH>() + i;
      if (!wgh::parse_frame(vp[i], vl[i], &F, nullptr, nullptr, mbw, mbh, D)) { bad.store(i); return; }
      if (F.width != width || F.height != height) { F.err = "batch decode needs identical dimensions"; bad.store(i); return; }
      D->stream_off = off[i];
      memcpy(ctx->hd_streams.as<uint8_t>() + off[i], vp[i], vl[i]);
      ctx->hd_ftype.as<uint8_t>()[i] = (uint8_t)F.filter_type;
      // The device decoder's 32-bit window relies on value >> bits <= range, which every boolean-coded partition satisfies
      // unless its FIRST byte is 0xff (no encoder emits that: value < range + 1 = 255 at the start); such a (corrupt) stream
      // is parsed by the host decoder, whose wide window defines what comes out of it (tests/test_decoder_model.py).
      if (vl[i] > 10 && vp[i][10] == 0xff) irregular.store(1);
      for (int p = 0; p <= D->last_part; ++p)
        if (D->part_len[p] > 0 && vp[i][D->part_off[p]] == 0xff) irregular.store(1);
    });
    if (bad.load() >= 0) FAIL(WGPU_ERR_BITSTREAM, std::string("image ") + std::to_string(bad.load()) + ": " + (frames[bad.load()].err ? frames[bad.load()].err : "parse error"));
    if (irregular.load()) ctx->d_dev_parsed = false;
    stream_bytes = off[n];
  }
  if (ctx->d_dev_parsed) {
    CK(cudaMemcpyAsync(ctx->d_streams.p, ctx->hd_streams.p, stream_bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)stream_bytes;
    CK(cudaMemcpyAsync(ctx->d_hdrs.p, ctx->hd_hdrs.p, (size_t)n * sizeof(wgh::DecHeaderH), cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n * sizeof(wgh::DecHeaderH));
    CK(cudaMemcpyAsync(ctx->d_ftype.p, ctx->hd_ftype.p, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n);
    // no zero fill of d_coeffs: the parser writes the macroblocks with a non-zero transform code, the reconstruction reads no others
    CK(cudaMemsetAsync(ctx->d_perr.p, 0, (size_t)n * 4, ctx->stream));
    // a truncated partition makes the parser stop early: what it did not reach must not be stale or uninitialised memory
    CK(cudaMemsetAsync(ctx->d_meta.p, 0, (size_t)n * nmb * sizeof(wg::MBMeta), ctx->stream));
    wg::DecParseParams DP;
    DP.streams = ctx->d_streams.as<uint8_t>(); DP.hdr = ctx->d_hdrs.as<wg::DecHeader>(); DP.bmodes = ctx->t_bmodes.as<uint8_t>();
    DP.coeffs = ctx->d_coeffs.as<int16_t>(); DP.meta = ctx->d_meta.as<wg::MBMeta>(); DP.err = ctx->d_perr.as<int>();
    DP.n_images = n; DP.mb_w = mbw; DP.mb_h = mbh;
    // one single-thread block per image: a chain issues an instruction every ~4.4 cycles, so an SM's four schedulers carry
    // ~16 of them at full speed; 14 KB of shared memory per block caps the stacking there (the parsers of several contexts
    // in flight share the GPU: 8 x 256 images in bench.py's decode leg)
    const size_t smem = std::max(wg::dec_parse_smem(mbw), (size_t)14 * 1024);
    wg::dec_parse_kernel<<<n, 1, smem, ctx->stream>>>(DP);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->hd_perr.p, ctx->d_perr.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->xfer_d2h += (uint64_t)((size_t)n * 4);
    ctx->d_any_filter = false;
    for (int i = 0; i < n; ++i) ctx->d_any_filter |= frames[i].filter_type > 0;
    ctx->d_n = n; ctx->d_w = width; ctx->d_h = height; ctx->d_mbw = mbw; ctx->d_mbh = mbh;
    ctx->d_ready = true;
    if (width_out) *width_out = width;
    if (height_out) *height_out = height;
    return WGPU_OK;
  }
  // host: boolean decoding of headers, modes and coefficient tokens (serial per image, parallel across images)
  parallel_for(n, threads_of(ctx), [&](int i) {
    const uint8_t* p; size_t pl;
    wgh::DecFrame& F = frames[i];
    if (!wgh::find_vp8(streams[i], lens[i], &p, &pl)) { F.err = "webp: no VP8 chunk"; bad.store(i); return; }
    if (!wgh::parse_frame(p, pl, &F, ctx->hd_coeffs.as<int16_t>() + (size_t)i * nmb * 384, ctx->hd_meta.as<wgh::MBMetaH>() + (size_t)i * nmb,
                          mbw, mbh)) { bad.store(i); return; }
    if (F.width != width || F.height != height) { F.err = "batch decode needs identical dimensions"; bad.store(i); return; }
    ctx->hd_ftype.as<uint8_t>()[i] = (uint8_t)F.filter_type;
  });
  if (bad.load() >= 0) FAIL(WGPU_ERR_BITSTREAM, std::string("image ") + std::to_string(bad.load()) + ": " + (frames[bad.load()].err ? frames[bad.load()].err : "parse error"));
  CK(cudaMemcpyAsync(ctx->d_coeffs.p, ctx->hd_coeffs.p, (size_t)n * nmb * 768, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)((size_t)n * nmb * 768);
  CK(cudaMemcpyAsync(ctx->d_meta.p, ctx->hd_meta.p, (size_t)n * nmb * sizeof(wg::MBMeta), cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)((size_t)n * nmb * sizeof(wg::MBMeta));
  CK(cudaMemcpyAsync(ctx->d_ftype.p, ctx->hd_ftype.p, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)((size_t)n);
  ctx->d_any_filter = false;
  for (int i = 0; i < n; ++i) ctx->d_any_filter |= frames[i].filter_type > 0;
  ctx->d_n = n; ctx->d_w = width; ctx->d_h = height; ctx->d_mbw = mbw; ctx->d_mbh = mbh;
  ctx->d_ready = true;
  if (width_out) *width_out = width;
  if (height_out) *height_out = height;
  return WGPU_OK;
}

int wgpu_dec_device(wgpu_ctx* ctx, int want_nrgba);
int wgpu_dec_reconstruct(wgpu_ctx* ctx, int n, int width, int height, const wgpu_mb_data* mbs, const uint8_t* filter_type, int want_nrgba) {
  if (!ctx) return WGPU_ERR_INVALID;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!mbs || !filter_type || n <= 0 || width <= 0 || height <= 0 || width > 16383 || height > 16383) FAIL(WGPU_ERR_INVALID, "wgpu_dec_reconstruct: bad arguments");
    static_assert(sizeof(wgpu_mb_data) == 768 + sizeof(wgh::MBMetaH), "wgpu_mb_data = Coeffs + the 32-byte side record");
    CK(cudaSetDevice(ctx->dev));
    ctx->d_ready = false;
    const int mbw = (width + 15) >> 4, mbh = (height + 15) >> 4;
    const size_t nmb = (size_t)mbw * mbh;
    RESERVE(ctx->hd_coeffs, (size_t)n * nmb * 768);
    RESERVE(ctx->hd_meta, (size_t)n * nmb * sizeof(wgh::MBMetaH));
    RESERVE(ctx->hd_ftype, (size_t)n);
    RESERVE(ctx->d_coeffs, (size_t)n * nmb * 768);
    RESERVE(ctx->d_meta, (size_t)n * nmb * sizeof(wg::MBMeta));
    RESERVE(ctx->d_ftype, (size_t)n);
    RESERVE(ctx->dy, (size_t)n * nmb * 256); RESERVE(ctx->du, (size_t)n * nmb * 64); RESERVE(ctx->dv, (size_t)n * nmb * 64);
    // array of records -> the two arrays the kernels read (coefficients, side data), into the pinned staging buffers
    int16_t* hc = ctx->hd_coeffs.as<int16_t>();
    wgh::MBMetaH* hm = ctx->hd_meta.as<wgh::MBMetaH>();
    std::atomic<int> bad(-1);
    parallel_for(n, threads_of(ctx), [&](int i) {
      if (filter_type[i] > 2) { bad.store(i); return; }
      for (size_t k = 0; k < nmb; ++k) {
        const wgpu_mb_data& m = mbs[(size_t)i * nmb + k];
        memcpy(hc + ((size_t)i * nmb + k) * 384, m.coeffs, 768);
        memcpy(&hm[(size_t)i * nmb + k], &m.non_zero_y, sizeof(wgh::MBMetaH));
        if (m.uv_mode > 3 || (m.is_i4x4 ? false : m.imodes[0] > 3)) bad.store(i);
        if (m.is_i4x4) for (int b = 0; b < 16; ++b) if (m.imodes[b] > 9) bad.store(i);
      }
      ctx->hd_ftype.as<uint8_t>()[i] = filter_type[i];
    });
    if (bad.load() >= 0) FAIL(WGPU_ERR_INVALID, std::string("image ") + std::to_string(bad.load()) + ": prediction mode or filter type out of range");
    CK(cudaMemcpyAsync(ctx->d_coeffs.p, ctx->hd_coeffs.p, (size_t)n * nmb * 768, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n * nmb * 768);
    CK(cudaMemcpyAsync(ctx->d_meta.p, ctx->hd_meta.p, (size_t)n * nmb * sizeof(wg::MBMeta), cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n * nmb * sizeof(wg::MBMeta));
    CK(cudaMemcpyAsync(ctx->d_ftype.p, ctx->hd_ftype.p, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n);
    ctx->d_any_filter = false;
    for (int i = 0; i < n; ++i) ctx->d_any_filter |= filter_type[i] > 0;
    ctx->d_dev_parsed = false;
    ctx->d_n = n; ctx->d_w = width; ctx->d_h = height; ctx->d_mbw = mbw; ctx->d_mbh = mbh;
    ctx->d_ready = true;
  }
  return wgpu_dec_device(ctx, want_nrgba);
}

int wgpu_dec_device(wgpu_ctx* ctx, int want_nrgba) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->d_ready) FAIL(WGPU_ERR_INVALID, "wgpu_dec_device called before wgpu_dec_parse");
  CK(cudaSetDevice(ctx->dev));
  const int n = ctx->d_n;
  const size_t nmb = (size_t)ctx->d_mbw * ctx->d_mbh;
  wg::DecKernelParams P;
  P.coeffs = ctx->d_coeffs.as<int16_t>(); P.meta = ctx->d_meta.as<wg::MBMeta>();
  P.y = ctx->dy.as<uint8_t>(); P.u = ctx->du.as<uint8_t>(); P.v = ctx->dv.as<uint8_t>();
  P.y_plane = nmb * 256; P.uv_plane = nmb * 64; P.filter_type = ctx->d_ftype.as<uint8_t>();
  P.n_images = n; P.mb_w = ctx->d_mbw; P.mb_h = ctx->d_mbh;
  int rc;
  if (want_nrgba) RESERVE(ctx->d_nrgba, (size_t)n * ctx->d_w * ctx->d_h * 4);
  // The device step is 2 x (mbW + 2 mbH - 2) wave launches of 3-25 us each plus the upsampler: launch-bound.  It is captured
  // once per (shape, buffers, options) into a CUDA graph and replayed; any change of the key re-captures.
  const uint64_t key[8] = {(uint64_t)n | ((uint64_t)ctx->d_mbw << 32), (uint64_t)ctx->d_mbh | ((uint64_t)ctx->d_w << 20) | ((uint64_t)ctx->d_h << 40),
                           (uint64_t)(ctx->d_any_filter ? 1 : 0) | (want_nrgba ? 2 : 0), (uint64_t)(uintptr_t)ctx->d_coeffs.p,
                           (uint64_t)(uintptr_t)ctx->d_meta.p, (uint64_t)(uintptr_t)ctx->dy.p ^ ((uint64_t)(uintptr_t)ctx->du.p << 1) ^ ((uint64_t)(uintptr_t)ctx->dv.p << 2),
                           (uint64_t)(uintptr_t)ctx->d_nrgba.p, (uint64_t)(uintptr_t)ctx->d_ftype.p};
  auto launch_all = [&]() -> int {
    int r;
    if ((r = dec_launch_recon(ctx, P))) return r;
    if (ctx->d_any_filter && (r = dec_launch_filter(ctx, P))) return r;
    if (want_nrgba && (r = launch_upsample(ctx, n, ctx->d_w, ctx->d_h, P.y, ctx->d_mbw * 16, P.u, P.v, ctx->d_mbw * 8, P.y_plane, P.uv_plane, nullptr,
                                           ctx->d_nrgba.as<uint8_t>()))) return r;
    return WGPU_OK;
  };
  static const bool use_graph = getenv_int("WGPU_DEC_GRAPH", 1) != 0;
  if (!use_graph) {
    if ((rc = launch_all())) return rc;
  } else {
    if (!ctx->d_graph || memcmp(key, ctx->d_graph_key, sizeof(key)) != 0) {
      if (ctx->d_graph) { cudaGraphExecDestroy(ctx->d_graph); ctx->d_graph = nullptr; }
      const uint64_t launches_before = ctx->launches;
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
      rc = launch_all();
      cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
      if (rc || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        if (rc) return rc;
        ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce);
        return WGPU_ERR_CUDA;
      }
      ce = cudaGraphInstantiate(&ctx->d_graph, g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) { ctx->d_graph = nullptr; ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce); return WGPU_ERR_CUDA; }
      memcpy(ctx->d_graph_key, key, sizeof(key));
      ctx->d_graph_launches = ctx->launches - launches_before;
      ctx->launches = launches_before;  // captured, not launched
    }
    CK(cudaGraphLaunch(ctx->d_graph, ctx->stream));
    ctx->launches += ctx->d_graph_launches;  // kernels the replay runs
  }
  ctx->d_has_nrgba = want_nrgba != 0;
  return WGPU_OK;
}

int wgpu_dec_fetch(wgpu_ctx* ctx, uint8_t* y, uint8_t* u, uint8_t* v, size_t y_plane_stride, size_t uv_plane_stride, uint8_t* nrgba,
                   size_t nrgba_image_stride) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->d_ready) FAIL(WGPU_ERR_INVALID, "wgpu_dec_fetch called before wgpu_dec_parse");
  CK(cudaSetDevice(ctx->dev));
  const int n = ctx->d_n;
  const size_t nmb = (size_t)ctx->d_mbw * ctx->d_mbh, yp = nmb * 256, uvp = nmb * 64, img = (size_t)ctx->d_w * ctx->d_h * 4;
  if ((y && y_plane_stride < yp) || ((u || v) && uv_plane_stride < uvp)) FAIL(WGPU_ERR_TOO_SMALL, "plane stride smaller than the padded plane");
  if (nrgba && (!ctx->d_has_nrgba || nrgba_image_stride < img)) FAIL(WGPU_ERR_TOO_SMALL, "nrgba not produced or image stride too small");
  if (ctx->d_dev_parsed) {  // the device parser reports truncated partitions here, the first point where the host waits -- and
    // BEFORE anything is copied out: planes reconstructed from a partially parsed frame never reach the caller
    CK(cudaStreamSynchronize(ctx->stream));
    const int* pe = ctx->hd_perr.as<int>();
    for (int i = 0; i < n; ++i)
      if (pe[i]) FAIL(WGPU_ERR_BITSTREAM, std::string("image ") + std::to_string(i) + ": vp8: premature end of data");
  }
  if (y) { CK(cudaMemcpy2DAsync(y, y_plane_stride, ctx->dy.p, yp, yp, n, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)yp * n; }
  if (u) { CK(cudaMemcpy2DAsync(u, uv_plane_stride, ctx->du.p, uvp, uvp, n, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)uvp * n; }
  if (v) { CK(cudaMemcpy2DAsync(v, uv_plane_stride, ctx->dv.p, uvp, uvp, n, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)uvp * n; }
  if (nrgba) { CK(cudaMemcpy2DAsync(nrgba, nrgba_image_stride, ctx->d_nrgba.p, img, img, n, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)img * n; }
  CK(cudaStreamSynchronize(ctx->stream));
  return WGPU_OK;
}

int wgpu_decode_batch(wgpu_ctx* ctx, const uint8_t* const* streams, const size_t* lens, int n, uint8_t* y, uint8_t* u, uint8_t* v,
                      size_t y_plane_stride, size_t uv_plane_stride, uint8_t* nrgba, size_t nrgba_image_stride) {
  int rc = wgpu_dec_parse(ctx, streams, lens, n, nullptr, nullptr);
  if (rc) return rc;
  if ((rc = wgpu_dec_device(ctx, nrgba != nullptr))) return rc;
  return wgpu_dec_fetch(ctx, y, u, v, y_plane_stride, uv_plane_stride, nrgba, nrgba_image_stride);
}

// ======================================================================================== stage-level
int wgpu_import_rgba(wgpu_ctx* ctx, const uint8_t* rgba, int n, int width, int height, int stride, size_t image_stride, int has_alpha,
                     uint8_t* y, uint8_t* u, uint8_t* v) {
  if (!ctx) return WGPU_ERR_INVALID;
  int rc = wgpu_enc_upload(ctx, rgba, n, width, height, stride, image_stride);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  wgpu_enc_options_default(&ctx->e_opt, 75);
  ctx->e_opt.has_alpha = has_alpha & 1;
  ctx->e_opt.use_sharp_yuv = (has_alpha >> 1) & 1;
  ctx->e_opt.dither_amp = (has_alpha >> 8) & 0x1ff;  // bits 8.. of has_alpha carry the dithering amplitude (stage-level entry)
  if ((rc = enc_reserve(ctx))) return rc;
  if ((rc = enc_launch_import(ctx))) return rc;
  const size_t nmb = (size_t)ctx->e_mbw * ctx->e_mbh;
  if (y) { CK(cudaMemcpyAsync(y, ctx->sy.p, (size_t)n * nmb * 256, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)((size_t)n * nmb * 256); }
  if (u) { CK(cudaMemcpyAsync(u, ctx->su.p, (size_t)n * nmb * 64, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)((size_t)n * nmb * 64); }
  if (v) { CK(cudaMemcpyAsync(v, ctx->sv.p, (size_t)n * nmb * 64, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)((size_t)n * nmb * 64); }
  CK(cudaStreamSynchronize(ctx->stream));
  return WGPU_OK;
}

int wgpu_cleanup_transparent(wgpu_ctx* ctx, const uint8_t* nrgba, int n, int width, int height, int stride, size_t image_stride, uint8_t* out) {
  if (!ctx) return WGPU_ERR_INVALID;
  if (!out) { std::lock_guard<std::mutex> lk(ctx->mu); FAIL(WGPU_ERR_INVALID, "wgpu_cleanup_transparent: nil output"); }
  int rc = wgpu_enc_upload(ctx, nrgba, n, width, height, stride, image_stride);  // rows land 16-byte aligned in ctx->rgba
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->e_uploaded = false;  // the staged pixels are about to change: not an encoder input any more
  wg::CleanupParams P;
  P.px = ctx->rgba.as<uint8_t>(); P.stride = ctx->e_rgba_stride; P.image_stride = (size_t)height * ctx->e_rgba_stride;
  P.n = n; P.width = width; P.height = height;
  const int block_rows = height / 8 + (height % 8 ? 1 : 0);
  wg::cleanup_transparent_kernel<<<(unsigned)(n * block_rows), 128, (size_t)(width / 8 + 1), ctx->stream>>>(P);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpy2DAsync(out, (size_t)4 * width, ctx->rgba.p, ctx->e_rgba_stride, (size_t)4 * width, (size_t)n * height, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += (uint64_t)4 * width * height * n;
  CK(cudaStreamSynchronize(ctx->stream));
  return WGPU_OK;
}

int wgpu_upsample_nrgba(wgpu_ctx* ctx, int n, int width, int height, const uint8_t* y, int y_stride, const uint8_t* u, const uint8_t* v,
                        int uv_stride, size_t y_plane_stride, size_t uv_plane_stride, const uint8_t* alpha, uint8_t* nrgba) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!y || !u || !v || !nrgba || n <= 0 || width <= 0 || height <= 0 || y_stride < width || uv_stride < (width + 1) / 2)
    FAIL(WGPU_ERR_INVALID, "wgpu_upsample_nrgba: bad arguments");
  CK(cudaSetDevice(ctx->dev));
  const int ch = (height + 1) / 2;
  const size_t ysz = (size_t)(n - 1) * y_plane_stride + (size_t)(height - 1) * y_stride + width;
  const size_t csz = (size_t)(n - 1) * uv_plane_stride + (size_t)(ch - 1) * uv_stride + (width + 1) / 2;
  RESERVE(ctx->dy, ysz); RESERVE(ctx->du, csz); RESERVE(ctx->dv, csz);
  RESERVE(ctx->d_nrgba, (size_t)n * width * height * 4);
  ctx->d_ready = false; ctx->d_has_nrgba = false;  // the decoder's result buffers are reused: a later wgpu_dec_fetch must not return them
  CK(cudaMemcpyAsync(ctx->dy.p, y, ysz, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)(ysz);
  CK(cudaMemcpyAsync(ctx->du.p, u, csz, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)(csz);
  CK(cudaMemcpyAsync(ctx->dv.p, v, csz, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)(csz);
  const uint8_t* d_alpha = nullptr;
  if (alpha) {
    RESERVE(ctx->d_alpha, (size_t)n * width * height);
    CK(cudaMemcpyAsync(ctx->d_alpha.p, alpha, (size_t)n * width * height, cudaMemcpyHostToDevice, ctx->stream));
    ctx->xfer_h2d += (uint64_t)((size_t)n * width * height);
    d_alpha = ctx->d_alpha.as<uint8_t>();
  }
  int rc = launch_upsample(ctx, n, width, height, ctx->dy.as<uint8_t>(), y_stride, ctx->du.as<uint8_t>(), ctx->dv.as<uint8_t>(), uv_stride,
                           y_plane_stride, uv_plane_stride, d_alpha, ctx->d_nrgba.as<uint8_t>());
  if (rc) return rc;
  CK(cudaMemcpyAsync(nrgba, ctx->d_nrgba.p, (size_t)n * width * height * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += (uint64_t)((size_t)n * width * height * 4);
  CK(cudaStreamSynchronize(ctx->stream));
  return WGPU_OK;
}

static int launch_metrics(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int width, int height, int stride, size_t plane_stride, bool want_ssim) {
  wg::MetricsParams mp;
  mp.a = a; mp.b = b; mp.plane_stride = plane_stride; mp.stride = stride; mp.width = width; mp.height = height; mp.n = n;
  mp.tiles_x = (width + wg::SS_TW - 1) / wg::SS_TW; mp.tiles_y = (height + wg::SS_TH - 1) / wg::SS_TH;
  const int tiles = mp.tiles_x * mp.tiles_y;
  RESERVE(ctx->m_sse, (size_t)n * 8); RESERVE(ctx->m_ssim, (size_t)n * 8);
  if (!want_ssim) {  // SSE / PSNR only: the streaming kernel
    CK(cudaMemsetAsync(ctx->m_sse.p, 0, (size_t)n * 8, ctx->stream));
    const long long per_img = (long long)((width + 15) / 16) * height;
    const unsigned bx = (unsigned)std::max<long long>(1, std::min<long long>((per_img + 255) / 256, std::max(1, ctx->sm_count * 16 / std::max(1, n))));
    wg::sse_kernel<<<dim3(bx, (unsigned)n), 256, 0, ctx->stream>>>(mp, ctx->m_sse.as<unsigned long long>());
    ctx->launches++;
    CK(cudaGetLastError());
    return WGPU_OK;
  }
  RESERVE(ctx->m_sse_part, (size_t)n * tiles * 8); RESERVE(ctx->m_ssim_part, (size_t)n * tiles * 8);
  mp.sse_part = ctx->m_sse_part.as<unsigned long long>(); mp.ssim_part = ctx->m_ssim_part.as<double>();
  wg::SsimSepParams sp;
  sp.a = a; sp.b = b; sp.plane_stride = plane_stride; sp.stride = stride; sp.width = width; sp.height = height; sp.n = n;
  sp.tiles_x = mp.tiles_x; sp.tiles_y = mp.tiles_y; sp.sse_part = mp.sse_part; sp.ssim_part = mp.ssim_part;
  wg::ssim_sep_kernel<<<(unsigned)(n * tiles), 256, 0, ctx->stream>>>(sp);
  wg::metrics_reduce_kernel<<<n, 256, 0, ctx->stream>>>(mp.sse_part, mp.ssim_part, tiles, ctx->m_sse.as<unsigned long long>(), ctx->m_ssim.as<double>());
  ctx->launches += 2;
  CK(cudaGetLastError());
  return WGPU_OK;
}

int wgpu_plane_metrics(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int width, int height, int stride, size_t plane_stride,
                       uint64_t* sse, double* ssim_sum) {
  if (!ctx) return WGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!a || !b || n <= 0 || width <= 0 || height <= 0 || stride < width) FAIL(WGPU_ERR_INVALID, "wgpu_plane_metrics: bad arguments");
  CK(cudaSetDevice(ctx->dev));
  const size_t sz = (size_t)(n - 1) * plane_stride + (size_t)(height - 1) * stride + width;
  RESERVE(ctx->m_a, sz); RESERVE(ctx->m_b, sz);
  CK(cudaMemcpyAsync(ctx->m_a.p, a, sz, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)(sz);
  CK(cudaMemcpyAsync(ctx->m_b.p, b, sz, cudaMemcpyHostToDevice, ctx->stream));
  ctx->xfer_h2d += (uint64_t)(sz);
  int rc = launch_metrics(ctx, n, ctx->m_a.as<uint8_t>(), ctx->m_b.as<uint8_t>(), width, height, stride, plane_stride, ssim_sum != nullptr);
  if (rc) return rc;
  if (sse) { CK(cudaMemcpyAsync(sse, ctx->m_sse.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)((size_t)n * 8); }
  if (ssim_sum) { CK(cudaMemcpyAsync(ssim_sum, ctx->m_ssim.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream)); ctx->xfer_d2h += (uint64_t)((size_t)n * 8); }
  CK(cudaStreamSynchronize(ctx->stream));
  return WGPU_OK;
}

double wgpu_psnr_from_sse(uint64_t sse, uint64_t count) {  // PSNRFromSSE (internal/dsp/ssim.go:163)
  if (sse == 0 || count == 0) return 99.0;
  return 10.0 * log10(255.0 * 255.0 / ((double)sse / (double)count));
}

// ======================================================================================== dsp batch surface
namespace {
struct Scratch {  // per-call device scratch for the per-block operator surface (parity-test entry points)
  std::vector<void*> ptrs;
  ~Scratch() { for (void* p : ptrs) cudaFree(p); }
  void* get(size_t bytes) { void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr; ptrs.push_back(p); return p; }
};
}  // namespace
#define DSP_BEGIN                                                                      \
  if (!ctx) return WGPU_ERR_INVALID;                                                   \
  std::lock_guard<std::mutex> lk(ctx->mu);                                             \
  if (n <= 0) FAIL(WGPU_ERR_INVALID, "dsp batch: n must be positive");                 \
  CK(cudaSetDevice(ctx->dev));                                                         \
  Scratch S;                                                                           \
  const unsigned grid = (unsigned)((n + 127) / 128)
#define DSP_IN(var, host, bytes)                                                        \
  auto* var = reinterpret_cast<decltype(host)>(S.get(bytes));                           \
  if (!var) FAIL(WGPU_ERR_NOMEM, "dsp batch: device allocation failed");                \
  CK(cudaMemcpyAsync((void*)var, host, bytes, cudaMemcpyHostToDevice, ctx->stream))
#define DSP_OUT(var, type, bytes)                                                       \
  type* var = reinterpret_cast<type*>(S.get(bytes));                                    \
  if (!var) FAIL(WGPU_ERR_NOMEM, "dsp batch: device allocation failed")
#define DSP_END(dev, host, bytes)                                                       \
  ctx->launches++;                                                                      \
  CK(cudaGetLastError());                                                               \
  CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream))
#define DSP_SYNC CK(cudaStreamSynchronize(ctx->stream)); return WGPU_OK

int wgpu_dsp_ftransform_batch(wgpu_ctx* ctx, int n, const uint8_t* src, const uint8_t* ref, int16_t* out) {
  DSP_BEGIN;
  DSP_IN(d_src, src, (size_t)n * 16); DSP_IN(d_ref, ref, (size_t)n * 16); DSP_OUT(d_out, int16_t, (size_t)n * 32);
  wg::dsp_ftransform_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_src, d_ref, d_out);
  DSP_END(d_out, out, (size_t)n * 32);
  DSP_SYNC;
}
int wgpu_dsp_itransform_batch(wgpu_ctx* ctx, int n, const uint8_t* ref, const int16_t* in, uint8_t* dst) {
  DSP_BEGIN;
  DSP_IN(d_ref, ref, (size_t)n * 16); DSP_IN(d_in, in, (size_t)n * 32); DSP_OUT(d_out, uint8_t, (size_t)n * 16);
  wg::dsp_itransform_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_ref, d_in, d_out);
  DSP_END(d_out, dst, (size_t)n * 16);
  DSP_SYNC;
}
int wgpu_dsp_fwht_batch(wgpu_ctx* ctx, int n, const int16_t* in, int16_t* out) {
  DSP_BEGIN;
  DSP_IN(d_in, in, (size_t)n * 32); DSP_OUT(d_out, int16_t, (size_t)n * 32);
  wg::dsp_fwht_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_in, d_out);
  DSP_END(d_out, out, (size_t)n * 32);
  DSP_SYNC;
}
int wgpu_dsp_iwht_batch(wgpu_ctx* ctx, int n, const int16_t* in, int16_t* out) {
  DSP_BEGIN;
  DSP_IN(d_in, in, (size_t)n * 32); DSP_OUT(d_out, int16_t, (size_t)n * 32);
  wg::dsp_iwht_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_in, d_out);
  DSP_END(d_out, out, (size_t)n * 32);
  DSP_SYNC;
}
int wgpu_dsp_sse4x4_batch(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  DSP_BEGIN;
  DSP_IN(d_a, a, (size_t)n * 16); DSP_IN(d_b, b, (size_t)n * 16); DSP_OUT(d_out, int32_t, (size_t)n * 4);
  wg::dsp_sse4x4_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_a, d_b, d_out);
  DSP_END(d_out, out, (size_t)n * 4);
  DSP_SYNC;
}
int wgpu_dsp_tdisto4x4_batch(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  DSP_BEGIN;
  DSP_IN(d_a, a, (size_t)n * 16); DSP_IN(d_b, b, (size_t)n * 16); DSP_OUT(d_out, int32_t, (size_t)n * 4);
  wg::dsp_tdisto4x4_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_a, d_b, d_out);
  DSP_END(d_out, out, (size_t)n * 4);
  DSP_SYNC;
}
int wgpu_dsp_pred4_batch(wgpu_ctx* ctx, int n, const uint8_t* ctx13, uint8_t* out) {
  DSP_BEGIN;
  (void)grid;
  DSP_IN(d_c, ctx13, (size_t)n * 13); DSP_OUT(d_out, uint8_t, (size_t)n * 160);
  wg::dsp_pred4_kernel<<<(unsigned)((n * 10 + 127) / 128), 128, 0, ctx->stream>>>(n * 10, d_c, d_out);
  DSP_END(d_out, out, (size_t)n * 160);
  DSP_SYNC;
}
int wgpu_dsp_pred_square_batch(wgpu_ctx* ctx, int n, int size, const uint8_t* ctx_px, uint8_t* out) {
  DSP_BEGIN;
  (void)grid;
  if (size != 16 && size != 8) FAIL(WGPU_ERR_INVALID, "dsp pred_square: size must be 16 or 8");
  const size_t cs = 1 + 2 * (size_t)size;
  DSP_IN(d_c, ctx_px, (size_t)n * cs); DSP_OUT(d_out, uint8_t, (size_t)n * 7 * size * size);
  wg::dsp_pred_square_kernel<<<(unsigned)((n * 7 + 15) / 16), 128, 0, ctx->stream>>>(n * 7, size, d_c, d_out);
  DSP_END(d_out, out, (size_t)n * 7 * size * size);
  DSP_SYNC;
}
static wg::SegQuant make_seg_quant(int dc_q, int ac_q, int type, int sharpen) {
  wgh::SegQuant h;
  wgh::expand_quant(&h, dc_q, ac_q, type);
  static const int kSharp[16] = {0, 30, 60, 90, 30, 60, 90, 90, 60, 90, 90, 90, 90, 90, 90, 90};
  if (sharpen) for (int i = 0; i < 16; ++i) h.sharpen[i] = (int16_t)((kSharp[i] * (i == 0 ? dc_q : ac_q)) >> 11);
  wg::SegQuant d;
  static_assert(sizeof(d) == sizeof(h), "SegQuant layout");
  memcpy(&d, &h, sizeof(d));
  return d;
}
static wg::TabPtrs tab_ptrs(const wgpu_ctx* ctx) {
  wg::TabPtrs t;
  t.lc = ctx->t_lc.as<uint16_t>(); t.eob = ctx->t_eob.as<uint16_t>(); t.lfc = ctx->t_lfc.as<uint16_t>();
  return t;
}
int wgpu_dsp_quantize_batch(wgpu_ctx* ctx, int n, const int16_t* in, int dc_q, int ac_q, int type, int sharpen, int first, int16_t* out,
                            int32_t* nz) {
  DSP_BEGIN;
  if (dc_q <= 0 || ac_q <= 0 || type < 0 || type > 2) FAIL(WGPU_ERR_INVALID, "dsp quantize: bad quantiser");
  DSP_IN(d_in, in, (size_t)n * 32); DSP_OUT(d_out, int16_t, (size_t)n * 32); DSP_OUT(d_nz, int32_t, (size_t)n * 4);
  wg::dsp_quantize_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_in, make_seg_quant(dc_q, ac_q, type, sharpen), first, d_out, d_nz);
  DSP_END(d_out, out, (size_t)n * 32);
  CK(cudaMemcpyAsync(nz, d_nz, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += (uint64_t)((size_t)n * 4);
  DSP_SYNC;
}
int wgpu_dsp_trellis_batch(wgpu_ctx* ctx, int n, const int16_t* in, int dc_q, int ac_q, int qtype, int sharpen, int first, int ctx_type,
                           const int32_t* ctx0, int lambda, int16_t* out, int32_t* nz) {
  DSP_BEGIN;
  if (dc_q <= 0 || ac_q <= 0 || qtype < 0 || qtype > 2 || ctx_type < 0 || ctx_type > 3) FAIL(WGPU_ERR_INVALID, "dsp trellis: bad arguments");
  DSP_IN(d_in, in, (size_t)n * 32); DSP_IN(d_ctx, ctx0, (size_t)n * 4);
  DSP_OUT(d_out, int16_t, (size_t)n * 32); DSP_OUT(d_nz, int32_t, (size_t)n * 4);
  wg::dsp_trellis_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_in, make_seg_quant(dc_q, ac_q, qtype, sharpen), first, ctx_type, d_ctx, lambda,
                                                        tab_ptrs(ctx), d_out, d_nz);
  DSP_END(d_out, out, (size_t)n * 32);
  CK(cudaMemcpyAsync(nz, d_nz, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->xfer_d2h += (uint64_t)((size_t)n * 4);
  DSP_SYNC;
}
int wgpu_dsp_token_cost_batch(wgpu_ctx* ctx, int n, const int16_t* levels, const int32_t* nz, int ctx_type, const int32_t* ctx0, int first,
                              int32_t* out) {
  DSP_BEGIN;
  if (ctx_type < 0 || ctx_type > 3) FAIL(WGPU_ERR_INVALID, "dsp token cost: bad type");
  DSP_IN(d_lv, levels, (size_t)n * 32); DSP_IN(d_nz, nz, (size_t)n * 4); DSP_IN(d_ctx, ctx0, (size_t)n * 4);
  DSP_OUT(d_out, int32_t, (size_t)n * 4);
  wg::dsp_token_cost_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_lv, d_nz, ctx_type, d_ctx, first, tab_ptrs(ctx), d_out);
  DSP_END(d_out, out, (size_t)n * 4);
  DSP_SYNC;
}

// ======================================================================================== measurement
int wgpu_dsp_sse16x16_batch(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  DSP_BEGIN;
  DSP_IN(d_a, a, (size_t)n * 256); DSP_IN(d_b, b, (size_t)n * 256); DSP_OUT(d_out, int32_t, (size_t)n * 4);
  wg::dsp_sse16x16_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_a, d_b, d_out);
  DSP_END(d_out, out, (size_t)n * 4);
  DSP_SYNC;
}
int wgpu_dsp_tdisto16x16_batch(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  DSP_BEGIN;
  DSP_IN(d_a, a, (size_t)n * 256); DSP_IN(d_b, b, (size_t)n * 256); DSP_OUT(d_out, int32_t, (size_t)n * 4);
  wg::dsp_tdisto16x16_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_a, d_b, d_out);
  DSP_END(d_out, out, (size_t)n * 4);
  DSP_SYNC;
}
int wgpu_dsp_dequant_batch(wgpu_ctx* ctx, int n, const int16_t* in, int dc_q, int ac_q, int16_t* out) {
  DSP_BEGIN;
  DSP_IN(d_in, in, (size_t)n * 32); DSP_OUT(d_out, int16_t, (size_t)n * 32);
  wg::SegQuant sq;
  memset(&sq, 0, sizeof(sq));
  sq.quant = ac_q; sq.dc_quant = dc_q;
  wg::dsp_dequant_kernel<<<grid, 128, 0, ctx->stream>>>(n, d_in, sq, d_out);
  DSP_END(d_out, out, (size_t)n * 32);
  DSP_SYNC;
}
int wgpu_dsp_boolcode_batch(wgpu_ctx* ctx, int n, const uint16_t* tokens, const unsigned long long* totals, uint8_t* out, size_t out_stride,
                            unsigned int* sizes, int* rounds) {
  DSP_BEGIN;
  (void)grid;
  if (!tokens || !totals || !out || !sizes) FAIL(WGPU_ERR_INVALID, "wgpu_dsp_boolcode_batch: null argument");
  std::vector<unsigned long long> bases(2 * (size_t)n);
  unsigned long long all = 0, oall = 0, src = 0;
  for (int i = 0; i < n; ++i) {
    if (totals[i] + 32 > out_stride * 8ull) FAIL(WGPU_ERR_TOO_SMALL, "wgpu_dsp_boolcode_batch: out_stride below a partition's worst case");
    bases[i] = all; all += (totals[i] + 7) & ~7ull;
    bases[n + i] = oall; oall += (totals[i] + 16 + 15) & ~15ull;
  }
  uint16_t* d_tok = reinterpret_cast<uint16_t*>(S.get((size_t)(all + 512) * 2));
  uint8_t* d_out = reinterpret_cast<uint8_t*>(S.get((size_t)oall + 64));
  unsigned long long* d_base = reinterpret_cast<unsigned long long*>(S.get((size_t)n * 24));
  unsigned int* d_size = reinterpret_cast<unsigned int*>(S.get((size_t)n * 4));
  if (!d_tok || !d_out || !d_base || !d_size) FAIL(WGPU_ERR_NOMEM, "dsp batch: device allocation failed");
  for (int i = 0; i < n; ++i) {
    if (totals[i]) CK(cudaMemcpyAsync(d_tok + bases[i], tokens + src, (size_t)totals[i] * 2, cudaMemcpyHostToDevice, ctx->stream));
    src += totals[i];
  }
  CK(cudaMemcpyAsync(d_base, bases.data(), (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_base + 2 * n, totals, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  wg::BoolCodeParams B;
  B.tokens = d_tok; B.img_base = d_base; B.img_total = d_base + 2 * n; B.out = d_out; B.out_base = d_base + n; B.out_size = d_size; B.n_images = n;
  wg::BcpParams BP;
  int rc = launch_boolcode_par(ctx, B, totals, (size_t)n, &BP);
  if (rc) return rc;
  CK(cudaMemcpyAsync(sizes, d_size, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  int extra = 0;
  {
    const uint32_t* changed = ctx->h_bcp.as<uint32_t>() + n + 1;
    int used = 0;
    for (int r = 0; r < kBcpRounds; ++r) if (changed[r]) used = r + 1;
    extra = used;
    const unsigned launches0 = (unsigned)ctx->launches;
    bool redone = false;
    if ((rc = finish_boolcode_par(ctx, BP, &redone))) return rc;
    if (redone) {
      extra = kBcpRounds + (int)(ctx->launches - launches0) - 3;
      CK(cudaMemcpyAsync(sizes, d_size, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
  }
  if (rounds) *rounds = extra;
  for (int i = 0; i < n; ++i) {
    if (sizes[i] > out_stride) FAIL(WGPU_ERR_TOO_SMALL, "wgpu_dsp_boolcode_batch: out_stride too small");
    if (sizes[i]) CK(cudaMemcpyAsync(out + (size_t)i * out_stride, d_out + bases[n + i], sizes[i], cudaMemcpyDeviceToHost, ctx->stream));
  }
  DSP_SYNC;
}
int wgpu_dsp_ftransform2_batch(wgpu_ctx* ctx, int n, const uint8_t* src, const uint8_t* ref, int16_t* out) {
  return wgpu_dsp_ftransform_batch(ctx, 2 * n, src, ref, out);  // FTransform2 = two adjacent blocks (dsp.go:14)
}
int wgpu_dsp_dec_transform_batch(wgpu_ctx* ctx, int n, int kind, const int16_t* in, const uint8_t* ref, uint8_t* dst) {
  DSP_BEGIN;
  if (kind < 0 || kind > 4) FAIL(WGPU_ERR_INVALID, "wgpu_dsp_dec_transform_batch: kind 0..4");
  const size_t per = kind < 3 ? 16 : 64;
  DSP_IN(d_in, in, (size_t)n * per * 2); DSP_IN(d_ref, ref, (size_t)n * per); DSP_OUT(d_out, uint8_t, (size_t)n * per);
  wg::dsp_dec_transform_kernel<<<grid, 128, 0, ctx->stream>>>(n, kind, d_in, d_ref, d_out);
  DSP_END(d_out, dst, (size_t)n * per);
  DSP_SYNC;
}
int wgpu_dsp_filter_batch(wgpu_ctx* ctx, int n, int kind, const uint8_t* tiles_in, int thresh, int ithresh, int hev_thresh, uint8_t* tiles_out) {
  DSP_BEGIN;
  (void)grid;
  if (kind < 0 || kind > 11) FAIL(WGPU_ERR_INVALID, "wgpu_dsp_filter_batch: kind 0..11");
  DSP_IN(d_t, tiles_in, (size_t)n * 576);
  const int tasks = n * (kind >= 8 ? 8 : 16);
  wg::dsp_filter_kernel<<<(unsigned)((tasks + 127) / 128), 128, 0, ctx->stream>>>(tasks, kind, const_cast<uint8_t*>(d_t), thresh, ithresh, hev_thresh);
  DSP_END(d_t, tiles_out, (size_t)n * 576);
  DSP_SYNC;
}
int wgpu_dsp_upsample_line_pair_batch(wgpu_ctx* ctx, int n, int width, const uint8_t* top_y, const uint8_t* bot_y, const uint8_t* top_u, const uint8_t* top_v,
                                      const uint8_t* bot_u, const uint8_t* bot_v, const uint8_t* alpha_top, const uint8_t* alpha_bot, int channels,
                                      uint8_t* top_dst, uint8_t* bot_dst) {
  DSP_BEGIN;
  (void)grid;
  if (width <= 0 || (channels != 3 && channels != 4) || !top_y || !top_u || !top_v || !bot_u || !bot_v || !top_dst || (bot_y && !bot_dst))
    FAIL(WGPU_ERR_INVALID, "wgpu_dsp_upsample_line_pair_batch: bad arguments");
  const size_t yb = (size_t)n * width, cb = (size_t)n * ((width + 1) / 2);
  DSP_IN(d_ty, top_y, yb); DSP_IN(d_tu, top_u, cb); DSP_IN(d_tv, top_v, cb); DSP_IN(d_bu, bot_u, cb); DSP_IN(d_bv, bot_v, cb);
  const uint8_t* d_by = nullptr; const uint8_t* d_at = nullptr; const uint8_t* d_ab = nullptr;
  if (bot_y) { DSP_IN(t_, bot_y, yb); d_by = t_; }
  if (alpha_top) { DSP_IN(t_, alpha_top, yb); d_at = t_; }
  if (alpha_bot) { DSP_IN(t_, alpha_bot, yb); d_ab = t_; }
  DSP_OUT(d_td, uint8_t, yb * channels); DSP_OUT(d_bd, uint8_t, yb * channels);
  wg::dsp_upsample_pair_kernel<<<(unsigned)((yb + 127) / 128), 128, 0, ctx->stream>>>(n, width, d_ty, d_by, d_tu, d_tv, d_bu, d_bv, d_at, d_ab, channels, d_td, d_bd);
  DSP_END(d_td, top_dst, yb * channels);
  if (bot_y) { DSP_END(d_bd, bot_dst, yb * channels); }
  DSP_SYNC;
}

int wgpu_timer_begin(wgpu_ctx* ctx) {
  if (!ctx) return WGPU_ERR_INVALID;
  CK(cudaSetDevice(ctx->dev));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  return WGPU_OK;
}
int wgpu_timer_end(wgpu_ctx* ctx, float* ms) {
  if (!ctx || !ms) return WGPU_ERR_INVALID;
  CK(cudaSetDevice(ctx->dev));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaEventSynchronize(ctx->ev1));
  CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return WGPU_OK;
}
uint64_t wgpu_launch_count(const wgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
