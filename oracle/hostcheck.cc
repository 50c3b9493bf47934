// TEST INFRASTRUCTURE ONLY.  Drives the PRODUCT's host-side code (webp_b200/csrc/host_enc.h: segment planning and
// the token / bool-coding serialiser that consumes the GPU's per-macroblock output) from oracle-produced per-MB data,
// so that host logic is parity-checked and timed on CPU without a GPU.  Nothing in webp_b200/ links this.
#include "encoder.h"
#include "../webp_b200/csrc/host_enc.h"
#include "../webp_b200/csrc/host_dec.h"
#include "../webp_b200/csrc/sharp_kernels.cuh"
#include "../webp_b200/csrc/ssim_sep.cuh"
#include "../webp_b200/csrc/enc_phased.cuh"
#include <algorithm>
#include <random>
#include <chrono>

static long g_bcp_parked = 0;
#define BCP_NOTE_PARKED() (++g_bcp_parked)
#include "../webp_b200/csrc/boolcode_par.cuh"
extern "C" {
long hostcheck_boolcode_parked() { return g_bcp_parked; }
struct OrcEncCfg2 {
  int quality, method, sns_strength, filter_strength, filter_sharpness, filter_type, partitions, segments, preprocessing, has_alpha, passes, dither_amp;
};
// Returns the RIFF size produced by the product serialiser (out receives it) or <0; *same = 1 when it equals the
// oracle's bytes; *ms_per_rep = serialiser time per repetition.
long hostcheck_serialize(const uint8_t* rgba, int stride, int w, int h, const OrcEncCfg2* c, uint8_t* out, long cap, int reps,
                         int* same, double* ms_per_rep) {
  orc::EncodeConfig e;
  e.quality = c->quality; e.method = c->method; e.sns_strength = c->sns_strength; e.filter_strength = c->filter_strength;
  e.filter_sharpness = c->filter_sharpness; e.filter_type = c->filter_type; e.partitions = c->partitions; e.segments = c->segments;
  e.preprocessing = c->preprocessing;
  e.pass = c->passes > 0 ? c->passes : 1;
  e.dither_amp = c->dither_amp & 0xffff;
  e.force_serial = (c->dither_amp >> 16) & 1;  // test hook: GOMAXPROCS == 1 semantics
  orc::Encoder* enc = new orc::Encoder();
  enc->init(rgba, stride, w, h, e, c->has_alpha);
  std::vector<uint8_t> ref = orc::riff_wrap(enc->encode_frame());
  const int nmb = enc->mb_w * enc->mb_h;
  static thread_local orc::ProbaStats st;
  enc->record_all_tokens(st);
  std::vector<uint32_t> stats(4 * 8 * 3 * 11 * 2);
  memcpy(stats.data(), st, stats.size() * 4);
  std::vector<uint8_t> hdr((size_t)nmb * 48), segmap(nmb);
  std::vector<int16_t> coeffs((size_t)nmb * 400);
  for (int i = 0; i < nmb; ++i) {
    const orc::MBInfo& m = enc->mb_info[i];
    uint8_t* hd = &hdr[(size_t)i * 48];
    hd[0] = (uint8_t)m.mb_type; hd[1] = m.i16_mode; hd[2] = m.uv_mode; hd[3] = m.segment; hd[4] = m.skip; hd[5] = m.nz_dc;
    memcpy(hd + 8, m.modes, 16); memcpy(hd + 24, m.nz_y, 16); memcpy(hd + 40, m.nz_uv, 8);
    memcpy(&coeffs[(size_t)i * 400], m.coeffs, 800);
  }
  wgpu_enc_options o;
  o.quality = c->quality; o.method = c->method; o.sns_strength = c->sns_strength; o.filter_strength = c->filter_strength;
  o.filter_sharpness = c->filter_sharpness; o.filter_type = c->filter_type; o.partitions = c->partitions; o.segments = c->segments;
  o.preprocessing = c->preprocessing; o.has_alpha = c->has_alpha; o.passes = c->passes;
  wgh::FramePlan fp;
  wgh::plan_frame(&fp, o, w, h, enc->alphas.data(), (long long)enc->global_uv_alpha * nmb, segmap.data());
  fp.num_parts = 1 << o.partitions;
  int seg_same = 1;
  for (int i = 0; i < nmb; ++i) seg_same &= (segmap[i] == enc->mb_info[i].segment);
  std::vector<uint8_t> riff;
  const auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < (reps > 0 ? reps : 1); ++r) {
    riff.clear();
    if (o.method < 3) wgh::serialize_frame_serial(fp, hdr.data(), coeffs.data(), segmap.data(), stats.data(), e.pass, &riff);
    else wgh::serialize_frame(fp, hdr.data(), coeffs.data(), segmap.data(), stats.data(), &riff);
  }
  *ms_per_rep = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (reps > 0 ? reps : 1);
  *same = seg_same && riff.size() == ref.size() && !memcmp(riff.data(), ref.data(), ref.size());
  if (getenv("HOSTCHECK_TOKENS") && o.partitions == 0) {  // token route: flat (bit, prob) stream + final probabilities, as the GPU hands them over
    std::vector<uint8_t> riff2;
    const auto t1 = std::chrono::steady_clock::now();
    for (int r = 0; r < (reps > 0 ? reps : 1); ++r) {
      riff2.clear();
      wgh::serialize_frame_tokens(fp, hdr.data(), segmap.data(), &enc->proba.bands[0][0][0][0], enc->tokens.data(), enc->tokens.size(), &riff2);
    }
    *ms_per_rep = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count() / (reps > 0 ? reps : 1);
    *same = *same && riff2.size() == ref.size() && !memcmp(riff2.data(), ref.data(), ref.size());
  }
  if (o.method >= 3 && (enc->mb_h < 4 || e.force_serial) && o.partitions == 0) {
    // serial RD path (with or without mid-stream refreshes): final optimizeProba on the state the refreshes left + tokens per
    // table or re-recorded -- the product's serialize_frame_tables / optimize_proba_host / build_cost_tables from oracle data
    enc->collect_all_stats(st);
    memcpy(stats.data(), st, stats.size() * 4);
    uint8_t final_proba[1056];
    memcpy(final_proba, &enc->proba_pre_final.bands[0][0][0][0], 1056);
    const int updates = wgh::optimize_proba_host(*reinterpret_cast<const wgh::Stats*>(stats.data()), *reinterpret_cast<uint8_t (*)[4][8][3][11]>(final_proba));
    const int nt = updates > 0 ? 1 : (int)enc->hist_starts.size();
    std::vector<const uint8_t*> tabs(nt);
    const int zero = 0;
    if (updates > 0) tabs[0] = final_proba;
    else for (int k = 0; k < nt; ++k) tabs[k] = &enc->proba_hist[k].bands[0][0][0][0];
    std::vector<uint8_t> riff3;
    wgh::serialize_frame_tables(fp, hdr.data(), coeffs.data(), segmap.data(), final_proba, nt, updates > 0 ? &zero : enc->hist_starts.data(), tabs.data(), &riff3);
    // cost tables of the default probabilities == what the library uploads at start-up (sanity of build_cost_tables' layout)
    static uint16_t lc[4 * 8 * 3 * 68], eobc[96];
    wgh::build_cost_tables(wgh::kCoeffsProba0, lc, eobc);
    const bool tab_ok = eobc[0] == wgh::kEntropyCost[wgh::kCoeffsProba0[0]] && lc[0] == wgh::kEntropyCost[255 - wgh::kCoeffsProba0[0]] + wgh::kEntropyCost[wgh::kCoeffsProba0[1]];
    riff = riff3;
    *same = seg_same && tab_ok && riff.size() == ref.size() && !memcmp(riff.data(), ref.data(), ref.size());
  }
  delete enc;
  if ((long)riff.size() > cap) return -2;
  memcpy(out, riff.data(), riff.size());
  return (long)riff.size();
}

// The PRODUCT's host macroblock parser (webp_b200/csrc/host_dec.h parse_frame, the route small batches and irregular
// partitions take) on one file: 0 and the per-macroblock arrays it hands to the reconstruction kernels, or <0 when it
// rejects the stream (-1 no VP8 chunk, -2 headers / macroblock data).  dims receives width, height, mb_w, mb_h, filter_type.
int hostcheck_parse(const uint8_t* data, long len, int16_t* coeffs, uint8_t* meta, long nmb_cap, int* dims) {
  const uint8_t* vp8; size_t n;
  if (!wgh::find_vp8(data, (size_t)len, &vp8, &n)) return -1;
  wgh::DecFrame F;
  int w = 0, h = 0; const char* err = nullptr;
  if (!wgh::peek_dims(vp8, n, &w, &h, &err)) return -2;
  const int mb_w = (w + 15) >> 4, mb_h = (h + 15) >> 4;
  if ((long)mb_w * mb_h > nmb_cap) return -4;
  if (!wgh::parse_frame(vp8, n, &F, coeffs, reinterpret_cast<wgh::MBMetaH*>(meta), mb_w, mb_h)) return -2;
  dims[0] = F.width; dims[1] = F.height; dims[2] = F.mb_w; dims[3] = F.mb_h; dims[4] = F.filter_type;
  return 0;
}

// The PRODUCT's SharpYUV import (webp_b200/csrc/sharp_kernels.cuh) run on the CPU in the kernels' schedule: the same per-sample
// functions the three kernels call, "threads" of a phase visited in a shuffled order (order_seed) so that any dependence
// between samples of one phase shows, barriers where the kernels have them.  Outputs padded planes as the encoder gets them.
int hostcheck_sharp(const uint8_t* rgba, int stride, int n, int width, int height, unsigned order_seed, uint8_t* y, uint8_t* u, uint8_t* v, int* iterations, int variant) {
  std::vector<uint32_t> tab(wg::kSharpG2L + wg::kSharpL2G);
  wg::sharp_build_tables(tab.data());
  wg::SharpParams P;
  P.rgba = rgba; P.image_stride = (size_t)stride * height; P.stride = stride; P.n = n; P.width = width; P.height = height;
  P.w = (width + 1) & ~1; P.h = (height + 1) & ~1; P.uv_w = P.w >> 1; P.uv_h = P.h >> 1;
  std::vector<uint16_t> by((size_t)n * P.w * P.h), ty(by.size());
  std::vector<int16_t> buv((size_t)n * 3 * P.uv_w * P.uv_h), tuv(buv.size());
  P.best_y = by.data(); P.target_y = ty.data(); P.best_uv = buv.data(); P.target_uv = tuv.data();
  P.g2l = tab.data(); P.l2g = tab.data() + wg::kSharpG2L;
  const int pad_w = ((width + 15) >> 4) * 16, pad_h = ((height + 15) >> 4) * 16;
  P.y = y; P.u = u; P.v = v; P.y_plane = (size_t)pad_w * pad_h; P.uv_plane = P.y_plane / 4; P.pad_w = pad_w; P.pad_h = pad_h;
  P.iterations = iterations;
  std::mt19937 rng(order_seed);
  auto shuffled = [&](long long count) {
    std::vector<long long> o((size_t)count);
    for (long long i = 0; i < count; ++i) o[(size_t)i] = i;
    if (order_seed) std::shuffle(o.begin(), o.end(), rng);
    return o;
  };
  const long long per1 = (long long)P.uv_w * P.uv_h;
  for (long long t : shuffled(per1 * n)) {  // sharp_init_kernel
    const int img = (int)(t / per1), rem = (int)(t % per1);
    wg::sharp_init_item(P, img, rem / P.uv_w, rem % P.uv_w, P.g2l, P.l2g);
  }
  std::vector<int> keep((size_t)3 * P.uv_w);
  for (int img = 0; img < n; ++img) {  // sharp_refine_kernel: one CTA per image
    const unsigned long long threshold = 3ull * (unsigned long long)P.w * (unsigned long long)P.h;
    unsigned long long prev_sum = ~0ull;
    int iters = 0;
    for (int iter = 0; iter < 4; ++iter) {
      unsigned long long sum = 0;
      ++iters;
      if (variant == 1) {  // sharp_refine_ring_kernel: ring of 4 residual rows, operands fetched one row pair ahead
        const int uv_w = P.uv_w, row_len = 3 * uv_w;
        std::vector<int16_t> ring((size_t)4 * row_len, (int16_t)0x7fff);
        std::vector<wg::SharpOperands> op(uv_w), nx(uv_w);
        for (long long e : shuffled(row_len)) wg::sharp_ring_preload(P, img, (int)e, ring.data());
        for (long long i : shuffled(uv_w)) wg::sharp_fetch_operands(P, img, 0, (int)i, op[(size_t)i]);
        // __syncthreads()
        for (int jp = 0; jp < P.uv_h; ++jp) {
          if (jp + 1 < P.uv_h)
            for (long long i : shuffled(uv_w)) wg::sharp_fetch_operands(P, img, jp + 1, (int)i, nx[(size_t)i]);
          const int16_t* cur = ring.data() + (size_t)(jp & 3) * row_len;
          const int16_t* prev = jp > 0 ? ring.data() + (size_t)((jp - 1) & 3) * row_len : cur;
          const int16_t* next = jp < P.uv_h - 1 ? ring.data() + (size_t)((jp + 1) & 3) * row_len : cur;
          for (long long i : shuffled(uv_w))
            sum += wg::sharp_refine_item_ring(P, img, jp, (int)i, prev, cur, next, op[(size_t)i], P.g2l, P.l2g, &keep[(size_t)3 * i]);
          // __syncthreads()
          for (long long i : shuffled(uv_w)) wg::sharp_commit_item_ring(P, img, jp, (int)i, ring.data(), op[(size_t)i], &keep[(size_t)3 * i]);
          // __syncthreads()
          op.swap(nx);
        }
      } else
      for (int jp = 0; jp < P.uv_h; ++jp) {
        for (long long i : shuffled(P.uv_w)) sum += wg::sharp_refine_item(P, img, jp, (int)i, P.g2l, P.l2g, &keep[(size_t)3 * i]);
        // __syncthreads()
        for (long long i : shuffled(P.uv_w)) wg::sharp_commit_item(P, img, jp, (int)i, &keep[(size_t)3 * i]);
        // __syncthreads()
      }
      if (wg::sharp_stop(iter, sum, prev_sum, threshold)) break;
      prev_sum = sum;
    }
    if (iterations) iterations[img] = iters;
  }
  const long long per3 = (long long)(pad_w / 2) * (pad_h / 2);
  for (long long t : shuffled(per3 * n)) {  // sharp_finish_kernel
    const int img = (int)(t / per3), rem = (int)(t % per3);
    wg::sharp_finish_item(P, img, rem / (pad_w / 2), rem % (pad_w / 2));
  }
  return 0;
}

// The PRODUCT's phase-synchronous mode search (webp_b200/csrc/enc_phased.cuh, the row-parallel RD path) run on the CPU in the
// kernel's schedule: waves in order, the CTAs of a wave one after the other, every phase as a loop over the CTA's threads in
// shuffled order (order_seed != 0), a barrier where the kernel has one.  Inputs are what the GPU path feeds the kernel (source
// planes, the product's own segment plan and folded cost tables); the per-macroblock output (48-byte header, 400 levels) and
// the reconstruction planes must equal the oracle encoder's.  Returns the number of differing macroblocks; first_bad receives
// {macroblock index, what: 1 header, 2 levels, 3 reconstruction, byte offset}.
extern "C++" {
template <int M, int NT>
static void run_phased_waves(const wg::EncKernelParams& P, const wg::CostTabs& T, const uint16_t* i4cost, unsigned order_seed) {
  std::vector<wg::PhMB> mbs(M);
  std::vector<int> order(NT);
  for (int i = 0; i < NT; ++i) order[i] = i;
  std::mt19937 rng(order_seed);
  const int waves = P.mb_w + 2 * (P.mb_h - 1);
  for (int w = 0; w < waves; ++w) {
    const int y_lo = std::max(0, (w - (P.mb_w - 1) + 1) >> 1), y_hi = std::min(P.mb_h - 1, w >> 1);
    const long long tasks = (long long)(y_hi - y_lo + 1) * P.n_images;
    if (tasks <= 0) continue;
    std::vector<long long> ctas((size_t)((tasks + M - 1) / M));
    for (size_t i = 0; i < ctas.size(); ++i) ctas[i] = (long long)i;
    if (order_seed) std::shuffle(ctas.begin(), ctas.end(), rng);
    for (long long c : ctas) {
      if (order_seed) std::shuffle(order.begin(), order.end(), rng);
      memset((void*)mbs.data(), 0xA5, sizeof(wg::PhMB) * M);  // shared memory starts undefined
      wg::ph_run_cta<M, NT>(P, mbs.data(), T, i4cost, w, c * M, 0, order_seed ? order.data() : nullptr);
    }
  }
}
}  // extern "C++"
int hostcheck_modesearch(const uint8_t* rgba, int stride, int w, int h, const OrcEncCfg2* c, unsigned order_seed, int m_per_cta, int* first_bad) {
  orc::EncodeConfig e;
  e.quality = c->quality; e.method = c->method; e.sns_strength = c->sns_strength; e.filter_strength = c->filter_strength;
  e.filter_sharpness = c->filter_sharpness; e.filter_type = c->filter_type; e.partitions = c->partitions; e.segments = c->segments;
  e.preprocessing = c->preprocessing;
  e.pass = 1;
  orc::Encoder* enc = new orc::Encoder();
  enc->init(rgba, stride, w, h, e, c->has_alpha);
  enc->encode_frame();
  const int mbw = enc->mb_w, mbh = enc->mb_h, nmb = mbw * mbh;
  if (c->method < 3 || mbh < 4) { delete enc; return -1; }  // not the row-parallel RD path
  wgpu_enc_options o;
  memset(&o, 0, sizeof(o));
  o.quality = c->quality; o.method = c->method; o.sns_strength = c->sns_strength; o.filter_strength = c->filter_strength;
  o.filter_sharpness = c->filter_sharpness; o.filter_type = c->filter_type; o.partitions = c->partitions; o.segments = c->segments;
  o.preprocessing = c->preprocessing; o.has_alpha = c->has_alpha; o.passes = 1;
  std::vector<uint8_t> segmap(nmb);
  wgh::FramePlan fp;
  wgh::plan_frame(&fp, o, w, h, enc->alphas.data(), (long long)enc->global_uv_alpha * nmb, segmap.data());
  static_assert(sizeof(wgh::SegParams) == sizeof(wg::SegParams), "SegParams layout");
  wg::ImageParams ip;
  memcpy(&ip, fp.dev, sizeof(ip));
  static uint16_t lc[wg::LC_SIZE], eobc[wg::EOB_SIZE], i4costs[1000];
  wgh::build_cost_tables(wgh::kCoeffsProba0, lc, eobc);
  wgh::compute_i4_costs(i4costs);
  std::vector<uint8_t> ry((size_t)nmb * 256, 0xEE), ru((size_t)nmb * 64, 0xEE), rv((size_t)nmb * 64, 0xEE), hdr((size_t)nmb * 48, 0xEE);
  std::vector<int16_t> coeffs((size_t)nmb * 400, 0x7777);
  std::vector<uint32_t> ctxw(nmb, 0xEEEEEEEEu);
  wg::EncKernelParams P;
  memset(&P, 0, sizeof(P));
  P.src_y = enc->src_y.data(); P.src_u = enc->src_u.data(); P.src_v = enc->src_v.data();
  P.rec_y = ry.data(); P.rec_u = ru.data(); P.rec_v = rv.data();
  P.segment = segmap.data(); P.img = &ip; P.ctx = ctxw.data();
  P.out_hdr = hdr.data(); P.out_coeffs = coeffs.data();
  P.i4_costs = i4costs; P.lc = lc; P.eob = eobc; P.lfc = wgh::kLevelFixedCosts;
  P.n_images = 1; P.width = w; P.height = h; P.mb_w = mbw; P.mb_h = mbh;
  P.method = c->method; P.max_i4_modes = c->quality < 50 ? 2 : 3;
  P.y_plane = (size_t)nmb * 256; P.uv_plane = (size_t)nmb * 64;
  wg::CostTabs T;
  static uint16_t lfc_near[wg::LFC_NEAR];  // what the kernel stages in shared memory
  memcpy(lfc_near, wgh::kLevelFixedCosts, sizeof(lfc_near));
  T.lc = lc; T.eob = eobc; T.lfc = lfc_near; T.lfc_hi = wgh::kLevelFixedCosts;
  if (m_per_cta == 8) run_phased_waves<8, 128>(P, T, i4costs, order_seed);
  else if (m_per_cta == 12) run_phased_waves<12, 192>(P, T, i4costs, order_seed);
  else run_phased_waves<16, 256>(P, T, i4costs, order_seed);
  int bad = 0;
  for (int i = 0; i < nmb; ++i) {
    const orc::MBInfo& m = enc->mb_info[i];
    uint8_t hd[48];
    memset(hd, 0, 48);
    hd[0] = (uint8_t)m.mb_type; hd[1] = m.i16_mode; hd[2] = m.uv_mode; hd[3] = m.segment; hd[4] = m.skip; hd[5] = m.nz_dc;
    memcpy(hd + 8, m.modes, 16); memcpy(hd + 24, m.nz_y, 16); memcpy(hd + 40, m.nz_uv, 8);
    if (m.mb_type == 0) { hd[1] = m.i16_mode; memset(hd + 8, 0, 16); }
    int what = 0, off = 0;
    if (memcmp(hd, &hdr[(size_t)i * 48], 48)) { what = 1; while (hd[off] == hdr[(size_t)i * 48 + off]) ++off; }
    else if (memcmp(m.coeffs, &coeffs[(size_t)i * 400], 800)) { what = 2; while (m.coeffs[off] == coeffs[(size_t)i * 400 + off]) ++off; }
    else {
      const int mx = i % mbw, my = i / mbw;
      // the oracle (like the reference's export, encode_parallel.go:1410) writes back only the visible part of the luma block
      const int wy = std::min(16, w - mx * 16), hy = std::min(16, h - my * 16);
      for (int r = 0; r < hy && !what; ++r)
        if (memcmp(&enc->y_plane[(size_t)(my * 16 + r) * enc->y_stride + mx * 16], &ry[(size_t)(my * 16 + r) * mbw * 16 + mx * 16], wy)) { what = 3; off = r; }
      for (int r = 0; r < 8 && !what; ++r)
        if (memcmp(&enc->u_plane[(size_t)(my * 8 + r) * enc->uv_stride + mx * 8], &ru[(size_t)(my * 8 + r) * mbw * 8 + mx * 8], 8) ||
            memcmp(&enc->v_plane[(size_t)(my * 8 + r) * enc->uv_stride + mx * 8], &rv[(size_t)(my * 8 + r) * mbw * 8 + mx * 8], 8)) { what = 3; off = 16 + r; }
    }
    if (what && getenv("HOSTCHECK_DEBUG") && bad < 3) {
      fprintf(stderr, "mb %d (%d,%d) what %d off %d\n ref:", i, i % mbw, i / mbw, what, off);
      for (int k = 0; k < 48; ++k) fprintf(stderr, " %d", hd[k]);
      fprintf(stderr, "\n got:");
      for (int k = 0; k < 48; ++k) fprintf(stderr, " %d", hdr[(size_t)i * 48 + k]);
      fprintf(stderr, "\n");
    }
    if (what) {
      if (!bad && first_bad) { first_bad[0] = i; first_bad[1] = what; first_bad[2] = off; first_bad[3] = m.mb_type; }
      ++bad;
    }
  }
  int seg_bad = 0;
  for (int i = 0; i < nmb; ++i) seg_bad += segmap[i] != enc->mb_info[i].segment;
  delete enc;
  return bad + seg_bad;
}

// ---- chunk-parallel boolean coder (webp_b200/csrc/boolcode_par.cuh) run on the CPU: the kernels' per-chunk functions in a
// shuffled order per launch, rounds until the relaxation reports no change.  tokens = the partitions back to back (totals[i]
// tokens each); partition i is written to out + i * out_stride, its size to sizes[i]; returns the number of relaxation rounds.
int hostcheck_boolcode_par(const uint16_t* tokens, const unsigned long long* totals, int n, uint8_t* out, long out_stride,
                           unsigned* sizes, unsigned order_seed) {
  std::vector<unsigned long long> base(n), obase(n);
  std::vector<uint32_t> first(n + 1);
  unsigned long long all = 0;
  uint32_t nchunks = 0;
  for (int i = 0; i < n; ++i) {
    base[i] = all; all += (totals[i] + 7) & ~7ull;
    obase[i] = (unsigned long long)i * (unsigned long long)out_stride;
    first[i] = nchunks; nchunks += wg::bcp_chunks_of(totals[i]);
  }
  first[n] = nchunks;
  std::vector<uint16_t> tk(all + 8, 0);
  unsigned long long src = 0;
  for (int i = 0; i < n; ++i) { memcpy(&tk[base[i]], tokens + src, (size_t)totals[i] * 2); src += totals[i]; }
  std::vector<uint8_t> entry(nchunks, 0), walked(nchunks, 0);
  std::vector<uint32_t> shift(nchunks, 0), bitpos(nchunks, 0), head(nchunks, 0), hcarry(nchunks, 0);
  std::vector<uint16_t> tail(nchunks, 0);
  std::vector<unsigned> changed(4096, 0);
  std::vector<uint32_t> pending(nchunks, 0xffffffffu), any_pending(n, 0);
  wg::BcpParams P;
  P.tokens = tk.data(); P.img_base = base.data(); P.img_total = totals; P.chunk_first = first.data(); P.n_images = n; P.n_chunks = nchunks;
  P.entry = entry.data(); P.walked = walked.data(); P.shift_total = shift.data(); P.chunk_bit = bitpos.data(); P.head = head.data();
  P.head_carry = hcarry.data(); P.tail = tail.data(); P.changed = changed.data(); P.pending = pending.data(); P.any_pending = any_pending.data(); P.round = 0; P.out = out; P.out_base = obase.data();
  P.out_size = sizes;
  std::vector<uint32_t> order(nchunks);
  for (uint32_t k = 0; k < nchunks; ++k) order[k] = k;
  unsigned rng = order_seed * 2654435761u + 12345u;
  auto shuffle = [&]() {
    if (!order_seed) return;
    for (uint32_t k = nchunks; k > 1; --k) { rng = rng * 1664525u + 1013904223u; std::swap(order[k - 1], order[(rng >> 8) % k]); }
  };
  int rounds = 0;
  for (;; ++rounds) {
    if (rounds >= 4096) return -1;
    P.round = rounds;
    shuffle();
    for (uint32_t k = 0; k < nchunks; ++k) wg::bcp_state_chunk(P, order[k]);
    if (rounds > 0 && changed[rounds] == 0) break;
  }
  for (int i = 0; i < n; ++i) {
    uint32_t run = 0;
    for (uint32_t k = first[i]; k < first[i + 1]; ++k) { bitpos[k] = run; run += shift[k]; }
  }
  shuffle();
  for (uint32_t k = 0; k < nchunks; ++k) wg::bcp_bytes_chunk(P, order[k]);
  shuffle();
  for (uint32_t k = 0; k < nchunks; ++k) wg::bcp_join_boundary(P, order[k]);
  for (int i = 0; i < n; ++i) wg::bcp_join_fix_image(P, i);
  return rounds + 1;
}

// the reference-order coder over one flat token array (oracle BoolWriter = bitio/writer_bool.go): what the chunked coder must equal
long hostcheck_boolcode_serial(const uint16_t* tokens, unsigned long long n, uint8_t* out, long cap) {
  orc::BoolWriter bw;
  for (unsigned long long i = 0; i < n; ++i) bw.put_bit(tokens[i] & 1, tokens[i] >> 8);
  std::vector<uint8_t> r = bw.finish();
  if ((long)r.size() > cap) return -1;
  memcpy(out, r.data(), r.size());
  return (long)r.size();
}

// The boundary join of boolcode_par.cuh on hand-made records (one partition of nch chunks starting at stream bits chunk_bit[]):
// every boundary in the given order, then the parked ripples.  out holds the bytes the chunks wrote themselves.
long hostcheck_bcp_join(int nch, const uint32_t* chunk_bit, uint32_t* head, uint32_t* head_carry, const uint16_t* tail, uint8_t* out,
                        const uint32_t* order) {
  uint32_t first[2] = {0, (uint32_t)nch};
  unsigned long long obase = 0, total = 0, tbase = 0;
  std::vector<uint32_t> pending(nch, 0xffffffffu), any(1, 0);
  wg::BcpParams P;
  memset(&P, 0, sizeof(P));
  P.chunk_first = first; P.n_images = 1; P.n_chunks = (uint32_t)nch; P.img_total = &total; P.img_base = &tbase; P.chunk_bit = const_cast<uint32_t*>(chunk_bit);
  P.head = head; P.head_carry = head_carry; P.tail = const_cast<uint16_t*>(tail); P.pending = pending.data(); P.any_pending = any.data();
  P.out = out; P.out_base = &obase;
  const long parked0 = g_bcp_parked;
  for (int k = 0; k < nch; ++k) wg::bcp_join_boundary(P, order[k]);
  wg::bcp_join_fix_image(P, 0);
  return g_bcp_parked - parked0;
}
// The PRODUCT's separable SSE / SSIM (webp_b200/csrc/ssim_sep.cuh) run on the CPU in the kernel's schedule: stage, hpass and
// vpass of every tile, the threads / tasks of each phase in shuffled order (order_seed), a barrier where the kernel has one,
// the tile partials added in the fixed order of metrics_reduce_kernel's input.  force_unaligned = 1 takes the byte path.
int hostcheck_ssim(const uint8_t* a, const uint8_t* b, int stride, int width, int height, unsigned order_seed, int force_unaligned,
                   unsigned long long* sse_out, double* ssim_out) {
  std::mt19937 rng(order_seed);
  auto shuffled = [&](int count) {
    std::vector<int> o((size_t)count);
    for (int i = 0; i < count; ++i) o[(size_t)i] = i;
    if (order_seed) std::shuffle(o.begin(), o.end(), rng);
    return o;
  };
  const int tiles_x = (width + wg::SS_TW - 1) / wg::SS_TW, tiles_y = (height + wg::SS_TH - 1) / wg::SS_TH;
  const bool aligned = !force_unaligned && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)stride) & 3u) == 0;
  std::vector<uint32_t> sa(wg::SS_ROWS * wg::SS_SW), sb(sa.size());
  std::vector<wg::SsimH> sh((size_t)wg::SS_ROWS * wg::SS_TW);
  unsigned long long sse = 0;
  double ssim = 0.0;
  for (int ty = 0; ty < tiles_y; ++ty)
    for (int tx = 0; tx < tiles_x; ++tx) {
      const int x0 = tx * wg::SS_TW, y0 = ty * wg::SS_TH;
      for (auto& e : sh) e = wg::SsimH{0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu};
      for (int i : shuffled(wg::SS_ROWS * wg::SS_SW)) wg::ssim_stage_word(a, b, stride, width, height, x0, y0, aligned, i, sa.data(), sb.data());
      // __syncthreads()
      unsigned long long tile_sse = 0;
      for (int t : shuffled(wg::SS_HTASKS)) {
        const int r = t >> 3, cg = t & 7;
        const uint32_t d = wg::ssim_hpass(sa.data() + r * wg::SS_SW, sb.data() + r * wg::SS_SW, cg, sh.data() + (size_t)r * wg::SS_TW);
        if (r >= 3 && r < 3 + wg::SS_TH) tile_sse += d;
      }
      // __syncthreads()
      double lane_sum[256];
      for (int t : shuffled(256)) lane_sum[t] = wg::ssim_vpass(sh.data(), t & 31, t >> 5, x0, y0, width, height);
      double tile_ssim = 0.0;
      for (int w = 0; w < 8; ++w) {  // shuffle-down tree of a warp, then warps in order
        double v[32];
        for (int l = 0; l < 32; ++l) v[l] = lane_sum[32 * w + l];
        for (int o = 16; o > 0; o >>= 1)
          for (int l = 0; l < o; ++l) v[l] += v[l + o];
        tile_ssim += v[0];
      }
      sse += tile_sse;
      ssim += tile_ssim;
    }
  *sse_out = sse;
  *ssim_out = ssim;
  return 0;
}
}
