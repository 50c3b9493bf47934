// ORACLE -- TEST INFRASTRUCTURE ONLY (see vp8_common.h).
// Flat C entry points over the CPU restatement so tests/ and bench.py's cpu_baseline leg can
// drive it through ctypes.  Nothing in webp_b200/ links this.
#include "decoder.h"
#include "encoder.h"
#include <thread>
#include <atomic>

using namespace orc;

extern "C" {

// ---- whole-codec -------------------------------------------------------------------------
struct OrcEncCfg {
  int quality, method, sns_strength, filter_strength, filter_sharpness, filter_type, partitions, segments,
      preprocessing, has_alpha, passes, dither_amp, target_size;
  float target_psnr;
  int qmin, qmax;
  int use_sharp_yuv;
};
static EncodeConfig to_cfg(const OrcEncCfg* c) {
  EncodeConfig e;
  e.quality = c->quality; e.method = c->method; e.sns_strength = c->sns_strength;
  e.filter_strength = c->filter_strength; e.filter_sharpness = c->filter_sharpness;
  e.filter_type = c->filter_type; e.partitions = c->partitions; e.segments = c->segments;
  e.preprocessing = c->preprocessing;
  e.pass = c->passes > 0 ? c->passes : 1;
  e.dither_amp = c->dither_amp & 0xffff;
  e.force_serial = (c->dither_amp >> 16) & 1;  // test hook: GOMAXPROCS == 1 semantics
  e.target_size = c->target_size; e.target_psnr = c->target_psnr; e.qmin = c->qmin; e.qmax = c->qmax <= 0 ? 100 : c->qmax;
  e.use_sharp_yuv = c->use_sharp_yuv != 0;
  return e;
}
// Encode one RGBA image (parallel-path semantics).  Returns RIFF size or <0 (-1 unsupported
// config on this path, -2 output too small).  Optional taps may be NULL.
//   mb_hdr [nmb][8]  : mb_type, i16_mode, uv_mode, segment, skip, nz_dc, alpha(after clustering), 0
//   mb_modes [nmb][16], mb_nz [nmb][24] (16 Y + 8 UV), mb_coeffs [nmb][400]
//   recon_y/u/v : reconstructed (unfiltered) planes, padded to 16*mb_w x 16*mb_h (and half)
//   src_y/u/v   : imported source planes (same geometry); alphas [nmb] : pre-clustering alpha
//   seg_out [4][8] int32 : quant, fstrength, alpha, beta, lambda_i4, lambda_i16, lambda_uv, lambda_mode
long orc_encode(const uint8_t* rgba, int stride, int w, int h, const OrcEncCfg* cfg, uint8_t* out, long out_cap,
                uint8_t* mb_hdr, uint8_t* mb_modes, uint8_t* mb_nz, int16_t* mb_coeffs, uint8_t* recon_y,
                uint8_t* recon_u, uint8_t* recon_v, uint8_t* src_y, uint8_t* src_u, uint8_t* src_v,
                uint8_t* alphas, int32_t* seg_out) {
  if (((w + 15) >> 4) > 1024) return -1;
  Encoder* enc = new Encoder();
  enc->init(rgba, stride, w, h, to_cfg(cfg), cfg->has_alpha);
  std::vector<uint8_t> vp8 = enc->encode_frame();
  std::vector<uint8_t> riff = riff_wrap(vp8);
  const size_t nmb = enc->mb_info.size();
  for (size_t i = 0; i < nmb; ++i) {
    const MBInfo& m = enc->mb_info[i];
    if (mb_hdr) {
      uint8_t* hd = mb_hdr + 8 * i;
      hd[0] = (uint8_t)m.mb_type; hd[1] = m.i16_mode; hd[2] = m.uv_mode; hd[3] = m.segment; hd[4] = m.skip;
      hd[5] = m.nz_dc; hd[6] = (uint8_t)m.alpha; hd[7] = 0;
    }
    if (mb_modes) memcpy(mb_modes + 16 * i, m.modes, 16);
    if (mb_nz) { memcpy(mb_nz + 24 * i, m.nz_y, 16); memcpy(mb_nz + 24 * i + 16, m.nz_uv, 8); }
    if (mb_coeffs) memcpy(mb_coeffs + 400 * i, m.coeffs, 800);
  }
  if (recon_y) memcpy(recon_y, enc->y_plane.data(), enc->y_plane.size());
  if (recon_u) memcpy(recon_u, enc->u_plane.data(), enc->u_plane.size());
  if (recon_v) memcpy(recon_v, enc->v_plane.data(), enc->v_plane.size());
  if (src_y) memcpy(src_y, enc->src_y.data(), enc->src_y.size());
  if (src_u) memcpy(src_u, enc->src_u.data(), enc->src_u.size());
  if (src_v) memcpy(src_v, enc->src_v.data(), enc->src_v.size());
  if (alphas) memcpy(alphas, enc->alphas.data(), enc->alphas.size());
  if (seg_out)
    for (int s = 0; s < 4; ++s) {
      int32_t* o = seg_out + 8 * s;
      const SegmentInfo& d = enc->dqm[s];
      o[0] = d.quant; o[1] = d.fstrength; o[2] = d.alpha; o[3] = d.beta; o[4] = d.lambda_i4; o[5] = d.lambda_i16;
      o[6] = d.lambda_uv; o[7] = d.lambda_mode;
    }
  long ret = (long)riff.size();
  if (ret > out_cap) ret = -2; else memcpy(out, riff.data(), riff.size());
  delete enc;
  return ret;
}

// The final token stream of the (single) token partition, bit | prob << 8 per token as recordAllTokens leaves it
// (encode_token.go:20,304), and the partition it codes to.  Test infrastructure for the device boolean coder: returns the
// token count (tokens may be null to size the buffer); *part_len / part receive the coded partition (emitTokenPartition).
long orc_encode_tokens(const uint8_t* rgba, int stride, int w, int h, const OrcEncCfg* cfg, uint16_t* tokens, long cap,
                       uint8_t* part, long part_cap, long* part_len) {
  if (((w + 15) >> 4) > 1024) return -1;
  Encoder* enc = new Encoder();
  enc->init(rgba, stride, w, h, to_cfg(cfg), cfg->has_alpha);
  (void)enc->encode_frame();
  const long n = (long)enc->tokens.size();
  if (tokens && n <= cap) memcpy(tokens, enc->tokens.data(), (size_t)n * 2);
  if (part_len) {
    std::vector<uint8_t> p = enc->emit_token_partition(0);
    *part_len = (long)p.size();
    if (part && (long)p.size() <= part_cap) memcpy(part, p.data(), p.size());
  }
  delete enc;
  return n;
}

// VP8BitWriter over one flat token array (bit | prob << 8): PutBit per token, then Finish (bitio/writer_bool.go:58-150).
long orc_boolcode(const uint16_t* tokens, unsigned long long n, uint8_t* out, long cap) {
  BoolWriter bw;
  for (unsigned long long i = 0; i < n; ++i) bw.put_bit(tokens[i] & 1, tokens[i] >> 8);
  std::vector<uint8_t> r = bw.finish();
  if ((long)r.size() > cap) return -1;
  memcpy(out, r.data(), r.size());
  return (long)r.size();
}

// Encode n same-size images on `threads` host threads (bench cpu_baseline / --impl reference).
// sizes[i] receives each RIFF size; returns total bytes or <0.
long orc_encode_batch(const uint8_t* rgba, int n, int stride, int w, int h, const OrcEncCfg* cfg, int threads,
                      long* sizes) {
  std::atomic<int> next(0);
  std::atomic<long> total(0);
  std::atomic<int> bad(0);
  auto work = [&]() {
    std::vector<uint8_t> out((size_t)w * h * 2 + (1 << 16));
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) return;
      const long r = orc_encode(rgba + (size_t)i * stride * h, stride, w, h, cfg, out.data(), (long)out.size(), 0, 0,
                                0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
      if (r < 0) bad = 1;
      if (sizes) sizes[i] = r;
      total += r;
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(work);
  for (auto& t : th) t.join();
  return bad ? -1 : total.load();
}

// Integer-operation counts of ONE encode.  Only the build with -DORC_COUNT_OPS (liboracle_ops.so) counts;
// the plain library returns -1.  counts[OP_STAGES] receives units per stage, weights[OP_STAGES] the operations per unit
// (vp8_common.h kOpWeight); returns the number of stages.
int orc_encode_ops(const uint8_t* rgba, int stride, int w, int h, const OrcEncCfg* cfg, unsigned long long* counts, unsigned* weights) {
#ifdef ORC_COUNT_OPS
  for (int i = 0; i < OP_STAGES; ++i) g_op_count[i] = 0;
  std::vector<uint8_t> out((size_t)w * h * 2 + (1 << 16));
  const long r = orc_encode(rgba, stride, w, h, cfg, out.data(), (long)out.size(), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  if (r < 0) return -2;
  for (int i = 0; i < OP_STAGES; ++i) { counts[i] = g_op_count[i]; weights[i] = kOpWeight[i]; }
  return OP_STAGES;
#else
  (void)rgba; (void)stride; (void)w; (void)h; (void)cfg; (void)counts; (void)weights;
  return -1;
#endif
}
const char* orc_op_name(int i) { return (i >= 0 && i < OP_STAGES) ? kOpName[i] : ""; }

// Decode header only: 0 ok.
int orc_decode_info(const uint8_t* data, long len, int* w, int* h, int* mb_w, int* mb_h) {
  const uint8_t* vp8; size_t n;
  if (!find_vp8_chunk(data, (size_t)len, &vp8, &n)) return -1;
  Decoder d;
  if (!d.parse_headers(vp8, n)) return -2;
  *w = d.width; *h = d.height; *mb_w = d.mb_w; *mb_h = d.mb_h;
  return 0;
}
// Full decode into padded planes (stride 16*mb_w / 8*mb_w).  filter=0 skips the loop filter
// (returns the unfiltered reconstruction); filter=3 filters with libwebp's inner-edge rule
// (cross-validation of libwebp-made streams only, see Decoder::libwebp_inner_rule).  Optional taps: per-MB parsed data.
//   mb_coeffs [nmb][384], mb_meta [nmb][24]: is_i4, uvmode, skip, segment, f_limit, f_ilevel, f_inner, hev, imodes[16]
//   mb_nz [nmb][2] uint32: non_zero_y, non_zero_uv ; hdr_out[4]: filter_type, sharpness, level, use_segment
int orc_decode(const uint8_t* data, long len, int filter, uint8_t* y, uint8_t* u, uint8_t* v, int16_t* mb_coeffs,
               uint8_t* mb_meta, uint32_t* mb_nz, int32_t* hdr_out) {
  const uint8_t* vp8; size_t n;
  if (!find_vp8_chunk(data, (size_t)len, &vp8, &n)) return -1;
  Decoder* d = new Decoder();
  int rc = 0;
  if (!d->parse_headers(vp8, n)) { delete d; return -2; }
  d->libwebp_inner_rule = (filter & 2) != 0;
  d->y_stride = 16 * d->mb_w; d->uv_stride = 8 * d->mb_w;
  d->mbs.assign((size_t)d->mb_w * d->mb_h, MBData());
  d->finfo.assign((size_t)d->mb_w * d->mb_h, FInfo{0, 0, 0, 0});
  d->y.assign((size_t)d->y_stride * 16 * d->mb_h, 0);
  d->u.assign((size_t)d->uv_stride * 8 * d->mb_h, 0);
  d->v.assign((size_t)d->uv_stride * 8 * d->mb_h, 0);
  d->precompute_filter_strengths();
  if (!d->parse_all()) rc = -3;
  if (rc == 0) {
    d->reconstruct_all();
    if (filter && d->filter_type > 0)
      for (int my = 0; my < d->mb_h; ++my)
        for (int mx = 0; mx < d->mb_w; ++mx) d->do_filter(mx, my);
    if (y) memcpy(y, d->y.data(), d->y.size());
    if (u) memcpy(u, d->u.data(), d->u.size());
    if (v) memcpy(v, d->v.data(), d->v.size());
    const size_t nmb = d->mbs.size();
    for (size_t i = 0; i < nmb; ++i) {
      const MBData& m = d->mbs[i];
      if (mb_coeffs) memcpy(mb_coeffs + 384 * i, m.coeffs, 768);
      if (mb_meta) {
        uint8_t* o = mb_meta + 24 * i;
        const FInfo& f = d->finfo[i];
        o[0] = m.is_i4x4; o[1] = m.uvmode; o[2] = m.skip; o[3] = m.segment;
        o[4] = f.f_limit; o[5] = f.f_ilevel; o[6] = f.f_inner; o[7] = f.hev_thresh;
        memcpy(o + 8, m.imodes, 16);
      }
      if (mb_nz) { mb_nz[2 * i] = m.non_zero_y; mb_nz[2 * i + 1] = m.non_zero_uv; }
    }
    if (hdr_out) { hdr_out[0] = d->filter_type; hdr_out[1] = d->f_sharpness; hdr_out[2] = d->f_level; hdr_out[3] = d->use_segment; }
  }
  delete d;
  return rc;
}

// cleanupTransparentAreaLossy on an NRGBA image, in place (encode.go:788-890): 8x8 blocks; transparent pixels of a partly
// transparent block take the average colour of its opaque pixels (integer division), fully transparent blocks of a block
// row are flattened to the colour of the first pixel of their run; right / bottom remainders are smoothened only.
static bool orc_smoothen_block(uint8_t* px, int stride, int bx, int by, int w, int h) {
  int cnt = 0, sr = 0, sg = 0, sb = 0;
  for (int y = by; y < by + h; ++y)
    for (int x = bx; x < bx + w; ++x) {
      const uint8_t* c = px + (size_t)y * stride + 4 * x;
      if (c[3] != 0) { cnt++; sr += c[0]; sg += c[1]; sb += c[2]; }
    }
  if (cnt == 0) return true;
  if (cnt < w * h) {
    const uint8_t ar = (uint8_t)(sr / cnt), ag = (uint8_t)(sg / cnt), ab = (uint8_t)(sb / cnt);
    for (int y = by; y < by + h; ++y)
      for (int x = bx; x < bx + w; ++x) {
        uint8_t* c = px + (size_t)y * stride + 4 * x;
        if (c[3] == 0) { c[0] = ar; c[1] = ag; c[2] = ab; }
      }
  }
  return false;
}
void orc_cleanup_transparent(uint8_t* px, int stride, int width, int height) {
  const int B = 8;
  for (int by = 0; by + B <= height; by += B) {
    uint8_t cr = 0, cg = 0, cb = 0;
    bool need_reset = true;
    for (int bx = 0; bx + B <= width; bx += B) {
      if (orc_smoothen_block(px, stride, bx, by, B, B)) {
        if (need_reset) {
          const uint8_t* c = px + (size_t)by * stride + 4 * bx;
          cr = c[0]; cg = c[1]; cb = c[2];
          need_reset = false;
        }
        for (int y = by; y < by + B; ++y)
          for (int x = bx; x < bx + B; ++x) {
            uint8_t* c = px + (size_t)y * stride + 4 * x;
            c[0] = cr; c[1] = cg; c[2] = cb; c[3] = 0;
          }
      } else {
        need_reset = true;
      }
    }
    const int rem = width % B;
    if (rem > 0) orc_smoothen_block(px, stride, width - rem, by, rem, B);
  }
  const int rem_h = height % B;
  if (rem_h > 0) {
    const int by = height - rem_h;
    for (int bx = 0; bx + B <= width; bx += B) orc_smoothen_block(px, stride, bx, by, B, rem_h);
    const int rem = width % B;
    if (rem > 0) orc_smoothen_block(px, stride, width - rem, by, rem, rem_h);
  }
}

// ---- stage-level ---------------------------------------------------------------------------
// RGBA -> padded YUV420 planes (encode.go:671 importImage).
void orc_import_rgba(const uint8_t* rgba, int stride, int w, int h, int has_alpha, uint8_t* y, uint8_t* u, uint8_t* v) {
  Encoder* enc = new Encoder();
  EncodeConfig c;
  c.dither_amp = has_alpha >> 8;  // bits 8.. carry the dithering amplitude for tests
  c.use_sharp_yuv = (has_alpha & 2) != 0;  // bit 1: SharpYUV planes
  has_alpha &= 1;
  enc->cfg = c; enc->width = w; enc->height = h; enc->mb_w = (w + 15) >> 4; enc->mb_h = (h + 15) >> 4;
  enc->y_stride = enc->mb_w * 16; enc->uv_stride = enc->mb_w * 8;
  enc->y_plane.assign((size_t)enc->y_stride * enc->mb_h * 16, 0);
  enc->u_plane.assign((size_t)enc->uv_stride * enc->mb_h * 8, 0);
  enc->v_plane.assign((size_t)enc->uv_stride * enc->mb_h * 8, 0);
  enc->import_image(rgba, stride, has_alpha);
  memcpy(y, enc->y_plane.data(), enc->y_plane.size());
  memcpy(u, enc->u_plane.data(), enc->u_plane.size());
  memcpy(v, enc->v_plane.data(), enc->v_plane.size());
  delete enc;
}
// sharpyuv.Convert with the WebP matrix and sRGB transfer on RGBA input: tight planes; returns the refinement passes run.
int orc_sharp_yuv(const uint8_t* rgba, int stride, int w, int h, uint8_t* y, uint8_t* u, uint8_t* v) {
  int iters = 0;
  sharp::convert(rgba, stride, w, h, y, w, u, v, (w + 1) >> 1, &iters);
  return iters;
}
// buildNRGBA (webp.go:379)
void orc_build_nrgba(int w, int h, const uint8_t* y, int ys, const uint8_t* u, const uint8_t* v, int uvs,
                     const uint8_t* alpha, uint8_t* out) {
  build_nrgba(w, h, y, ys, u, v, uvs, alpha, out);
}
// Plane metrics: SSE (ssim.go:172) and the sum over all pixels of SSIMGet / SSIMGetClipped
// (interior / border windows, as libwebp's plane accumulation).
uint64_t orc_plane_sse(const uint8_t* a, int sa, const uint8_t* b, int sb, int w, int h) {
  uint64_t s = 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) { const int d = a[x + (size_t)y * sa] - b[x + (size_t)y * sb]; s += (uint64_t)(d * d); }
  return s;
}
double orc_plane_ssim(const uint8_t* a, int sa, const uint8_t* b, int sb, int w, int h) {
  double sum = 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      if (x >= 3 && y >= 3 && x + 3 < w && y + 3 < h)
        sum += ssim_get(a + (x - 3) + (size_t)(y - 3) * sa, sa, b + (x - 3) + (size_t)(y - 3) * sb, sb);
      else
        sum += ssim_get_clipped(a, sa, b, sb, x, y, w, h);
    }
  return sum;
}
double orc_psnr_from_sse(uint64_t sse, uint64_t count) { return psnr_from_sse(sse, count); }
// per-pixel SSIM map (float64 [h][w]) for tolerance checks
void orc_plane_ssim_map(const uint8_t* a, int sa, const uint8_t* b, int sb, int w, int h, double* out) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      out[(size_t)y * w + x] = (x >= 3 && y >= 3 && x + 3 < w && y + 3 < h)
                                   ? ssim_get(a + (x - 3) + (size_t)(y - 3) * sa, sa, b + (x - 3) + (size_t)(y - 3) * sb, sb)
                                   : ssim_get_clipped(a, sa, b, sb, x, y, w, h);
}

// ---- dsp primitives, batched over n blocks (SoA) for property tests -----------------------
// blocks are 4x4 (or 16x16) u8 tiles stored densely (16 or 256 bytes each).
static void load4(const uint8_t* s, uint8_t* d) { for (int j = 0; j < 4; ++j) memcpy(d + j * BPS, s + 4 * j, 4); }
void orc_ftransform_batch(int n, const uint8_t* src, const uint8_t* ref, int16_t* out) {
  uint8_t a[4 * BPS], b[4 * BPS];
  for (int i = 0; i < n; ++i) { load4(src + 16 * i, a); load4(ref + 16 * i, b); ftransform(a, b, out + 16 * i); }
}
void orc_itransform_batch(int n, const uint8_t* ref, const int16_t* in, uint8_t* dst) {
  uint8_t a[4 * BPS], b[4 * BPS];
  for (int i = 0; i < n; ++i) {
    load4(ref + 16 * i, a);
    itransform_one(a, in + 16 * i, b);
    for (int j = 0; j < 4; ++j) memcpy(dst + 16 * i + 4 * j, b + j * BPS, 4);
  }
}
void orc_fwht_batch(int n, const int16_t* in, int16_t* out) { for (int i = 0; i < n; ++i) ftransform_wht(in + 16 * i, out + 16 * i); }
void orc_iwht_batch(int n, const int16_t* in, int16_t* out) {
  int16_t tmp[256];
  for (int i = 0; i < n; ++i) { transform_wht(in + 16 * i, tmp); for (int k = 0; k < 16; ++k) out[16 * i + k] = tmp[16 * k]; }
}
void orc_sse4x4_batch(int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  uint8_t x[4 * BPS], y[4 * BPS];
  for (int i = 0; i < n; ++i) { load4(a + 16 * i, x); load4(b + 16 * i, y); out[i] = sse4x4(x, y); }
}
void orc_tdisto4x4_batch(int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  uint8_t x[4 * BPS], y[4 * BPS];
  for (int i = 0; i < n; ++i) { load4(a + 16 * i, x); load4(b + 16 * i, y); out[i] = tdisto4x4(x, y); }
}
// 4x4 predictors: ctx[i] = 13 bytes {tl, t0..t7, l0..l3}; out 16 bytes per (block, mode) for modes 0..9
void orc_pred4_batch(int n, const uint8_t* ctx, uint8_t* out) {
  uint8_t buf[6 * BPS];
  for (int i = 0; i < n; ++i) {
    const uint8_t* c = ctx + 13 * i;
    for (int mode = 0; mode < 10; ++mode) {
      memset(buf, 0, sizeof(buf));
      const int off = BPS + 8;
      buf[off - BPS - 1] = c[0];
      memcpy(buf + off - BPS, c + 1, 8);
      for (int j = 0; j < 4; ++j) buf[off - 1 + j * BPS] = c[9 + j];
      pred_luma4(mode, buf, off);
      for (int j = 0; j < 4; ++j) memcpy(out + (i * 10 + mode) * 16 + 4 * j, buf + off + j * BPS, 4);
    }
  }
}
// 16x16 / 8x8 predictors: ctx = {tl, top[size], left[size]}; out size*size per (block, mode 0..6)
void orc_pred_square_batch(int n, int size, const uint8_t* ctx, uint8_t* out) {
  uint8_t buf[18 * BPS];
  const int cs = 1 + 2 * size;
  for (int i = 0; i < n; ++i) {
    const uint8_t* c = ctx + cs * i;
    for (int mode = 0; mode < 7; ++mode) {
      memset(buf, 0, sizeof(buf));
      const int off = BPS + 8;
      buf[off - BPS - 1] = c[0];
      memcpy(buf + off - BPS, c + 1, size);
      for (int j = 0; j < size; ++j) buf[off - 1 + j * BPS] = c[1 + size + j];
      pred_square(mode, buf, off, size);
      for (int j = 0; j < size; ++j) memcpy(out + ((size_t)(i * 7 + mode) * size + j) * size, buf + off + j * BPS, size);
    }
  }
}
// quantize (encode_quant.go:16) with a full SegmentQuant derived from (dc_q, ac_q, bias type, sharpen on/off)
static SegmentQuant make_sq(int dc_q, int ac_q, int type, int sharpen) {
  SegmentQuant sq;
  Encoder::init_segment_quant(&sq, dc_q, ac_q, type);
  for (int i = 0; i < 16; ++i) sq.sharpen[i] = sharpen ? (int16_t)((kFreqSharpening[i] * (i == 0 ? dc_q : ac_q)) >> 11) : 0;
  return sq;
}
void orc_quantize_batch(int n, const int16_t* in, int dc_q, int ac_q, int type, int sharpen, int first, int16_t* out,
                        int32_t* nz) {
  const SegmentQuant sq = make_sq(dc_q, ac_q, type, sharpen);
  for (int i = 0; i < n; ++i) nz[i] = quantize_coeffs(in + 16 * i, out + 16 * i, &sq, first);
}
void orc_trellis_batch(int n, const int16_t* in, int dc_q, int ac_q, int qtype, int sharpen, int first, int ctx_type,
                       const int32_t* ctx0, int lambda, int16_t* out, int32_t* nz) {
  const SegmentQuant sq = make_sq(dc_q, ac_q, qtype, sharpen);
  Proba p;
  reset_proba(&p);
  for (int i = 0; i < n; ++i) nz[i] = trellis_quantize_block(in + 16 * i, out + 16 * i, &sq, first, ctx_type, ctx0[i], &p, lambda);
}
void orc_token_cost_batch(int n, const int16_t* levels, const int32_t* nz, int ctx_type, const int32_t* ctx0, int first,
                          int32_t* out) {
  Proba p;
  reset_proba(&p);
  for (int i = 0; i < n; ++i) out[i] = token_cost(levels + 16 * i, nz[i], ctx_type, &p, ctx0[i], first);
}
void orc_fixed_costs_i4(uint16_t* out) {
  Encoder* e = new Encoder();
  e->compute_fixed_costs_i4();
  memcpy(out, e->fixed_costs_i4, sizeof(e->fixed_costs_i4));
  delete e;
}
void orc_gamma_tables(uint32_t* g2l, uint32_t* l2g) {
  memcpy(g2l, gamma_tables().gamma_to_linear, 256 * 4);
  memcpy(l2g, gamma_tables().linear_to_gamma, 34 * 4);
}
int orc_quality_to_qindex(int q) { return Encoder::quality_to_qindex(q); }
const uint16_t* orc_level_fixed_costs() { return kLevelFixedCosts; }
const uint16_t* orc_entropy_cost() { return kEntropyCost; }

// ---- the rest of the dsp surface, batched like the product's twins (same layouts as include/webpgpu.h)
void orc_sse16x16_batch(int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  uint8_t x[16 * BPS], y[16 * BPS];
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < 16; ++j) { memcpy(x + j * BPS, a + 256 * (size_t)i + 16 * j, 16); memcpy(y + j * BPS, b + 256 * (size_t)i + 16 * j, 16); }
    out[i] = sse16x16(x, y);
  }
}
void orc_tdisto16x16_batch(int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  uint8_t x[16 * BPS], y[16 * BPS];
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < 16; ++j) { memcpy(x + j * BPS, a + 256 * (size_t)i + 16 * j, 16); memcpy(y + j * BPS, b + 256 * (size_t)i + 16 * j, 16); }
    out[i] = tdisto16x16(x, y);
  }
}
void orc_dequant_batch(int n, const int16_t* in, int dc_q, int ac_q, int16_t* out) {
  SegmentQuant sq;
  memset(&sq, 0, sizeof(sq));
  sq.quant = ac_q; sq.dc_quant = dc_q;
  for (int i = 0; i < n; ++i) dequant_coeffs(in + 16 * (size_t)i, out + 16 * (size_t)i, &sq);
}
void orc_dec_transform_batch(int n, int kind, const int16_t* in, const uint8_t* ref, uint8_t* dst) {
  uint8_t buf[8 * BPS];
  for (int i = 0; i < n; ++i) {
    if (kind < 3) {
      for (int j = 0; j < 4; ++j) memcpy(buf + j * BPS, ref + 16 * (size_t)i + 4 * j, 4);
      if (kind == 0) transform_one(in + 16 * (size_t)i, buf); else if (kind == 1) transform_dc(in + 16 * (size_t)i, buf); else transform_ac3(in + 16 * (size_t)i, buf);
      for (int j = 0; j < 4; ++j) memcpy(dst + 16 * (size_t)i + 4 * j, buf + j * BPS, 4);
    } else {  // transformUV / transformDCUV (transforms.go:186-216): the four blocks of an 8x8 tile
      for (int j = 0; j < 8; ++j) memcpy(buf + j * BPS, ref + 64 * (size_t)i + 8 * j, 8);
      for (int b = 0; b < 4; ++b) {
        uint8_t* d = buf + (b >> 1) * 4 * BPS + (b & 1) * 4;
        if (kind == 3) transform_one(in + 64 * (size_t)i + 16 * b, d); else transform_dc(in + 64 * (size_t)i + 16 * b, d);
      }
      for (int j = 0; j < 8; ++j) memcpy(dst + 64 * (size_t)i + 8 * j, buf + j * BPS, 8);
    }
  }
}
void orc_filter_batch(int n, int kind, uint8_t* tiles, int thresh, int ithresh, int hev_t) {
  for (int i = 0; i < n; ++i) {
    uint8_t* p = tiles + 576 * (size_t)i + 4 * 24 + 4;
    const int S = 24;
    switch (kind) {
      case 0: simple_filter(p, S, 1, 16, thresh); break;                                        // SimpleVFilter16 (filter.go:93)
      case 1: simple_filter(p, 1, S, 16, thresh); break;                                        // SimpleHFilter16
      case 2: for (int k = 1; k <= 3; ++k) simple_filter(p + k * 4 * S, S, 1, 16, thresh); break;  // SimpleVFilter16i
      case 3: for (int k = 1; k <= 3; ++k) simple_filter(p + k * 4, 1, S, 16, thresh); break;      // SimpleHFilter16i
      case 4: filter_loop26(p, S, 1, 16, thresh, ithresh, hev_t); break;                        // VFilter16
      case 5: filter_loop26(p, 1, S, 16, thresh, ithresh, hev_t); break;                        // HFilter16
      case 6: for (int k = 1; k <= 3; ++k) filter_loop24(p + k * 4 * S, S, 1, 16, thresh, ithresh, hev_t); break;  // VFilter16i
      case 7: for (int k = 1; k <= 3; ++k) filter_loop24(p + k * 4, 1, S, 16, thresh, ithresh, hev_t); break;      // HFilter16i
      case 8: filter_loop26(p, S, 1, 8, thresh, ithresh, hev_t); break;                         // VFilter8 (one plane)
      case 9: filter_loop26(p, 1, S, 8, thresh, ithresh, hev_t); break;                         // HFilter8
      case 10: filter_loop24(p + 4 * S, S, 1, 8, thresh, ithresh, hev_t); break;                // VFilter8i
      default: filter_loop24(p + 4, 1, S, 8, thresh, ithresh, hev_t); break;                    // HFilter8i
    }
  }
}
void orc_upsample_line_pair_batch(int n, int width, const uint8_t* top_y, const uint8_t* bot_y, const uint8_t* top_u, const uint8_t* top_v,
                                  const uint8_t* bot_u, const uint8_t* bot_v, const uint8_t* alpha_top, const uint8_t* alpha_bot, int channels,
                                  uint8_t* top_dst, uint8_t* bot_dst) {
  const int cw = (width + 1) / 2;
  std::vector<uint8_t> t4((size_t)width * 4), b4((size_t)width * 4);
  for (int i = 0; i < n; ++i) {
    const size_t o = (size_t)i * width, c = (size_t)i * cw;
    upsample_line_pair_nrgba(top_y + o, bot_y ? bot_y + o : nullptr, top_u + c, top_v + c, bot_u + c, bot_v + c, t4.data(), b4.data(),
                             alpha_top ? alpha_top + o : nullptr, alpha_bot ? alpha_bot + o : nullptr, width);
    for (int x = 0; x < width; ++x)
      for (int k = 0; k < channels; ++k) {
        top_dst[(o + x) * channels + k] = t4[4 * x + k];
        if (bot_y) bot_dst[(o + x) * channels + k] = b4[4 * x + k];
      }
  }
}
// VP8Random (internal/dsp/random.go): state after InitRandom(dithering) -- index1, index2, amp (random_test.go:5-50)
void orc_random_init(float dithering, int* out3) {
  const int amp = dithering < 0.0f ? 0 : (dithering > 1.0f ? 256 : (int)(256.0f * dithering));  // InitRandom (random.go:39-50)
  Encoder::Random rg(amp);
  out3[0] = rg.i1; out3[1] = rg.i2; out3[2] = rg.amp;
}

}  // extern "C"
