// Host side of the encoder above the GPU stages (the part the reference keeps on the host:
// segment clustering + quantiser setup between the analysis and mode-search kernels, and the
// token / probability / boolean-coding serialiser that consumes the GPU's per-macroblock
// modes and quantised levels).  Mirrors, in batch form:
//   assignSegments / setSegmentParams / simplifySegments / setupFilterStrength / setupSegment
//       internal/lossy/encode_analysis.go:122-237,737-903, internal/lossy/encode.go:1012-1320
//   recordMBTokens / collectAllStats / optimizeProba       internal/lossy/encode_frame.go:647, encode_proba.go
//   emitPartition0 / writeMBModes / assembleFrame           internal/lossy/encode_syntax.go
//   BoolWriter                                              internal/bitio/writer_bool.go
//   writeRIFFSimple                                         encode.go:968
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <vector>
#include "../../include/webpgpu.h"

namespace wgh {

#include "vp8_tables.inc"

static const uint8_t kBands[17] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0};
static const uint8_t kZigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
static const int8_t kI4Tree[18] = {0, 1, -1, 2, -2, 3, 4, 6, -3, 5, -4, -5, -6, 7, -7, 8, -8, -9};
static const uint8_t kCat3[] = {173, 148, 140, 0}, kCat4[] = {176, 155, 140, 135, 0},
                     kCat5[] = {180, 157, 141, 134, 130, 0},
                     kCat6[] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0};
static const uint8_t* const kCats[4] = {kCat3, kCat4, kCat5, kCat6};

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int bit_cost(int bit, uint8_t p) { return bit ? kEntropyCost[255 - p] : kEntropyCost[p]; }

// Device-visible per-segment parameters; layout must match wg::SegParams (vp8_dev.cuh).
struct SegQuant { int quant, iquant, bias, dc_quant, dc_iquant, dc_bias; int16_t sharpen[16]; };
struct SegParams { SegQuant y1, y2, uv; int lambda_i4, lambda_i16, lambda_uv, lambda_mode, tlambda_i4, tlambda_i16, tlambda_sd, pad; };

struct SegHost { int quant, fstrength, alpha, beta; };

// Everything the serialiser needs besides the per-MB arrays.
struct FramePlan {
  int width, height, mb_w, mb_h;
  int num_segments;
  SegHost seg[4];
  SegParams dev[4];
  bool seg_use, seg_update_map;
  int8_t seg_quantizer[4], seg_fstrength[4];
  uint8_t seg_proba[3];
  bool f_simple;
  int f_level, f_sharpness;
  int dq_uv_dc, dq_uv_ac;
  int num_parts;  // 1, 2, 4 or 8 token partitions (encode.go:475)
  int global_uv_alpha;
};

static inline double quality_to_compression(int q) {  // encode.go:1039
  if (q <= 0) return 0.0;
  if (q >= 100) return 1.0;
  const double c = (double)q / 100.0;
  return pow(c < 0.75 ? c * (2.0 / 3.0) : 2.0 * c - 1.0, 1.0 / 3.0);
}

static inline void expand_quant(SegQuant* sq, int dcq, int acq, int type) {  // encode.go:1169
  static const int kBias[3][2] = {{96, 110}, {96, 108}, {110, 115}};
  sq->dc_quant = dcq; sq->dc_iquant = (1 << 17) / dcq; sq->dc_bias = kBias[type][0] << 9;
  sq->quant = acq; sq->iquant = (1 << 17) / acq; sq->bias = kBias[type][1] << 9;
  memset(sq->sharpen, 0, sizeof(sq->sharpen));
}
static inline void setup_segment(FramePlan* fp, const wgpu_enc_options& o, int i) {  // encode.go:1084
  static const int kSharp[16] = {0, 30, 60, 90, 30, 60, 90, 90, 60, 90, 90, 90, 90, 90, 90, 90};
  const int q = fp->seg[i].quant;
  SegParams* d = &fp->dev[i];
  const int y1dc = kDcTable[clampi(q, 0, 127)], y1ac = kAcTable[clampi(q, 0, 127)];
  int y2dc = kDcTable[clampi(q, 0, 127)] * 2;
  if (y2dc < 8) y2dc = 8;
  const int y2ac = kAcTable2[clampi(q, 0, 127)];
  const int uvdc = kDcTable[clampi(q + fp->dq_uv_dc, 0, 117)], uvac = kAcTable[clampi(q + fp->dq_uv_ac, 0, 127)];
  expand_quant(&d->y1, y1dc, y1ac, 0);
  expand_quant(&d->y2, y2dc, y2ac, 1);
  expand_quant(&d->uv, uvdc, uvac, 2);
  const int qi4 = (y1dc + 15 * y1ac + 8) >> 4, qi16 = (y2dc + 15 * y2ac + 8) >> 4, quv = (uvdc + 15 * uvac + 8) >> 4;
  auto max1 = [](int v) { return v > 1 ? v : 1; };
  d->lambda_i4 = max1((3 * qi4 * qi4) >> 7);
  d->lambda_i16 = max1(3 * qi16 * qi16);
  d->lambda_uv = max1((3 * quv * quv) >> 6);
  d->lambda_mode = max1((qi4 * qi4) >> 7);
  d->tlambda_i4 = max1((7 * qi4 * qi4) >> 3);
  d->tlambda_i16 = max1((qi16 * qi16) >> 2);
  d->tlambda_sd = (o.method >= 4 && o.sns_strength > 0) ? (o.sns_strength * qi4) >> 5 : 0;
  d->pad = 0;
  for (int k = 0; k < 16; ++k) d->y1.sharpen[k] = (int16_t)((kSharp[k] * (k == 0 ? y1dc : y1ac)) >> 11);
}

// setSegmentParams (encode_analysis.go:122) with setupFilterStrength, simplifySegments and setupSegment: everything that
// depends on the quality.  Called once by plan_frame and again by the rate-control loop (adjustQuantForTarget,
// lossy/encode.go:1580) with the segment count left by the previous call; the per-segment alpha / beta stay.
static inline void set_segment_params(FramePlan* fp, const wgpu_enc_options& o, int num_segs, uint8_t* segment_out) {
  const int total = fp->mb_w * fp->mb_h;
  const int global_uv_alpha = fp->global_uv_alpha;
  const int sns = o.sns_strength < 0 ? 0 : o.sns_strength;
  const double amp = 0.9 * (double)sns / 100.0 / 128.0;
  const double c_base = quality_to_compression(o.quality);
  for (int i = 0; i < num_segs; ++i) {
    const double c = pow(c_base, 1.0 - amp * (double)fp->seg[i].alpha);
    fp->seg[i].quant = clampi((int)(127.0 * (1.0 - c)), 0, 127);
  }
  for (int i = num_segs; i < 4; ++i) fp->seg[i].quant = fp->seg[0].quant;
  int dq = (global_uv_alpha - 64) * 10 / 70;
  dq = dq * sns / 100;
  fp->dq_uv_ac = clampi(dq, -4, 6);
  fp->dq_uv_dc = clampi(-4 * sns / 100, -15, 15);
  // setupFilterStrength (encode.go:1276)
  fp->f_simple = (o.filter_type == 0);
  fp->f_sharpness = clampi(o.filter_sharpness, 0, 7);
  if (o.filter_strength <= 0) {
    fp->f_level = 0;
  } else {
    const int level0 = 5 * o.filter_strength, cfg_segs = clampi(o.segments, 1, 4);
    for (int i = 0; i < cfg_segs; ++i) {
      const int qstep = kAcTable[clampi(fp->seg[i].quant, 0, 127)] >> 2;
      int f = kLevelsFromDelta[fp->f_sharpness * 64 + clampi(qstep, 0, 63)] * level0 / (256 + fp->seg[i].beta);
      if (f < 2) f = 0;
      if (f > 63) f = 63;
      fp->seg[i].fstrength = f;
    }
    fp->f_level = fp->seg[0].fstrength;
  }
  if (num_segs > 1) {  // simplifySegments (encode_analysis.go:197)
    int smap[4] = {0, 1, 2, 3}, nf = 1;
    for (int s1 = 1; s1 < num_segs; ++s1) {
      bool found = false;
      for (int s2 = 0; s2 < nf; ++s2)
        if (fp->seg[s1].quant == fp->seg[s2].quant && fp->seg[s1].fstrength == fp->seg[s2].fstrength) { smap[s1] = s2; found = true; break; }
      if (!found) { smap[s1] = nf; if (nf != s1) fp->seg[nf] = fp->seg[s1]; nf++; }
    }
    if (nf < num_segs) {
      for (int i = 0; i < total; ++i) segment_out[i] = (uint8_t)smap[segment_out[i]];
      for (int i = nf; i < num_segs; ++i) fp->seg[i] = fp->seg[nf - 1];
    }
    num_segs = nf;
  }
  fp->num_segments = num_segs;
  for (int i = 0; i < 4; ++i) setup_segment(fp, o, i);
}
static inline void build_segment_header(FramePlan* fp, const wgpu_enc_options& o, int num_segs) {  // encode_analysis.go:852
  fp->seg_use = num_segs > 1;
  fp->seg_update_map = fp->seg_use;
  if (fp->seg_use)  // entries past num_segs keep what an earlier call left, as the reference's do
    for (int i = 0; i < num_segs; ++i) {
      fp->seg_quantizer[i] = (int8_t)clampi(fp->seg[i].quant, -127, 127);
      const int q0 = kAcTable[clampi(fp->seg[0].quant, 0, 127)] >> 2, qi = kAcTable[clampi(fp->seg[i].quant, 0, 127)] >> 2;
      fp->seg_fstrength[i] = (int8_t)clampi((qi - q0) * o.filter_strength / 100, -63, 63);
    }
}

// Between the analysis and mode-search kernels: cluster alphas, derive every per-segment parameter.
// alphas: per-MB mixed alpha (0..255) from the analysis kernel; uv_alpha_sum: sum of per-MB UV alphas.
// segment_out: per-MB segment id (uploaded back to the GPU).
static inline void plan_frame(FramePlan* fp, const wgpu_enc_options& o, int width, int height, const uint8_t* alphas,
                              long long uv_alpha_sum, uint8_t* segment_out) {
  fp->width = width; fp->height = height;
  fp->mb_w = (width + 15) >> 4; fp->mb_h = (height + 15) >> 4;
  const int total = fp->mb_w * fp->mb_h;
  int num_segs = clampi(o.segments, 1, 4);
  memset(fp->seg, 0, sizeof(fp->seg));
  const int global_uv_alpha = (int)uv_alpha_sum / total;
  if (num_segs <= 1) {
    memset(segment_out, 0, total);
  } else {  // assignSegments (encode_analysis.go:737): k-means on the 256-bin alpha histogram
    int histo[256] = {0};
    for (int i = 0; i < total; ++i) histo[alphas[i]]++;
    int min_a = 0, max_a = 255;
    while (min_a <= 255 && histo[min_a] == 0) min_a++;
    while (max_a > min_a && histo[max_a] == 0) max_a--;
    const int range_a = max_a - min_a;
    int centers[4] = {0, 0, 0, 0}, amap[256] = {0}, wavg = 0;
    for (int k = 0; k < num_segs; ++k) centers[k] = min_a + ((2 * k + 1) * range_a) / (2 * num_segs);
    for (int iter = 0; iter < 6; ++iter) {
      int accum[4] = {0, 0, 0, 0}, dist[4] = {0, 0, 0, 0}, n = 0;
      for (int a = min_a; a <= max_a; ++a) {
        if (!histo[a]) continue;
        while (n + 1 < num_segs && abs(a - centers[n + 1]) < abs(a - centers[n])) n++;
        amap[a] = n;
        dist[n] += a * histo[a];
        accum[n] += histo[a];
      }
      int displaced = 0, tw = 0;
      wavg = 0;
      for (int s = 0; s < num_segs; ++s)
        if (accum[s] > 0) {
          const int nc = (dist[s] + accum[s] / 2) / accum[s];
          displaced += abs(centers[s] - nc);
          centers[s] = nc;
          wavg += nc * accum[s];
          tw += accum[s];
        }
      if (tw > 0) wavg = (wavg + tw / 2) / tw;
      if (displaced < 5) break;
    }
    for (int i = 0; i < total; ++i) segment_out[i] = (uint8_t)amap[alphas[i]];
    if (o.segments > 1 && (o.preprocessing & 1) && fp->mb_w >= 3 && fp->mb_h >= 3) {  // smoothSegmentMap (:76)
      const int w = fp->mb_w, h = fp->mb_h;
      std::vector<uint8_t> tmp(segment_out, segment_out + total);
      for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) {
          int cnt[4] = {0, 0, 0, 0};
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) cnt[segment_out[(y + dy) * w + x + dx]]++;
          uint8_t best = tmp[y * w + x];
          for (int s = 0; s < 4; ++s) if (cnt[s] >= 5) best = (uint8_t)s;
          tmp[y * w + x] = best;
        }
      for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) segment_out[y * w + x] = tmp[y * w + x];
    }
    int min_c = centers[0], max_c = centers[0];
    for (int s = 1; s < num_segs; ++s) { if (centers[s] < min_c) min_c = centers[s]; if (centers[s] > max_c) max_c = centers[s]; }
    int range_c = max_c - min_c;
    if (range_c == 0) range_c = 1;
    for (int s = 0; s < num_segs; ++s) {
      fp->seg[s].alpha = clampi(255 * (centers[s] - wavg) / range_c, -127, 127);
      fp->seg[s].beta = clampi(255 * (centers[s] - min_c) / range_c, 0, 255);
    }
  }
  fp->global_uv_alpha = global_uv_alpha;
  memset(fp->seg_quantizer, 0, 4);
  memset(fp->seg_fstrength, 0, 4);
  set_segment_params(fp, o, num_segs, segment_out);
  num_segs = fp->num_segments;
  build_segment_header(fp, o, num_segs);
  // setSegmentProbas (encode_analysis.go:874)
  int counts[4] = {0, 0, 0, 0};
  for (int i = 0; i < total; ++i) counts[segment_out[i]]++;
  auto gp = [](int a, int b) -> uint8_t { const int t = a + b; return t == 0 ? 255 : (uint8_t)((255 * a + t / 2) / t); };
  fp->seg_proba[0] = gp(counts[0] + counts[1], counts[2] + counts[3]);
  fp->seg_proba[1] = gp(counts[0], counts[1]);
  fp->seg_proba[2] = gp(counts[2], counts[3]);
  if (fp->seg_proba[0] == 255 && fp->seg_proba[1] == 255 && fp->seg_proba[2] == 255) {
    fp->seg_update_map = false;
    memset(segment_out, 0, total);
  }
}

// ---- boolean coder (bitio/writer_bool.go:23-210) appending to a byte vector; the hot state lives in locals of put()
struct BoolEnc {
  int32_t range = 254, value = 0;
  int run = 0, nb_bits = -8;
  std::vector<uint8_t>* out;
  uint8_t* p = nullptr;   // write cursor into out->data()
  uint8_t* lim = nullptr; // end of the reserved region
  explicit BoolEnc(std::vector<uint8_t>* o) : out(o) { grow(4096); }
  void grow(size_t extra) {
    const size_t used = p ? (size_t)(p - out->data()) : out->size();
    out->resize(used + extra + used / 2);
    p = out->data() + used;
    lim = out->data() + out->size();
  }
  void flush() {
    const int s = 8 + nb_bits;
    const int32_t bits = value >> s;
    value -= bits << s;
    nb_bits -= 8;
    if ((bits & 0xff) != 0xff) {
      if ((size_t)(lim - p) < (size_t)run + 2) grow((size_t)run + 4096);
      if ((bits & 0x100) && p > out->data()) p[-1]++;
      if (run > 0) { memset(p, (bits & 0x100) ? 0x00 : 0xff, (size_t)run); p += run; run = 0; }
      *p++ = (uint8_t)bits;
    } else {
      run++;
    }
  }
  inline void put(int bit, int prob) {
    // PutBit (writer_bool.go:58-76), branch-free except for the byte flush: the bit selects between the two sub-ranges
    // with a mask, and the renormalisation shift is 0 whenever range >= 127 (kNorm[range] == clz(range + 1) - 24).
    const int32_t split = (range * prob) >> 8;
    const int32_t mask = -(int32_t)(bit != 0);
    value += (split + 1) & mask;
    const int32_t r = split ^ ((split ^ (range - split - 1)) & mask);
    const int shift = __builtin_clz((unsigned)(r + 1)) - 24;
    range = ((r + 1) << shift) - 1;  // kNewRange
    value <<= shift;
    nb_bits += shift;
    if (nb_bits > 0) flush();
  }
  inline void put_uniform(int bit) { put(bit, 128); }  // identical arithmetic to PutBitUniform: (r*128)>>8 == r>>1
  void put_bits(uint32_t v, int n) { for (uint32_t m = 1u << (n - 1); m; m >>= 1) put_uniform((v & m) ? 1 : 0); }
  void put_signed(int v, int n) {
    put_uniform(v != 0);
    if (v == 0) return;
    if (v < 0) put_bits(((uint32_t)(-v) << 1) | 1, n + 1); else put_bits((uint32_t)v << 1, n + 1);
  }
  void finish() {
    put_bits(0, 9 - nb_bits);
    nb_bits = 0;
    flush();
    out->resize((size_t)(p - out->data()));
  }
};

// Per-MB arrays as produced by the GPU (see enc_kernels.cuh): hdr [48] = mb_type,i16,uv,segment,skip,
// nz_dc,0,0, modes[16], nz_y[16], nz_uv[8]; coeffs [400] int16.
struct MBView {
  const uint8_t* hdr;
  const int16_t* coeffs;
  int mb_type() const { return hdr[0]; }
  int skip() const { return hdr[4]; }
};

// Walks the blocks of one MB in bitstream order with the NZ-context bookkeeping of
// recordMBTokens (encode_frame.go:647); f(coeffs, nz, type, first, ctx).
template <class F>
static inline void walk_mb(const MBView& m, uint32_t* top_nz, uint32_t* left_nz, uint8_t* top_dc, uint8_t* left_dc, F f) {
  const uint32_t top = *top_nz, left = *left_nz;
  int first = 0, type = 3;
  if (m.mb_type() == 0) {
    const int nz_dc = m.hdr[5];
    f(m.coeffs + 384, nz_dc, 1, 0, (int)*top_dc + (int)*left_dc);
    *top_dc = *left_dc = (nz_dc > 0);
    first = 1;
    type = 0;
  }
  uint32_t tnz = top & 0x0f, lnz = left & 0x0f;
  for (int y = 0; y < 4; ++y) {
    uint32_t l = lnz & 1;
    for (int x = 0; x < 4; ++x) {
      const int b = y * 4 + x, nz = m.hdr[24 + b];
      f(m.coeffs + b * 16, nz, type, first, (int)(l + (tnz & 1)));
      l = nz > first;
      tnz = (tnz >> 1) | (l << 7);
    }
    tnz >>= 4;
    lnz = (lnz >> 1) | (l << 7);
  }
  uint32_t out_t = tnz, out_l = lnz >> 4;
  for (int ch = 0; ch < 4; ch += 2) {
    tnz = (top >> (4 + ch)) & 0x0f;
    lnz = (left >> (4 + ch)) & 0x0f;
    for (int y = 0; y < 2; ++y) {
      uint32_t l = lnz & 1;
      for (int x = 0; x < 2; ++x) {
        const int k = (ch / 2) * 4 + y * 2 + x, nz = m.hdr[40 + k];
        f(m.coeffs + (16 + k) * 16, nz, 2, 0, (int)(l + (tnz & 1)));
        l = nz > 0;
        tnz = (tnz >> 1) | (l << 3);
      }
      tnz >>= 2;
      lnz = (lnz >> 1) | (l << 5);
    }
    out_t |= (tnz << 4) << ch;
    out_l |= (lnz & 0xf0) << ch;
  }
  *top_nz = out_t;
  *left_nz = out_l;
}

typedef int Stats[4][8][3][11][2];
// collectCoeffStats (encode_proba.go:10-113) on the host: only the serial-path serialiser needs it (the mid-stream
// probability refreshes look at partially encoded frames); the parallel path takes its statistics from the GPU.
static inline void stat_block(const int16_t* c, int n_coeffs, int type, int first, int ctx, Stats st) {
  int n = first;
  if (n_coeffs <= first) { st[type][kBands[n]][ctx][0][0]++; return; }
  while (n < 16) {
    int b = kBands[n];
    if (n >= n_coeffs) { st[type][b][ctx][0][0]++; return; }
    st[type][b][ctx][0][1]++;
    for (;;) {
      const int v = abs((int)c[kZigzag[n]]);
      b = kBands[n];
      int(*q)[2] = st[type][b][ctx];
      if (v == 0) { q[1][0]++; if (++n >= 16) return; ctx = 0; continue; }
      q[1][1]++;
      if (v == 1) { q[2][0]++; }
      else {
        q[2][1]++;
        if (v <= 4) { q[3][0]++; if (v == 2) q[4][0]++; else { q[4][1]++; q[5][v == 3 ? 0 : 1]++; } }
        else if (v <= 10) { q[3][1]++; q[6][0]++; q[7][v <= 6 ? 0 : 1]++; }
        else {
          q[3][1]++; q[6][1]++;
          const int cat = v <= 18 ? 0 : v <= 34 ? 1 : v <= 66 ? 2 : 3;
          q[8][cat >> 1]++;
          q[9 + (cat >> 1)][cat & 1]++;
        }
      }
      ctx = (v == 1) ? 1 : 2;
      n++;
      break;
    }
  }
}

template <class Sink>
static inline void code_block(Sink& bw, const uint8_t (*pb)[3][11] /*[band][ctx][p]*/, const int16_t* c, int n_coeffs,
                              int first, int ctx) {  // RecordCoeffs + recordLevelVP8 (encode_token.go:115-300)
  int n = first;
  if (n_coeffs <= first) { bw.put(0, pb[kBands[n]][ctx][0]); return; }
  while (n < 16) {
    const uint8_t* p = pb[kBands[n]][ctx];
    if (n >= n_coeffs) { bw.put(0, p[0]); return; }
    bw.put(1, p[0]);
    for (;;) {
      int v = c[kZigzag[n]];
      const int sign = v < 0;
      if (sign) v = -v;
      if (v == 0) { bw.put(0, p[1]); if (++n >= 16) return; p = pb[kBands[n]][0]; continue; }
      bw.put(1, p[1]);
      if (v == 1) { bw.put(0, p[2]); }
      else {
        bw.put(1, p[2]);
        if (v <= 4) {
          bw.put(0, p[3]);
          if (v == 2) bw.put(0, p[4]); else { bw.put(1, p[4]); bw.put(v == 4, p[5]); }
        } else if (v <= 10) {
          bw.put(1, p[3]);
          bw.put(0, p[6]);
          if (v <= 6) { bw.put(0, p[7]); bw.put(v - 5, 159); }
          else { bw.put(1, p[7]); bw.put((v - 7) >> 1, 165); bw.put((v - 7) & 1, 145); }
        } else {
          bw.put(1, p[3]);
          bw.put(1, p[6]);
          const int cat = v <= 18 ? 0 : v <= 34 ? 1 : v <= 66 ? 2 : 3;
          bw.put(cat >> 1, p[8]);
          bw.put(cat & 1, p[9 + (cat >> 1)]);
          const int extra = v - (3 + (8 << cat));
          const uint8_t* tab = kCats[cat];
          int nbits = 0;
          while (tab[nbits]) nbits++;
          for (int i = 0; i < nbits; ++i) bw.put((extra >> (nbits - 1 - i)) & 1, tab[i]);
        }
      }
      bw.put(sign, 128);
      ctx = (v == 1) ? 1 : 2;
      n++;
      break;
    }
  }
}

static inline bool i4_subtree_has(int node, int mode) {
  if (node <= 0) return -node == mode;
  return i4_subtree_has(kI4Tree[2 * node], mode) || i4_subtree_has(kI4Tree[2 * node + 1], mode);
}
// Bit path of each 4x4 mode through kI4Tree (writeI4Mode, encode_syntax.go:474): per mode up to 4 (prob index, bit) steps.
struct I4Path { uint8_t n; uint8_t idx[8]; uint8_t bit[8]; };
struct I4Paths {
  I4Path tab[10];
  I4Paths() {
    for (int mode = 0; mode < 10; ++mode) {
      I4Path& t = tab[mode];
      t.n = 0;
      int bit = i4_subtree_has(kI4Tree[0], mode) ? 0 : 1;
      t.idx[t.n] = 0; t.bit[t.n++] = (uint8_t)bit;
      int i = kI4Tree[bit];
      while (i > 0) {
        bit = i4_subtree_has(kI4Tree[2 * i], mode) ? 0 : 1;
        t.idx[t.n] = (uint8_t)i; t.bit[t.n++] = (uint8_t)bit;
        i = kI4Tree[2 * i + bit];
      }
    }
  }
};
static inline const I4Path* i4_paths() {
  static const I4Paths paths;  // thread-safe one-time construction
  return paths.tab;
}
// VP8FixedCostsI4 (encode_analysis.go:1497): cost of signalling each 4x4 mode given (top, left).
static inline void compute_i4_costs(uint16_t* out /*[10][10][10]*/) {
  for (int top = 0; top < 10; ++top)
    for (int left = 0; left < 10; ++left) {
      const uint8_t* prob = &kBModesProba[(top * 10 + left) * 9];
      for (int mode = 0; mode < 10; ++mode) {
        int cost = 0, bit = i4_subtree_has(kI4Tree[0], mode) ? 0 : 1;
        cost += bit_cost(bit, prob[0]);
        int i = kI4Tree[bit];
        while (i > 0) {
          bit = i4_subtree_has(kI4Tree[2 * i], mode) ? 0 : 1;
          cost += bit_cost(bit, prob[i]);
          i = kI4Tree[2 * i + bit];
        }
        out[(top * 10 + left) * 10 + mode] = (uint16_t)cost;
      }
    }
}

// Partition 0 (encode_syntax.go:47-102,349-410): frame header bits, probability updates, per-MB segment/skip/modes.
static inline void emit_partition0(const FramePlan& fp, const uint8_t* mb_hdr, const uint8_t* segment_map, const uint8_t* proba /*[1056]*/,
                                   int num_skip, int skip_proba, std::vector<uint8_t>* part0) {
  const int mb_w = fp.mb_w, mb_h = fp.mb_h;
  BoolEnc bw(part0);
  bw.put_uniform(0);
  bw.put_uniform(0);
  bw.put_uniform(fp.seg_use);
  if (fp.seg_use) {
    bw.put_uniform(fp.seg_update_map);
    bw.put_uniform(1);
    bw.put_uniform(1);
    for (int i = 0; i < 4; ++i) {
      const int q = fp.seg_quantizer[i];
      if (q) { bw.put_uniform(1); bw.put_bits((uint32_t)abs(q), 7); bw.put_uniform(q < 0); } else bw.put_uniform(0);
    }
    for (int i = 0; i < 4; ++i) {
      const int f = fp.seg_fstrength[i];
      if (f) { bw.put_uniform(1); bw.put_bits((uint32_t)abs(f), 6); bw.put_uniform(f < 0); } else bw.put_uniform(0);
    }
    if (fp.seg_update_map)
      for (int i = 0; i < 3; ++i) {
        if (fp.seg_proba[i] != 255) { bw.put_uniform(1); bw.put_bits(fp.seg_proba[i], 8); } else bw.put_uniform(0);
      }
  }
  bw.put_uniform(fp.f_simple);
  bw.put_bits((uint32_t)fp.f_level, 6);
  bw.put_bits((uint32_t)fp.f_sharpness, 3);
  bw.put_uniform(0);   // use_lf_delta
  bw.put_bits((uint32_t)(fp.num_parts == 8 ? 3 : fp.num_parts == 4 ? 2 : fp.num_parts == 2 ? 1 : 0), 2);
  bw.put_bits((uint32_t)fp.seg[0].quant, 7);
  bw.put_signed(0, 4);
  bw.put_signed(0, 4);
  bw.put_signed(0, 4);
  bw.put_signed(fp.dq_uv_dc, 4);
  bw.put_signed(fp.dq_uv_ac, 4);
  bw.put_uniform(0);
  for (int i = 0; i < 4 * 8 * 3 * 11; ++i) {
    const uint8_t pr = proba[i];
    if (pr != kCoeffsProba0[i]) { bw.put(1, kCoeffsUpdateProba[i]); bw.put_bits(pr, 8); } else bw.put(0, kCoeffsUpdateProba[i]);
  }
  if (num_skip > 0) { bw.put_uniform(1); bw.put_bits((uint32_t)skip_proba, 8); } else bw.put_uniform(0);
  std::vector<uint8_t> tm(mb_w * 4, 0);
  const I4Path* paths = i4_paths();
  for (int my = 0; my < mb_h; ++my) {
    uint8_t lm[4] = {0, 0, 0, 0};
    for (int mx = 0; mx < mb_w; ++mx) {
      const int idx = my * mb_w + mx;
      const uint8_t* h = mb_hdr + (size_t)idx * 48;
      uint8_t* top = &tm[4 * mx];
      if (fp.seg_use && fp.seg_update_map) {
        const int id = segment_map[idx];
        bw.put((id >> 1) & 1, fp.seg_proba[0]);
        bw.put(id & 1, id >= 2 ? fp.seg_proba[2] : fp.seg_proba[1]);
      }
      if (num_skip > 0) bw.put(h[4] ? 1 : 0, skip_proba);
      if (h[0] == 0) {
        bw.put(1, 145);
        const int m = h[1];
        if (m == 0) { bw.put(0, 156); bw.put(0, 163); }
        else if (m == 2) { bw.put(0, 156); bw.put(1, 163); }
        else if (m == 3) { bw.put(1, 156); bw.put(0, 128); }
        else { bw.put(1, 156); bw.put(1, 128); }
        memset(top, m, 4);
        memset(lm, m, 4);
      } else {
        bw.put(0, 145);
        for (int y = 0; y < 4; ++y) {
          int ym = lm[y];
          for (int x = 0; x < 4; ++x) {
            const int mode = h[8 + y * 4 + x];
            const uint8_t* prob = &kBModesProba[(top[x] * 10 + ym) * 9];
            const I4Path& pt = paths[mode];
            for (int k = 0; k < pt.n; ++k) bw.put(pt.bit[k], prob[pt.idx[k]]);
            ym = mode;
            top[x] = (uint8_t)mode;
          }
          lm[y] = (uint8_t)ym;
        }
      }
      const int uv = h[2];
      if (uv == 0) bw.put(0, 142);
      else if (uv == 2) { bw.put(1, 142); bw.put(0, 114); }
      else if (uv == 3) { bw.put(1, 142); bw.put(1, 114); bw.put(0, 183); }
      else { bw.put(1, 142); bw.put(1, 114); bw.put(1, 183); }
    }
  }
  bw.finish();
}

// Frame tag + picture header (encode_syntax.go:118-172) and the simple RIFF container (encode.go:968-997) around
// [hdr_pos + 30, end) = partition 0 + partition sizes + partitions already appended to *riff.
static inline void finish_riff(const FramePlan& fp, size_t hdr_pos, size_t part0_size, std::vector<uint8_t>* riff) {
  uint8_t* f = riff->data() + hdr_pos + 20;
  const uint32_t tag = (1u << 4) | ((uint32_t)part0_size << 5);
  f[0] = (uint8_t)tag; f[1] = (uint8_t)(tag >> 8); f[2] = (uint8_t)(tag >> 16);
  f[3] = 0x9d; f[4] = 0x01; f[5] = 0x2a;
  f[6] = (uint8_t)fp.width; f[7] = (uint8_t)((fp.width & 0x3fff) >> 8);
  f[8] = (uint8_t)fp.height; f[9] = (uint8_t)((fp.height & 0x3fff) >> 8);
  const uint32_t payload = (uint32_t)(riff->size() - hdr_pos - 20);
  if (payload & 1) riff->push_back(0);
  uint8_t* r = riff->data() + hdr_pos;
  const uint32_t riff_size = 4 + 8 + payload + (payload & 1);
  memcpy(r, "RIFF", 4);
  r[4] = (uint8_t)riff_size; r[5] = (uint8_t)(riff_size >> 8); r[6] = (uint8_t)(riff_size >> 16); r[7] = (uint8_t)(riff_size >> 24);
  memcpy(r + 8, "WEBPVP8 ", 8);
  r[16] = (uint8_t)payload; r[17] = (uint8_t)(payload >> 8); r[18] = (uint8_t)(payload >> 16); r[19] = (uint8_t)(payload >> 24);
}

static inline int count_skips(const uint8_t* mb_hdr, int total) {
  int n = 0;
  for (int idx = 0; idx < total; ++idx) n += mb_hdr[(size_t)idx * 48 + 4] != 0;
  return n;
}

// Host partition 0 for the routes that do not generate it on the device (part0_kernels.cuh); the host emits partition 0
// and writes RIFF(20) + frame header(10) around [partition 0][token partition] laid out in place at `dst`.
static inline void emit_partition0_of(const FramePlan& fp, const uint8_t* mb_hdr, const uint8_t* segment_map, const uint8_t* proba /*[1056]*/,
                                      std::vector<uint8_t>* part0) {
  const int total = fp.mb_w * fp.mb_h;
  const int num_skip = count_skips(mb_hdr, total);
  const int skip_proba = num_skip > 0 ? (total - num_skip) * 255 / total : 0;
  part0->reserve((size_t)total * 4 + 2048);
  emit_partition0(fp, mb_hdr, segment_map, proba, num_skip, skip_proba, part0);
}
static inline size_t frame_file_size(size_t part0_size, size_t coded_size) {
  const size_t payload = 10 + part0_size + coded_size;
  return 20 + payload + (payload & 1);
}
static inline void write_frame_headers(const FramePlan& fp, uint8_t* dst, size_t part0_size, size_t coded_size) {
  uint8_t* f = dst + 20;
  const uint32_t tag = (1u << 4) | ((uint32_t)part0_size << 5);
  f[0] = (uint8_t)tag; f[1] = (uint8_t)(tag >> 8); f[2] = (uint8_t)(tag >> 16);
  f[3] = 0x9d; f[4] = 0x01; f[5] = 0x2a;
  f[6] = (uint8_t)fp.width; f[7] = (uint8_t)((fp.width & 0x3fff) >> 8);
  f[8] = (uint8_t)fp.height; f[9] = (uint8_t)((fp.height & 0x3fff) >> 8);
  const uint32_t payload = (uint32_t)(10 + part0_size + coded_size);
  if (payload & 1) dst[20 + payload] = 0;
  const uint32_t riff_size = 4 + 8 + payload + (payload & 1);
  memcpy(dst, "RIFF", 4);
  dst[4] = (uint8_t)riff_size; dst[5] = (uint8_t)(riff_size >> 8); dst[6] = (uint8_t)(riff_size >> 16); dst[7] = (uint8_t)(riff_size >> 24);
  memcpy(dst + 8, "WEBPVP8 ", 8);
  dst[16] = (uint8_t)payload; dst[17] = (uint8_t)(payload >> 8); dst[18] = (uint8_t)(payload >> 16); dst[19] = (uint8_t)(payload >> 24);
}

// Boolean-code two independent token streams in lock step: the coder is a serial dependency chain of ~20 cycles per
// token, so interleaving two images in one thread nearly doubles throughput (measured 7.4 -> 3.9 ns/token).
static inline void code_token_streams(const uint16_t* ta, size_t na, std::vector<uint8_t>* oa, const uint16_t* tb, size_t nb,
                                      std::vector<uint8_t>* ob) {
  BoolEnc a(oa);
  if (!tb) {
    for (size_t i = 0; i < na; ++i) { const uint32_t t = ta[i]; a.put((int)(t & 1), (int)(t >> 8)); }
    a.finish();
    return;
  }
  BoolEnc b(ob);
  const size_t n = na < nb ? na : nb;
  for (size_t i = 0; i < n; ++i) {
    const uint32_t t = ta[i], u = tb[i];
    a.put((int)(t & 1), (int)(t >> 8));
    b.put((int)(u & 1), (int)(u >> 8));
  }
  for (size_t i = n; i < na; ++i) { const uint32_t t = ta[i]; a.put((int)(t & 1), (int)(t >> 8)); }
  for (size_t i = n; i < nb; ++i) { const uint32_t u = tb[i]; b.put((int)(u & 1), (int)(u >> 8)); }
  a.finish();
  b.finish();
}

// Single-partition route: the GPU has already produced the final (bit, prob) token stream (token_kernels.cuh) and the
// optimised probabilities; the host writes partition 0 and runs the boolean coder over the flat token array
// (EmitTokens / PutBitBatchPacked, encode_token.go:304, bitio/writer_bool.go:106).
// `coded` = the already boolean-coded token partition of this image (code_token_streams).
static inline void assemble_frame_tokens(const FramePlan& fp, const uint8_t* mb_hdr, const uint8_t* segment_map, const uint8_t* proba /*[1056]*/,
                                         const std::vector<uint8_t>& coded, std::vector<uint8_t>* riff) {
  const int total = fp.mb_w * fp.mb_h;
  const int num_skip = count_skips(mb_hdr, total);
  const int skip_proba = num_skip > 0 ? (total - num_skip) * 255 / total : 0;
  std::vector<uint8_t> part0;
  part0.reserve((size_t)total * 4 + 2048);
  emit_partition0(fp, mb_hdr, segment_map, proba, num_skip, skip_proba, &part0);
  const size_t hdr_pos = riff->size();
  riff->resize(hdr_pos + 20 + 10);
  riff->insert(riff->end(), part0.begin(), part0.end());
  riff->insert(riff->end(), coded.begin(), coded.end());
  finish_riff(fp, hdr_pos, part0.size(), riff);
}
static inline void serialize_frame_tokens(const FramePlan& fp, const uint8_t* mb_hdr, const uint8_t* segment_map, const uint8_t* proba /*[1056]*/,
                                          const uint16_t* tokens, size_t n_tokens, std::vector<uint8_t>* riff) {
  const int total = fp.mb_w * fp.mb_h;
  const int num_skip = count_skips(mb_hdr, total);
  const int skip_proba = num_skip > 0 ? (total - num_skip) * 255 / total : 0;
  std::vector<uint8_t> part0;
  part0.reserve((size_t)total * 4 + 2048);
  emit_partition0(fp, mb_hdr, segment_map, proba, num_skip, skip_proba, &part0);
  const size_t hdr_pos = riff->size();
  riff->resize(hdr_pos + 20 + 10);
  riff->insert(riff->end(), part0.begin(), part0.end());
  std::vector<uint8_t> part;
  part.reserve(n_tokens / 4 + 4096);
  {
    BoolEnc bw(&part);
    for (size_t i = 0; i < n_tokens; ++i) { const uint32_t t = tokens[i]; bw.put((int)(t & 1), (int)(t >> 8)); }
    bw.finish();
  }
  riff->insert(riff->end(), part.begin(), part.end());
  finish_riff(fp, hdr_pos, part0.size(), riff);
}

// Multi-partition route (Partitions > 0): levels come back from the GPU and the host walks them, because the
// reference's partitioned emission depends on its per-MB token start table, stale entries included (see below).
static inline void serialize_frame(const FramePlan& fp, const uint8_t* mb_hdr /*[nmb][48]*/,
                                   const int16_t* mb_coeffs /*[nmb][400]*/, const uint8_t* segment_map,
                                   const uint32_t* stats /*[4][8][3][11][2]*/, std::vector<uint8_t>* riff) {
  const int mb_w = fp.mb_w, mb_h = fp.mb_h, total = mb_w * mb_h;
  // statistics come from the GPU (mb_stats_kernel == collectMBStats); only the skip count is taken here
  const int (*st)[8][3][11][2] = reinterpret_cast<const int (*)[8][3][11][2]>(stats);
  std::vector<uint32_t> top_nz(mb_w);
  std::vector<uint8_t> top_dc(mb_w);
  const int num_skip = count_skips(mb_hdr, total);
  const int skip_proba = num_skip > 0 ? (total - num_skip) * 255 / total : 0;
  // optimizeProba (encode_proba.go:117)
  uint8_t proba[4][8][3][11];
  memcpy(proba, kCoeffsProba0, sizeof(proba));
  bool any_update = false;
  for (int t = 0; t < 4; ++t)
    for (int b = 0; b < 8; ++b)
      for (int c = 0; c < 3; ++c)
        for (int p = 0; p < 11; ++p) {
          const int c0 = st[t][b][c][p][0], c1 = st[t][b][c][p][1], tot = c0 + c1;
          if (!tot) continue;
          const int new_p = c1 > 0 ? 255 - (int)((long long)c1 * 255 / tot) : 255;
          const int idx = ((t * 8 + b) * 3 + c) * 11 + p;
          const int old_p = kCoeffsProba0[idx];
          const uint8_t up = kCoeffsUpdateProba[idx];
          auto bc = [&](int pr) -> long long { pr = clampi(pr, 1, 255); return (long long)c1 * bit_cost(1, (uint8_t)pr) + (long long)c0 * bit_cost(0, (uint8_t)pr); };
          if (bc(old_p) + bit_cost(0, up) > bc(new_p) + bit_cost(1, up) + 8 * 256) { proba[t][b][c][p] = (uint8_t)new_p; any_update = true; }
        }
  std::vector<uint8_t> part0;
  part0.reserve((size_t)total * 4 + 2048);
  emit_partition0(fp, mb_hdr, segment_map, &proba[0][0][0][0], num_skip, skip_proba, &part0);
  const size_t hdr_pos = riff->size();
  riff->resize(hdr_pos + 20 + 10);
  riff->insert(riff->end(), part0.begin(), part0.end());
  {
    // The reference records (bit, prob) tokens with a per-MB start index that is only written for non-skipped MBs
    // (encode_token.go:90), then emits [start[i], start[i+1]) for the rows of each partition (encode_token.go:322-361).
    // Restated literally, stale entries included.  (With one partition this degenerates to "emit everything in order".)
    struct TokSink {
      std::vector<uint16_t>* v;
      void put(int bit, int prob) { v->push_back((uint16_t)((bit & 1) | (prob << 8))); }
    };
    std::vector<uint16_t> toks;
    toks.reserve((size_t)total * 64);
    std::vector<size_t> start((size_t)total + 1, 0);
    for (int pass = 0; pass < 2; ++pass) {  // the first recording (default probabilities) leaves the stale starts behind
      toks.clear();
      TokSink sink{&toks};
      uint8_t p0[4][8][3][11];
      memcpy(p0, kCoeffsProba0, sizeof(p0));
      const uint8_t (*pp)[8][3][11] = pass == 0 ? p0 : proba;
      for (int my = 0; my < mb_h; ++my) {
        uint32_t left_nz = 0;
        uint8_t left_dc = 0;
        if (my == 0) { std::fill(top_nz.begin(), top_nz.end(), 0u); std::fill(top_dc.begin(), top_dc.end(), 0); }
        for (int mx = 0; mx < mb_w; ++mx) {
          const int idx = my * mb_w + mx;
          const MBView m{mb_hdr + (size_t)idx * 48, mb_coeffs + (size_t)idx * 400};
          if (m.skip()) {
            top_nz[mx] = 0; left_nz = 0;
            if (m.mb_type() == 0) { top_dc[mx] = 0; left_dc = 0; }
            continue;
          }
          start[idx] = toks.size();
          walk_mb(m, &top_nz[mx], &left_nz, &top_dc[mx], &left_dc,
                  [&](const int16_t* c, int nz, int type, int first, int ctx) { code_block(sink, pp[type], c, nz, first, ctx > 2 ? 2 : ctx); });
        }
      }
      if (pass == 0 && !any_update) break;  // rerecord happens only when optimizeProba changed something
    }
    start[total] = toks.size();
    const int np = fp.num_parts < 1 ? 1 : fp.num_parts;
    std::vector<std::vector<uint8_t>> parts(np);
    for (int pi = 0; pi < np; ++pi) {
      BoolEnc bw(&parts[pi]);
      if (np == 1) {
        for (size_t t = 0; t < toks.size(); ++t) bw.put(toks[t] & 1, toks[t] >> 8);
      } else {
        for (int idx = 0; idx < total; ++idx) {
          if (((idx / mb_w) & (np - 1)) != pi) continue;
          for (size_t t = start[idx]; t < start[idx + 1]; ++t) bw.put(toks[t] & 1, toks[t] >> 8);
        }
      }
      bw.finish();
    }
    for (int pi = 0; pi + 1 < np; ++pi) {
      const size_t sz = parts[pi].size();
      riff->push_back((uint8_t)sz); riff->push_back((uint8_t)(sz >> 8)); riff->push_back((uint8_t)(sz >> 16));
    }
    for (int pi = 0; pi < np; ++pi) riff->insert(riff->end(), parts[pi].begin(), parts[pi].end());
  }
  finish_riff(fp, hdr_pos, part0.size(), riff);
}

// optimizeProba (encode_proba.go:117-156): compares against CoeffsProba0, only ever SETS entries of `proba`; returns the
// number of entries set.  Costs are int64 (Go int).
static inline int optimize_proba_host(const Stats st, uint8_t proba[4][8][3][11]) {
  int updates = 0;
  for (int t = 0; t < 4; ++t)
    for (int b = 0; b < 8; ++b)
      for (int c = 0; c < 3; ++c)
        for (int p = 0; p < 11; ++p) {
          const int c0 = st[t][b][c][p][0], c1 = st[t][b][c][p][1], tot = c0 + c1;
          if (!tot) continue;
          const int new_p = c1 > 0 ? 255 - (int)((long long)c1 * 255 / tot) : 255;
          const int idx = ((t * 8 + b) * 3 + c) * 11 + p;
          const int old_p = kCoeffsProba0[idx];
          const uint8_t up = kCoeffsUpdateProba[idx];
          auto bc = [&](int pr) -> long long { pr = clampi(pr, 1, 255); return (long long)c1 * bit_cost(1, (uint8_t)pr) + (long long)c0 * bit_cost(0, (uint8_t)pr); };
          if (bc(old_p) + bit_cost(0, up) > bc(new_p) + bit_cost(1, up) + 8 * 256) { proba[t][b][c][p] = (uint8_t)new_p; ++updates; }
        }
  return updates;
}

// Serial RD path with mid-stream probability refreshes (encode_frame.go:35-57), single token partition: tokens are recorded
// while the frame is encoded, each macroblock under the probabilities in force when it was encoded (tabs[k] from macroblock
// index tab_start[k] on); after a successful final optimizeProba they are all re-recorded under the final table (then
// n_tabs == 1).  Partition 0 carries part0_proba (what the decoder will use).
static inline void serialize_frame_tables(const FramePlan& fp, const uint8_t* mb_hdr, const int16_t* mb_coeffs, const uint8_t* segment_map,
                                          const uint8_t* part0_proba /*[1056]*/, int n_tabs, const int* tab_start, const uint8_t* const* tabs,
                                          std::vector<uint8_t>* riff) {
  const int mb_w = fp.mb_w, mb_h = fp.mb_h, total = mb_w * mb_h;
  const int num_skip = count_skips(mb_hdr, total);
  const int skip_proba = num_skip > 0 ? (total - num_skip) * 255 / total : 0;
  std::vector<uint8_t> part0;
  part0.reserve((size_t)total * 4 + 2048);
  emit_partition0(fp, mb_hdr, segment_map, part0_proba, num_skip, skip_proba, &part0);
  const size_t hdr_pos = riff->size();
  riff->resize(hdr_pos + 20 + 10);
  riff->insert(riff->end(), part0.begin(), part0.end());
  std::vector<uint8_t> part;
  part.reserve((size_t)total * 32 + 4096);
  {
    BoolEnc bw(&part);
    std::vector<uint32_t> top_nz(mb_w, 0u);
    std::vector<uint8_t> top_dc(mb_w, 0);
    int k = 0;
    for (int my = 0; my < mb_h; ++my) {
      uint32_t left_nz = 0;
      uint8_t left_dc = 0;
      for (int mx = 0; mx < mb_w; ++mx) {
        const int idx = my * mb_w + mx;
        while (k + 1 < n_tabs && idx >= tab_start[k + 1]) ++k;
        const uint8_t (*pp)[8][3][11] = reinterpret_cast<const uint8_t (*)[8][3][11]>(tabs[k]);
        const MBView m{mb_hdr + (size_t)idx * 48, mb_coeffs + (size_t)idx * 400};
        if (m.skip()) {
          top_nz[mx] = 0; left_nz = 0;
          if (m.mb_type() == 0) { top_dc[mx] = 0; left_dc = 0; }
          continue;
        }
        walk_mb(m, &top_nz[mx], &left_nz, &top_dc[mx], &left_dc,
                [&](const int16_t* c, int nz, int type, int first, int ctx) { code_block(bw, pp[type], c, nz, first, ctx > 2 ? 2 : ctx); });
      }
    }
    bw.finish();
  }
  riff->insert(riff->end(), part.begin(), part.end());
  finish_riff(fp, hdr_pos, part0.size(), riff);
}

// Folded per-(type, band, ctx, level) token costs of a probability table (vp8_dev.cuh CostTabs; TokenCostForCoeffs,
// encode_quant.go:170, with variableLevelCost :248 folded in up to level 67).
static inline void build_cost_tables(const uint8_t* proba /*[1056]*/, uint16_t* lc /*[4*8*3*68]*/, uint16_t* eobc /*[96]*/) {
  for (int tbc = 0; tbc < 4 * 8 * 3; ++tbc) {
    const uint8_t* p = proba + tbc * 11;
    eobc[tbc] = kEntropyCost[p[0]];
    const int not_eob = kEntropyCost[255 - p[0]];
    lc[tbc * 68] = (uint16_t)(not_eob + kEntropyCost[p[1]]);
    for (int v = 1; v < 68; ++v) {
      int pattern = kLevelCodes[2 * (v - 1)], bits = kLevelCodes[2 * (v - 1) + 1], cost = 0;
      for (int i = 2; pattern; ++i, bits >>= 1, pattern >>= 1)
        if (pattern & 1) cost += bit_cost(bits & 1, p[i]);
      lc[tbc * 68 + v] = (uint16_t)(not_eob + kEntropyCost[255 - p[1]] + cost);
    }
  }
}

// Serial-path serialiser (Method < 3: statLoop + encodeFrame, encode.go:1334-1400, encode_frame.go:15-108).  The GPU's
// mode decisions on this path do not depend on the coefficient probabilities, so one device pass gives the content
// of mbInfo for every reference pass; what the host restates is the evolution of enc.proba:
//   * statLoop pass 0 refreshes the probabilities before macroblocks M, 2M+1, 3M+2, ... (M = max(total >> 3, 96),
//     encode_frame.go:35-57) from statistics over the WHOLE mbInfo array (collectAllStats, encode_proba.go:171): the
//     macroblocks encoded so far plus ZERO-STATE entries (non-skipped I16, all levels zero) for the rest;
//   * every later refresh / end-of-pass optimisation / the final one sees the complete frame.  optimizeProba compares
//     against CoeffsProba0 and only sets entries, so with identical statistics it is idempotent: all of those collapse
//     into one application of the full-frame statistics (which the GPU already accumulated), and the tokens recorded
//     inline in the main pass equal the ones a re-recording would produce.
// tests/test_oracle.py::test_host_serialisers_reproduce_oracle_bytes checks this against the literal multi-pass oracle.
static inline void serialize_frame_serial(const FramePlan& fp, const uint8_t* mb_hdr, const int16_t* mb_coeffs, const uint8_t* segment_map,
                                          const uint32_t* full_stats /*[4][8][3][11][2] from the GPU*/, int passes, std::vector<uint8_t>* riff) {
  (void)passes;  // Pass only repeats idempotent work on this path (see above)
  const int mb_w = fp.mb_w, mb_h = fp.mb_h, total = mb_w * mb_h;
  uint8_t proba[4][8][3][11];
  memcpy(proba, kCoeffsProba0, sizeof(proba));
  static thread_local Stats pre, st;
  static const uint8_t zero_hdr[48] = {0};
  static const int16_t zero_coeffs[400] = {0};
  int max_count = total >> 3;
  if (max_count < 96) max_count = 96;
  {  // statLoop pass 0: prefix statistics over the real macroblocks, zero-state tail added at every refresh point
    memset(pre, 0, sizeof(pre));
    std::vector<uint32_t> tnz(mb_w, 0u), tnz2(mb_w);
    std::vector<uint8_t> tdc(mb_w, 0), tdc2(mb_w);
    uint32_t left_nz = 0;
    uint8_t left_dc = 0;
    int refresh_cnt = max_count;
    for (int idx = 0; idx < total; ++idx) {
      const int my = idx / mb_w, mx = idx - my * mb_w;
      if (mx == 0) { left_nz = 0; left_dc = 0; }
      if (--refresh_cnt < 0) {
        memcpy(st, pre, sizeof(st));
        tnz2 = tnz; tdc2 = tdc;
        uint32_t l2 = left_nz;
        uint8_t ld2 = left_dc;
        const MBView z{zero_hdr, zero_coeffs};
        for (int j = idx; j < total; ++j) {
          const int jx = j % mb_w;
          if (jx == 0) { l2 = 0; ld2 = 0; }
          walk_mb(z, &tnz2[jx], &l2, &tdc2[jx], &ld2,
                  [&](const int16_t* c, int nz, int type, int first, int ctx) { stat_block(c, nz, type, first, ctx > 2 ? 2 : ctx, st); });
        }
        optimize_proba_host(st, proba);
        refresh_cnt = max_count;
      }
      const MBView m{mb_hdr + (size_t)idx * 48, mb_coeffs + (size_t)idx * 400};
      if (m.skip()) {
        tnz[mx] = 0; left_nz = 0;
        if (m.mb_type() == 0) { tdc[mx] = 0; left_dc = 0; }
        continue;
      }
      walk_mb(m, &tnz[mx], &left_nz, &tdc[mx], &left_dc,
              [&](const int16_t* c, int nz, int type, int first, int ctx) { stat_block(c, nz, type, first, ctx > 2 ? 2 : ctx, pre); });
    }
  }
  // end of statLoop pass 0 == every later optimisation: the full-frame statistics
  optimize_proba_host(*reinterpret_cast<const Stats*>(full_stats), proba);
  // main pass: tokens with the (now stable) probabilities
  struct TokSink {
    std::vector<uint16_t>* v;
    void put(int bit, int prob) { v->push_back((uint16_t)((bit & 1) | (prob << 8))); }
  };
  std::vector<uint16_t> toks;
  toks.reserve((size_t)total * 64);
  std::vector<size_t> start((size_t)total + 1, 0);
  {
    TokSink sink{&toks};
    std::vector<uint32_t> tnz(mb_w, 0u);
    std::vector<uint8_t> tdc(mb_w, 0);
    for (int my = 0; my < mb_h; ++my) {
      uint32_t left_nz = 0;
      uint8_t left_dc = 0;
      for (int mx = 0; mx < mb_w; ++mx) {
        const int idx = my * mb_w + mx;
        const MBView m{mb_hdr + (size_t)idx * 48, mb_coeffs + (size_t)idx * 400};
        if (m.skip()) {
          tnz[mx] = 0; left_nz = 0;
          if (m.mb_type() == 0) { tdc[mx] = 0; left_dc = 0; }
          continue;
        }
        start[idx] = toks.size();
        walk_mb(m, &tnz[mx], &left_nz, &tdc[mx], &left_dc,
                [&](const int16_t* c, int nz, int type, int first, int ctx) { code_block(sink, proba[type], c, nz, first, ctx > 2 ? 2 : ctx); });
      }
    }
  }
  const int num_skip = count_skips(mb_hdr, total);
  const int skip_proba = num_skip > 0 ? (total - num_skip) * 255 / total : 0;
  start[total] = toks.size();
  std::vector<uint8_t> part0;
  part0.reserve((size_t)total * 4 + 2048);
  emit_partition0(fp, mb_hdr, segment_map, &proba[0][0][0][0], num_skip, skip_proba, &part0);
  const size_t hdr_pos = riff->size();
  riff->resize(hdr_pos + 20 + 10);
  riff->insert(riff->end(), part0.begin(), part0.end());
  const int np = fp.num_parts < 1 ? 1 : fp.num_parts;
  std::vector<std::vector<uint8_t>> parts(np);
  for (int pi = 0; pi < np; ++pi) {
    BoolEnc bw(&parts[pi]);
    if (np == 1) {
      for (size_t t = 0; t < toks.size(); ++t) bw.put(toks[t] & 1, toks[t] >> 8);
    } else {  // EmitTokensPartitioned with the reference's per-MB start table (encode_token.go:322-361)
      for (int idx = 0; idx < total; ++idx) {
        if (((idx / mb_w) & (np - 1)) != pi) continue;
        for (size_t t = start[idx]; t < start[idx + 1]; ++t) bw.put(toks[t] & 1, toks[t] >> 8);
      }
    }
    bw.finish();
  }
  for (int pi = 0; pi + 1 < np; ++pi) {
    const size_t sz = parts[pi].size();
    riff->push_back((uint8_t)sz); riff->push_back((uint8_t)(sz >> 8)); riff->push_back((uint8_t)(sz >> 16));
  }
  for (int pi = 0; pi < np; ++pi) riff->insert(riff->end(), parts[pi].begin(), parts[pi].end());
  finish_riff(fp, hdr_pos, part0.size(), riff);
}

}  // namespace wgh
