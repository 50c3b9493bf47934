// ORACLE -- TEST INFRASTRUCTURE ONLY (see vp8_common.h).
// Restates the reference VP8 lossy encoder on the path taken by webp.Encode with
// Method>=3, single pass, GOMAXPROCS>1, mbH>=4 ("parallel path", internal/lossy/encode.go:1356):
//   importImage            internal/lossy/encode.go:671
//   analysis / segments    internal/lossy/encode_analysis.go:29-903, encode.go:1012-1320
//   mode search + residual internal/lossy/encode_parallel.go:168-1496
//   quantize / trellis     internal/lossy/encode_quant.go, encode_trellis.go
//   tokens / proba / emit  internal/lossy/encode_token.go, encode_proba.go, encode_syntax.go,
//                          internal/bitio/writer_bool.go
// Macroblocks are visited in raster order, which satisfies the same left/top/top-right
// dependencies as the reference's row-pipelined workers (results are order-independent).
#pragma once
#include "dsp.h"
#include "sharpyuv.h"
#include <vector>

namespace orc {

// ------------------------------------------------------------------ writer_bool.go
struct BoolWriter {
  int32_t range = 254, value = 0;
  int run = 0, nb_bits = -8;
  std::vector<uint8_t> buf;
  static int knorm(int r) { int s = 0; while (((r + 1) << s) < 128) ++s; return r == 0 ? 7 : s; }
  void flush() {
    const int s = 8 + nb_bits;
    const int32_t bits = value >> s;
    value -= bits << s;
    nb_bits -= 8;
    if ((bits & 0xff) != 0xff) {
      if (bits & 0x100) { if (!buf.empty()) buf.back()++; }
      if (run > 0) {
        const uint8_t val = (bits & 0x100) ? 0x00 : 0xff;
        for (; run > 0; --run) buf.push_back(val);
      }
      buf.push_back((uint8_t)(bits & 0xff));
    } else {
      run++;
    }
  }
  int put_bit(int bit, int prob) {
    const int32_t split = (range * prob) >> 8;
    if (bit) { value += split + 1; range -= split + 1; } else { range = split; }
    if (range < 127) {
      const int shift = knorm(range);
      range = ((range + 1) << shift) - 1;
      value <<= shift;
      nb_bits += shift;
      if (nb_bits > 0) flush();
    }
    return bit;
  }
  int put_bit_uniform(int bit) {
    const int32_t split = range >> 1;
    if (bit) { value += split + 1; range -= split + 1; } else { range = split; }
    if (range < 127) {
      range = ((range + 1) << knorm(range)) - 1;  // kNewRange
      value <<= 1;
      nb_bits += 1;
      if (nb_bits > 0) flush();
    }
    return bit;
  }
  void put_bits(uint32_t v, int n) {
    for (uint32_t mask = 1u << (n - 1); mask; mask >>= 1) put_bit_uniform((v & mask) ? 1 : 0);
  }
  void put_signed_bits(int v, int n) {
    if (!put_bit_uniform(v != 0)) return;
    if (v < 0) put_bits(((uint32_t)(-v) << 1) | 1, n + 1); else put_bits((uint32_t)v << 1, n + 1);
  }
  std::vector<uint8_t>& finish() {
    put_bits(0, 9 - nb_bits);
    nb_bits = 0;
    flush();
    return buf;
  }
};

// ------------------------------------------------------------------ config / per-segment data
struct EncodeConfig {  // internal/lossy/encode.go:46-86 (DefaultConfig)
  int quality = 75, method = 4, sns_strength = 50, filter_strength = 60, filter_sharpness = 0;
  int filter_type = 1, partitions = 0, segments = 4, pass = 1, preprocessing = 0;
  bool force_serial = false;  // GOMAXPROCS == 1 semantics: serial encodeFrame even where the reference would go row-parallel
  int dither_amp = 0;  // VP8Random.amp = int(256 * Dithering) (dsp/random.go:39-50); 0 == no dithering
  int target_size = 0;      // bytes; > 0 -> size search (doSearch, encode.go:1338)
  float target_psnr = 0.f;  // dB; > 0 -> "PSNR search" (the measured PSNR is always 99.0)
  int qmin = 0, qmax = 100; // resolveQMax(-1) == 100 (encode.go:305-306)
  bool use_sharp_yuv = false;  // EncoderOptions.UseSharpYUV: planes from sharpyuv.Convert + importYCbCr (encode.go:531-535)
};
struct SegmentQuant {  // encode.go:311-323
  int quant, iquant, bias, dc_quant, dc_iquant, dc_bias;
  int16_t sharpen[16];
};
struct SegmentInfo {  // encode.go:278-306
  SegmentQuant y1, y2, uv;
  int lambda_i4, lambda_i16, lambda_uv, lambda_mode, tlambda_i4, tlambda_i16, tlambda_uv, tlambda_sd;
  int quant, fstrength, alpha, beta;
};
struct MBInfo {  // encode.go:241-275
  int mb_type = 0;  // 0 = i16, 1 = i4
  uint8_t uv_mode = 0, segment = 0, skip = 0, i16_mode = 0;
  int alpha = 0;
  uint8_t modes[16] = {0};
  int16_t coeffs[400] = {0};
  uint32_t non_zero_y = 0, non_zero_uv = 0;
  uint8_t nz_y[16] = {0}, nz_uv[8] = {0}, nz_dc = 0;
};
typedef int ProbaStats[NUM_TYPES][NUM_BANDS][NUM_CTX][NUM_PROBAS][2];

static const int kFreqSharpening[16] = {0, 30, 60, 90, 30, 60, 90, 90, 60, 90, 90, 90, 90, 90, 90, 90};
static const int kBiasMatrices[3][2] = {{96, 110}, {96, 108}, {110, 115}};
static const int kModeFixedCost16[4] = {663, 919, 872, 919};  // encode_analysis.go:1481
static const int kModeFixedCostUV[4] = {302, 984, 439, 642};  // encode_analysis.go:1485
static const int kWeightTrellis[16] = {30, 27, 19, 11, 27, 24, 17, 10, 19, 17, 12, 8, 11, 10, 8, 6};
static const int kReverseZigzag[16] = {0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15};

// ------------------------------------------------------------------ quantization (encode_quant.go)
static inline int quantize_coeffs(const int16_t* in, int16_t* out, const SegmentQuant* sq, int first) {
  ORC_COUNT(OP_QUANT_COEFF, 16 - first);
  int max_zz = -1;
  if (first == 0) {
    int v = in[0];
    int sign = 1;
    if (v < 0) { sign = -1; v = -v; }
    v += sq->sharpen[0];
    if (v < 0) v = 0;
    int coeff = (int)(((uint32_t)v * (uint32_t)sq->dc_iquant + (uint32_t)sq->dc_bias) >> 17);
    if (coeff > 2047) coeff = 2047;
    out[0] = (int16_t)(sign * coeff);
    if (coeff) max_zz = 0;
  } else {
    out[0] = 0;
  }
  const uint32_t iq = (uint32_t)sq->iquant, bias = (uint32_t)sq->bias;
  for (int n = 1; n < 16; ++n) {
    int v = in[n];
    int sign = 1;
    if (v < 0) { sign = -1; v = -v; }
    v += sq->sharpen[n];
    if (v < 0) v = 0;
    int coeff = (int)(((uint32_t)v * iq + bias) >> 17);
    if (coeff > 2047) coeff = 2047;
    out[n] = (int16_t)(sign * coeff);
    if (coeff && kReverseZigzag[n] > max_zz) max_zz = kReverseZigzag[n];
  }
  return max_zz + 1;
}
static inline void dequant_coeffs(const int16_t* in, int16_t* out, const SegmentQuant* sq) {
  ORC_COUNT(OP_DEQUANT_BLOCK, 1);
  out[0] = (int16_t)(in[0] * sq->dc_quant);
  for (int n = 1; n < 16; ++n) out[n] = (int16_t)(in[n] * sq->quant);
}
static inline uint64_t rd_score(int disto, int rate, int lambda) {
  ORC_COUNT(OP_MODE_SCORE, 1);
  return (uint64_t)(int64_t)rate * (uint64_t)(int64_t)lambda + 256 * (uint64_t)(int64_t)disto;
}
static inline int variable_level_cost(int level, const uint8_t* probas) {  // encode_quant.go:248
  int idx = level - 1;
  if (idx >= 67) idx = 66;
  int pattern = kLevelCodes[2 * idx], bits = kLevelCodes[2 * idx + 1];
  int cost = 0;
  for (int i = 2; pattern; ++i) {
    if (pattern & 1) cost += bit_cost(bits & 1, probas[i]);
    bits >>= 1;
    pattern >>= 1;
  }
  return cost;
}
// TokenCostForCoeffs (encode_quant.go:170)
static inline int token_cost(const int16_t* coeffs, int nz_count, int type, const Proba* proba, int ctx0, int first) {
  if (nz_count <= first) return kEntropyCost[proba->bands[type][kBands[first]][ctx0][0]];
  const int last = nz_count - 1;
  int cost = 0, ctx = ctx0;
  ORC_COUNT(OP_TOKEN_COEFF, (last < 15 ? last + 2 : 16) - first);
  for (int n = first; n < 16; ++n) {
    const uint8_t* pp = proba->bands[type][kBands[n]][ctx];
    int v = coeffs[kZigzag[n]];
    if (v < 0) v = -v;
    if (n > last) { cost += kEntropyCost[pp[0]]; break; }
    cost += kEntropyCost[255 - pp[0]];
    if (v == 0) {
      cost += kEntropyCost[pp[1]];
      ctx = 0;
    } else {
      cost += kEntropyCost[255 - pp[1]];
      if (v == 1) {
        cost += kLevelFixedCosts[1] + kEntropyCost[pp[2]];
        ctx = 1;
      } else if (v == 2) {
        cost += kLevelFixedCosts[2] + kEntropyCost[255 - pp[2]] + kEntropyCost[pp[3]] + kEntropyCost[pp[4]];
        ctx = 2;
      } else {
        cost += kLevelFixedCosts[v] + variable_level_cost(v, pp);
        ctx = 2;
      }
    }
  }
  return cost;
}

// TrellisQuantizeBlock (encode_trellis.go:23-324).  in/out may alias.
static inline int trellis_quantize_block(const int16_t* in_, int16_t* out, const SegmentQuant* sq, int first,
                                         int type, int initial_ctx, const Proba* proba, int lambda) {
  {  // all-zero pre-scan with neutral bias (encode_trellis.go:39-98)
    bool non_zero = false;
    for (int n = first; n < 16 && !non_zero; ++n) {
      ORC_COUNT(OP_TRELLIS_PRESCAN_COEFF, 1);
      int raw = in_[kZigzag[n]];
      if (raw < 0) raw = -raw;
      int c = raw + sq->sharpen[kZigzag[n]];
      if (c < 0) c = 0;
      const int iq = (n == 0) ? sq->dc_iquant : sq->iquant;
      if (((int64_t)c * iq >> 17) > 0) non_zero = true;
    }
    if (!non_zero) {
      for (int i = 0; i < 16; ++i) out[i] = 0;
      return 0;
    }
  }
  int16_t in[16];
  memcpy(in, in_, sizeof(in));
  for (int i = 0; i < 16; ++i) out[i] = 0;
  if (initial_ctx > 2) initial_ctx = 2;
  struct State { int64_t score; int16_t level; int prev_ctx; bool valid; };
  struct Path { int16_t level; int prev_ctx; bool valid; };
  State prev[3], curr[3];
  Path path[16][3];
  memset(path, 0, sizeof(path));
  for (int c = 0; c < 3; ++c) prev[c] = State{0, 0, 0, false};
  prev[initial_ctx] = State{0, 0, 0, true};
  const int first_band = kBands[first];
  const int skip_rate = bit_cost(0, proba->bands[type][first_band][initial_ctx][0]);
  int64_t best_terminal = (int64_t)skip_rate * lambda;
  int best_last_n = -1, best_last_ctx = -1;
  const int64_t lam = lambda;
  for (int n = first; n < 16; ++n) {
    ORC_COUNT(OP_TRELLIS_POS, 1);
    const int zig = kZigzag[n];
    const int band = kBands[n + 1];  // sic: next position's band (encode_trellis.go:151)
    int raw = in[zig];
    int sign = 1;
    if (raw < 0) { sign = -1; raw = -raw; }
    int coeff0 = raw + sq->sharpen[zig];
    if (coeff0 < 0) coeff0 = 0;
    const int quant = (n == 0) ? sq->dc_quant : sq->quant;
    const int iquant = (n == 0) ? sq->dc_iquant : sq->iquant;
    int L0 = (int)(((int64_t)coeff0 * iquant) >> 17);
    if (L0 > 2047) L0 = 2047;
    int thresh_level = (int)(((uint32_t)coeff0 * (uint32_t)iquant + 65536u) >> 17);
    if (thresh_level > 2047) thresh_level = 2047;
    const int64_t weight = kWeightTrellis[zig];
    const int64_t coeff0sq = (int64_t)coeff0 * coeff0;
    const uint8_t(*band_probas)[NUM_PROBAS] = proba->bands[type][band];
    const int64_t kMaxScore = (int64_t)1 << 60;
    for (int c = 0; c < 3; ++c) { curr[c].valid = false; curr[c].score = kMaxScore; }
    const bool has_l0 = L0 > 0 && L0 <= thresh_level;
    const bool has_l1 = L0 + 1 <= 2047 && L0 + 1 <= thresh_level;
    int64_t delta_d0 = 0, delta_d1 = 0;
    int next_ctx0 = 0, next_ctx1 = 0, fixed_l0 = 0, fixed_l1 = 0;
    int16_t signed_l0 = 0, signed_l1 = 0;
    if (has_l0) {
      const int64_t new_err = coeff0 - (int64_t)L0 * quant;
      delta_d0 = weight * (new_err * new_err - coeff0sq);
      next_ctx0 = L0 > 2 ? 2 : L0;
      signed_l0 = (int16_t)(sign * L0);
      fixed_l0 = kLevelFixedCosts[L0];
    }
    if (has_l1) {
      const int L1 = L0 + 1;
      const int64_t new_err = coeff0 - (int64_t)L1 * quant;
      delta_d1 = weight * (new_err * new_err - coeff0sq);
      next_ctx1 = L1 > 2 ? 2 : L1;
      signed_l1 = (int16_t)(sign * L1);
      fixed_l1 = kLevelFixedCosts[L1];
    }
    const int64_t disto_l0 = 256 * delta_d0, disto_l1 = 256 * delta_d1;
    for (int pc = 0; pc < 3; ++pc) {
      if (!prev[pc].valid) continue;
      ORC_COUNT(OP_TRELLIS_TRANS, 1 + (has_l0 ? 1 : 0) + (has_l1 ? 1 : 0));
      const int64_t prev_score = prev[pc].score;
      const uint8_t* p = band_probas[pc];
      const int not_eob = kEntropyCost[255 - p[0]];
      const int rate0 = not_eob + kEntropyCost[p[1]];
      const int64_t total = prev_score + (int64_t)rate0 * lam;
      if (!curr[0].valid || total < curr[0].score) curr[0] = State{total, 0, pc, true};
      if (has_l0 || has_l1) {
        const int non_zero = not_eob + kEntropyCost[255 - p[1]];
        if (has_l0) {
          const int rate = non_zero + fixed_l0 + variable_level_cost(L0, p);
          const int64_t ts = prev_score + (int64_t)rate * lam + disto_l0;
          if (!curr[next_ctx0].valid || ts < curr[next_ctx0].score) curr[next_ctx0] = State{ts, signed_l0, pc, true};
        }
        if (has_l1) {
          const int rate = non_zero + fixed_l1 + variable_level_cost(L0 + 1, p);
          const int64_t ts = prev_score + (int64_t)rate * lam + disto_l1;
          if (!curr[next_ctx1].valid || ts < curr[next_ctx1].score) curr[next_ctx1] = State{ts, signed_l1, pc, true};
        }
      }
    }
    for (int c = 0; c < 3; ++c)
      if (curr[c].valid) path[n][c] = Path{curr[c].level, curr[c].prev_ctx, true};
    for (int c = 1; c < 3; ++c) {
      if (!curr[c].valid) continue;
      ORC_COUNT(OP_TRELLIS_TERMINAL, 1);
      int64_t eob_score = curr[c].score;
      if (n < 15) eob_score += (int64_t)kEntropyCost[proba->bands[type][band][c][0]] * lam;
      if (eob_score < best_terminal) {
        best_terminal = eob_score;
        best_last_n = n;
        best_last_ctx = c;
      }
    }
    for (int c = 0; c < 3; ++c) prev[c] = curr[c];
  }
  if (best_last_n < 0) return 0;
  int ctx = best_last_ctx, last = 0;
  for (int n = best_last_n; n >= first; --n) {
    ORC_COUNT(OP_TRELLIS_BACKTRACK, 1);
    if (path[n][ctx].valid) {
      const int zig = kZigzag[n];
      out[zig] = path[n][ctx].level;
      if (out[zig] != 0 && last == 0) last = n + 1;
      ctx = path[n][ctx].prev_ctx;
    }
  }
  for (int n = 0; n < first; ++n) out[kZigzag[n]] = 0;
  return last;
}

// ------------------------------------------------------------------ the encoder
struct Encoder {
  EncodeConfig cfg;
  int width = 0, height = 0, mb_w = 0, mb_h = 0;
  std::vector<uint8_t> y_plane, u_plane, v_plane;  // padded source, overwritten by reconstruction
  int y_stride = 0, uv_stride = 0;
  std::vector<MBInfo> mb_info;
  SegmentInfo dqm[4];
  Proba proba;
  std::vector<Proba> proba_hist;   // test taps: probability tables in force over the last serial pass, from macroblock hist_starts[k] on
  std::vector<int> hist_starts;
  Proba proba_pre_final;           // state before the final optimizeProba
  // segment / filter headers
  bool seg_use = false, seg_update_map = false;
  int8_t seg_quantizer[4] = {0, 0, 0, 0}, seg_fstrength[4] = {0, 0, 0, 0};
  bool f_simple = false;
  int f_level = 0, f_sharpness = 0;
  int num_parts = 1, num_segments = 1, base_quant = 0;
  int dq_y1_dc = 0, dq_y2_dc = 0, dq_y2_ac = 0, dq_uv_dc = 0, dq_uv_ac = 0;
  int global_uv_alpha = 0;
  uint8_t skip_proba = 0;
  int num_skip = 0;
  uint16_t fixed_costs_i4[10][10][10];
  // tokens: bit | prob<<8
  std::vector<uint16_t> tokens;
  std::vector<size_t> mb_start;
  // per-MB work buffers (encode_parallel.go:118-143)
  uint8_t yuv_in[YUV_SIZE], yuv_out[YUV_SIZE], yuv_out2[YUV_SIZE];
  int16_t tmp_best_dq[16], tmp_best_q[16];
  int tmp_best_nz = 0;
  // optional taps for tests
  std::vector<uint8_t> src_y, src_u, src_v;  // copy of the imported planes (before reconstruction)
  std::vector<uint8_t> alphas;                // per-MB mixed alpha before clustering

  // ---- encode_analysis.go:1497-1533 + encode_syntax.go:474
  static bool subtree_contains(int node_or_leaf, int mode) {
    if (node_or_leaf <= 0) return -node_or_leaf == mode;
    return subtree_contains(kYModesIntra4[2 * node_or_leaf], mode) ||
           subtree_contains(kYModesIntra4[2 * node_or_leaf + 1], mode);
  }
  void compute_fixed_costs_i4() {
    for (int top = 0; top < 10; ++top)
      for (int left = 0; left < 10; ++left) {
        const uint8_t* prob = &kBModesProba[(top * 10 + left) * 9];
        for (int mode = 0; mode < 10; ++mode) {
          int cost = 0;
          int bit = subtree_contains(kYModesIntra4[0], mode) ? 0 : 1;
          cost += bit_cost(bit, prob[0]);
          int i = kYModesIntra4[bit];
          while (i > 0) {
            bit = subtree_contains(kYModesIntra4[2 * i], mode) ? 0 : 1;
            cost += bit_cost(bit, prob[i]);
            i = kYModesIntra4[2 * i + bit];
          }
          fixed_costs_i4[top][left][mode] = (uint16_t)cost;
        }
      }
  }

  // ---- encode.go:452-494 (NewEncoder) with importImage encode.go:671-942 for *image.RGBA/NRGBA
  void init(const uint8_t* rgba, int stride, int w, int h, const EncodeConfig& c, int has_alpha) {
    cfg = c;
    width = w; height = h;
    mb_w = (w + 15) >> 4; mb_h = (h + 15) >> 4;
    num_parts = 1 << cfg.partitions;
    if (num_parts > 8) num_parts = 8;
    y_stride = mb_w * 16; uv_stride = mb_w * 8;
    y_plane.assign((size_t)y_stride * mb_h * 16, 0);
    u_plane.assign((size_t)uv_stride * mb_h * 8, 0);
    v_plane.assign((size_t)uv_stride * mb_h * 8, 0);
    mb_info.assign((size_t)mb_w * mb_h, MBInfo());
    memset(dqm, 0, sizeof(dqm));
    memset(top_derr, 0, sizeof(top_derr));
    memset(left_derr, 0, sizeof(left_derr));
    compute_fixed_costs_i4();
    import_image(rgba, stride, has_alpha);
    init_segments();
    reset_proba(&proba);
    tokens.clear();
    mb_start.assign((size_t)mb_w * mb_h + 1, 0);
  }

  // VP8Random (internal/dsp/random.go:17-79): subtractive generator over a 55-entry table, lags 55 / 24
  struct Random {
    uint32_t tab[55];
    int i1 = 0, i2 = 31, amp = 0;
    explicit Random(int a) : amp(a) { memcpy(tab, kRandomTable, sizeof(tab)); }
    int bits(int num_bits) {
      int64_t diff = (int64_t)tab[i1] - (int64_t)tab[i2];
      if (diff < 0) diff += (int64_t)1 << 31;
      tab[i1] = (uint32_t)diff;
      if (++i1 == 55) i1 = 0;
      if (++i2 == 55) i2 = 0;
      int d = (int)((int32_t)((uint32_t)diff << 1) >> (32 - num_bits));
      d = (d * amp) >> 8;
      return d + (1 << (num_bits - 1));
    }
  };
  // NewEncoderFromYUV's importYCbCr (internal/lossy/encode.go:544-585): sharp planes, edges replicated up to the macroblock grid
  void import_sharp(const uint8_t* pix, int stride) {
    const int w = width, h = height, pad_w = mb_w * 16, pad_h = mb_h * 16, uv_w = (w + 1) >> 1, uv_h = (h + 1) >> 1;
    std::vector<uint8_t> ty((size_t)w * h), tu((size_t)uv_w * uv_h), tv((size_t)uv_w * uv_h);
    sharp::convert(pix, stride, w, h, ty.data(), w, tu.data(), tv.data(), uv_w);
    for (int y = 0; y < pad_h; ++y)
      for (int x = 0; x < pad_w; ++x) y_plane[(size_t)y * y_stride + x] = ty[(size_t)(y >= h ? h - 1 : y) * w + (x >= w ? w - 1 : x)];
    for (int y = 0; y < pad_h / 2; ++y)
      for (int x = 0; x < pad_w / 2; ++x) {
        const size_t s = (size_t)(y >= uv_h ? uv_h - 1 : y) * uv_w + (x >= uv_w ? uv_w - 1 : x);
        u_plane[(size_t)y * uv_stride + x] = tu[s];
        v_plane[(size_t)y * uv_stride + x] = tv[s];
      }
    src_y = y_plane; src_u = u_plane; src_v = v_plane;
  }
  void import_image(const uint8_t* pix, int stride, int has_alpha) {
    if (cfg.use_sharp_yuv) return import_sharp(pix, stride);
    const int w = width, h = height, pad_w = mb_w * 16, pad_h = mb_h * 16;
    Random rg(cfg.dither_amp);
    const bool dither = cfg.dither_amp > 0;  // amp 0 gives the fixed rounding back, so "Dithering > 0 but amp == 0" is the same
    for (int y = 0; y < pad_h; ++y) {  // encode.go:757-792
      const int sy = y >= h ? h - 1 : y;
      const uint8_t* row = pix + (size_t)sy * stride;
      uint8_t* dst = &y_plane[(size_t)y * y_stride];
      if (dither) {  // serial dithered branch (encode.go:793-809): every padded pixel draws its own random rounding
        for (int x = 0; x < pad_w; ++x) {
          const int sx = x >= w ? w - 1 : x;
          dst[x] = (uint8_t)((16839 * row[4 * sx] + 33059 * row[4 * sx + 1] + 6420 * row[4 * sx + 2] + rg.bits(16) + (16 << 16)) >> 16);
        }
        continue;
      }
      for (int x = 0; x < w; ++x) dst[x] = rgb_to_y(row[4 * x], row[4 * x + 1], row[4 * x + 2]);
      for (int x = w; x < pad_w; ++x) dst[x] = dst[w - 1];
    }
    const GammaTables& gt = gamma_tables();
    const int uv_width = (pad_w + 1) >> 1;
    std::vector<uint8_t> pr(2 * pad_w), pg(2 * pad_w), pb(2 * pad_w), pa(2 * pad_w, 0xff);
    for (int y = 0; y < pad_h / 2; ++y) {  // encode.go:836-902
      for (int r = 0; r < 2; ++r) {
        int sy = 2 * y + r;
        if (sy >= h) sy = h - 1;
        const uint8_t* row = pix + (size_t)sy * stride;
        for (int x = 0; x < pad_w; ++x) {
          const int sx = x >= w ? w - 1 : x;
          pr[r * pad_w + x] = row[4 * sx];
          pg[r * pad_w + x] = row[4 * sx + 1];
          pb[r * pad_w + x] = row[4 * sx + 2];
          if (has_alpha) pa[r * pad_w + x] = row[4 * sx + 3];
        }
      }
      // AccumulateRGBA (yuv.go:486) + ConvertRGBA32ToUV (yuv.go:553); pad_w is even.
      for (int i = 0; i < uv_width; ++i) {
        const int j = 2 * i;
        const uint32_t total_a = pa[j] + pa[j + 1] + pa[pad_w + j] + pa[pad_w + j + 1];
        int rv, gv, bv;
        if (total_a == 4 * 0xff || total_a == 0) {
#define SUM4(p) (gt.gamma_to_linear[p[j]] + gt.gamma_to_linear[p[j + 1]] + gt.gamma_to_linear[p[pad_w + j]] + \
                 gt.gamma_to_linear[p[pad_w + j + 1]])
          rv = linear_to_gamma(SUM4(pr), 0);
          gv = linear_to_gamma(SUM4(pg), 0);
          bv = linear_to_gamma(SUM4(pb), 0);
#undef SUM4
        } else {
          const uint8_t al[4] = {pa[j], pa[j + 1], pa[pad_w + j], pa[pad_w + j + 1]};
          const uint8_t sr[4] = {pr[j], pr[j + 1], pr[pad_w + j], pr[pad_w + j + 1]};
          const uint8_t sg[4] = {pg[j], pg[j + 1], pg[pad_w + j], pg[pad_w + j + 1]};
          const uint8_t sb[4] = {pb[j], pb[j + 1], pb[pad_w + j], pb[pad_w + j + 1]};
          rv = linear_to_gamma_weighted(sr, al, total_a);
          gv = linear_to_gamma_weighted(sg, al, total_a);
          bv = linear_to_gamma_weighted(sb, al, total_a);
        }
        // AccumulateRGBA stores uint16 (yuv.go:516)
        rv = (uint16_t)rv; gv = (uint16_t)gv; bv = (uint16_t)bv;
        // ConvertRGBA32ToUV / ...Dithered (yuv.go:553-576): U draws before V
        const int ru = dither ? rg.bits(18) : (1 << 17);
        u_plane[(size_t)y * uv_stride + i] = rgb_to_u(rv, gv, bv, ru);
        const int rvv = dither ? rg.bits(18) : (1 << 17);
        v_plane[(size_t)y * uv_stride + i] = rgb_to_v(rv, gv, bv, rvv);
      }
    }
    src_y = y_plane; src_u = u_plane; src_v = v_plane;
  }

  // ---- encode.go:1039-1063
  static double quality_to_compression(int quality) {
    if (quality <= 0) return 0.0;
    if (quality >= 100) return 1.0;
    const double c = (double)quality / 100.0;
    const double linear_c = (c < 0.75) ? c * (2.0 / 3.0) : 2.0 * c - 1.0;
    return pow(linear_c, 1.0 / 3.0);
  }
  static int quality_to_qindex(int quality) {
    return clampi((int)(127.0 * (1.0 - quality_to_compression(quality))), 0, 127);
  }
  void init_segments() {  // encode.go:1012
    const int q = quality_to_qindex(cfg.quality);
    num_segments = clampi(cfg.segments, 1, 4);
    dq_uv_dc = dq_uv_ac = 0;
    for (int i = 0; i < 4; ++i) setup_segment(i, q);
  }
  static void init_segment_quant(SegmentQuant* sq, int dc_quant, int ac_quant, int bias_type) {  // encode.go:1169
    sq->dc_quant = dc_quant;
    sq->dc_iquant = (1 << 17) / dc_quant;
    sq->dc_bias = kBiasMatrices[bias_type][0] << 9;
    sq->quant = ac_quant;
    sq->iquant = (1 << 17) / ac_quant;
    sq->bias = kBiasMatrices[bias_type][1] << 9;
  }
  void setup_segment(int idx, int q) {  // encode.go:1084
    SegmentInfo* seg = &dqm[idx];
    seg->quant = q;
    const int y1dc = kDcTable[clampi(q + dq_y1_dc, 0, 127)];
    const int y1ac = kAcTable[clampi(q, 0, 127)];
    init_segment_quant(&seg->y1, y1dc, y1ac, 0);
    int y2dc = kDcTable[clampi(q + dq_y2_dc, 0, 127)] * 2;
    if (y2dc < 8) y2dc = 8;
    const int y2ac = kAcTable2[clampi(q + dq_y2_ac, 0, 127)];
    init_segment_quant(&seg->y2, y2dc, y2ac, 1);
    const int uvdc = kDcTable[clampi(q + dq_uv_dc, 0, 117)];
    const int uvac = kAcTable[clampi(q + dq_uv_ac, 0, 127)];
    init_segment_quant(&seg->uv, uvdc, uvac, 2);
    const int q_i4 = (y1dc + 15 * y1ac + 8) >> 4;
    const int q_i16 = (y2dc + 15 * y2ac + 8) >> 4;
    const int q_uv = (uvdc + 15 * uvac + 8) >> 4;
#define MAX1(x) ((x) > 1 ? (x) : 1)
    seg->lambda_i4 = MAX1((3 * q_i4 * q_i4) >> 7);
    seg->lambda_i16 = MAX1(3 * q_i16 * q_i16);
    seg->lambda_uv = MAX1((3 * q_uv * q_uv) >> 6);
    seg->lambda_mode = MAX1((1 * q_i4 * q_i4) >> 7);
    seg->tlambda_i4 = MAX1((7 * q_i4 * q_i4) >> 3);
    seg->tlambda_i16 = MAX1((q_i16 * q_i16) >> 2);
    seg->tlambda_uv = MAX1((q_uv * q_uv) << 1);
#undef MAX1
    seg->tlambda_sd = (cfg.method >= 4 && cfg.sns_strength > 0) ? (cfg.sns_strength * q_i4) >> 5 : 0;
    for (int i = 0; i < 16; ++i) {
      const int qq = (i == 0) ? seg->y1.dc_quant : seg->y1.quant;
      seg->y1.sharpen[i] = (int16_t)((kFreqSharpening[i] * qq) >> 11);
      seg->y2.sharpen[i] = 0;
      seg->uv.sharpen[i] = 0;
    }
  }

  // ---- analysis (encode_analysis.go:29-903)
  static int alpha_from_histogram(const int* distribution) {  // encode_analysis.go:580-598
    int max_value = 0, last_non_zero = 1;
    for (int k = 0; k <= 31; ++k)
      if (distribution[k] > 0) {
        if (distribution[k] > max_value) max_value = distribution[k];
        last_non_zero = k;
      }
    int alpha = 0;
    if (max_value > 1) alpha = 2 * 255 * last_non_zero / max_value;
    return alpha > 255 ? 255 : alpha;
  }
  static void histo_add(const int16_t* c, int* distribution) {
    for (int k = 0; k < 16; ++k) {
      int v = abs((int)c[k]) >> 3;
      if (v > 31) v = 31;
      distribution[v]++;
    }
  }
  int compute_mb_alpha(int mx, int my) {  // encode_analysis.go:407-548
    uint8_t src[16 * BPS], pred[16 * BPS];
    const int x0 = mx * 16, y0 = my * 16;
    for (int j = 0; j < 16; ++j) {
      const int sy = (y0 + j >= height) ? height - 1 : y0 + j;
      for (int i = 0; i < 16; ++i) {
        const int sx = (x0 + i >= width) ? width - 1 : x0 + i;
        src[j * BPS + i] = y_plane[(size_t)sy * y_stride + sx];
      }
    }
    int best_alpha = 256;
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == TM_PRED && (mx == 0 || my == 0)) continue;
      if (mode == DC_PRED) {
        int dc_val = 128, sum = 0, count = 0;
        if (my > 0)
          for (int i = 0; i < 16; ++i) {
            const int sx = (x0 + i >= width) ? width - 1 : x0 + i;
            sum += y_plane[(size_t)(y0 - 1) * y_stride + sx];
            count++;
          }
        if (mx > 0)
          for (int j = 0; j < 16; ++j) {
            const int sy = (y0 + j >= height) ? height - 1 : y0 + j;
            sum += y_plane[(size_t)sy * y_stride + x0 - 1];
            count++;
          }
        if (count > 0) dc_val = (sum + count / 2) / count;
        for (int j = 0; j < 16; ++j) memset(pred + j * BPS, dc_val, 16);
      } else {
        int top[16], left[16];
        for (int i = 0; i < 16; ++i) {
          const int sx = (x0 + i >= width) ? width - 1 : x0 + i;
          top[i] = y_plane[(size_t)(y0 - 1) * y_stride + sx];
        }
        const int top_left = y_plane[(size_t)(y0 - 1) * y_stride + x0 - 1];
        for (int j = 0; j < 16; ++j) {
          const int sy = (y0 + j >= height) ? height - 1 : y0 + j;
          left[j] = y_plane[(size_t)sy * y_stride + x0 - 1];
        }
        for (int j = 0; j < 16; ++j)
          for (int i = 0; i < 16; ++i) pred[j * BPS + i] = clip8(top[i] + left[j] - top_left);
      }
      int distribution[32] = {0};
      int16_t c[16];
      for (int by = 0; by < 4; ++by)
        for (int bx = 0; bx < 4; ++bx) {
          const int off = by * 4 * BPS + bx * 4;
          ftransform(src + off, pred + off, c);
          histo_add(c, distribution);
        }
      const int alpha = alpha_from_histogram(distribution);
      if (alpha < best_alpha) best_alpha = alpha;
    }
    return best_alpha > 255 ? 255 : best_alpha;
  }
  int compute_mb_uv_alpha(int mx, int my) {  // encode_analysis.go:613-728
    uint8_t src_u_[8 * BPS], src_v_[8 * BPS], pred_u[8 * BPS], pred_v[8 * BPS];
    const int ux0 = mx * 8, uy0 = my * 8;
    for (int j = 0; j < 8; ++j)
      for (int i = 0; i < 8; ++i) {
        src_u_[j * BPS + i] = u_plane[(size_t)(uy0 + j) * uv_stride + ux0 + i];
        src_v_[j * BPS + i] = v_plane[(size_t)(uy0 + j) * uv_stride + ux0 + i];
      }
    int dc_u = 128, dc_v = 128, sum_u = 0, sum_v = 0, count = 0;
    if (my > 0)
      for (int i = 0; i < 8; ++i) {
        sum_u += u_plane[(size_t)(uy0 - 1) * uv_stride + ux0 + i];
        sum_v += v_plane[(size_t)(uy0 - 1) * uv_stride + ux0 + i];
        count++;
      }
    if (mx > 0)
      for (int j = 0; j < 8; ++j) {
        sum_u += u_plane[(size_t)(uy0 + j) * uv_stride + ux0 - 1];
        sum_v += v_plane[(size_t)(uy0 + j) * uv_stride + ux0 - 1];
        count++;
      }
    if (count > 0) {
      dc_u = (sum_u + count / 2) / count;
      dc_v = (sum_v + count / 2) / count;
    }
    for (int j = 0; j < 8; ++j) {
      memset(pred_u + j * BPS, dc_u, 8);
      memset(pred_v + j * BPS, dc_v, 8);
    }
    int distribution[32] = {0};
    int16_t c[16];
    for (int by = 0; by < 2; ++by)
      for (int bx = 0; bx < 2; ++bx) {
        const int off = by * 4 * BPS + bx * 4;
        ftransform(src_u_ + off, pred_u + off, c);
        histo_add(c, distribution);
        ftransform(src_v_ + off, pred_v + off, c);
        histo_add(c, distribution);
      }
    return alpha_from_histogram(distribution);
  }

  void analysis() {  // encode_analysis.go:29
    const int num_segs = clampi(cfg.segments, 1, 4);
    const int total = mb_w * mb_h;
    std::vector<int> al(total, 0);
    int64_t uv_alpha_sum = 0;
    alphas.assign(total, 0);
    for (int my = 0; my < mb_h; ++my)
      for (int mx = 0; mx < mb_w; ++mx) {
        const int idx = my * mb_w + mx;
        const int luma_alpha = compute_mb_alpha(mx, my);
        const int uv_alpha = compute_mb_uv_alpha(mx, my);
        int mixed = 255 - ((3 * luma_alpha + uv_alpha + 2) >> 2);
        mixed = clampi(mixed, 0, 255);
        al[idx] = mixed;
        alphas[idx] = (uint8_t)mixed;
        mb_info[idx].alpha = mixed;
        uv_alpha_sum += uv_alpha;
      }
    global_uv_alpha = (int)uv_alpha_sum / total;
    if (num_segs <= 1) {
      for (auto& m : mb_info) m.segment = 0;
      dqm[0].alpha = 0;
      dqm[0].beta = 0;
    } else {
      assign_segments(al, num_segs);
    }
    set_segment_params(num_segs);
    build_segment_header(num_segments);
  }

  void assign_segments(const std::vector<int>& al, int num_segs) {  // encode_analysis.go:737
    int histo[256] = {0};
    for (int a : al) histo[a]++;
    int min_a = 0;
    while (min_a <= 255 && histo[min_a] == 0) min_a++;
    int max_a = 255;
    while (max_a > min_a && histo[max_a] == 0) max_a--;
    const int range_a = max_a - min_a;
    int centers[4] = {0, 0, 0, 0};
    for (int k = 0; k < num_segs; ++k) centers[k] = min_a + ((2 * k + 1) * range_a) / (2 * num_segs);
    int alpha_map[256] = {0};
    int weighted_avg = 0;
    for (int iter = 0; iter < 6; ++iter) {
      int accum[4] = {0, 0, 0, 0}, dist_accum[4] = {0, 0, 0, 0};
      int n = 0;
      for (int a = min_a; a <= max_a; ++a) {
        if (histo[a] == 0) continue;
        while (n + 1 < num_segs && abs(a - centers[n + 1]) < abs(a - centers[n])) n++;
        alpha_map[a] = n;
        dist_accum[n] += a * histo[a];
        accum[n] += histo[a];
      }
      int displaced = 0, total_weight = 0;
      weighted_avg = 0;
      for (int s = 0; s < num_segs; ++s)
        if (accum[s] > 0) {
          const int new_center = (dist_accum[s] + accum[s] / 2) / accum[s];
          displaced += abs(centers[s] - new_center);
          centers[s] = new_center;
          weighted_avg += new_center * accum[s];
          total_weight += accum[s];
        }
      if (total_weight > 0) weighted_avg = (weighted_avg + total_weight / 2) / total_weight;
      if (displaced < 5) break;
    }
    for (auto& m : mb_info) {
      const int a = m.alpha;
      m.segment = (uint8_t)alpha_map[a];
      m.alpha = centers[alpha_map[a]];
    }
    if (cfg.segments > 1 && (cfg.preprocessing & 1)) smooth_segment_map();
    int min_c = centers[0], max_c = centers[0];
    for (int s = 1; s < num_segs; ++s) {
      if (centers[s] < min_c) min_c = centers[s];
      if (centers[s] > max_c) max_c = centers[s];
    }
    int range_c = max_c - min_c;
    if (range_c == 0) range_c = 1;
    for (int s = 0; s < num_segs; ++s) {
      dqm[s].alpha = clampi(255 * (centers[s] - weighted_avg) / range_c, -127, 127);
      dqm[s].beta = clampi(255 * (centers[s] - min_c) / range_c, 0, 255);
    }
  }
  void smooth_segment_map() {  // encode_analysis.go:76
    const int w = mb_w, h = mb_h;
    if (w < 3 || h < 3) return;
    std::vector<uint8_t> tmp(w * h);
    for (int i = 0; i < w * h; ++i) tmp[i] = mb_info[i].segment;
    for (int y = 1; y < h - 1; ++y)
      for (int x = 1; x < w - 1; ++x) {
        int cnt[4] = {0, 0, 0, 0};
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) cnt[mb_info[(y + dy) * w + (x + dx)].segment]++;
        uint8_t best = tmp[y * w + x];
        for (int s = 0; s < 4; ++s) if (cnt[s] >= 5) best = (uint8_t)s;
        tmp[y * w + x] = best;
      }
    for (int y = 1; y < h - 1; ++y)
      for (int x = 1; x < w - 1; ++x) mb_info[y * w + x].segment = tmp[y * w + x];
  }
  void set_segment_params(int num_segs) {  // encode_analysis.go:122
    const int sns = cfg.sns_strength < 0 ? 0 : cfg.sns_strength;
    const double amp = 0.9 * (double)sns / 100.0 / 128.0;
    const double c_base = quality_to_compression(cfg.quality);
    for (int i = 0; i < num_segs; ++i) {
      const double expn = 1.0 - amp * (double)dqm[i].alpha;
      const double c = pow(c_base, expn);
      dqm[i].quant = clampi((int)(127.0 * (1.0 - c)), 0, 127);
    }
    base_quant = dqm[0].quant;
    for (int i = num_segs; i < 4; ++i) dqm[i].quant = base_quant;
    int dq = (global_uv_alpha - 64) * (6 - (-4)) / (100 - 30);
    dq = dq * sns / 100;
    dq_uv_ac = clampi(dq, -4, 6);
    dq_uv_dc = clampi(-4 * sns / 100, -15, 15);
    dq_y1_dc = dq_y2_dc = dq_y2_ac = 0;
    setup_filter_strength();
    if (num_segs > 1) num_segs = simplify_segments(num_segs);
    num_segments = num_segs;
    for (int i = 0; i < 4; ++i) setup_segment(i, dqm[i].quant);
  }
  void setup_filter_strength() {  // encode.go:1276
    f_simple = (cfg.filter_type == 0);
    f_sharpness = clampi(cfg.filter_sharpness, 0, 7);
    if (cfg.filter_strength <= 0) { f_level = 0; return; }
    const int level0 = 5 * cfg.filter_strength;
    const int num_segs = clampi(cfg.segments, 1, 4);
    for (int i = 0; i < num_segs; ++i) {
      SegmentInfo* m = &dqm[i];
      const int qstep = kAcTable[clampi(m->quant, 0, 127)] >> 2;
      const int base_strength = kLevelsFromDelta[f_sharpness * 64 + clampi(qstep, 0, 63)];
      int f = base_strength * level0 / (256 + m->beta);
      if (f < 2) f = 0;
      if (f > 63) f = 63;
      m->fstrength = f;
    }
    f_level = dqm[0].fstrength;
  }
  int simplify_segments(int num_segs) {  // encode_analysis.go:197
    int seg_map[4] = {0, 1, 2, 3};
    int num_final = 1;
    for (int s1 = 1; s1 < num_segs; ++s1) {
      bool found = false;
      for (int s2 = 0; s2 < num_final; ++s2)
        if (dqm[s1].quant == dqm[s2].quant && dqm[s1].fstrength == dqm[s2].fstrength) {
          seg_map[s1] = s2;
          found = true;
          break;
        }
      if (!found) {
        seg_map[s1] = num_final;
        if (num_final != s1) dqm[num_final] = dqm[s1];
        num_final++;
      }
    }
    if (num_final < num_segs) {
      for (auto& m : mb_info) m.segment = (uint8_t)seg_map[m.segment];
      for (int i = num_final; i < num_segs; ++i) dqm[i] = dqm[num_final - 1];
    }
    return num_final;
  }
  void build_segment_header(int num_segs) {  // encode_analysis.go:852
    seg_use = num_segs > 1;
    seg_update_map = seg_use;
    if (seg_use)
      for (int i = 0; i < num_segs; ++i) {
        seg_quantizer[i] = (int8_t)clampi(dqm[i].quant, -127, 127);
        const int qstep0 = kAcTable[clampi(dqm[0].quant, 0, 127)] >> 2;
        const int qstep_i = kAcTable[clampi(dqm[i].quant, 0, 127)] >> 2;
        seg_fstrength[i] = (int8_t)clampi((qstep_i - qstep0) * cfg.filter_strength / 100, -63, 63);
      }
  }
  void set_segment_probas() {  // encode_analysis.go:874
    int counts[4] = {0, 0, 0, 0};
    for (auto& m : mb_info) counts[m.segment]++;
    auto get_proba = [](int a, int b) -> uint8_t {
      const int total = a + b;
      return total == 0 ? 255 : (uint8_t)((255 * a + total / 2) / total);
    };
    proba.segments[0] = get_proba(counts[0] + counts[1], counts[2] + counts[3]);
    proba.segments[1] = get_proba(counts[0], counts[1]);
    proba.segments[2] = get_proba(counts[2], counts[3]);
    if (proba.segments[0] == 255 && proba.segments[1] == 255 && proba.segments[2] == 255) {
      seg_update_map = false;
      for (auto& m : mb_info) m.segment = 0;
    }
  }

  // ---- per-MB path (encode_parallel.go:250-336)
  static void import_block(const uint8_t* src, int src_stride, uint8_t* dst, int w, int h, int size) {
    // encode_iterator.go:145 (src already points at the MB's top-left sample)
    for (int j = 0; j < h; ++j) {
      memcpy(dst + j * BPS, src + (size_t)j * src_stride, w);
      for (int i = w; i < size; ++i) dst[j * BPS + i] = dst[j * BPS + w - 1];
    }
    for (int j = h; j < size; ++j) memcpy(dst + j * BPS, dst + (h - 1) * BPS, size);
  }
  static int check_mode(int mx, int my, int mode) {
    if (mode == B_DC_PRED) {
      if (mx == 0) return my == 0 ? B_DC_PRED_NOTOPLEFT : B_DC_PRED_NOLEFT;
      if (my == 0) return B_DC_PRED_NOTOP;
    }
    return mode;
  }
  static bool is_flat_source16(const uint8_t* src) {  // encode_analysis.go:358
    const uint8_t v = src[0];
    for (int j = 0; j < 16; ++j)
      for (int i = 0; i < 16; ++i) if (src[j * BPS + i] != v) return false;
    return true;
  }
  static bool is_flat(const int16_t* levels, int num_blocks, int thresh) {  // encode_analysis.go:374
    int score = 0;
    for (int b = 0; b < num_blocks; ++b)
      for (int i = 1; i < 16; ++i)
        if (levels[b * 16 + i] != 0) { score++; if (score > thresh) return false; }
    return true;
  }
  static bool needs_top4(int mode) {
    return mode == B_VE_PRED || mode == B_VR_PRED || mode == B_LD_PRED || mode == B_VL_PRED || mode == B_HD_PRED ||
           mode == B_RD_PRED || mode == B_TM_PRED;
  }
  static bool needs_left4(int mode) {
    return mode == B_HE_PRED || mode == B_HU_PRED || mode == B_HD_PRED || mode == B_RD_PRED || mode == B_TM_PRED;
  }

  // pickBestI16ModeRDParallel (encode_parallel.go:624)
  void pick_best_i16(int mx, int my, const SegmentInfo* seg, uint32_t top_nz, uint32_t left_nz, int top_nz_dc,
                     int left_nz_dc, int* best_mode, int* best_rate, int* best_disto) {
    uint64_t best_score = ~(uint64_t)0;
    *best_mode = DC_PRED; *best_rate = 0; *best_disto = 0;
    const uint8_t* src = yuv_in;
    uint8_t* pred = yuv_out2;
    const bool src_flat = is_flat_source16(src + Y_OFF);
    memcpy(pred, yuv_out, U_OFF);
    const uint32_t init_tnz = top_nz & 0x0f, init_lnz = left_nz & 0x0f;
    int dc_ctx = top_nz_dc + left_nz_dc;
    if (dc_ctx > 2) dc_ctx = 2;
    for (int mode = 0; mode < 4; ++mode) {
      if (mode == V_PRED && my == 0) continue;
      if (mode == H_PRED && mx == 0) continue;
      if (mode == TM_PRED && (mx == 0 || my == 0)) continue;
      pred_luma16(check_mode(mx, my, mode), pred, Y_OFF);
      int16_t dc_coeffs[16] = {0}, all_q[16][16], c[16], q[16];
      int total_rate = kModeFixedCost16[mode];
      uint32_t tnz = init_tnz, lnz = init_lnz;
      for (int by = 0; by < 4; ++by) {
        uint32_t l = lnz & 1;
        for (int bx = 0; bx < 4; ++bx) {
          const int b = by * 4 + bx, off = Y_OFF + by * 4 * BPS + bx * 4;
          int ctx = (int)l + (int)(tnz & 1);
          if (ctx > 2) ctx = 2;
          ftransform(src + off, pred + off, c);
          dc_coeffs[b] = c[0];
          c[0] = 0;
          const int nz = quantize_coeffs(c, q, &seg->y1, 1);
          memcpy(all_q[b], q, sizeof(q));
          total_rate += token_cost(q, nz, 0, &proba, ctx, 1);
          l = nz > 0;
          tnz = (tnz >> 1) | (l << 7);
        }
        tnz >>= 4;
        lnz = (lnz >> 1) | (l << 7);
      }
      ftransform_wht(dc_coeffs, c);
      const int nz_dc = quantize_coeffs(c, q, &seg->y2, 0);
      total_rate += token_cost(q, nz_dc, 1, &proba, dc_ctx, 0);
      int16_t wht_dq[16], wht_buf[256], dq[16];
      dequant_coeffs(q, wht_dq, &seg->y2);
      transform_wht(wht_dq, wht_buf);
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        dequant_coeffs(all_q[b], dq, &seg->y1);
        dq[0] = wht_buf[b * 16];
        itransform_one(pred + off, dq, pred + off);
      }
      int disto = sse16x16(src + Y_OFF, pred + Y_OFF);
      if (seg->tlambda_sd > 0) disto += (seg->tlambda_sd * tdisto16x16(src + Y_OFF, pred + Y_OFF) + 128) >> 8;
      if (src_flat && is_flat(&all_q[0][0], 16, 0)) disto *= 2;
      const uint64_t score = rd_score(disto, total_rate, seg->lambda_i16);
      if (score < best_score) {
        best_score = score;
        *best_mode = mode; *best_rate = total_rate; *best_disto = disto;
      }
    }
  }

  // pickBestI4ModeRD(Trellis)Parallel (encode_parallel.go:842,942); pred_buf = yuv_out2, in place.
  void pick_best_i4(int off, const SegmentInfo* seg, int top_mode, int left_mode, bool has_top, bool has_left,
                    int nz_ctx, int max_modes, bool trellis, int* best_mode, int* best_rate, int* best_disto) {
    uint64_t best_score = ~(uint64_t)0;
    *best_mode = B_DC_PRED; *best_rate = 0; *best_disto = 0;
    const uint8_t* src = yuv_in + off;
    uint8_t* pred_buf = yuv_out2;
    struct Cand { int mode, sse; } cand[10];
    int n_cand = 0;
    for (int mode = 0; mode < 10; ++mode) {
      if (!has_top && needs_top4(mode)) continue;
      if (!has_left && needs_left4(mode)) continue;
      pred_luma4(mode, pred_buf, off);
      cand[n_cand].mode = mode;
      cand[n_cand].sse = sse4x4(src, pred_buf + off);
      n_cand++;
    }
    int K = max_modes;
    if (n_cand <= K) K = n_cand;
    for (int i = 0; i < K; ++i) {
      int min_idx = i;
      for (int j = i + 1; j < n_cand; ++j) if (cand[j].sse < cand[min_idx].sse) min_idx = j;
      if (min_idx != i) { Cand t = cand[i]; cand[i] = cand[min_idx]; cand[min_idx] = t; }
    }
    for (int i = 0; i < K; ++i) {
      const int mode = cand[i].mode;
      int16_t c[16], q[16], dq[16];
      uint8_t recon[4 * BPS];
      pred_luma4(mode, pred_buf, off);
      ftransform(src, pred_buf + off, c);
      const int nz = trellis ? trellis_quantize_block(c, q, &seg->y1, 0, 3, nz_ctx, &proba, seg->tlambda_i4)
                             : quantize_coeffs(c, q, &seg->y1, 0);
      dequant_coeffs(q, dq, &seg->y1);
      itransform_one(pred_buf + off, dq, recon);
      int disto = sse4x4(src, recon);
      if (seg->tlambda_sd > 0) disto += (seg->tlambda_sd * tdisto4x4(src, recon) + 128) >> 8;
      if (256 * (uint64_t)disto >= best_score) continue;
      int rate = 0;
      if (mode > 0 && is_flat(q, 1, 3)) rate = 140;
      rate += token_cost(q, nz, 3, &proba, nz_ctx, 0);
      rate += fixed_costs_i4[top_mode][left_mode][mode];
      const uint64_t score = rd_score(disto, rate, seg->lambda_i4);
      if (score < best_score) {
        best_score = score;
        *best_mode = mode; *best_rate = rate; *best_disto = disto;
        memcpy(tmp_best_dq, dq, sizeof(dq));
        memcpy(tmp_best_q, q, sizeof(q));
        tmp_best_nz = nz;
      }
    }
  }

  // tryI4ModesRDParallel (encode_parallel.go:738)
  uint64_t try_i4_modes(int mx, int my, MBInfo* info, const SegmentInfo* seg, uint8_t* modes, const uint8_t* top_modes,
                        const uint8_t* left_modes, uint64_t i16_score, uint32_t top_nz, uint32_t left_nz) {
    int total_rate = 0, total_disto = 0, total_header_bits = 0;
    uint8_t top_m[4];
    for (int i = 0; i < 4; ++i) top_m[i] = (my > 0) ? top_modes[mx * 4 + i] : B_DC_PRED;
    memcpy(yuv_out2, yuv_out, YUV_SIZE);
    uint32_t tnz = top_nz & 0x0f, lnz = left_nz & 0x0f, l = 0;
    bool early_exit = false;
    const int max_modes = cfg.quality < 50 ? 2 : 3;  // getMaxI4RDModes (encode_parallel.go:931)
    for (int by = 0; by < 4 && !early_exit; ++by) {
      l = lnz & 1;
      for (int bx = 0; bx < 4; ++bx) {
        const int b = by * 4 + bx;
        const int top_mode = (by == 0) ? top_m[bx] : modes[b - 4];
        const int left_mode = (bx == 0) ? left_modes[by] : modes[b - 1];
        const int off = Y_OFF + by * 4 * BPS + bx * 4;
        const bool has_top = (my > 0 || by > 0), has_left = (mx > 0 || bx > 0);
        int nz_ctx = (int)l + (int)(tnz & 1);
        if (nz_ctx > 2) nz_ctx = 2;
        int best_mode, rate, disto;
        pick_best_i4(off, seg, top_mode, left_mode, has_top, has_left, nz_ctx, max_modes, cfg.method >= 4,
                     &best_mode, &rate, &disto);
        modes[b] = (uint8_t)best_mode;
        total_rate += rate;
        total_disto += disto;
        total_header_bits += fixed_costs_i4[top_mode][left_mode][best_mode];
        memcpy(info->coeffs + b * 16, tmp_best_q, 32);
        const int nz = tmp_best_nz;
        info->nz_y[b] = (uint8_t)nz;
        if (rd_score(total_disto, total_rate + 211, seg->lambda_mode) >= i16_score) { early_exit = true; break; }
        if (total_header_bits > 15000) { early_exit = true; break; }
        pred_luma4(best_mode, yuv_out2, off);
        itransform_one(yuv_out2 + off, tmp_best_dq, yuv_out2 + off);
        l = nz > 0;
        tnz = (tnz >> 1) | (l << 7);
      }
      tnz >>= 4;
      lnz = (lnz >> 1) | (l << 7);
    }
    if (early_exit) return ~(uint64_t)0;
    return rd_score(total_disto, total_rate + 211, seg->lambda_mode);
  }

  // pickBestUVModeRDParallel (encode_parallel.go:1030)
  int pick_best_uv(int mx, int my, const SegmentInfo* seg, uint32_t top_nz, uint32_t left_nz) {
    uint64_t best_score = ~(uint64_t)0;
    int best_mode = DC_PRED;
    const uint8_t* src = yuv_in;
    uint8_t* pred = yuv_out2;
    memcpy(pred + U_OFF, yuv_out + U_OFF, YUV_SIZE - U_OFF);
    for (int mode = 0; mode < 4; ++mode) {
      if (mode == V_PRED && my == 0) continue;
      if (mode == H_PRED && mx == 0) continue;
      if (mode == TM_PRED && (mx == 0 || my == 0)) continue;
      const int am = check_mode(mx, my, mode);
      pred_chroma8(am, pred, U_OFF);
      pred_chroma8(am, pred, V_OFF);
      int total_rate = kModeFixedCostUV[mode];
      int16_t uv_levels[128];
      int uv_idx = 0;
      for (int ch = 0; ch < 4; ch += 2) {
        uint32_t tnz = (top_nz >> (4 + ch)) & 0x0f, lnz = (left_nz >> (4 + ch)) & 0x0f;
        const int plane_off = ch ? V_OFF : U_OFF;
        for (int by = 0; by < 2; ++by) {
          uint32_t l = lnz & 1;
          for (int bx = 0; bx < 2; ++bx) {
            const int off = plane_off + by * 4 * BPS + bx * 4;
            int ctx = (int)l + (int)(tnz & 1);
            if (ctx > 2) ctx = 2;
            int16_t c[16], q[16], dq[16];
            ftransform(src + off, pred + off, c);
            const int nz = quantize_coeffs(c, q, &seg->uv, 0);
            total_rate += token_cost(q, nz, 2, &proba, ctx, 0);
            memcpy(uv_levels + uv_idx * 16, q, 32);
            uv_idx++;
            dequant_coeffs(q, dq, &seg->uv);
            itransform_one(pred + off, dq, pred + off);
            l = nz > 0;
            tnz = (tnz >> 1) | (l << 3);
          }
          tnz >>= 2;
          lnz = (lnz >> 1) | (l << 5);
        }
      }
      if (mode > 0 && is_flat(uv_levels, 8, 2)) total_rate += 140 * 8;
      int disto = 0;
      for (int by = 0; by < 2; ++by)
        for (int bx = 0; bx < 2; ++bx) {
          const int off = by * 4 * BPS + bx * 4;
          disto += sse4x4(src + U_OFF + off, pred + U_OFF + off);
          disto += sse4x4(src + V_OFF + off, pred + V_OFF + off);
        }
      const uint64_t score = rd_score(disto, total_rate, seg->lambda_uv);
      if (score < best_score) { best_score = score; best_mode = mode; }
    }
    return best_mode;
  }

  // encodeFrameParallel Phase A, one MB (encode_parallel.go:293-335)
  struct RowCtx {
    uint8_t left_y[16], left_u[8], left_v[8], left_modes[4];
    uint8_t top_left_y, top_left_u, top_left_v;
    uint32_t left_nz;
    uint8_t left_nz_dc;
  };
  std::vector<uint8_t> top_y, top_u, top_v, top_modes, top_nz_dc;
  std::vector<uint32_t> top_nz;

  void encode_mb(int mx, int my, RowCtx& rc) {
    MBInfo* info = &mb_info[(size_t)my * mb_w + mx];
    const SegmentInfo* seg = &dqm[info->segment];
    // 1. import (encode_parallel.go:431)
    {
      const int x = mx * 16, y = my * 16;
      const int ww = width - x > 16 ? 16 : width - x, hh = height - y > 16 ? 16 : height - y;
      import_block(&y_plane[(size_t)y * y_stride + x], y_stride, yuv_in + Y_OFF, ww, hh, 16);
      const int uvw = (ww + 1) >> 1, uvh = (hh + 1) >> 1;
      import_block(&u_plane[(size_t)my * 8 * uv_stride + mx * 8], uv_stride, yuv_in + U_OFF, uvw, uvh, 8);
      import_block(&v_plane[(size_t)my * 8 * uv_stride + mx * 8], uv_stride, yuv_in + V_OFF, uvw, uvh, 8);
    }
    // 2. prediction context (encode_parallel.go:455)
    {
      uint8_t* o = yuv_out;
      for (int i = 0; i < 16; ++i) o[Y_OFF - BPS + i] = my > 0 ? top_y[mx * 16 + i] : 127;
      for (int i = 0; i < 4; ++i)
        o[Y_OFF - BPS + 16 + i] = my > 0 ? (mx < mb_w - 1 ? top_y[(mx + 1) * 16 + i] : top_y[mx * 16 + 15]) : 127;
      for (int r = 1; r <= 3; ++r) memcpy(o + Y_OFF - BPS + 16 + r * 4 * BPS, o + Y_OFF - BPS + 16, 4);
      o[Y_OFF - BPS - 1] = (mx > 0 && my > 0) ? rc.top_left_y : (my > 0 ? 129 : 127);
      for (int j = 0; j < 16; ++j) o[Y_OFF - 1 + j * BPS] = mx > 0 ? rc.left_y[j] : 129;
      for (int i = 0; i < 8; ++i) {
        o[U_OFF - BPS + i] = my > 0 ? top_u[mx * 8 + i] : 127;
        o[V_OFF - BPS + i] = my > 0 ? top_v[mx * 8 + i] : 127;
      }
      o[U_OFF - BPS - 1] = (mx > 0 && my > 0) ? rc.top_left_u : (my > 0 ? 129 : 127);
      o[V_OFF - BPS - 1] = (mx > 0 && my > 0) ? rc.top_left_v : (my > 0 ? 129 : 127);
      for (int j = 0; j < 8; ++j) {
        o[U_OFF - 1 + j * BPS] = mx > 0 ? rc.left_u[j] : 129;
        o[V_OFF - 1 + j * BPS] = mx > 0 ? rc.left_v[j] : 129;
      }
    }
    const uint32_t tnz_val = top_nz[mx], lnz_val = rc.left_nz;
    const int tnz_dc = top_nz_dc[mx], lnz_dc = rc.left_nz_dc;
    // 3. mode decision (encode_parallel.go:563-598)
    bool pred_cached, i4_cached = false;
    {
      int best16, rate16, disto16;
      pick_best_i16(mx, my, seg, tnz_val, lnz_val, tnz_dc, lnz_dc, &best16, &rate16, &disto16);
      const uint64_t score16 = rd_score(disto16, rate16, seg->lambda_mode);
      uint8_t modes4[16] = {0};
      const uint64_t score4 = try_i4_modes(mx, my, info, seg, modes4, top_modes.data(), rc.left_modes, score16,
                                           tnz_val, lnz_val);
      if (score4 < score16) {
        info->mb_type = 1;
        memcpy(info->modes, modes4, 16);
        pred_cached = false;
        if (cfg.method >= 4) {
          i4_cached = true;
          for (int j = 0; j < 16; ++j) memcpy(yuv_out + Y_OFF + j * BPS, yuv_out2 + Y_OFF + j * BPS, 16);
        }
      } else {
        info->mb_type = 0;
        info->i16_mode = (uint8_t)best16;
        pred_luma16(check_mode(mx, my, best16), yuv_out, Y_OFF);
        pred_cached = true;
      }
      const int best_uv = pick_best_uv(mx, my, seg, tnz_val, lnz_val);
      info->uv_mode = (uint8_t)best_uv;
      pred_chroma8(check_mode(mx, my, best_uv), yuv_out, U_OFF);
      pred_chroma8(check_mode(mx, my, best_uv), yuv_out, V_OFF);
    }
    (void)pred_cached;
    // 4. residuals (encode_parallel.go:1164-1355)
    if (info->mb_type == 0) {
      int16_t dc_coeffs[16];
      uint32_t nz_y = 0, tnz = tnz_val & 0x0f, lnz = lnz_val & 0x0f;
      for (int by = 0; by < 4; ++by) {
        uint32_t l = lnz & 1;
        for (int bx = 0; bx < 4; ++bx) {
          const int b = by * 4 + bx, off = Y_OFF + by * 4 * BPS + bx * 4;
          int16_t* c = info->coeffs + b * 16;
          ftransform(yuv_in + off, yuv_out + off, c);
          dc_coeffs[b] = c[0];
          c[0] = 0;
          int nz;
          if (cfg.method >= 4) {
            int ctx = (int)l + (int)(tnz & 1);
            if (ctx > 2) ctx = 2;
            nz = trellis_quantize_block(c, c, &seg->y1, 1, 0, ctx, &proba, seg->tlambda_i16);
          } else {
            nz = quantize_coeffs(c, c, &seg->y1, 1);
          }
          info->nz_y[b] = (uint8_t)nz;
          if (nz > 0) nz_y |= 1u << b;
          l = nz > 0;
          tnz = (tnz >> 1) | (l << 7);
        }
        tnz >>= 4;
        lnz = (lnz >> 1) | (l << 7);
      }
      int16_t wht[16];
      ftransform_wht(dc_coeffs, wht);
      const int nz_dc = quantize_coeffs(wht, info->coeffs + 384, &seg->y2, 0);
      info->nz_dc = (uint8_t)nz_dc;
      if (nz_dc > 0) nz_y |= 1u << 24;
      info->non_zero_y = nz_y;
    } else if (i4_cached) {
      uint32_t nz_y = 0;
      for (int b = 0; b < 16; ++b) if (info->nz_y[b] > 0) nz_y |= 1u << b;
      info->non_zero_y = nz_y;
    } else {
      uint32_t nz_y = 0;
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        int16_t* c = info->coeffs + b * 16;
        int16_t dq[16];
        pred_luma4(info->modes[b], yuv_out, off);
        ftransform(yuv_in + off, yuv_out + off, c);
        const int nz = quantize_coeffs(c, c, &seg->y1, 0);
        info->nz_y[b] = (uint8_t)nz;
        if (nz > 0) nz_y |= 1u << b;
        dequant_coeffs(c, dq, &seg->y1);
        itransform_one(yuv_out + off, dq, yuv_out + off);
      }
      info->non_zero_y = nz_y;
    }
    {  // UV residuals (encode_parallel.go:1295): no DC error diffusion on this path
      uint32_t nz_uv = 0;
      for (int ch = 0; ch < 2; ++ch)
        for (int b = 0; b < 4; ++b) {
          const int off = (ch ? V_OFF : U_OFF) + (b >> 1) * 4 * BPS + (b & 1) * 4;
          int16_t* c = info->coeffs + (16 + ch * 4 + b) * 16;
          ftransform(yuv_in + off, yuv_out + off, c);
          const int nz = quantize_coeffs(c, c, &seg->uv, 0);
          info->nz_uv[ch * 4 + b] = (uint8_t)nz;
          if (nz > 0) nz_uv |= 1u << (ch * 4 + b);
        }
      info->non_zero_uv = nz_uv;
    }
    // 5. skip
    info->skip = (info->non_zero_y == 0 && info->non_zero_uv == 0);
    // 6. reconstruct (encode_parallel.go:1358)
    if (info->mb_type == 0) {
      int16_t wht_dq[16], wht_buf[256], dq[16];
      dequant_coeffs(info->coeffs + 384, wht_dq, &seg->y2);
      transform_wht(wht_dq, wht_buf);
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        dequant_coeffs(info->coeffs + b * 16, dq, &seg->y1);
        dq[0] = wht_buf[b * 16];
        itransform_one(yuv_out + off, dq, yuv_out + off);
      }
    }
    for (int b = 0; b < 4; ++b) {
      int16_t dq[16];
      const int o = (b >> 1) * 4 * BPS + (b & 1) * 4;
      dequant_coeffs(info->coeffs + (16 + b) * 16, dq, &seg->uv);
      itransform_one(yuv_out + U_OFF + o, dq, yuv_out + U_OFF + o);
      dequant_coeffs(info->coeffs + (20 + b) * 16, dq, &seg->uv);
      itransform_one(yuv_out + V_OFF + o, dq, yuv_out + V_OFF + o);
    }
    // 7. export (encode_parallel.go:1410)
    {
      const int x = mx * 16, y = my * 16;
      const int wy = x + 16 > width ? width - x : 16, hy = y + 16 > height ? height - y : 16;
      for (int j = 0; j < hy; ++j) memcpy(&y_plane[(size_t)(y + j) * y_stride + x], yuv_out + Y_OFF + j * BPS, wy);
      for (int j = 0; j < 8; ++j) {
        memcpy(&u_plane[(size_t)(my * 8 + j) * uv_stride + mx * 8], yuv_out + U_OFF + j * BPS, 8);
        memcpy(&v_plane[(size_t)(my * 8 + j) * uv_stride + mx * 8], yuv_out + V_OFF + j * BPS, 8);
      }
      rc.top_left_y = top_y[mx * 16 + 15];
      rc.top_left_u = top_u[mx * 8 + 7];
      rc.top_left_v = top_v[mx * 8 + 7];
      memcpy(&top_y[mx * 16], yuv_out + Y_OFF + 15 * BPS, 16);
      memcpy(&top_u[mx * 8], yuv_out + U_OFF + 7 * BPS, 8);
      memcpy(&top_v[mx * 8], yuv_out + V_OFF + 7 * BPS, 8);
      for (int j = 0; j < 16; ++j) rc.left_y[j] = yuv_out[Y_OFF + j * BPS + 15];
      for (int j = 0; j < 8; ++j) {
        rc.left_u[j] = yuv_out[U_OFF + j * BPS + 7];
        rc.left_v[j] = yuv_out[V_OFF + j * BPS + 7];
      }
      if (info->mb_type == 1) {
        for (int i = 0; i < 4; ++i) {
          top_modes[mx * 4 + i] = info->modes[12 + i];
          rc.left_modes[i] = info->modes[3 + 4 * i];
        }
      } else {
        for (int i = 0; i < 4; ++i) top_modes[mx * 4 + i] = rc.left_modes[i] = B_DC_PRED;
      }
    }
    // 8. NZ context (encode_parallel.go:341)
    update_nz(info, &top_nz[mx], &rc.left_nz, &top_nz_dc[mx], &rc.left_nz_dc);
  }

  // Shared NZ-context walk: updateNZContextParallel / recordMBTokens / collectMBStats.
  // visit(coeffs, nz, type, first, ctx) is called per block in bitstream order when non-null.
  template <class F>
  static void walk_mb(const MBInfo* info, uint32_t* top_nz, uint32_t* left_nz, uint8_t* top_nz_dc,
                      uint8_t* left_nz_dc, F visit) {
    const uint32_t top = *top_nz, left = *left_nz;
    uint32_t out_t, out_l;
    int first = 0, type = 3;
    if (info->mb_type == 0) {
      int dc_ctx = *top_nz_dc + *left_nz_dc;
      if (dc_ctx > 2) dc_ctx = 2;
      visit(info->coeffs + 384, (int)info->nz_dc, 1, 0, dc_ctx);
      *top_nz_dc = *left_nz_dc = (info->nz_dc > 0);
      first = 1;
      type = 0;
    }
    {
      uint32_t tnz = top & 0x0f, lnz = left & 0x0f;
      for (int y = 0; y < 4; ++y) {
        uint32_t l = lnz & 1;
        for (int x = 0; x < 4; ++x) {
          const int b = y * 4 + x;
          int ctx = (int)l + (int)(tnz & 1);
          if (ctx > 2) ctx = 2;
          const int nz = info->nz_y[b];
          visit(info->coeffs + b * 16, nz, type, first, ctx);
          l = nz > first;
          tnz = (tnz >> 1) | (l << 7);
        }
        tnz >>= 4;
        lnz = (lnz >> 1) | (l << 7);
      }
      out_t = tnz;
      out_l = lnz >> 4;
    }
    for (int ch = 0; ch < 4; ch += 2) {
      uint32_t tnz = (top >> (4 + ch)) & 0x0f, lnz = (left >> (4 + ch)) & 0x0f;
      for (int y = 0; y < 2; ++y) {
        uint32_t l = lnz & 1;
        for (int x = 0; x < 2; ++x) {
          const int uv_idx = (ch / 2) * 4 + y * 2 + x;
          int ctx = (int)l + (int)(tnz & 1);
          if (ctx > 2) ctx = 2;
          const int nz = info->nz_uv[uv_idx];
          visit(info->coeffs + (16 + uv_idx) * 16, nz, 2, 0, ctx);
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
  static void update_nz(const MBInfo* info, uint32_t* top_nz, uint32_t* left_nz, uint8_t* top_nz_dc,
                        uint8_t* left_nz_dc) {
    walk_mb(info, top_nz, left_nz, top_nz_dc, left_nz_dc, [](const int16_t*, int, int, int, int) {});
  }

  // ---- tokens (encode_token.go:115-300)
  void rec(int bit, int prob) { tokens.push_back((uint16_t)((bit & 1) | (prob << 8))); }
  void record_level(int level, const uint8_t* p) {
    if (level == 1) { rec(0, p[2]); return; }
    rec(1, p[2]);
    if (level <= 4) {
      rec(0, p[3]);
      if (level == 2) rec(0, p[4]);
      else { rec(1, p[4]); rec(level == 3 ? 0 : 1, p[5]); }
    } else if (level <= 10) {
      rec(1, p[3]);
      rec(0, p[6]);
      if (level <= 6) { rec(0, p[7]); rec(level - 5, 159); }
      else { rec(1, p[7]); const int v = level - 7; rec(v >> 1, 165); rec(v & 1, 145); }
    } else {
      rec(1, p[3]);
      rec(1, p[6]);
      const int cat = level <= 18 ? 0 : level <= 34 ? 1 : level <= 66 ? 2 : 3;
      const int bit1 = cat >> 1, bit0 = cat & 1;
      rec(bit1, p[8]);
      rec(bit0, p[9 + bit1]);
      const int v = level - (3 + (8 << cat));
      const uint8_t* tab = kCat3456[cat];
      int nbits = 0;
      while (tab[nbits]) nbits++;
      for (int i = 0; i < nbits; ++i) rec((v >> (nbits - 1 - i)) & 1, tab[i]);
    }
  }
  void record_coeffs(const int16_t* coeffs, int n_coeffs, int type, int first, int ctx) {
    int n = first;
    if (n_coeffs <= first) { rec(0, proba.bands[type][kBands[n]][ctx][0]); return; }
    while (n < 16) {
      const uint8_t* p = proba.bands[type][kBands[n]][ctx];
      if (n >= n_coeffs) { rec(0, p[0]); return; }
      rec(1, p[0]);
      for (;;) {
        int v = coeffs[kZigzag[n]];
        int sign = 0;
        if (v < 0) { v = -v; sign = 1; }
        if (v == 0) {
          rec(0, p[1]);
          if (++n >= 16) return;
          p = proba.bands[type][kBands[n]][0];
          continue;
        }
        rec(1, p[1]);
        record_level(v, p);
        rec(sign, 128);
        ctx = (v == 1) ? 1 : 2;
        n++;
        break;
      }
    }
  }
  // ---- stats (encode_proba.go:10-113)
  static void collect_coeff_stats(const int16_t* coeffs, int n_coeffs, int type, int first, int ctx, ProbaStats st) {
    int n = first;
    if (n_coeffs <= first) { st[type][kBands[n]][ctx][0][0]++; return; }
    while (n < 16) {
      int b = kBands[n];
      if (n >= n_coeffs) { st[type][b][ctx][0][0]++; return; }
      st[type][b][ctx][0][1]++;
      for (;;) {
        int v = abs((int)coeffs[kZigzag[n]]);
        b = kBands[n];
        if (v == 0) {
          st[type][b][ctx][1][0]++;
          if (++n >= 16) return;
          ctx = 0;
          continue;
        }
        st[type][b][ctx][1][1]++;
        if (v == 1) {
          st[type][b][ctx][2][0]++;
        } else {
          st[type][b][ctx][2][1]++;
          if (v <= 4) {
            st[type][b][ctx][3][0]++;
            if (v == 2) st[type][b][ctx][4][0]++;
            else { st[type][b][ctx][4][1]++; st[type][b][ctx][5][v == 3 ? 0 : 1]++; }
          } else if (v <= 10) {
            st[type][b][ctx][3][1]++;
            st[type][b][ctx][6][0]++;
            st[type][b][ctx][7][v <= 6 ? 0 : 1]++;
          } else {
            st[type][b][ctx][3][1]++;
            st[type][b][ctx][6][1]++;
            const int cat = v <= 18 ? 0 : v <= 34 ? 1 : v <= 66 ? 2 : 3;
            st[type][b][ctx][8][cat >> 1]++;
            st[type][b][ctx][9 + (cat >> 1)][cat & 1]++;
          }
        }
        ctx = (v == 1) ? 1 : 2;
        n++;
        break;
      }
    }
  }
  static int64_t branch_cost(int cnt0, int cnt1, int p) {  // Go int is 64-bit: counts x costs exceed 2^31 on large frames
    p = clampi(p, 1, 255);
    return (int64_t)cnt1 * bit_cost(1, (uint8_t)p) + (int64_t)cnt0 * bit_cost(0, (uint8_t)p);
  }
  int optimize_proba(ProbaStats st) {  // encode_proba.go:117
    int num_updates = 0;
    for (int t = 0; t < 4; ++t)
      for (int b = 0; b < 8; ++b)
        for (int c = 0; c < 3; ++c)
          for (int p = 0; p < 11; ++p) {
            const int cnt0 = st[t][b][c][p][0], cnt1 = st[t][b][c][p][1], total = cnt0 + cnt1;
            if (total == 0) continue;
            const int new_p = cnt1 > 0 ? 255 - (int)((int64_t)cnt1 * 255 / total) : 255;
            const int idx = ((t * 8 + b) * 3 + c) * 11 + p;
            const int old_p = kCoeffsProba0[idx];
            const uint8_t up = kCoeffsUpdateProba[idx];
            const int64_t old_cost = branch_cost(cnt0, cnt1, old_p) + bit_cost(0, up);
            const int64_t new_cost = branch_cost(cnt0, cnt1, new_p) + bit_cost(1, up) + 8 * 256;
            if (old_cost > new_cost) {
              proba.bands[t][b][c][p] = (uint8_t)new_p;
              num_updates++;
            }
          }
    return num_updates;
  }

  // recordAllTokens (encode_parallel.go:1503) + rerecordAllTokens (encode_proba.go:317)
  void record_all_tokens(ProbaStats st) {
    std::vector<uint32_t> tnz(mb_w, 0), s_tnz(mb_w, 0);
    std::vector<uint8_t> tnz_dc(mb_w, 0), s_tnz_dc(mb_w, 0);
    tokens.clear();
    num_skip = 0;
    if (st) memset(st, 0, sizeof(ProbaStats));
    for (int my = 0; my < mb_h; ++my) {
      uint32_t lnz = 0, s_lnz = 0;
      uint8_t lnz_dc = 0, s_lnz_dc = 0;
      for (int mx = 0; mx < mb_w; ++mx) {
        const size_t idx = (size_t)my * mb_w + mx;
        const MBInfo* info = &mb_info[idx];
        if (info->skip) {
          num_skip++;
          tnz[mx] = 0; lnz = 0;
          s_tnz[mx] = 0; s_lnz = 0;
          if (info->mb_type == 0) { tnz_dc[mx] = 0; lnz_dc = 0; s_tnz_dc[mx] = 0; s_lnz_dc = 0; }
          continue;
        }
        mb_start[idx] = tokens.size();
        walk_mb(info, &tnz[mx], &lnz, &tnz_dc[mx], &lnz_dc,
                [&](const int16_t* c, int nz, int type, int first, int ctx) { record_coeffs(c, nz, type, first, ctx); });
        if (st)
          walk_mb(info, &s_tnz[mx], &s_lnz, &s_tnz_dc[mx], &s_lnz_dc,
                  [&](const int16_t* c, int nz, int type, int first, int ctx) {
                    collect_coeff_stats(c, nz, type, first, ctx, st);
                  });
      }
    }
  }

  // ---- emit (encode_syntax.go)
  void write_i4_mode(BoolWriter& bw, int mode, const uint8_t* prob) {
    int bit = subtree_contains(kYModesIntra4[0], mode) ? 0 : 1;
    bw.put_bit(bit, prob[0]);
    int i = kYModesIntra4[bit];
    while (i > 0) {
      bit = subtree_contains(kYModesIntra4[2 * i], mode) ? 0 : 1;
      bw.put_bit(bit, prob[i]);
      i = kYModesIntra4[2 * i + bit];
    }
  }
  std::vector<uint8_t> emit_partition0() {
    BoolWriter bw;
    bw.put_bit_uniform(0);
    bw.put_bit_uniform(0);
    // segment header (encode_syntax.go:175)
    bw.put_bit_uniform(seg_use);
    if (seg_use) {
      bw.put_bit_uniform(seg_update_map);
      bw.put_bit_uniform(1);
      bw.put_bit_uniform(1);  // absolute delta
      for (int i = 0; i < 4; ++i) {
        const int q = seg_quantizer[i];
        if (q != 0) { bw.put_bit_uniform(1); bw.put_bits((uint32_t)abs(q), 7); bw.put_bit_uniform(q < 0); }
        else bw.put_bit_uniform(0);
      }
      for (int i = 0; i < 4; ++i) {
        const int f = seg_fstrength[i];
        if (f != 0) { bw.put_bit_uniform(1); bw.put_bits((uint32_t)abs(f), 6); bw.put_bit_uniform(f < 0); }
        else bw.put_bit_uniform(0);
      }
      if (seg_update_map)
        for (int i = 0; i < 3; ++i) {
          if (proba.segments[i] != 255) { bw.put_bit_uniform(1); bw.put_bits(proba.segments[i], 8); }
          else bw.put_bit_uniform(0);
        }
    }
    // filter header (encode_syntax.go:240); use_lf_delta is always false (encode.go:1285)
    bw.put_bit_uniform(f_simple);
    bw.put_bits((uint32_t)f_level, 6);
    bw.put_bits((uint32_t)f_sharpness, 3);
    bw.put_bit_uniform(0);
    const int log2_parts = num_parts == 2 ? 1 : num_parts == 4 ? 2 : num_parts == 8 ? 3 : 0;
    bw.put_bits((uint32_t)log2_parts, 2);
    // quant (encode_syntax.go:307)
    bw.put_bits((uint32_t)dqm[0].quant, 7);
    bw.put_signed_bits(dq_y1_dc, 4);
    bw.put_signed_bits(dq_y2_dc, 4);
    bw.put_signed_bits(dq_y2_ac, 4);
    bw.put_signed_bits(dq_uv_dc, 4);
    bw.put_signed_bits(dq_uv_ac, 4);
    bw.put_bit_uniform(0);  // refresh
    for (int t = 0; t < 4; ++t)
      for (int b = 0; b < 8; ++b)
        for (int c = 0; c < 3; ++c)
          for (int p = 0; p < 11; ++p) {
            const int idx = ((t * 8 + b) * 3 + c) * 11 + p;
            const uint8_t prob = proba.bands[t][b][c][p];
            if (prob != kCoeffsProba0[idx]) { bw.put_bit(1, kCoeffsUpdateProba[idx]); bw.put_bits(prob, 8); }
            else bw.put_bit(0, kCoeffsUpdateProba[idx]);
          }
    if (num_skip > 0) { bw.put_bit_uniform(1); bw.put_bits(skip_proba, 8); } else bw.put_bit_uniform(0);
    // MB modes (encode_syntax.go:349)
    std::vector<uint8_t> tm(mb_w * 4, 0);
    for (int my = 0; my < mb_h; ++my) {
      uint8_t lm[4] = {0, 0, 0, 0};
      for (int mx = 0; mx < mb_w; ++mx) {
        const MBInfo* info = &mb_info[(size_t)my * mb_w + mx];
        uint8_t* top = &tm[4 * mx];
        if (seg_use && seg_update_map) {
          const int id = info->segment;
          bw.put_bit((id >> 1) & 1, proba.segments[0]);
          bw.put_bit(id & 1, id >= 2 ? proba.segments[2] : proba.segments[1]);
        }
        if (num_skip > 0) bw.put_bit(info->skip ? 1 : 0, skip_proba);
        if (info->mb_type == 0) {
          bw.put_bit(1, 145);
          switch (info->i16_mode) {
            case DC_PRED: bw.put_bit(0, 156); bw.put_bit(0, 163); break;
            case V_PRED: bw.put_bit(0, 156); bw.put_bit(1, 163); break;
            case H_PRED: bw.put_bit(1, 156); bw.put_bit(0, 128); break;
            case TM_PRED: bw.put_bit(1, 156); bw.put_bit(1, 128); break;
          }
          for (int i = 0; i < 4; ++i) top[i] = lm[i] = info->i16_mode;
        } else {
          bw.put_bit(0, 145);
          for (int y = 0; y < 4; ++y) {
            int ymode = lm[y];
            for (int x = 0; x < 4; ++x) {
              const int mode = info->modes[y * 4 + x];
              write_i4_mode(bw, mode, &kBModesProba[(top[x] * 10 + ymode) * 9]);
              ymode = mode;
              top[x] = (uint8_t)mode;
            }
            lm[y] = (uint8_t)ymode;
          }
        }
        switch (info->uv_mode) {
          case DC_PRED: bw.put_bit(0, 142); break;
          case V_PRED: bw.put_bit(1, 142); bw.put_bit(0, 114); break;
          case H_PRED: bw.put_bit(1, 142); bw.put_bit(1, 114); bw.put_bit(0, 183); break;
          case TM_PRED: bw.put_bit(1, 142); bw.put_bit(1, 114); bw.put_bit(1, 183); break;
        }
      }
    }
    return bw.finish();
  }
  std::vector<uint8_t> emit_token_partition(int part_idx) {  // encode_token.go:304-361
    BoolWriter bw;
    if (num_parts <= 1) {
      for (uint16_t t : tokens) bw.put_bit(t & 1, t >> 8);
    } else {
      const size_t total_mb = (size_t)mb_w * mb_h;
      mb_start[total_mb] = tokens.size();
      for (size_t idx = 0; idx < total_mb; ++idx) {
        if ((int)((idx / mb_w) & (num_parts - 1)) != part_idx) continue;
        const size_t s = mb_start[idx], e = mb_start[idx + 1];  // NB: stale for skipped MBs, as in the reference
        for (size_t t = s; t < e; ++t) bw.put_bit(tokens[t] & 1, tokens[t] >> 8);
      }
    }
    return bw.finish();
  }

  // ================================================================================================
  // Serial path for Method < 3 (encode.go:1334-1336,1356: statLoop + encodeFrame; no RD, no trellis, no
  // error diffusion).  Mode decisions here do not depend on the coefficient probabilities, but the
  // probability state does evolve through the mid-stream refreshes of every pass (encode_frame.go:35-57)
  // and is restated literally, including statistics taken over not-yet-encoded (zero-state) macroblocks.
  uint8_t yuv_p[33 * BPS] = {0};  // I4 scoring scratch whose borders are never filled
  bool skip_tokens = false, skip_export = false;
  std::vector<uint32_t> s_top_nz;
  std::vector<uint8_t> s_top_nz_dc;
  uint32_t s_left_nz = 0;
  uint8_t s_left_nz_dc = 0;

  void mb_import(int mx, int my) {  // MBIterator.Import (encode_iterator.go:115)
    const int x = mx * 16, y = my * 16;
    const int ww = width - x > 16 ? 16 : width - x, hh = height - y > 16 ? 16 : height - y;
    import_block(&y_plane[(size_t)y * y_stride + x], y_stride, yuv_in + Y_OFF, ww, hh, 16);
    const int uvw = (ww + 1) >> 1, uvh = (hh + 1) >> 1;
    import_block(&u_plane[(size_t)my * 8 * uv_stride + mx * 8], uv_stride, yuv_in + U_OFF, uvw, uvh, 8);
    import_block(&v_plane[(size_t)my * 8 * uv_stride + mx * 8], uv_stride, yuv_in + V_OFF, uvw, uvh, 8);
  }
  void mb_fill_ctx(int mx, int my, const RowCtx& rc) {  // FillPredContext (encode_iterator.go:262-377)
    uint8_t* o = yuv_out;
    for (int i = 0; i < 16; ++i) o[Y_OFF - BPS + i] = my > 0 ? top_y[mx * 16 + i] : 127;
    for (int i = 0; i < 4; ++i)
      o[Y_OFF - BPS + 16 + i] = my > 0 ? (mx < mb_w - 1 ? top_y[(mx + 1) * 16 + i] : top_y[mx * 16 + 15]) : 127;
    for (int r = 1; r <= 3; ++r) memcpy(o + Y_OFF - BPS + 16 + r * 4 * BPS, o + Y_OFF - BPS + 16, 4);
    o[Y_OFF - BPS - 1] = (mx > 0 && my > 0) ? rc.top_left_y : (my > 0 ? 129 : 127);
    for (int j = 0; j < 16; ++j) o[Y_OFF - 1 + j * BPS] = mx > 0 ? rc.left_y[j] : 129;
    for (int i = 0; i < 8; ++i) {
      o[U_OFF - BPS + i] = my > 0 ? top_u[mx * 8 + i] : 127;
      o[V_OFF - BPS + i] = my > 0 ? top_v[mx * 8 + i] : 127;
    }
    o[U_OFF - BPS - 1] = (mx > 0 && my > 0) ? rc.top_left_u : (my > 0 ? 129 : 127);
    o[V_OFF - BPS - 1] = (mx > 0 && my > 0) ? rc.top_left_v : (my > 0 ? 129 : 127);
    for (int j = 0; j < 8; ++j) {
      o[U_OFF - 1 + j * BPS] = mx > 0 ? rc.left_u[j] : 129;
      o[V_OFF - 1 + j * BPS] = mx > 0 ? rc.left_v[j] : 129;
    }
  }
  void mb_export(int mx, int my, RowCtx& rc) {  // MBIterator.Export (encode_iterator.go:181-250)
    if (!skip_export) {
      const int x = mx * 16, y = my * 16;
      const int wy = x + 16 > width ? width - x : 16, hy = y + 16 > height ? height - y : 16;
      for (int j = 0; j < hy; ++j) memcpy(&y_plane[(size_t)(y + j) * y_stride + x], yuv_out + Y_OFF + j * BPS, wy);
      for (int j = 0; j < 8; ++j) {
        memcpy(&u_plane[(size_t)(my * 8 + j) * uv_stride + mx * 8], yuv_out + U_OFF + j * BPS, 8);
        memcpy(&v_plane[(size_t)(my * 8 + j) * uv_stride + mx * 8], yuv_out + V_OFF + j * BPS, 8);
      }
    }
    rc.top_left_y = top_y[mx * 16 + 15];
    rc.top_left_u = top_u[mx * 8 + 7];
    rc.top_left_v = top_v[mx * 8 + 7];
    memcpy(&top_y[mx * 16], yuv_out + Y_OFF + 15 * BPS, 16);
    memcpy(&top_u[mx * 8], yuv_out + U_OFF + 7 * BPS, 8);
    memcpy(&top_v[mx * 8], yuv_out + V_OFF + 7 * BPS, 8);
    for (int j = 0; j < 16; ++j) rc.left_y[j] = yuv_out[Y_OFF + j * BPS + 15];
    for (int j = 0; j < 8; ++j) {
      rc.left_u[j] = yuv_out[U_OFF + j * BPS + 7];
      rc.left_v[j] = yuv_out[V_OFF + j * BPS + 7];
    }
  }

  // pickBestMode, Method < 3 branch (encode_frame.go:165-189) with PickBestI16Mode / tryI4Modes / PickBestUVMode
  // (encode_analysis.go:911-1070, encode_frame.go:193-237).
  void pick_best_mode_fast(int mx, int my, MBInfo* info, const SegmentInfo* seg, RowCtx& rc) {
    // I16: SSE of the prediction + fixed mode cost
    uint64_t best16 = ~(uint64_t)0;
    int mode16 = DC_PRED;
    for (int mode = 0; mode < 4; ++mode) {
      if (mode == V_PRED && my == 0) continue;
      if (mode == H_PRED && mx == 0) continue;
      if (mode == TM_PRED && (mx == 0 || my == 0)) continue;
      pred_luma16(check_mode(mx, my, mode), yuv_out, Y_OFF);
      const uint64_t sc = rd_score(sse16x16(yuv_in + Y_OFF, yuv_out + Y_OFF), kModeFixedCost16[mode], seg->lambda_i16);
      if (sc < best16) { best16 = sc; mode16 = mode; }
    }
    if ((mx == 0 || my == 0) && is_flat_source16(yuv_in + Y_OFF)) mode16 = (mx == 0) ? DC_PRED : V_PRED;  // score keeps the loop's value
    // I4 (Method >= 2): every mode is predicted into yuv_p, whose row 0 / column 0 / columns 17.. are never written
    uint64_t best4 = ~(uint64_t)0;
    uint8_t modes4[16] = {0};
    if (cfg.method >= 2) {
      uint64_t total = 0;
      uint8_t top_m[4];
      for (int i = 0; i < 4; ++i) top_m[i] = (my > 0) ? top_modes[mx * 4 + i] : B_DC_PRED;
      for (int by = 0; by < 4; ++by)
        for (int bx = 0; bx < 4; ++bx) {
          const int b = by * 4 + bx;
          const int top_mode = (by == 0) ? top_m[bx] : modes4[b - 4];
          const int left_mode = (bx == 0) ? rc.left_modes[by] : modes4[b - 1];
          const int src_off = Y_OFF + by * 4 * BPS + bx * 4, pred_off = BPS + 1 + by * 4 * BPS + bx * 4;
          const bool has_top = (my > 0 || by > 0), has_left = (mx > 0 || bx > 0);
          uint64_t bs = ~(uint64_t)0;
          int bm = B_DC_PRED;
          for (int mode = 0; mode < 10; ++mode) {  // PickBestI4Mode (encode_analysis.go:967)
            if (!has_top && needs_top4(mode)) continue;
            if (!has_left && needs_left4(mode)) continue;
            pred_luma4(mode, yuv_p, pred_off);
            const uint64_t sc = rd_score(sse4x4(yuv_in + src_off, yuv_p + pred_off), fixed_costs_i4[top_mode][left_mode][mode], seg->lambda_i4);
            if (sc < bs) { bs = sc; bm = mode; }
          }
          modes4[b] = (uint8_t)bm;
          total += bs;
        }
      for (int i = 0; i < 4; ++i) {  // saved whether or not I4 wins (encode_frame.go:229-231)
        top_modes[mx * 4 + i] = modes4[12 + i];
        rc.left_modes[i] = modes4[3 + 4 * i];
      }
      total += (uint64_t)seg->lambda_mode * 211;
      best4 = total;
    }
    if (best4 < best16) {
      info->mb_type = 1;
      memcpy(info->modes, modes4, 16);
    } else {
      info->mb_type = 0;
      info->i16_mode = (uint8_t)mode16;
    }
    // UV
    uint64_t bestuv = ~(uint64_t)0;
    int modeuv = DC_PRED;
    for (int mode = 0; mode < 4; ++mode) {
      if (mode == V_PRED && my == 0) continue;
      if (mode == H_PRED && mx == 0) continue;
      if (mode == TM_PRED && (mx == 0 || my == 0)) continue;
      const int am = check_mode(mx, my, mode);
      pred_chroma8(am, yuv_out, U_OFF);
      pred_chroma8(am, yuv_out, V_OFF);
      int disto = 0;
      for (int by = 0; by < 2; ++by)
        for (int bx = 0; bx < 2; ++bx) {
          const int off = by * 4 * BPS + bx * 4;
          disto += sse4x4(yuv_in + U_OFF + off, yuv_out + U_OFF + off) + sse4x4(yuv_in + V_OFF + off, yuv_out + V_OFF + off);
        }
      const uint64_t sc = rd_score(disto, kModeFixedCostUV[mode], seg->lambda_uv);
      if (sc < bestuv) { bestuv = sc; modeuv = mode; }
    }
    info->uv_mode = (uint8_t)modeuv;
  }

  // One macroblock of encodeFrame (encode_frame.go:44-98) on the Method < 3 path.
  void encode_mb_serial(int mx, int my, RowCtx& rc) {
    const size_t idx = (size_t)my * mb_w + mx;
    MBInfo* info = &mb_info[idx];
    const SegmentInfo* seg = &dqm[info->segment];
    mb_import(mx, my);
    mb_fill_ctx(mx, my, rc);
    pick_best_mode_fast(mx, my, info, seg, rc);
    // encodeResiduals (encode_frame.go:350-645): plain quantisation everywhere on this path
    if (info->mb_type == 0) {
      pred_luma16(check_mode(mx, my, info->i16_mode), yuv_out, Y_OFF);
      int16_t dc_coeffs[16];
      uint32_t nz_y = 0;
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        int16_t* c = info->coeffs + b * 16;
        ftransform(yuv_in + off, yuv_out + off, c);
        dc_coeffs[b] = c[0];
        c[0] = 0;
        const int nz = quantize_coeffs(c, c, &seg->y1, 1);
        info->nz_y[b] = (uint8_t)nz;
        if (nz > 0) nz_y |= 1u << b;
      }
      int16_t wht[16];
      ftransform_wht(dc_coeffs, wht);
      const int nz_dc = quantize_coeffs(wht, info->coeffs + 384, &seg->y2, 0);
      info->nz_dc = (uint8_t)nz_dc;
      if (nz_dc > 0) nz_y |= 1u << 24;
      info->non_zero_y = nz_y;
    } else {
      uint32_t nz_y = 0;
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        int16_t* c = info->coeffs + b * 16;
        int16_t dq[16];
        pred_luma4(info->modes[b], yuv_out, off);
        ftransform(yuv_in + off, yuv_out + off, c);
        const int nz = quantize_coeffs(c, c, &seg->y1, 0);
        info->nz_y[b] = (uint8_t)nz;
        if (nz > 0) nz_y |= 1u << b;
        dequant_coeffs(c, dq, &seg->y1);
        itransform_one(yuv_out + off, dq, yuv_out + off);
      }
      info->non_zero_y = nz_y;
    }
    {
      const int am = check_mode(mx, my, info->uv_mode);
      pred_chroma8(am, yuv_out, U_OFF);
      pred_chroma8(am, yuv_out, V_OFF);
      uint32_t nz_uv = 0;
      for (int ch = 0; ch < 2; ++ch)
        for (int b = 0; b < 4; ++b) {
          const int off = (ch ? V_OFF : U_OFF) + (b >> 1) * 4 * BPS + (b & 1) * 4;
          int16_t* c = info->coeffs + (16 + ch * 4 + b) * 16;
          ftransform(yuv_in + off, yuv_out + off, c);
          const int nz = quantize_coeffs(c, c, &seg->uv, 0);
          info->nz_uv[ch * 4 + b] = (uint8_t)nz;
          if (nz > 0) nz_uv |= 1u << (ch * 4 + b);
        }
      info->non_zero_uv = nz_uv;
    }
    info->skip = (info->non_zero_y == 0 && info->non_zero_uv == 0);
    if (info->skip) {
      num_skip++;
      s_top_nz[mx] = 0; s_left_nz = 0;
      if (info->mb_type == 0) { s_top_nz_dc[mx] = 0; s_left_nz_dc = 0; }
    } else if (skip_tokens) {
      update_nz(info, &s_top_nz[mx], &s_left_nz, &s_top_nz_dc[mx], &s_left_nz_dc);
    } else {
      mb_start[idx] = tokens.size();
      walk_mb(info, &s_top_nz[mx], &s_left_nz, &s_top_nz_dc[mx], &s_left_nz_dc,
              [&](const int16_t* c, int nz, int type, int first, int ctx) { record_coeffs(c, nz, type, first, ctx); });
    }
    // reconstructMB (encode_frame.go:855): I4 is already reconstructed
    if (info->mb_type == 0) {
      int16_t wht_dq[16], wht_buf[256], dq[16];
      dequant_coeffs(info->coeffs + 384, wht_dq, &seg->y2);
      transform_wht(wht_dq, wht_buf);
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        dequant_coeffs(info->coeffs + b * 16, dq, &seg->y1);
        dq[0] = wht_buf[b * 16];
        itransform_one(yuv_out + off, dq, yuv_out + off);
      }
    }
    for (int b = 0; b < 4; ++b) {
      int16_t dq[16];
      const int o = (b >> 1) * 4 * BPS + (b & 1) * 4;
      dequant_coeffs(info->coeffs + (16 + b) * 16, dq, &seg->uv);
      itransform_one(yuv_out + U_OFF + o, dq, yuv_out + U_OFF + o);
      dequant_coeffs(info->coeffs + (20 + b) * 16, dq, &seg->uv);
      itransform_one(yuv_out + V_OFF + o, dq, yuv_out + V_OFF + o);
    }
    mb_export(mx, my, rc);
  }

  // ---- serial RD path (Method >= 3 when the reference does not go row-parallel: mbH < 4 or GOMAXPROCS == 1, encode.go:1356)
  int8_t top_derr[1024][2][2];  // enc.topDerr[x][ch] (encode.go:150); sized for mb_w <= 1024
  int8_t left_derr[2][2];       // enc.leftDerr[ch]: NOT reset at row starts
  static int quantize_single(int16_t* v, const SegmentQuant* sq) {  // encode_frame.go:510
    int V = *v, sign = 1;
    if (V < 0) { sign = -1; V = -V; }
    const int zthresh = ((1 << 17) - 1 - sq->dc_bias) / sq->dc_iquant;  // DCZthresh (encode.go:1175)
    if (V > zthresh) {
      const int qV = (int)(((uint32_t)V * (uint32_t)sq->dc_iquant + (uint32_t)sq->dc_bias) >> 17) * sq->dc_quant;
      const int err = V - qV;
      *v = (int16_t)(sign * qV);
      return (sign * err) >> 1;
    }
    *v = 0;
    return (sign * V) >> 1;
  }
  // PickBestI4ModeRD (encode_analysis.go:1216): every eligible mode, plain quantisation, in mode order
  void pick_best_i4_all(int off, const SegmentInfo* seg, int top_mode, int left_mode, bool has_top, bool has_left, int nz_ctx,
                        int* best_mode, int* best_rate, int* best_disto) {
    uint64_t best_score = ~(uint64_t)0;
    *best_mode = B_DC_PRED; *best_rate = 0; *best_disto = 0;
    const uint8_t* src = yuv_in + off;
    uint8_t* pred_buf = yuv_out2;
    for (int mode = 0; mode < 10; ++mode) {
      if (!has_top && needs_top4(mode)) continue;
      if (!has_left && needs_left4(mode)) continue;
      int16_t c[16], q[16], dq[16];
      uint8_t recon[4 * BPS];
      pred_luma4(mode, pred_buf, off);
      ftransform(src, pred_buf + off, c);
      const int nz = quantize_coeffs(c, q, &seg->y1, 0);
      dequant_coeffs(q, dq, &seg->y1);
      itransform_one(pred_buf + off, dq, recon);
      int disto = sse4x4(src, recon);
      if (seg->tlambda_sd > 0) disto += (seg->tlambda_sd * tdisto4x4(src, recon) + 128) >> 8;
      if (256 * (uint64_t)disto >= best_score) continue;
      int rate = 0;
      if (mode > 0 && is_flat(q, 1, 3)) rate = 140;
      rate += token_cost(q, nz, 3, &proba, nz_ctx, 0);
      rate += fixed_costs_i4[top_mode][left_mode][mode];
      const uint64_t score = rd_score(disto, rate, seg->lambda_i4);
      if (score < best_score) {
        best_score = score;
        *best_mode = mode; *best_rate = rate; *best_disto = disto;
        memcpy(tmp_best_dq, dq, sizeof(dq));
        memcpy(tmp_best_q, q, sizeof(q));
        tmp_best_nz = nz;
      }
    }
  }
  // tryI4ModesRD (encode_frame.go:242-347): like the parallel twin, but Method 3 scores all modes, and the mode-cost
  // context is saved here -- only when the search ran to completion, whatever the I4-vs-I16 outcome.
  uint64_t try_i4_modes_serial(int mx, int my, MBInfo* info, const SegmentInfo* seg, uint8_t* modes, RowCtx& rc, uint64_t i16_score,
                               uint32_t top_nz_v, uint32_t left_nz_v) {
    int total_rate = 0, total_disto = 0, total_header_bits = 0;
    uint8_t top_m[4];
    for (int i = 0; i < 4; ++i) top_m[i] = (my > 0) ? top_modes[mx * 4 + i] : B_DC_PRED;
    memcpy(yuv_out2, yuv_out, YUV_SIZE);
    uint32_t tnz = top_nz_v & 0x0f, lnz = left_nz_v & 0x0f, l = 0;
    bool early_exit = false;
    const int max_modes = cfg.quality < 50 ? 2 : 3;
    for (int by = 0; by < 4 && !early_exit; ++by) {
      l = lnz & 1;
      for (int bx = 0; bx < 4; ++bx) {
        const int b = by * 4 + bx;
        const int top_mode = (by == 0) ? top_m[bx] : modes[b - 4];
        const int left_mode = (bx == 0) ? rc.left_modes[by] : modes[b - 1];
        const int off = Y_OFF + by * 4 * BPS + bx * 4;
        const bool has_top = (my > 0 || by > 0), has_left = (mx > 0 || bx > 0);
        int nz_ctx = (int)l + (int)(tnz & 1);
        if (nz_ctx > 2) nz_ctx = 2;
        int best_mode, rate, disto;
        if (cfg.method >= 4) pick_best_i4(off, seg, top_mode, left_mode, has_top, has_left, nz_ctx, max_modes, true, &best_mode, &rate, &disto);
        else pick_best_i4_all(off, seg, top_mode, left_mode, has_top, has_left, nz_ctx, &best_mode, &rate, &disto);
        modes[b] = (uint8_t)best_mode;
        total_rate += rate;
        total_disto += disto;
        total_header_bits += fixed_costs_i4[top_mode][left_mode][best_mode];
        memcpy(info->coeffs + b * 16, tmp_best_q, 32);
        const int nz = tmp_best_nz;
        info->nz_y[b] = (uint8_t)nz;
        if (rd_score(total_disto, total_rate + 211, seg->lambda_mode) >= i16_score) { early_exit = true; break; }
        if (total_header_bits > 15000) { early_exit = true; break; }
        pred_luma4(best_mode, yuv_out2, off);
        itransform_one(yuv_out2 + off, tmp_best_dq, yuv_out2 + off);
        l = nz > 0;
        tnz = (tnz >> 1) | (l << 7);
      }
      tnz >>= 4;
      lnz = (lnz >> 1) | (l << 7);
    }
    if (early_exit) return ~(uint64_t)0;
    for (int i = 0; i < 4; ++i) {
      top_modes[mx * 4 + i] = modes[12 + i];
      rc.left_modes[i] = modes[3 + 4 * i];
    }
    return rd_score(total_disto, total_rate + 211, seg->lambda_mode);
  }
  void encode_mb_serial_rd(int mx, int my, RowCtx& rc) {
    const size_t idx = (size_t)my * mb_w + mx;
    MBInfo* info = &mb_info[idx];
    const SegmentInfo* seg = &dqm[info->segment];
    mb_import(mx, my);
    mb_fill_ctx(mx, my, rc);
    const uint32_t tnz_val = s_top_nz[mx], lnz_val = s_left_nz;
    bool i4_cached = false;
    {  // pickBestMode, Method >= 3 branch (encode_frame.go:122-164)
      int best16, rate16, disto16;
      pick_best_i16(mx, my, seg, tnz_val, lnz_val, s_top_nz_dc[mx], s_left_nz_dc, &best16, &rate16, &disto16);
      const uint64_t score16 = rd_score(disto16, rate16, seg->lambda_mode);
      uint8_t modes4[16] = {0};
      const uint64_t score4 = try_i4_modes_serial(mx, my, info, seg, modes4, rc, score16, tnz_val, lnz_val);
      if (score4 < score16) {
        info->mb_type = 1;
        memcpy(info->modes, modes4, 16);
        if (cfg.method >= 4) {
          i4_cached = true;
          for (int j = 0; j < 16; ++j) memcpy(yuv_out + Y_OFF + j * BPS, yuv_out2 + Y_OFF + j * BPS, 16);
        }
      } else {
        info->mb_type = 0;
        info->i16_mode = (uint8_t)best16;
        pred_luma16(check_mode(mx, my, best16), yuv_out, Y_OFF);
      }
      const int best_uv = pick_best_uv(mx, my, seg, tnz_val, lnz_val);
      info->uv_mode = (uint8_t)best_uv;
      pred_chroma8(check_mode(mx, my, best_uv), yuv_out, U_OFF);
      pred_chroma8(check_mode(mx, my, best_uv), yuv_out, V_OFF);
    }
    // encodeResiduals (encode_frame.go:350-645)
    if (info->mb_type == 0) {
      int16_t dc_coeffs[16];
      uint32_t nz_y = 0, tnz = tnz_val & 0x0f, lnz = lnz_val & 0x0f;
      for (int by = 0; by < 4; ++by) {
        uint32_t l = lnz & 1;
        for (int bx = 0; bx < 4; ++bx) {
          const int b = by * 4 + bx, off = Y_OFF + by * 4 * BPS + bx * 4;
          int16_t* c = info->coeffs + b * 16;
          ftransform(yuv_in + off, yuv_out + off, c);
          dc_coeffs[b] = c[0];
          c[0] = 0;
          int nz;
          if (cfg.method >= 4) {
            int ctx = (int)l + (int)(tnz & 1);
            if (ctx > 2) ctx = 2;
            nz = trellis_quantize_block(c, c, &seg->y1, 1, 0, ctx, &proba, seg->tlambda_i16);
          } else {
            nz = quantize_coeffs(c, c, &seg->y1, 1);
          }
          info->nz_y[b] = (uint8_t)nz;
          if (nz > 0) nz_y |= 1u << b;
          l = nz > 0;
          tnz = (tnz >> 1) | (l << 7);
        }
        tnz >>= 4;
        lnz = (lnz >> 1) | (l << 7);
      }
      int16_t wht[16];
      ftransform_wht(dc_coeffs, wht);
      const int nz_dc = quantize_coeffs(wht, info->coeffs + 384, &seg->y2, 0);
      info->nz_dc = (uint8_t)nz_dc;
      if (nz_dc > 0) nz_y |= 1u << 24;
      info->non_zero_y = nz_y;
    } else if (i4_cached) {
      uint32_t nz_y = 0;
      for (int b = 0; b < 16; ++b) if (info->nz_y[b] > 0) nz_y |= 1u << b;
      info->non_zero_y = nz_y;
    } else {
      uint32_t nz_y = 0;
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        int16_t* c = info->coeffs + b * 16;
        int16_t dq[16];
        pred_luma4(info->modes[b], yuv_out, off);
        ftransform(yuv_in + off, yuv_out + off, c);
        const int nz = quantize_coeffs(c, c, &seg->y1, 0);
        info->nz_y[b] = (uint8_t)nz;
        if (nz > 0) nz_y |= 1u << b;
        dequant_coeffs(c, dq, &seg->y1);
        itransform_one(yuv_out + off, dq, yuv_out + off);
      }
      info->non_zero_y = nz_y;
    }
    {  // encodeUVResiduals with DC error diffusion (useDerr = Method >= 3, encode_frame.go:569-645)
      for (int ch = 0; ch < 2; ++ch)
        for (int b = 0; b < 4; ++b) {
          const int off = (ch ? V_OFF : U_OFF) + (b >> 1) * 4 * BPS + (b & 1) * 4;
          ftransform(yuv_in + off, yuv_out + off, info->coeffs + (16 + ch * 4 + b) * 16);
        }
      int8_t derr[2][3];
      for (int ch = 0; ch < 2; ++ch) {  // correctDCValues (encode_frame.go:529)
        const int8_t* top = top_derr[mx][ch];
        const int8_t* left = left_derr[ch];
        int16_t* c0 = info->coeffs + (16 + ch * 4 + 0) * 16;
        int16_t* c1 = info->coeffs + (16 + ch * 4 + 1) * 16;
        int16_t* c2 = info->coeffs + (16 + ch * 4 + 2) * 16;
        int16_t* c3 = info->coeffs + (16 + ch * 4 + 3) * 16;
        *c0 += (int16_t)((7 * (int)top[0] + 8 * (int)left[0]) >> 3);
        const int err0 = quantize_single(c0, &seg->uv);
        *c1 += (int16_t)((7 * (int)top[1] + 8 * err0) >> 3);
        const int err1 = quantize_single(c1, &seg->uv);
        *c2 += (int16_t)((7 * err0 + 8 * (int)left[1]) >> 3);
        const int err2 = quantize_single(c2, &seg->uv);
        *c3 += (int16_t)((7 * err1 + 8 * err2) >> 3);
        const int err3 = quantize_single(c3, &seg->uv);
        derr[ch][0] = (int8_t)err1; derr[ch][1] = (int8_t)err2; derr[ch][2] = (int8_t)err3;
      }
      uint32_t nz_uv = 0;
      for (int ch = 0; ch < 2; ++ch)
        for (int b = 0; b < 4; ++b) {
          int16_t* c = info->coeffs + (16 + ch * 4 + b) * 16;
          const int nz = quantize_coeffs(c, c, &seg->uv, 0);
          info->nz_uv[ch * 4 + b] = (uint8_t)nz;
          if (nz > 0) nz_uv |= 1u << (ch * 4 + b);
        }
      for (int ch = 0; ch < 2; ++ch) {  // storeDiffusionErrors (encode_frame.go:557)
        int8_t* top = top_derr[mx][ch];
        int8_t* left = left_derr[ch];
        left[0] = derr[ch][0];
        left[1] = (int8_t)((3 * (int)derr[ch][2]) >> 2);
        top[0] = derr[ch][1];
        top[1] = (int8_t)(derr[ch][2] - left[1]);
      }
      info->non_zero_uv = nz_uv;
    }
    info->skip = (info->non_zero_y == 0 && info->non_zero_uv == 0);
    if (info->skip) {
      num_skip++;
      s_top_nz[mx] = 0; s_left_nz = 0;
      if (info->mb_type == 0) { s_top_nz_dc[mx] = 0; s_left_nz_dc = 0; }
    } else if (skip_tokens) {
      update_nz(info, &s_top_nz[mx], &s_left_nz, &s_top_nz_dc[mx], &s_left_nz_dc);
    } else {
      mb_start[idx] = tokens.size();
      walk_mb(info, &s_top_nz[mx], &s_left_nz, &s_top_nz_dc[mx], &s_left_nz_dc,
              [&](const int16_t* c, int nz, int type, int first, int ctx) { record_coeffs(c, nz, type, first, ctx); });
    }
    if (info->mb_type == 0) {  // reconstructMB
      int16_t wht_dq[16], wht_buf[256], dq[16];
      dequant_coeffs(info->coeffs + 384, wht_dq, &seg->y2);
      transform_wht(wht_dq, wht_buf);
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        dequant_coeffs(info->coeffs + b * 16, dq, &seg->y1);
        dq[0] = wht_buf[b * 16];
        itransform_one(yuv_out + off, dq, yuv_out + off);
      }
    }
    for (int b = 0; b < 4; ++b) {
      int16_t dq[16];
      const int o = (b >> 1) * 4 * BPS + (b & 1) * 4;
      dequant_coeffs(info->coeffs + (16 + b) * 16, dq, &seg->uv);
      itransform_one(yuv_out + U_OFF + o, dq, yuv_out + U_OFF + o);
      dequant_coeffs(info->coeffs + (20 + b) * 16, dq, &seg->uv);
      itransform_one(yuv_out + V_OFF + o, dq, yuv_out + V_OFF + o);
    }
    mb_export(mx, my, rc);
  }

  // collectAllStats (encode_proba.go:171-313): the whole mb_info array, whatever it currently holds
  void collect_all_stats(ProbaStats st) {
    memset(st, 0, sizeof(ProbaStats));
    std::vector<uint32_t> tnz(mb_w, 0);
    std::vector<uint8_t> tdc(mb_w, 0);
    for (int my = 0; my < mb_h; ++my) {
      uint32_t lnz = 0;
      uint8_t ldc = 0;
      for (int mx = 0; mx < mb_w; ++mx) {
        const MBInfo* info = &mb_info[(size_t)my * mb_w + mx];
        if (info->skip) {
          tnz[mx] = 0; lnz = 0;
          if (info->mb_type == 0) { tdc[mx] = 0; ldc = 0; }
          continue;
        }
        walk_mb(info, &tnz[mx], &lnz, &tdc[mx], &ldc,
                [&](const int16_t* c, int nz, int type, int first, int ctx) { collect_coeff_stats(c, nz, type, first, ctx, st); });
      }
    }
  }
  void refresh_probas() {  // encode_frame.go:113-117
    static thread_local ProbaStats st;
    collect_all_stats(st);
    optimize_proba(st);
  }
  // encodeFrame (encode_frame.go:15-108)
  void encode_frame_serial_pass() {
    proba_hist.assign(1, proba); hist_starts.assign(1, 0);  // test taps (hostcheck.cc)
    top_y.assign(mb_w * 16, 127); top_u.assign(mb_w * 8, 127); top_v.assign(mb_w * 8, 127);
    top_modes.assign(mb_w * 4, B_DC_PRED);
    s_top_nz.assign(mb_w, 0); s_top_nz_dc.assign(mb_w, 0);
    s_left_nz = 0; s_left_nz_dc = 0;
    num_skip = 0;
    // topDerr / leftDerr are zeroed when the encoder is created or recycled (resetForReuse), not per pass or per row
    const int total_mb = mb_w * mb_h;
    int max_count = total_mb >> 3;
    if (max_count < 96) max_count = 96;
    int refresh_cnt = max_count;
    for (int my = 0; my < mb_h; ++my) {
      RowCtx rc;  // resetLeftContext (encode_iterator.go:71)
      memset(rc.left_y, 129, 16); memset(rc.left_u, 129, 8); memset(rc.left_v, 129, 8);
      memset(rc.left_modes, B_DC_PRED, 4);
      rc.top_left_y = rc.top_left_u = rc.top_left_v = 127;
      rc.left_nz = 0; rc.left_nz_dc = 0;
      for (int mx = 0; mx < mb_w; ++mx) {
        if (mx == 0) { s_left_nz = 0; s_left_nz_dc = 0; }
        if (--refresh_cnt < 0) {
          refresh_probas(); refresh_cnt = max_count;
          proba_hist.push_back(proba); hist_starts.push_back(my * mb_w + mx);  // test tap: what the following macroblocks are recorded under
        }
        if (cfg.method >= 3) encode_mb_serial_rd(mx, my, rc); else encode_mb_serial(mx, my, rc);
      }
    }
    if (num_skip > 0) skip_proba = (uint8_t)((total_mb - num_skip) * 255 / total_mb);
  }
  void stat_loop(int passes) {  // encode.go:1405-1437
    int n = passes < 1 ? 1 : (passes > 10 ? 10 : passes);
    for (int pass = 0; pass < n; ++pass) {
      skip_tokens = true; skip_export = true;
      encode_frame_serial_pass();
      skip_tokens = false; skip_export = false;
      static thread_local ProbaStats st;
      collect_all_stats(st);
      if (optimize_proba(st) == 0) break;
    }
  }
  void rerecord_all_tokens() { record_all_tokens(nullptr); }  // encode_proba.go:317 (same walk, current probabilities)

  std::vector<uint8_t> assemble_frame() {  // emitFrame / assembleFrame (encode_syntax.go:27-172)
    std::vector<uint8_t> part0 = emit_partition0();
    std::vector<std::vector<uint8_t>> parts(num_parts);
    for (int i = 0; i < num_parts; ++i) parts[i] = emit_token_partition(i);
    std::vector<uint8_t> out;
    const uint32_t tag = (0u) | (0u << 1) | (1u << 4) | ((uint32_t)part0.size() << 5);
    out.push_back((uint8_t)tag); out.push_back((uint8_t)(tag >> 8)); out.push_back((uint8_t)(tag >> 16));
    out.push_back(0x9d); out.push_back(0x01); out.push_back(0x2a);
    out.push_back((uint8_t)(width & 0xff)); out.push_back((uint8_t)((width & 0x3fff) >> 8));
    out.push_back((uint8_t)(height & 0xff)); out.push_back((uint8_t)((height & 0x3fff) >> 8));
    out.insert(out.end(), part0.begin(), part0.end());
    for (int i = 0; i + 1 < num_parts; ++i) {
      const size_t sz = parts[i].size();
      out.push_back((uint8_t)sz); out.push_back((uint8_t)(sz >> 8)); out.push_back((uint8_t)(sz >> 16));
    }
    for (auto& p : parts) out.insert(out.end(), p.begin(), p.end());
    return out;
  }

  // ---- rate control (encode.go:1440-1590): passStats, computeNextQ, adjustQuantForTarget
  struct PassStats { bool is_first; double dq, q, last_q, qmin, qmax, value, last_value, target; bool do_size; } rc;
  bool rc_init = false;
  double compute_next_q() {  // encode.go:1505
    double dq;
    if (rc.is_first) {
      dq = rc.value > rc.target ? -rc.dq : rc.dq;
      rc.is_first = false;
    } else if (rc.value != rc.last_value) {
      const double slope = (rc.target - rc.value) / (rc.last_value - rc.value);
      dq = slope * (rc.last_q - rc.q);
    } else {
      dq = 0;
    }
    if (dq < -30) dq = -30;
    if (dq > 30) dq = 30;
    rc.dq = dq; rc.last_q = rc.q; rc.last_value = rc.value;
    rc.q = rc.q + dq;
    if (rc.q < rc.qmin) rc.q = rc.qmin;
    if (rc.q > rc.qmax) rc.q = rc.qmax;
    return rc.q;
  }
  bool adjust_quant_for_target() {  // encode.go:1544
    if (!rc_init) {  // initPassStats (encode.go:1455)
      rc.do_size = cfg.target_size > 0;
      rc.target = rc.do_size ? (double)cfg.target_size : (cfg.target_psnr > 0 ? (double)cfg.target_psnr : 40.0);
      rc.qmin = cfg.qmin; rc.qmax = cfg.qmax <= 0 ? 100.0 : (double)cfg.qmax;
      double q = cfg.quality;
      if (q < rc.qmin) q = rc.qmin;
      if (q > rc.qmax) q = rc.qmax;
      rc.is_first = true; rc.dq = 10.0; rc.q = q; rc.last_q = q; rc.value = 0; rc.last_value = 0;
      rc_init = true;
    }
    if (rc.do_size) {
      rc.value = (double)assemble_frame().size();  // trial frame with the probabilities as they stand (not yet optimised)
    } else {
      // MBEncInfo.Disto is read here but never written anywhere in the reference: total distortion 0 -> 99 dB
      rc.value = 99.0;
    }
    if (std::fabs(rc.dq) <= 0.4 && !rc.is_first) return true;
    const double next_q = compute_next_q();
    cfg.quality = (int)(next_q + 0.5);
    set_segment_params(num_segments);
    build_segment_header(num_segments);
    y_plane = src_y; u_plane = src_u; v_plane = src_v;  // restoreSourcePixels (encode.go:1606)
    return false;
  }

  // EncodeFrame (encode.go:1324); returns the raw VP8 frame.  Method >= 3 -> row-parallel path semantics
  // (GOMAXPROCS > 1, mbH >= 4); Method < 3 -> statLoop + serial encodeFrame.
  std::vector<uint8_t> encode_frame() {
    analysis();
    set_segment_probas();
    if (cfg.method < 3) stat_loop(cfg.pass);
    // doSearch (TargetSize / TargetPSNR) always takes the serial encodeFrame, at least three passes (encode.go:1338-1374)
    const bool do_search = cfg.target_size > 0 || cfg.target_psnr > 0;
    if (cfg.method < 3 || mb_h < 4 || cfg.force_serial || do_search) {  // useParallel == false (encode.go:1356)
      int max_passes = cfg.pass > 1 ? cfg.pass : 1;
      if (do_search && max_passes < 3) max_passes = 3;
      for (int pass = 0; pass < max_passes; ++pass) {
        tokens.clear();
        encode_frame_serial_pass();
        if (!do_search) break;
        if (adjust_quant_for_target()) break;
      }
      static thread_local ProbaStats st3;
      collect_all_stats(st3);
      proba_pre_final = proba;  // test tap
      if (optimize_proba(st3) > 0) rerecord_all_tokens();
      return assemble_frame();
    }
    // Phase A
    top_y.assign(mb_w * 16, 127); top_u.assign(mb_w * 8, 127); top_v.assign(mb_w * 8, 127);
    top_modes.assign(mb_w * 4, B_DC_PRED);
    top_nz.assign(mb_w, 0); top_nz_dc.assign(mb_w, 0);
    for (int my = 0; my < mb_h; ++my) {
      RowCtx rc;
      memset(rc.left_y, 129, 16); memset(rc.left_u, 129, 8); memset(rc.left_v, 129, 8);
      memset(rc.left_modes, B_DC_PRED, 4);
      rc.top_left_y = rc.top_left_u = rc.top_left_v = 127;
      rc.left_nz = 0; rc.left_nz_dc = 0;
      for (int mx = 0; mx < mb_w; ++mx) encode_mb(mx, my, rc);
    }
    // Phase B + final probabilities
    static thread_local ProbaStats st;
    record_all_tokens(st);
    const int total_mb = mb_w * mb_h;
    if (num_skip > 0) skip_proba = (uint8_t)((total_mb - num_skip) * 255 / total_mb);
    if (optimize_proba(st) > 0) record_all_tokens(nullptr);
    return assemble_frame();
  }
};

// writeRIFFSimple (encode.go:968-997)
static inline std::vector<uint8_t> riff_wrap(const std::vector<uint8_t>& vp8) {
  const uint32_t payload = (uint32_t)vp8.size(), padded = payload + (payload & 1);
  const uint32_t riff_size = 4 + 8 + padded;
  std::vector<uint8_t> out(8 + riff_size, 0);
  memcpy(&out[0], "RIFF", 4);
  out[4] = (uint8_t)riff_size; out[5] = (uint8_t)(riff_size >> 8); out[6] = (uint8_t)(riff_size >> 16); out[7] = (uint8_t)(riff_size >> 24);
  memcpy(&out[8], "WEBP", 4);
  memcpy(&out[12], "VP8 ", 4);
  out[16] = (uint8_t)payload; out[17] = (uint8_t)(payload >> 8); out[18] = (uint8_t)(payload >> 16); out[19] = (uint8_t)(payload >> 24);
  if (payload) memcpy(&out[20], vp8.data(), payload);
  return out;
}

}  // namespace orc
