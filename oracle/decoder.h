// ORACLE -- TEST INFRASTRUCTURE ONLY (see vp8_common.h).
// Restates the reference VP8 lossy decoder: internal/lossy/decode.go, decode_tree.go,
// decode_mb.go, decode_quant.go, decode_frame.go and internal/bitio/reader_bool.go.
#pragma once
#include "dsp.h"
#include <vector>
#include <string>

namespace orc {

// Boolean decoder (reader_bool.go:27-230).  Bit-serial restatement of the same arithmetic
// (range kept as range-1 in [127,254]); byte-at-a-time refill; reads zeros past the end.
struct BoolReader {
  const uint8_t* buf = 0;
  size_t size = 0, pos = 0;
  uint64_t value = 0;
  uint32_t range = 254;
  int bits = -8;
  bool eof = false;
  void init(const uint8_t* d, size_t n) {
    buf = d; size = n; pos = 0; value = 0; range = 254; bits = -8; eof = false;
    load();
  }
  void load() {  // loadNewBytes / loadFinalBytes (bitio/reader_bool.go:40-75): 56 bits at a time while 8 bytes remain, then byte-wise,
                 // one virtual zero byte past the end, then EOF.  For a well-formed partition the width of the load does not
                 // matter; it does for a corrupt one that starts with 0xff (value >> bits > range from the first bit on).
    if (pos + 8 <= size) {
      uint64_t in = 0;
      for (int i = 0; i < 7; ++i) in = (in << 8) | buf[pos + i];
      value = in | (value << 56);
      pos += 7;
      bits += 56;
    } else if (pos < size) {
      bits += 8;
      value = (uint64_t)buf[pos++] | (value << 8);
    } else if (!eof) {
      value <<= 8;
      bits += 8;
      eof = true;
    } else {
      bits = 0;
    }
  }
  int get_bit(int prob) {
    uint32_t r = range;
    if (bits < 0) load();
    const int p = bits;
    const uint32_t split = (r * (uint32_t)prob) >> 8;
    const uint32_t v = (uint32_t)(value >> p);
    int bit;
    if (v > split) {
      r -= split;
      value -= (uint64_t)(split + 1) << p;
      bit = 1;
    } else {
      r = split + 1;
      bit = 0;
    }
    int shift = 0;
    while ((r << shift) < 128) ++shift;  // 7 ^ floor(log2(r))
    r <<= shift;
    bits -= shift;
    range = r - 1;
    return bit;
  }
  uint32_t get_value(int n) {
    uint32_t v = 0;
    for (int i = n - 1; i >= 0; --i) v |= (uint32_t)get_bit(0x80) << i;
    return v;
  }
  int get_signed_value(int n) {
    const int v = (int)get_value(n);
    return get_bit(0x80) ? -v : v;
  }
  int get_signed(int v) { return get_bit(0x80) ? -v : v; }
};

struct QuantMatrix {  // decode_quant.go:5-11
  int y1[2], y2[2], uv[2];
};
struct FInfo {  // decode.go:107-112
  uint8_t f_limit, f_ilevel, f_inner, hev_thresh;
};
struct MBData {  // decode.go:121-131
  int16_t coeffs[384];
  uint8_t is_i4x4;
  uint8_t imodes[16];
  uint8_t uvmode;
  uint32_t non_zero_y, non_zero_uv;
  uint8_t skip, segment;
};

struct Decoder {
  int width = 0, height = 0, mb_w = 0, mb_h = 0;
  // headers
  bool use_segment = false, update_map = false, absolute_delta = true;
  int8_t seg_quant[4] = {0, 0, 0, 0}, seg_filter[4] = {0, 0, 0, 0};
  bool f_simple = false, use_lf_delta = false;
  int f_level = 0, f_sharpness = 0, ref_lf_delta[4] = {0, 0, 0, 0}, mode_lf_delta[4] = {0, 0, 0, 0};
  int filter_type = 0;
  int num_parts_minus_one = 0;
  Proba proba;
  bool use_skip_proba = false;
  uint8_t skip_p = 0;
  QuantMatrix dqm[4];
  FInfo fstrengths[4][2];
  BoolReader br, parts[8];
  // frame-sized outputs
  std::vector<MBData> mbs;
  std::vector<FInfo> finfo;
  std::vector<uint8_t> y, u, v;  // strides 16*mb_w / 8*mb_w
  int y_stride = 0, uv_stride = 0;
  std::string err;
  // Cross-validation aid ONLY: libwebp treats an MB whose residuals parse to all-zero as
  // skipped when deciding inner-edge filtering (vp8_dec.c VP8DecodeMB: skip = ParseResiduals()).
  // The reference does not (decode_mb.go:277-296: parseResiduals has no result, FInner |= !skipflag),
  // so it filters inner edges of such MBs.  Streams written by the reference encoder never contain
  // that case (all-zero MBs always carry the skip flag), so both rules agree on them.
  bool libwebp_inner_rule = false;

  bool fail(const char* m) { err = m; return false; }

  // decode.go:245-329
  bool parse_headers(const uint8_t* data, size_t len) {
    if (len < 4) return fail("vp8: truncated header");
    const uint32_t bits = data[0] | (data[1] << 8) | (data[2] << 16);
    const bool key = !(bits & 1);
    const int profile = (bits >> 1) & 7;
    const bool show = (bits >> 4) & 1;
    const uint32_t part_len = bits >> 5;
    if (profile > 3) return fail("vp8: bad profile");
    if (!show) return fail("vp8: frame not displayable");
    if (!key) return fail("vp8: not a keyframe");
    const uint8_t* buf = data + 3;
    size_t n = len - 3;
    if (n < 7) return fail("vp8: truncated picture header");
    if (buf[0] != 0x9d || buf[1] != 0x01 || buf[2] != 0x2a) return fail("vp8: bad signature");
    width = (buf[3] | (buf[4] << 8)) & 0x3fff;
    height = (buf[5] | (buf[6] << 8)) & 0x3fff;
    buf += 7; n -= 7;
    if (width == 0 || height == 0) return fail("vp8: zero dimensions");
    mb_w = (width + 15) >> 4;
    mb_h = (height + 15) >> 4;
    reset_proba(&proba);
    absolute_delta = true;
    if (part_len > n) return fail("vp8: bad partition length");
    br.init(buf, part_len);
    const uint8_t* token_buf = buf + part_len;
    const size_t token_len = n - part_len;
    br.get_bit(0x80);  // colorspace
    br.get_bit(0x80);  // clamp type
    // segment header (decode.go:332)
    use_segment = br.get_bit(0x80);
    if (use_segment) {
      update_map = br.get_bit(0x80);
      if (br.get_bit(0x80)) {
        absolute_delta = br.get_bit(0x80);
        for (int s = 0; s < 4; ++s) seg_quant[s] = br.get_bit(0x80) ? (int8_t)br.get_signed_value(7) : 0;
        for (int s = 0; s < 4; ++s) seg_filter[s] = br.get_bit(0x80) ? (int8_t)br.get_signed_value(6) : 0;
      }
      if (update_map)
        for (int s = 0; s < 3; ++s) proba.segments[s] = br.get_bit(0x80) ? (uint8_t)br.get_value(8) : 255;
    } else {
      update_map = false;
    }
    if (br.eof) return fail("vp8: premature EOF in segment header");
    // filter header (decode.go:376)
    f_simple = br.get_bit(0x80);
    f_level = br.get_value(6);
    f_sharpness = br.get_value(3);
    use_lf_delta = br.get_bit(0x80);
    if (use_lf_delta && br.get_bit(0x80)) {
      for (int i = 0; i < 4; ++i) if (br.get_bit(0x80)) ref_lf_delta[i] = br.get_signed_value(6);
      for (int i = 0; i < 4; ++i) if (br.get_bit(0x80)) mode_lf_delta[i] = br.get_signed_value(6);
    }
    filter_type = (f_level == 0) ? 0 : (f_simple ? 1 : 2);
    // partitions (decode.go:409)
    num_parts_minus_one = (1 << br.get_value(2)) - 1;
    const int last = num_parts_minus_one;
    if (token_len < (size_t)3 * last) return fail("vp8: not enough data for partition sizes");
    const uint8_t* part_start = token_buf + 3 * last;
    size_t size_left = token_len - 3 * last;
    const uint8_t* sz = token_buf;
    for (int p = 0; p < last; ++p, sz += 3) {
      size_t psize = sz[0] | (sz[1] << 8) | (sz[2] << 16);
      if (psize > size_left) return fail("vp8: partition size exceeds remaining data");
      parts[p].init(part_start, psize);
      part_start += psize;
      size_left -= psize;
    }
    parts[last].init(part_start, size_left);
    // quantizer (decode_quant.go:27)
    {
      const int base_q0 = br.get_value(7);
      const int dqy1_dc = br.get_bit(0x80) ? br.get_signed_value(4) : 0;
      const int dqy2_dc = br.get_bit(0x80) ? br.get_signed_value(4) : 0;
      const int dqy2_ac = br.get_bit(0x80) ? br.get_signed_value(4) : 0;
      const int dquv_dc = br.get_bit(0x80) ? br.get_signed_value(4) : 0;
      const int dquv_ac = br.get_bit(0x80) ? br.get_signed_value(4) : 0;
      for (int i = 0; i < 4; ++i) {
        int q;
        if (use_segment) {
          q = seg_quant[i];
          if (!absolute_delta) q += base_q0;
        } else {
          if (i > 0) { dqm[i] = dqm[0]; continue; }
          q = base_q0;
        }
        QuantMatrix& m = dqm[i];
        m.y1[0] = kDcTable[clampi(q + dqy1_dc, 0, 127)];
        m.y1[1] = kAcTable[clampi(q, 0, 127)];
        m.y2[0] = kDcTable[clampi(q + dqy2_dc, 0, 127)] * 2;
        m.y2[1] = (kAcTable[clampi(q + dqy2_ac, 0, 127)] * 101581) >> 16;
        if (m.y2[1] < 8) m.y2[1] = 8;
        m.uv[0] = kDcTable[clampi(q + dquv_dc, 0, 117)];
        m.uv[1] = kAcTable[clampi(q + dquv_ac, 0, 127)];
      }
    }
    br.get_bit(0x80);  // update_proba flag (ignored)
    // coefficient probabilities (decode_tree.go:7)
    for (int t = 0; t < 4; ++t)
      for (int b = 0; b < 8; ++b)
        for (int c = 0; c < 3; ++c)
          for (int p = 0; p < 11; ++p) {
            const int idx = ((t * 8 + b) * 3 + c) * 11 + p;
            proba.bands[t][b][c][p] = br.get_bit(kCoeffsUpdateProba[idx]) ? (uint8_t)br.get_value(8)
                                                                          : kCoeffsProba0[idx];
          }
    use_skip_proba = br.get_bit(0x80);
    if (use_skip_proba) skip_p = (uint8_t)br.get_value(8);
    return true;
  }

  // decode_frame.go:220-280
  void precompute_filter_strengths() {
    if (filter_type <= 0) return;
    for (int s = 0; s < 4; ++s) {
      int base_level;
      if (use_segment) {
        base_level = seg_filter[s];
        if (!absolute_delta) base_level += f_level;
      } else {
        base_level = f_level;
      }
      for (int i4 = 0; i4 <= 1; ++i4) {
        FInfo& info = fstrengths[s][i4];
        info = FInfo{0, 0, 0, 0};
        int level = base_level;
        if (use_lf_delta) {
          level += ref_lf_delta[0];
          if (i4) level += mode_lf_delta[0];
        }
        level = clampi(level, 0, 63);
        if (level > 0) {
          int ilevel = level;
          if (f_sharpness > 0) {
            ilevel >>= (f_sharpness > 4) ? 2 : 1;
            if (ilevel > 9 - f_sharpness) ilevel = 9 - f_sharpness;
          }
          if (ilevel < 1) ilevel = 1;
          info.f_ilevel = (uint8_t)ilevel;
          info.f_limit = (uint8_t)(2 * level + ilevel);
          info.hev_thresh = (level >= 40) ? 2 : (level >= 15) ? 1 : 0;
        } else {
          info.f_limit = 0;
        }
        info.f_inner = (uint8_t)i4;
      }
    }
  }

  // decode_mb.go:111 (getCoeffsInline)
  static int get_coeffs(BoolReader& br, const uint8_t bands[NUM_BANDS][NUM_CTX][NUM_PROBAS], int ctx, int dq0,
                        int dq1, int n, int16_t* out) {
    const uint8_t* p = bands[kBands[n]][ctx];
    for (; n < 16; ++n) {
      if (!br.get_bit(p[0])) return n;
      while (!br.get_bit(p[1])) {
        p = bands[kBands[++n]][0];
        if (n == 16) return 16;
      }
      const uint8_t(*p_ctx)[NUM_PROBAS] = bands[kBands[n + 1]];
      int v;
      if (!br.get_bit(p[2])) {
        v = 1;
        p = p_ctx[1];
      } else {
        if (!br.get_bit(p[3])) {
          if (!br.get_bit(p[4])) v = 2;
          else v = 3 + br.get_bit(p[5]);
        } else if (!br.get_bit(p[6])) {
          if (!br.get_bit(p[7])) {
            v = 5 + br.get_bit(159);
          } else {
            v = 7 + 2 * br.get_bit(165);
            v += br.get_bit(145);
          }
        } else {
          const int bit1 = br.get_bit(p[8]);
          const int bit0 = br.get_bit(p[9 + bit1]);
          const int cat = 2 * bit1 + bit0;
          v = 0;
          for (const uint8_t* tab = kCat3456[cat]; *tab; ++tab) v += v + br.get_bit(*tab);
          v += 3 + (8 << cat);
        }
        p = p_ctx[2];
      }
      out[kZigzag[n]] = (int16_t)(br.get_signed(v) * (n > 0 ? dq1 : dq0));
    }
    return 16;
  }
  static uint32_t nz_code_bits(uint32_t nz_coeffs, int nz, int dc_nz) {  // decode_mb.go:254
    nz_coeffs <<= 2;
    nz_coeffs |= (nz > 3) ? 3 : (nz > 1) ? 2 : dc_nz;
    return nz_coeffs;
  }

  bool decode_frame(const uint8_t* data, size_t len) {
    if (!parse_headers(data, len)) return false;
    y_stride = 16 * mb_w;
    uv_stride = 8 * mb_w;
    mbs.assign((size_t)mb_w * mb_h, MBData());
    finfo.assign((size_t)mb_w * mb_h, FInfo{0, 0, 0, 0});
    y.assign((size_t)y_stride * 16 * mb_h, 0);
    u.assign((size_t)uv_stride * 8 * mb_h, 0);
    v.assign((size_t)uv_stride * 8 * mb_h, 0);
    precompute_filter_strengths();
    if (!parse_all()) return false;
    reconstruct_all();
    if (filter_type > 0)
      for (int my = 0; my < mb_h; ++my)
        for (int mx = 0; mx < mb_w; ++mx) do_filter(mx, my);
    return true;
  }

  // Parsing of modes + residuals for the whole frame (decode_tree.go:35, decode_mb.go:267-430).
  bool parse_all() {
    std::vector<uint8_t> intra_t(4 * mb_w, B_DC_PRED);
    std::vector<uint8_t> top_nz(mb_w, 0), top_nz_dc(mb_w, 0);
    for (int my = 0; my < mb_h; ++my) {
      BoolReader& tbr = parts[my & num_parts_minus_one];
      uint8_t intra_l[4] = {B_DC_PRED, B_DC_PRED, B_DC_PRED, B_DC_PRED};
      // -- intra mode row (partition 0)
      for (int mx = 0; mx < mb_w; ++mx) {
        MBData& blk = mbs[(size_t)my * mb_w + mx];
        uint8_t* top = &intra_t[4 * mx];
        if (update_map) {
          blk.segment = !br.get_bit(proba.segments[0]) ? (uint8_t)br.get_bit(proba.segments[1])
                                                       : (uint8_t)(br.get_bit(proba.segments[2]) + 2);
        } else {
          blk.segment = 0;
        }
        if (use_skip_proba) blk.skip = (uint8_t)br.get_bit(skip_p);
        blk.is_i4x4 = !br.get_bit(145);
        if (!blk.is_i4x4) {
          const int ymode = br.get_bit(156) ? (br.get_bit(128) ? TM_PRED : H_PRED)
                                            : (br.get_bit(163) ? V_PRED : DC_PRED);
          blk.imodes[0] = (uint8_t)ymode;
          memset(top, ymode, 4);
          memset(intra_l, ymode, 4);
        } else {
          uint8_t* modes = blk.imodes;
          for (int yy = 0; yy < 4; ++yy) {
            int ymode = intra_l[yy];
            for (int xx = 0; xx < 4; ++xx) {
              const uint8_t* prob = &kBModesProba[(top[xx] * 10 + ymode) * 9];
              int i = kYModesIntra4[br.get_bit(prob[0])];
              while (i > 0) i = kYModesIntra4[2 * i + br.get_bit(prob[i])];
              ymode = -i;
              top[xx] = (uint8_t)ymode;
              modes[yy * 4 + xx] = (uint8_t)ymode;
            }
            intra_l[yy] = (uint8_t)ymode;
          }
        }
        blk.uvmode = !br.get_bit(142) ? DC_PRED : !br.get_bit(114) ? V_PRED : br.get_bit(183) ? TM_PRED : H_PRED;
      }
      if (br.eof) return fail("vp8: premature end of data");
      // -- residuals (token partition)
      uint8_t left_nz = 0, left_nz_dc = 0;
      for (int mx = 0; mx < mb_w; ++mx) {
        MBData& blk = mbs[(size_t)my * mb_w + mx];
        const bool skip = use_skip_proba ? blk.skip : false;
        if (!skip) {
          parse_residuals(blk, tbr, top_nz[mx], left_nz, top_nz_dc[mx], left_nz_dc);
        } else {
          left_nz = top_nz[mx] = 0;
          if (!blk.is_i4x4) left_nz_dc = top_nz_dc[mx] = 0;
          blk.non_zero_y = blk.non_zero_uv = 0;
        }
        if (filter_type > 0) {
          FInfo f = fstrengths[blk.segment & 3][blk.is_i4x4];
          bool eff_skip = skip;
          if (libwebp_inner_rule && !skip) eff_skip = !(blk.non_zero_y | blk.non_zero_uv);
          f.f_inner = f.f_inner || !eff_skip;
          finfo[(size_t)my * mb_w + mx] = f;
        }
        if (tbr.eof) return fail("vp8: premature end of data");
      }
    }
    return true;
  }

  void parse_residuals(MBData& blk, BoolReader& tbr, uint8_t& mb_nz, uint8_t& left_nz, uint8_t& mb_nz_dc,
                       uint8_t& left_nz_dc) {  // decode_mb.go:313
    const QuantMatrix& q = dqm[blk.segment & 3];
    int16_t* dst = blk.coeffs;
    memset(dst, 0, sizeof(blk.coeffs));
    uint32_t non_zero_y = 0, non_zero_uv = 0;
    int first;
    int ac_type;
    if (!blk.is_i4x4) {
      int16_t dc[16] = {0};
      const int ctx = mb_nz_dc + left_nz_dc;
      const int nz = get_coeffs(tbr, proba.bands[1], ctx, q.y2[0], q.y2[1], 0, dc);
      mb_nz_dc = left_nz_dc = (nz > 0);
      if (nz > 1) {
        transform_wht(dc, dst);
      } else {
        const int16_t dc0 = (int16_t)((dc[0] + 3) >> 3);
        for (int i = 0; i < 256; i += 16) dst[i] = dc0;
      }
      first = 1;
      ac_type = 0;
    } else {
      first = 0;
      ac_type = 3;
    }
    uint8_t tnz = mb_nz & 0x0f, lnz = left_nz & 0x0f;
    for (int yy = 0; yy < 4; ++yy) {
      uint8_t l = lnz & 1;
      uint32_t nz_coeffs = 0;
      for (int xx = 0; xx < 4; ++xx) {
        const int ctx = l + (tnz & 1);
        const int nz = get_coeffs(tbr, proba.bands[ac_type], ctx, q.y1[0], q.y1[1], first, dst);
        l = (nz > first);
        tnz = (uint8_t)((tnz >> 1) | (l << 7));
        nz_coeffs = nz_code_bits(nz_coeffs, nz, dst[0] != 0);
        dst += 16;
      }
      tnz >>= 4;
      lnz = (uint8_t)((lnz >> 1) | (l << 7));
      non_zero_y = (non_zero_y << 8) | nz_coeffs;
    }
    uint8_t out_t_nz = tnz, out_l_nz = lnz >> 4;
    for (int ch = 0; ch < 4; ch += 2) {
      uint32_t nz_coeffs = 0;
      tnz = mb_nz >> (4 + ch);
      lnz = left_nz >> (4 + ch);
      for (int yy = 0; yy < 2; ++yy) {
        uint8_t l = lnz & 1;
        for (int xx = 0; xx < 2; ++xx) {
          const int ctx = l + (tnz & 1);
          const int nz = get_coeffs(tbr, proba.bands[2], ctx, q.uv[0], q.uv[1], 0, dst);
          l = (nz > 0);
          tnz = (uint8_t)((tnz >> 1) | (l << 3));
          nz_coeffs = nz_code_bits(nz_coeffs, nz, dst[0] != 0);
          dst += 16;
        }
        tnz >>= 2;
        lnz = (uint8_t)((lnz >> 1) | (l << 5));
      }
      non_zero_uv |= nz_coeffs << (4 * ch);
      out_t_nz |= (uint8_t)((tnz << 4) << ch);
      out_l_nz |= (uint8_t)((lnz & 0xf0) << ch);
    }
    mb_nz = out_t_nz;
    left_nz = out_l_nz;
    blk.non_zero_y = non_zero_y;
    blk.non_zero_uv = non_zero_uv;
  }

  static int check_mode(int mx, int my, int mode) {  // decode_frame.go:6
    if (mode == B_DC_PRED) {
      if (mx == 0) return my == 0 ? B_DC_PRED_NOTOPLEFT : B_DC_PRED_NOLEFT;
      if (my == 0) return B_DC_PRED_NOTOP;
    }
    return mode;
  }
  static void do_transform(uint32_t bits, const int16_t* src, uint8_t* dst) {  // decode_frame.go:22
    switch (bits >> 30) {
      case 3: transform_one(src, dst); break;
      case 2: transform_ac3(src, dst); break;
      case 1: transform_dc(src, dst); break;
      default: break;
    }
  }
  static void do_uv_transform(uint32_t bits, const int16_t* src, uint8_t* dst) {  // decode_frame.go:46
    if (bits & 0xff) {
      if (bits & 0xaa) {
        transform_one(src, dst);
        transform_one(src + 16, dst + 4);
        transform_one(src + 32, dst + 4 * BPS);
        transform_one(src + 48, dst + 4 * BPS + 4);
      } else {
        if (src[0]) transform_dc(src, dst);
        if (src[16]) transform_dc(src + 16, dst + 4);
        if (src[32]) transform_dc(src + 32, dst + 4 * BPS);
        if (src[48]) transform_dc(src + 48, dst + 4 * BPS + 4);
      }
    }
  }

  // decode_frame.go:83-217 over the whole frame.  Prediction uses UNFILTERED neighbours,
  // so all rows are reconstructed before any filtering (equivalent to the row-interleaved order).
  void reconstruct_all() {
    uint8_t buf[YUV_SIZE];
    struct Top { uint8_t y[16], u[8], v[8]; };
    std::vector<Top> yuv_t(mb_w);
    for (int my = 0; my < mb_h; ++my) {
      memset(buf, 0, sizeof(buf));
      for (int j = 0; j < 16; ++j) buf[Y_OFF + j * BPS - 1] = 129;
      for (int j = 0; j < 8; ++j) buf[U_OFF + j * BPS - 1] = buf[V_OFF + j * BPS - 1] = 129;
      if (my > 0) {
        buf[Y_OFF - 1 - BPS] = buf[U_OFF - 1 - BPS] = buf[V_OFF - 1 - BPS] = 129;
      } else {
        memset(buf + Y_OFF - BPS - 1, 127, 16 + 4 + 1);
        memset(buf + U_OFF - BPS - 1, 127, 8 + 1);
        memset(buf + V_OFF - BPS - 1, 127, 8 + 1);
      }
      for (int mx = 0; mx < mb_w; ++mx) {
        const MBData& blk = mbs[(size_t)my * mb_w + mx];
        if (mx > 0) {
          for (int j = -1; j < 16; ++j) memcpy(buf + Y_OFF + j * BPS - 4, buf + Y_OFF + j * BPS + 12, 4);
          for (int j = -1; j < 8; ++j) {
            memcpy(buf + U_OFF + j * BPS - 4, buf + U_OFF + j * BPS + 4, 4);
            memcpy(buf + V_OFF + j * BPS - 4, buf + V_OFF + j * BPS + 4, 4);
          }
        }
        Top& top = yuv_t[mx];
        uint32_t bits = blk.non_zero_y;
        if (my > 0) {
          memcpy(buf + Y_OFF - BPS, top.y, 16);
          memcpy(buf + U_OFF - BPS, top.u, 8);
          memcpy(buf + V_OFF - BPS, top.v, 8);
        }
        if (blk.is_i4x4) {
          uint8_t* top_right = buf + Y_OFF - BPS + 16;
          if (my > 0) {
            if (mx >= mb_w - 1) memset(top_right, top.y[15], 4);
            else memcpy(top_right, yuv_t[mx + 1].y, 4);
          }
          for (int r = 1; r <= 3; ++r) memcpy(top_right + r * 4 * BPS, top_right, 4);
          for (int n = 0; n < 16; ++n, bits <<= 2) {
            const int off = Y_OFF + (n & 3) * 4 + (n >> 2) * 4 * BPS;
            pred_luma4(blk.imodes[n], buf, off);
            do_transform(bits, blk.coeffs + n * 16, buf + off);
          }
        } else {
          pred_luma16(check_mode(mx, my, blk.imodes[0]), buf, Y_OFF);
          if (bits != 0)
            for (int n = 0; n < 16; ++n, bits <<= 2)
              do_transform(bits, blk.coeffs + n * 16, buf + Y_OFF + (n & 3) * 4 + (n >> 2) * 4 * BPS);
        }
        const uint32_t bits_uv = blk.non_zero_uv;
        const int pm = check_mode(mx, my, blk.uvmode);
        pred_chroma8(pm, buf, U_OFF);
        pred_chroma8(pm, buf, V_OFF);
        do_uv_transform(bits_uv >> 0, blk.coeffs + 16 * 16, buf + U_OFF);
        do_uv_transform(bits_uv >> 8, blk.coeffs + 20 * 16, buf + V_OFF);
        if (my < mb_h - 1) {
          memcpy(top.y, buf + Y_OFF + 15 * BPS, 16);
          memcpy(top.u, buf + U_OFF + 7 * BPS, 8);
          memcpy(top.v, buf + V_OFF + 7 * BPS, 8);
        }
        for (int j = 0; j < 16; ++j)
          memcpy(&y[(size_t)(my * 16 + j) * y_stride + mx * 16], buf + Y_OFF + j * BPS, 16);
        for (int j = 0; j < 8; ++j) {
          memcpy(&u[(size_t)(my * 8 + j) * uv_stride + mx * 8], buf + U_OFF + j * BPS, 8);
          memcpy(&v[(size_t)(my * 8 + j) * uv_stride + mx * 8], buf + V_OFF + j * BPS, 8);
        }
      }
    }
  }

  // decode_frame.go:293-342
  void do_filter(int mx, int my) {
    const FInfo& f = finfo[(size_t)my * mb_w + mx];
    const int limit = f.f_limit;
    if (limit == 0) return;
    const int ilevel = f.f_ilevel;
    const int ys = y_stride;
    uint8_t* yp = &y[(size_t)my * 16 * ys + mx * 16];
    if (filter_type == 1) {
      if (mx > 0) simple_filter(yp, 1, ys, 16, limit + 4);
      if (f.f_inner) for (int k = 1; k <= 3; ++k) simple_filter(yp + 4 * k, 1, ys, 16, limit);
      if (my > 0) simple_filter(yp, ys, 1, 16, limit + 4);
      if (f.f_inner) for (int k = 1; k <= 3; ++k) simple_filter(yp + 4 * k * ys, ys, 1, 16, limit);
    } else {
      const int us = uv_stride;
      uint8_t* up = &u[(size_t)my * 8 * us + mx * 8];
      uint8_t* vp = &v[(size_t)my * 8 * us + mx * 8];
      const int hev_t = f.hev_thresh;
      if (mx > 0) {
        filter_loop26(yp, 1, ys, 16, limit + 4, ilevel, hev_t);
        filter_loop26(up, 1, us, 8, limit + 4, ilevel, hev_t);
        filter_loop26(vp, 1, us, 8, limit + 4, ilevel, hev_t);
      }
      if (f.f_inner) {
        for (int k = 1; k <= 3; ++k) filter_loop24(yp + 4 * k, 1, ys, 16, limit, ilevel, hev_t);
        filter_loop24(up + 4, 1, us, 8, limit, ilevel, hev_t);
        filter_loop24(vp + 4, 1, us, 8, limit, ilevel, hev_t);
      }
      if (my > 0) {
        filter_loop26(yp, ys, 1, 16, limit + 4, ilevel, hev_t);
        filter_loop26(up, us, 1, 8, limit + 4, ilevel, hev_t);
        filter_loop26(vp, us, 1, 8, limit + 4, ilevel, hev_t);
      }
      if (f.f_inner) {
        for (int k = 1; k <= 3; ++k) filter_loop24(yp + 4 * k * ys, ys, 1, 16, limit, ilevel, hev_t);
        filter_loop24(up + 4 * us, us, 1, 8, limit, ilevel, hev_t);
        filter_loop24(vp + 4 * us, us, 1, 8, limit, ilevel, hev_t);
      }
    }
  }
};

// RIFF container: locate the "VP8 " chunk payload (internal/container; simple + VP8X layouts).
static inline bool find_vp8_chunk(const uint8_t* data, size_t len, const uint8_t** out, size_t* out_len) {
  if (len >= 12 && !memcmp(data, "RIFF", 4) && !memcmp(data + 8, "WEBP", 4)) {
    size_t pos = 12;
    while (pos + 8 <= len) {
      const uint32_t sz = data[pos + 4] | (data[pos + 5] << 8) | (data[pos + 6] << 16) | ((uint32_t)data[pos + 7] << 24);
      if (!memcmp(data + pos, "VP8 ", 4)) {
        if (pos + 8 + sz > len) return false;
        *out = data + pos + 8;
        *out_len = sz;
        return true;
      }
      pos += 8 + sz + (sz & 1);
    }
    return false;
  }
  *out = data;  // raw VP8 frame
  *out_len = len;
  return true;
}

}  // namespace orc
