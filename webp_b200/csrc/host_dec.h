// Host side of the decoder below the GPU stages: the part the reference keeps serial per image
// (boolean decoder, headers, intra modes, coefficient tokens, dequantisation, inverse WHT).  Produces the
// frame-sized per-macroblock arrays the reconstruction / loop-filter kernels consume.  Mirrors:
//   parseHeaders / segment / filter / partitions   internal/lossy/decode.go:245-438
//   ParseQuant                                     internal/lossy/decode_quant.go:27
//   parseProba / parseIntraModeRow                 internal/lossy/decode_tree.go:7,35
//   decodeMB / parseResiduals / getCoeffsInline    internal/lossy/decode_mb.go:111-430
//   precomputeFilterStrengths                      internal/lossy/decode_frame.go:220
//   BoolReader                                     internal/bitio/reader_bool.go:27-230
//   RIFF chunk walk                                internal/container (VP8 / VP8X simple layouts)
#pragma once
#include <stdint.h>
#include <string.h>
#include "host_enc.h"

namespace wgh {

// Layout must match wg::MBMeta (dec_kernels.cuh).
struct MBMetaH {
  uint32_t non_zero_y, non_zero_uv;
  uint8_t imodes[16];
  uint8_t is_i4, uvmode, skip, segment;
  uint8_t f_limit, f_ilevel, f_inner, hev_thresh;
};
static_assert(sizeof(MBMetaH) == 32, "MBMetaH layout");

// VP8 boolean decoder: 56-bit refills, byte-wise tail, zero-extension past the end (one virtual byte, then EOF).
struct BoolDec {
  const uint8_t* p = nullptr; const uint8_t* end = nullptr; const uint8_t* start = nullptr; size_t len = 0;
  uint64_t value = 0;
  uint32_t range = 254;
  int bits = -8;
  bool eof = false;
  void init(const uint8_t* d, size_t n) { p = d; end = d + n; start = d; len = n; value = 0; range = 254; bits = -8; eof = false; refill(); }
  inline void refill() {
    if (end - p >= 8) {
      uint64_t w;
      memcpy(&w, p, 8);
      w = __builtin_bswap64(w) >> 8;  // 7 bytes, big endian
      p += 7;
      value = (value << 56) | w;
      bits += 56;
    } else if (p < end) {
      value = (value << 8) | *p++;
      bits += 8;
    } else if (!eof) {
      value <<= 8;
      bits += 8;
      eof = true;
    } else {
      bits = 0;
    }
  }
  inline int get(int prob) {
    uint32_t r = range;
    if (bits < 0) refill();
    const int pos = bits;
    const uint32_t split = (r * (uint32_t)prob) >> 8;
    const uint32_t v = (uint32_t)(value >> pos);
    int bit;
    if (v > split) { r -= split; value -= (uint64_t)(split + 1) << pos; bit = 1; }
    else { r = split + 1; bit = 0; }
    const int shift = 7 ^ (31 - __builtin_clz(r));
    r <<= shift;
    bits -= shift;
    range = r - 1;
    return bit;
  }
  inline uint32_t value_bits(int n) { uint32_t v = 0; while (n-- > 0) v |= (uint32_t)get(0x80) << n; return v; }
  inline int signed_bits(int n) { const int v = (int)value_bits(n); return get(0x80) ? -v : v; }
};

// What the device macroblock parser (dec_parse.cuh, wg::DecHeader -- same layout) needs from the frame headers.
struct DecHeaderH {
  unsigned long long br_value;
  uint32_t br_range;
  int32_t br_bits;
  uint32_t br_pos, br_end;
  uint32_t part_off[8], part_len[8];
  unsigned long long stream_off;
  int32_t dq[4][6];
  uint8_t fs[4][2][4];
  uint8_t proba[1056];
  uint8_t seg_proba[3];
  uint8_t br_eof, update_map, use_skip, skip_p, last_part, filter_type;
  uint8_t pad[3];
};

struct DecFrame {
  int width = 0, height = 0, mb_w = 0, mb_h = 0;
  int filter_type = 0;
  const char* err = nullptr;
};

// Locate the VP8 payload inside a RIFF/WebP file (or accept a raw VP8 frame).  1 = found, 0 = no VP8 chunk, -1 = the file
// uses a feature outside the lossy path: an ALPH chunk or the VP8X alpha flag (webp.Decode would return NRGBA with the real
// alpha, webp.go:323-350 -- decoding the VP8 chunk alone would silently hand back an opaque picture), or animation.
static inline int find_vp8_ex(const uint8_t* d, size_t n, const uint8_t** out, size_t* out_n) {
  if (n >= 12 && !memcmp(d, "RIFF", 4) && !memcmp(d + 8, "WEBP", 4)) {
    size_t pos = 12;
    while (pos + 8 <= n) {
      const uint32_t sz = d[pos + 4] | (d[pos + 5] << 8) | (d[pos + 6] << 16) | ((uint32_t)d[pos + 7] << 24);
      if (!memcmp(d + pos, "VP8X", 4) && pos + 9 <= n && (d[pos + 8] & (0x10 | 0x02))) return -1;  // alpha / animation flags
      if (!memcmp(d + pos, "ALPH", 4) || !memcmp(d + pos, "ANIM", 4) || !memcmp(d + pos, "ANMF", 4)) return -1;
      if (!memcmp(d + pos, "VP8 ", 4)) {
        if (pos + 8 + (size_t)sz > n) return 0;
        *out = d + pos + 8; *out_n = sz;
        return 1;
      }
      pos += 8 + (size_t)sz + (sz & 1);
    }
    return 0;
  }
  *out = d; *out_n = n;
  return 1;
}
static inline bool find_vp8(const uint8_t* d, size_t n, const uint8_t** out, size_t* out_n) { return find_vp8_ex(d, n, out, out_n) == 1; }
static const char* const kErrOutsideLossyPath = "webp: ALPH chunk / alpha or animation flag: outside the GPU lossy path";
// Frame tag + picture header only (webp.DecodeConfig path).
static inline bool peek_dims(const uint8_t* d, size_t n, int* w, int* h, const char** err) {
  if (n < 10) { *err = "vp8: truncated header"; return false; }
  const uint32_t tag = d[0] | (d[1] << 8) | (d[2] << 16);
  if (tag & 1) { *err = "vp8: not a keyframe"; return false; }
  if (((tag >> 1) & 7) > 3) { *err = "vp8: bad profile"; return false; }
  if (!((tag >> 4) & 1)) { *err = "vp8: frame not displayable"; return false; }
  if (d[3] != 0x9d || d[4] != 0x01 || d[5] != 0x2a) { *err = "vp8: bad signature"; return false; }
  *w = (d[6] | (d[7] << 8)) & 0x3fff;
  *h = (d[8] | (d[9] << 8)) & 0x3fff;
  if (*w == 0 || *h == 0) { *err = "vp8: zero dimensions"; return false; }
  return true;
}

// Token reader for one 4x4 block (getCoeffsInline, decode_mb.go:111): writes dequantised values in raster order.
static inline int read_block(BoolDec& br, const uint8_t (*bands)[3][11], int ctx, int dq_dc, int dq_ac, int n, int16_t* out) {
  const uint8_t* p = bands[kBands[n]][ctx];
  for (; n < 16; ++n) {
    if (!br.get(p[0])) return n;
    while (!br.get(p[1])) {
      p = bands[kBands[++n]][0];
      if (n == 16) return 16;
    }
    const uint8_t (*next)[11] = bands[kBands[n + 1]];
    int v;
    if (!br.get(p[2])) { v = 1; p = next[1]; }
    else {
      if (!br.get(p[3])) { v = !br.get(p[4]) ? 2 : 3 + br.get(p[5]); }
      else if (!br.get(p[6])) {
        if (!br.get(p[7])) v = 5 + br.get(159);
        else { v = 7 + 2 * br.get(165); v += br.get(145); }
      } else {
        const int b1 = br.get(p[8]), b0 = br.get(p[9 + b1]), cat = 2 * b1 + b0;
        v = 0;
        for (const uint8_t* t = kCats[cat]; *t; ++t) v += v + br.get(*t);
        v += 3 + (8 << cat);
      }
      p = next[2];
    }
    if (br.get(0x80)) v = -v;
    out[kZigzag[n]] = (int16_t)(v * (n > 0 ? dq_ac : dq_dc));
  }
  return 16;
}
static inline uint32_t push_code(uint32_t acc, int nz, int dc_nz) { return (acc << 2) | (uint32_t)(nz > 3 ? 3 : nz > 1 ? 2 : dc_nz); }

// inverse WHT producing the sixteen block DCs at stride 16 (transformWHT, internal/dsp/transforms.go:223)
static inline void inverse_wht(const int16_t* in, int16_t* out) {
  int t[16];
  for (int i = 0; i < 4; ++i) {
    const int a0 = in[i] + in[12 + i], a1 = in[4 + i] + in[8 + i], a2 = in[4 + i] - in[8 + i], a3 = in[i] - in[12 + i];
    t[i] = a0 + a1; t[8 + i] = a0 - a1; t[4 + i] = a3 + a2; t[12 + i] = a3 - a2;
  }
  for (int i = 0; i < 4; ++i) {
    const int dc = t[4 * i] + 3;
    const int a0 = dc + t[4 * i + 3], a1 = t[4 * i + 1] + t[4 * i + 2], a2 = t[4 * i + 1] - t[4 * i + 2], a3 = dc - t[4 * i + 3];
    int16_t* o = out + 64 * i;
    o[0] = (int16_t)((a0 + a1) >> 3); o[16] = (int16_t)((a3 + a2) >> 3); o[32] = (int16_t)((a0 - a1) >> 3); o[48] = (int16_t)((a3 - a2) >> 3);
  }
}

// Parse one VP8 key frame completely.  coeffs [nmb][384], meta [nmb] (caller-provided, e.g. pinned staging).
// When coeffs == nullptr only the headers are parsed (dimensions / filter type).
// With dev_hdr != nullptr only the headers are parsed here and the per-macroblock part is left to the device parser, which
// continues from the partition-0 decoder state stored in *dev_hdr (BoolDec refills 7 bytes at a time, so the state is
// rewound to byte granularity: value keeps `bits` + 8 significant bits, the rest is re-read on the device).
static inline bool parse_frame(const uint8_t* data, size_t len, DecFrame* F, int16_t* coeffs, MBMetaH* meta, int expect_mb_w,
                               int expect_mb_h, DecHeaderH* dev_hdr = nullptr) {
  if (!peek_dims(data, len, &F->width, &F->height, &F->err)) return false;
  const uint32_t part0_len = (data[0] | (data[1] << 8) | (data[2] << 16)) >> 5;
  const uint8_t* buf = data + 10;
  const size_t n = len - 10;
  F->mb_w = (F->width + 15) >> 4;
  F->mb_h = (F->height + 15) >> 4;
  if (part0_len > n) { F->err = "vp8: bad partition length"; return false; }
  BoolDec br;
  br.init(buf, part0_len);
  const uint8_t* tok = buf + part0_len;
  const size_t tok_len = n - part0_len;
  br.get(0x80);  // colourspace
  br.get(0x80);  // clamping type
  // segment header (decode.go:332)
  uint8_t seg_proba[3] = {255, 255, 255};
  int8_t seg_q[4] = {0, 0, 0, 0}, seg_f[4] = {0, 0, 0, 0};
  bool update_map = false, absolute = true;
  const bool use_seg = br.get(0x80);
  if (use_seg) {
    update_map = br.get(0x80);
    if (br.get(0x80)) {
      absolute = br.get(0x80);
      for (int s = 0; s < 4; ++s) seg_q[s] = br.get(0x80) ? (int8_t)br.signed_bits(7) : 0;
      for (int s = 0; s < 4; ++s) seg_f[s] = br.get(0x80) ? (int8_t)br.signed_bits(6) : 0;
    }
    if (update_map)
      for (int s = 0; s < 3; ++s) seg_proba[s] = br.get(0x80) ? (uint8_t)br.value_bits(8) : 255;
  }
  if (br.eof) { F->err = "vp8: premature EOF in segment header"; return false; }
  // filter header (decode.go:376)
  const bool f_simple = br.get(0x80);
  const int f_level = br.value_bits(6), f_sharp = br.value_bits(3);
  const bool use_lf_delta = br.get(0x80);
  int ref_delta[4] = {0, 0, 0, 0}, mode_delta[4] = {0, 0, 0, 0};
  if (use_lf_delta && br.get(0x80)) {
    for (int i = 0; i < 4; ++i) if (br.get(0x80)) ref_delta[i] = br.signed_bits(6);
    for (int i = 0; i < 4; ++i) if (br.get(0x80)) mode_delta[i] = br.signed_bits(6);
  }
  F->filter_type = f_level == 0 ? 0 : (f_simple ? 1 : 2);
  // token partitions (decode.go:409)
  const int last = (1 << br.value_bits(2)) - 1;
  if (tok_len < (size_t)3 * last) { F->err = "vp8: not enough data for partition sizes"; return false; }
  BoolDec parts[8];
  {
    const uint8_t* start = tok + 3 * last;
    size_t left = tok_len - 3 * last;
    for (int p = 0; p < last; ++p) {
      const size_t sz = tok[3 * p] | (tok[3 * p + 1] << 8) | (tok[3 * p + 2] << 16);
      if (sz > left) { F->err = "vp8: partition size exceeds remaining data"; return false; }
      parts[p].init(start, sz);
      start += sz;
      left -= sz;
    }
    parts[last].init(start, left);
  }
  // quantisers (decode_quant.go:27)
  int dq[4][6];  // y1 dc/ac, y2 dc/ac, uv dc/ac
  {
    const int base = br.value_bits(7);
    const int d_y1dc = br.get(0x80) ? br.signed_bits(4) : 0, d_y2dc = br.get(0x80) ? br.signed_bits(4) : 0;
    const int d_y2ac = br.get(0x80) ? br.signed_bits(4) : 0, d_uvdc = br.get(0x80) ? br.signed_bits(4) : 0;
    const int d_uvac = br.get(0x80) ? br.signed_bits(4) : 0;
    for (int i = 0; i < 4; ++i) {
      int q;
      if (use_seg) { q = seg_q[i]; if (!absolute) q += base; }
      else if (i > 0) { memcpy(dq[i], dq[0], sizeof(dq[0])); continue; }
      else q = base;
      dq[i][0] = kDcTable[clampi(q + d_y1dc, 0, 127)];
      dq[i][1] = kAcTable[clampi(q, 0, 127)];
      dq[i][2] = kDcTable[clampi(q + d_y2dc, 0, 127)] * 2;
      dq[i][3] = (kAcTable[clampi(q + d_y2ac, 0, 127)] * 101581) >> 16;
      if (dq[i][3] < 8) dq[i][3] = 8;
      dq[i][4] = kDcTable[clampi(q + d_uvdc, 0, 117)];
      dq[i][5] = kAcTable[clampi(q + d_uvac, 0, 127)];
    }
  }
  br.get(0x80);  // refresh-entropy flag, unused for key frames
  uint8_t proba[4][8][3][11];
  for (int i = 0; i < 4 * 8 * 3 * 11; ++i)
    (&proba[0][0][0][0])[i] = br.get(kCoeffsUpdateProba[i]) ? (uint8_t)br.value_bits(8) : kCoeffsProba0[i];
  const bool use_skip = br.get(0x80);
  const int skip_p = use_skip ? (int)br.value_bits(8) : 0;
  if (!coeffs && !dev_hdr) return true;
  if (F->mb_w != expect_mb_w || F->mb_h != expect_mb_h) { F->err = "batch decode needs identical dimensions"; return false; }
  // per-(segment, i4) filter strengths (decode_frame.go:220)
  uint8_t fs[4][2][4];  // limit, ilevel, inner, hev
  memset(fs, 0, sizeof(fs));
  if (F->filter_type > 0)
    for (int s = 0; s < 4; ++s) {
      int base_level = use_seg ? seg_f[s] + (absolute ? 0 : f_level) : f_level;
      for (int i4 = 0; i4 <= 1; ++i4) {
        int level = base_level;
        if (use_lf_delta) { level += ref_delta[0]; if (i4) level += mode_delta[0]; }
        level = clampi(level, 0, 63);
        if (level > 0) {
          int il = level;
          if (f_sharp > 0) { il >>= (f_sharp > 4) ? 2 : 1; if (il > 9 - f_sharp) il = 9 - f_sharp; }
          if (il < 1) il = 1;
          fs[s][i4][0] = (uint8_t)(2 * level + il);
          fs[s][i4][1] = (uint8_t)il;
          fs[s][i4][3] = level >= 40 ? 2 : level >= 15 ? 1 : 0;
        }
        fs[s][i4][2] = (uint8_t)i4;
      }
    }
  if (dev_hdr) {
    DecHeaderH& D = *dev_hdr;
    memset(&D, 0, sizeof(D));
    // hand the decoder over at a byte boundary: drop the whole bytes still unread in `value`
    int bits = br.bits;
    uint64_t value = br.value;
    const uint8_t* p = br.p;
    if (!br.eof) {
      while (bits >= 8) { value >>= 8; bits -= 8; --p; }
    }
    D.br_value = value; D.br_range = br.range; D.br_bits = bits; D.br_eof = br.eof;
    D.br_pos = (uint32_t)(p - data); D.br_end = (uint32_t)(br.end - data);
    for (int p2 = 0; p2 <= last; ++p2) { D.part_off[p2] = (uint32_t)(parts[p2].start - data); D.part_len[p2] = (uint32_t)parts[p2].len; }
    memcpy(D.dq, dq, sizeof(dq));
    memcpy(D.fs, fs, sizeof(fs));
    memcpy(D.proba, proba, sizeof(proba));
    memcpy(D.seg_proba, seg_proba, 3);
    D.update_map = update_map; D.use_skip = use_skip; D.skip_p = (uint8_t)skip_p; D.last_part = (uint8_t)last; D.filter_type = (uint8_t)F->filter_type;
    return true;
  }
  // macroblocks: intra modes from partition 0, residuals from partition (row & last)
  const int mb_w = F->mb_w, mb_h = F->mb_h;
  std::vector<uint8_t> top_modes((size_t)4 * mb_w, 0), top_nz(mb_w, 0), top_dc(mb_w, 0);
  for (int my = 0; my < mb_h; ++my) {
    uint8_t left_modes[4] = {0, 0, 0, 0};
    MBMetaH* row = meta + (size_t)my * mb_w;
    for (int mx = 0; mx < mb_w; ++mx) {  // parseIntraModeRow (decode_tree.go:35)
      MBMetaH& m = row[mx];
      uint8_t* top = &top_modes[4 * mx];
      m.segment = update_map ? (uint8_t)(!br.get(seg_proba[0]) ? br.get(seg_proba[1]) : br.get(seg_proba[2]) + 2) : 0;
      m.skip = use_skip ? (uint8_t)br.get(skip_p) : 0;
      m.is_i4 = !br.get(145);
      memset(m.imodes, 0, 16);
      if (!m.is_i4) {
        const int ym = br.get(156) ? (br.get(128) ? 1 : 3) : (br.get(163) ? 2 : 0);
        m.imodes[0] = (uint8_t)ym;
        memset(top, ym, 4);
        memset(left_modes, ym, 4);
      } else {
        for (int y = 0; y < 4; ++y) {
          int ym = left_modes[y];
          for (int x = 0; x < 4; ++x) {
            const uint8_t* prob = &kBModesProba[(top[x] * 10 + ym) * 9];
            int i = kI4Tree[br.get(prob[0])];
            while (i > 0) i = kI4Tree[2 * i + br.get(prob[i])];
            ym = -i;
            top[x] = (uint8_t)ym;
            m.imodes[4 * y + x] = (uint8_t)ym;
          }
          left_modes[y] = (uint8_t)ym;
        }
      }
      m.uvmode = !br.get(142) ? 0 : !br.get(114) ? 2 : br.get(183) ? 1 : 3;
    }
    if (br.eof) { F->err = "vp8: premature end of data"; return false; }
    BoolDec& tb = parts[my & last];
    uint8_t left_nz = 0, left_dc = 0;
    for (int mx = 0; mx < mb_w; ++mx) {
      MBMetaH& m = row[mx];
      int16_t* dst = coeffs + ((size_t)my * mb_w + mx) * 384;
      memset(dst, 0, 384 * sizeof(int16_t));
      const bool skip = use_skip && m.skip;
      if (skip) {
        left_nz = top_nz[mx] = 0;
        if (!m.is_i4) left_dc = top_dc[mx] = 0;
        m.non_zero_y = m.non_zero_uv = 0;
      } else {  // parseResiduals (decode_mb.go:313)
        const int* q = dq[m.segment & 3];
        const uint8_t tnz_in = top_nz[mx], lnz_in = left_nz;
        int first = 0, type = 3;
        if (!m.is_i4) {
          int16_t dc[16] = {0};
          const int nz = read_block(tb, proba[1], top_dc[mx] + left_dc, q[2], q[3], 0, dc);
          top_dc[mx] = left_dc = (nz > 0);
          if (nz > 1) inverse_wht(dc, dst);
          else { const int16_t d0 = (int16_t)((dc[0] + 3) >> 3); for (int i = 0; i < 256; i += 16) dst[i] = d0; }
          first = 1;
          type = 0;
        }
        uint8_t tnz = tnz_in & 0x0f, lnz = lnz_in & 0x0f;
        uint32_t nzy = 0, nzuv = 0;
        int16_t* d = dst;
        for (int y = 0; y < 4; ++y) {
          uint8_t l = lnz & 1;
          uint32_t acc = 0;
          for (int x = 0; x < 4; ++x) {
            const int nz = read_block(tb, proba[type], l + (tnz & 1), q[0], q[1], first, d);
            l = nz > first;
            tnz = (uint8_t)((tnz >> 1) | (l << 7));
            acc = push_code(acc, nz, d[0] != 0);
            d += 16;
          }
          tnz >>= 4;
          lnz = (uint8_t)((lnz >> 1) | (l << 7));
          nzy = (nzy << 8) | acc;
        }
        uint8_t out_t = tnz, out_l = lnz >> 4;
        for (int ch = 0; ch < 4; ch += 2) {
          uint32_t acc = 0;
          tnz = tnz_in >> (4 + ch);
          lnz = lnz_in >> (4 + ch);
          for (int y = 0; y < 2; ++y) {
            uint8_t l = lnz & 1;
            for (int x = 0; x < 2; ++x) {
              const int nz = read_block(tb, proba[2], l + (tnz & 1), q[4], q[5], 0, d);
              l = nz > 0;
              tnz = (uint8_t)((tnz >> 1) | (l << 3));
              acc = push_code(acc, nz, d[0] != 0);
              d += 16;
            }
            tnz >>= 2;
            lnz = (uint8_t)((lnz >> 1) | (l << 5));
          }
          nzuv |= acc << (4 * ch);
          out_t |= (uint8_t)((tnz << 4) << ch);
          out_l |= (uint8_t)((lnz & 0xf0) << ch);
        }
        top_nz[mx] = out_t;
        left_nz = out_l;
        m.non_zero_y = nzy;
        m.non_zero_uv = nzuv;
      }
      const uint8_t* f = fs[m.segment & 3][m.is_i4];
      m.f_limit = f[0]; m.f_ilevel = f[1]; m.hev_thresh = f[3];
      m.f_inner = (uint8_t)(f[2] || !skip);  // FInner |= !skip (decode_mb.go:291)
      if (tb.eof) { F->err = "vp8: premature end of data"; return false; }
    }
  }
  return true;
}

}  // namespace wgh
