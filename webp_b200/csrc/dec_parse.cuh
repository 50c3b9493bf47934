// Decoder macroblock parser on the device (sm_100a): intra modes from partition 0 and coefficient tokens from the token
// partitions, i.e. parseIntraModeRow (internal/lossy/decode_tree.go:35) and parseResiduals / getCoeffs
// (decode_mb.go:111-313) over the boolean decoder of internal/bitio/reader_bool.go -- SURVEY 8(f) rank 2, the mirror of the
// encoder's boolean coder (whose range recurrence depends on the tokens only and so runs chunk-parallel, boolcode_par.cuh;
// the decoder's depends on the bits it has yet to decode).  The frame HEADERS stay on the host (host_dec.h::parse_frame: a few hundred bits per image); the host
// hands over the partition-0 decoder state right after them plus the tables they define (DecHeader).
//
// Boolean decoding is a serial chain per partition with data-dependent control flow at every bit, so there is nothing for
// SIMT to share inside an image: one single-thread block per image walks the macroblocks in raster order, and the batch
// supplies the parallelism (256 images = 256 warps = 3 % of the warp slots; several batches' parsers overlap on the GPU).  What it
// buys is the host: 768 B of dequantised coefficients per macroblock no longer cross PCIe (1.26 GB per 256-image batch ->
// 46 MB of compressed bytes), and 32 host cores no longer parse for 8 GPUs.
//
// Output layout is exactly what host_dec.h::parse_frame writes (coefficients dequantised, WHT applied, [nmb][384]; MBMeta).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dec_kernels.cuh"

namespace wg {

// Filled by host_dec.h::parse_frame (same layout as wgh::DecHeaderH).
struct DecHeader {
  unsigned long long br_value;  // partition-0 decoder state after the frame headers
  uint32_t br_range;
  int32_t br_bits;
  uint32_t br_pos, br_end;      // next byte / end of partition 0, offsets into this image's VP8 frame
  uint32_t part_off[8], part_len[8];
  unsigned long long stream_off;  // offset of the VP8 frame in the packed stream buffer
  int32_t dq[4][6];             // per segment: y1 dc/ac, y2 dc/ac, uv dc/ac
  uint8_t fs[4][2][4];          // per (segment, is_i4): limit, ilevel, inner, hev_thresh
  uint8_t proba[1056];
  uint8_t seg_proba[3];
  uint8_t br_eof, update_map, use_skip, skip_p, last_part, filter_type;
  uint8_t pad[3];
};

__device__ __constant__ int8_t c_i4tree[18] = {0, 1, -1, 2, -2, 3, 4, 6, -3, 5, -4, -5, -6, 7, -7, 8, -8, -9};
__device__ __constant__ uint8_t c_dcat3[4] = {173, 148, 140, 0};
__device__ __constant__ uint8_t c_dcat4[5] = {176, 155, 140, 135, 0};
__device__ __constant__ uint8_t c_dcat5[6] = {180, 157, 141, 134, 130, 0};
__device__ __constant__ uint8_t c_dcat6[12] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0};

// VP8 boolean decoder (bitio/reader_bool.go), the device twin of wgh::BoolDec: a 32-bit window refilled 16 bits at a time
// (the 64-bit window of the host costs multi-instruction shifts at every bit here), byte-wise tail, one virtual zero byte
// past the end and then EOF -- the same bits and the same EOF point as the host's 56-bit refills.
struct DBoolDec {
  const uint8_t* p; const uint8_t* end;
  uint32_t value;
  uint32_t range;  // the reference's range - 1 form, normalised: 127 .. 254
  int bits;
  bool eof;
  __device__ __forceinline__ void init(const uint8_t* d, uint32_t n) { p = d; end = d + n; value = 0; range = 254; bits = -8; eof = false; refill(); }
  // adopt a state handed over by the host at byte granularity (bits < 8, value < 2^17)
  __device__ __forceinline__ void adopt(const uint8_t* cur, const uint8_t* e, unsigned long long v, uint32_t r, int b, bool f) {
    p = cur; end = e; range = r; eof = f;
    // the host may hand over with EOF already hit and up to 55 unread (zero) bits: keep the low window only
    while (b > 15) { v >>= 8; b -= 8; }
    value = (uint32_t)v; bits = b;
  }
  __device__ __forceinline__ void refill() {
    if (end - p >= 2) {
      value = (value << 16) | ((uint32_t)__ldg(p) << 8) | (uint32_t)__ldg(p + 1);
      p += 2;
      bits += 16;
    } else if (p < end) {
      value = (value << 8) | __ldg(p++);
      bits += 8;
    } else if (!eof) {
      value <<= 8;
      bits += 8;
      eof = true;
    } else {
      bits = 0;
    }
  }
  // One decoded bit; p24 = probability << 24.  The dependent chain per bit is range -> split -> compare -> new range ->
  // its leading zeros -> shift, and it is kept to six operations: split = (range * prob) >> 8 is the upper word of
  // range * p24 (one IMAD.HI instead of multiply + shift), and the renormalised (r << shift) - 1 is one funnel shift of
  // r - 1 with ones coming in from below (no separate decrement).
  __device__ __forceinline__ int get24(uint32_t p24) {
    if (bits < 0) refill();
    const int pos = bits;
    const uint32_t split = __umulhi(range, p24);
    const uint32_t v = value >> pos;
    const bool bit = v > split;
    const uint32_t r = bit ? range - split : split + 1;  // the new range + 1: what the shift is read from
    if (bit) value -= (split + 1) << pos;
    const uint32_t rm = r - 1;                            // the new range itself, before renormalisation (off the chain)
    const int shift = 7 ^ (31 - __clz(r));
    range = __funnelshift_l(0xffffffffu, rm, shift);
    bits -= shift;
    return bit;
  }
  __device__ __forceinline__ int get(int prob) { return get24((uint32_t)prob << 24); }
};

// getCoeffs (decode_mb.go:111): one block's tokens, dequantised into out[zigzag]; returns the position after the last
// coefficient read.  P = probabilities of this block's type in shared memory, rows [band][ctx] padded to 16 bytes so that
// a row is ONE 128-bit load into registers per coefficient instead of a dependent byte load per bit; *dc_nz = whether
// out[0] was written non-zero (as the int16 it is stored as).
#define WG_PB(row, i) __byte_perm((i) < 4 ? (row).x : (i) < 8 ? (row).y : (row).z, 0, ((((i) & 3)) << 12) | 0x444)  /* probability i << 24 */
__device__ __forceinline__ int dread_block(DBoolDec& br, const uint4* P, int ctx, int dq_dc, int dq_ac, int n, int16_t* out, bool* dc_nz) {
  uint4 p = P[c_bands[n] * 3 + ctx];
  for (; n < 16; ++n) {
    if (!br.get24(WG_PB(p, 0))) return n;
    while (!br.get24(WG_PB(p, 1))) {
      p = P[c_bands[++n] * 3 + 0];
      if (n == 16) return 16;
    }
    const uint4* next = P + c_bands[n + 1] * 3;
    const uint4 n1 = next[1], n2 = next[2];  // both candidates for the next row, issued before the level is known
    int v;
    if (!br.get24(WG_PB(p, 2))) { v = 1; p = n1; }
    else {
      if (!br.get24(WG_PB(p, 3))) { v = !br.get24(WG_PB(p, 4)) ? 2 : 3 + br.get24(WG_PB(p, 5)); }
      else if (!br.get24(WG_PB(p, 6))) {
        if (!br.get24(WG_PB(p, 7))) v = 5 + br.get(159);
        else { v = 7 + 2 * br.get(165); v += br.get(145); }
      } else {
        const int b1 = br.get24(WG_PB(p, 8)), b0 = br.get24(b1 ? WG_PB(p, 10) : WG_PB(p, 9)), cat = 2 * b1 + b0;
        const uint8_t* t = cat == 0 ? c_dcat3 : cat == 1 ? c_dcat4 : cat == 2 ? c_dcat5 : c_dcat6;
        v = 0;
        for (; *t; ++t) v += v + br.get(*t);
        v += 3 + (8 << cat);
      }
      p = n2;
    }
    if (br.get(0x80)) v = -v;
    const int16_t q = (int16_t)(v * (n > 0 ? dq_ac : dq_dc));
    if (n == 0) *dc_nz = q != 0;
    out[c_zigzag[n]] = q;
  }
  return 16;
}

// transformWHT (dsp/transforms.go:223): the sixteen block DCs at stride 16; returns which of them are non-zero
__device__ __forceinline__ uint32_t dinverse_wht(const int16_t* in, int16_t* out) {
  int t[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a0 = in[i] + in[12 + i], a1 = in[4 + i] + in[8 + i], a2 = in[4 + i] - in[8 + i], a3 = in[i] - in[12 + i];
    t[i] = a0 + a1; t[8 + i] = a0 - a1; t[4 + i] = a3 + a2; t[12 + i] = a3 - a2;
  }
  uint32_t mask = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int dc = t[4 * i] + 3;
    const int a0 = dc + t[4 * i + 3], a1 = t[4 * i + 1] + t[4 * i + 2], a2 = t[4 * i + 1] - t[4 * i + 2], a3 = dc - t[4 * i + 3];
    int16_t* o = out + 64 * i;
    const int16_t v0 = (int16_t)((a0 + a1) >> 3), v1 = (int16_t)((a3 + a2) >> 3), v2 = (int16_t)((a0 - a1) >> 3), v3 = (int16_t)((a3 - a2) >> 3);
    o[0] = v0; o[16] = v1; o[32] = v2; o[48] = v3;
    mask |= ((uint32_t)(v0 != 0) | ((uint32_t)(v1 != 0) << 1) | ((uint32_t)(v2 != 0) << 2) | ((uint32_t)(v3 != 0) << 3)) << (4 * i);
  }
  return mask;
}

struct DecParseParams {
  const uint8_t* streams;    // packed VP8 frames
  const DecHeader* hdr;      // [n]
  const uint8_t* bmodes;     // kBModesProba [10][10][9]
  int16_t* coeffs;           // [n][nmb][384]; written for macroblocks with a non-zero transform code only (nothing else is read)
  MBMeta* meta;              // [n][nmb]
  int* err;                  // [n] 0 ok, 1 premature end of data
  int n_images, mb_w, mb_h;
};

// Dynamic shared memory: the image's header (probabilities, quantisers, filter strengths), the intra-mode probabilities,
// the per-column contexts (4 top modes + NZ flags + DC flag per macroblock column) and the token-partition decoder states.
// ONE thread per block.  The chain has nothing for a second lane to do, and a block that cannot diverge is compiled without
// the convergence bookkeeping of a 32-lane warp: with __launch_bounds__(1) ptxas drops the BSSY / BSYNC pair around the refill
// test of every decoded bit and makes the bit's own branch a uniform one (87 of them in this kernel; 149 -> 62 BSSY), which
// is two to three instructions off a ~20-instruction dependent chain.  The thread also moves a coded macroblock out of shared
// memory itself: 48 independent 128-bit stores against the ~400 decoded bits of such a macroblock.
__global__ void __launch_bounds__(1) dec_parse_kernel(const DecParseParams P) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ __align__(16) int16_t s_co[384];  // the macroblock being parsed: zero between macroblocks
  uint4* s_co4 = reinterpret_cast<uint4*>(s_co);
  DecHeader* H = reinterpret_cast<DecHeader*>(s_dyn);
  uint4* s_prob = reinterpret_cast<uint4*>(s_dyn + ((sizeof(DecHeader) + 15) & ~(size_t)15));  // [4][8][3] rows of 11 (+5) bytes
  uint8_t* s_bmodes = reinterpret_cast<uint8_t*>(s_prob + 96);
  uint8_t* top_modes = s_bmodes + 912;
  uint8_t* top_nz = top_modes + 4 * P.mb_w;
  uint8_t* top_dc = top_nz + P.mb_w;
  DBoolDec* parts = reinterpret_cast<DBoolDec*>(s_dyn + ((((top_dc + P.mb_w) - s_dyn) + 15) & ~(size_t)15));
  const int img = blockIdx.x;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.hdr + img);
    uint32_t* dst = reinterpret_cast<uint32_t*>(H);
    for (int i = 0; i < (int)(sizeof(DecHeader) / 4); ++i) dst[i] = src[i];
    const uint32_t* bm = reinterpret_cast<const uint32_t*>(P.bmodes);  // 900 bytes, 4-byte aligned table
    for (int i = 0; i < 225; ++i) reinterpret_cast<uint32_t*>(s_bmodes)[i] = bm[i];
    for (int r = 0; r < 96; ++r) {
      uint8_t* d = reinterpret_cast<uint8_t*>(s_prob + r);
      for (int k = 0; k < 16; ++k) d[k] = k < 11 ? H->proba[r * 11 + k] : 0;
    }
    for (int i = 0; i < 6 * P.mb_w; ++i) top_modes[i] = 0;
  }
  for (int i = 0; i < 48; ++i) s_co4[i] = make_uint4(0, 0, 0, 0);
  const uint8_t* frame = P.streams + H->stream_off;
  DBoolDec br;
  br.adopt(frame + H->br_pos, frame + H->br_end, H->br_value, H->br_range, H->br_bits, H->br_eof != 0);
  const int last = H->last_part;
  for (int p = 0; p <= last; ++p) parts[p].init(frame + H->part_off[p], H->part_len[p]);
  const int mb_w = P.mb_w, mb_h = P.mb_h;
  const size_t nmb = (size_t)mb_w * mb_h;
  const bool update_map = H->update_map, use_skip = H->use_skip;
  const int skip_p = H->skip_p;
  for (int my = 0; my < mb_h; ++my) {
    uint8_t left_modes[4] = {0, 0, 0, 0};
    MBMeta* row = P.meta + (size_t)img * nmb + (size_t)my * mb_w;
    for (int mx = 0; mx < mb_w; ++mx) {  // parseIntraModeRow (decode_tree.go:35)
      MBMeta m;
      uint8_t* top = top_modes + 4 * mx;
      m.segment = update_map ? (uint8_t)(!br.get(H->seg_proba[0]) ? br.get(H->seg_proba[1]) : br.get(H->seg_proba[2]) + 2) : 0;
      m.skip = use_skip ? (uint8_t)br.get(skip_p) : 0;
      m.is_i4 = !br.get(145);
#pragma unroll
      for (int i = 0; i < 16; ++i) m.imodes[i] = 0;
      if (!m.is_i4) {
        const int ym = br.get(156) ? (br.get(128) ? 1 : 3) : (br.get(163) ? 2 : 0);
        m.imodes[0] = (uint8_t)ym;
        for (int i = 0; i < 4; ++i) { top[i] = (uint8_t)ym; left_modes[i] = (uint8_t)ym; }
      } else {
        for (int y = 0; y < 4; ++y) {
          int ym = left_modes[y];
          for (int x = 0; x < 4; ++x) {
            const uint8_t* prob = s_bmodes + (top[x] * 10 + ym) * 9;
            int i = c_i4tree[br.get(prob[0])];
            while (i > 0) i = c_i4tree[2 * i + br.get(prob[i])];
            ym = -i;
            top[x] = (uint8_t)ym;
            m.imodes[4 * y + x] = (uint8_t)ym;
          }
          left_modes[y] = (uint8_t)ym;
        }
      }
      m.uvmode = !br.get(142) ? 0 : !br.get(114) ? 2 : br.get(183) ? 1 : 3;
      m.non_zero_y = m.non_zero_uv = 0;
      m.f_limit = m.f_ilevel = m.f_inner = m.hev_thresh = 0;
      row[mx] = m;
    }
    if (br.eof) { P.err[img] = 1; return; }
    DBoolDec tb = parts[my & last];
    uint8_t left_nz = 0, left_dc = 0;
    for (int mx = 0; mx < mb_w; ++mx) {
      MBMeta& m = row[mx];
      const int is_i4 = m.is_i4, segment = m.segment & 3;
      int16_t* dst = s_co;
      const bool skip = use_skip && m.skip;
      uint32_t nzy = 0, nzuv = 0;
      if (skip) {
        left_nz = top_nz[mx] = 0;
        if (!is_i4) left_dc = top_dc[mx] = 0;
      } else {  // parseResiduals (decode_mb.go:313)
        const int* q = H->dq[segment];
        const uint8_t tnz_in = top_nz[mx], lnz_in = left_nz;
        int first = 0, type = 3;
        uint32_t wht_nz = 0;  // I16: which blocks got a non-zero DC from the WHT
        if (!is_i4) {
          int16_t dc[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) dc[i] = 0;
          bool unused;
          const int nz = dread_block(tb, s_prob + 1 * 24, top_dc[mx] + left_dc, q[2], q[3], 0, dc, &unused);
          top_dc[mx] = left_dc = (nz > 0);
          if (nz > 1) {
            wht_nz = dinverse_wht(dc, dst);
          } else {
            const int16_t d0 = (int16_t)((dc[0] + 3) >> 3);
            for (int i = 0; i < 256; i += 16) dst[i] = d0;
            wht_nz = d0 != 0 ? 0xffffu : 0u;
          }
          first = 1;
          type = 0;
        }
        uint8_t tnz = tnz_in & 0x0f, lnz = lnz_in & 0x0f;
        int16_t* d = dst;
        for (int y = 0; y < 4; ++y) {
          uint8_t l = lnz & 1;
          uint32_t acc = 0;
          for (int x = 0; x < 4; ++x) {
            bool dc_nz = (wht_nz >> (4 * y + x)) & 1u;
            const int nz = dread_block(tb, s_prob + type * 24, l + (tnz & 1), q[0], q[1], first, d, &dc_nz);
            l = nz > first;
            tnz = (uint8_t)((tnz >> 1) | (l << 7));
            acc = (acc << 2) | (uint32_t)(nz > 3 ? 3 : nz > 1 ? 2 : dc_nz);
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
              bool dc_nz = false;
              const int nz = dread_block(tb, s_prob + 2 * 24, l + (tnz & 1), q[4], q[5], 0, d, &dc_nz);
              l = nz > 0;
              tnz = (uint8_t)((tnz >> 1) | (l << 3));
              acc = (acc << 2) | (uint32_t)(nz > 3 ? 3 : nz > 1 ? 2 : dc_nz);
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
      }
      m.non_zero_y = nzy;
      m.non_zero_uv = nzuv;
      const uint8_t* f = H->fs[segment][is_i4];
      m.f_limit = f[0]; m.f_ilevel = f[1]; m.hev_thresh = f[3];
      m.f_inner = (uint8_t)(f[2] || !skip);  // FInner |= !skip (decode_mb.go:291)
      if (tb.eof) { P.err[img] = 1; return; }
      if (nzy | nzuv) {
        // the reconstruction reads the coefficients of blocks with a non-zero code only (recon_wave_kernel), so an all-zero
        // macroblock leaves nothing behind and the array needs no zero fill: 768 B per coded macroblock, written once
        uint4* dst4 = reinterpret_cast<uint4*>(P.coeffs + ((size_t)img * nmb + (size_t)my * mb_w + mx) * 384);
#pragma unroll 8
        for (int i = 0; i < 48; ++i) { dst4[i] = s_co4[i]; s_co4[i] = make_uint4(0, 0, 0, 0); }
      }
    }
    parts[my & last] = tb;
  }
}

inline size_t dec_parse_smem(int mb_w) {
  size_t s = (sizeof(DecHeader) + 15) & ~(size_t)15;
  s += 96 * 16 + 912 + 6 * (size_t)mb_w;
  s = (s + 15) & ~(size_t)15;
  return s + 8 * sizeof(DBoolDec);
}

}  // namespace wg
