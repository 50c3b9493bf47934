// Device-side token generation: after the mode search, the GPU turns the per-macroblock levels
// into the final (bit, probability) token stream of each image, so the host only runs the serial boolean coder over a
// flat array.  Restates, per macroblock and in bitstream order,
//   recordMBTokens / RecordCoeffs / recordLevelVP8     internal/lossy/encode_frame.go:647, encode_token.go:115-300
//   optimizeProba                                      internal/lossy/encode_proba.go:117-156
// Token = bit | prob << 8 (uint16), exactly the reference's Token{Bit, Prob} pair (encode_token.go:20).
#pragma once
#include "enc_kernels.cuh"

namespace wg {

struct TokenParams {
  const uint8_t* hdr;        // [n][nmb][48]
  const int16_t* coeffs;     // [n][nmb][400]
  const uint32_t* ctxw;      // [n][nmb] NZ context words written by the mode search
  const uint8_t* proba;      // [n][1056] final probabilities (emit only)
  uint32_t* mb_tokens;       // [n][nmb] token count per macroblock
  unsigned long long* mb_offset;  // [n][nmb] exclusive prefix within the image (raster order)
  unsigned long long* img_total;  // [n]
  const unsigned long long* img_base;  // [n] offset of each image in the compact token buffer (emit only)
  uint16_t* tokens;          // compact token buffer (emit only)
  int n_images, mb_w, mb_h;
};

__device__ __constant__ uint8_t c_cat3[3] = {173, 148, 140};
__device__ __constant__ uint8_t c_cat4[4] = {176, 155, 140, 135};
__device__ __constant__ uint8_t c_cat5[5] = {180, 157, 141, 134, 130};
__device__ __constant__ uint8_t c_cat6[11] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129};

// One block's tokens.  P = probabilities of this block's type, [band][ctx][11].  EMIT=false only counts.
template <bool EMIT>
__device__ __forceinline__ int walk_block_tokens(const int16_t* lev, int n_coeffs, const uint8_t* P, int first, int ctx, uint16_t* out) {
  int cnt = 0;
  auto rec = [&](int bit, int prob) {
    if (EMIT) out[cnt] = (uint16_t)((bit & 1) | (prob << 8));
    ++cnt;
  };
  auto pr = [&](int band, int c, int k) -> int { return EMIT ? (int)P[(band * 3 + c) * 11 + k] : 0; };
  int n = first;
  if (n_coeffs <= first) { rec(0, pr(c_bands[n], ctx, 0)); return cnt; }
  while (n < 16) {
    int band = c_bands[n];
    if (n >= n_coeffs) { rec(0, pr(band, ctx, 0)); return cnt; }
    rec(1, pr(band, ctx, 0));
    for (;;) {
      int v = lev[c_zigzag[n]];
      const int sign = v < 0;
      v = abs(v);
      if (v == 0) {
        rec(0, pr(band, ctx, 1));
        if (++n >= 16) return cnt;
        band = c_bands[n];
        ctx = 0;
        continue;
      }
      rec(1, pr(band, ctx, 1));
      if (v == 1) {
        rec(0, pr(band, ctx, 2));
      } else {
        rec(1, pr(band, ctx, 2));
        if (v <= 4) {
          rec(0, pr(band, ctx, 3));
          if (v == 2) rec(0, pr(band, ctx, 4));
          else { rec(1, pr(band, ctx, 4)); rec(v == 3 ? 0 : 1, pr(band, ctx, 5)); }
        } else if (v <= 10) {
          rec(1, pr(band, ctx, 3));
          rec(0, pr(band, ctx, 6));
          if (v <= 6) { rec(0, pr(band, ctx, 7)); rec(v - 5, 159); }
          else { rec(1, pr(band, ctx, 7)); rec((v - 7) >> 1, 165); rec((v - 7) & 1, 145); }
        } else {
          rec(1, pr(band, ctx, 3));
          rec(1, pr(band, ctx, 6));
          const int cat = v <= 18 ? 0 : v <= 34 ? 1 : v <= 66 ? 2 : 3;
          rec(cat >> 1, pr(band, ctx, 8));
          rec(cat & 1, pr(band, ctx, 9 + (cat >> 1)));
          const int extra = v - (3 + (8 << cat));
          const int nbits = cat == 0 ? 3 : cat == 1 ? 4 : cat == 2 ? 5 : 11;
          const uint8_t* tab = cat == 0 ? c_cat3 : cat == 1 ? c_cat4 : cat == 2 ? c_cat5 : c_cat6;
          for (int i = 0; i < nbits; ++i) rec((extra >> (nbits - 1 - i)) & 1, tab[i]);
        }
      }
      rec(sign, 128);
      ctx = (v == 1) ? 1 : 2;
      ++n;
      break;
    }
  }
  return cnt;
}

// 8 lanes per macroblock, no wavefront: every context needed is already in ctxw / hdr.
// Bitstream order of the blocks: [WHT (I16 only)], Y0..15, U0..3, V0..3 -> slot 0 is the WHT block, slots 1..24 the rest.
template <bool EMIT>
__global__ void __launch_bounds__(128) token_kernel(const TokenParams P) {
  __shared__ uint16_t s_cnt[16][26];
  const int lane = threadIdx.x & 31, gl = lane & 7;
  const int grp = threadIdx.x >> 3;  // 16 macroblocks per CTA
  const int nmb = P.mb_w * P.mb_h;
  const long long total = (long long)nmb * P.n_images;
  const long long task = (long long)blockIdx.x * 16 + grp;
  const bool active = task < total;
  const int img = active ? (int)(task / nmb) : 0;
  const int mb = active ? (int)(task - (long long)img * nmb) : 0;
  const int my = mb / P.mb_w, mx = mb - my * P.mb_w;
  const uint8_t* hdr = P.hdr + ((size_t)img * nmb + mb) * 48;
  const int16_t* co = P.coeffs + ((size_t)img * nmb + mb) * 400;
  const uint32_t* ctxw = P.ctxw + (size_t)img * nmb;
  const bool i4 = hdr[0] != 0, skip = hdr[4] != 0;
  uint32_t top_nz = 0, left_nz = 0;
  int top_dc = 0, left_dc = 0;
  if (active && my > 0) { const uint32_t cw = ctxw[mb - P.mb_w]; top_nz = cw & 0xff; top_dc = (cw >> 16) & 1; }
  if (active && mx > 0) { const uint32_t cw = ctxw[mb - 1]; left_nz = (cw >> 8) & 0xff; left_dc = (cw >> 17) & 1; }
  uint16_t* cnt = s_cnt[grp];
  const uint8_t* proba = EMIT ? P.proba + (size_t)img * 1056 : nullptr;
  // per-slot geometry: slot s -> (levels, nz, type, first, ctx)
  auto slot_info = [&](int s, const int16_t*& lev, int& nz, int& type, int& first, int& ctx) {
    if (s == 0) { lev = co + 384; nz = hdr[5]; type = 1; first = 0; ctx = min(top_dc + left_dc, 2); return; }
    const int b = s - 1;
    lev = co + b * 16;
    nz = hdr[24 + b];
    if (b < 16) {
      const int bx = b & 3, by = b >> 2;
      type = i4 ? 3 : 0; first = i4 ? 0 : 1;
      const int l = bx > 0 ? (hdr[24 + b - 1] > first) : (int)((left_nz >> by) & 1);
      const int t = by > 0 ? (hdr[24 + b - 4] > first) : (int)((top_nz >> bx) & 1);
      ctx = l + t;
    } else {
      const int k = b - 16, ch = k >> 2, bx = k & 1, by = (k >> 1) & 1;
      const uint32_t tn = (top_nz >> (4 + 2 * ch)) & 3, ln = (left_nz >> (4 + 2 * ch)) & 3;
      type = 2; first = 0;
      const int l = bx > 0 ? (hdr[24 + b - 1] > 0) : (int)((ln >> by) & 1);
      const int t = by > 0 ? (hdr[24 + b - 2] > 0) : (int)((tn >> bx) & 1);
      ctx = l + t;
    }
  };
  const int s_lo = i4 ? 1 : 0;
  if (active && !skip) {
    for (int s = s_lo + gl; s < 25; s += 8) {
      const int16_t* lev; int nz, type, first, ctx;
      slot_info(s, lev, nz, type, first, ctx);
      cnt[s] = (uint16_t)walk_block_tokens<false>(lev, nz, nullptr, first, ctx, nullptr);
    }
  }
  __syncwarp();
  if (!EMIT) {
    if (active && gl == 0) {
      uint32_t t = 0;
      if (!skip) for (int s = s_lo; s < 25; ++s) t += cnt[s];
      P.mb_tokens[(size_t)img * nmb + mb] = t;
    }
    return;
  }
  if (active && !skip) {
    uint16_t* base = P.tokens + P.img_base[img] + P.mb_offset[(size_t)img * nmb + mb];
    for (int s = s_lo + gl; s < 25; s += 8) {
      uint32_t off = 0;
      for (int k = s_lo; k < s; ++k) off += cnt[k];
      const int16_t* lev; int nz, type, first, ctx;
      slot_info(s, lev, nz, type, first, ctx);
      walk_block_tokens<true>(lev, nz, proba + type * 264, first, ctx, base + off);
    }
  }
}

// Exclusive prefix of the per-macroblock token counts in raster order, one CTA per image.
__global__ void __launch_bounds__(256) token_scan_kernel(const TokenParams P) {
  __shared__ unsigned long long s_part[256];
  const int img = blockIdx.x, nmb = P.mb_w * P.mb_h;
  const uint32_t* in = P.mb_tokens + (size_t)img * nmb;
  unsigned long long* out = P.mb_offset + (size_t)img * nmb;
  const int per = (nmb + 255) / 256;
  const int lo = min(threadIdx.x * per, nmb), hi = min(lo + per, nmb);
  unsigned long long s = 0;
  for (int i = lo; i < hi; ++i) s += in[i];
  s_part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 256; ++i) { const unsigned long long v = s_part[i]; s_part[i] = run; run += v; }
    P.img_total[img] = run;
  }
  __syncthreads();
  unsigned long long run = s_part[threadIdx.x];
  for (int i = lo; i < hi; ++i) { out[i] = run; run += in[i]; }
}

// optimizeProba (encode_proba.go:117-156): one thread per probability slot, int64 costs as Go's int.
struct ProbaParams {
  const unsigned int* stats;  // [n][1056][2]
  const uint8_t* proba0;      // CoeffsProba0 [1056]
  const uint8_t* update;      // CoeffsUpdateProba [1056]
  const uint16_t* ecost;      // VP8EntropyCost [256]
  uint8_t* proba;             // [n][1056]
  int n_images;
};
__global__ void __launch_bounds__(352) optimize_proba_kernel(const ProbaParams P) {
  const int img = blockIdx.x;
  for (int idx = threadIdx.x; idx < 1056; idx += blockDim.x) {
    const long long c0 = P.stats[((size_t)img * 1056 + idx) * 2], c1 = P.stats[((size_t)img * 1056 + idx) * 2 + 1];
    const long long tot = c0 + c1;
    int out = P.proba0[idx];
    if (tot > 0) {
      const int new_p = c1 > 0 ? 255 - (int)(c1 * 255 / tot) : 255;
      const int old_p = P.proba0[idx], up = P.update[idx];
      auto bc = [&](int p) -> long long { p = min(max(p, 1), 255); return c1 * P.ecost[255 - p] + c0 * P.ecost[p]; };
      const long long old_cost = bc(old_p) + P.ecost[up];
      const long long new_cost = bc(new_p) + P.ecost[255 - up] + 8 * 256;
      if (old_cost > new_cost) out = new_p;
    }
    P.proba[(size_t)img * 1056 + idx] = (uint8_t)out;
  }
}

// The boolean coder of the token partitions is boolcode_par.cuh (chunk-parallel).  Its arguments, as webpgpu.cu fills them:
struct BoolCodeParams {
  const uint16_t* tokens;              // compact token buffer
  const unsigned long long* img_base;  // [n] token offset of each image (multiple of 8 tokens)
  const unsigned long long* img_total; // [n] token count
  uint8_t* out;                        // coded partitions
  const unsigned long long* out_base;  // [n] byte offset of each image's partition in `out` (capacity >= total + 16)
  unsigned int* out_size;              // [n] coded size in bytes
  int n_images;
};

}  // namespace wg
