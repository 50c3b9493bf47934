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

// Boolean coder on the device: the token partition of one image per warp.  Restates VP8BitWriter PutBit / Flush / Finish
// (internal/bitio/writer_bool.go:58-104,140-150) over the flat (bit | prob << 8) token array emitted above.  The coder is
// a serial dependency chain per partition (range -> split -> renormalise), so one lane codes while the warp streams the
// tokens through shared memory with 128-bit loads; parallelism comes from the batch (one warp per image partition), and
// the kernel is small enough (<= 32 registers, 2 KB shared) to run beside the next batch's mode-search waves.  The carry
// of Flush (buf[pos-1]++) is applied to a byte held back in a register instead of a read-modify-write in HBM.
struct BoolCodeParams {
  const uint16_t* tokens;              // compact token buffer
  const unsigned long long* img_base;  // [n] token offset of each image (multiple of 8 tokens)
  const unsigned long long* img_total; // [n] token count
  uint8_t* out;                        // coded partitions
  const unsigned long long* out_base;  // [n] byte offset of each image's partition in `out` (capacity >= total + 16)
  unsigned int* out_size;              // [n] coded size in bytes
  const int* order;                    // [n] partition handled by (block, warp pair) slot k = order[k]: longest first
  int n_images;
};
// Shared-memory layout of one coder block (32 partitions): per-lane regions of 64 tokens / events with a 144-byte stride,
// which keeps both the 16-byte cp.async destinations aligned and the per-lane 128-bit reads conflict-free.
constexpr int BOOLCODE_T = 64;            // tokens per lane per chunk
constexpr int BOOLCODE_STRIDE = 144;      // bytes per lane region
constexpr int BOOLCODE_SMEM = 4 * 32 * BOOLCODE_STRIDE + 32 * 40 + 32 * 8 * 2;  // per warp pair: 2 token + 2 event buffers, closing events, bases/totals
constexpr int BOOLCODE_MAX_PAIRS = 4;     // warp pairs (x 32 partitions) a block may carry; 1 measured best (packing slows the chains)

__global__ void __launch_bounds__(256) boolcode_kernel(const BoolCodeParams P) {
  // SIMT across partitions: a block codes 32 partitions (slots 32b .. 32b+31 of the longest-first order), one per lane, with
  // two warps pipelined over 64-token chunks through shared memory:
  //   warp 0 (range warp): cp.async streams each lane's tokens in; every lane runs the range recurrence of ITS partition
  //                        R -> split -> sub-range -> renormalise and leaves (value increment, shift) per token;
  //   warp 1 (byte warp):  every lane folds its events into the pending value, four tokens per step, and writes its
  //                        bytes (Flush with its carry / 0xff run; the only divergent part).
  // The range recurrence is the irreducible serial chain of a partition (IMAD -> SHF -> IADD -> FLO -> SHF per token), so
  // the time of a batch is the chain of its longest partition whatever the layout -- what SIMT buys is that the whole batch
  // of 256 partitions occupies 16 warps instead of 512, and can sit on a few SMs beside the next batch's mode search.
  // Deferring Flush by up to four tokens is exact: value is a big-number accumulator whose carries ripple inside the
  // 64-bit register exactly as Flush would have applied them to the held-back byte.
  constexpr int T = BOOLCODE_T, STRIDE = BOOLCODE_STRIDE;
  // A block carries blockDim / 64 such warp pairs (32 partitions each), every pair on its own named barrier: the pairs of a
  // block land on different SM sub-partitions, so packing four of them halves the SMs the coder takes from the waves twice.
  extern __shared__ __align__(16) unsigned char s_dyn_all[];
  const int pair = threadIdx.x >> 6;
  unsigned char* s_dyn = s_dyn_all + (size_t)pair * BOOLCODE_SMEM;
  auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); };
  unsigned char* s_tok = s_dyn;                                   // [2][32][STRIDE]
  unsigned char* s_ev = s_dyn + 2 * 32 * STRIDE;                  // [2][32][STRIDE]
  uint16_t* s_fin = reinterpret_cast<uint16_t*>(s_dyn + 4 * 32 * STRIDE);  // [32][20]
  unsigned long long* s_base = reinterpret_cast<unsigned long long*>(s_dyn + 4 * 32 * STRIDE + 32 * 40);  // [32] token offset
  unsigned long long* s_total = s_base + 32;                                                                // [32] token count
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 1;
  const int slot = (blockIdx.x * (blockDim.x >> 6) + pair) * 32 + lane;
  const int img = slot < P.n_images ? (P.order ? P.order[slot] : slot) : -1;
  const unsigned long long total = img >= 0 ? P.img_total[img] : 0ull;
  if (warp == 0) { s_base[lane] = img >= 0 ? P.img_base[img] : 0ull; s_total[lane] = total; }
  unsigned long long max_total = total;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, max_total, o); max_total = v > max_total ? v : max_total; }
  const long long n_chunks = (long long)((max_total + T - 1) / T);
  pair_sync();
  // ---- range warp state: R = range + 1 in [128, 255]
  int R = 255;
  auto step = [&](uint32_t tok, bool valid) -> uint32_t {  // one PutBit on the range side; returns the event for the byte side
    const int prob = (int)(tok >> 8);
    const bool bit = tok & 1u;
    // split = ((R - 1) * prob) >> 8; sub-range r + 1 = bit ? R - 1 - split : split + 1.  Both cases are one multiply-add on R:
    //   r + 1 = (R * A + Cc) >> 8  with (A, Cc) = bit ? (256 - prob, prob - 1) : (prob, 256 - prob),
    // and a lane past the end of its partition takes (256, 0): the identity.  The pair depends on the token only, so the
    // chain per token is IMAD -> FLO -> SHF: the leading-one index is taken on the 16-bit product itself (it is the index of
    // r + 1 plus 8), and the renormalised range is ((t & 0xff00) << 7) >> index.
    //   (measured and rejected: the same recurrence as a 64 KB shared-memory table rtab[token][range]: 2x slower, a byte load
    //   per token whose address hangs on the previous load)
    const int A = valid ? (bit ? 256 - prob : prob) : 256, Cc = valid ? (bit ? prob - 1 : 256 - prob) : 0;
    const int Rold = R;
    const int t = Rold * A + Cc;
    const int kk = 31 - __clz(t);  // >= 8
    R = ((t & 0xff00) << 7) >> kk;
    // off the chain: the value increment (bit ? split + 1 : 0, as -sx of the former formulation) and the shift 7 - k = 15 - kk
    const int sx = (Rold * (bit ? -prob : prob) + (bit ? prob - 1 : -prob)) >> 8;
    return valid ? (uint32_t)((bit ? -sx : 0) | ((15 - kk) << 12)) : 0u;  // an all-zero event is a no-op on the byte side
  };
  auto issue_loads = [&](long long c) {  // chunk c of all 32 partitions -> token buffer c & 1 (8 lanes x 16 bytes per partition)
    unsigned char* dst_buf = s_tok + (size_t)(c & 1) * 32 * STRIDE;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int st = 4 * k + (lane >> 3), part = lane & 7;
      if ((unsigned long long)c * T + part * 8 < s_total[st]) {
        const uint16_t* src = P.tokens + s_base[st] + (unsigned long long)c * T + part * 8;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dst_buf + st * STRIDE + part * 16);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ---- byte warp state
  uint8_t* out = img >= 0 ? P.out + P.out_base[img] : nullptr;
  unsigned long long value = 0;
  int run = 0, nb_bits = -8, last = -1;
  unsigned int pos = 0;
  auto flush = [&]() {
    const int s = 8 + nb_bits;
    const int bits = (int)(value >> s);
    value -= (unsigned long long)bits << s;
    nb_bits -= 8;
    if ((bits & 0xff) != 0xff) {
      const int carry = (bits >> 8) & 1;
      if (last >= 0) out[pos++] = (uint8_t)(last + carry);
      if (run > 0) {
        const uint8_t fill = carry ? 0x00 : 0xff;
#pragma unroll 1
        for (; run > 0; --run) out[pos++] = fill;
      }
      last = bits & 0xff;
    } else {
      ++run;
    }
  };
  auto fold4 = [&](uint32_t x, uint32_t y) {  // four events: x = e0 | e1 << 16, y = e2 | e3 << 16
    // ((((v + a0) << s0) + a1) << s1 ...) == (v << S0) + (a0 << S0) + (a1 << S1) + (a2 << S2) + (a3 << S3), Sj = sj + ... + s3
    const int s3 = (int)(y >> 28), s2 = s3 + (int)((y >> 12) & 7u), s1 = s2 + (int)(x >> 28), s0 = s1 + (int)((x >> 12) & 7u);
    const unsigned long long add = ((unsigned long long)(x & 0x1ffu) << s0) + ((unsigned long long)((x >> 16) & 0x1ffu) << s1) +
                                   ((unsigned long long)(y & 0x1ffu) << s2) + ((unsigned long long)((y >> 16) & 0x1ffu) << s3);
    value = (value << s0) + add;
    nb_bits += s0;
    // All complete bytes of this step at once (1-4 of them under a carry bit).  Fast path: no 0xff among them and no 0xff
    // run pending -> the held-back byte takes the carry and goes out, the new bytes follow, the last one is held back, and
    // the pending value keeps its low bits (an AND: the extraction below stays off the value's dependency chain).
    // Anything involving 0xff goes through Flush byte by byte (rare, divergent).
    const int nbytes = (nb_bits + 7) >> 3;  // <= 0 when nothing is due
    if (nbytes > 0) {
      const int s_low = 16 + nb_bits - 8 * nbytes;  // what Flush leaves pending after the last of these bytes
      const unsigned long long cb = value >> s_low;  // carry bit + nbytes bytes
      const uint32_t bytes = (uint32_t)cb;            // the bytes, most significant first from bit 8 * nbytes - 1 down
      const uint32_t ff = bytes & (bytes >> 1), f2 = ff & (ff >> 2), f4 = f2 & (f2 >> 4);  // bit 8j set iff byte j == 0xff
      const uint32_t mask = nbytes == 4 ? 0x01010101u : ((1u << (8 * nbytes)) - 1u) & 0x01010101u;
      if (run > 0 || (f4 & mask) != 0) {
        while (nb_bits > 0) flush();
      } else {
        const int carry = (int)(cb >> (8 * nbytes)) & 1;
        if (last >= 0) out[pos++] = (uint8_t)(last + carry);
        if (nbytes > 1) out[pos++] = (uint8_t)(bytes >> (8 * nbytes - 8));
        if (nbytes > 2) out[pos++] = (uint8_t)(bytes >> (8 * nbytes - 16));
        if (nbytes > 3) out[pos++] = (uint8_t)(bytes >> 8);
        last = (int)(bytes & 0xffu);
        value &= (1ull << s_low) - 1ull;
        nb_bits -= 8 * nbytes;
      }
    }
  };
  if (warp == 0 && n_chunks > 0) issue_loads(0);
  for (long long c = 0; c <= n_chunks; ++c) {
    if (warp == 0) {
      if (c < n_chunks) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (c + 1 < n_chunks) issue_loads(c + 1);
        const uint4* tk = reinterpret_cast<const uint4*>(s_tok + (size_t)(c & 1) * 32 * STRIDE + lane * STRIDE);
        uint4* ev = reinterpret_cast<uint4*>(s_ev + (size_t)(c & 1) * 32 * STRIDE + lane * STRIDE);
        const unsigned long long done = (unsigned long long)c * T;
        const int cnt = total > done ? (total - done < (unsigned long long)T ? (int)(total - done) : T) : 0;
        uint4 nxt = tk[0];
#pragma unroll 1
        for (int q = 0; q < T / 8; ++q) {
          const uint4 cur = nxt;
          nxt = tk[(q + 1) & (T / 8 - 1)];  // one group ahead: the shared-memory latency stays off the chain
          const int i0 = q * 8;
          uint4 e;
          e.x = step(cur.x & 0xffffu, i0 + 0 < cnt) | (step(cur.x >> 16, i0 + 1 < cnt) << 16);
          e.y = step(cur.y & 0xffffu, i0 + 2 < cnt) | (step(cur.y >> 16, i0 + 3 < cnt) << 16);
          e.z = step(cur.z & 0xffffu, i0 + 4 < cnt) | (step(cur.z >> 16, i0 + 5 < cnt) << 16);
          e.w = step(cur.w & 0xffffu, i0 + 6 < cnt) | (step(cur.w >> 16, i0 + 7 < cnt) << 16);
          ev[q] = e;
        }
      } else {
        // Finish = PutBits(0, 9 - nbBits): up to 17 zero bits at probability 128, whose shifts still depend on the range
        for (int i = 0; i < 17; ++i) s_fin[lane * 20 + i] = (uint16_t)step(128u << 8, true);
      }
    } else if (c > 0) {
      const uint4* ev = reinterpret_cast<const uint4*>(s_ev + (size_t)((c - 1) & 1) * 32 * STRIDE + lane * STRIDE);
      uint4 nxt = ev[0];
#pragma unroll 1
      for (int q = 0; q < T / 8; ++q) {
        const uint4 cur = nxt;
        nxt = ev[(q + 1) & (T / 8 - 1)];
        fold4(cur.x, cur.y);
        fold4(cur.z, cur.w);
      }
    }
    pair_sync();
  }
  if (warp == 1 && img >= 0) {
    const int n_fin = 9 - nb_bits;  // the closing bits were prepared by the range warp in the last pipeline step
    for (int i = 0; i < n_fin; ++i) {
      const uint32_t e = s_fin[lane * 20 + i];
      const int shift = (int)(e >> 12);
      value <<= shift;
      nb_bits += shift;
      if (nb_bits > 0) flush();
    }
    nb_bits = 0;
    flush();
    if (last >= 0) out[pos++] = (uint8_t)last;
    P.out_size[img] = pos;
  }
}

}  // namespace wg
