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
  int n_images;
};
__global__ void __launch_bounds__(64) boolcode_kernel(const BoolCodeParams P) {
  // Two warps per partition, pipelined over 512-token chunks through shared memory:
  //   warp 0 (range warp): streams the tokens in (128-bit loads, all lanes), then lane 0 runs the range recurrence
  //                        R -> split -> sub-range -> renormalise and leaves (value increment, shift) per token;
  //   warp 1 (byte warp):  lane 0 folds those into the pending value, four tokens per step, and writes the bytes
  //                        (Flush with its carry / 0xff run), several per step when due.
  // The range recurrence is the irreducible serial chain (IMAD -> SHF -> IADD -> FLO -> SHF per token); everything else is
  // kept off it.  Deferring Flush by up to four tokens is exact: value is a big-number accumulator whose carries ripple
  // inside the 64-bit register exactly as Flush would have applied them to the held-back byte.
  constexpr int CHUNK = 512;
  __shared__ uint4 s_tok[CHUNK / 8];
  __shared__ uint16_t s_fin[20];
  __shared__ __align__(16) uint16_t s_ev[2][CHUNK];  // per token: bits 0-8 value increment (0 or split + 1), bits 12-14 shift
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long total = P.img_total[img];
  const uint4* tk = reinterpret_cast<const uint4*>(P.tokens + P.img_base[img]);
  uint8_t* out = P.out + P.out_base[img];
  const long long n_chunks = (long long)((total + CHUNK - 1) / CHUNK);
  // range warp state: R = range + 1 in [128, 255]
  int R = 255;
  uint4 r0 = make_uint4(0, 0, 0, 0), r1v = r0;
  if (warp == 0 && n_chunks > 0) { r0 = __ldg(tk + lane); r1v = __ldg(tk + 32 + lane); }
  auto step = [&](uint32_t tok) -> uint32_t {  // one PutBit on the range side; returns the event for the byte side
    const int prob = (int)(tok >> 8);
    const bool bit = tok & 1u;
    // split = ((R - 1) * prob) >> 8; sub-range r + 1 = bit ? R - 1 - split : split + 1.  With sx = bit ? ~split : split
    // (one IMAD + arithmetic shift, since ~(t >> 8) == (~t) >> 8): r + 1 = (bit ? R : 1) + sx.
    const int ma = bit ? -prob : prob, mc = bit ? prob - 1 : -prob;
    const int sx = (R * ma + mc) >> 8;
    const int r1 = (bit ? R : 1) + sx;
    const int k = 31 - __clz(r1);  // kNorm[r] = 7 - k, kNewRange[r] + 1 = r1 << (7 - k)
    R = (r1 << 7) >> k;
    return (uint32_t)((bit ? -sx : 0) | ((7 - k) << 12));
  };
  // byte warp state
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
  auto emit = [&](uint32_t ev) {  // PutBit's value side: value += increment; value <<= shift; nbBits += shift; Flush when > 0
    const int shift = (int)(ev >> 12);
    value = (value + (ev & 0x1ffu)) << shift;
    nb_bits += shift;
    if (nb_bits > 0) flush();
  };
  for (long long c = 0; c <= n_chunks; ++c) {
    if (warp == 0) {
      if (c < n_chunks) {
        s_tok[lane] = r0; s_tok[32 + lane] = r1v;
        if (c + 1 < n_chunks) { r0 = __ldg(tk + (c + 1) * (CHUNK / 8) + lane); r1v = __ldg(tk + (c + 1) * (CHUNK / 8) + 32 + lane); }
        __syncwarp();
        if (lane == 0) {
          const unsigned long long left = total - (unsigned long long)c * CHUNK;
          const int cnt = left < (unsigned long long)CHUNK ? (int)left : CHUNK;
          const uint2* t2 = reinterpret_cast<const uint2*>(s_tok);
          uint2* ev2 = reinterpret_cast<uint2*>(s_ev[c & 1]);
          const int groups = cnt >> 2;
          uint2 nxt = t2[0];
#pragma unroll 1
          for (int g = 0; g < groups; ++g) {
            const uint2 cur = nxt;
            nxt = t2[(g + 1) & (CHUNK / 4 - 1)];  // one group ahead: the shared-memory latency stays off the chain
            const uint32_t e0 = step(cur.x & 0xffffu), e1 = step(cur.x >> 16), e2 = step(cur.y & 0xffffu), e3 = step(cur.y >> 16);
            ev2[g] = make_uint2(e0 | (e1 << 16), e2 | (e3 << 16));
          }
          const uint16_t* t = reinterpret_cast<const uint16_t*>(s_tok);
          for (int i = groups * 4; i < cnt; ++i) s_ev[c & 1][i] = (uint16_t)step(t[i]);
        }
        __syncwarp();
      } else if (lane == 0) {
        // Finish = PutBits(0, 9 - nbBits): up to 17 zero bits at probability 128, whose shifts still depend on the range
        for (int i = 0; i < 17; ++i) s_fin[i] = (uint16_t)step(128u << 8);
      }
    } else if (c > 0 && lane == 0) {
      const unsigned long long left = total - (unsigned long long)(c - 1) * CHUNK;
      const int cnt = left < (unsigned long long)CHUNK ? (int)left : CHUNK;
      const uint2* ev2 = reinterpret_cast<const uint2*>(s_ev[(c - 1) & 1]);
      const int groups = cnt >> 2;
      uint2 nxt = ev2[0];
#pragma unroll 1
      for (int g = 0; g < groups; ++g) {
        const uint2 cur = nxt;
        nxt = ev2[(g + 1) & (CHUNK / 4 - 1)];
        // ((((v + a0) << s0) + a1) << s1 ...) == (v << S0) + (a0 << S0) + (a1 << S1) + (a2 << S2) + (a3 << S3), Sj = sj + ... + s3
        const int s3 = (int)(cur.y >> 28), s2 = s3 + (int)((cur.y >> 12) & 7u), s1 = s2 + (int)(cur.x >> 28), s0 = s1 + (int)((cur.x >> 12) & 7u);
        const unsigned long long add = ((unsigned long long)(cur.x & 0x1ffu) << s0) + ((unsigned long long)((cur.x >> 16) & 0x1ffu) << s1) +
                                       ((unsigned long long)(cur.y & 0x1ffu) << s2) + ((unsigned long long)((cur.y >> 16) & 0x1ffu) << s3);
        value = (value << s0) + add;
        nb_bits += s0;
        while (nb_bits > 0) flush();
      }
      const uint16_t* ev = s_ev[(c - 1) & 1];
      for (int i = groups * 4; i < cnt; ++i) emit(ev[i]);
    }
    __syncthreads();
  }
  if (warp == 1 && lane == 0) {
    const int n_fin = 9 - nb_bits;  // the closing bits were prepared by the range warp in the last pipeline step
    for (int i = 0; i < n_fin; ++i) emit(s_fin[i]);
    nb_bits = 0;
    flush();
    if (last >= 0) out[pos++] = (uint8_t)last;
    P.out_size[img] = pos;
  }
}

}  // namespace wg
