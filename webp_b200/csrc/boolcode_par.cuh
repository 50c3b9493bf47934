// Chunk-parallel boolean coder for the token partitions (VP8BitWriter PutBit / Flush / Finish, internal/bitio/
// writer_bool.go:58-104,140-150, over the flat bit | prob << 8 token arrays of token_kernels.cuh).
//
// The coder of one partition is two recurrences.  The RANGE recurrence R -> R' depends on the tokens only and lives on 128
// states; the CODE VALUE is a big-number sum: token i adds (bit ? split + 1 : 0) with its most significant bit at stream
// bit E_i = sum of the renormalisation shifts before it.  So a partition is cut into chunks of BCP_L tokens and
//   1. bcp_state_kernel   finds every chunk's entry range by relaxation: each chunk is walked from a guessed entry state
//                         (a warm-up over the tokens before it), hands its exit state to the next chunk, and is walked again
//                         only if its entry state changed.  Distinct entry states merge within a chunk, so corrections die
//                         out after a few rounds; the host repeats rounds until none changes anything (exact fixed point:
//                         chunk 0 enters at 255 and every chunk was walked from its predecessor's final exit state).
//                         The same walk leaves the chunk's total shift.
//   2. bcp_scan_kernel    prefix sums of the shifts: the stream bit G_c at which every chunk starts.
//   3. bcp_bytes_kernel   walks every chunk again, now with the byte side of the writer (value, nb_bits, the held-back byte
//                         and the 0xff run, all in registers) started on the partition's byte grid at G_c.  A chunk writes
//                         the bytes only it contributes to; the two bytes it shares with its predecessor's tail (an addend
//                         is 8 bits wide) and the carry out of its first byte go to a boundary record, as do the two tail
//                         bytes it leaves in its successor's first bytes.  The last chunk runs Finish.
//   4. bcp_join_kernel    every chunk boundary: tail + head (+ carries), ripple into the bytes above; bcp_join_fix_kernel
//                         applies the rare ripples that ran through a whole chunk.
// One lane per chunk everywhere: a 256-image batch of 1536x1024 frames is ~220k chunks, i.e. the whole GPU for a few hundred
// microseconds instead of one 160 ms dependency chain per partition.  BCP_L >= 127 * 32 guarantees that a full chunk shifts
// by >= 32 bits (the range shrinks with every token until it renormalises, so at most 127 tokens pass without a shift):
// boundary zones of consecutive chunks never touch.
// The functions are host + device so that the identical code is checked on the CPU against the reference-order coder
// (hostcheck test harness, tests/test_oracle.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BCP_HD __host__ __device__ __forceinline__
#else
#define BCP_HD inline
#endif

#if !defined(BCP_NOTE_PARKED)
#define BCP_NOTE_PARKED() ((void)0)  // the CPU test harness counts how often the rare path runs
#endif

namespace wg {

constexpr int BCP_L = 4096;     // tokens per chunk (the last chunk of a partition takes the remainder: BCP_L .. 2 * BCP_L - 1)
constexpr int BCP_WARM = 256;   // tokens walked before a chunk to guess its entry state

struct BcpParams {
  const uint16_t* tokens;               // compact token buffer
  const unsigned long long* img_base;   // [n] token offset of each partition (multiple of 8 tokens)
  const unsigned long long* img_total;  // [n] token count
  const uint32_t* chunk_first;          // [n + 1] index of each partition's first chunk
  int n_images;
  uint32_t n_chunks;
  uint8_t* entry;                       // [chunks] entry state R (range + 1, 128..255) as currently known
  uint8_t* walked;                      // [chunks] entry state the chunk was last walked from (0 = never)
  uint32_t* shift_total;                // [chunks] sum of the shifts of the chunk's tokens, from `walked`
  uint32_t* chunk_bit;                  // [chunks] stream bit at which the chunk starts
  uint32_t* head;                       // [chunks] first two bytes of the chunk, h0 | h1 << 16 (16-bit fields: ripples add in)
  uint32_t* head_carry;                 // [chunks] carries out of the chunk's first byte
  uint16_t* tail;                       // [chunks] t0 | t1 << 8: what the chunk leaves in the first two bytes of the next
  unsigned int* changed;                // [rounds] entries changed in each relaxation round
  uint32_t* pending;                    // [chunks] a ripple parked at the chunk's boundary zone: byte index << 2 | carry
  uint32_t* any_pending;                // [n] the partition has parked ripples (zeroed with `changed`)
  int round;
  uint8_t* out;                         // coded partitions
  const unsigned long long* out_base;   // [n] byte offset of each partition in `out`
  unsigned int* out_size;               // [n] coded size in bytes
};

BCP_HD int bcp_top_bit(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)v);
#else
  return 31 - __builtin_clz(v);
#endif
}
// One PutBit on the range side.  R = range + 1 in [128, 255].  split = ((R - 1) * prob) >> 8; the sub-range is
// bit ? R - 1 - split : split + 1 = (R * A + C) >> 8 with (A, C) = bit ? (256 - prob, prob - 1) : (prob, 256 - prob); the leading-one
// index of the 16-bit product gives the shift, the renormalised range is ((t & 0xff00) << 7) >> index.
BCP_HD int bcp_step(int& R, uint32_t tok, uint32_t* add) {
  const int prob = (int)(tok >> 8);
  const bool bit = tok & 1u;
  const int A = bit ? 256 - prob : prob, Cc = bit ? prob - 1 : 256 - prob;
  if (add) *add = bit ? (uint32_t)((((R - 1) * prob) >> 8) + 1) : 0u;
  const int t = R * A + Cc;
  const int kk = bcp_top_bit((uint32_t)t);  // >= 8
  R = ((t & 0xff00) << 7) >> kk;
  return 15 - kk;
}

// chunk geometry of a partition of `total` tokens
BCP_HD uint32_t bcp_chunks_of(unsigned long long total) { return total >= 2ull * BCP_L ? (uint32_t)(total / BCP_L) : 1u; }
BCP_HD void bcp_chunk_range(unsigned long long total, uint32_t nch, uint32_t c, unsigned long long* lo, unsigned long long* hi) {
  *lo = (unsigned long long)c * BCP_L;
  *hi = c + 1 == nch ? total : *lo + BCP_L;
}
// which partition a chunk belongs to: the last i with chunk_first[i] <= ck
BCP_HD int bcp_image_of(const uint32_t* chunk_first, int n, uint32_t ck) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (chunk_first[mid] <= ck) lo = mid; else hi = mid - 1;
  }
  return lo;
}
// writer state on the partition's byte grid at stream bit G: nb_bits = G - 8 - 8 * (bytes flushed so far)
BCP_HD void bcp_grid_at(uint32_t G, int* nb_bits, uint32_t* flushed) {
  if (G == 0) { *nb_bits = -8; *flushed = 0; return; }
  *nb_bits = (int)((G - 1) & 7u) - 7;
  *flushed = (uint32_t)(((long long)G - 8 - *nb_bits) >> 3);
}

// eight tokens (one 128-bit load on the device)
struct BcpGroup { uint32_t x, y, z, w; };
BCP_HD BcpGroup bcp_load8(const uint16_t* tk, unsigned long long i) {
  BcpGroup q;
#if defined(__CUDA_ARCH__)
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(tk + i));
  q.x = v.x; q.y = v.y; q.z = v.z; q.w = v.w;
#else
  q.x = tk[i] | ((uint32_t)tk[i + 1] << 16); q.y = tk[i + 2] | ((uint32_t)tk[i + 3] << 16);
  q.z = tk[i + 4] | ((uint32_t)tk[i + 5] << 16); q.w = tk[i + 6] | ((uint32_t)tk[i + 7] << 16);
#endif
  return q;
}
// walk tokens [lo, hi) of a partition from R: exit state and total shift.  lo is a multiple of 8; the group after the one
// being walked is already in flight (a lane's loads are 8 KB apart from its neighbours': nothing coalesces, latency is all)
BCP_HD void bcp_walk(const uint16_t* tk, unsigned long long lo, unsigned long long hi, int* R_io, uint32_t* shift_out) {
  int R = *R_io;
  uint32_t S = 0;
  unsigned long long i = lo;
  const unsigned long long full = lo + ((hi - lo) & ~7ull);
  BcpGroup nxt = {0, 0, 0, 0};
  if (i < full) nxt = bcp_load8(tk, i);
  for (; i < full; i += 8) {
    const BcpGroup q = nxt;
    if (i + 8 < full) nxt = bcp_load8(tk, i + 8);
    S += bcp_step(R, q.x & 0xffffu, nullptr); S += bcp_step(R, q.x >> 16, nullptr);
    S += bcp_step(R, q.y & 0xffffu, nullptr); S += bcp_step(R, q.y >> 16, nullptr);
    S += bcp_step(R, q.z & 0xffffu, nullptr); S += bcp_step(R, q.z >> 16, nullptr);
    S += bcp_step(R, q.w & 0xffffu, nullptr); S += bcp_step(R, q.w >> 16, nullptr);
  }
  for (; i < hi; ++i) S += bcp_step(R, tk[i], nullptr);
  *R_io = R;
  *shift_out = S;
}

// ---- 1. one relaxation round for chunk ck
BCP_HD void bcp_state_chunk(const BcpParams& P, uint32_t ck) {
  const int img = bcp_image_of(P.chunk_first, P.n_images, ck);
  const uint32_t first = P.chunk_first[img], nch = P.chunk_first[img + 1] - first, c = ck - first;
  const unsigned long long total = P.img_total[img];
  const uint16_t* tk = P.tokens + P.img_base[img];
  unsigned long long lo, hi;
  bcp_chunk_range(total, nch, c, &lo, &hi);
  int R;
  if (P.round == 0) {
    R = 255;
    if (c > 0) {  // guess: the state a walk from 255 reaches over the BCP_WARM tokens before the chunk
      uint32_t s;
      bcp_walk(tk, lo - BCP_WARM, lo, &R, &s);
    } else {
      P.entry[ck] = 255;  // the only entry its own chunk writes; all others belong to the chunk before them
    }
  } else {
    R = P.entry[ck];
    if (R == P.walked[ck]) return;
  }
  const int R_in = R;
  uint32_t S;
  bcp_walk(tk, lo, hi, &R, &S);
  P.shift_total[ck] = S;
  P.walked[ck] = (uint8_t)R_in;  // what this walk really started from: a later round compares it with the entry
  if (c + 1 < nch && (P.round == 0 || P.entry[ck + 1] != (uint8_t)R)) {
    P.entry[ck + 1] = (uint8_t)R;
#if defined(__CUDA_ARCH__)
    atomicAdd(&P.changed[P.round], 1u);
#else
    P.changed[P.round] += 1u;
#endif
  }
}

// ---- 3. byte side of one chunk
struct BcpWriter {
  unsigned long long value;
  int nb_bits, run, last;
  uint32_t pos;         // index of the next byte to be written
  uint32_t icpt;        // bytes [icpt, icpt + 2) go to the head record (chunks after the first)
  bool intercept;
  uint32_t h0, h1;
  uint32_t head_carry;
  uint8_t* out;
  BCP_HD void emit(uint32_t b) {
    if (intercept && pos - icpt < 2u) { if (pos == icpt) h0 = b; else h1 = b; } else out[pos] = (uint8_t)b;
    ++pos;
  }
  BCP_HD void carry_in(int carry) {  // a carry into the held-back byte and the 0xff run behind it, then write them
    if (last >= 0) emit((uint32_t)(last + carry)); else head_carry += (uint32_t)carry;
    const uint32_t fill = carry ? 0x00u : 0xffu;
    for (; run > 0; --run) emit(fill);
  }
  BCP_HD void flush() {  // Flush (writer_bool.go:82-104) with the carry applied to a held-back byte
    const int s = 8 + nb_bits;
    const int bits = (int)(value >> s);
    value -= (unsigned long long)bits << s;
    nb_bits -= 8;
    if ((bits & 0xff) != 0xff) {
      carry_in((bits >> 8) & 1);
      last = bits & 0xff;
    } else {
      ++run;
    }
  }
  BCP_HD void put(uint32_t add, int shift) {
    value = (value + add) << shift;
    nb_bits += shift;
    if (nb_bits > 0) flush();
  }
};

BCP_HD void bcp_bytes_chunk(const BcpParams& P, uint32_t ck) {
  const int img = bcp_image_of(P.chunk_first, P.n_images, ck);
  const uint32_t first = P.chunk_first[img], nch = P.chunk_first[img + 1] - first, c = ck - first;
  const unsigned long long total = P.img_total[img];
  const uint16_t* tk = P.tokens + P.img_base[img];
  unsigned long long lo, hi;
  bcp_chunk_range(total, nch, c, &lo, &hi);
  BcpWriter W;
  W.value = 0; W.run = 0; W.last = -1; W.head_carry = 0; W.h0 = W.h1 = 0;
  bcp_grid_at(P.chunk_bit[ck], &W.nb_bits, &W.pos);
  W.icpt = W.pos; W.intercept = c > 0;
  W.out = P.out + P.out_base[img];
  int R = c > 0 ? (int)P.entry[ck] : 255;
  unsigned long long i = lo;
  const unsigned long long full = lo + ((hi - lo) & ~7ull);
  BcpGroup nxt = {0, 0, 0, 0};
  if (i < full) nxt = bcp_load8(tk, i);
  for (; i < full; i += 8) {
    const BcpGroup q = nxt;
    if (i + 8 < full) nxt = bcp_load8(tk, i + 8);
    uint32_t a; int s;
    s = bcp_step(R, q.x & 0xffffu, &a); W.put(a, s); s = bcp_step(R, q.x >> 16, &a); W.put(a, s);
    s = bcp_step(R, q.y & 0xffffu, &a); W.put(a, s); s = bcp_step(R, q.y >> 16, &a); W.put(a, s);
    s = bcp_step(R, q.z & 0xffffu, &a); W.put(a, s); s = bcp_step(R, q.z >> 16, &a); W.put(a, s);
    s = bcp_step(R, q.w & 0xffffu, &a); W.put(a, s); s = bcp_step(R, q.w >> 16, &a); W.put(a, s);
  }
  for (; i < hi; ++i) { uint32_t a; const int s = bcp_step(R, tk[i], &a); W.put(a, s); }
  if (c + 1 < nch) {
    // the pending bits (< 2^(16 + nb_bits), plus a carry) are the chunk's share of the next two bytes: align them to the byte
    // grid, hand the carry to the held-back byte, write what is held back, leave the two bytes to the join
    const unsigned long long v = W.value << (-W.nb_bits);
    W.carry_in((int)((v >> 16) & 1u));
    P.tail[ck] = (uint16_t)(((v >> 8) & 0xffu) | ((v & 0xffu) << 8));
  } else {
    // Finish (writer_bool.go:140-150): PutBits(0, 9 - nbBits) at probability 128, then nbBits = 0 and a last Flush
    const int n_fin = 9 - W.nb_bits;
    for (int k = 0; k < n_fin; ++k) { const int s = bcp_step(R, 128u << 8, nullptr); W.put(0u, s); }
    W.nb_bits = 0;
    W.flush();
    if (W.last >= 0) W.emit((uint32_t)W.last);
    P.out_size[img] = W.pos;
    P.tail[ck] = 0;
  }
  P.head[ck] = W.h0 | (W.h1 << 16);
  P.head_carry[ck] = W.head_carry;
}

// ---- 4. the boundary between chunk ck - 1 and chunk ck (ck is not the first chunk of its partition): tail + head + carries
// give the two shared bytes; what is left over ripples into the bytes above, which belong to chunk ck - 1 alone down to ITS
// boundary zone -- a ripple that gets that far (a chunk-long run of 0xff) is parked in pending[ck - 1] and applied once every
// boundary byte is in memory (bcp_join_fix_image), so that all boundaries can be resolved side by side.
BCP_HD void bcp_join_boundary(const BcpParams& P, uint32_t ck) {
  const int img = bcp_image_of(P.chunk_first, P.n_images, ck);
  const uint32_t first = P.chunk_first[img], c = ck - first;
  if (c == 0) return;
  uint8_t* out = P.out + P.out_base[img];
  int nb; uint32_t F;
  bcp_grid_at(P.chunk_bit[ck], &nb, &F);
  const uint32_t hd = P.head[ck], tl = P.tail[ck - 1];
  const uint32_t x1 = (tl >> 8) + (hd >> 16);
  const uint32_t x0 = (tl & 0xffu) + (hd & 0xffffu) + (x1 >> 8);
  uint32_t cy = (x0 >> 8) + P.head_carry[ck];
  out[F] = (uint8_t)x0;
  out[F + 1] = (uint8_t)x1;
  uint32_t lo_zone = 0;
  const bool zone = c - 1 > 0;
  if (zone) { int nb2; bcp_grid_at(P.chunk_bit[ck - 1], &nb2, &lo_zone); }
  uint32_t p = F - 1, parked = 0;
  while (cy) {
    if (zone && p < lo_zone + 2) { parked = (p << 2) | cy; BCP_NOTE_PARKED(); break; }
    const uint32_t b = out[p] + cy;
    out[p] = (uint8_t)b;
    cy = b >> 8;
    if (p == 0) break;
    --p;
  }
  P.pending[ck - 1] = parked;
  if (parked) P.any_pending[img] = 1u;
}
// the parked ripples of one partition: plain big-number carries now that every byte is in memory
BCP_HD void bcp_join_fix_image(const BcpParams& P, int img) {
  if (!P.any_pending[img]) return;
  const uint32_t first = P.chunk_first[img], nch = P.chunk_first[img + 1] - first;
  uint8_t* out = P.out + P.out_base[img];
  for (uint32_t c = nch - 1; c-- > 1;) {  // chunks nch - 2 .. 1 can hold a parked ripple
    const uint32_t v = P.pending[first + c];
    if (!v) continue;
    uint32_t p = v >> 2, cy = v & 3u;
    while (cy) {
      const uint32_t b = out[p] + cy;
      out[p] = (uint8_t)b;
      cy = b >> 8;
      if (p == 0) break;
      --p;
    }
  }
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(128) bcp_state_kernel(const BcpParams P) {
  const uint32_t ck = blockIdx.x * blockDim.x + threadIdx.x;
  if (ck < P.n_chunks) bcp_state_chunk(P, ck);
}
// exclusive prefix of the chunk shifts of every partition: one warp per partition
__global__ void __launch_bounds__(128) bcp_scan_kernel(const BcpParams P) {
  const int img = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (img >= P.n_images) return;
  const uint32_t first = P.chunk_first[img], nch = P.chunk_first[img + 1] - first;
  uint32_t run = 0;
  for (uint32_t base = 0; base < nch; base += 32) {
    const uint32_t k = base + lane;
    const uint32_t v = k < nch ? P.shift_total[first + k] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (k < nch) P.chunk_bit[first + k] = run + inc - v;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
}
__global__ void __launch_bounds__(128) bcp_bytes_kernel(const BcpParams P) {
  const uint32_t ck = blockIdx.x * blockDim.x + threadIdx.x;
  if (ck < P.n_chunks) bcp_bytes_chunk(P, ck);
}
__global__ void __launch_bounds__(128) bcp_join_kernel(const BcpParams P) {
  const uint32_t ck = blockIdx.x * blockDim.x + threadIdx.x;
  if (ck < P.n_chunks) bcp_join_boundary(P, ck);
}
__global__ void __launch_bounds__(64) bcp_join_fix_kernel(const BcpParams P) {
  const int img = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (img < P.n_images) bcp_join_fix_image(P, img);
}
// Frames for the host: [partition 0][token partition] of image 0, of image 1, ... back to back, so that one copy brings a
// batch's coded bytes over.  sizes[k] / in_base[k]: k < n the token partition of image k, k >= n partition 0 of image k - n.
struct FramePackParams {
  const uint8_t* coded; const unsigned long long* in_base; const unsigned int* sizes; uint8_t* packed; unsigned long long* offsets; int n;
};
// one block: exclusive prefix of size(p0_i) + size(tok_i) over the images
__global__ void __launch_bounds__(1024) frame_pack_scan_kernel(const FramePackParams P) {
  __shared__ unsigned long long s_part[1024];
  const int per = (P.n + 1023) / 1024, lo = min((int)threadIdx.x * per, P.n), hi = min(lo + per, P.n);
  unsigned long long sum = 0;
  for (int i = lo; i < hi; ++i) sum += (unsigned long long)P.sizes[i] + P.sizes[P.n + i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; ++i) { const unsigned long long v = s_part[i]; s_part[i] = run; run += v; }
  }
  __syncthreads();
  unsigned long long run = s_part[threadIdx.x];
  for (int i = lo; i < hi; ++i) { P.offsets[i] = run; run += (unsigned long long)P.sizes[i] + P.sizes[P.n + i]; }
}
// grid = (n, chunks of 16 KB)
__global__ void __launch_bounds__(256) frame_pack_copy_kernel(const FramePackParams P) {
  const int img = blockIdx.x;
  const unsigned int p0 = P.sizes[P.n + img], tk = P.sizes[img], total = p0 + tk;
  const unsigned int begin = blockIdx.y * 16384u;
  if (begin >= total) return;
  const unsigned int end = min(begin + 16384u, total);
  const uint8_t* a = P.coded + P.in_base[P.n + img];
  const uint8_t* b = P.coded + P.in_base[img];
  uint8_t* dst = P.packed + P.offsets[img];
  for (unsigned int i = begin + threadIdx.x; i < end; i += 256) dst[i] = i < p0 ? a[i] : b[i - p0];
}
#endif

}  // namespace wg
