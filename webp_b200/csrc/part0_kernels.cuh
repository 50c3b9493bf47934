// Partition 0 of a frame as a token stream generated on the device: the frame header bits (segment header, filter header,
// quantiser, coefficient probability updates, skip probability -- encode_syntax.go:118-330 as emit_partition0 in host_enc.h
// walks them) followed by the per-macroblock intra modes (segment id, skip flag, I16 / I4 modes through kI4Tree, chroma mode;
// encode_syntax.go:420-520).  Every bit becomes a (bit | prob << 8) token exactly as the VP8BitWriter would receive it, so the
// partition is coded by the same chunk-parallel coder as the token partitions (boolcode_par.cuh) and the host writes none of it.
// A macroblock's tokens depend on its own header and the modes of its left / top neighbours only: one thread per macroblock.
#pragma once
#include "token_kernels.cuh"

namespace wg {

struct P0Plan {  // the FramePlan fields partition 0 carries (host_enc.h FramePlan)
  uint8_t seg_use, seg_update_map, f_simple, f_level, f_sharpness, parts_code, base_quant, pad0;
  int8_t seg_quantizer[4], seg_fstrength[4];
  uint8_t seg_proba[3], pad1;
  int16_t dq_uv_dc, dq_uv_ac;
};
struct I4PathDev { uint8_t n; uint8_t idx[8]; uint8_t bit[8]; };  // == wgh::I4Path

struct P0Params {
  const uint8_t* hdr;        // [n][nmb][48]
  const uint8_t* segment;    // [n][nmb]
  const uint8_t* proba;      // [n][1056] final probabilities
  const uint8_t* proba0;     // kCoeffsProba0 [1056]
  const uint8_t* update;     // kCoeffsUpdateProba [1056]
  const uint8_t* bmodes;     // kBModesProba [10][10][9]
  const I4PathDev* i4paths;  // [10]
  const P0Plan* plan;        // [n]
  uint32_t* info;            // [n][4]: skipped macroblocks, skip probability, header tokens, 0
  uint32_t* mb_tokens;       // [n][nmb] mode tokens per macroblock
  const unsigned long long* mb_offset;  // [n][nmb] exclusive prefix (emit only)
  const unsigned long long* img_base;   // [n] offset of each image's partition-0 tokens in `tokens` (emit only)
  uint16_t* tokens;
  int n_images, mb_w, mb_h;
};

struct P0Sink {  // counts, and writes when `out` is set
  uint16_t* out;
  uint32_t cnt;
  __device__ __forceinline__ void put(int bit, int prob) { if (out) out[cnt] = (uint16_t)((bit & 1) | (prob << 8)); ++cnt; }
  __device__ __forceinline__ void put_uniform(int bit) { put(bit, 128); }
  __device__ __forceinline__ void put_bits(uint32_t v, int n) { for (uint32_t m = 1u << (n - 1); m; m >>= 1) put_uniform((v & m) ? 1 : 0); }
  __device__ __forceinline__ void put_signed(int v, int n) {  // PutSignedBits (bitio/writer_bool.go:152)
    put_uniform(v != 0);
    if (v == 0) return;
    if (v < 0) put_bits(((uint32_t)(-v) << 1) | 1u, n + 1); else put_bits((uint32_t)v << 1, n + 1);
  }
};

// the header part, in the order of emit_partition0 (host_enc.h) / encode_syntax.go:118-330
__device__ inline uint32_t p0_header(const P0Plan& fp, const uint8_t* proba, const uint8_t* proba0, const uint8_t* update, int num_skip,
                                     int skip_proba, uint16_t* out) {
  P0Sink bw{out, 0};
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
  bw.put_uniform(0);
  bw.put_bits((uint32_t)fp.parts_code, 2);
  bw.put_bits((uint32_t)fp.base_quant, 7);
  bw.put_signed(0, 4);
  bw.put_signed(0, 4);
  bw.put_signed(0, 4);
  bw.put_signed(fp.dq_uv_dc, 4);
  bw.put_signed(fp.dq_uv_ac, 4);
  bw.put_uniform(0);
  for (int i = 0; i < 1056; ++i) {
    const int pr = proba[i], up = update[i];
    if (pr != proba0[i]) { bw.put(1, up); bw.put_bits((uint32_t)pr, 8); } else bw.put(0, up);
  }
  if (num_skip > 0) { bw.put_uniform(1); bw.put_bits((uint32_t)skip_proba, 8); } else bw.put_uniform(0);
  return bw.cnt;
}

// skipped macroblocks, skip probability and header token count of every image: one block per image
__global__ void __launch_bounds__(256) p0_info_kernel(const P0Params P) {
  __shared__ int s_cnt[8];
  const int img = blockIdx.x, nmb = P.mb_w * P.mb_h;
  const uint8_t* H = P.hdr + (size_t)img * nmb * 48;
  int c = 0;
  for (int i = threadIdx.x; i < nmb; i += 256) c += H[(size_t)i * 48 + 4] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int num_skip = 0;
    for (int w = 0; w < 8; ++w) num_skip += s_cnt[w];
    const int skip_proba = num_skip > 0 ? (nmb - num_skip) * 255 / nmb : 0;
    uint32_t* info = P.info + (size_t)img * 4;
    info[0] = (uint32_t)num_skip; info[1] = (uint32_t)skip_proba;
    info[2] = p0_header(P.plan[img], P.proba + (size_t)img * 1056, P.proba0, P.update, num_skip, skip_proba, nullptr);
    info[3] = 0;
  }
}

// grid = (ceil(nmb / 128), n): the mode tokens of one macroblock per thread; EMIT also writes the header (thread 0 of block 0)
template <bool EMIT>
__global__ void __launch_bounds__(128) p0_mb_kernel(const P0Params P) {
  const int img = blockIdx.y, nmb = P.mb_w * P.mb_h;
  const int idx = blockIdx.x * 128 + threadIdx.x;
  const uint32_t* info = P.info + (size_t)img * 4;
  const int num_skip = (int)info[0], skip_proba = (int)info[1];
  const P0Plan& fp = P.plan[img];
  uint16_t* base = EMIT ? P.tokens + P.img_base[img] : nullptr;
  if (EMIT && idx == 0) p0_header(fp, P.proba + (size_t)img * 1056, P.proba0, P.update, num_skip, skip_proba, base);
  if (idx >= nmb) return;
  const int my = idx / P.mb_w, mx = idx - my * P.mb_w;
  const uint8_t* H = P.hdr + (size_t)img * nmb * 48;
  const uint8_t* h = H + (size_t)idx * 48;
  P0Sink bw{EMIT ? base + info[2] + P.mb_offset[(size_t)img * nmb + idx] : nullptr, 0};
  if (fp.seg_use && fp.seg_update_map) {
    const int id = P.segment[(size_t)img * nmb + idx];
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
  } else {
    bw.put(0, 145);
    // context modes: the bottom row of the macroblock above and the right column of the one to the left -- an I16 macroblock
    // stands for four copies of its mode, the frame edge for mode 0 (the writer's top / left arrays, encode_syntax.go:455-470)
    int top[4] = {0, 0, 0, 0}, lm[4] = {0, 0, 0, 0};
    if (my > 0) {
      const uint8_t* t = h - (size_t)P.mb_w * 48;
      for (int x = 0; x < 4; ++x) top[x] = t[0] ? t[8 + 12 + x] : t[1];
    }
    if (mx > 0) {
      const uint8_t* l = h - 48;
      for (int y = 0; y < 4; ++y) lm[y] = l[0] ? l[8 + 4 * y + 3] : l[1];
    }
    for (int y = 0; y < 4; ++y) {
      int ym = lm[y];
      for (int x = 0; x < 4; ++x) {
        const int mode = h[8 + y * 4 + x];
        const uint8_t* prob = P.bmodes + (top[x] * 10 + ym) * 9;
        const I4PathDev& pt = P.i4paths[mode];
        for (int k = 0; k < pt.n; ++k) bw.put(pt.bit[k], prob[pt.idx[k]]);
        ym = mode;
        top[x] = mode;
      }
    }
  }
  const int uv = h[2];
  if (uv == 0) bw.put(0, 142);
  else if (uv == 2) { bw.put(1, 142); bw.put(0, 114); }
  else if (uv == 3) { bw.put(1, 142); bw.put(1, 114); bw.put(0, 183); }
  else { bw.put(1, 142); bw.put(1, 114); bw.put(1, 183); }
  if (!EMIT) P.mb_tokens[(size_t)img * nmb + idx] = bw.cnt;
}

}  // namespace wg
