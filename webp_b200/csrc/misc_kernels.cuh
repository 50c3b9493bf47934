// Plane metrics (SSE / SSIM, internal/dsp/ssim.go:12-181) and the batched per-block dsp operator surface
// (internal/dsp/dsp.go:12-37 function variables and their *Direct twins), sm_100a.
#pragma once
#include "dec_kernels.cuh"

namespace wg {

// ---- SSE over whole plane pairs (the SSE + SSIM kernel is ssim_sep.cuh; metrics_reduce_kernel below adds its tile partials).
struct MetricsParams {
  const uint8_t* a; const uint8_t* b;
  size_t plane_stride;
  int stride, width, height, n;
  int tiles_x, tiles_y;
  unsigned long long* sse_part;  // [n][tiles]
  double* ssim_part;             // [n][tiles]
};
// SSE alone (dsp.SSE, ssim.go:172; PSNR for TargetPSNR): a pure stream, 2 B/px.  16 bytes of each plane per thread and step
// (128-bit loads where the row allows), |a - b| per byte and the dot product of the differences with themselves (dp4a).
__global__ void __launch_bounds__(256) sse_kernel(const MetricsParams P, unsigned long long* sse) {
  __shared__ unsigned long long s_part[8];
  const int chunks = (P.width + 15) >> 4;
  const long long per_img = (long long)chunks * P.height;
  const int img = blockIdx.y;
  const uint8_t* pa = P.a + (size_t)img * P.plane_stride;
  const uint8_t* pb = P.b + (size_t)img * P.plane_stride;
  unsigned long long acc = 0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < per_img; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / chunks), c = (int)(t - (long long)r * chunks) * 16;
    const uint8_t* ra = pa + (size_t)r * P.stride + c;
    const uint8_t* rb = pb + (size_t)r * P.stride + c;
    uint32_t s = 0;
    if (c + 16 <= P.width && (((uintptr_t)ra | (uintptr_t)rb) & 15) == 0) {
      const uint4 va = *reinterpret_cast<const uint4*>(ra), vb = *reinterpret_cast<const uint4*>(rb);
      const uint32_t d0 = __vabsdiffu4(va.x, vb.x), d1 = __vabsdiffu4(va.y, vb.y), d2 = __vabsdiffu4(va.z, vb.z), d3 = __vabsdiffu4(va.w, vb.w);
      s = __dp4a(d0, d0, s); s = __dp4a(d1, d1, s); s = __dp4a(d2, d2, s); s = __dp4a(d3, d3, s);
    } else {
      for (int k = 0; k < 16 && c + k < P.width; ++k) { const int d = (int)ra[k] - (int)rb[k]; s += (uint32_t)(d * d); }
    }
    acc += s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tsum = 0;
    for (int i = 0; i < 8; ++i) tsum += s_part[i];
    if (tsum) atomicAdd(sse + img, tsum);  // integer sum: the order of the CTAs does not matter
  }
}
// one CTA per image: fixed-order reduction of the tile partials
__global__ void __launch_bounds__(256) metrics_reduce_kernel(const unsigned long long* sse_part, const double* ssim_part,
                                                             int tiles, unsigned long long* sse, double* ssim) {
  __shared__ unsigned long long s_s[256];
  __shared__ double s_q[256];
  const int img = blockIdx.x;
  unsigned long long a = 0;
  double q = 0.0;
  for (int i = threadIdx.x; i < tiles; i += 256) { a += sse_part[(size_t)img * tiles + i]; q += ssim_part[(size_t)img * tiles + i]; }
  s_s[threadIdx.x] = a; s_q[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s_s[threadIdx.x] += s_s[threadIdx.x + o]; s_q[threadIdx.x] += s_q[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { sse[img] = s_s[0]; ssim[img] = s_q[0]; }
}

// ---- cleanupTransparentAreaLossy (encode.go:788-890) on NRGBA images, in place.  One CTA per (image, row of 8x8 blocks), a
// warp per block (two pixels per lane): transparent pixels of a partly transparent block take the average colour of its
// opaque pixels; fully transparent blocks are flattened to the colour of the first pixel of their run in the block row
// (the reference carries it from block to block: a run starts after any block that is not fully transparent).  The right
// remainder and the bottom remainder rows are smoothened only.
struct CleanupParams {
  uint8_t* px;           // [n][height][stride] NRGBA
  size_t image_stride;
  int stride, n, width, height;
};
__global__ void __launch_bounds__(128) cleanup_transparent_kernel(const CleanupParams P) {
  extern __shared__ uint8_t s_transp[];  // [full blocks of the row] 1 = fully transparent
  const int rows_full = P.height / 8, rem_h = P.height % 8;
  const int block_rows = rows_full + (rem_h ? 1 : 0);
  const int img = blockIdx.x / block_rows, brow = blockIdx.x % block_rows;
  const int by = brow * 8, bh = brow < rows_full ? 8 : rem_h;
  const int nfull = P.width / 8, rem_w = P.width % 8, nblk = nfull + (rem_w ? 1 : 0);
  uint8_t* base = P.px + (size_t)img * P.image_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < nblk; b += 4) {
    const int bx = b * 8, bw = b < nfull ? 8 : rem_w;
    int cnt = 0, sr = 0, sg = 0, sb = 0;
    uint32_t pix[2];
    bool in[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = lane + 32 * k, x = i & 7, y = i >> 3;
      in[k] = x < bw && y < bh;
      pix[k] = in[k] ? *reinterpret_cast<const uint32_t*>(base + (size_t)(by + y) * P.stride + 4 * (bx + x)) : 0u;
      if (in[k] && (pix[k] >> 24) != 0) { cnt++; sr += pix[k] & 0xff; sg += (pix[k] >> 8) & 0xff; sb += (pix[k] >> 16) & 0xff; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o); sr += __shfl_xor_sync(0xffffffffu, sr, o);
      sg += __shfl_xor_sync(0xffffffffu, sg, o); sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if (cnt > 0 && cnt < bw * bh) {
      const uint32_t avg = (uint32_t)(sr / cnt) | ((uint32_t)(sg / cnt) << 8) | ((uint32_t)(sb / cnt) << 16);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = lane + 32 * k, x = i & 7, y = i >> 3;
        if (in[k] && (pix[k] >> 24) == 0) *reinterpret_cast<uint32_t*>(base + (size_t)(by + y) * P.stride + 4 * (bx + x)) = avg;
      }
    }
    if (lane == 0 && b < nfull) s_transp[b] = (cnt == 0);
  }
  __syncthreads();
  if (bh < 8) return;  // bottom remainder: smoothing only
  for (int b = warp; b < nfull; b += 4) {
    if (!s_transp[b]) continue;
    int start = b;
    while (start > 0 && s_transp[start - 1]) --start;
    const uint32_t c = *reinterpret_cast<const uint32_t*>(base + (size_t)by * P.stride + 4 * (start * 8)) & 0x00ffffffu;  // untouched so far
    // (that pixel is transparent already, so the warp flattening the run's first block rewrites it with the same value)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = lane + 32 * k, x = i & 7, y = i >> 3;
      *reinterpret_cast<uint32_t*>(base + (size_t)(by + y) * P.stride + 4 * (b * 8 + x)) = c;
    }
  }
}

// ---- batched dsp surface: one thread per 4x4 block, dense 16-byte tiles ----------------------------------
__device__ __forceinline__ void load16u8(const uint8_t* p, int* d) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { d[4 * j] = w[j] & 0xff; d[4 * j + 1] = (w[j] >> 8) & 0xff; d[4 * j + 2] = (w[j] >> 16) & 0xff; d[4 * j + 3] = w[j] >> 24; }
}
__device__ __forceinline__ void store16u8(uint8_t* p, const int* d) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = (uint32_t)d[4 * j] | ((uint32_t)d[4 * j + 1] << 8) | ((uint32_t)d[4 * j + 2] << 16) | ((uint32_t)d[4 * j + 3] << 24);
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void load16s16(const int16_t* p, int* d) {
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i] = p[i];
}
__device__ __forceinline__ void store16s16(int16_t* p, const int* d) {
#pragma unroll
  for (int i = 0; i < 16; ++i) p[i] = (int16_t)d[i];
}
#define WG_TID int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return
__global__ void dsp_ftransform_kernel(int n, const uint8_t* src, const uint8_t* ref, int16_t* out) {
  WG_TID; int s[16], r[16], c[16]; load16u8(src + 16 * (size_t)i, s); load16u8(ref + 16 * (size_t)i, r); ftransform(s, r, c); store16s16(out + 16 * (size_t)i, c);
}
__global__ void dsp_itransform_kernel(int n, const uint8_t* ref, const int16_t* in, uint8_t* dst) {
  WG_TID; int r[16], c[16], d[16]; load16u8(ref + 16 * (size_t)i, r); load16s16(in + 16 * (size_t)i, c); itransform(r, c, d); store16u8(dst + 16 * (size_t)i, d);
}
__global__ void dsp_fwht_kernel(int n, const int16_t* in, int16_t* out) {
  WG_TID; int a[16], b[16]; load16s16(in + 16 * (size_t)i, a); fwht(a, b); store16s16(out + 16 * (size_t)i, b);
}
__global__ void dsp_iwht_kernel(int n, const int16_t* in, int16_t* out) {
  WG_TID; int a[16], b[16]; load16s16(in + 16 * (size_t)i, a); iwht(a, b); store16s16(out + 16 * (size_t)i, b);
}
__global__ void dsp_sse4x4_kernel(int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  WG_TID; int x[16], y[16]; load16u8(a + 16 * (size_t)i, x); load16u8(b + 16 * (size_t)i, y); out[i] = sse16(x, y);
}
__global__ void dsp_tdisto4x4_kernel(int n, const uint8_t* a, const uint8_t* b, int32_t* out) {
  WG_TID; int x[16], y[16]; load16u8(a + 16 * (size_t)i, x); load16u8(b + 16 * (size_t)i, y); out[i] = tdisto4x4(x, y);
}
__global__ void dsp_pred4_kernel(int n, const uint8_t* ctx13, uint8_t* out) {  // n = blocks * 10 (one thread per block x mode)
  WG_TID;
  const int blk = i / 10, mode = i - blk * 10;
  int e[13], d[16];
#pragma unroll
  for (int k = 0; k < 13; ++k) e[k] = ctx13[13 * (size_t)blk + k];
  pred4(mode, e, d);
  store16u8(out + 16 * (size_t)i, d);
}
// one 8-lane group per (block, mode): stages the context into a BPS-strided work buffer and runs the cooperative predictor
__global__ void __launch_bounds__(128) dsp_pred_square_kernel(int n_tasks, int size, const uint8_t* ctx_px, uint8_t* out) {
  __shared__ __align__(16) uint8_t s_buf[16][18 * BPS];
  const int grp = threadIdx.x >> 3, gl = threadIdx.x & 7;
  const int task = blockIdx.x * 16 + grp;
  const bool active = task < n_tasks;
  const int blk = active ? task / 7 : 0, mode = active ? task - blk * 7 : 0;
  uint8_t* buf = s_buf[grp];
  const int off = BPS + 8, cs = 1 + 2 * size;
  const uint8_t* c = ctx_px + (size_t)blk * cs;
  if (active) {
    if (gl == 0) buf[off - BPS - 1] = c[0];
    for (int i = gl; i < size; i += 8) { buf[off - BPS + i] = c[1 + i]; buf[off - 1 + i * BPS] = c[1 + size + i]; }
  }
  __syncwarp();
  if (active) pred_square_coop<8>(gl, mode, buf, off, size);
  __syncwarp();
  if (active)
    for (int i = gl; i < size * size; i += 8) out[(size_t)task * size * size + i] = buf[off + (i / size) * BPS + (i % size)];
}
__global__ void dsp_quantize_kernel(int n, const int16_t* in, SegQuant sq, int first, int16_t* out, int32_t* nz) {
  WG_TID; int c[16], q[16]; load16s16(in + 16 * (size_t)i, c); nz[i] = quantize_block(c, q, sq, first); store16s16(out + 16 * (size_t)i, q);
}
struct TabPtrs { const uint16_t* lc; const uint16_t* eob; const uint16_t* lfc; };
__global__ void dsp_trellis_kernel(int n, const int16_t* in, SegQuant sq, int first, int ctx_type, const int32_t* ctx0, int lambda,
                                   TabPtrs tp, int16_t* out, int32_t* nz) {
  __shared__ int16_t s_io[128][16];  // the trellis works in place in shared memory (zigzag-indexed access)
  WG_TID;
  CostTabs T; T.lc = tp.lc; T.eob = tp.eob; T.lfc = tp.lfc; T.lfc_hi = tp.lfc;
  int16_t* io = s_io[threadIdx.x];
#pragma unroll
  for (int k = 0; k < 16; ++k) io[k] = in[16 * (size_t)i + k];
  // rate * lambda stays a 32-bit product in the latency-oriented trellis (encoder lambdas are <= 21567)
  nz[i] = lambda <= 50000 ? trellis_block_v3(io, sq, first, ctx_type, ctx0[i], lambda, T) : trellis_block_smem(io, sq, first, ctx_type, ctx0[i], lambda, T);
#pragma unroll
  for (int k = 0; k < 16; ++k) out[16 * (size_t)i + k] = io[k];
}
__global__ void dsp_token_cost_kernel(int n, const int16_t* levels, const int32_t* nzc, int ctx_type, const int32_t* ctx0, int first,
                                      TabPtrs tp, int32_t* out) {
  WG_TID;
  CostTabs T; T.lc = tp.lc; T.eob = tp.eob; T.lfc = tp.lfc; T.lfc_hi = tp.lfc;
  int q[16]; load16s16(levels + 16 * (size_t)i, q);
  out[i] = token_cost(q, nzc[i], ctx_type, ctx0[i], first, T);
}
// ---- the rest of the per-block operator surface, batched: 16x16 SSE / TDisto, DequantCoeffs, the decoder's
// short-cut transforms, the loop-filter set on 24x24 tiles, UpsampleLinePair (RGB and NRGBA)
__global__ void dsp_sse16x16_kernel(int n, const uint8_t* a, const uint8_t* b, int32_t* out) {  // ssim.go:220
  WG_TID;
  const uint8_t* pa = a + 256 * (size_t)i; const uint8_t* pb = b + 256 * (size_t)i;
  int s = 0;
  for (int k = 0; k < 256; ++k) { const int d = (int)pa[k] - (int)pb[k]; s += d * d; }
  out[i] = s;
}
__global__ void dsp_tdisto16x16_kernel(int n, const uint8_t* a, const uint8_t* b, int32_t* out) {  // tDisto16x16Go (ssim.go:327)
  WG_TID;
  const uint8_t* pa = a + 256 * (size_t)i; const uint8_t* pb = b + 256 * (size_t)i;
  int d = 0;
  for (int blk = 0; blk < 16; ++blk) {
    int x[16], y[16];
    for (int k = 0; k < 16; ++k) {
      const int off = ((blk >> 2) * 4 + (k >> 2)) * 16 + (blk & 3) * 4 + (k & 3);
      x[k] = pa[off]; y[k] = pb[off];
    }
    d += tdisto4x4(x, y);
  }
  out[i] = d;
}
__global__ void dsp_dequant_kernel(int n, const int16_t* in, SegQuant sq, int16_t* out) {  // dequantCoeffsGo (encode_quant.go:81)
  WG_TID; int q[16], dq[16]; load16s16(in + 16 * (size_t)i, q); dequant_block(q, dq, sq); store16s16(out + 16 * (size_t)i, dq);
}
// decoder transforms on one 4x4 block (transforms.go:37-216): kind 0 transformOne, 1 transformDC ((dc + 4) >> 3 everywhere),
// 2 transformAC3 (only in[0], in[1], in[4]); kinds 3 / 4 = transformUV / transformDCUV: four blocks of an 8x8 tile
__device__ __forceinline__ void dec_transform_block(int kind, const int* in, const int* ref, int* dst) {
  if (kind == 1) {
    const int add = (in[0] + 4) >> 3;
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[k] = clip8(ref[k] + add);
  } else if (kind == 2) {
    const int a = in[0] + 4, c4 = mul2(in[4]), d4 = mul1(in[4]), c1 = mul2(in[1]), d1 = mul1(in[1]);
    const int rowv[4] = {a + d4, a + c4, a - c4, a - d4};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dst[4 * j + 0] = clip8(ref[4 * j + 0] + ((rowv[j] + d1) >> 3));
      dst[4 * j + 1] = clip8(ref[4 * j + 1] + ((rowv[j] + c1) >> 3));
      dst[4 * j + 2] = clip8(ref[4 * j + 2] + ((rowv[j] - c1) >> 3));
      dst[4 * j + 3] = clip8(ref[4 * j + 3] + ((rowv[j] - d1) >> 3));
    }
  } else {
    itransform(ref, in, dst);
  }
}
__global__ void dsp_dec_transform_kernel(int n, int kind, const int16_t* in, const uint8_t* ref, uint8_t* dst) {
  WG_TID;
  if (kind < 3) {
    int c[16], r[16], d[16];
    load16s16(in + 16 * (size_t)i, c); load16u8(ref + 16 * (size_t)i, r);
    dec_transform_block(kind, c, r, d);
    store16u8(dst + 16 * (size_t)i, d);
    return;
  }
  for (int blk = 0; blk < 4; ++blk) {  // 8x8 tile (stride 8), coefficients [4][16]
    int c[16], r[16], d[16];
    load16s16(in + 64 * (size_t)i + 16 * blk, c);
    const size_t o = 64 * (size_t)i + (blk >> 1) * 32 + (blk & 1) * 4;
    for (int k = 0; k < 16; ++k) r[k] = ref[o + (k >> 2) * 8 + (k & 3)];
    dec_transform_block(kind == 3 ? 0 : 1, c, r, d);
    for (int k = 0; k < 16; ++k) dst[o + (k >> 2) * 8 + (k & 3)] = (uint8_t)d[k];
  }
}
// loop-filter set (filter.go:93-242) on 24x24 tiles: the 16x16 (or 8x8) block sits at (4, 4), so edge 0 has its four
// samples of context.  kind: 0 SimpleVFilter16, 1 SimpleHFilter16, 2 SimpleVFilter16i, 3 SimpleHFilter16i, 4 VFilter16,
// 5 HFilter16, 6 VFilter16i, 7 HFilter16i, 8 VFilter8, 9 HFilter8, 10 VFilter8i, 11 HFilter8i (one chroma plane per tile).
// One thread = one sample position along the edge; the inner edges of a position are walked in order as the reference does.
__global__ void dsp_filter_kernel(int n_tasks, int kind, uint8_t* tiles, int thresh, int ithresh, int hev_t) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  const int size = kind >= 8 ? 8 : 16;
  const int tile = t / size, i = t - tile * size;
  uint8_t* base = tiles + (size_t)tile * 576 + 4 * 24 + 4;
  const bool vertical = (kind & 1) == 0;                 // V filters cross a horizontal edge: step across = one row
  const int across = vertical ? 24 : 1, along = vertical ? 1 : 24;
  const bool simple = kind < 4, inner = (kind & 2) != 0;
  const int k0 = inner ? 1 : 0, k1 = inner ? (size == 16 ? 3 : 1) : 0;
  for (int k = k0; k <= k1; ++k) {
    uint8_t* p = base + i * along + k * 4 * across;
    if (simple) simple_edge(p, across, thresh);
    else complex_edge(p, across, thresh, ithresh, hev_t, !inner);
  }
}
// UpsampleLinePair / UpsampleLinePairNRGBA (upsample.go:45,130) on n independent line pairs of `width` pixels: one thread per
// output column, both rows.  channels = 3 (RGB) or 4 (NRGBA, alpha rows optional).  bot_y may be null (last row of an odd height).
__global__ void dsp_upsample_pair_kernel(int n, int width, const uint8_t* top_y, const uint8_t* bot_y, const uint8_t* top_u, const uint8_t* top_v,
                                         const uint8_t* bot_u, const uint8_t* bot_v, const uint8_t* alpha_top, const uint8_t* alpha_bot, int channels,
                                         uint8_t* top_dst, uint8_t* bot_dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * width) return;
  const int pair = (int)(t / width), x = (int)(t - (long long)pair * width);
  const int cw = (width + 1) >> 1;
  const uint8_t* tu = top_u + (size_t)pair * cw; const uint8_t* tv = top_v + (size_t)pair * cw;
  const uint8_t* bu = bot_u + (size_t)pair * cw; const uint8_t* bv = bot_v + (size_t)pair * cw;
  auto ld = [](const uint8_t* u, const uint8_t* v, int k) { return (uint32_t)u[k] | ((uint32_t)v[k] << 16); };
  uint32_t uvt, uvb;
  const int last_pair = (width - 1) >> 1;
  if (x == 0) {
    const uint32_t tl = ld(tu, tv, 0), l = ld(bu, bv, 0);
    uvt = (3 * tl + l + 0x00020002u) >> 2; uvb = (3 * l + tl + 0x00020002u) >> 2;
  } else if (((x + 1) >> 1) > last_pair) {  // last pixel of an even width
    const uint32_t tl = ld(tu, tv, last_pair), l = ld(bu, bv, last_pair);
    uvt = (3 * tl + l + 0x00020002u) >> 2; uvb = (3 * l + tl + 0x00020002u) >> 2;
  } else {
    const int k = (x + 1) >> 1;
    const uint32_t tl = ld(tu, tv, k - 1), tt = ld(tu, tv, k), l = ld(bu, bv, k - 1), cur = ld(bu, bv, k);
    const uint32_t avg = tl + tt + l + cur + 0x00080008u;
    const uint32_t diag12 = (avg + 2 * (tt + l)) >> 3, diag03 = (avg + 2 * (tl + cur)) >> 3;
    if (x & 1) { uvt = (diag12 + tl) >> 1; uvb = (diag03 + l) >> 1; } else { uvt = (diag03 + tt) >> 1; uvb = (diag12 + cur) >> 1; }
  }
  const size_t o = ((size_t)pair * width + x);
  const uint32_t pt = yuv_to_rgba(top_y[o], uvt & 0xff, (uvt >> 16) & 0xff, alpha_top ? alpha_top[o] : 255);
  for (int c = 0; c < channels; ++c) top_dst[o * channels + c] = (uint8_t)(pt >> (8 * c));
  if (bot_y) {
    const uint32_t pb = yuv_to_rgba(bot_y[o], uvb & 0xff, (uvb >> 16) & 0xff, alpha_bot ? alpha_bot[o] : 255);
    for (int c = 0; c < channels; ++c) bot_dst[o * channels + c] = (uint8_t)(pb >> (8 * c));
  }
}
#undef WG_TID

}  // namespace wg
