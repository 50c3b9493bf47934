// Plane metrics (SSE / SSIM, internal/dsp/ssim.go:12-181) and the batched per-block dsp operator surface
// (internal/dsp/dsp.go:12-37 function variables and their *Direct twins), sm_100a.
#pragma once
#include "dec_kernels.cuh"

namespace wg {

// ---- SSE + SSIM over whole plane pairs.  One CTA = a 32x8 tile of windows; both planes' tiles (with the
// 3-pixel apron) are staged in shared memory so each source byte is read from HBM once (2 B/px algorithmic).
struct MetricsParams {
  const uint8_t* a; const uint8_t* b;
  size_t plane_stride;
  int stride, width, height, n;
  int tiles_x, tiles_y;
  unsigned long long* sse_part;  // [n][tiles]
  double* ssim_part;             // [n][tiles]
};
// ssimCalculation (ssim.go:48) on integer window statistics; the only floating-point op is the final divide.
__device__ __forceinline__ double ssim_calc(uint32_t xm, uint32_t ym, uint32_t xxm, uint32_t xym, uint32_t yym, uint32_t N) {
  const unsigned long long w2 = (unsigned long long)N * N;
  const unsigned long long C1 = 20 * w2, C2 = 60 * w2, C3 = 64 * w2;
  const unsigned long long xmxm = (unsigned long long)xm * xm, ymym = (unsigned long long)ym * ym;
  if (xmxm + ymym < C3) return 1.0;
  const long long xmym = (long long)xm * (long long)ym;
  const long long sxy = (long long)xym * (long long)N - xmym;
  const unsigned long long sxx = (unsigned long long)xxm * N - xmxm;
  const unsigned long long syy = (unsigned long long)yym * N - ymym;
  const unsigned long long sxy_pos = sxy > 0 ? (unsigned long long)sxy : 0ull;
  const unsigned long long num_s = (2 * sxy_pos + C2) >> 8;
  const unsigned long long den_s = (sxx + syy + C2) >> 8;
  const unsigned long long fnum = (2 * (unsigned long long)xmym + C1) * num_s;
  const unsigned long long fden = (xmxm + ymym + C1) * den_s;
  if (fden == 0) return 1.0;
  return (double)fnum / (double)fden;
}
__global__ void __launch_bounds__(256) metrics_kernel(const MetricsParams P) {
  constexpr int TW = 32, TH = 8, AW = TW + 6, AH = TH + 6;
  __shared__ uint8_t sa[AH][AW + 2], sb[AH][AW + 2];
  __shared__ unsigned long long s_sse[8];
  __shared__ double s_ssim[8];
  const int tiles = P.tiles_x * P.tiles_y;
  const int img = blockIdx.x / tiles, tile = blockIdx.x - img * tiles;
  const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
  const int x0 = tx * TW, y0 = ty * TH;
  const uint8_t* pa = P.a + (size_t)img * P.plane_stride;
  const uint8_t* pb = P.b + (size_t)img * P.plane_stride;
  for (int i = threadIdx.x; i < AW * AH; i += 256) {
    const int r = i / AW, c = i - r * AW;
    const int gx = x0 + c - 3, gy = y0 + r - 3;
    uint8_t va = 0, vb = 0;
    if (gx >= 0 && gx < P.width && gy >= 0 && gy < P.height) {
      va = pa[(size_t)gy * P.stride + gx];
      vb = pb[(size_t)gy * P.stride + gx];
    }
    sa[r][c] = va;
    sb[r][c] = vb;
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int gx = x0 + lx, gy = y0 + ly;
  unsigned long long sse = 0;
  double ssim = 0.0;
  if (gx < P.width && gy < P.height) {
    const int d = (int)sa[ly + 3][lx + 3] - (int)sb[ly + 3][lx + 3];
    sse = (unsigned long long)(d * d);
    uint32_t w = 0, xm = 0, ym = 0, xxm = 0, xym = 0, yym = 0;
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
      const int yy = gy + dy - 3;
      if (yy < 0 || yy >= P.height) continue;
      const uint32_t wy = dy < 4 ? dy + 1 : 7 - dy;
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const int xx = gx + dx - 3;
        if (xx < 0 || xx >= P.width) continue;
        const uint32_t ww = wy * (dx < 4 ? dx + 1 : 7 - dx);
        const uint32_t x = sa[ly + dy][lx + dx], y = sb[ly + dy][lx + dx];
        w += ww; xm += ww * x; ym += ww * y; xxm += ww * x * x; xym += ww * x * y; yym += ww * y * y;
      }
    }
    ssim = ssim_calc(xm, ym, xxm, xym, yym, w);  // interior windows have w == 256 (SSIMGet), borders are SSIMGetClipped
  }
  // deterministic in-CTA reduction: lanes, then warps in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sse += __shfl_down_sync(0xffffffffu, sse, o);
    ssim += __shfl_down_sync(0xffffffffu, ssim, o);
  }
  if (lx == 0) { s_sse[ly] = sse; s_ssim[ly] = ssim; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long ts = 0;
    double tq = 0.0;
    for (int i = 0; i < 8; ++i) { ts += s_sse[i]; tq += s_ssim[i]; }
    P.sse_part[blockIdx.x] = ts;
    P.ssim_part[blockIdx.x] = tq;
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
#undef WG_TID

}  // namespace wg
