// SSE + SSIM over whole plane pairs (dsp.SSIM / SSIMGet / SSIMGetClipped / ssimCalculation, internal/dsp/ssim.go:12-160),
// separable form, sm_100a.  The 7x7 window weights are the outer product of (1,2,3,4,3,2,1) with itself and that triangle
// is a box of 4 convolved with a box of 4, so the five window statistics (sum x, sum y, sum xx, sum xy, sum yy) are two
// running box sums along a row and two down a column: ~13 additions per statistic and four outputs instead of 49
// multiply-adds per statistic and output (the direct form, metrics_kernel in round 1: 90 % of issue slots, 13-26 ms per
// 256 x 1536x1024).  Clipped windows need no special case: samples outside the picture are staged as zero (they add
// nothing to any sum) and the weight total N of a window is the product of its row and column weight totals.
//
// One CTA = a 32 x 56 tile of windows.  Phases (barrier between them):
//   stage   both planes' tile + 3-sample apron (4 on the left so that words stay aligned) as 32-bit words
//   hpass   task = (row, 4 adjacent columns): row sums of the five statistics, packed 16 B per sample into shared memory
//           (columns XOR-swizzled so that the 128-bit stores of this phase and the loads of the next are conflict-free);
//           the squared difference of the task's own samples goes to the SSE
//   vpass   thread = (column, 7 adjacent rows): column sums over 13 staged rows, then ssimCalculation per window
// The phase functions are host+device so that a CPU test harness can run them in this schedule (tests/test_oracle.py).
#pragma once
#include <stdint.h>
#include <stddef.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WG_SHD __host__ __device__ __forceinline__
#else
#define WG_SHD inline
#endif

namespace wg {

enum { SS_TW = 32, SS_TH = 56, SS_ROWS = SS_TH + 6, SS_SW = 10, SS_RPT = 7, SS_HTASKS = SS_ROWS * 8 };

struct alignas(16) SsimH { uint32_t xy, xx, xym, yy; };  // (sum x) | (sum y) << 16, sum xx, sum xy, sum yy of one row (16 B)

// ssimCalculation (ssim.go:48) on integer window statistics.  fnum / fden are products of two integers that a double holds
// exactly, so one rounded multiplication gives the same double as the reference's conversion of the 64-bit product.
WG_SHD double ssim_calc_sep(uint32_t xm, uint32_t ym, uint32_t xxm, uint32_t xym, uint32_t yym, uint32_t N) {
  const uint32_t w2 = N * N;  // N <= 256
  const unsigned long long C1 = 20ull * w2, C2 = 60ull * w2, C3 = 64ull * w2;
  const uint32_t xmxm = xm * xm, ymym = ym * ym;  // xm, ym <= 255 * 256
  const unsigned long long e = (unsigned long long)xmxm + ymym;
  if (e < C3) return 1.0;
  const uint32_t xmym = xm * ym;
  const uint32_t xyn = xym * N, xxn = xxm * N, yyn = yym * N;  // <= 255^2 * 256^2 < 2^32
  const uint32_t sxy_pos = xyn > xmym ? xyn - xmym : 0u;
  const uint32_t sxx = xxn - xmxm, syy = yyn - ymym;
  const unsigned long long num_s = (2ull * sxy_pos + C2) >> 8;
  const unsigned long long den_s = ((unsigned long long)sxx + syy + C2) >> 8;
  const double fnum = (double)(2ull * xmym + C1) * (double)num_s;
  const double fden = (double)(e + C1) * (double)den_s;
  if (fden == 0.0) return 1.0;
  return fnum / fden;
}

// weight total of the triangle (1,2,3,4,3,2,1) centred on pos over the samples inside [0, size)
WG_SHD uint32_t ssim_tri_weight(int pos, int size) {
  if (pos >= 3 && pos + 3 < size) return 16u;
  uint32_t w = 0;
  for (int d = -3; d <= 3; ++d)
    if (pos + d >= 0 && pos + d < size) w += (uint32_t)(4 - (d < 0 ? -d : d));
  return w;
}

// stage: word i of the tile (row i / SS_SW, picture columns x0 - 4 + 4 * (i % SS_SW) ...), zero outside the picture
WG_SHD void ssim_stage_word(const uint8_t* pa, const uint8_t* pb, int stride, int width, int height, int x0, int y0, bool aligned, int i,
                            uint32_t* sa, uint32_t* sb) {
  const int r = i / SS_SW, wc = i - r * SS_SW;
  const int gy = y0 + r - 3, gx = x0 + 4 * wc - 4;
  uint32_t va = 0, vb = 0;
  if (gy >= 0 && gy < height && gx >= 0 && gx < width) {
    const uint8_t* ra = pa + (size_t)gy * stride + gx;
    const uint8_t* rb = pb + (size_t)gy * stride + gx;
    if (aligned && gx + 4 <= width) {
      va = *reinterpret_cast<const uint32_t*>(ra);
      vb = *reinterpret_cast<const uint32_t*>(rb);
    } else {
      for (int k = 0; k < 4 && gx + k < width; ++k) { va |= (uint32_t)ra[k] << (8 * k); vb |= (uint32_t)rb[k] << (8 * k); }
    }
  }
  sa[i] = va; sb[i] = vb;
}

// sums of COUNT consecutive triangles over v[0 .. COUNT + 5]: box of 4, then box of 4 of the boxes
template <int COUNT>
WG_SHD void ssim_triangles(const uint32_t* v, uint32_t* o) {
  uint32_t b[COUNT + 3];
  b[0] = v[0] + v[1] + v[2] + v[3];
#pragma unroll
  for (int k = 1; k < COUNT + 3; ++k) b[k] = b[k - 1] + v[k + 3] - v[k - 1];
  o[0] = b[0] + b[1] + b[2] + b[3];
#pragma unroll
  for (int j = 1; j < COUNT; ++j) o[j] = o[j - 1] + b[j + 3] - b[j - 1];
}

// hpass: staged row (SS_SW words per plane), column group cg (tile columns 4 cg .. 4 cg + 3) -> four SsimH of that row;
// returns the squared difference of the group's own four samples
WG_SHD uint32_t ssim_hpass(const uint32_t* sa_row, const uint32_t* sb_row, int cg, SsimH* hrow) {
  uint32_t wa[3], wb[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { wa[k] = sa_row[cg + k]; wb[k] = sb_row[cg + k]; }
  // tile column c sits at staged byte c + 4; its window spans staged bytes c + 1 .. c + 7: bytes 1 .. 10 of the three words
  uint32_t x[10], y[10], xx[10], xy[10], yy[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    x[i] = (wa[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xffu;
    y[i] = (wb[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xffu;
    xx[i] = x[i] * x[i]; xy[i] = x[i] * y[i]; yy[i] = y[i] * y[i];
  }
  uint32_t ox[4], oy[4], oxx[4], oxy[4], oyy[4];
  ssim_triangles<4>(x, ox); ssim_triangles<4>(y, oy); ssim_triangles<4>(xx, oxx); ssim_triangles<4>(xy, oxy); ssim_triangles<4>(yy, oyy);
  uint32_t sse = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * cg + j;
    SsimH h;
    h.xy = ox[j] | (oy[j] << 16); h.xx = oxx[j]; h.xym = oxy[j]; h.yy = oyy[j];
    hrow[c ^ (c >> 3)] = h;
    sse += xx[j + 3] + yy[j + 3] - 2u * xy[j + 3];  // (x - y)^2 of tile column c
  }
  return sse;
}

// vpass: column c, row group g (tile rows 7 g .. 7 g + 6): sum of ssimCalculation over the group's windows inside the picture
WG_SHD double ssim_vpass(const SsimH* h /* [SS_ROWS][SS_TW] */, int c, int g, int x0, int y0, int width, int height) {
  if (x0 + c >= width || y0 + SS_RPT * g >= height) return 0.0;
  uint32_t vx[SS_RPT + 6], vy[SS_RPT + 6], vxx[SS_RPT + 6], vxy[SS_RPT + 6], vyy[SS_RPT + 6];
  const int cs = c ^ (c >> 3);
#pragma unroll
  for (int k = 0; k < SS_RPT + 6; ++k) {
    const SsimH e = h[(SS_RPT * g + k) * SS_TW + cs];
    vx[k] = e.xy & 0xffffu; vy[k] = e.xy >> 16; vxx[k] = e.xx; vxy[k] = e.xym; vyy[k] = e.yy;
  }
  uint32_t xm[SS_RPT], ym[SS_RPT], xxm[SS_RPT], xym[SS_RPT], yym[SS_RPT];
  ssim_triangles<SS_RPT>(vx, xm); ssim_triangles<SS_RPT>(vy, ym); ssim_triangles<SS_RPT>(vxx, xxm); ssim_triangles<SS_RPT>(vxy, xym);
  ssim_triangles<SS_RPT>(vyy, yym);
  const uint32_t wx = ssim_tri_weight(x0 + c, width);
  double sum = 0.0;
#pragma unroll
  for (int m = 0; m < SS_RPT; ++m) {
    const int gy = y0 + SS_RPT * g + m;
    if (gy < height) sum += ssim_calc_sep(xm[m], ym[m], xxm[m], xym[m], yym[m], wx * ssim_tri_weight(gy, height));
  }
  return sum;
}

#ifdef __CUDACC__
struct SsimSepParams {
  const uint8_t* a; const uint8_t* b;
  size_t plane_stride;
  int stride, width, height, n;
  int tiles_x, tiles_y;
  unsigned long long* sse_part;  // [n][tiles]
  double* ssim_part;             // [n][tiles]
};
__global__ void __launch_bounds__(256, 3) ssim_sep_kernel(const SsimSepParams P) {
  __shared__ __align__(16) uint32_t s_a[SS_ROWS * SS_SW];
  __shared__ __align__(16) uint32_t s_b[SS_ROWS * SS_SW];
  __shared__ __align__(16) SsimH s_h[SS_ROWS * SS_TW];
  __shared__ unsigned long long s_sse[8];
  __shared__ double s_ssim[8];
  const int tiles = P.tiles_x * P.tiles_y;
  const int img = blockIdx.x / tiles, tile = blockIdx.x - img * tiles;
  const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
  const int x0 = tx * SS_TW, y0 = ty * SS_TH;
  const uint8_t* pa = P.a + (size_t)img * P.plane_stride;
  const uint8_t* pb = P.b + (size_t)img * P.plane_stride;
  const bool aligned = ((reinterpret_cast<uintptr_t>(pa) | reinterpret_cast<uintptr_t>(pb) | (uintptr_t)P.stride) & 3u) == 0;
  for (int i = threadIdx.x; i < SS_ROWS * SS_SW; i += 256) ssim_stage_word(pa, pb, P.stride, P.width, P.height, x0, y0, aligned, i, s_a, s_b);
  __syncthreads();
  unsigned long long sse = 0;
  for (int t = threadIdx.x; t < SS_HTASKS; t += 256) {
    const int r = t >> 3, cg = t & 7;
    const uint32_t d = ssim_hpass(s_a + r * SS_SW, s_b + r * SS_SW, cg, s_h + r * SS_TW);
    if (r >= 3 && r < 3 + SS_TH) sse += d;  // the tile's own rows; samples outside the picture are 0 in both planes
  }
  __syncthreads();
  double ssim = ssim_vpass(s_h, threadIdx.x & 31, threadIdx.x >> 5, x0, y0, P.width, P.height);
  // deterministic in-CTA reduction: lanes, then warps in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sse += __shfl_down_sync(0xffffffffu, sse, o);
    ssim += __shfl_down_sync(0xffffffffu, ssim, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sse[threadIdx.x >> 5] = sse; s_ssim[threadIdx.x >> 5] = ssim; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long ts = 0;
    double tq = 0.0;
    for (int i = 0; i < 8; ++i) { ts += s_sse[i]; tq += s_ssim[i]; }
    P.sse_part[blockIdx.x] = ts;
    P.ssim_part[blockIdx.x] = tq;
  }
}
#endif

}  // namespace wg
