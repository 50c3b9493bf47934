// SharpYUV import (EncoderOptions.UseSharpYUV): RGB -> YUV420 planes by iterative refinement of a luma plane W and three
// half-resolution residual planes (R-W, G-W, B-W), then the WebP matrix.  Mirrors
//   convertSharp / importOneRow / storeGray / updateW / updateChroma   sharpyuv/sharpyuv.go:190-361
//   interpolateTwoRows / sharpYUVUpdateY / sharpYUVUpdateRGB           sharpyuv/sharpyuv.go:363-389
//   convertWRGBToYUV                                                   sharpyuv/sharpyuv.go:391-431
//   gamma tables (sRGB), fixed-point interpolation                     sharpyuv/gamma.go:47-131
//   importYCbCr (edge replication up to the macroblock grid)           internal/lossy/encode.go:544-585
// Three kernels: sharp_init (one thread per 2x2 block, all images), sharp_refine (one CTA per image: the refinement sweeps a
// frame top to bottom and every row pair reads the chroma row above AFTER its update, so row pairs are a serial chain; the
// samples of one row pair are independent and go across the CTA's threads), sharp_finish (one thread per padded 2x2 block).
// sharp_refine_ring_kernel is the same sweep with the step's independent operands fetched one row pair ahead and the residual
// rows in a shared-memory ring: the one the library launches (11.2 vs 17.6 ms per 256 x 1536x1024); sharp_refine_kernel<1024>
// stays for rows wider than the ring.
// The per-sample code is host+device so that a CPU test harness can run the same functions in the kernels' schedule.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#ifdef __CUDACC__
#define WG_SHD __host__ __device__ __forceinline__
#else
#define WG_SHD inline
#endif

namespace wg {

struct SharpParams {
  const uint8_t* rgba; size_t image_stride; int stride;  // RGBA rows, alpha ignored (encode.go:1190-1215 drops it)
  int n, width, height;      // picture size
  int w, h, uv_w, uv_h;      // rounded up to even / halved (sharpyuv.go:193-202)
  uint16_t* best_y; uint16_t* target_y;  // [n][h][w]          10-bit samples (8-bit input << 2)
  int16_t* best_uv; int16_t* target_uv;  // [n][uv_h][3][uv_w] residuals R-W, G-W, B-W, channel rows interleaved per line
  const uint32_t* g2l;  // [1026] gammaToLinearTab
  const uint32_t* l2g;  // [514]  linearToGammaTab
  uint8_t* y; uint8_t* u; uint8_t* v; size_t y_plane, uv_plane; int pad_w, pad_h;  // encoder source planes
  int* iterations;  // [n] refinement passes run (test tap; may be null)
};

enum { kSharpMax = 1023, kSharpG2L = 1026, kSharpL2G = 514 };
// gammaToLinearTab then linearToGammaTab (sharpyuv/gamma.go:47-91), in doubles as the reference builds them.  Host only.
inline void sharp_build_tables(uint32_t* tab /* [kSharpG2L + kSharpL2G] */) {
  const double a = 0.09929682680944, thresh = 0.018053968510807, final_scale = 65536.0, gamma_f = 1.0 / 0.45;
  for (int v = 0; v <= 1024; ++v) {
    const double g = (1.0 / 1024.0) * v;
    const double value = g <= thresh * 4.5 ? g / 4.5 : pow((1.0 / (1.0 + a)) * (g + a), gamma_f);
    tab[v] = (uint32_t)(value * final_scale + 0.5);
  }
  tab[1025] = tab[1024];
  for (int v = 0; v <= 512; ++v) {
    const double g = (1.0 / 512.0) * v;
    const double value = g <= thresh ? 4.5 * g : (1.0 + a) * pow(g, 1.0 / gamma_f) - a;
    tab[kSharpG2L + v] = (uint32_t)(final_scale * value + 0.5);
  }
  tab[kSharpG2L + 513] = tab[kSharpG2L + 512];
}
WG_SHD uint32_t sharp_l2g(uint32_t value, const uint32_t* l2g) {  // fromLinearSrgb at 10 bits: table step 128, values >> 6
  const uint32_t pos = value >> 7, x = value & 127;
  const uint32_t v0 = l2g[pos] >> 6, v1 = l2g[pos + 1] >> 6;
  return (v0 + (((v1 - v0) * x + 64) >> 7)) & 0xffff;
}
WG_SHD int sharp_gray(int r, int g, int b) {  // rgbToGray; inputs <= 65536, the sum fits 32 bits unsigned
  return (int)((13933u * (uint32_t)r + 46871u * (uint32_t)g + 4732u * (uint32_t)b + (1u << 15)) >> 16);
}
WG_SHD int sharp_clip(int v) { return v < 0 ? 0 : v > kSharpMax ? kSharpMax : v; }
WG_SHD int sharp_w_of(int r, int g, int b, const uint32_t* g2l, const uint32_t* l2g) {  // updateW for one pixel
  return (int)sharp_l2g((uint32_t)sharp_gray((int)g2l[r], (int)g2l[g], (int)g2l[b]), l2g);
}
WG_SHD int sharp_scale_down(int a, int b, int c, int d, const uint32_t* g2l, const uint32_t* l2g) {
  return (int)sharp_l2g((g2l[a] + g2l[b] + g2l[c] + g2l[d] + 2) >> 2, l2g);
}
// px[row][col][channel] of one 2x2 block (10-bit) -> residual triple (updateChroma)
WG_SHD void sharp_chroma_of(const int (*px)[2][3], const uint32_t* g2l, const uint32_t* l2g, int out[3]) {
  int c[3];
  for (int k = 0; k < 3; ++k) c[k] = sharp_scale_down(px[0][0][k], px[0][1][k], px[1][0][k], px[1][1][k], g2l, l2g);
  const int gray = sharp_gray(c[0], c[1], c[2]);
  for (int k = 0; k < 3; ++k) out[k] = (int16_t)(c[k] - gray);
}

// ---- phase 1: one 2x2 block (cx, cy) of image img
WG_SHD void sharp_init_item(const SharpParams& P, int img, int cy, int cx, const uint32_t* g2l, const uint32_t* l2g) {
  const uint8_t* base = P.rgba + (size_t)img * P.image_stride;
  int px[2][2][3];
  for (int r = 0; r < 2; ++r) {
    const int sy = 2 * cy + r < P.height ? 2 * cy + r : P.height - 1;  // odd height: the last row twice
    for (int c = 0; c < 2; ++c) {
      const int sx = 2 * cx + c < P.width ? 2 * cx + c : P.width - 1;  // odd width: the last column twice
      const uint8_t* p = base + (size_t)sy * P.stride + 4 * sx;
      for (int k = 0; k < 3; ++k) px[r][c][k] = p[k] << 2;
    }
  }
  const size_t yo = ((size_t)img * P.h + 2 * cy) * P.w + 2 * cx;
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c) {
      P.best_y[yo + (size_t)r * P.w + c] = (uint16_t)sharp_gray(px[r][c][0], px[r][c][1], px[r][c][2]);  // storeGray
      P.target_y[yo + (size_t)r * P.w + c] = (uint16_t)sharp_w_of(px[r][c][0], px[r][c][1], px[r][c][2], g2l, l2g);
    }
  int uv[3];
  sharp_chroma_of(px, g2l, l2g, uv);
  const size_t uo = ((size_t)img * P.uv_h + cy) * 3 * P.uv_w + cx;
  for (int k = 0; k < 3; ++k) {
    P.target_uv[uo + (size_t)k * P.uv_w] = (int16_t)uv[k];
    P.best_uv[uo + (size_t)k * P.uv_w] = (int16_t)uv[k];
  }
}

// ---- phase 2, first half of a row pair: chroma sample i of row pair jp.  Reads best_uv rows jp-1 (already updated in this
// pass), jp and jp+1 (not yet) at i-1, i, i+1; updates its own four best_y samples; returns the new residual triple, which
// the caller stores only after every sample of the row pair has read the old row jp.  Returns sum |target_y - w|.
WG_SHD uint32_t sharp_refine_item(const SharpParams& P, int img, int jp, int i, const uint32_t* g2l, const uint32_t* l2g, int new_uv[3]) {
  const int uv_w = P.uv_w;
  const int16_t* cur = P.best_uv + ((size_t)img * P.uv_h + jp) * 3 * uv_w;
  const int16_t* prev = jp > 0 ? cur - 3 * uv_w : cur;
  const int16_t* next = jp < P.uv_h - 1 ? cur + 3 * uv_w : cur;
  const size_t yo = ((size_t)img * P.h + 2 * jp) * P.w + 2 * i;
  const int il = i > 0 ? i - 1 : 0, ir = i < uv_w - 1 ? i + 1 : uv_w - 1;
  int px[2][2][3];
  int by[2][2];
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c) by[r][c] = P.best_y[yo + (size_t)r * P.w + c];
  for (int k = 0; k < 3; ++k) {
    const int16_t* cu = cur + k * uv_w;
    const int a = cu[i];
    for (int r = 0; r < 2; ++r) {
      const int16_t* o = (r ? next : prev) + k * uv_w;  // the other chroma row: above for the upper luma row, below for the lower
      const int b = o[i];
      int vl, vr;
      if (i == 0) vl = (a * 3 + b + 2) >> 2;                                 // filter2 at x = 0
      else vl = (a * 9 + cu[il] * 3 + b * 3 + o[il] + 8) >> 4;              // x = 2i   (v1 of the pair i-1, i)
      if (i == uv_w - 1) vr = (a * 3 + b + 2) >> 2;                          // filter2 at x = w-1
      else vr = (a * 9 + cu[ir] * 3 + b * 3 + o[ir] + 8) >> 4;              // x = 2i+1 (v0 of the pair i, i+1)
      px[r][0][k] = sharp_clip(by[r][0] + vl);
      px[r][1][k] = sharp_clip(by[r][1] + vr);
    }
  }
  uint32_t diff = 0;
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c) {  // updateW + sharpYUVUpdateY
      const size_t o = yo + (size_t)r * P.w + c;
      const int d = (int)P.target_y[o] - sharp_w_of(px[r][c][0], px[r][c][1], px[r][c][2], g2l, l2g);
      P.best_y[o] = (uint16_t)sharp_clip(by[r][c] + d);
      diff += (uint32_t)(d < 0 ? -d : d);
    }
  int uv[3];
  sharp_chroma_of(px, g2l, l2g, uv);  // updateChroma + sharpYUVUpdateRGB (int16 wrap-around arithmetic as in the reference)
  const int16_t* tgt = P.target_uv + ((size_t)img * P.uv_h + jp) * 3 * uv_w;
  for (int k = 0; k < 3; ++k)
    new_uv[k] = (int16_t)(cur[k * uv_w + i] + (int16_t)(tgt[k * uv_w + i] - (int16_t)uv[k]));
  return diff;
}
WG_SHD void sharp_commit_item(const SharpParams& P, int img, int jp, int i, const int new_uv[3]) {
  int16_t* cur = P.best_uv + ((size_t)img * P.uv_h + jp) * 3 * P.uv_w;
  for (int k = 0; k < 3; ++k) cur[k * P.uv_w + i] = (int16_t)new_uv[k];
}
// stop rule after pass `iter` (sharpyuv.go:280-288)
WG_SHD bool sharp_stop(int iter, unsigned long long sum, unsigned long long prev_sum, unsigned long long threshold) {
  return iter > 0 && (sum < threshold || sum > prev_sum);
}

// ---- phase 2, pipelined (the default for rows of up to 256 * SHARP_ITEMS chroma samples): the operands of a row pair that do not depend on the row pair above it
// (its W and target samples, its target residuals, the residual row two below) are fetched one step ahead, and the residual rows
// live in a 4-slot ring in shared memory (slot = row & 3), so the dependent path of a step has no global-memory round trip.
struct SharpOperands {
  uint32_t by[2], ty[2];  // two 10-bit samples per word: W and target W of the 2x2 block, upper row then lower row
  int16_t tuv[3];         // target residuals
  int16_t ahead[3];       // residuals of row pair jp + 2 (still as the previous sweep left them), for the ring
};
WG_SHD void sharp_fetch_operands(const SharpParams& P, int img, int jp, int i, SharpOperands& o) {
  const size_t yo = ((size_t)img * P.h + 2 * jp) * P.w + 2 * i;  // even: 4-byte aligned pairs
  for (int r = 0; r < 2; ++r) {
    o.by[r] = *reinterpret_cast<const uint32_t*>(P.best_y + yo + (size_t)r * P.w);
    o.ty[r] = *reinterpret_cast<const uint32_t*>(P.target_y + yo + (size_t)r * P.w);
  }
  const int16_t* tgt = P.target_uv + ((size_t)img * P.uv_h + jp) * 3 * P.uv_w + i;
  for (int k = 0; k < 3; ++k) o.tuv[k] = tgt[k * P.uv_w];
  if (jp + 2 < P.uv_h) {
    const int16_t* a = P.best_uv + ((size_t)img * P.uv_h + jp + 2) * 3 * P.uv_w + i;
    for (int k = 0; k < 3; ++k) o.ahead[k] = a[k * P.uv_w];
  }
}
// Same arithmetic as sharp_refine_item; prev / cur / next are ring rows ([3][uv_w]), the rest comes from `o`.  Stores the four W
// samples; the caller stores new_uv (ring + global) after the barrier.
WG_SHD uint32_t sharp_refine_item_ring(const SharpParams& P, int img, int jp, int i, const int16_t* prev, const int16_t* cur, const int16_t* next,
                                       const SharpOperands& o, const uint32_t* g2l, const uint32_t* l2g, int new_uv[3]) {
  const int uv_w = P.uv_w;
  const int il = i > 0 ? i - 1 : 0, ir = i < uv_w - 1 ? i + 1 : uv_w - 1;
  int px[2][2][3];
  int by[2][2], ty[2][2];
  for (int r = 0; r < 2; ++r) {
    by[r][0] = (int)(o.by[r] & 0xffff); by[r][1] = (int)(o.by[r] >> 16);
    ty[r][0] = (int)(o.ty[r] & 0xffff); ty[r][1] = (int)(o.ty[r] >> 16);
  }
  for (int k = 0; k < 3; ++k) {
    const int16_t* cu = cur + k * uv_w;
    const int a = cu[i];
    for (int r = 0; r < 2; ++r) {
      const int16_t* q = (r ? next : prev) + k * uv_w;
      const int b = q[i];
      const int vl = i == 0 ? (a * 3 + b + 2) >> 2 : (a * 9 + cu[il] * 3 + b * 3 + q[il] + 8) >> 4;
      const int vr = i == uv_w - 1 ? (a * 3 + b + 2) >> 2 : (a * 9 + cu[ir] * 3 + b * 3 + q[ir] + 8) >> 4;
      px[r][0][k] = sharp_clip(by[r][0] + vl);
      px[r][1][k] = sharp_clip(by[r][1] + vr);
    }
  }
  uint32_t diff = 0;
  const size_t yo = ((size_t)img * P.h + 2 * jp) * P.w + 2 * i;
  for (int r = 0; r < 2; ++r) {
    uint32_t packed = 0;
    for (int c = 0; c < 2; ++c) {
      const int d = ty[r][c] - sharp_w_of(px[r][c][0], px[r][c][1], px[r][c][2], g2l, l2g);
      packed |= (uint32_t)sharp_clip(by[r][c] + d) << (16 * c);
      diff += (uint32_t)(d < 0 ? -d : d);
    }
    *reinterpret_cast<uint32_t*>(P.best_y + yo + (size_t)r * P.w) = packed;
  }
  int uv[3];
  sharp_chroma_of(px, g2l, l2g, uv);
  for (int k = 0; k < 3; ++k) new_uv[k] = (int16_t)(cur[k * uv_w + i] + (int16_t)(o.tuv[k] - (int16_t)uv[k]));
  return diff;
}
// after the barrier: the new residuals of row pair jp into its ring slot and into global memory, row pair jp + 2 into the ring
WG_SHD void sharp_commit_item_ring(const SharpParams& P, int img, int jp, int i, int16_t* ring, const SharpOperands& o, const int new_uv[3]) {
  const int uv_w = P.uv_w;
  int16_t* slot = ring + (size_t)(jp & 3) * 3 * uv_w;
  int16_t* g = P.best_uv + ((size_t)img * P.uv_h + jp) * 3 * uv_w;
  for (int k = 0; k < 3; ++k) { slot[k * uv_w + i] = (int16_t)new_uv[k]; g[k * uv_w + i] = (int16_t)new_uv[k]; }
  if (jp + 2 < P.uv_h) {
    int16_t* ahead = ring + (size_t)((jp + 2) & 3) * 3 * uv_w;
    for (int k = 0; k < 3; ++k) ahead[k * uv_w + i] = o.ahead[k];
  }
}
// start of a sweep: row pairs 0 and 1 into ring slots 0 and 1 (element e of 3 * uv_w)
WG_SHD void sharp_ring_preload(const SharpParams& P, int img, int e, int16_t* ring) {
  const int16_t* g = P.best_uv + (size_t)img * P.uv_h * 3 * P.uv_w;
  ring[e] = g[e];
  if (P.uv_h > 1) ring[3 * P.uv_w + e] = g[3 * P.uv_w + e];
}

// ---- phase 3: padded 2x2 block (cx, cy) of the encoder's source planes (convertWRGBToYUV + importYCbCr)
WG_SHD int sharp_matrix(int r, int g, int b, int c0, int c1, int c2, int offset) {  // offsets << 2, rounder 1 << 17, >> 18
  const long long v = (long long)c0 * r + (long long)c1 * g + (long long)c2 * b + ((long long)offset << 2) + (1ll << 17);
  const int o = (int)(v >> 18);
  return o < 0 ? 0 : o > 255 ? 255 : o;
}
WG_SHD void sharp_finish_item(const SharpParams& P, int img, int cy, int cx) {
  const int uvh_pic = (P.height + 1) >> 1, uvw_pic = (P.width + 1) >> 1;
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c) {
      const int py = 2 * cy + r, pxx = 2 * cx + c;
      const int sy = py < P.height ? py : P.height - 1, sx = pxx < P.width ? pxx : P.width - 1;
      const int wv = P.best_y[((size_t)img * P.h + sy) * P.w + sx];
      const int16_t* q = P.best_uv + ((size_t)img * P.uv_h + (sy >> 1)) * 3 * P.uv_w + (sx >> 1);
      P.y[(size_t)img * P.y_plane + (size_t)py * P.pad_w + pxx] =
          (uint8_t)sharp_matrix(q[0] + wv, q[P.uv_w] + wv, q[2 * P.uv_w] + wv, 16839, 33059, 6420, 16 << 16);
    }
  const int sy = cy < uvh_pic ? cy : uvh_pic - 1, sx = cx < uvw_pic ? cx : uvw_pic - 1;
  const int16_t* q = P.best_uv + ((size_t)img * P.uv_h + sy) * 3 * P.uv_w + sx;
  const size_t o = (size_t)img * P.uv_plane + (size_t)cy * (P.pad_w >> 1) + cx;
  P.u[o] = (uint8_t)sharp_matrix(q[0], q[P.uv_w], q[2 * P.uv_w], -9719, -19081, 28800, 128 << 16);
  P.v[o] = (uint8_t)sharp_matrix(q[0], q[P.uv_w], q[2 * P.uv_w], 28800, -24116, -4684, 128 << 16);
}

#ifdef __CUDACC__
enum { SHARP_ITEMS = 8 };  // a CTA keeps up to 8 chroma samples of a row pair per thread in registers

__device__ __forceinline__ void sharp_load_tables(const SharpParams& P, uint32_t* s_g2l, uint32_t* s_l2g) {
  for (int i = threadIdx.x; i < 1026; i += blockDim.x) s_g2l[i] = P.g2l[i];
  for (int i = threadIdx.x; i < 514; i += blockDim.x) s_l2g[i] = P.l2g[i];
  __syncthreads();
}
__global__ void __launch_bounds__(256) sharp_init_kernel(const SharpParams P) {
  __shared__ uint32_t s_g2l[1026], s_l2g[514];
  sharp_load_tables(P, s_g2l, s_l2g);
  const long long per_img = (long long)P.uv_w * P.uv_h, total = per_img * P.n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(t / per_img);
    const int rem = (int)(t - (long long)img * per_img);
    const int cy = rem / P.uv_w;
    sharp_init_item(P, img, cy, rem - cy * P.uv_w, s_g2l, s_l2g);
  }
}
// One CTA per image.  Row pairs in order; per row pair every thread computes its samples (new residuals kept in registers),
// barrier, stores them, barrier.  SHARP_THREADS * SHARP_ITEMS >= uv_w (256 threads up to 4096 pixels wide, 1024 beyond).
template <int SHARP_THREADS>
__global__ void __launch_bounds__(SHARP_THREADS) sharp_refine_kernel(const SharpParams P) {
  __shared__ uint32_t s_g2l[1026], s_l2g[514];
  __shared__ unsigned long long s_part[SHARP_THREADS / 32];
  __shared__ unsigned long long s_sum;
  sharp_load_tables(P, s_g2l, s_l2g);
  const int img = blockIdx.x;
  const unsigned long long threshold = 3ull * (unsigned long long)P.w * (unsigned long long)P.h;
  unsigned long long prev_sum = ~0ull;
  int iters = 0;
  for (int iter = 0; iter < 4; ++iter) {
    unsigned long long mine = 0;
    ++iters;
    for (int jp = 0; jp < P.uv_h; ++jp) {
      int keep[SHARP_ITEMS][3];
#pragma unroll
      for (int k = 0; k < SHARP_ITEMS; ++k) {
        const int i = threadIdx.x + k * SHARP_THREADS;
        if (i < P.uv_w) mine += sharp_refine_item(P, img, jp, i, s_g2l, s_l2g, keep[k]);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SHARP_ITEMS; ++k) {
        const int i = threadIdx.x + k * SHARP_THREADS;
        if (i < P.uv_w) sharp_commit_item(P, img, jp, i, keep[k]);
      }
      __syncthreads();
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long s = 0;
      for (int k = 0; k < SHARP_THREADS / 32; ++k) s += s_part[k];
      s_sum = s;
    }
    __syncthreads();
    const unsigned long long sum = s_sum;
    if (sharp_stop(iter, sum, prev_sum, threshold)) break;
    prev_sum = sum;
  }
  if (P.iterations && threadIdx.x == 0) P.iterations[img] = iters;
}
// Pipelined variant: 256 threads, ITEMS samples per thread (256 * ITEMS >= uv_w), dynamic shared memory = ring of 4 residual rows.
template <int ITEMS>
__global__ void __launch_bounds__(256) sharp_refine_ring_kernel(const SharpParams P) {
  extern __shared__ __align__(16) int16_t s_ring[];  // [4][3][uv_w]
  __shared__ uint32_t s_g2l[1026], s_l2g[514];
  __shared__ unsigned long long s_part[8];
  __shared__ unsigned long long s_sum;
  sharp_load_tables(P, s_g2l, s_l2g);
  const int img = blockIdx.x, uv_w = P.uv_w, row_len = 3 * uv_w;
  const unsigned long long threshold = 3ull * (unsigned long long)P.w * (unsigned long long)P.h;
  unsigned long long prev_sum = ~0ull;
  int iters = 0;
  for (int iter = 0; iter < 4; ++iter) {
    unsigned long long mine = 0;
    ++iters;
    for (int e = threadIdx.x; e < row_len; e += 256) sharp_ring_preload(P, img, e, s_ring);
    SharpOperands op[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < uv_w) sharp_fetch_operands(P, img, 0, i, op[k]);
    }
    __syncthreads();
    for (int jp = 0; jp < P.uv_h; ++jp) {
      SharpOperands nx[ITEMS];
      if (jp + 1 < P.uv_h) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
          const int i = threadIdx.x + k * 256;
          if (i < uv_w) sharp_fetch_operands(P, img, jp + 1, i, nx[k]);  // in flight while this step computes
        }
      }
      const int16_t* cur = s_ring + (jp & 3) * row_len;
      const int16_t* prev = jp > 0 ? s_ring + ((jp - 1) & 3) * row_len : cur;
      const int16_t* next = jp < P.uv_h - 1 ? s_ring + ((jp + 1) & 3) * row_len : cur;
      int keep[ITEMS][3];
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < uv_w) mine += sharp_refine_item_ring(P, img, jp, i, prev, cur, next, op[k], s_g2l, s_l2g, keep[k]);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < uv_w) sharp_commit_item_ring(P, img, jp, i, s_ring, op[k], keep[k]);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) op[k] = nx[k];
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long s = 0;
      for (int k = 0; k < 8; ++k) s += s_part[k];
      s_sum = s;
    }
    __syncthreads();
    const unsigned long long sum = s_sum;
    if (sharp_stop(iter, sum, prev_sum, threshold)) break;
    prev_sum = sum;
  }
  if (P.iterations && threadIdx.x == 0) P.iterations[img] = iters;
}
__global__ void __launch_bounds__(256) sharp_finish_kernel(const SharpParams P) {
  const int qw = P.pad_w >> 1, qh = P.pad_h >> 1;
  const long long per_img = (long long)qw * qh, total = per_img * P.n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(t / per_img);
    const int rem = (int)(t - (long long)img * per_img);
    const int cy = rem / qw;
    sharp_finish_item(P, img, cy, rem - cy * qw);
  }
}
#endif  // __CUDACC__

}  // namespace wg
