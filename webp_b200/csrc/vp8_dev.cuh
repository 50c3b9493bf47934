// Device-side VP8 primitives (sm_100a).  Register-resident 4x4 blocks: every routine works on
// int[16] in raster order so the transforms fully unroll into IADD3/IMAD/SHF with no local memory.
// Arithmetic follows the reference's scalar Go code (file:line cited per function); all
// intermediate ranges fit int32 (|coeff| <= 2048+sharpen, residual <= 255).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WG_HD __host__ __device__ __forceinline__
#else
#include <algorithm>
#define WG_HD inline
#endif

namespace wg {
#ifndef __CUDACC__
// by value (std::min / std::max return references to their arguments, temporaries included)
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
#endif

#include "vp8_tables.inc"  // host-visible copies (static const); device copies below

// ---- tables in device global memory (L1/L2 cached); hot kernels stage them into shared memory
// The per-block routines below are host+device (WG_HD) so that a CPU harness (the hostcheck test harness) can run the kernels' code
// in the kernels' schedule; on the host the small tables are plain static arrays.
#define WG_TAB_ZIGZAG {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15}
#define WG_TAB_BANDS {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0}
#define WG_TAB_REV_ZIGZAG {0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15}
#define WG_TAB_WEIGHT_Y {38, 32, 20, 9, 32, 28, 17, 7, 20, 17, 10, 4, 9, 7, 4, 2}
#define WG_TAB_WEIGHT_TRELLIS {30, 27, 19, 11, 27, 24, 17, 10, 19, 17, 12, 8, 11, 10, 8, 6}
#ifdef __CUDACC__
__device__ __constant__ uint8_t c_zigzag_d[16] = WG_TAB_ZIGZAG;
__device__ __constant__ uint8_t c_bands_d[17] = WG_TAB_BANDS;
__device__ __constant__ uint8_t c_rev_zigzag_d[16] = WG_TAB_REV_ZIGZAG;
__device__ __constant__ uint8_t c_weight_trellis_d[16] = WG_TAB_WEIGHT_TRELLIS;
#endif
static const uint8_t c_zigzag_h[16] = WG_TAB_ZIGZAG;
static const uint8_t c_bands_h[17] = WG_TAB_BANDS;
static const uint8_t c_rev_zigzag_h[16] = WG_TAB_REV_ZIGZAG;
static const uint8_t c_weight_trellis_h[16] = WG_TAB_WEIGHT_TRELLIS;
#ifdef __CUDA_ARCH__
#define c_zigzag c_zigzag_d
#define c_bands c_bands_d
#define c_rev_zigzag c_rev_zigzag_d
#define c_weight_trellis c_weight_trellis_d
#else
#define c_zigzag c_zigzag_h
#define c_bands c_bands_h
#define c_rev_zigzag c_rev_zigzag_h
#define c_weight_trellis c_weight_trellis_h
#endif

// Cost tables a kernel uses (global or shared memory).  On the reference's row-parallel path the coefficient
// probabilities are the constant defaults (encode_parallel.go:1558-1567), so everything TokenCostForCoeffs and
// TrellisQuantizeBlock look up per coefficient is folded on the host into one table per (type, band, ctx, level):
//   lc[((type*8+band)*3+ctx)*68 + min(v,67)] = cost(not EOB) + (v == 0 ? cost(zero) : cost(non-zero) + variableLevelCost(v))
//   eob[(type*8+band)*3+ctx]                 = cost(EOB)
// built from VP8EntropyCost (internal/dsp/cost.go:6), vp8LevelCodes (internal/lossy/encode_quant.go:226) and
// CoeffsProba0 (internal/lossy/proba.go:45); lfc = VP8LevelFixedCosts[2048] (internal/dsp/cost.go:33) stays separate.
enum { LC_LEVELS = 68, LC_SIZE = 4 * 8 * 3 * LC_LEVELS, EOB_SIZE = 4 * 8 * 3 };
struct CostTabs {
  const uint16_t* lc;
  const uint16_t* eob;
  const uint16_t* lfc;     // VP8LevelFixedCosts: at least the first LFC_NEAR entries (the phased kernel stages only those)
  const uint16_t* lfc_hi;  // the whole table (levels >= LFC_NEAR: very low quality, rare)
};
enum { LFC_NEAR = 512 };
WG_HD int lfc_at(const CostTabs& T, int v) { return v < LFC_NEAR ? T.lfc[v] : T.lfc_hi[v]; }

// SegmentQuant (internal/lossy/encode.go:311)
struct SegQuant {
  int quant, iquant, bias, dc_quant, dc_iquant, dc_bias;
  int16_t sharpen[16];
};
// SegmentInfo subset used on the device (internal/lossy/encode.go:278)
struct SegParams {
  SegQuant y1, y2, uv;
  int lambda_i4, lambda_i16, lambda_uv, lambda_mode, tlambda_i4, tlambda_i16, tlambda_sd;
  int flags;  // seg[0] only, serial rate-control passes: bits 0-7 per-image max I4 RD modes (0 = launch default), bit 8 = image parked
};

WG_HD int clip8(int v) { return min(max(v, 0), 255); }
WG_HD int mul1(int a) { return ((a * 20091) >> 16) + a; }  // transforms.go:20
WG_HD int mul2(int a) { return (a * 35468) >> 16; }        // transforms.go:25

// fTransform (internal/dsp/transforms.go:371): src - ref -> 16 coefficients
WG_HD void ftransform(const int* src, const int* ref, int* out) {
  int tmp[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int d0 = src[4 * j + 0] - ref[4 * j + 0], d1 = src[4 * j + 1] - ref[4 * j + 1];
    const int d2 = src[4 * j + 2] - ref[4 * j + 2], d3 = src[4 * j + 3] - ref[4 * j + 3];
    const int a0 = d0 + d3, a1 = d1 + d2, a2 = d1 - d2, a3 = d0 - d3;
    tmp[4 * j + 0] = (a0 + a1) * 8;
    tmp[4 * j + 1] = (a2 * 2217 + a3 * 5352 + 1812) >> 9;
    tmp[4 * j + 2] = (a0 - a1) * 8;
    tmp[4 * j + 3] = (a3 * 2217 - a2 * 5352 + 937) >> 9;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a0 = tmp[0 + i] + tmp[12 + i], a1 = tmp[4 + i] + tmp[8 + i];
    const int a2 = tmp[4 + i] - tmp[8 + i], a3 = tmp[0 + i] - tmp[12 + i];
    out[0 + i] = (a0 + a1 + 7) >> 4;
    out[4 + i] = ((a2 * 2217 + a3 * 5352 + 12000) >> 16) + (a3 != 0);
    out[8 + i] = (a0 - a1 + 7) >> 4;
    out[12 + i] = (a3 * 2217 - a2 * 5352 + 51000) >> 16;
  }
}
// iTransformOne (internal/dsp/transforms.go:265): dst = clip(ref + IDCT(in)); dst may alias ref
WG_HD void itransform(const int* ref, const int* in, int* dst) {
  int tmp[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = in[i] + in[8 + i], b = in[i] - in[8 + i];
    const int c = mul2(in[4 + i]) - mul1(in[12 + i]);
    const int d = mul1(in[4 + i]) + mul2(in[12 + i]);
    tmp[i] = a + d;
    tmp[4 + i] = b + c;
    tmp[8 + i] = b - c;
    tmp[12 + i] = a - d;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int dc = tmp[4 * j] + 4;
    const int a = dc + tmp[4 * j + 2], b = dc - tmp[4 * j + 2];
    const int c = mul2(tmp[4 * j + 1]) - mul1(tmp[4 * j + 3]);
    const int d = mul1(tmp[4 * j + 1]) + mul2(tmp[4 * j + 3]);
    const int r0 = ref[4 * j + 0], r1 = ref[4 * j + 1], r2 = ref[4 * j + 2], r3 = ref[4 * j + 3];
    dst[4 * j + 0] = clip8(r0 + ((a + d) >> 3));
    dst[4 * j + 1] = clip8(r1 + ((b + c) >> 3));
    dst[4 * j + 2] = clip8(r2 + ((b - c) >> 3));
    dst[4 * j + 3] = clip8(r3 + ((a - d) >> 3));
  }
}
// fTransformWHT (transforms.go:500): flat 4x4 DCs -> 16
WG_HD void fwht(const int* in, int* out) {
  int tmp[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a0 = in[i * 4 + 0] + in[i * 4 + 2], a1 = in[i * 4 + 1] + in[i * 4 + 3];
    const int a2 = in[i * 4 + 1] - in[i * 4 + 3], a3 = in[i * 4 + 0] - in[i * 4 + 2];
    tmp[0 + i * 4] = a0 + a1;
    tmp[1 + i * 4] = a3 + a2;
    tmp[2 + i * 4] = a3 - a2;
    tmp[3 + i * 4] = a0 - a1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a0 = tmp[0 + i] + tmp[8 + i], a1 = tmp[4 + i] + tmp[12 + i];
    const int a2 = tmp[4 + i] - tmp[12 + i], a3 = tmp[0 + i] - tmp[8 + i];
    out[0 + i] = (a0 + a1) >> 1;
    out[4 + i] = (a3 + a2) >> 1;
    out[8 + i] = (a3 - a2) >> 1;
    out[12 + i] = (a0 - a1) >> 1;
  }
}
// transformWHT (transforms.go:223); out[b] = DC of block b (the reference's stride-16 layout, flattened);
// values are truncated to int16 as the reference stores them.
WG_HD void iwht(const int* in, int* out) {
  int tmp[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a0 = in[0 + i] + in[12 + i], a1 = in[4 + i] + in[8 + i];
    const int a2 = in[4 + i] - in[8 + i], a3 = in[0 + i] - in[12 + i];
    tmp[0 + i] = a0 + a1;
    tmp[8 + i] = a0 - a1;
    tmp[4 + i] = a3 + a2;
    tmp[12 + i] = a3 - a2;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int dc = tmp[i * 4 + 0] + 3;
    const int a0 = dc + tmp[i * 4 + 3], a1 = tmp[i * 4 + 1] + tmp[i * 4 + 2];
    const int a2 = tmp[i * 4 + 1] - tmp[i * 4 + 2], a3 = dc - tmp[i * 4 + 3];
    out[i * 4 + 0] = (int)(int16_t)((a0 + a1) >> 3);
    out[i * 4 + 1] = (int)(int16_t)((a3 + a2) >> 3);
    out[i * 4 + 2] = (int)(int16_t)((a0 - a1) >> 3);
    out[i * 4 + 3] = (int)(int16_t)((a3 - a2) >> 3);
  }
}
WG_HD int sse16(const int* a, const int* b) {  // ssim.go:188
  int s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int d = a[i] - b[i];
    s += d * d;
  }
  return s;
}
WG_HD int ttransform(const int* in) {  // ssim.go:266 with kWeightY
  int tmp[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a0 = in[4 * i + 0] + in[4 * i + 2], a1 = in[4 * i + 1] + in[4 * i + 3];
    const int a2 = in[4 * i + 1] - in[4 * i + 3], a3 = in[4 * i + 0] - in[4 * i + 2];
    tmp[0 + i * 4] = a0 + a1;
    tmp[1 + i * 4] = a3 + a2;
    tmp[2 + i * 4] = a3 - a2;
    tmp[3 + i * 4] = a0 - a1;
  }
  const int w[16] = {38, 32, 20, 9, 32, 28, 17, 7, 20, 17, 10, 4, 9, 7, 4, 2};
  int sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a0 = tmp[0 + i] + tmp[8 + i], a1 = tmp[4 + i] + tmp[12 + i];
    const int a2 = tmp[4 + i] - tmp[12 + i], a3 = tmp[0 + i] - tmp[8 + i];
    sum += w[0 + i] * abs(a0 + a1) + w[4 + i] * abs(a3 + a2) + w[8 + i] * abs(a3 - a2) + w[12 + i] * abs(a0 - a1);
  }
  return sum;
}
WG_HD int tdisto4x4(const int* a, const int* b) {  // ssim.go:315
  return abs(ttransform(b) - ttransform(a)) >> 5;
}

// PredLuma4Direct (internal/dsp/predict_lossy.go:185-451).  e[0]=top-left, e[1..8]=top[0..7], e[9..12]=left[0..3]
WG_HD int avg3(int a, int b, int c) { return (a + 2 * b + c + 2) >> 2; }
WG_HD int avg2(int a, int b) { return (a + b + 1) >> 1; }
WG_HD void pred4(int mode, const int* e, int* d) {
  const int tl = e[0], t0 = e[1], t1 = e[2], t2 = e[3], t3 = e[4], t4 = e[5], t5 = e[6], t6 = e[7], t7 = e[8];
  const int l0 = e[9], l1 = e[10], l2 = e[11], l3 = e[12];
#define D(x, y) d[(x) + 4 * (y)]
  switch (mode) {
    case 0: {
      const int v = (t0 + t1 + t2 + t3 + l0 + l1 + l2 + l3 + 4) >> 3;
#pragma unroll
      for (int i = 0; i < 16; ++i) d[i] = v;
    } break;
    case 1: {
      const int l[4] = {l0, l1, l2, l3}, t[4] = {t0, t1, t2, t3};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) D(i, j) = clip8(l[j] + t[i] - tl);
    } break;
    case 2: {
      const int v0 = avg3(tl, t0, t1), v1 = avg3(t0, t1, t2), v2 = avg3(t1, t2, t3), v3 = avg3(t2, t3, t4);
#pragma unroll
      for (int j = 0; j < 4; ++j) { D(0, j) = v0; D(1, j) = v1; D(2, j) = v2; D(3, j) = v3; }
    } break;
    case 3: {
      const int v[4] = {avg3(tl, l0, l1), avg3(l0, l1, l2), avg3(l1, l2, l3), avg3(l2, l3, l3)};
#pragma unroll
      for (int j = 0; j < 4; ++j) { D(0, j) = v[j]; D(1, j) = v[j]; D(2, j) = v[j]; D(3, j) = v[j]; }
    } break;
    case 4:
      D(0, 3) = avg3(l3, l2, l1);
      D(0, 2) = D(1, 3) = avg3(l2, l1, l0);
      D(0, 1) = D(1, 2) = D(2, 3) = avg3(l1, l0, tl);
      D(0, 0) = D(1, 1) = D(2, 2) = D(3, 3) = avg3(l0, tl, t0);
      D(1, 0) = D(2, 1) = D(3, 2) = avg3(tl, t0, t1);
      D(2, 0) = D(3, 1) = avg3(t0, t1, t2);
      D(3, 0) = avg3(t1, t2, t3);
      break;
    case 5:
      D(0, 0) = D(1, 2) = avg2(tl, t0);
      D(1, 0) = D(2, 2) = avg2(t0, t1);
      D(2, 0) = D(3, 2) = avg2(t1, t2);
      D(3, 0) = avg2(t2, t3);
      D(0, 1) = D(1, 3) = avg3(l0, tl, t0);
      D(1, 1) = D(2, 3) = avg3(tl, t0, t1);
      D(2, 1) = D(3, 3) = avg3(t0, t1, t2);
      D(3, 1) = avg3(t1, t2, t3);
      D(0, 2) = avg3(l1, l0, tl);
      D(0, 3) = avg3(l2, l1, l0);
      break;
    case 6:
      D(0, 0) = avg3(t0, t1, t2);
      D(1, 0) = D(0, 1) = avg3(t1, t2, t3);
      D(2, 0) = D(1, 1) = D(0, 2) = avg3(t2, t3, t4);
      D(3, 0) = D(2, 1) = D(1, 2) = D(0, 3) = avg3(t3, t4, t5);
      D(3, 1) = D(2, 2) = D(1, 3) = avg3(t4, t5, t6);
      D(3, 2) = D(2, 3) = avg3(t5, t6, t7);
      D(3, 3) = avg3(t6, t7, t7);
      break;
    case 7:
      D(0, 0) = avg2(t0, t1);
      D(1, 0) = D(0, 2) = avg2(t1, t2);
      D(2, 0) = D(1, 2) = avg2(t2, t3);
      D(3, 0) = D(2, 2) = avg2(t3, t4);
      D(0, 1) = avg3(t0, t1, t2);
      D(1, 1) = D(0, 3) = avg3(t1, t2, t3);
      D(2, 1) = D(1, 3) = avg3(t2, t3, t4);
      D(3, 1) = D(2, 3) = avg3(t3, t4, t5);
      D(3, 2) = avg3(t4, t5, t6);
      D(3, 3) = avg3(t5, t6, t7);
      break;
    case 8:
      D(0, 0) = D(2, 1) = avg2(tl, l0);
      D(1, 0) = D(3, 1) = avg3(l0, tl, t0);
      D(2, 0) = avg3(tl, t0, t1);
      D(3, 0) = avg3(t0, t1, t2);
      D(0, 1) = D(2, 2) = avg2(l0, l1);
      D(1, 1) = D(3, 2) = avg3(tl, l0, l1);
      D(0, 2) = D(2, 3) = avg2(l1, l2);
      D(1, 2) = D(3, 3) = avg3(l0, l1, l2);
      D(0, 3) = avg2(l2, l3);
      D(1, 3) = avg3(l1, l2, l3);
      break;
    default:
      D(0, 0) = avg2(l0, l1);
      D(1, 0) = avg3(l0, l1, l2);
      D(2, 0) = D(0, 1) = avg2(l1, l2);
      D(3, 0) = D(1, 1) = avg3(l1, l2, l3);
      D(2, 1) = D(0, 2) = avg2(l2, l3);
      D(3, 1) = D(1, 2) = avg3(l2, l3, l3);
      D(2, 2) = D(3, 2) = D(0, 3) = D(1, 3) = D(2, 3) = D(3, 3) = l3;
      break;
  }
#undef D
}

// quantizeCoeffsGo (internal/lossy/encode_quant.go:16): returns zigzag last-nz + 1
WG_HD int quantize_block(const int* in, int* out, const SegQuant& sq, int first) {
  int max_zz = -1;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    if (n == 0 && first != 0) { out[0] = 0; continue; }
    int v = in[n];
    const bool neg = v < 0;
    v = abs(v) + sq.sharpen[n];
    if (v < 0) v = 0;
    const uint32_t iq = (n == 0) ? (uint32_t)sq.dc_iquant : (uint32_t)sq.iquant;
    const uint32_t bias = (n == 0) ? (uint32_t)sq.dc_bias : (uint32_t)sq.bias;
    int coeff = (int)(((uint32_t)v * iq + bias) >> 17);
    coeff = min(coeff, 2047);
    out[n] = neg ? -coeff : coeff;
    if (coeff) max_zz = max(max_zz, (int)c_rev_zigzag[n]);
  }
  return max_zz + 1;
}
// dequantCoeffsGo (encode_quant.go:81): int16 truncation kept
WG_HD void dequant_block(const int* in, int* out, const SegQuant& sq) {
  out[0] = (int)(int16_t)(in[0] * sq.dc_quant);
#pragma unroll
  for (int n = 1; n < 16; ++n) out[n] = (int)(int16_t)(in[n] * sq.quant);
}

// Rolled, in-place variants for data that already lives in shared memory.  Small loop bodies matter more than a few
// extra LDS here: the mode-search kernel is instruction-fetch bound when primitives are unrolled and inlined at every
// call site (L0 I-cache ~6 KB, L1.5 32 KB; profiles/README.md).
template <class LevT>
WG_HD int quantize_smem(LevT* io, const SegQuant& sq, int first) {  // quantizeCoeffsGo, encode_quant.go:16
  int max_zz = -1;
  if (first != 0) io[0] = 0;
#pragma unroll 1
  for (int n = first; n < 16; ++n) {  // raster index; position 0 is the only one `first` can exclude
    int v = io[n];
    const bool neg = v < 0;
    v = max(abs(v) + sq.sharpen[n], 0);
    const uint32_t iq = (n == 0) ? (uint32_t)sq.dc_iquant : (uint32_t)sq.iquant;
    const uint32_t bias = (n == 0) ? (uint32_t)sq.dc_bias : (uint32_t)sq.bias;
    const int coeff = min((int)(((uint32_t)v * iq + bias) >> 17), 2047);
    io[n] = (LevT)(neg ? -coeff : coeff);
    if (coeff) max_zz = max(max_zz, (int)c_rev_zigzag[n]);
  }
  return max_zz + 1;
}
template <class LevT>
WG_HD int token_cost_smem(const LevT* lev, int nz_count, int type, int ctx0, int first, const CostTabs& T) {
  const uint16_t* lc = T.lc + type * (8 * 3 * LC_LEVELS);
  const uint16_t* eob = T.eob + type * (8 * 3);
  if (nz_count <= first) return eob[c_bands[first] * 3 + ctx0];
  int cost = 0, ctx = ctx0;
#pragma unroll 1
  for (int n = first; n < nz_count; ++n) {
    const int v = abs((int)lev[c_zigzag[n]]);
    cost += lc[(c_bands[n] * 3 + ctx) * LC_LEVELS + min(v, LC_LEVELS - 1)] + T.lfc[v];
    ctx = min(v, 2);
  }
  if (nz_count < 16) cost += eob[c_bands[nz_count] * 3 + ctx];
  return cost;
}

// TokenCostForCoeffs (encode_quant.go:170); levels in raster order.  Per coefficient: one folded-table lookup
// (+ the fixed level cost); note the reference charges the not-EOB bit at every position up to the last non-zero.
WG_HD int token_cost(const int* lev, int nz_count, int type, int ctx0, int first, const CostTabs& T) {
  const uint16_t* lc = T.lc + type * (8 * 3 * LC_LEVELS);
  const uint16_t* eob = T.eob + type * (8 * 3);
  if (nz_count <= first) return eob[c_bands[first] * 3 + ctx0];
  const int last = nz_count - 1;
  int cost = 0, ctx = ctx0;
  constexpr int kZig[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
  constexpr int kBnd[17] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0};
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    if (n >= first && n <= last) {
      const int v = abs(lev[kZig[n]]);
      cost += lc[(kBnd[n] * 3 + ctx) * LC_LEVELS + min(v, LC_LEVELS - 1)] + lfc_at(T, v);
      ctx = min(v, 2);
    }
  }
  if (last < 15) cost += eob[c_bands[last + 1] * 3 + ctx];
  return cost;
}

// TrellisQuantizeBlock (internal/lossy/encode_trellis.go:23-324), in place on 16 int16 values in SHARED memory
// (raster order: coefficients in, levels out).  The position loop is ROLLED (a fully unrolled 16 x 3 x 3 Viterbi is
// ~5.5k instructions and thrashes the instruction cache, profiles/r1_mode_search_c_*.md); shared memory provides the
// zigzag-indexed access, the survivor path is 10 bits per position in three 64-bit registers:
//   bits 0-2  ctx0: prev ctx (2) + valid          bits 3-5  ctx1: prev ctx + valid (level is L0 if L0 == 1 else L0+1)
//   bits 6-9  ctx2: prev ctx + valid + "level is L0+1"
// Scores are int64 as in the reference.  An unreachable state carries kBig (the reference's kMaxScore = 1<<60); real
// scores stay below 2^50, so >= kThr means "not valid" and transitions out of an unreachable state can never beat a
// reachable one -- the reference's `if (!prev[pc].valid) continue` without a branch.
WG_HD int trellis_block_smem(int16_t* io, const SegQuant& sq, int first, int type, int initial_ctx,
                                                  int lambda, const CostTabs& T) {
  const int quant_ac = sq.quant, quant_dc = sq.dc_quant;
  const unsigned iq_ac = (unsigned)sq.iquant, iq_dc = (unsigned)sq.dc_iquant;
  uint32_t neg_mask = 0;  // bit i: raster coefficient i is negative
  {
    bool non_zero = false;  // all-zero pre-scan with neutral bias (encode_trellis.go:39-98)
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
      const int raw = io[i];
      const int c0 = max(abs(raw) + sq.sharpen[i], 0);
      neg_mask |= (raw < 0 ? 1u : 0u) << i;
      io[i] = (int16_t)c0;  // <= 32767 + sharpen is clamped by the int16 store only for inputs the encoder never produces
      if (i > 0 || first == 0) non_zero |= (((unsigned)c0 * (i == 0 ? iq_dc : iq_ac)) >> 17) > 0;
    }
    if (!non_zero) {
#pragma unroll
      for (int i = 0; i < 16; ++i) io[i] = 0;
      return 0;
    }
  }
  initial_ctx = min(initial_ctx, 2);
  const uint16_t* lc = T.lc + type * (8 * 3 * LC_LEVELS);
  const uint16_t* eob = T.eob + type * (8 * 3);
  const long long kBig = 1ll << 60, kThr = 1ll << 59;
  long long ps0 = initial_ctx == 0 ? 0 : kBig, ps1 = initial_ctx == 1 ? 0 : kBig, ps2 = initial_ctx == 2 ? 0 : kBig;
  const long long lam = lambda;
  long long best_terminal = (long long)eob[c_bands[first] * 3 + initial_ctx] * lam;
  int best_last_n = -1, best_last_ctx = -1;
  unsigned long long pw0 = 0, pw1 = 0, pw2 = 0;  // survivor paths: positions 0-5, 6-11, 12-15
#pragma unroll 1
  for (int n = first; n < 16; ++n) {
    const int zig = c_zigzag[n];
    const int band = c_bands[n + 1];  // sic: the next position's band (encode_trellis.go:151)
    const int coeff0 = io[zig];
    const int quant = (n == 0) ? quant_dc : quant_ac;
    const unsigned iquant = (n == 0) ? iq_dc : iq_ac;
    const int L0 = min((int)(((unsigned)coeff0 * iquant) >> 17), 2047);
    const int thresh = min((int)(((unsigned)coeff0 * iquant + 65536u) >> 17), 2047);
    io[zig] = (int16_t)L0;  // the backtrack only needs the base level
    const bool has_l0 = L0 > 0 && L0 <= thresh;
    const bool has_l1 = L0 + 1 <= 2047 && L0 + 1 <= thresh;
    const int c0sq = coeff0 * coeff0;
    const int e0 = coeff0 - L0 * quant, e1 = coeff0 - (L0 + 1) * quant;
    const int wt256 = c_weight_trellis[zig] * 256;
    const long long disto_l0 = (long long)(e0 * e0 - c0sq) * wt256;
    const long long disto_l1 = (long long)(e1 * e1 - c0sq) * wt256;
    const int lfc0 = T.lfc[L0], lfc1 = T.lfc[min(L0 + 1, 2047)];
    const int li0 = min(L0, LC_LEVELS - 1), li1 = min(L0 + 1, LC_LEVELS - 1);
    // ctx 1 receives at most one candidate (L0 == 1: level L0; L0 == 0: level L0+1); ctx 2 receives level L0 when
    // L0 >= 2, then level L0+1 when L0 >= 1, in that order (strict '<' keeps the earlier on ties, as the reference does)
    const bool c1_from_l0 = (L0 == 1);
    const bool c1_ok = c1_from_l0 ? has_l0 : (L0 == 0 && has_l1);
    const int c1_li = c1_from_l0 ? li0 : li1, c1_lfc = c1_from_l0 ? lfc0 : lfc1;
    const long long c1_d = c1_from_l0 ? disto_l0 : disto_l1;
    const bool c2a_ok = has_l0 && L0 >= 2, c2b_ok = has_l1 && L0 >= 1;
    long long cs0 = 2 * kBig, cs1 = 2 * kBig, cs2 = 2 * kBig;
    uint32_t p_0 = 0, p_1 = 0, p_2 = 0;  // prev ctx of the survivor; p_2 bit 3 = took level L0+1
    const uint16_t* row = lc + band * 3 * LC_LEVELS;
#pragma unroll
    for (int pc = 0; pc < 3; ++pc) {
      const long long prev = pc == 0 ? ps0 : (pc == 1 ? ps1 : ps2);
      const uint16_t* r = row + pc * LC_LEVELS;
      const long long t0 = prev + (long long)r[0] * lam;
      const long long t1 = c1_ok ? prev + (long long)((int)r[c1_li] + c1_lfc) * lam + c1_d : kBig;
      const long long t2a = c2a_ok ? prev + (long long)((int)r[li0] + lfc0) * lam + disto_l0 : kBig;
      const long long t2b = c2b_ok ? prev + (long long)((int)r[li1] + lfc1) * lam + disto_l1 : kBig;
      const bool u0 = t0 < cs0;
      cs0 = u0 ? t0 : cs0; p_0 = u0 ? pc : p_0;
      const bool u1 = t1 < cs1;
      cs1 = u1 ? t1 : cs1; p_1 = u1 ? pc : p_1;
      const bool u2a = t2a < cs2;
      cs2 = u2a ? t2a : cs2; p_2 = u2a ? pc : p_2;
      const bool u2b = t2b < cs2;
      cs2 = u2b ? t2b : cs2; p_2 = u2b ? (pc | 8u) : p_2;
    }
    const bool v0 = cs0 < kThr, v1 = cs1 < kThr, v2 = cs2 < kThr;
    const unsigned long long ent = (unsigned long long)((p_0 | (v0 ? 4u : 0u)) | ((p_1 | (v1 ? 4u : 0u)) << 3) |
                                                        (((p_2 & 3u) | (v2 ? 4u : 0u) | (p_2 & 8u)) << 6));
    const int sh = (n % 6) * 10;
    if (n < 6) pw0 |= ent << sh; else if (n < 12) pw1 |= ent << sh; else pw2 |= ent << sh;
    {
      const uint16_t* eb = eob + band * 3;
      const long long s1 = cs1 + (n < 15 ? (long long)eb[1] * lam : 0ll);  // unreachable states stay >= kThr and never win
      const bool b1 = s1 < best_terminal;
      best_terminal = b1 ? s1 : best_terminal; best_last_n = b1 ? n : best_last_n; best_last_ctx = b1 ? 1 : best_last_ctx;
      const long long s2 = cs2 + (n < 15 ? (long long)eb[2] * lam : 0ll);
      const bool b2 = s2 < best_terminal;
      best_terminal = b2 ? s2 : best_terminal; best_last_n = b2 ? n : best_last_n; best_last_ctx = b2 ? 2 : best_last_ctx;
    }
    ps0 = v0 ? cs0 : kBig; ps1 = v1 ? cs1 : kBig; ps2 = v2 ? cs2 : kBig;
  }
  int ctx = best_last_ctx, last = 0;
  if (first == 1) io[0] = 0;
#pragma unroll 1
  for (int n = 15; n >= first; --n) {
    const int zig = c_zigzag[n];
    const unsigned long long w = n < 6 ? pw0 : (n < 12 ? pw1 : pw2);
    const uint32_t ent = (uint32_t)(w >> ((n % 6) * 10)) & 0x3ffu;
    const uint32_t e = ctx == 0 ? (ent & 7u) : (ctx == 1 ? ((ent >> 3) & 7u) : ((ent >> 6) & 15u));
    const bool take = n <= best_last_n && (e & 4u);
    const int L0 = io[zig];
    // level by landing ctx: ctx0 -> 0; ctx1 -> 1 (L0 if L0 == 1 else L0+1, both equal 1); ctx2 -> L0 or L0+1
    const int mag = ctx == 0 ? 0 : (ctx == 1 ? 1 : L0 + (int)((e >> 3) & 1u));
    const int lv = ((neg_mask >> zig) & 1u) ? -mag : mag;
    io[zig] = (int16_t)(take ? lv : 0);
    last = (take && lv != 0 && last == 0) ? n + 1 : last;
    ctx = take ? (int)(e & 3u) : ctx;
  }
  return last;
}

// TrellisQuantizeBlock again (same contract as trellis_block_smem: 16 int16 in shared memory, raster order, coefficients in,
// levels out, returns last + 1), organised for the latency of one call -- a macroblock's chain of dependent trellis calls is
// what a wave of the mode search waits for.  Everything about a position that does not depend on the Viterbi state (base
// level, which of the candidate levels {0, L0, L0+1} exist, their distortion and fixed-cost terms, where their table costs
// sit) is computed one position AHEAD, and the recurrence runs per candidate LEVEL (minimum over the three previous contexts)
// before the two non-zero levels scatter to their landing context min(level, 2).  A straightforward 64-bit version of that
// was ~275 SASS instructions per position, most of them 64-bit compare / select pairs and index bookkeeping.  Here every score is
// kept multiplied by 64, which frees the low six bits of the 64-bit word for the bookkeeping the comparisons have to carry:
//   * a transition key is (score * 64 | prev ctx): the minimum over the three previous contexts is then a plain signed
//     minimum, and equal scores resolve to the lower context exactly as the reference's strict '<' in loop order does;
//   * bit 3 marks "level L0 + 1" on the way into context 2, and key(L0+1) < key(L0) is the reference's order of visits
//     (prev ctx major, then level) on equal scores;
//   * the running best terminal carries 1 + 2n + (ctx - 1): a later terminal only wins when strictly smaller.
// rate * (lambda * 64) + score is one IMAD.WIDE.U32.  The survivor entries (10 bits per position) go through a 160-bit shift
// register of five 32-bit words (funnel shifts) instead of variable 64-bit shifts.  Scaled scores stay below 2^47.
WG_HD uint32_t wg_fshl(uint32_t lo, uint32_t hi, int s) {  // upper word of (hi:lo) << s, 0 < s < 32
#ifdef __CUDA_ARCH__
  return __funnelshift_l(lo, hi, s);
#else
  return (hi << s) | (lo >> (32 - s));
#endif
}
WG_HD uint32_t wg_fshr(uint32_t lo, uint32_t hi, int s) {  // lower word of (hi:lo) >> s, 0 < s < 32
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return (lo >> s) | (hi << (32 - s));
#endif
}
WG_HD long long wg_mad_wide(uint32_t a, uint32_t b, long long c) {  // a * b + c, one IMAD.WIDE.U32
#ifdef __CUDA_ARCH__
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((unsigned long long)c));
  return (long long)r;
#else
  return c + (long long)((unsigned long long)a * b);
#endif
}
struct TrellisPos3 {  // what a position contributes, independent of the Viterbi state
  int L0;
  uint32_t flags;      // 1: level L0 exists, 2: level L0+1 exists
  long long k0, k1;    // 64 * (fixed-cost * lambda + weighted distortion delta) of level L0 / L0+1
  int iA, iB;          // offsets of the cost-table entries of level L0 / L0+1 in the row of previous context 0
};
WG_HD void trellis_prep3(const int16_t* io, int n, int quant_dc, int quant_ac, unsigned iq_dc, unsigned iq_ac, const CostTabs& T,
                         uint32_t lam64, TrellisPos3& P) {
  const int zig = c_zigzag[n];
  const int coeff0 = io[zig];
  const int quant = (n == 0) ? quant_dc : quant_ac;
  const unsigned iquant = (n == 0) ? iq_dc : iq_ac;
  const int L0 = min((int)(((unsigned)coeff0 * iquant) >> 17), 2047);
  const int thresh = min((int)(((unsigned)coeff0 * iquant + 65536u) >> 17), 2047);
  P.L0 = L0;
  P.flags = ((L0 > 0 && L0 <= thresh) ? 1u : 0u) | ((L0 + 1 <= 2047 && L0 + 1 <= thresh) ? 2u : 0u);
  const int c0sq = coeff0 * coeff0;
  const int e0 = coeff0 - L0 * quant, e1 = coeff0 - (L0 + 1) * quant;
  const int wt = c_weight_trellis[zig] * (256 * 64);
  P.k0 = wg_mad_wide((uint32_t)lfc_at(T, L0), lam64, (long long)(e0 * e0 - c0sq) * wt);
  P.k1 = wg_mad_wide((uint32_t)lfc_at(T, min(L0 + 1, 2047)), lam64, (long long)(e1 * e1 - c0sq) * wt);
  P.iA = min(L0, LC_LEVELS - 1);
  P.iB = min(L0 + 1, LC_LEVELS - 1);
}
WG_HD long long trellis_min3(long long a, long long b, long long c) {
  const long long m = b < a ? b : a;
  return c < m ? c : m;
}
WG_HD int wg_hi32(long long v) { return (int)(v >> 32); }
WG_HD int trellis_block_v3(int16_t* io, const SegQuant& sq, int first, int type, int initial_ctx, int lambda, const CostTabs& T) {
  const int quant_ac = sq.quant, quant_dc = sq.dc_quant;
  const unsigned iq_ac = (unsigned)sq.iquant, iq_dc = (unsigned)sq.dc_iquant;
  uint32_t neg_mask = 0;  // bit i: raster coefficient i is negative
  {
    bool non_zero = false;  // all-zero pre-scan with neutral bias (encode_trellis.go:39-98)
    int c0[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int raw = io[i];
      c0[i] = max(abs(raw) + sq.sharpen[i], 0);
      neg_mask |= (raw < 0 ? 1u : 0u) << i;
      if (i > 0) non_zero |= (((unsigned)c0[i] * iq_ac) >> 17) > 0;
      else non_zero |= first == 0 && (((unsigned)c0[0] * iq_dc) >> 17) > 0;
    }
    if (!non_zero) {
#pragma unroll
      for (int i = 0; i < 16; ++i) io[i] = 0;
      return 0;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) io[i] = (int16_t)c0[i];
  }
  initial_ctx = min(initial_ctx, 2);
  const uint16_t* lc = T.lc + type * (8 * 3 * LC_LEVELS);
  const uint16_t* eob = T.eob + type * (8 * 3);
  const long long kBig = 1ll << 56, kMask = ~63ll;  // keys below 2^55 are reachable states
  long long ps0 = initial_ctx == 0 ? 0 : kBig, ps1 = initial_ctx == 1 ? 0 : kBig, ps2 = initial_ctx == 2 ? 0 : kBig;
  const uint32_t lam64 = (uint32_t)lambda << 6;
  long long best = (long long)((unsigned long long)eob[c_bands[first] * 3 + initial_ctx] * lam64);  // order tag 0: "no coefficient"
  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0;  // survivor entries, newest in the low 10 bits of w0
  const int kThrHi = 1 << 23;  // key < kThr  <=>  its upper word < 2^23 (scores may be negative)
  TrellisPos3 cur;
  trellis_prep3(io, first, quant_dc, quant_ac, iq_dc, iq_ac, T, lam64, cur);
#pragma unroll 1
  for (int n = first; n < 16; ++n) {
    // this position's table costs: addresses known since the previous iteration, issued first ...
    const int band = c_bands[n + 1];  // sic: the next position's band (encode_trellis.go:151)
    const uint16_t* row = lc + band * 3 * LC_LEVELS;
    const uint32_t r00 = row[0], r01 = row[LC_LEVELS], r02 = row[2 * LC_LEVELS];
    const uint32_t rA0 = row[cur.iA], rA1 = row[LC_LEVELS + cur.iA], rA2 = row[2 * LC_LEVELS + cur.iA];
    const uint32_t rB0 = row[cur.iB], rB1 = row[LC_LEVELS + cur.iB], rB2 = row[2 * LC_LEVELS + cur.iB];
    const uint32_t eb1 = (n < 15) ? eob[band * 3 + 1] : 0u, eb2 = (n < 15) ? eob[band * 3 + 2] : 0u;
    // ... then the next position's state-independent half while those loads are in flight
    TrellisPos3 nxt;
    trellis_prep3(io, min(n + 1, 15), quant_dc, quant_ac, iq_dc, iq_ac, T, lam64, nxt);
    const int L0 = cur.L0;
    io[c_zigzag[n]] = (int16_t)L0;  // the backtrack only needs the base level (position n + 1 was read above)
    const long long q0 = ps0, q1 = ps1 | 1, q2 = ps2 | 2;
    const long long key0 = trellis_min3(wg_mad_wide(r00, lam64, q0), wg_mad_wide(r01, lam64, q1), wg_mad_wide(r02, lam64, q2));
    const long long keyA = trellis_min3(wg_mad_wide(rA0, lam64, q0), wg_mad_wide(rA1, lam64, q1), wg_mad_wide(rA2, lam64, q2)) + cur.k0;
    const long long keyB = trellis_min3(wg_mad_wide(rB0, lam64, q0), wg_mad_wide(rB1, lam64, q1), wg_mad_wide(rB2, lam64, q2)) + cur.k1;
    const bool hasA = (cur.flags & 1u) != 0, hasB = (cur.flags & 2u) != 0;
    // level L0 lands in context min(L0, 2), level L0 + 1 in min(L0 + 1, 2)
    const bool a1 = hasA && L0 == 1, b1 = hasB && L0 == 0;  // at most one of them
    const bool a2 = hasA && L0 >= 2, b2 = hasB && L0 >= 1;
    const long long key1 = a1 ? keyA : (b1 ? keyB : kBig);
    const bool take_b = b2 && (!a2 || keyB < keyA);
    const long long key2 = take_b ? (keyB | 8) : (a2 ? keyA : kBig);
    const bool v0 = wg_hi32(key0) < kThrHi, v1 = wg_hi32(key1) < kThrHi, v2 = wg_hi32(key2) < kThrHi;
    const uint32_t l0 = (uint32_t)key0, l1 = (uint32_t)key1, l2 = (uint32_t)key2;
    const uint32_t ent = ((l0 & 3u) | (v0 ? 4u : 0u)) | (((l1 & 3u) | (v1 ? 4u : 0u)) << 3) | (((l2 & 11u) | (v2 ? 4u : 0u)) << 6);
    w4 = wg_fshl(w3, w4, 10); w3 = wg_fshl(w2, w3, 10); w2 = wg_fshl(w1, w2, 10); w1 = wg_fshl(w0, w1, 10); w0 = (w0 << 10) | ent;
    const long long s1 = key1 & kMask, s2 = key2 & kMask;
    {  // terminals: unreachable states stay >= kThr and never win
      const long long t1 = wg_mad_wide(eb1, lam64, s1) | (long long)(2 * n + 1);
      best = t1 < best ? t1 : best;
      const long long t2 = wg_mad_wide(eb2, lam64, s2) | (long long)(2 * n + 2);
      best = t2 < best ? t2 : best;
    }
    ps0 = v0 ? (key0 & kMask) : kBig; ps1 = v1 ? s1 : kBig; ps2 = v2 ? s2 : kBig;
    cur = nxt;
  }
  const int tag = (int)((uint32_t)best & 63u);
  const int best_last_n = tag ? (tag - 1) >> 1 : -1;
  int ctx = tag ? 1 + ((tag - 1) & 1) : -1, last = 0;
  if (first == 1) io[0] = 0;
#pragma unroll 1
  for (int n = 15; n >= first; --n) {
    const int zig = c_zigzag[n];
    const uint32_t ent = w0 & 0x3ffu;
    w0 = wg_fshr(w0, w1, 10); w1 = wg_fshr(w1, w2, 10); w2 = wg_fshr(w2, w3, 10); w3 = wg_fshr(w3, w4, 10); w4 >>= 10;
    const uint32_t e = ctx == 0 ? (ent & 7u) : (ctx == 1 ? ((ent >> 3) & 7u) : ((ent >> 6) & 15u));
    const bool take = n <= best_last_n && (e & 4u);
    const int L0 = io[zig];
    const int mag = ctx == 0 ? 0 : (ctx == 1 ? 1 : L0 + (int)((e >> 3) & 1u));
    const int lv = ((neg_mask >> zig) & 1u) ? -mag : mag;
    io[zig] = (int16_t)(take ? lv : 0);
    last = (take && lv != 0 && last == 0) ? n + 1 : last;
    ctx = take ? (int)(e & 3u) : ctx;
  }
  return last;
}

}  // namespace wg
