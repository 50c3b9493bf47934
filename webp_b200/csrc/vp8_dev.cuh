// Device-side VP8 primitives (sm_100a).  Register-resident 4x4 blocks: every routine works on
// int[16] in raster order so the transforms fully unroll into IADD3/IMAD/SHF with no local memory.
// Arithmetic follows the reference's scalar Go code (file:line cited per function); all
// intermediate ranges fit int32 (|coeff| <= 2048+sharpen, residual <= 255).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace wg {

#include "vp8_tables.inc"  // host-visible copies (static const); device copies below

// ---- tables in device global memory (L1/L2 cached); hot kernels stage them into shared memory
__device__ __constant__ uint8_t c_zigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
__device__ __constant__ uint8_t c_bands[17] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0};
__device__ __constant__ uint8_t c_rev_zigzag[16] = {0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15};
__device__ __constant__ uint8_t c_weight_y[16] = {38, 32, 20, 9, 32, 28, 17, 7, 20, 17, 10, 4, 9, 7, 4, 2};
__device__ __constant__ uint8_t c_weight_trellis[16] = {30, 27, 19, 11, 27, 24, 17, 10, 19, 17, 12, 8, 11, 10, 8, 6};

// Pointers to the cost tables a kernel uses (global or shared memory).
struct CostTabs {
  const uint16_t* ecost;   // VP8EntropyCost[256]               internal/dsp/cost.go:6
  const uint16_t* lfc;     // VP8LevelFixedCosts[2048]          internal/dsp/cost.go:33
  const uint16_t* lcodes;  // vp8LevelCodes[67][2]              internal/lossy/encode_quant.go:226
  const uint8_t* proba;    // [4][8][3][11] coefficient probas  internal/lossy/proba.go:45
};

// SegmentQuant (internal/lossy/encode.go:311)
struct SegQuant {
  int quant, iquant, bias, dc_quant, dc_iquant, dc_bias;
  int16_t sharpen[16];
};
// SegmentInfo subset used on the device (internal/lossy/encode.go:278)
struct SegParams {
  SegQuant y1, y2, uv;
  int lambda_i4, lambda_i16, lambda_uv, lambda_mode, tlambda_i4, tlambda_i16, tlambda_sd, pad;
};

__device__ __forceinline__ int clip8(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int mul1(int a) { return ((a * 20091) >> 16) + a; }  // transforms.go:20
__device__ __forceinline__ int mul2(int a) { return (a * 35468) >> 16; }        // transforms.go:25

// fTransform (internal/dsp/transforms.go:371): src - ref -> 16 coefficients
__device__ __forceinline__ void ftransform(const int* src, const int* ref, int* out) {
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
__device__ __forceinline__ void itransform(const int* ref, const int* in, int* dst) {
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
__device__ __forceinline__ void fwht(const int* in, int* out) {
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
__device__ __forceinline__ void iwht(const int* in, int* out) {
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
__device__ __forceinline__ int sse16(const int* a, const int* b) {  // ssim.go:188
  int s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int d = a[i] - b[i];
    s += d * d;
  }
  return s;
}
__device__ __forceinline__ int ttransform(const int* in) {  // ssim.go:266 with kWeightY
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
__device__ __forceinline__ int tdisto4x4(const int* a, const int* b) {  // ssim.go:315
  return abs(ttransform(b) - ttransform(a)) >> 5;
}

// PredLuma4Direct (internal/dsp/predict_lossy.go:185-451).  e[0]=top-left, e[1..8]=top[0..7], e[9..12]=left[0..3]
__device__ __forceinline__ int avg3(int a, int b, int c) { return (a + 2 * b + c + 2) >> 2; }
__device__ __forceinline__ int avg2(int a, int b) { return (a + b + 1) >> 1; }
__device__ __forceinline__ void pred4(int mode, const int* e, int* d) {
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
__device__ __forceinline__ int quantize_block(const int* in, int* out, const SegQuant& sq, int first) {
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
__device__ __forceinline__ void dequant_block(const int* in, int* out, const SegQuant& sq) {
  out[0] = (int)(int16_t)(in[0] * sq.dc_quant);
#pragma unroll
  for (int n = 1; n < 16; ++n) out[n] = (int)(int16_t)(in[n] * sq.quant);
}

// variableLevelCost (encode_quant.go:248)
__device__ __forceinline__ int variable_level_cost(int level, const uint8_t* p, const CostTabs& T) {
  int idx = min(level - 1, 66);
  int pattern = T.lcodes[2 * idx], bits = T.lcodes[2 * idx + 1];
  int cost = 0;
  for (int i = 2; pattern; ++i) {
    if (pattern & 1) cost += T.ecost[(bits & 1) ? 255 - p[i] : p[i]];
    bits >>= 1;
    pattern >>= 1;
  }
  return cost;
}
// fastVariableLevelCost (encode_trellis.go:328) -- same values, short paths for 1..4
__device__ __forceinline__ int fast_variable_level_cost(int level, const uint8_t* p, const CostTabs& T) {
  const uint16_t* e = T.ecost;
  switch (level) {
    case 1: return e[p[2]];
    case 2: return e[255 - p[2]] + e[p[3]] + e[p[4]];
    case 3: return e[255 - p[2]] + e[p[3]] + e[255 - p[4]] + e[p[5]];
    case 4: return e[255 - p[2]] + e[p[3]] + e[255 - p[4]] + e[255 - p[5]];
    default: return variable_level_cost(level, p, T);
  }
}
// TokenCostForCoeffs (encode_quant.go:170); levels in raster order
__device__ __forceinline__ int token_cost(const int* lev, int nz_count, int type, int ctx0, int first, const CostTabs& T) {
  const uint8_t* pt = T.proba + type * (8 * 3 * 11);
  if (nz_count <= first) return T.ecost[pt[(c_bands[first] * 3 + ctx0) * 11]];
  const int last = nz_count - 1;
  int cost = 0, ctx = ctx0;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    if (n < first) continue;
    const uint8_t* pp = pt + (c_bands[n] * 3 + ctx) * 11;
    if (n > last) { cost += T.ecost[pp[0]]; break; }
    const int v = abs(lev[c_zigzag[n]]);
    cost += T.ecost[255 - pp[0]];
    if (v == 0) {
      cost += T.ecost[pp[1]];
      ctx = 0;
    } else {
      cost += T.ecost[255 - pp[1]] + T.lfc[v] + fast_variable_level_cost(v, pp, T);
      ctx = (v == 1) ? 1 : 2;
    }
  }
  return cost;
}

// TrellisQuantizeBlock (internal/lossy/encode_trellis.go:23-324).  in/out raster order.
// Path storage is packed: per position 3 x (level int16, prev_ctx/valid byte).
__device__ __noinline__ int trellis_block(const int* in, int* out, const SegQuant& sq, int first, int type,
                                          int initial_ctx, int lambda, const CostTabs& T) {
  {  // all-zero pre-scan with neutral bias (encode_trellis.go:39-98)
    bool non_zero = false;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      if (n < first) continue;
      const int zig = c_zigzag[n];
      int c = abs(in[zig]) + sq.sharpen[zig];
      if (c < 0) c = 0;
      const int iq = (n == 0) ? sq.dc_iquant : sq.iquant;
      non_zero |= (((unsigned)c * (unsigned)iq) >> 17) > 0;  // c*iq < 2^31 (c<=~4200, iq<=32768)
    }
    if (!non_zero) {
#pragma unroll
      for (int i = 0; i < 16; ++i) out[i] = 0;
      return 0;
    }
  }
  if (initial_ctx > 2) initial_ctx = 2;
  const uint8_t* pt = T.proba + type * (8 * 3 * 11);
  const long long kMaxScore = 1ll << 60;
  long long ps[3];  // prev scores; invalid = kInvalid
  const long long kInvalid = 0x7fffffffffffffffll;
  ps[0] = ps[1] = ps[2] = kInvalid;
  ps[initial_ctx] = 0;
  short path_level[16][3];
  signed char path_prev[16][3];  // -1 = invalid
  const int skip_rate = T.ecost[pt[(c_bands[first] * 3 + initial_ctx) * 11]];
  long long best_terminal = (long long)skip_rate * lambda;
  int best_last_n = -1, best_last_ctx = -1;
  const long long lam = lambda;
  for (int n = first; n < 16; ++n) {
    const int zig = c_zigzag[n];
    const int band = c_bands[n + 1];  // sic (encode_trellis.go:151)
    int raw = in[zig];
    const bool neg = raw < 0;
    raw = abs(raw);
    int coeff0 = raw + sq.sharpen[zig];
    if (coeff0 < 0) coeff0 = 0;
    const int quant = (n == 0) ? sq.dc_quant : sq.quant;
    const int iquant = (n == 0) ? sq.dc_iquant : sq.iquant;
    int L0 = (int)(((unsigned)coeff0 * (unsigned)iquant) >> 17);
    L0 = min(L0, 2047);
    int thresh_level = (int)(((unsigned)coeff0 * (unsigned)iquant + 65536u) >> 17);
    thresh_level = min(thresh_level, 2047);
    const long long weight = c_weight_trellis[zig];
    const long long coeff0sq = (long long)coeff0 * coeff0;
    const uint8_t* band_probas = pt + band * 33;
    long long cs[3] = {kMaxScore, kMaxScore, kMaxScore};
    bool cv[3] = {false, false, false};
    short cl[3] = {0, 0, 0};
    signed char cp[3] = {-1, -1, -1};
    const bool has_l0 = L0 > 0 && L0 <= thresh_level;
    const bool has_l1 = L0 + 1 <= 2047 && L0 + 1 <= thresh_level;
    long long disto_l0 = 0, disto_l1 = 0;
    int next_ctx0 = 0, next_ctx1 = 0, fixed_l0 = 0, fixed_l1 = 0;
    if (has_l0) {
      const long long e = coeff0 - L0 * quant;
      disto_l0 = 256 * (weight * (e * e - coeff0sq));
      next_ctx0 = min(L0, 2);
      fixed_l0 = T.lfc[L0];
    }
    if (has_l1) {
      const long long e = coeff0 - (L0 + 1) * quant;
      disto_l1 = 256 * (weight * (e * e - coeff0sq));
      next_ctx1 = min(L0 + 1, 2);
      fixed_l1 = T.lfc[L0 + 1];
    }
    const short sl0 = (short)(neg ? -L0 : L0), sl1 = (short)(neg ? -(L0 + 1) : (L0 + 1));
#pragma unroll
    for (int pc = 0; pc < 3; ++pc) {
      if (ps[pc] == kInvalid) continue;
      const long long prev_score = ps[pc];
      const uint8_t* p = band_probas + pc * 11;
      const int not_eob = T.ecost[255 - p[0]];
      const int rate0 = not_eob + T.ecost[p[1]];
      const long long total = prev_score + (long long)rate0 * lam;
      if (!cv[0] || total < cs[0]) { cs[0] = total; cl[0] = 0; cp[0] = (signed char)pc; cv[0] = true; }
      if (has_l0 || has_l1) {
        const int non_zero = not_eob + T.ecost[255 - p[1]];
        if (has_l0) {
          const int rate = non_zero + fixed_l0 + fast_variable_level_cost(L0, p, T);
          const long long ts = prev_score + (long long)rate * lam + disto_l0;
          // next_ctx0 in {1,2}
          if (next_ctx0 == 1) { if (!cv[1] || ts < cs[1]) { cs[1] = ts; cl[1] = sl0; cp[1] = (signed char)pc; cv[1] = true; } }
          else               { if (!cv[2] || ts < cs[2]) { cs[2] = ts; cl[2] = sl0; cp[2] = (signed char)pc; cv[2] = true; } }
        }
        if (has_l1) {
          const int rate = non_zero + fixed_l1 + fast_variable_level_cost(L0 + 1, p, T);
          const long long ts = prev_score + (long long)rate * lam + disto_l1;
          if (next_ctx1 == 1) { if (!cv[1] || ts < cs[1]) { cs[1] = ts; cl[1] = sl1; cp[1] = (signed char)pc; cv[1] = true; } }
          else               { if (!cv[2] || ts < cs[2]) { cs[2] = ts; cl[2] = sl1; cp[2] = (signed char)pc; cv[2] = true; } }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      path_level[n][c] = cl[c];
      path_prev[n][c] = cv[c] ? cp[c] : (signed char)-1;
    }
#pragma unroll
    for (int c = 1; c < 3; ++c) {
      if (!cv[c]) continue;
      long long eob_score = cs[c];
      if (n < 15) eob_score += (long long)T.ecost[band_probas[c * 11]] * lam;
      if (eob_score < best_terminal) { best_terminal = eob_score; best_last_n = n; best_last_ctx = c; }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) ps[c] = cv[c] ? cs[c] : kInvalid;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) out[i] = 0;
  if (best_last_n < 0) return 0;
  int ctx = best_last_ctx, last = 0;
  for (int n = best_last_n; n >= first; --n) {
    if (path_prev[n][ctx] >= 0) {
      const int zig = c_zigzag[n];
      const int lv = path_level[n][ctx];
      // out[] is register-resident: select by unrolled compare instead of a dynamic index
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i == zig) out[i] = lv;
      if (lv != 0 && last == 0) last = n + 1;
      ctx = path_prev[n][ctx];
    }
  }
  return last;
}

}  // namespace wg
