// ORACLE -- TEST INFRASTRUCTURE ONLY (see vp8_common.h).
// Restates internal/dsp of the reference: 4x4 transforms, intra predictors, SSE/TDisto,
// loop-filter taps, YUV<->RGB, fancy upsampler, SSIM/PSNR.  Work buffers use BPS=32 stride.
#pragma once
#include "vp8_common.h"
#include <math.h>

namespace orc {

// ---------------------------------------------------------------- transforms.go
static inline int mul1(int a) { return ((a * 20091) >> 16) + a; }  // transforms.go:20
static inline int mul2(int a) { return (a * 35468) >> 16; }        // transforms.go:25

// iTransformOne (transforms.go:265): dst = clip(ref + IDCT(in)); ref/dst BPS-strided, may alias.
static inline void itransform_one(const uint8_t* ref, const int16_t* in, uint8_t* dst) {
  ORC_COUNT(OP_ITRANSFORM, 1);
  int tmp[16];
  for (int i = 0; i < 4; ++i) {  // vertical pass, column i
    const int a = in[i] + in[8 + i];
    const int b = in[i] - in[8 + i];
    const int c = mul2(in[4 + i]) - mul1(in[12 + i]);
    const int d = mul1(in[4 + i]) + mul2(in[12 + i]);
    tmp[i] = a + d;
    tmp[4 + i] = b + c;
    tmp[8 + i] = b - c;
    tmp[12 + i] = a - d;
  }
  for (int j = 0; j < 4; ++j) {  // horizontal pass, row j
    const int dc = tmp[4 * j] + 4;
    const int a = dc + tmp[4 * j + 2];
    const int b = dc - tmp[4 * j + 2];
    const int c = mul2(tmp[4 * j + 1]) - mul1(tmp[4 * j + 3]);
    const int d = mul1(tmp[4 * j + 1]) + mul2(tmp[4 * j + 3]);
    const uint8_t r0 = ref[j * BPS + 0], r1 = ref[j * BPS + 1], r2 = ref[j * BPS + 2], r3 = ref[j * BPS + 3];
    dst[j * BPS + 0] = clip8(r0 + ((a + d) >> 3));
    dst[j * BPS + 1] = clip8(r1 + ((b + c) >> 3));
    dst[j * BPS + 2] = clip8(r2 + ((b - c) >> 3));
    dst[j * BPS + 3] = clip8(r3 + ((a - d) >> 3));
  }
}
// Decoder transforms (transforms.go:37-216). transformOne == itransform_one with ref==dst.
static inline void transform_one(const int16_t* in, uint8_t* dst) { itransform_one(dst, in, dst); }
static inline void transform_dc(const int16_t* in, uint8_t* dst) {  // transforms.go:148 / decode_frame.go:31
  const int add = (in[0] + 4) >> 3;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) dst[j * BPS + i] = clip8(dst[j * BPS + i] + add);
}
static inline void transform_ac3(const int16_t* in, uint8_t* dst) {  // transforms.go:170
  const int a = in[0] + 4;
  const int c4 = mul2(in[4]), d4 = mul1(in[4]);
  const int c1 = mul2(in[1]), d1 = mul1(in[1]);
  const int rowv[4] = {a + d4, a + c4, a - c4, a - d4};
  for (int j = 0; j < 4; ++j) {
    const int dc = rowv[j];
    dst[j * BPS + 0] = clip8(dst[j * BPS + 0] + ((dc + d1) >> 3));
    dst[j * BPS + 1] = clip8(dst[j * BPS + 1] + ((dc + c1) >> 3));
    dst[j * BPS + 2] = clip8(dst[j * BPS + 2] + ((dc - c1) >> 3));
    dst[j * BPS + 3] = clip8(dst[j * BPS + 3] + ((dc - d1) >> 3));
  }
}
// transformWHT (transforms.go:223): out has stride 16 between DCs.
static inline void transform_wht(const int16_t* in, int16_t* out) {
  ORC_COUNT(OP_IWHT, 1);
  int tmp[16];
  for (int i = 0; i < 4; ++i) {
    const int a0 = in[0 + i] + in[12 + i];
    const int a1 = in[4 + i] + in[8 + i];
    const int a2 = in[4 + i] - in[8 + i];
    const int a3 = in[0 + i] - in[12 + i];
    tmp[0 + i] = a0 + a1;
    tmp[8 + i] = a0 - a1;
    tmp[4 + i] = a3 + a2;
    tmp[12 + i] = a3 - a2;
  }
  for (int i = 0; i < 4; ++i) {
    const int dc = tmp[i * 4 + 0] + 3;
    const int a0 = dc + tmp[i * 4 + 3];
    const int a1 = tmp[i * 4 + 1] + tmp[i * 4 + 2];
    const int a2 = tmp[i * 4 + 1] - tmp[i * 4 + 2];
    const int a3 = dc - tmp[i * 4 + 3];
    const int base = i * 4 * 16;
    out[base + 0 * 16] = (int16_t)((a0 + a1) >> 3);
    out[base + 1 * 16] = (int16_t)((a3 + a2) >> 3);
    out[base + 2 * 16] = (int16_t)((a0 - a1) >> 3);
    out[base + 3 * 16] = (int16_t)((a3 - a2) >> 3);
  }
}
// fTransform (transforms.go:371)
static inline void ftransform(const uint8_t* src, const uint8_t* ref, int16_t* out) {
  ORC_COUNT(OP_FTRANSFORM, 1);
  int tmp[16];
  for (int j = 0; j < 4; ++j) {
    const int d0 = src[j * BPS + 0] - ref[j * BPS + 0];
    const int d1 = src[j * BPS + 1] - ref[j * BPS + 1];
    const int d2 = src[j * BPS + 2] - ref[j * BPS + 2];
    const int d3 = src[j * BPS + 3] - ref[j * BPS + 3];
    const int a0 = d0 + d3, a1 = d1 + d2, a2 = d1 - d2, a3 = d0 - d3;
    tmp[4 * j + 0] = (a0 + a1) * 8;
    tmp[4 * j + 1] = (a2 * 2217 + a3 * 5352 + 1812) >> 9;
    tmp[4 * j + 2] = (a0 - a1) * 8;
    tmp[4 * j + 3] = (a3 * 2217 - a2 * 5352 + 937) >> 9;
  }
  for (int i = 0; i < 4; ++i) {
    const int a0 = tmp[0 + i] + tmp[12 + i];
    const int a1 = tmp[4 + i] + tmp[8 + i];
    const int a2 = tmp[4 + i] - tmp[8 + i];
    const int a3 = tmp[0 + i] - tmp[12 + i];
    out[0 + i] = (int16_t)((a0 + a1 + 7) >> 4);
    out[4 + i] = (int16_t)(((a2 * 2217 + a3 * 5352 + 12000) >> 16) + (a3 != 0));
    out[8 + i] = (int16_t)((a0 - a1 + 7) >> 4);
    out[12 + i] = (int16_t)((a3 * 2217 - a2 * 5352 + 51000) >> 16);
  }
}
// fTransformWHT (transforms.go:500): flat 4x4 DC array in, 16 out.
static inline void ftransform_wht(const int16_t* in, int16_t* out) {
  ORC_COUNT(OP_FWHT, 1);
  int tmp[16];
  for (int i = 0; i < 4; ++i) {
    const int a0 = in[i * 4 + 0] + in[i * 4 + 2];
    const int a1 = in[i * 4 + 1] + in[i * 4 + 3];
    const int a2 = in[i * 4 + 1] - in[i * 4 + 3];
    const int a3 = in[i * 4 + 0] - in[i * 4 + 2];
    tmp[0 + i * 4] = a0 + a1;
    tmp[1 + i * 4] = a3 + a2;
    tmp[2 + i * 4] = a3 - a2;
    tmp[3 + i * 4] = a0 - a1;
  }
  for (int i = 0; i < 4; ++i) {
    const int a0 = tmp[0 + i] + tmp[8 + i];
    const int a1 = tmp[4 + i] + tmp[12 + i];
    const int a2 = tmp[4 + i] - tmp[12 + i];
    const int a3 = tmp[0 + i] - tmp[8 + i];
    out[0 + i] = (int16_t)((a0 + a1) >> 1);
    out[4 + i] = (int16_t)((a3 + a2) >> 1);
    out[8 + i] = (int16_t)((a3 - a2) >> 1);
    out[12 + i] = (int16_t)((a0 - a1) >> 1);
  }
}

// ---------------------------------------------------------------- predict_lossy.go
static inline uint8_t avg3(int a, int b, int c) { return (uint8_t)((a + 2 * b + c + 2) >> 2); }
static inline uint8_t avg2(int a, int b) { return (uint8_t)((a + b + 1) >> 1); }
static inline void fill_block(uint8_t* d, int size, int v) {
  for (int j = 0; j < size; ++j) memset(d + j * BPS, v, size);
}
// mode: 0 DC, 1 TM, 2 VE, 3 HE, 4 DC-noTop, 5 DC-noLeft, 6 DC-noTopLeft (predict_lossy.go:27-181)
static inline void pred_square(int mode, uint8_t* buf, int off, int size) {
  uint8_t* d = buf + off;
  const int shift = (size == 16) ? 4 : 3;  // log2(size)
  switch (mode) {
    case 0: {
      int dc = 0;
      ORC_COUNT(OP_PRED_DC_SUM, 2 * size + 2);
      for (int i = 0; i < size; ++i) dc += d[i - BPS] + d[-1 + i * BPS];
      fill_block(d, size, (dc + size) >> (shift + 1));
    } break;
    case 1: {
      const int tl = d[-1 - BPS];
      ORC_COUNT(OP_PRED_TM_PIXEL, size * size);
      for (int j = 0; j < size; ++j) {
        const int base = d[-1 + j * BPS] - tl;
        for (int i = 0; i < size; ++i) d[i + j * BPS] = clip8(base + d[i - BPS]);
      }
    } break;
    case 2:
      for (int j = 0; j < size; ++j) memcpy(d + j * BPS, d - BPS, size);
      break;
    case 3:
      for (int j = 0; j < size; ++j) memset(d + j * BPS, d[-1 + j * BPS], size);
      break;
    case 4: {
      int dc = 0;
      ORC_COUNT(OP_PRED_DC_SUM, size + 2);
      for (int i = 0; i < size; ++i) dc += d[-1 + i * BPS];
      fill_block(d, size, (dc + (size >> 1)) >> shift);
    } break;
    case 5: {
      int dc = 0;
      ORC_COUNT(OP_PRED_DC_SUM, size + 2);
      for (int i = 0; i < size; ++i) dc += d[i - BPS];
      fill_block(d, size, (dc + (size >> 1)) >> shift);
    } break;
    default:
      fill_block(d, size, 128);
  }
}
static inline void pred_luma16(int mode, uint8_t* buf, int off) { pred_square(mode, buf, off, 16); }
static inline void pred_chroma8(int mode, uint8_t* buf, int off) { pred_square(mode, buf, off, 8); }

// PredLuma4Direct (predict_lossy.go:185-451)
static inline void pred_luma4(int mode, uint8_t* buf, int off) {
  ORC_COUNT(OP_PRED4, 1);
  uint8_t* d = buf + off;
#define DST(x, y) d[(x) + (y)*BPS]
  const int tl = d[-1 - BPS];
  const int t0 = d[0 - BPS], t1 = d[1 - BPS], t2 = d[2 - BPS], t3 = d[3 - BPS];
  const int l0 = d[-1], l1 = d[-1 + BPS], l2 = d[-1 + 2 * BPS], l3 = d[-1 + 3 * BPS];
  switch (mode) {
    case B_DC_PRED: {
      const int dc = t0 + t1 + t2 + t3 + l0 + l1 + l2 + l3;
      fill_block(d, 4, (dc + 4) >> 3);
    } break;
    case B_TM_PRED:
      for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) DST(i, j) = clip8(d[-1 + j * BPS] + d[i - BPS] - tl);
      break;
    case B_VE_PRED: {
      const int t4 = d[4 - BPS];
      const uint8_t v[4] = {avg3(tl, t0, t1), avg3(t0, t1, t2), avg3(t1, t2, t3), avg3(t2, t3, t4)};
      for (int j = 0; j < 4; ++j) memcpy(d + j * BPS, v, 4);
    } break;
    case B_HE_PRED: {
      const uint8_t v[4] = {avg3(tl, l0, l1), avg3(l0, l1, l2), avg3(l1, l2, l3), avg3(l2, l3, l3)};
      for (int j = 0; j < 4; ++j) memset(d + j * BPS, v[j], 4);
    } break;
    case B_RD_PRED:
      DST(0, 3) = avg3(l3, l2, l1);
      DST(0, 2) = DST(1, 3) = avg3(l2, l1, l0);
      DST(0, 1) = DST(1, 2) = DST(2, 3) = avg3(l1, l0, tl);
      DST(0, 0) = DST(1, 1) = DST(2, 2) = DST(3, 3) = avg3(l0, tl, t0);
      DST(1, 0) = DST(2, 1) = DST(3, 2) = avg3(tl, t0, t1);
      DST(2, 0) = DST(3, 1) = avg3(t0, t1, t2);
      DST(3, 0) = avg3(t1, t2, t3);
      break;
    case B_VR_PRED:
      DST(0, 0) = DST(1, 2) = avg2(tl, t0);
      DST(1, 0) = DST(2, 2) = avg2(t0, t1);
      DST(2, 0) = DST(3, 2) = avg2(t1, t2);
      DST(3, 0) = avg2(t2, t3);
      DST(0, 1) = DST(1, 3) = avg3(l0, tl, t0);
      DST(1, 1) = DST(2, 3) = avg3(tl, t0, t1);
      DST(2, 1) = DST(3, 3) = avg3(t0, t1, t2);
      DST(3, 1) = avg3(t1, t2, t3);
      DST(0, 2) = avg3(l1, l0, tl);
      DST(0, 3) = avg3(l2, l1, l0);
      break;
    case B_LD_PRED: {
      const int A = t0, B = t1, C = t2, D = t3, E = d[4 - BPS], F = d[5 - BPS], G = d[6 - BPS], H = d[7 - BPS];
      DST(0, 0) = avg3(A, B, C);
      DST(1, 0) = DST(0, 1) = avg3(B, C, D);
      DST(2, 0) = DST(1, 1) = DST(0, 2) = avg3(C, D, E);
      DST(3, 0) = DST(2, 1) = DST(1, 2) = DST(0, 3) = avg3(D, E, F);
      DST(3, 1) = DST(2, 2) = DST(1, 3) = avg3(E, F, G);
      DST(3, 2) = DST(2, 3) = avg3(F, G, H);
      DST(3, 3) = avg3(G, H, H);
    } break;
    case B_VL_PRED: {
      const int A = t0, B = t1, C = t2, D = t3, E = d[4 - BPS], F = d[5 - BPS], G = d[6 - BPS], H = d[7 - BPS];
      DST(0, 0) = avg2(A, B);
      DST(1, 0) = DST(0, 2) = avg2(B, C);
      DST(2, 0) = DST(1, 2) = avg2(C, D);
      DST(3, 0) = DST(2, 2) = avg2(D, E);
      DST(0, 1) = avg3(A, B, C);
      DST(1, 1) = DST(0, 3) = avg3(B, C, D);
      DST(2, 1) = DST(1, 3) = avg3(C, D, E);
      DST(3, 1) = DST(2, 3) = avg3(D, E, F);
      DST(3, 2) = avg3(E, F, G);
      DST(3, 3) = avg3(F, G, H);
    } break;
    case B_HD_PRED:
      DST(0, 0) = DST(2, 1) = avg2(tl, l0);
      DST(1, 0) = DST(3, 1) = avg3(l0, tl, t0);
      DST(2, 0) = avg3(tl, t0, t1);
      DST(3, 0) = avg3(t0, t1, t2);
      DST(0, 1) = DST(2, 2) = avg2(l0, l1);
      DST(1, 1) = DST(3, 2) = avg3(tl, l0, l1);
      DST(0, 2) = DST(2, 3) = avg2(l1, l2);
      DST(1, 2) = DST(3, 3) = avg3(l0, l1, l2);
      DST(0, 3) = avg2(l2, l3);
      DST(1, 3) = avg3(l1, l2, l3);
      break;
    case B_HU_PRED:
      DST(0, 0) = avg2(l0, l1);
      DST(1, 0) = avg3(l0, l1, l2);
      DST(2, 0) = DST(0, 1) = avg2(l1, l2);
      DST(3, 0) = DST(1, 1) = avg3(l1, l2, l3);
      DST(2, 1) = DST(0, 2) = avg2(l2, l3);
      DST(3, 1) = DST(1, 2) = avg3(l2, l3, l3);
      DST(2, 2) = DST(3, 2) = DST(0, 3) = DST(1, 3) = DST(2, 3) = DST(3, 3) = (uint8_t)l3;
      break;
  }
#undef DST
}

// ---------------------------------------------------------------- ssim.go:188-335
static inline int sse4x4(const uint8_t* a, const uint8_t* b) {
  ORC_COUNT(OP_SSE4X4, 1);
  int s = 0;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      const int d = a[i + j * BPS] - b[i + j * BPS];
      s += d * d;
    }
  return s;
}
static inline int sse16x16(const uint8_t* a, const uint8_t* b) {
  ORC_COUNT(OP_SSE4X4, 16);
  int s = 0;
  for (int j = 0; j < 16; ++j)
    for (int i = 0; i < 16; ++i) {
      const int d = a[i + j * BPS] - b[i + j * BPS];
      s += d * d;
    }
  return s;
}
static const uint16_t kWeightY[16] = {38, 32, 20, 9, 32, 28, 17, 7, 20, 17, 10, 4, 9, 7, 4, 2};
static inline int ttransform(const uint8_t* in, const uint16_t* w) {  // ssim.go:266
  ORC_COUNT(OP_TTRANSFORM, 1);
  int tmp[16];
  for (int i = 0; i < 4; ++i) {
    const uint8_t* p = in + i * BPS;
    const int a0 = p[0] + p[2], a1 = p[1] + p[3], a2 = p[1] - p[3], a3 = p[0] - p[2];
    tmp[0 + i * 4] = a0 + a1;
    tmp[1 + i * 4] = a3 + a2;
    tmp[2 + i * 4] = a3 - a2;
    tmp[3 + i * 4] = a0 - a1;
  }
  int sum = 0;
  for (int i = 0; i < 4; ++i) {
    const int a0 = tmp[0 + i] + tmp[8 + i];
    const int a1 = tmp[4 + i] + tmp[12 + i];
    const int a2 = tmp[4 + i] - tmp[12 + i];
    const int a3 = tmp[0 + i] - tmp[8 + i];
    const int b0 = a0 + a1, b1 = a3 + a2, b2 = a3 - a2, b3 = a0 - a1;
    sum += w[0 + i] * abs(b0) + w[4 + i] * abs(b1) + w[8 + i] * abs(b2) + w[12 + i] * abs(b3);
  }
  return sum;
}
static inline int tdisto4x4(const uint8_t* a, const uint8_t* b) {  // ssim.go:315
  const int d = ttransform(b, kWeightY) - ttransform(a, kWeightY);
  return abs(d) >> 5;
}
static inline int tdisto16x16(const uint8_t* a, const uint8_t* b) {  // ssim.go:327
  int d = 0;
  for (int y = 0; y < 16 * BPS; y += 4 * BPS)
    for (int x = 0; x < 16; x += 4) d += tdisto4x4(a + x + y, b + x + y);
  return d;
}

// ---------------------------------------------------------------- filter.go / decode_frame.go:360-558
static inline int sclip1(int v) { return v < -128 ? -128 : (v > 127 ? 127 : v); }  // cliptables.go:25
static inline int sclip2(int v) { return v < -16 ? -16 : (v > 15 ? 15 : v); }      // cliptables.go:28
static inline void do_filter2(uint8_t* p, int step) {
  const int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
  const int a = 3 * (q0 - p0) + sclip1(p1 - q1);
  const int a1 = sclip2((a + 4) >> 3);
  const int a2 = sclip2((a + 3) >> 3);
  p[-step] = clip8(p0 + a2);
  p[0] = clip8(q0 - a1);
}
static inline void do_filter4(uint8_t* p, int step) {
  const int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
  const int a = 3 * (q0 - p0);
  const int a1 = sclip2((a + 4) >> 3);
  const int a2 = sclip2((a + 3) >> 3);
  const int a3 = (a1 + 1) >> 1;
  p[-2 * step] = clip8(p1 + a3);
  p[-step] = clip8(p0 + a2);
  p[0] = clip8(q0 - a1);
  p[step] = clip8(q1 - a3);
}
static inline void do_filter6(uint8_t* p, int step) {
  const int p2 = p[-3 * step], p1 = p[-2 * step], p0 = p[-step];
  const int q0 = p[0], q1 = p[step], q2 = p[2 * step];
  const int a = sclip1(3 * (q0 - p0) + sclip1(p1 - q1));
  const int a1 = (27 * a + 63) >> 7;
  const int a2 = (18 * a + 63) >> 7;
  const int a3 = (9 * a + 63) >> 7;
  p[-3 * step] = clip8(p2 + a3);
  p[-2 * step] = clip8(p1 + a2);
  p[-step] = clip8(p0 + a1);
  p[0] = clip8(q0 - a1);
  p[step] = clip8(q1 - a2);
  p[2 * step] = clip8(q2 - a3);
}
static inline bool needs_filter(const uint8_t* p, int step, int t) {
  const int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
  return 4 * abs(p0 - q0) + abs(p1 - q1) <= t;
}
static inline bool needs_filter2(const uint8_t* p, int step, int t, int it) {  // decode_frame.go:484
  const int p3 = p[-4 * step], p2 = p[-3 * step], p1 = p[-2 * step], p0 = p[-step];
  const int q0 = p[0], q1 = p[step], q2 = p[2 * step], q3 = p[3 * step];
  if (4 * abs(p0 - q0) + abs(p1 - q1) > t) return false;
  return abs(p3 - p2) <= it && abs(p2 - p1) <= it && abs(p1 - p0) <= it && abs(q3 - q2) <= it &&
         abs(q2 - q1) <= it && abs(q1 - q0) <= it;
}
static inline bool hev(const uint8_t* p, int step, int thresh) {
  const int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
  return abs(p1 - p0) > thresh || abs(q1 - q0) > thresh;
}
// simple filter over `size` samples: hstride = step across the edge, vstride = step along it
static inline void simple_filter(uint8_t* p, int hstride, int vstride, int size, int thresh) {
  const int thresh2 = 2 * thresh + 1;
  for (int i = 0; i < size; ++i, p += vstride)
    if (needs_filter(p, hstride, thresh2)) do_filter2(p, hstride);
}
// FilterLoop26 (MB edges) / FilterLoop24 (inner edges): decode_frame.go:387-454
static inline void filter_loop26(uint8_t* p, int hstride, int vstride, int size, int thresh, int ithresh,
                                 int hev_t) {
  const int thresh2 = 2 * thresh + 1;
  for (int i = 0; i < size; ++i, p += vstride) {
    if (!needs_filter2(p, hstride, thresh2, ithresh)) continue;
    if (hev(p, hstride, hev_t)) do_filter2(p, hstride); else do_filter6(p, hstride);
  }
}
static inline void filter_loop24(uint8_t* p, int hstride, int vstride, int size, int thresh, int ithresh,
                                 int hev_t) {
  const int thresh2 = 2 * thresh + 1;
  for (int i = 0; i < size; ++i, p += vstride) {
    if (!needs_filter2(p, hstride, thresh2, ithresh)) continue;
    if (hev(p, hstride, hev_t)) do_filter2(p, hstride); else do_filter4(p, hstride);
  }
}

// ---------------------------------------------------------------- yuv.go
static inline int mult_hi(int v, int coeff) { return (v * coeff) >> 8; }  // yuv.go:38
static inline uint8_t yuv_clip(int v) {  // yuv.go:71-104: clip to [0, (256<<6)-1] then >>6
  return (v < 0) ? 0 : (v > 16383) ? 255 : (uint8_t)(v >> 6);
}
static inline uint8_t yuv_to_r(int y, int v) { return yuv_clip(mult_hi(y, 19077) + mult_hi(v, 26149) - 14234); }
static inline uint8_t yuv_to_g(int y, int u, int v) {
  return yuv_clip(mult_hi(y, 19077) - mult_hi(u, 6419) - mult_hi(v, 13320) + 8708);
}
static inline uint8_t yuv_to_b(int y, int u) { return yuv_clip(mult_hi(y, 19077) + mult_hi(u, 33050) - 17685); }

static inline uint8_t rgb_to_y(int r, int g, int b) {  // yuv.go:151
  return (uint8_t)((16839 * r + 33059 * g + 6420 * b + (1 << 15) + (16 << 16)) >> 16);
}
static inline uint8_t clip_uv(int uv, int rounding) {  // yuv.go:138
  uv = (uv + rounding + (128 << 18)) >> 18;
  if ((uv & ~0xff) == 0) return (uint8_t)uv;
  return uv < 0 ? 0 : 255;
}
static inline uint8_t rgb_to_u(int r, int g, int b, int rounding) {
  return clip_uv(-9719 * r - 19081 * g + 28800 * b, rounding);
}
static inline uint8_t rgb_to_v(int r, int g, int b, int rounding) {
  return clip_uv(28800 * r - 24116 * g - 4684 * b, rounding);
}

// Gamma tables (yuv.go:176-249): kGamma=0.80, 12-bit linear, 32-entry interpolation table.
struct GammaTables {
  uint32_t gamma_to_linear[256];
  uint32_t linear_to_gamma[34];
  GammaTables() {
    for (int i = 0; i < 256; ++i) {
      const double v = (double)i / 255.0;
      const double lin = (v <= 0 ? 0.0 : pow(v, 0.80)) * 4095.0;
      gamma_to_linear[i] = (uint32_t)(lin + 0.5);
    }
    const double scale = 128.0 / 4095.0;
    for (int i = 0; i <= 32; ++i) {
      const double v = scale * (double)i;
      const double g = (v <= 0 ? 0.0 : pow(v, 1.0 / 0.80)) * 255.0;
      linear_to_gamma[i] = (uint32_t)(g + 0.5);
    }
    linear_to_gamma[33] = 255;
  }
};
static inline const GammaTables& gamma_tables() {
  static const GammaTables t;
  return t;
}
static inline int linear_to_gamma(uint32_t base_value, int shift) {  // yuv.go:237
  const GammaTables& t = gamma_tables();
  const int v = (int)base_value << shift;
  int tab_pos = v >> (7 + 2);
  if (tab_pos >= 32) tab_pos = 31;
  const int x = v & ((128 << 2) - 1);
  const int v0 = (int)t.linear_to_gamma[tab_pos];
  const int v1 = (int)t.linear_to_gamma[tab_pos + 1];
  const int y = v1 * x + v0 * ((128 << 2) - x);
  return (y + 64) >> 7;
}
static inline uint32_t inv_alpha(uint32_t a) { return a == 0 ? 0u : (1u << 19) / a; }  // yuv.go:343 (kInvAlpha)
static inline int linear_to_gamma_weighted(const uint8_t src[4], const uint8_t alpha[4], uint32_t total_a) {
  const GammaTables& t = gamma_tables();  // yuv.go:466
  const uint32_t sum = alpha[0] * t.gamma_to_linear[src[0]] + alpha[1] * t.gamma_to_linear[src[1]] +
                       alpha[2] * t.gamma_to_linear[src[2]] + alpha[3] * t.gamma_to_linear[src[3]];
  return linear_to_gamma((sum * inv_alpha(total_a)) >> (19 - 2), 0);
}

// ---------------------------------------------------------------- upsample.go:130
// One line pair -> NRGBA.  bot_y / bot_dst / alpha_* may be NULL.
static inline void put_nrgba(int y, uint32_t uv, uint8_t* dst, const uint8_t* alpha, int x) {
  const int u = uv & 0xff, v = (uv >> 16) & 0xff;
  dst[4 * x + 0] = yuv_to_r(y, v);
  dst[4 * x + 1] = yuv_to_g(y, u, v);
  dst[4 * x + 2] = yuv_to_b(y, u);
  dst[4 * x + 3] = alpha ? alpha[x] : 255;
}
static inline void upsample_line_pair_nrgba(const uint8_t* top_y, const uint8_t* bot_y, const uint8_t* top_u,
                                            const uint8_t* top_v, const uint8_t* bot_u, const uint8_t* bot_v,
                                            uint8_t* top_dst, uint8_t* bot_dst, const uint8_t* alpha_top,
                                            const uint8_t* alpha_bot, int width) {
  if (width <= 0) return;
  const int last_pixel_pair = (width - 1) >> 1;
  uint32_t tl_uv = top_u[0] | ((uint32_t)top_v[0] << 16);
  uint32_t l_uv = bot_u[0] | ((uint32_t)bot_v[0] << 16);
  put_nrgba(top_y[0], (3 * tl_uv + l_uv + 0x00020002u) >> 2, top_dst, alpha_top, 0);
  if (bot_y) put_nrgba(bot_y[0], (3 * l_uv + tl_uv + 0x00020002u) >> 2, bot_dst, alpha_bot, 0);
  for (int x = 1; x <= last_pixel_pair; ++x) {
    const uint32_t t_uv = top_u[x] | ((uint32_t)top_v[x] << 16);
    const uint32_t uv = bot_u[x] | ((uint32_t)bot_v[x] << 16);
    const uint32_t avg = tl_uv + t_uv + l_uv + uv + 0x00080008u;
    const uint32_t diag_12 = (avg + 2 * (t_uv + l_uv)) >> 3;
    const uint32_t diag_03 = (avg + 2 * (tl_uv + uv)) >> 3;
    put_nrgba(top_y[2 * x - 1], (diag_12 + tl_uv) >> 1, top_dst, alpha_top, 2 * x - 1);
    put_nrgba(top_y[2 * x], (diag_03 + t_uv) >> 1, top_dst, alpha_top, 2 * x);
    if (bot_y) {
      put_nrgba(bot_y[2 * x - 1], (diag_03 + l_uv) >> 1, bot_dst, alpha_bot, 2 * x - 1);
      put_nrgba(bot_y[2 * x], (diag_12 + uv) >> 1, bot_dst, alpha_bot, 2 * x);
    }
    tl_uv = t_uv;
    l_uv = uv;
  }
  if (!(width & 1)) {
    put_nrgba(top_y[width - 1], (3 * tl_uv + l_uv + 0x00020002u) >> 2, top_dst, alpha_top, width - 1);
    if (bot_y) put_nrgba(bot_y[width - 1], (3 * l_uv + tl_uv + 0x00020002u) >> 2, bot_dst, alpha_bot, width - 1);
  }
}
// buildNRGBA (webp.go:379-450)
static inline void build_nrgba(int width, int height, const uint8_t* yp, int ystride, const uint8_t* up,
                               const uint8_t* vp, int uvstride, const uint8_t* alpha, uint8_t* out) {
  const int ostride = 4 * width;
#define YR(r) (yp + (size_t)(r)*ystride)
#define UR(r) (up + (size_t)(r)*uvstride)
#define VR(r) (vp + (size_t)(r)*uvstride)
#define AR(r) (alpha ? alpha + (size_t)(r)*width : (const uint8_t*)0)
#define DR(r) (out + (size_t)(r)*ostride)
  upsample_line_pair_nrgba(YR(0), 0, UR(0), VR(0), UR(0), VR(0), DR(0), 0, AR(0), 0, width);
  if (height == 1) return;
  int y = 0;
  for (; y + 2 < height; y += 2) {
    const int ct = y / 2, cb = ct + 1;
    upsample_line_pair_nrgba(YR(y + 1), YR(y + 2), UR(ct), VR(ct), UR(cb), VR(cb), DR(y + 1), DR(y + 2),
                             AR(y + 1), AR(y + 2), width);
  }
  if (!(height & 1)) {
    const int lc = (height - 1) / 2;
    upsample_line_pair_nrgba(YR(height - 1), 0, UR(lc), VR(lc), UR(lc), VR(lc), DR(height - 1), 0,
                             AR(height - 1), 0, width);
  }
#undef YR
#undef UR
#undef VR
#undef AR
#undef DR
}

// ---------------------------------------------------------------- ssim.go:12-181
struct DistoStats {
  uint32_t w, xm, ym, xxm, xym, yym;
};
static inline double ssim_calculation(const DistoStats& s, uint32_t N) {  // ssim.go:48
  const uint64_t w2 = (uint64_t)N * N;
  const uint64_t C1 = 20 * w2, C2 = 60 * w2, C3 = 8 * 8 * w2;
  const uint64_t xmxm = (uint64_t)s.xm * s.xm, ymym = (uint64_t)s.ym * s.ym;
  if (xmxm + ymym < C3) return 1.0;
  const int64_t xmym = (int64_t)s.xm * (int64_t)s.ym;
  const int64_t sxy = (int64_t)s.xym * (int64_t)N - xmym;
  const uint64_t sxx = (uint64_t)s.xxm * N - xmxm;
  const uint64_t syy = (uint64_t)s.yym * N - ymym;
  const uint64_t sxy_pos = sxy > 0 ? (uint64_t)sxy : 0;
  const uint64_t num_s = (2 * sxy_pos + C2) >> 8;
  const uint64_t den_s = (sxx + syy + C2) >> 8;
  const uint64_t fnum = (2 * (uint64_t)xmym + C1) * num_s;
  const uint64_t fden = (xmxm + ymym + C1) * den_s;
  if (fden == 0) return 1.0;
  return (double)fnum / (double)fden;
}
static const uint32_t kSsimWeight[7] = {1, 2, 3, 4, 3, 2, 1};
static inline void ssim_acc(DistoStats& s, uint8_t x, uint8_t y, uint32_t w) {
  s.w += w;
  s.xm += w * x;
  s.ym += w * y;
  s.xxm += w * x * x;
  s.xym += w * x * y;
  s.yym += w * y * y;
}
static inline double ssim_get(const uint8_t* a, int sa, const uint8_t* b, int sb) {  // ssim.go:116
  DistoStats s = {0, 0, 0, 0, 0, 0};
  for (int y = 0; y <= 6; ++y)
    for (int x = 0; x <= 6; ++x) ssim_acc(s, a[x + y * sa], b[x + y * sb], kSsimWeight[x] * kSsimWeight[y]);
  return s.w == 0 ? 0.0 : ssim_calculation(s, 256);
}
static inline double ssim_get_clipped(const uint8_t* a, int sa, const uint8_t* b, int sb, int xo, int yo, int W,
                                      int H) {  // ssim.go:132
  DistoStats s = {0, 0, 0, 0, 0, 0};
  const int ymin = yo - 3 < 0 ? 0 : yo - 3, ymax = yo + 3 > H - 1 ? H - 1 : yo + 3;
  const int xmin = xo - 3 < 0 ? 0 : xo - 3, xmax = xo + 3 > W - 1 ? W - 1 : xo + 3;
  for (int y = ymin; y <= ymax; ++y)
    for (int x = xmin; x <= xmax; ++x)
      ssim_acc(s, a[x + y * sa], b[x + y * sb], kSsimWeight[3 + x - xo] * kSsimWeight[3 + y - yo]);
  return ssim_calculation(s, s.w);
}
static inline double psnr_from_sse(uint64_t sse, uint64_t count) {  // ssim.go:163
  if (sse == 0 || count == 0) return 99.0;
  return 10.0 * log10(255.0 * 255.0 / ((double)sse / (double)count));
}

}  // namespace orc
