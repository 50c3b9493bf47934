// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  CPU restatement of the reference's SharpYUV RGB -> YUV420 conversion
// as EncoderOptions.UseSharpYUV reaches it: WebP matrix, sRGB transfer, 8-bit input (encode.go:1174-1234).
//   gamma tables, GammaToLinear / LinearToGamma     sharpyuv/gamma.go:47-131, 360, 405
//   convertSharp: import, refinement loop, finish   sharpyuv/sharpyuv.go:190-431
//   WebP matrix                                     sharpyuv/csp.go:61-66
// Pinned against libsharpyuv (the C library the reference's testc/sharpyuv suite compares with) in tests/test_oracle.py.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

namespace sharp {

struct Tables {
  uint32_t g2l[1024 + 2];  // gammaToLinearTab (gamma.go:57-72)
  uint32_t l2g[512 + 2];   // linearToGammaTab (gamma.go:75-89)
};
inline const Tables& tables() {
  static Tables T = [] {
    Tables t;
    const double a = 0.09929682680944, thresh = 0.018053968510807, final_scale = 65536.0, gamma_f = 1.0 / 0.45;
    for (int v = 0; v <= 1024; ++v) {
      const double g = (1.0 / 1024.0) * v;
      const double value = g <= thresh * 4.5 ? g / 4.5 : pow((1.0 / (1.0 + a)) * (g + a), gamma_f);
      t.g2l[v] = (uint32_t)(value * final_scale + 0.5);
    }
    t.g2l[1025] = t.g2l[1024];
    for (int v = 0; v <= 512; ++v) {
      const double g = (1.0 / 512.0) * v;
      const double value = g <= thresh ? 4.5 * g : (1.0 + a) * pow(g, 1.0 / gamma_f) - a;
      t.l2g[v] = (uint32_t)(final_scale * value + 0.5);
    }
    t.l2g[513] = t.l2g[512];
    return t;
  }();
  return T;
}

// Working precision: 8-bit input shifted left by 2 (getPrecisionShift, sharpyuv.go:138), so every sample has 10 bits.
enum { kSfix = 2, kBits = 10, kMax = (1 << kBits) - 1 };

inline uint32_t gamma_to_linear(uint16_t v) { return tables().g2l[v]; }  // toLinearSrgb with shift 0 (gamma.go:117)
inline uint16_t linear_to_gamma(uint32_t value) {  // fromLinearSrgb -> fixedPointInterpolation(v, tab, 7, -6) (gamma.go:100-131)
  const Tables& t = tables();
  const uint32_t pos = value >> 7, x = value - (pos << 7);
  const uint32_t v0 = t.l2g[pos] >> 6, v1 = t.l2g[pos + 1] >> 6;
  return (uint16_t)(v0 + (((v1 - v0) * x + 64) >> 7));
}
inline int rgb_to_gray(int64_t r, int64_t g, int64_t b) { return (int)((13933 * r + 46871 * g + 4732 * b + (1 << 15)) >> 16); }
inline uint16_t clip_depth(int y) { return (uint16_t)(y < 0 ? 0 : y > kMax ? kMax : y); }
inline uint32_t scale_down(uint16_t a, uint16_t b, uint16_t c, uint16_t d) {  // sharpyuv.go:152
  return linear_to_gamma((gamma_to_linear(a) + gamma_to_linear(b) + gamma_to_linear(c) + gamma_to_linear(d) + 2) >> 2);
}
inline uint8_t clip_u8(int32_t v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

static const int32_t kRGBToY[4] = {16839, 33059, 6420, 16 << 16};
static const int32_t kRGBToU[4] = {-9719, -19081, 28800, 128 << 16};
static const int32_t kRGBToV[4] = {28800, -24116, -4684, 128 << 16};

// Rows hold one channel after the other: R[w] G[w] B[w] (sharpyuv.go:190).
inline void update_w(const uint16_t* src, uint16_t* dst, int w) {  // sharpyuv.go:334
  for (int i = 0; i < w; ++i)
    dst[i] = linear_to_gamma((uint32_t)rgb_to_gray(gamma_to_linear(src[i]), gamma_to_linear(src[i + w]), gamma_to_linear(src[i + 2 * w])));
}
inline void update_chroma(const uint16_t* s1, const uint16_t* s2, int16_t* dst, int uv_w) {  // sharpyuv.go:344
  const int w = 2 * uv_w;
  for (int i = 0; i < uv_w; ++i) {
    const int j = 2 * i;
    const int r = (int)scale_down(s1[j], s1[j + 1], s2[j], s2[j + 1]);
    const int g = (int)scale_down(s1[j + w], s1[j + w + 1], s2[j + w], s2[j + w + 1]);
    const int b = (int)scale_down(s1[j + 2 * w], s1[j + 2 * w + 1], s2[j + 2 * w], s2[j + 2 * w + 1]);
    const int gray = rgb_to_gray(r, g, b);
    dst[i] = (int16_t)(r - gray); dst[i + uv_w] = (int16_t)(g - gray); dst[i + 2 * uv_w] = (int16_t)(b - gray);
  }
}
inline uint16_t filter2(int a, int b, int w0) { return clip_depth(((a * 3 + b + 2) >> 2) + w0); }
inline void interpolate_two_rows(const uint16_t* best_y, const int16_t* prev_uv, const int16_t* cur_uv, const int16_t* next_uv, int w,
                                 uint16_t* out1, uint16_t* out2) {  // sharpyuv.go:363
  const int uv_w = w >> 1, len = (w - 1) >> 1;
  for (int k = 0; k < 3; ++k) {
    const int16_t *cu = cur_uv + k * uv_w, *pv = prev_uv + k * uv_w, *nx = next_uv + k * uv_w;
    uint16_t *o1 = out1 + k * w, *o2 = out2 + k * w;
    o1[0] = filter2(cu[0], pv[0], best_y[0]);
    o2[0] = filter2(cu[0], nx[0], best_y[w]);
    for (int i = 0; i < len; ++i) {
      const int a0 = cu[i], a1 = cu[i + 1], b0 = pv[i], b1 = pv[i + 1], c0 = nx[i], c1 = nx[i + 1];
      o1[2 * i + 1] = clip_depth(best_y[2 * i + 1] + ((a0 * 9 + a1 * 3 + b0 * 3 + b1 + 8) >> 4));
      o1[2 * i + 2] = clip_depth(best_y[2 * i + 2] + ((a1 * 9 + a0 * 3 + b1 * 3 + b0 + 8) >> 4));
      o2[2 * i + 1] = clip_depth(best_y[w + 2 * i + 1] + ((a0 * 9 + a1 * 3 + c0 * 3 + c1 + 8) >> 4));
      o2[2 * i + 2] = clip_depth(best_y[w + 2 * i + 2] + ((a1 * 9 + a0 * 3 + c1 * 3 + c0 + 8) >> 4));
    }
    if (!(w & 1)) {
      o1[w - 1] = filter2(cu[uv_w - 1], pv[uv_w - 1], best_y[w - 1]);
      o2[w - 1] = filter2(cu[uv_w - 1], nx[uv_w - 1], best_y[2 * w - 1]);
    }
  }
}

// rgba: 4 bytes per pixel (alpha ignored, as sharpYUVConvert drops it); y/u/v: tight planes of width x height and
// ceil(width/2) x ceil(height/2).  iterations_out (optional) = refinement passes actually run.
inline void convert(const uint8_t* rgba, int stride, int width, int height, uint8_t* y, int y_stride, uint8_t* u, uint8_t* v, int uv_stride,
                    int* iterations_out = nullptr) {
  const int w = (width + 1) & ~1, h = (height + 1) & ~1, uv_w = w >> 1, uv_h = h >> 1;
  std::vector<uint16_t> row1(3 * w), row2(3 * w), best_y((size_t)w * h), target_y((size_t)w * h), best_rgb_y(2 * w);
  std::vector<int16_t> best_uv((size_t)3 * uv_w * uv_h), target_uv((size_t)3 * uv_w * uv_h), best_rgb_uv(3 * uv_w);
  auto import_row = [&](int row, uint16_t* dst) {  // importOneRow (sharpyuv.go:307)
    const uint8_t* p = rgba + (size_t)row * stride;
    for (int i = 0; i < width; ++i)
      for (int k = 0; k < 3; ++k) dst[i + k * w] = (uint16_t)(p[4 * i + k] << kSfix);
    if (width < w)
      for (int k = 0; k < 3; ++k) dst[width + k * w] = dst[width - 1 + k * w];
  };
  for (int j = 0; j < height; j += 2) {  // phase 1 (sharpyuv.go:215-239)
    import_row(j, row1.data());
    if (j != height - 1) import_row(j + 1, row2.data());
    else row2 = row1;
    uint16_t* by = &best_y[(size_t)(j / 2) * 2 * w];
    uint16_t* ty = &target_y[(size_t)(j / 2) * 2 * w];
    for (int i = 0; i < w; ++i) {  // storeGray
      by[i] = (uint16_t)rgb_to_gray(row1[i], row1[i + w], row1[i + 2 * w]);
      by[i + w] = (uint16_t)rgb_to_gray(row2[i], row2[i + w], row2[i + 2 * w]);
    }
    update_w(row1.data(), ty, w);
    update_w(row2.data(), ty + w, w);
    int16_t* tuv = &target_uv[(size_t)(j / 2) * 3 * uv_w];
    update_chroma(row1.data(), row2.data(), tuv, uv_w);
    memcpy(&best_uv[(size_t)(j / 2) * 3 * uv_w], tuv, sizeof(int16_t) * 3 * uv_w);
  }
  const uint64_t threshold = (uint64_t)3 * w * h;  // phase 2 (sharpyuv.go:242-288)
  uint64_t prev_sum = ~(uint64_t)0;
  int iters = 0;
  for (int iter = 0; iter < 4; ++iter) {
    uint64_t sum = 0;
    ++iters;
    for (int j = 0; j < h; j += 2) {
      const size_t cur = (size_t)(j / 2) * 3 * uv_w;
      const size_t prev = j > 0 ? cur - 3 * uv_w : cur, next = j < h - 2 ? cur + 3 * uv_w : cur;
      uint16_t* by = &best_y[(size_t)j * w];
      const uint16_t* ty = &target_y[(size_t)j * w];
      interpolate_two_rows(by, &best_uv[prev], &best_uv[cur], &best_uv[next], w, row1.data(), row2.data());
      update_w(row1.data(), best_rgb_y.data(), w);
      update_w(row2.data(), best_rgb_y.data() + w, w);
      update_chroma(row1.data(), row2.data(), best_rgb_uv.data(), uv_w);
      for (int i = 0; i < 2 * w; ++i) {  // sharpYUVUpdateY (sharpyuv.go:374)
        const int d = (int)ty[i] - (int)best_rgb_y[i];
        by[i] = clip_depth((int)by[i] + d);
        sum += (uint64_t)(d < 0 ? -d : d);
      }
      for (int i = 0; i < 3 * uv_w; ++i)  // sharpYUVUpdateRGB
        best_uv[cur + i] = (int16_t)(best_uv[cur + i] + (int16_t)(target_uv[cur + i] - best_rgb_uv[i]));
    }
    if (iter > 0 && (sum < threshold || sum > prev_sum)) break;
    prev_sum = sum;
  }
  if (iterations_out) *iterations_out = iters;
  const int64_t rounder = (int64_t)1 << (16 + kSfix - 1);  // phase 3 (sharpyuv.go:391)
  for (int j = 0; j < height; ++j)
    for (int i = 0; i < width; ++i) {
      const size_t q = (size_t)(j / 2) * 3 * uv_w + (i >> 1);
      const int64_t wv = best_y[(size_t)j * w + i];
      const int64_t r = best_uv[q] + wv, g = best_uv[q + uv_w] + wv, b = best_uv[q + 2 * uv_w] + wv;
      y[(size_t)j * y_stride + i] = clip_u8((int32_t)((kRGBToY[0] * r + kRGBToY[1] * g + kRGBToY[2] * b + ((int64_t)kRGBToY[3] << kSfix) + rounder) >> (16 + kSfix)));
    }
  for (int j = 0; j < uv_h; ++j)
    for (int i = 0; i < uv_w; ++i) {
      const size_t q = (size_t)j * 3 * uv_w + i;
      const int64_t r = best_uv[q], g = best_uv[q + uv_w], b = best_uv[q + 2 * uv_w];
      u[(size_t)j * uv_stride + i] = clip_u8((int32_t)((kRGBToU[0] * r + kRGBToU[1] * g + kRGBToU[2] * b + ((int64_t)kRGBToU[3] << kSfix) + rounder) >> (16 + kSfix)));
      v[(size_t)j * uv_stride + i] = clip_u8((int32_t)((kRGBToV[0] * r + kRGBToV[1] * g + kRGBToV[2] * b + ((int64_t)kRGBToV[3] << kSfix) + rounder) >> (16 + kSfix)));
    }
}

}  // namespace sharp
