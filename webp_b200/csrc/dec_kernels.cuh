// Decoder kernels (sm_100a): wavefront reconstruction, wavefront loop filter, fancy upsampler.
//
//   reconstruct : internal/lossy/decode_frame.go:83-217 (reconstructRow), doTransform :22, doUVTransform :46
//   loop filter : internal/lossy/decode_frame.go:283-342 (doFilter) + filter primitives :360-558
//   upsampler   : webp.go:379-450 (buildNRGBA), internal/dsp/upsample.go:130, internal/dsp/yuv.go:71-104
//
// Scheduling: macroblock (x, y) of every image runs in wave t = x + 2y (prediction needs left, top and
// top-right neighbours; the filter of (x, y) rewrites pixels that (x+1, y-1) must have finished with), one
// launch per wave over the whole batch.  Prediction reads UNFILTERED neighbours, so all reconstruction
// waves finish before the first filter wave.
#pragma once
#include "enc_kernels.cuh"

namespace wg {

// Per-macroblock side data produced by the host parser (MBData + FInfo, internal/lossy/decode.go:107-131).
struct MBMeta {
  uint32_t non_zero_y, non_zero_uv;  // 2-bit transform codes, block 0 in bits 31..30 (decode_mb.go:254)
  uint8_t imodes[16];                // I4: sixteen B_* modes; I16: imodes[0] = 16x16 mode
  uint8_t is_i4, uvmode, skip, segment;
  uint8_t f_limit, f_ilevel, f_inner, hev_thresh;
};
static_assert(sizeof(MBMeta) == 32, "MBMeta layout");

struct DecKernelParams {
  const int16_t* coeffs;   // [n][nmb][384] dequantised, WHT already applied
  const MBMeta* meta;      // [n][nmb]
  uint8_t* y; uint8_t* u; uint8_t* v;
  size_t y_plane, uv_plane;
  const uint8_t* filter_type;  // [n] 0 none, 1 simple, 2 complex (decode.go:399)
  int n_images, mb_w, mb_h;
};

// ---- reconstruction: G lanes per macroblock, BPS-strided work buffer in shared memory
template <int G, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) recon_wave_kernel(const DecKernelParams P, int wave) {
  constexpr int MPW = 32 / G;
  __shared__ __align__(16) uint8_t s_buf[WARPS * MPW][YUV_SIZE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, gl = lane % G;
  const int y_lo = max(0, (wave - (P.mb_w - 1) + 1) >> 1), y_hi = min(P.mb_h - 1, wave >> 1);
  const int rows = y_hi - y_lo + 1;
  const long long total = (long long)rows * P.n_images;
  const long long task = ((long long)blockIdx.x * WARPS + warp) * MPW + g;
  const bool active = task < total;
  const int img = active ? (int)(task / rows) : 0;
  const int my = active ? y_lo + (int)(task % rows) : 0;
  const int mx = active ? wave - 2 * my : 0;
  const int nmb = P.mb_w * P.mb_h, mb_idx = my * P.mb_w + mx;
  const int ys = P.mb_w * 16, uvs = P.mb_w * 8;
  uint8_t* py = P.y + (size_t)img * P.y_plane;
  uint8_t* pu = P.u + (size_t)img * P.uv_plane;
  uint8_t* pv = P.v + (size_t)img * P.uv_plane;
  const MBMeta& M = P.meta[(size_t)img * nmb + mb_idx];
  const int16_t* co = P.coeffs + ((size_t)img * nmb + mb_idx) * 384;
  uint8_t* o = s_buf[warp * MPW + g];
  const int x0 = mx * 16, y0 = my * 16;

  if (active) {  // prediction context (decode_frame.go:91-160)
    for (int i = gl; i < 20; i += G) {
      int v = 127;
      if (my > 0) v = py[(size_t)(y0 - 1) * ys + ((i < 16 || mx < P.mb_w - 1) ? x0 + i : x0 + 15)];
      o[Y_OFF - BPS + i] = (uint8_t)v;
    }
    for (int j = gl; j < 16; j += G) o[Y_OFF - 1 + j * BPS] = mx > 0 ? py[(size_t)(y0 + j) * ys + x0 - 1] : 129;
    for (int i = gl; i < 16; i += G) {
      const int pl = i >> 3, c = i & 7;
      const uint8_t* rp = pl ? pv : pu;
      const int off = pl ? V_OFF : U_OFF;
      o[off - BPS + c] = my > 0 ? rp[(size_t)(my * 8 - 1) * uvs + mx * 8 + c] : 127;
      o[off - 1 + c * BPS] = mx > 0 ? rp[(size_t)(my * 8 + c) * uvs + mx * 8 - 1] : 129;
    }
    if (gl == 0) {
      const bool both = mx > 0 && my > 0;
      o[Y_OFF - BPS - 1] = both ? py[(size_t)(y0 - 1) * ys + x0 - 1] : (my > 0 ? 129 : 127);
      o[U_OFF - BPS - 1] = both ? pu[(size_t)(my * 8 - 1) * uvs + mx * 8 - 1] : (my > 0 ? 129 : 127);
      o[V_OFF - BPS - 1] = both ? pv[(size_t)(my * 8 - 1) * uvs + mx * 8 - 1] : (my > 0 ? 129 : 127);
    }
  }
  __syncwarp();
  if (active)
    for (int i = gl; i < 12; i += G) o[Y_OFF - BPS + 16 + (1 + i / 4) * 4 * BPS + (i & 3)] = o[Y_OFF - BPS + 16 + (i & 3)];
  __syncwarp();
  const uint32_t nzy = active ? M.non_zero_y : 0, nzuv = active ? M.non_zero_uv : 0;
  const bool i4 = active && M.is_i4;
  // chroma + I16 luma prediction, cooperative
  if (active) {
    const int pm = check_mode(mx, my, M.uvmode);
    pred_square_coop<G>(gl, pm, o, U_OFF, 8);
    pred_square_coop<G>(gl, pm, o, V_OFF, 8);
    if (!i4) pred_square_coop<G>(gl, check_mode(mx, my, M.imodes[0]), o, Y_OFF, 16);
  }
  __syncwarp();
  if (active) {
    // chroma residuals: 8 blocks; a zero transform code means an all-zero block (doUVTransform)
    for (int b = gl; b < 8; b += G) {
      const int pl = b >> 2, k = b & 3;
      const uint32_t code = ((nzuv >> (8 * pl)) >> (6 - 2 * k)) & 3;
      if (code) {
        const int off = (pl ? V_OFF : U_OFF) + (k >> 1) * 4 * BPS + (k & 1) * 4;
        int p[16], c[16], r[16];
        load4x4(o + off, p);
        const int16_t* cb = co + (16 + b) * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = cb[i];
        itransform(p, c, r);
        store4x4(o + off, r);
      }
    }
    if (!i4) {
      for (int b = gl; b < 16; b += G) {
        const uint32_t code = (nzy >> (30 - 2 * b)) & 3;
        if (code) {
          const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
          int p[16], c[16], r[16];
          load4x4(o + off, p);
#pragma unroll
          for (int i = 0; i < 16; ++i) c[i] = co[b * 16 + i];
          itransform(p, c, r);
          store4x4(o + off, r);
        }
      }
    } else if (gl == 0) {
      // I4: the sixteen sub-blocks form a dependency chain (each predicts from the previous ones)
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        int e[13], p[16], r[16];
        load_pred4_ctx(o + off, e);
        pred4(M.imodes[b], e, p);
        if ((nzy >> (30 - 2 * b)) & 3) {
          int c[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) c[i] = co[b * 16 + i];
          itransform(p, c, r);
          store4x4(o + off, r);
        } else {
          store4x4(o + off, p);
        }
      }
    }
  }
  __syncwarp();
  if (active) {
    for (int i = gl; i < 64; i += G) {
      const int r = i >> 2, c4 = (i & 3) * 4;
      *reinterpret_cast<uint32_t*>(py + (size_t)(y0 + r) * ys + x0 + c4) = *reinterpret_cast<const uint32_t*>(o + Y_OFF + r * BPS + c4);
    }
    for (int i = gl; i < 32; i += G) {
      const int pl = i >> 4, r = (i >> 1) & 7, c4 = (i & 1) * 4;
      uint8_t* rp = pl ? pv : pu;
      *reinterpret_cast<uint32_t*>(rp + (size_t)(my * 8 + r) * uvs + mx * 8 + c4) =
          *reinterpret_cast<const uint32_t*>(o + (pl ? V_OFF : U_OFF) + r * BPS + c4);
    }
  }
}

// ---- loop filter primitives (decode_frame.go:360-558; clip tables of internal/dsp/cliptables.go as arithmetic)
__device__ __forceinline__ int sclip1(int v) { return min(max(v, -128), 127); }
__device__ __forceinline__ int sclip2(int v) { return min(max(v, -16), 15); }
__device__ __forceinline__ void do_filter2(uint8_t* p, int s) {
  const int p1 = p[-2 * s], p0 = p[-s], q0 = p[0], q1 = p[s];
  const int a = 3 * (q0 - p0) + sclip1(p1 - q1);
  const int a1 = sclip2((a + 4) >> 3), a2 = sclip2((a + 3) >> 3);
  p[-s] = (uint8_t)clip8(p0 + a2);
  p[0] = (uint8_t)clip8(q0 - a1);
}
__device__ __forceinline__ void do_filter4(uint8_t* p, int s) {
  const int p1 = p[-2 * s], p0 = p[-s], q0 = p[0], q1 = p[s];
  const int a = 3 * (q0 - p0);
  const int a1 = sclip2((a + 4) >> 3), a2 = sclip2((a + 3) >> 3), a3 = (a1 + 1) >> 1;
  p[-2 * s] = (uint8_t)clip8(p1 + a3);
  p[-s] = (uint8_t)clip8(p0 + a2);
  p[0] = (uint8_t)clip8(q0 - a1);
  p[s] = (uint8_t)clip8(q1 - a3);
}
__device__ __forceinline__ void do_filter6(uint8_t* p, int s) {
  const int p2 = p[-3 * s], p1 = p[-2 * s], p0 = p[-s], q0 = p[0], q1 = p[s], q2 = p[2 * s];
  const int a = sclip1(3 * (q0 - p0) + sclip1(p1 - q1));
  const int a1 = (27 * a + 63) >> 7, a2 = (18 * a + 63) >> 7, a3 = (9 * a + 63) >> 7;
  p[-3 * s] = (uint8_t)clip8(p2 + a3);
  p[-2 * s] = (uint8_t)clip8(p1 + a2);
  p[-s] = (uint8_t)clip8(p0 + a1);
  p[0] = (uint8_t)clip8(q0 - a1);
  p[s] = (uint8_t)clip8(q1 - a2);
  p[2 * s] = (uint8_t)clip8(q2 - a3);
}
__device__ __forceinline__ void simple_edge(uint8_t* p, int s, int thresh) {  // one sample of SimpleV/HFilter16
  const int p1 = p[-2 * s], p0 = p[-s], q0 = p[0], q1 = p[s];
  if (4 * abs(p0 - q0) + abs(p1 - q1) <= 2 * thresh + 1) do_filter2(p, s);
}
// one sample of FilterLoop26 (mb_edge) / FilterLoop24 (inner)
__device__ __forceinline__ void complex_edge(uint8_t* p, int s, int thresh, int it, int hev_t, bool mb_edge) {
  const int p3 = p[-4 * s], p2 = p[-3 * s], p1 = p[-2 * s], p0 = p[-s];
  const int q0 = p[0], q1 = p[s], q2 = p[2 * s], q3 = p[3 * s];
  if (4 * abs(p0 - q0) + abs(p1 - q1) > 2 * thresh + 1) return;
  if (abs(p3 - p2) > it || abs(p2 - p1) > it || abs(p1 - p0) > it || abs(q3 - q2) > it || abs(q2 - q1) > it ||
      abs(q1 - q0) > it)
    return;
  if (abs(p1 - p0) > hev_t || abs(q1 - q0) > hev_t) do_filter2(p, s);
  else if (mb_edge) do_filter6(p, s);
  else do_filter4(p, s);
}

// 16 lanes per macroblock; tiles staged in shared memory with a 4-pixel apron above and to the left.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) filter_wave_kernel(const DecKernelParams P, int wave) {
  constexpr int LS = 20, CS = 12;  // tile strides (bytes)
  __shared__ __align__(16) uint8_t s_y[WARPS * 2][20 * LS];
  __shared__ __align__(16) uint8_t s_c[WARPS * 2][2][12 * CS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 4, gl = lane & 15;
  const int y_lo = max(0, (wave - (P.mb_w - 1) + 1) >> 1), y_hi = min(P.mb_h - 1, wave >> 1);
  const int rows = y_hi - y_lo + 1;
  const long long total = (long long)rows * P.n_images;
  const long long task = ((long long)blockIdx.x * WARPS + warp) * 2 + g;
  bool active = task < total;
  const int img = active ? (int)(task / rows) : 0;
  const int my = active ? y_lo + (int)(task % rows) : 0;
  const int mx = active ? wave - 2 * my : 0;
  const int nmb = P.mb_w * P.mb_h;
  const MBMeta& M = P.meta[(size_t)img * nmb + my * P.mb_w + mx];
  const int ftype = active ? P.filter_type[img] : 0;
  const int limit = active ? M.f_limit : 0;
  active = active && ftype > 0 && limit > 0;
  const int ilevel = M.f_ilevel, inner = M.f_inner, hev_t = M.hev_thresh;
  const int ys = P.mb_w * 16, uvs = P.mb_w * 8;
  uint8_t* py = P.y + (size_t)img * P.y_plane + (size_t)my * 16 * ys + mx * 16;
  uint8_t* pc[2] = {P.u + (size_t)img * P.uv_plane + (size_t)my * 8 * uvs + mx * 8,
                    P.v + (size_t)img * P.uv_plane + (size_t)my * 8 * uvs + mx * 8};
  uint8_t* ty = s_y[warp * 2 + g];
  const int r_lo = my > 0 ? -4 : 0, c_lo = mx > 0 ? -1 : 0;  // rows / 4-byte columns that exist
  const bool cplx = ftype == 2;
  if (active) {
    for (int i = gl; i < 100; i += 16) {
      const int r = i / 5 - 4, c = i % 5 - 1;
      if (r >= r_lo && c >= c_lo)
        *reinterpret_cast<uint32_t*>(ty + (r + 4) * LS + (c + 1) * 4) = *reinterpret_cast<const uint32_t*>(py + (ptrdiff_t)r * ys + c * 4);
    }
    if (cplx)
      for (int i = gl; i < 72; i += 16) {
        const int pl = i / 36, k = i % 36, r = k / 3 - 4, c = k % 3 - 1;
        if (r >= r_lo && c >= c_lo)
          *reinterpret_cast<uint32_t*>(s_c[warp * 2 + g][pl] + (r + 4) * CS + (c + 1) * 4) =
              *reinterpret_cast<const uint32_t*>(pc[pl] + (ptrdiff_t)r * uvs + c * 4);
      }
  }
  __syncwarp();
  uint8_t* Y = ty + 4 * LS + 4;                                     // luma sample (0,0)
  uint8_t* C = s_c[warp * 2 + g][gl >> 3] + 4 * CS + 4;             // chroma sample (0,0) of this lane's plane
  const int k8 = gl & 7;
  // 1. left macroblock edge
  if (active && mx > 0) {
    if (cplx) { complex_edge(Y + gl * LS, 1, limit + 4, ilevel, hev_t, true); complex_edge(C + k8 * CS, 1, limit + 4, ilevel, hev_t, true); }
    else simple_edge(Y + gl * LS, 1, limit + 4);
  }
  __syncwarp();
  // 2. inner vertical edges
  if (active && inner) {
    for (int k = 1; k <= 3; ++k) {
      if (cplx) complex_edge(Y + gl * LS + 4 * k, 1, limit, ilevel, hev_t, false);
      else simple_edge(Y + gl * LS + 4 * k, 1, limit);
    }
    if (cplx) complex_edge(C + k8 * CS + 4, 1, limit, ilevel, hev_t, false);
  }
  __syncwarp();
  // 3. top macroblock edge
  if (active && my > 0) {
    if (cplx) { complex_edge(Y + gl, LS, limit + 4, ilevel, hev_t, true); complex_edge(C + k8, CS, limit + 4, ilevel, hev_t, true); }
    else simple_edge(Y + gl, LS, limit + 4);
  }
  __syncwarp();
  // 4. inner horizontal edges
  if (active && inner) {
    for (int k = 1; k <= 3; ++k) {
      if (cplx) complex_edge(Y + gl + 4 * k * LS, LS, limit, ilevel, hev_t, false);
      else simple_edge(Y + gl + 4 * k * LS, LS, limit);
    }
    if (cplx) complex_edge(C + k8 + 4 * CS, CS, limit, ilevel, hev_t, false);
  }
  __syncwarp();
  if (active) {
    for (int i = gl; i < 100; i += 16) {
      const int r = i / 5 - 4, c = i % 5 - 1;
      if (r >= r_lo && c >= c_lo)
        *reinterpret_cast<uint32_t*>(py + (ptrdiff_t)r * ys + c * 4) = *reinterpret_cast<const uint32_t*>(ty + (r + 4) * LS + (c + 1) * 4);
    }
    if (cplx)
      for (int i = gl; i < 72; i += 16) {
        const int pl = i / 36, k = i % 36, r = k / 3 - 4, c = k % 3 - 1;
        if (r >= r_lo && c >= c_lo)
          *reinterpret_cast<uint32_t*>(pc[pl] + (ptrdiff_t)r * uvs + c * 4) =
              *reinterpret_cast<const uint32_t*>(s_c[warp * 2 + g][pl] + (r + 4) * CS + (c + 1) * 4);
      }
  }
}

// ---- fancy upsampler + YUV->RGBA (buildNRGBA, webp.go:379-450).  One thread = 4 pixels of one output row,
// written with a single 128-bit store.  The 9-3-3-1 diamond of upsample.go:130 per channel (the reference's
// packed-u32 arithmetic never carries between its two 16-bit fields, so per-channel evaluation is identical).
struct UpsampleParams {
  const uint8_t* y; const uint8_t* u; const uint8_t* v; const uint8_t* alpha;  // alpha may be null (A = 255)
  size_t y_plane, uv_plane, alpha_plane, out_image;
  int y_stride, uv_stride, width, height, n;
  uint8_t* out;  // [n][height][width][4]
};
__device__ __forceinline__ int yuv_clip6(int v) { return v < 0 ? 0 : (v > 16383 ? 255 : v >> 6); }
__device__ __forceinline__ uint32_t yuv_to_rgba(int y, int u, int v, int a) {  // yuv.go:71-104
  const int yy = (y * 19077) >> 8;
  const int r = yuv_clip6(yy + ((v * 26149) >> 8) - 14234);
  const int g = yuv_clip6(yy - ((u * 6419) >> 8) - ((v * 13320) >> 8) + 8708);
  const int b = yuv_clip6(yy + ((u * 33050) >> 8) - 17685);
  return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)a << 24);
}
// Fancy upsampling + YUV -> NRGBA (buildNRGBA webp.go:379-450, upsampleLinePairNRGBAGo upsample.go:130, yuv.go:71-104).
// One thread = 16 pixels of a LINE PAIR, as the reference walks the picture: luma rows 2p-1 and 2p share chroma rows p-1 and
// p (row 0 and, for even heights, the last row pair with a mirrored chroma row).  Per thread: one 128-bit luma load per row,
// one 64-bit load + 2 edge bytes per chroma row (U and V of both rows are used by BOTH luma rows), the 9-3-3-1 diamond in
// the reference's packed u | v << 16 arithmetic (avg, diag12 and diag03 once per 2x2 output quad), 4 x 128-bit stores per row.
// One pixel: y24 = luma << 24, uv = u | v << 16 (other bits are don't-care, as in the reference's packed arithmetic).
// MultHi(v, c) = (v * c) >> 8 is the upper word of (v << 24) * c (one IMAD.HI); VP8YUVToR/G/B's clip (yuv.go:71: < 0 -> 0,
// > 16383 -> 255, else >> 6) is an arithmetic shift followed by a saturating conversion, and cvt.pack.sat.u8.s32 converts
// and packs two channels per instruction.
__device__ __forceinline__ uint32_t ups_px(uint32_t y24, uint32_t uv, uint32_t a) {
  const uint32_t u24 = __byte_perm(uv, 0, 0x0444), v24 = __byte_perm(uv, 0, 0x2444);
  const int yy = (int)__umulhi(y24, 19077u);
  const int r = (yy + (int)__umulhi(v24, 26149u) - 14234) >> 6;
  const int g = (yy - (int)__umulhi(u24, 6419u) - (int)__umulhi(v24, 13320u) + 8708) >> 6;
  const int b = (yy + (int)__umulhi(u24, 33050u) - 17685) >> 6;
  uint32_t ba, px;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(ba) : "r"((int)a), "r"(b), "r"(0));   // b | a << 8
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(px) : "r"(g), "r"(r), "r"(ba));        // r | g << 8 | b << 16 | a << 24
  return px;
}
// blockDim = (bx, 256 / bx): thread (tx, ty) takes pixel group blockIdx.x * bx + tx of line pair blockIdx.y * by + ty of image
// blockIdx.z -- no index divisions.
template <bool HAS_ALPHA>
__global__ void __launch_bounds__(256, 4) upsample_nrgba_kernel(const UpsampleParams P) {
  const int gw = (P.width + 15) >> 4, npairs = (P.height >> 1) + 1;
  const int ch = (P.height + 1) >> 1, half_w = (P.width + 1) >> 1;
  {
    const int img = blockIdx.z, p = blockIdx.y * blockDim.y + threadIdx.y, gx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gx >= gw || p >= npairs) return;
    const int x0 = gx * 16;
    const int rt = 2 * p - 1, rb = 2 * p;
    const bool has_top = p >= 1, has_bot = rb < P.height;
    const int ca = max(p - 1, 0), cb = min(p, ch - 1);
    const uint8_t* ua = P.u + (size_t)img * P.uv_plane + (size_t)ca * P.uv_stride;
    const uint8_t* ub = P.u + (size_t)img * P.uv_plane + (size_t)cb * P.uv_stride;
    const uint8_t* va = P.v + (size_t)img * P.uv_plane + (size_t)ca * P.uv_stride;
    const uint8_t* vb = P.v + (size_t)img * P.uv_plane + (size_t)cb * P.uv_stride;
    // chroma samples x0/2 - 1 .. x0/2 + 8 of both rows (clamped at the row ends), packed u | v << 16
    uint32_t A[10], B[10];
    const int c0 = (x0 >> 1) - 1;
    const bool wide = c0 >= 0 && c0 + 9 < half_w && (((uintptr_t)(ua + c0 + 1) | (uintptr_t)(ub + c0 + 1) | (uintptr_t)(va + c0 + 1) | (uintptr_t)(vb + c0 + 1)) & 7) == 0;
    if (wide) {
      const uint2 wua = *reinterpret_cast<const uint2*>(ua + c0 + 1), wva = *reinterpret_cast<const uint2*>(va + c0 + 1);
      const uint2 wub = *reinterpret_cast<const uint2*>(ub + c0 + 1), wvb = *reinterpret_cast<const uint2*>(vb + c0 + 1);
      A[0] = ua[c0] | ((uint32_t)va[c0] << 16); B[0] = ub[c0] | ((uint32_t)vb[c0] << 16);
      A[9] = ua[c0 + 9] | ((uint32_t)va[c0 + 9] << 16); B[9] = ub[c0 + 9] | ((uint32_t)vb[c0 + 9] << 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        A[1 + j] = ((wua.x >> (8 * j)) & 0xffu) | (((wva.x >> (8 * j)) & 0xffu) << 16);
        A[5 + j] = ((wua.y >> (8 * j)) & 0xffu) | (((wva.y >> (8 * j)) & 0xffu) << 16);
        B[1 + j] = ((wub.x >> (8 * j)) & 0xffu) | (((wvb.x >> (8 * j)) & 0xffu) << 16);
        B[5 + j] = ((wub.y >> (8 * j)) & 0xffu) | (((wvb.y >> (8 * j)) & 0xffu) << 16);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        const int c = min(max(c0 + j, 0), half_w - 1);
        A[j] = ua[c] | ((uint32_t)va[c] << 16);
        B[j] = ub[c] | ((uint32_t)vb[c] << 16);
      }
    }
    // luma of both rows
    uint32_t yt[4] = {0, 0, 0, 0}, yb[4] = {0, 0, 0, 0};
    const uint8_t* ytr = P.y + (size_t)img * P.y_plane + (size_t)max(rt, 0) * P.y_stride + x0;
    const uint8_t* ybr = P.y + (size_t)img * P.y_plane + (size_t)min(rb, P.height - 1) * P.y_stride + x0;
    const bool full = x0 + 15 < P.width;
    if (full && (((uintptr_t)ytr | (uintptr_t)ybr) & 15) == 0) {
      const uint4 a = *reinterpret_cast<const uint4*>(ytr), b = *reinterpret_cast<const uint4*>(ybr);
      yt[0] = a.x; yt[1] = a.y; yt[2] = a.z; yt[3] = a.w; yb[0] = b.x; yb[1] = b.y; yb[2] = b.z; yb[3] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int xx = min(i, P.width - 1 - x0);
        yt[i >> 2] |= (uint32_t)ytr[xx] << (8 * (i & 3));
        yb[i >> 2] |= (uint32_t)ybr[xx] << (8 * (i & 3));
      }
    }
    const uint8_t* at = HAS_ALPHA ? P.alpha + (size_t)img * P.alpha_plane + (size_t)max(rt, 0) * P.width + x0 : nullptr;
    const uint8_t* ab = HAS_ALPHA ? P.alpha + (size_t)img * P.alpha_plane + (size_t)min(rb, P.height - 1) * P.width + x0 : nullptr;
    uint8_t* dt = P.out + (size_t)img * P.out_image + ((size_t)max(rt, 0) * P.width + x0) * 4;
    uint8_t* db = P.out + (size_t)img * P.out_image + ((size_t)min(rb, P.height - 1) * P.width + x0) * 4;
    const bool vec = full && (((uintptr_t)dt | (uintptr_t)db) & 15) == 0;
    // pixel i (x = x0 + i) sits in the quad between samples j = (i + 1) >> 1 and j + 1 of A / B: odd x takes the quad's first
    // output, even x its second (upsample.go:74-107); x == 0 and the last pixel of an even width have no horizontal neighbour
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      uint32_t pt[4], pb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = 4 * g4 + k, j = (i + 1) >> 1;
        const uint32_t tl = A[j], tt = A[j + 1], l = B[j], cur = B[j + 1];
        const uint32_t avg = tl + tt + l + cur + 0x00080008u;
        const uint32_t diag12 = (avg + 2 * (tt + l)) >> 3, diag03 = (avg + 2 * (tl + cur)) >> 3;
        uint32_t uvt = (i & 1) ? (diag12 + tl) >> 1 : (diag03 + tt) >> 1;
        uint32_t uvb = (i & 1) ? (diag03 + l) >> 1 : (diag12 + cur) >> 1;
        // x == 0 and the last pixel of an even width have no horizontal neighbour (upsample.go:74-107: (3 * a + b + 2) >> 2);
        // their missing sample is the clamped replica in A / B, and with tl == tt, l == cur (or tt == tl, cur == l) the
        // diamond IS that formula: ((a + b + 2) >> 1 + a) >> 1 == (3 * a + b + 2) >> 2 -- no special case
        const uint32_t ytv = __byte_perm(yt[g4], 0, (k << 12) | 0x444), ybv = __byte_perm(yb[g4], 0, (k << 12) | 0x444);  // luma << 24
        uint32_t alt = 255u, alb = 255u;
        if (HAS_ALPHA) { const int xa = min(i, P.width - 1 - x0); alt = at[xa]; alb = ab[xa]; }
        pt[k] = ups_px(ytv, uvt, alt);
        pb[k] = ups_px(ybv, uvb, alb);
      }
      if (vec) {
        if (has_top) *reinterpret_cast<uint4*>(dt + 16 * g4) = make_uint4(pt[0], pt[1], pt[2], pt[3]);
        if (has_bot) *reinterpret_cast<uint4*>(db + 16 * g4) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (x0 + 4 * g4 + k < P.width) {
            if (has_top) reinterpret_cast<uint32_t*>(dt)[4 * g4 + k] = pt[k];
            if (has_bot) reinterpret_cast<uint32_t*>(db)[4 * g4 + k] = pb[k];
          }
      }
    }
  }
}

}  // namespace wg
