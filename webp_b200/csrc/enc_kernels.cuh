// Encoder kernels (sm_100a): RGBA import, analysis, and the wavefront mode-search kernel.
//
// Mode search restates the reference's per-macroblock pipeline on its parallel path
// (internal/lossy/encode_parallel.go:293-335): import -> prediction context -> pickBestMode
// (I16 RD / I4 RD with SSE pre-screen and trellis / UV RD) -> residuals -> reconstruct -> export.
// Scheduling is B200-first: macroblock (x, y) of every image in the batch runs in wave
// t = x + 2y (left / top / top-right dependency), one launch per wave, G lanes of a
// warp per macroblock, block rows staged in shared memory in the reference's BPS=32 work-buffer
// layout (internal/lossy/constants.go:66-75).
#pragma once
#include "enc_common.cuh"

namespace wg {

struct I4Cand {
  unsigned long long score;
  int disto, rate, nz, mode;
  int16_t lev[16];
  uint8_t rec[16];
};
struct MBShared {
  uint8_t in[384];  // source macroblock, compact: Y 16x16 (stride 16), U 8x8 at 256, V 8x8 at 320 (stride 8)
  uint8_t out[YUV_SIZE];
  uint8_t out2[YUV_SIZE];
  int16_t lev[24][16];
  int dc[16];
  int16_t dcrec[16];
  uint8_t nz[24];
  int sse[10];
  uint8_t smode[12];
  uint8_t pred4s[10][16];  // the ten 4x4 predictions of the current sub-block (pre-screen -> RD candidates)
  I4Cand cand[3];
  int misc[4];
};

template <int G>
__device__ __forceinline__ int grp_sum(int v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
  return v;
}
template <int G>
__device__ __forceinline__ int grp_and(int v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v &= __shfl_xor_sync(0xffffffffu, v, o, G);
  return v;
}

__device__ __forceinline__ void load4x4(const uint8_t* p, int* d) {  // p 4-byte aligned, stride BPS
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p + j * BPS);
    d[4 * j + 0] = w & 0xff;
    d[4 * j + 1] = (w >> 8) & 0xff;
    d[4 * j + 2] = (w >> 16) & 0xff;
    d[4 * j + 3] = w >> 24;
  }
}
template <int STRIDE>
__device__ __forceinline__ void load4x4s(const uint8_t* p, int* d) {  // p 4-byte aligned, compact stride
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p + j * STRIDE);
    d[4 * j + 0] = w & 0xff;
    d[4 * j + 1] = (w >> 8) & 0xff;
    d[4 * j + 2] = (w >> 16) & 0xff;
    d[4 * j + 3] = w >> 24;
  }
}
// source block b of the compact MBShared::in: luma 0..15, chroma 16..23 (U 16..19, V 20..23)
__device__ __forceinline__ void load_src_block(const uint8_t* in, int b, int* d) {
  if (b < 16) load4x4s<16>(in + (b >> 2) * 64 + (b & 3) * 4, d);
  else { const int k = b - 16; load4x4s<8>(in + 256 + (k >> 2) * 64 + ((k >> 1) & 1) * 32 + (k & 1) * 4, d); }
}
__device__ __forceinline__ void store4x4(uint8_t* p, const int* d) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint32_t*>(p + j * BPS) =
        (uint32_t)d[4 * j] | ((uint32_t)d[4 * j + 1] << 8) | ((uint32_t)d[4 * j + 2] << 16) | ((uint32_t)d[4 * j + 3] << 24);
}
__device__ __forceinline__ void load_pred4_ctx(const uint8_t* p, int* e) {  // p = block origin in a BPS buffer
  e[0] = p[-BPS - 1];
  const uint32_t a = *reinterpret_cast<const uint32_t*>(p - BPS), b = *reinterpret_cast<const uint32_t*>(p - BPS + 4);
  e[1] = a & 0xff; e[2] = (a >> 8) & 0xff; e[3] = (a >> 16) & 0xff; e[4] = a >> 24;
  e[5] = b & 0xff; e[6] = (b >> 8) & 0xff; e[7] = (b >> 16) & 0xff; e[8] = b >> 24;
  e[9] = p[-1]; e[10] = p[-1 + BPS]; e[11] = p[-1 + 2 * BPS]; e[12] = p[-1 + 3 * BPS];
}

// Cooperative square predictor into buf at off (predict_lossy.go:27-181); modes 0..6.
template <int G>
__device__ __noinline__ void pred_square_coop(int gl, int mode, uint8_t* buf, int off, int size) {
  uint8_t* d = buf + off;
  const int words_per_row = size >> 2;
  const int nwords = size * words_per_row;
  int dcv = 128;
  if (mode == 0 || mode == 4 || mode == 5) {
    int s = 0;
    if (mode != 4) for (int i = 0; i < size; ++i) s += d[i - BPS];
    if (mode != 5) for (int i = 0; i < size; ++i) s += d[-1 + i * BPS];
    const int shift = (size == 16) ? 4 : 3;
    dcv = (mode == 0) ? (s + size) >> (shift + 1) : (s + (size >> 1)) >> shift;
  }
  const int tl = d[-1 - BPS];
  for (int w = gl; w < nwords; w += G) {
    const int j = w / words_per_row, i0 = (w % words_per_row) * 4;
    uint32_t v;
    if (mode == 1) {
      const int base = d[-1 + j * BPS] - tl;
      v = (uint32_t)clip8(base + d[i0 - BPS]) | ((uint32_t)clip8(base + d[i0 + 1 - BPS]) << 8) |
          ((uint32_t)clip8(base + d[i0 + 2 - BPS]) << 16) | ((uint32_t)clip8(base + d[i0 + 3 - BPS]) << 24);
    } else if (mode == 2) {
      v = *reinterpret_cast<const uint32_t*>(d + i0 - BPS);
    } else if (mode == 3) {
      v = (uint32_t)d[-1 + j * BPS] * 0x01010101u;
    } else {
      v = (uint32_t)dcv * 0x01010101u;
    }
    *reinterpret_cast<uint32_t*>(d + i0 + j * BPS) = v;
  }
}

// collectCoeffStats (internal/lossy/encode_proba.go:10-113) with red.global increments; levels int16 (or int) raster order.
template <class LevT>
__device__ __forceinline__ void stat_block_dev(const LevT* lev, int n_coeffs, int type, int first, int ctx, unsigned int* st, bool) {
  auto add = [&](int band, int c, int p, int bit) { atomicAdd(st + ((((type * 8 + band) * 3 + c) * 11 + p) << 1) + bit, 1u); };
  int n = first;
  if (n_coeffs <= first) { add(c_bands[n], ctx, 0, 0); return; }
  while (n < 16) {
    if (n >= n_coeffs) { add(c_bands[n], ctx, 0, 0); return; }
    add(c_bands[n], ctx, 0, 1);
    for (;;) {
      const int v = abs((int)lev[c_zigzag[n]]);
      const int b = c_bands[n];
      if (v == 0) {
        add(b, ctx, 1, 0);
        if (++n >= 16) return;
        ctx = 0;
        continue;
      }
      add(b, ctx, 1, 1);
      if (v == 1) {
        add(b, ctx, 2, 0);
      } else {
        add(b, ctx, 2, 1);
        if (v <= 4) {
          add(b, ctx, 3, 0);
          if (v == 2) add(b, ctx, 4, 0);
          else { add(b, ctx, 4, 1); add(b, ctx, 5, v == 3 ? 0 : 1); }
        } else if (v <= 10) {
          add(b, ctx, 3, 1); add(b, ctx, 6, 0); add(b, ctx, 7, v <= 6 ? 0 : 1);
        } else {
          add(b, ctx, 3, 1); add(b, ctx, 6, 1);
          const int cat = v <= 18 ? 0 : v <= 34 ? 1 : v <= 66 ? 2 : 3;
          add(b, ctx, 8, cat >> 1);
          add(b, ctx, 9 + (cat >> 1), cat & 1);
        }
      }
      ctx = (v == 1) ? 1 : 2;
      n++;
      break;
    }
  }
}

// Loads of data another macroblock produced (reconstruction borders, NZ context words, neighbour modes): they come through a
// kernel boundary (one launch per wave / macroblock index), so plain loads see them.
// CG: the raster-order chroma chain keeps running inside one launch, so what the previous macroblock wrote is read past L1.
template <bool CG = false, class Tp>
__device__ __forceinline__ Tp ldn(const Tp* p) { return CG ? __ldcg(p) : *p; }

// The mode search of MPW = 32/G macroblocks by one warp: macroblock `task_base + lane/G` of wave `wave`.
// PART (serial RD path only): 0 = the whole macroblock, 1 = luma only, 2 = chroma only (see EncKernelParams::serial_wave).
template <int G, bool FAST, bool SERIAL = false, int PART = 0>
__device__ __forceinline__ void encode_mb_group(const EncKernelParams& P, int wave, long long task_base, MBShared* s_mb_warp,
                                                const CostTabs& T_launch, const uint16_t* s_i4cost) {
  const int lane = threadIdx.x & 31;
  const int g = lane / G, gl = lane % G;
  // rows on this wave: x = wave - 2y in [0, mb_w)
  const int y_lo = max(0, (wave - (P.mb_w - 1) + 1) >> 1), y_hi = min(P.mb_h - 1, wave >> 1);
  const int rows = y_hi - y_lo + 1;
  const bool wave_map = !SERIAL || (PART == 1 && P.serial_wave != 0);  // tasks are (image, row of the wave) pairs, else images
  const long long total = wave_map ? (long long)rows * P.n_images : (long long)P.n_images;
  const long long task = task_base + g;
  // SERIAL (the reference's serial encodeFrame order): `wave` is the raster macroblock index, one macroblock per image per
  // launch.  Rate-control passes (adjustQuantForTarget) re-encode only the images that have not converged: bit 8 of
  // seg[0].flags parks an image, bits 0-7 carry its own getMaxI4RDModes (its quality moves with the search).
  const bool in_range = task < total && (!SERIAL || P.serial_gpw == 0 || g < P.serial_gpw);
  const int img_flags = ((SERIAL || FAST) && in_range) ? P.img[wave_map ? (int)(task / rows) : (int)task].seg[0].flags : 0;
  const bool active0 = task < total && !(img_flags & 0x100) && (!SERIAL || P.serial_gpw == 0 || g < P.serial_gpw);
  const int max_i4_modes = (img_flags & 0xff) ? (img_flags & 0xff) : P.max_i4_modes;
  const int img = active0 ? (wave_map ? (int)(task / rows) : (int)task) : 0;
  const CostTabs& T = T_launch;  // launch-wide tables, or this macroblock's image's own (the plane-split serial kernels)
  const int my = active0 ? (wave_map ? y_lo + (int)(task % rows) : wave / P.mb_w) : 0;
  const int mx = active0 ? (wave_map ? wave - 2 * my : wave - my * P.mb_w) : 0;
  const int nmb = P.mb_w * P.mb_h;
  const int mb_idx = my * P.mb_w + mx;
  const bool active = active0 && (!(SERIAL && PART == 1 && P.serial_wave != 0) || (mb_idx >= P.mb_begin && mb_idx < P.mb_end));
  MBShared& S = s_mb_warp[g];

  const int y_stride = P.mb_w * 16, uv_stride = P.mb_w * 8;
  const uint8_t* src_y = P.src_y + (size_t)img * P.y_plane;
  const uint8_t* src_u = P.src_u + (size_t)img * P.uv_plane;
  const uint8_t* src_v = P.src_v + (size_t)img * P.uv_plane;
  uint8_t* rec_y = P.rec_y + (size_t)img * P.y_plane;
  uint8_t* rec_u = P.rec_u + (size_t)img * P.uv_plane;
  uint8_t* rec_v = P.rec_v + (size_t)img * P.uv_plane;
  uint32_t* ctxw = (PART == 2 ? P.ctx_uv : P.ctx) + (size_t)img * nmb;  // the chroma chain keeps its half of the NZ context apart
  uint8_t* hdr = P.out_hdr + ((size_t)img * nmb + mb_idx) * 48;
  int16_t* oc = P.out_coeffs + ((size_t)img * nmb + mb_idx) * 400;
  const int segment = active ? P.segment[(size_t)img * nmb + mb_idx] : 0;
  const SegParams& seg = P.img[img].seg[segment];
  const bool trellis = P.method >= 4;

  // ---- 1. import source MB with edge replication (encode_iterator.go:145) + 2. prediction context
  if (active) {
    const int x0 = mx * 16, y0 = my * 16;
    const int ww = min(16, P.width - x0), hh = min(16, P.height - y0);
    if (ww == 16 && hh == 16) {  // interior macroblock: 32-bit loads
      for (int i = gl; i < 64; i += G)
        reinterpret_cast<uint32_t*>(S.in)[i] = *reinterpret_cast<const uint32_t*>(src_y + (size_t)(y0 + (i >> 2)) * y_stride + x0 + (i & 3) * 4);
      for (int i = gl; i < 32; i += G) {
        const uint8_t* sp = (i >> 4) ? src_v : src_u;
        reinterpret_cast<uint32_t*>(S.in + 256)[i] = *reinterpret_cast<const uint32_t*>(sp + (size_t)(my * 8 + ((i >> 1) & 7)) * uv_stride + mx * 8 + (i & 1) * 4);
      }
    } else {  // partial macroblock: replicate the last valid sample / row inside the block (importBlock)
      for (int i = gl; i < 256; i += G) {
        const int r = i >> 4, c = i & 15;
        S.in[r * 16 + c] = src_y[(size_t)(y0 + min(r, hh - 1)) * y_stride + x0 + min(c, ww - 1)];
      }
      const int uvw = (ww + 1) >> 1, uvh = (hh + 1) >> 1;
      for (int i = gl; i < 128; i += G) {
        const int pl = i >> 6, r = (i >> 3) & 7, c = i & 7;
        const uint8_t* sp = pl ? src_v : src_u;
        S.in[256 + pl * 64 + r * 8 + c] = sp[(size_t)(my * 8 + min(r, uvh - 1)) * uv_stride + mx * 8 + min(c, uvw - 1)];
      }
    }
    // context from the reconstruction planes (full padded MBs are stored, so partial MBs are exact)
    uint8_t* o = S.out;
    for (int i = gl; i < 20; i += G) {  // top row 16 + top-right 4 (encode_parallel.go:461-487)
      int v = 127;
      if (my > 0) {
        const int xx = (i < 16 || mx < P.mb_w - 1) ? x0 + i : x0 + 15;
        v = ldn<PART == 2>(&rec_y[(size_t)(y0 - 1) * y_stride + xx]);
      }
      o[Y_OFF - BPS + i] = (uint8_t)v;
    }
    for (int j = gl; j < 16; j += G) o[Y_OFF - 1 + j * BPS] = mx > 0 ? ldn<PART == 2>(&rec_y[(size_t)(y0 + j) * y_stride + x0 - 1]) : 129;
    for (int i = gl; i < 16; i += G) {
      const int pl = i >> 3, c = i & 7;
      const uint8_t* rp = pl ? rec_v : rec_u;
      const int off = pl ? V_OFF : U_OFF;
      o[off - BPS + c] = my > 0 ? ldn<PART == 2>(&rp[(size_t)(my * 8 - 1) * uv_stride + mx * 8 + c]) : 127;
      o[off - 1 + c * BPS] = mx > 0 ? ldn<PART == 2>(&rp[(size_t)(my * 8 + c) * uv_stride + mx * 8 - 1]) : 129;
    }
    if (gl == 0) {
      const bool both = mx > 0 && my > 0;
      o[Y_OFF - BPS - 1] = both ? ldn<PART == 2>(&rec_y[(size_t)(y0 - 1) * y_stride + x0 - 1]) : (my > 0 ? 129 : 127);
      o[U_OFF - BPS - 1] = both ? ldn<PART == 2>(&rec_u[(size_t)(my * 8 - 1) * uv_stride + mx * 8 - 1]) : (my > 0 ? 129 : 127);
      o[V_OFF - BPS - 1] = both ? ldn<PART == 2>(&rec_v[(size_t)(my * 8 - 1) * uv_stride + mx * 8 - 1]) : (my > 0 ? 129 : 127);
    }
  }
  __syncwarp();
  if (active) {  // replicate top-right under rows 3, 7, 11 (encode_parallel.go:489-495)
    for (int i = gl; i < 12; i += G) {
      const int r = 1 + i / 4, c = i & 3;
      S.out[Y_OFF - BPS + 16 + r * 4 * BPS + c] = S.out[Y_OFF - BPS + 16 + c];
    }
  }
  // neighbour contexts
  uint32_t top_nz = 0, left_nz = 0;
  int top_nz_dc = 0, left_nz_dc = 0;
  int top_modes[4] = {0, 0, 0, 0}, left_modes[4] = {0, 0, 0, 0};
  if (active) {
    if (my > 0) {
      const uint32_t cw = ldn<PART == 2>(&ctxw[mb_idx - P.mb_w]);
      top_nz = cw & 0xff;
      top_nz_dc = (cw >> 16) & 1;
      const uint8_t* th = hdr - (size_t)P.mb_w * 48;
      if (FAST || SERIAL) {
        const uint32_t tm = ldn<PART == 2>(&P.ctx2[(size_t)img * nmb + mb_idx - P.mb_w]);
        for (int i = 0; i < 4; ++i) top_modes[i] = (tm >> (4 * i)) & 15;
      } else if (ldn<PART == 2>(&th[0]) == 1) { top_modes[0] = ldn<PART == 2>(&th[8 + 12]); top_modes[1] = ldn<PART == 2>(&th[8 + 13]); top_modes[2] = ldn<PART == 2>(&th[8 + 14]); top_modes[3] = ldn<PART == 2>(&th[8 + 15]); }
    }
    if (mx > 0) {
      const uint32_t cw = ldn<PART == 2>(&ctxw[mb_idx - 1]);
      left_nz = (cw >> 8) & 0xff;
      left_nz_dc = (cw >> 17) & 1;
      const uint8_t* lh = hdr - 48;
      if (FAST || SERIAL) {
        const uint32_t lm = ldn<PART == 2>(&P.ctx2[(size_t)img * nmb + mb_idx - 1]);
        for (int i = 0; i < 4; ++i) left_modes[i] = (lm >> (16 + 4 * i)) & 15;
      } else if (ldn<PART == 2>(&lh[0]) == 1) { left_modes[0] = ldn<PART == 2>(&lh[8 + 3]); left_modes[1] = ldn<PART == 2>(&lh[8 + 7]); left_modes[2] = ldn<PART == 2>(&lh[8 + 11]); left_modes[3] = ldn<PART == 2>(&lh[8 + 15]); }
    }
  }
  __syncwarp();

  int best16 = 0;
  unsigned long long score16 = 0;
  unsigned long long score4 = ~0ull;
  uint32_t i4_nzmask = 0;
  uint32_t trial_modes = 0;  // FAST: bottom-row (bits 0-15) and right-column (16-31) trial modes for the neighbours
  if constexpr (PART != 2) {
  if constexpr (!FAST) {  // ---- 3a. I16 RD search (encode_parallel.go:624-735)
    int rate16 = 0, disto16 = 0;
    unsigned long long best_score = ~0ull;
    int src_flat = 0;
    if (active) {  // isFlatSource16 (encode_analysis.go:358)
      const int v0 = S.in[0];
      int ok = 1;
      for (int i = gl; i < 256; i += G) ok &= (S.in[i] == v0);
      src_flat = ok;
    }
    src_flat = grp_and<G>(src_flat);
    if (active) for (int i = gl; i < U_OFF / 4; i += G) reinterpret_cast<uint32_t*>(S.out2)[i] = reinterpret_cast<const uint32_t*>(S.out)[i];
    __syncwarp();
    const int dc_ctx = min(top_nz_dc + left_nz_dc, 2);
    for (int mode = 0; mode < 4; ++mode) {
      const bool allowed = active && !((mode == 2 && my == 0) || (mode == 3 && mx == 0) || (mode == 1 && (mx == 0 || my == 0)));
      if (allowed) pred_square_coop<G>(gl, check_mode(mx, my, mode), S.out2, Y_OFF, 16);
      __syncwarp();
      if (allowed) {
        for (int b = gl; b < 16; b += G) {
          const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
          int s[16], p[16], c[16];
          load_src_block(S.in, b, s);
          load4x4(S.out2 + off, p);
          ftransform(s, p, c);
          S.dc[b] = c[0];
#pragma unroll
          for (int i = 0; i < 16; ++i) S.lev[b][i] = (int16_t)c[i];
          S.nz[b] = quantize_smem(S.lev[b], seg.y1, 1);
        }
      }
      __syncwarp();
      int rate = 0;
      if (allowed) {
        if (gl == 0) {  // WHT of the 16 DCs, quantise, cost, reconstruct DCs
          int d[16], w[16], q[16], dq[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) d[i] = S.dc[i];
          fwht(d, w);
#pragma unroll
          for (int i = 0; i < 16; ++i) S.dc[i] = w[i];
          const int nz_dc = quantize_smem(S.dc, seg.y2, 0);
          rate += kModeFixedCost16(mode) + token_cost_smem(S.dc, nz_dc, 1, dc_ctx, 0, T);
#pragma unroll
          for (int i = 0; i < 16; ++i) q[i] = S.dc[i];
          dequant_block(q, dq, seg.y2);
          iwht(dq, d);
#pragma unroll
          for (int i = 0; i < 16; ++i) S.dcrec[i] = d[i];
        }
        for (int b = gl; b < 16; b += G) {
          const int bx = b & 3, by = b >> 2;
          const int l = bx > 0 ? (S.nz[b - 1] > 0) : ((left_nz >> by) & 1);
          const int t = by > 0 ? (S.nz[b - 4] > 0) : ((top_nz >> bx) & 1);
          rate += token_cost_smem(S.lev[b], S.nz[b], 0, l + t, 1, T);
        }
      }
      __syncwarp();
      int disto = 0, td = 0, any_ac = 0;
      if (allowed) {
        for (int b = gl; b < 16; b += G) {
          const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
          int q[16], dq[16], p[16], r[16], s[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) q[i] = S.lev[b][i];
          dequant_block(q, dq, seg.y1);
          dq[0] = S.dcrec[b];
          load4x4(S.out2 + off, p);
          itransform(p, dq, r);
          load_src_block(S.in, b, s);
          disto += sse16(s, r);
          if (seg.tlambda_sd > 0) td += tdisto4x4(s, r);
          any_ac |= (S.nz[b] > 0);
        }
      }
      rate = grp_sum<G>(rate);
      disto = grp_sum<G>(disto);
      td = grp_sum<G>(td);
      any_ac = grp_sum<G>(any_ac);
      if (allowed) {
        if (seg.tlambda_sd > 0) disto += (seg.tlambda_sd * td + 128) >> 8;
        if (src_flat && any_ac == 0) disto *= 2;
        const unsigned long long score = rd_score(disto, rate, seg.lambda_i16);
        if (score < best_score) { best_score = score; best16 = mode; rate16 = rate; disto16 = disto; }
      }
      __syncwarp();
    }
    score16 = rd_score(disto16, rate16, seg.lambda_mode);
  }
  else {  // ---- 3a (Method < 3). PickBestI16Mode (encode_analysis.go:911-961): prediction SSE + fixed mode cost
    unsigned long long best_score = ~0ull;
    if (active) for (int i = gl; i < U_OFF / 4; i += G) reinterpret_cast<uint32_t*>(S.out2)[i] = reinterpret_cast<const uint32_t*>(S.out)[i];
    __syncwarp();
    for (int mode = 0; mode < 4; ++mode) {
      const bool allowed = active && !((mode == 2 && my == 0) || (mode == 3 && mx == 0) || (mode == 1 && (mx == 0 || my == 0)));
      if (allowed) pred_square_coop<G>(gl, check_mode(mx, my, mode), S.out2, Y_OFF, 16);
      __syncwarp();
      int disto = 0;
      if (allowed)
        for (int b = gl; b < 16; b += G) {
          int s_[16], p_[16];
          load_src_block(S.in, b, s_);
          load4x4(S.out2 + Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4, p_);
          disto += sse16(s_, p_);
        }
      disto = grp_sum<G>(disto);
      if (allowed) {
        const unsigned long long score = rd_score(disto, kModeFixedCost16(mode), seg.lambda_i16);
        if (score < best_score) { best_score = score; best16 = mode; }
      }
      __syncwarp();
    }
    score16 = best_score;
    {  // checkerboard avoidance for flat border blocks (encode_analysis.go:950-959); the score keeps the loop's value
      const bool border = active && (mx == 0 || my == 0);
      int ok = 1;
      if (border) {
        const int v0 = S.in[0];
        for (int i = gl; i < 256; i += G) ok &= (S.in[i] == v0);
      }
      ok = grp_and<G>(ok);
      if (border && ok) best16 = (mx == 0) ? 0 : 2;
    }
  }

  if constexpr (!FAST) {  // ---- 3b. I4 RD search (encode_parallel.go:738-1027)
    if (active) for (int i = gl; i < YUV_SIZE / 4; i += G) reinterpret_cast<uint32_t*>(S.out2)[i] = reinterpret_cast<const uint32_t*>(S.out)[i];
    __syncwarp();
    int total_rate = 0, total_disto = 0, total_hdr = 0;
    bool alive = active;
    uint32_t nzmask = 0;   // bit b: block b has nz > 0
    uint32_t modes_lo = 0, modes_hi = 0;  // 16 x 4-bit modes
    for (int b = 0; b < 16; ++b) {
      const int bx = b & 3, by = b >> 2;
      const int off = Y_OFF + by * 4 * BPS + bx * 4;
      auto get_mode = [&](int k) -> int { return (k < 8) ? (modes_lo >> (4 * k)) & 15 : (modes_hi >> (4 * (k - 8))) & 15; };
      const int top_mode = by == 0 ? top_modes[bx] : get_mode(b - 4);
      const int left_mode = bx == 0 ? left_modes[by] : get_mode(b - 1);
      const bool has_top = my > 0 || by > 0, has_left = mx > 0 || bx > 0;
      const int l = bx > 0 ? ((nzmask >> (b - 1)) & 1) : ((left_nz >> by) & 1);
      const int t = by > 0 ? ((nzmask >> (b - 4)) & 1) : ((top_nz >> bx) & 1);
      const int nz_ctx = l + t;
      int e[13], s[16];
      int n_cand = 0;
      if (alive) {
        load_pred4_ctx(S.out2 + off, e);
        load_src_block(S.in, b, s);
        // pre-screen: prediction SSE of every eligible mode (encode_parallel.go:955-966), written as the compact
        // candidate list the reference builds (position = number of eligible modes before this one)
        uint32_t elig = 0;
#pragma unroll
        for (int m = 0; m < 10; ++m)
          if (!((!has_top && needs_top4(m)) || (!has_left && needs_left4(m)))) elig |= 1u << m;
        n_cand = __popc(elig);
        for (int m = gl; m < 10; m += G) {
          if ((elig >> m) & 1) {
            int p[16];
            pred4(m, e, p);
            const int pos = __popc(elig & ((1u << m) - 1));
            S.sse[pos] = sse16(s, p);
            S.smode[pos] = (uint8_t)m;
            uint32_t* ps = reinterpret_cast<uint32_t*>(S.pred4s[m]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ps[j] = (uint32_t)p[4 * j] | ((uint32_t)p[4 * j + 1] << 8) | ((uint32_t)p[4 * j + 2] << 16) | ((uint32_t)p[4 * j + 3] << 24);
          }
        }
        if (gl == G - 1 && seg.tlambda_sd > 0) S.misc[2] = ttransform(s);  // source half of TDisto, shared by the candidates
      }
      __syncwarp();
      // serial path, Method 3: PickBestI4ModeRD scores every eligible mode in mode order, no pre-screen (encode_analysis.go:1216)
      const bool all_modes = SERIAL && P.method == 3;
      const int K = alive ? (all_modes ? n_cand : min(max_i4_modes, n_cand)) : 0;
      if (alive && gl == 0 && !all_modes) {  // the reference's selection sort of the first K entries, literally (encode_parallel.go:969-983)
#pragma unroll 1
        for (int i = 0; i < K; ++i) {
          int mi = i, mv = S.sse[i];
#pragma unroll 1
          for (int j = i + 1; j < n_cand; ++j) { const int v = S.sse[j]; if (v < mv) { mv = v; mi = j; } }
          if (mi != i) {
            const int ts = S.sse[i]; const uint8_t tm = S.smode[i];
            S.sse[i] = mv; S.smode[i] = S.smode[mi];
            S.sse[mi] = ts; S.smode[mi] = tm;
          }
        }
      }
      __syncwarp();
      // full RD on the K candidates, one lane each, three candidates per round (one round on the row-parallel path)
      I4Cand& BEST = *reinterpret_cast<I4Cand*>(&S.lev[16][0]);  // chroma level slots are idle during the luma search
      unsigned long long bs = ~0ull;
      bool have_best = false;
      for (int r0 = 0; r0 < (SERIAL ? 12 : 3); r0 += 3) {
      if (alive) {
        for (int kk = gl; kk < 3 && r0 + kk < K; kk += G) {
          const int k = kk;
          const int mode = S.smode[r0 + kk];
          int p[16], c[16], q[16], dq[16], r[16];
          load4x4s<4>(S.pred4s[mode], p);
          ftransform(s, p, c);
          I4Cand& C = S.cand[k];
          int nz;
          if (trellis) {  // in place on the candidate's level slot in shared memory
#pragma unroll
            for (int i = 0; i < 16; ++i) C.lev[i] = (int16_t)c[i];
            nz = trellis_block_smem(C.lev, seg.y1, 0, 3, nz_ctx, seg.tlambda_i4, T);
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = C.lev[i];
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) C.lev[i] = (int16_t)c[i];
            nz = quantize_smem(C.lev, seg.y1, 0);
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = C.lev[i];
          }
          dequant_block(q, dq, seg.y1);
          itransform(p, dq, r);
          int disto = sse16(s, r);
          if (seg.tlambda_sd > 0) disto += (seg.tlambda_sd * (abs(ttransform(r) - S.misc[2]) >> 5) + 128) >> 8;
          int rate = 0;
          if (mode > 0) {  // isFlat(levels, 1, 3)  (encode_analysis.go:374)
            int cnt = 0;
#pragma unroll
            for (int i = 1; i < 16; ++i) cnt += (q[i] != 0);
            if (cnt <= 3) rate = 140;
          }
          rate += token_cost_smem(C.lev, nz, 3, nz_ctx, 0, T);
          rate += s_i4cost[(top_mode * 10 + left_mode) * 10 + mode];
          C.score = rd_score(disto, rate, seg.lambda_i4);
          C.disto = disto; C.rate = rate; C.nz = nz; C.mode = mode;
#pragma unroll
          for (int i = 0; i < 16; ++i) { C.lev[i] = (int16_t)q[i]; C.rec[i] = (uint8_t)r[i]; }
        }
      }
      __syncwarp();
      if (alive) {
        for (int k = 0; k < 3 && r0 + k < K; ++k) {
          const I4Cand& Ck = S.cand[k];
          if (256ull * (unsigned long long)Ck.disto >= bs) continue;
          if (Ck.score < bs) {
            bs = Ck.score;
            have_best = true;
            for (int i = gl; i < (int)(sizeof(I4Cand) / 4); i += G) reinterpret_cast<uint32_t*>(&BEST)[i] = reinterpret_cast<const uint32_t*>(&Ck)[i];
          }
        }
      }
      __syncwarp();
      }
      if (alive) {
        (void)have_best;
        const I4Cand& C = BEST;
        const int bm = C.mode;
        if (b < 8) modes_lo |= (uint32_t)bm << (4 * b); else modes_hi |= (uint32_t)bm << (4 * (b - 8));
        total_rate += C.rate;
        total_disto += C.disto;
        total_hdr += s_i4cost[(top_mode * 10 + left_mode) * 10 + bm];
        for (int i = gl; i < 16; i += G) { oc[b * 16 + i] = C.lev[i]; S.lev[b][i] = C.lev[i]; }
        if (gl == 0) { hdr[24 + b] = (uint8_t)C.nz; S.nz[b] = C.nz; }
        if (C.nz > 0) nzmask |= 1u << b;
        if (rd_score(total_disto, total_rate + 211, seg.lambda_mode) >= score16 || total_hdr > 15000) {
          alive = false;
        } else {
          for (int i = gl; i < 16; i += G) S.out2[off + (i >> 2) * BPS + (i & 3)] = C.rec[i];
        }
      }
      __syncwarp();
    }
    if (alive) score4 = rd_score(total_disto, total_rate + 211, seg.lambda_mode);
    i4_nzmask = nzmask;
    if constexpr (SERIAL) {
      // tryI4ModesRD saves the mode-cost context only when the search ran to completion -- whatever wins afterwards; an
      // early exit leaves the iterator's topModes / leftModes as they were (encode_frame.go:337-340)
      if (alive)
        trial_modes = ((modes_hi >> 16) & 0xffffu) | (((modes_lo >> 12) & 15) << 16) | (((modes_lo >> 28) & 15) << 20) |
                      (((modes_hi >> 12) & 15) << 24) | (((modes_hi >> 28) & 15) << 28);
      else
        trial_modes = (uint32_t)top_modes[0] | ((uint32_t)top_modes[1] << 4) | ((uint32_t)top_modes[2] << 8) | ((uint32_t)top_modes[3] << 12) |
                      ((uint32_t)left_modes[0] << 16) | ((uint32_t)left_modes[1] << 20) | ((uint32_t)left_modes[2] << 24) | ((uint32_t)left_modes[3] << 28);
    }
    // ---- decision (encode_parallel.go:572-592)
    const bool use_i4 = active && score4 < score16;
    if (active) {
      if (use_i4) {
        for (int i = gl; i < 64; i += G)
          *reinterpret_cast<uint32_t*>(S.out + Y_OFF + (i >> 2) * BPS + (i & 3) * 4) =
              *reinterpret_cast<const uint32_t*>(S.out2 + Y_OFF + (i >> 2) * BPS + (i & 3) * 4);
        for (int i = gl; i < 16; i += G) hdr[8 + i] = (uint8_t)((i < 8) ? (modes_lo >> (4 * i)) & 15 : (modes_hi >> (4 * (i - 8))) & 15);
      } else {
        pred_square_coop<G>(gl, check_mode(mx, my, best16), S.out, Y_OFF, 16);
        for (int i = gl; i < 16; i += G) hdr[8 + i] = 0;
      }
    }
    __syncwarp();
    S.misc[0] = use_i4;  // uniform within the group
  }
  else {  // ---- 3b (Method < 3). tryI4Modes / PickBestI4Mode (encode_frame.go:193-237, encode_analysis.go:967-1010)
    // Every mode is predicted into a scratch block buffer whose borders are never filled (the reference's yuvP,
    // SURVEY F7): row 0, column 0 and columns 17.. stay zero, the interior holds the LAST evaluated mode of the
    // blocks already visited.  S.out2 plays that buffer (block origin BPS + 1, as in the reference).
    uint32_t modes_lo = 0, modes_hi = 0;
    unsigned long long total = 0;
    const bool do_i4 = P.method >= 2;
    if (active && do_i4) for (int i = gl; i < (17 * BPS) / 4; i += G) reinterpret_cast<uint32_t*>(S.out2)[i] = 0u;
    __syncwarp();
    if (do_i4) {
      for (int b = 0; b < 16; ++b) {
        const int bx = b & 3, by = b >> 2;
        auto get_mode = [&](int k) -> int { return (k < 8) ? (modes_lo >> (4 * k)) & 15 : (modes_hi >> (4 * (k - 8))) & 15; };
        const int top_mode = by == 0 ? top_modes[bx] : get_mode(b - 4);
        const int left_mode = bx == 0 ? left_modes[by] : get_mode(b - 1);
        const bool has_top = my > 0 || by > 0, has_left = mx > 0 || bx > 0;
        uint8_t* pp = S.out2 + BPS + 1 + by * 4 * BPS + bx * 4;
        uint32_t elig = 0;
#pragma unroll
        for (int m = 0; m < 10; ++m)
          if (!((!has_top && needs_top4(m)) || (!has_left && needs_left4(m)))) elig |= 1u << m;
        const int last_mode = 31 - __clz(elig);
        unsigned long long best = ~0ull;  // (score << 4 | mode): lexicographic min == strict '<' in mode order
        int keep[16];
        bool have_keep = false;
        if (active) {
          int e[13], s_[16];
          e[0] = pp[-BPS - 1];
#pragma unroll
          for (int i = 0; i < 8; ++i) e[1 + i] = pp[-BPS + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) e[9 + j] = pp[-1 + j * BPS];
          load_src_block(S.in, b, s_);
          for (int m = gl; m < 10; m += G) {
            if (!((elig >> m) & 1)) continue;
            int p_[16];
            pred4(m, e, p_);
            const unsigned long long sc = rd_score(sse16(s_, p_), s_i4cost[(top_mode * 10 + left_mode) * 10 + m], seg.lambda_i4);
            const unsigned long long key = (sc << 4) | (unsigned)m;
            best = key < best ? key : best;
            if (m == last_mode) {
#pragma unroll
              for (int i = 0; i < 16; ++i) keep[i] = p_[i];
              have_keep = true;
            }
          }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o, G);
          best = other < best ? other : best;
        }
        __syncwarp();  // all lanes have read the context before the block area is overwritten
        if (active && have_keep) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pp[(i >> 2) * BPS + (i & 3)] = (uint8_t)keep[i];
        }
        __syncwarp();
        if (active) {
          const int bm = (int)(best & 15);
          if (b < 8) modes_lo |= (uint32_t)bm << (4 * b); else modes_hi |= (uint32_t)bm << (4 * (b - 8));
          total += best >> 4;
        }
      }
      total += (unsigned long long)seg.lambda_mode * 211ull;
      score4 = total;
      // saved for the neighbours whether or not I4 wins (encode_frame.go:229-231)
      trial_modes = ((modes_hi >> 16) & 0xffffu) | (((modes_lo >> 12) & 15) << 16) | (((modes_lo >> 28) & 15) << 20) |
                    (((modes_hi >> 12) & 15) << 24) | (((modes_hi >> 28) & 15) << 28);
    }
    const bool use_i4 = active && do_i4 && score4 < score16;
    if (active) {
      if (use_i4) {
        for (int i = gl; i < 16; i += G) hdr[8 + i] = (uint8_t)((i < 8) ? (modes_lo >> (4 * i)) & 15 : (modes_hi >> (4 * (i - 8))) & 15);
      } else {
        pred_square_coop<G>(gl, check_mode(mx, my, best16), S.out, Y_OFF, 16);
        for (int i = gl; i < 16; i += G) hdr[8 + i] = 0;
      }
    }
    __syncwarp();
    S.misc[0] = use_i4;
    S.misc[3] = (int)modes_lo; S.misc[2] = (int)modes_hi;  // for the residual pass below
  }
  }
  __syncwarp();
  const bool use_i4 = PART != 2 && S.misc[0] != 0;

  int best_uv = 0;
  if constexpr (PART == 1) {
  } else if constexpr (!FAST) {  // ---- 3c. UV RD search (encode_parallel.go:1030-1116)
    unsigned long long best_score = ~0ull;
    if (active) {
      // the chroma-only pass has no luma search before it that left the whole work buffer (borders included) in out2
      const int from = PART == 2 ? 0 : U_OFF;
      for (int i = gl; i < (YUV_SIZE - from) / 4; i += G)
        reinterpret_cast<uint32_t*>(S.out2 + from)[i] = reinterpret_cast<const uint32_t*>(S.out + from)[i];
    }
    __syncwarp();
    for (int mode = 0; mode < 4; ++mode) {
      const bool allowed = active && !((mode == 2 && my == 0) || (mode == 3 && mx == 0) || (mode == 1 && (mx == 0 || my == 0)));
      if (allowed) {
        const int am = check_mode(mx, my, mode);
        pred_square_coop<G>(gl, am, S.out2, U_OFF, 8);
        pred_square_coop<G>(gl, am, S.out2, V_OFF, 8);
      }
      __syncwarp();
      int disto = 0, ac_cnt = 0;
      if (allowed) {
        for (int b = gl; b < 8; b += G) {
          const int off = ((b & 4) ? V_OFF : U_OFF) + ((b >> 1) & 1) * 4 * BPS + (b & 1) * 4;
          int s[16], p[16], c[16], q[16], dq[16], r[16];
          load_src_block(S.in, 16 + b, s);
          load4x4(S.out2 + off, p);
          ftransform(s, p, c);
#pragma unroll
          for (int i = 0; i < 16; ++i) S.lev[16 + b][i] = (int16_t)c[i];
          S.nz[16 + b] = quantize_smem(S.lev[16 + b], seg.uv, 0);
#pragma unroll
          for (int i = 0; i < 16; ++i) q[i] = S.lev[16 + b][i];
#pragma unroll
          for (int i = 1; i < 16; ++i) ac_cnt += (q[i] != 0);
          dequant_block(q, dq, seg.uv);
          itransform(p, dq, r);
          disto += sse16(s, r);
        }
      }
      __syncwarp();
      int rate = 0;
      if (allowed) {
        for (int b = gl; b < 8; b += G) {
          const int ch = b >> 2, bx = b & 1, by = (b >> 1) & 1;
          const uint32_t tn = (top_nz >> (4 + 2 * ch)) & 3, ln = (left_nz >> (4 + 2 * ch)) & 3;
          const int l = bx > 0 ? (S.nz[16 + b - 1] > 0) : ((ln >> by) & 1);
          const int t = by > 0 ? (S.nz[16 + b - 2] > 0) : ((tn >> bx) & 1);
          rate += token_cost_smem(S.lev[16 + b], S.nz[16 + b], 2, l + t, 0, T);
        }
      }
      rate = grp_sum<G>(rate);
      disto = grp_sum<G>(disto);
      ac_cnt = grp_sum<G>(ac_cnt);
      if (allowed) {
        rate += kModeFixedCostUV(mode);
        if (mode > 0 && ac_cnt <= 2) rate += 140 * 8;
        const unsigned long long score = rd_score(disto, rate, seg.lambda_uv);
        if (score < best_score) { best_score = score; best_uv = mode; }
      }
      __syncwarp();
    }
  }
  else {  // ---- 3c (Method < 3). PickBestUVMode (encode_analysis.go:1015-1070): SSE of both planes + fixed mode cost
    unsigned long long best_score = ~0ull;
    if (active)
      for (int i = gl; i < (YUV_SIZE - U_OFF) / 4; i += G)
        reinterpret_cast<uint32_t*>(S.out2 + U_OFF)[i] = reinterpret_cast<const uint32_t*>(S.out + U_OFF)[i];
    __syncwarp();
    for (int mode = 0; mode < 4; ++mode) {
      const bool allowed = active && !((mode == 2 && my == 0) || (mode == 3 && mx == 0) || (mode == 1 && (mx == 0 || my == 0)));
      if (allowed) {
        const int am = check_mode(mx, my, mode);
        pred_square_coop<G>(gl, am, S.out2, U_OFF, 8);
        pred_square_coop<G>(gl, am, S.out2, V_OFF, 8);
      }
      __syncwarp();
      int disto = 0;
      if (allowed)
        for (int b = gl; b < 8; b += G) {
          int s_[16], p_[16];
          load_src_block(S.in, 16 + b, s_);
          load4x4(S.out2 + ((b & 4) ? V_OFF : U_OFF) + ((b >> 1) & 1) * 4 * BPS + (b & 1) * 4, p_);
          disto += sse16(s_, p_);
        }
      disto = grp_sum<G>(disto);
      if (allowed) {
        const unsigned long long score = rd_score(disto, kModeFixedCostUV(mode), seg.lambda_uv);
        if (score < best_score) { best_score = score; best_uv = mode; }
      }
      __syncwarp();
    }
  }

  // ---- 4. final residuals + 6. reconstruction (encode_parallel.go:1164-1407)
  uint32_t nzy_flags = 0;  // 16 bits: Y block has nz>0 (i16: AC), bit 24: DC block
  int nz_dc = 0;
  if constexpr (PART != 2) {
  if (active && !use_i4) {
    // I16: forward transform all blocks against the cached prediction in S.out
    for (int b = gl; b < 16; b += G) {
      const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
      int s[16], p[16], c[16];
      load_src_block(S.in, b, s);
      load4x4(S.out + off, p);
      ftransform(s, p, c);
      S.dc[b] = c[0];
      c[0] = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) S.lev[b][i] = (int16_t)c[i];  // raw coefficients for now
    }
  }
  __syncwarp();
  if (!trellis) {
    if (active && !use_i4)
      for (int b = gl; b < 16; b += G) {
        S.nz[b] = quantize_smem(S.lev[b], seg.y1, 1);
      }
    __syncwarp();
  } else {
    // trellis with the NZ-context chain: anti-diagonal wavefront over the 16 blocks (7 steps)
    for (int d = 0; d < 7; ++d) {
      if (active && !use_i4) {
        const int by_lo = max(0, d - 3), by_hi = min(3, d);
        for (int k = gl; k <= by_hi - by_lo; k += G) {
          const int by = by_lo + k, bx = d - by, b = by * 4 + bx;
          const int l = bx > 0 ? (S.nz[b - 1] > 0) : ((left_nz >> by) & 1);
          const int t = by > 0 ? (S.nz[b - 4] > 0) : ((top_nz >> bx) & 1);
          S.nz[b] = trellis_block_smem(S.lev[b], seg.y1, 1, 0, l + t, seg.tlambda_i16, T);
        }
      }
      __syncwarp();
    }
  }
  if (active && !use_i4) {
    if (gl == 0) {
      int d[16], w[16], q[16], dq[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) d[i] = S.dc[i];
      fwht(d, w);
#pragma unroll
      for (int i = 0; i < 16; ++i) S.dc[i] = w[i];
      const int n = quantize_smem(S.dc, seg.y2, 0);  // S.dc now holds the WHT levels
      S.misc[1] = n;
#pragma unroll
      for (int i = 0; i < 16; ++i) { q[i] = S.dc[i]; oc[384 + i] = (int16_t)q[i]; }
      dequant_block(q, dq, seg.y2);
      iwht(dq, d);
#pragma unroll
      for (int i = 0; i < 16; ++i) S.dcrec[i] = d[i];
    }
  }
  __syncwarp();
  if (active && !use_i4) {
    nz_dc = S.misc[1];
    for (int b = gl; b < 16; b += G) {
      const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
      int q[16], dq[16], p[16], r[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { q[i] = S.lev[b][i]; oc[b * 16 + i] = (int16_t)q[i]; }
      hdr[24 + b] = (uint8_t)S.nz[b];
      dequant_block(q, dq, seg.y1);
      dq[0] = S.dcrec[b];
      load4x4(S.out + off, p);
      itransform(p, dq, r);
      store4x4(S.out + off, r);
    }
    for (int b = 0; b < 16; ++b) nzy_flags |= (uint32_t)(S.nz[b] > 0) << b;
  }
  if constexpr (FAST) {
    // encodeI4Residuals without cached coefficients (encode_frame.go:439-497): predict from the real context, transform,
    // plain quantisation, immediate reconstruction; the sixteen blocks are a dependency chain, one lane walks them.
    if (active && use_i4 && gl == 0) {
      const uint32_t mlo = (uint32_t)S.misc[3], mhi = (uint32_t)S.misc[2];
      uint32_t nzmask = 0;
      for (int b = 0; b < 16; ++b) {
        const int off = Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
        const int mode = (b < 8) ? (mlo >> (4 * b)) & 15 : (mhi >> (4 * (b - 8))) & 15;
        int e[13], s_[16], p_[16], c[16], q[16], dq[16], r[16];
        load_pred4_ctx(S.out + off, e);
        load_src_block(S.in, b, s_);
        pred4(mode, e, p_);
        ftransform(s_, p_, c);
#pragma unroll
        for (int i = 0; i < 16; ++i) S.lev[b][i] = (int16_t)c[i];
        const int nz = quantize_smem(S.lev[b], seg.y1, 0);
#pragma unroll
        for (int i = 0; i < 16; ++i) { q[i] = S.lev[b][i]; oc[b * 16 + i] = (int16_t)q[i]; }
        S.nz[b] = (uint8_t)nz;
        hdr[24 + b] = (uint8_t)nz;
        if (nz > 0) nzmask |= 1u << b;
        dequant_block(q, dq, seg.y1);
        itransform(p_, dq, r);
        store4x4(S.out + off, r);
      }
      S.misc[1] = (int)nzmask;
    }
    __syncwarp();
    if (active && use_i4) i4_nzmask = (uint32_t)S.misc[1];
  }
  if (active && use_i4) {
    nzy_flags = i4_nzmask;
    for (int i = gl; i < 16; i += G) oc[384 + i] = 0;  // no WHT block on I4 macroblocks
  }
  }
  if constexpr (PART != 1) {
  // UV: predict with the winner, transform, quantise, reconstruct
  if (active) {
    const int am = check_mode(mx, my, best_uv);
    pred_square_coop<G>(gl, am, S.out, U_OFF, 8);
    pred_square_coop<G>(gl, am, S.out, V_OFF, 8);
  }
  __syncwarp();
  if (active) {
    for (int b = gl; b < 8; b += G) {
      const int off = ((b & 4) ? V_OFF : U_OFF) + ((b >> 1) & 1) * 4 * BPS + (b & 1) * 4;
      int s[16], p[16], c[16];
      load_src_block(S.in, 16 + b, s);
      load4x4(S.out + off, p);
      ftransform(s, p, c);
#pragma unroll
      for (int i = 0; i < 16; ++i) S.lev[16 + b][i] = (int16_t)c[i];
    }
  }
  __syncwarp();
  if constexpr (SERIAL) {
    // DC error diffusion (useDerr = Method >= 3 on the serial path): correctDCValues before the quantisation,
    // storeDiffusionErrors after it (encode_frame.go:500-566).  topDerr is per column, leftDerr is NOT reset per row.
    if (active && gl == 0) {
      int8_t* td = P.top_derr + ((size_t)img * P.mb_w + mx) * 4;
      int8_t* ld = P.left_derr + (size_t)img * 4;
      const SegQuant& sq = seg.uv;
      const int zthresh = ((1 << 17) - 1 - sq.dc_bias) / sq.dc_iquant;  // DCZthresh (encode.go:1175)
      auto quantize_single = [&](int16_t* v) -> int {  // quantizeSingle (encode_frame.go:510)
        int V = *v, sign = 1;
        if (V < 0) { sign = -1; V = -V; }
        if (V > zthresh) {
          const int qV = (int)(((uint32_t)V * (uint32_t)sq.dc_iquant + (uint32_t)sq.dc_bias) >> 17) * sq.dc_quant;
          *v = (int16_t)(sign * qV);
          return (sign * (V - qV)) >> 1;
        }
        *v = 0;
        return (sign * V) >> 1;
      };
      for (int ch = 0; ch < 2; ++ch) {
        const int t0 = td[ch * 2], t1 = td[ch * 2 + 1], l0 = ld[ch * 2], l1 = ld[ch * 2 + 1];
        int16_t* c0 = &S.lev[16 + ch * 4 + 0][0];
        int16_t* c1 = &S.lev[16 + ch * 4 + 1][0];
        int16_t* c2 = &S.lev[16 + ch * 4 + 2][0];
        int16_t* c3 = &S.lev[16 + ch * 4 + 3][0];
        *c0 = (int16_t)(*c0 + (int16_t)((7 * t0 + 8 * l0) >> 3));
        const int e0 = quantize_single(c0);
        *c1 = (int16_t)(*c1 + (int16_t)((7 * t1 + 8 * e0) >> 3));
        const int e1 = quantize_single(c1);
        *c2 = (int16_t)(*c2 + (int16_t)((7 * e0 + 8 * l1) >> 3));
        const int e2 = quantize_single(c2);
        *c3 = (int16_t)(*c3 + (int16_t)((7 * e1 + 8 * e2) >> 3));
        const int e3 = quantize_single(c3);
        const int8_t d0 = (int8_t)e1, d1 = (int8_t)e2, d2 = (int8_t)e3;  // info.Derr[ch] (int8 truncation as in the reference)
        const int8_t nl1 = (int8_t)((3 * (int)d2) >> 2);
        ld[ch * 2] = d0; ld[ch * 2 + 1] = nl1;
        td[ch * 2] = d1; td[ch * 2 + 1] = (int8_t)(d2 - nl1);
      }
    }
    __syncwarp();
  }
  if (active) {
    for (int b = gl; b < 8; b += G) {
      const int off = ((b & 4) ? V_OFF : U_OFF) + ((b >> 1) & 1) * 4 * BPS + (b & 1) * 4;
      int p[16], q[16], dq[16], r[16];
      load4x4(S.out + off, p);
      const int nz = quantize_smem(S.lev[16 + b], seg.uv, 0);
      S.nz[16 + b] = nz;
      hdr[40 + b] = (uint8_t)nz;
#pragma unroll
      for (int i = 0; i < 16; ++i) { q[i] = S.lev[16 + b][i]; oc[(16 + b) * 16 + i] = (int16_t)q[i]; }
      dequant_block(q, dq, seg.uv);
      itransform(p, dq, r);
      store4x4(S.out + off, r);
    }
  }
  __syncwarp();

  }
  // ---- 6b. token statistics for the final probability optimisation (collectMBStats, encode_parallel.go:1606-1707;
  // collectCoeffStats, encode_proba.go:10-113).  Contexts are the ones this kernel already holds: a skipped MB
  // contributes nothing and leaves all-zero flags behind, exactly like the reset in recordAllTokens.
  uint32_t nzuv_all = 0;
  if (PART != 1 && active) for (int b = 0; b < 8; ++b) nzuv_all |= (uint32_t)(S.nz[16 + b] > 0) << b;
  const bool mb_skip = active && (nzy_flags == 0) && (use_i4 || nz_dc == 0) && nzuv_all == 0;
  if (active && !mb_skip && P.stats != nullptr) {
    unsigned int* st = P.stats + (size_t)img * STATS_SIZE;
    if (!use_i4 && gl == G - 1) stat_block_dev(S.dc, nz_dc, 1, 0, min(top_nz_dc + left_nz_dc, 2), st, true);
    for (int b = gl; b < 24; b += G) {
      int type, first, l, t;
      if (b < 16) {
        const int bx = b & 3, by = b >> 2;
        type = use_i4 ? 3 : 0; first = use_i4 ? 0 : 1;
        l = bx > 0 ? (S.nz[b - 1] > 0) : ((left_nz >> by) & 1);
        t = by > 0 ? (S.nz[b - 4] > 0) : ((top_nz >> bx) & 1);
      } else {
        const int k = b - 16, ch = k >> 2, bx = k & 1, by = (k >> 1) & 1;
        const uint32_t tn = (top_nz >> (4 + 2 * ch)) & 3, ln = (left_nz >> (4 + 2 * ch)) & 3;
        type = 2; first = 0;
        l = bx > 0 ? (S.nz[b - 1] > 0) : ((ln >> by) & 1);
        t = by > 0 ? (S.nz[b - 2] > 0) : ((tn >> bx) & 1);
      }
      stat_block_dev(S.lev[b], S.nz[b], type, first, l + t, st, false);
    }
  }

  // ---- 7. export: reconstruction planes, header, NZ context (encode_parallel.go:341-428,1410-1496)
  if (active) {
    const int x0 = mx * 16, y0 = my * 16;
    if (PART != 2)
      for (int i = gl; i < 64; i += G) {
        const int r = i >> 2, c4 = (i & 3) * 4;
        *reinterpret_cast<uint32_t*>(rec_y + (size_t)(y0 + r) * y_stride + x0 + c4) =
            *reinterpret_cast<const uint32_t*>(S.out + Y_OFF + r * BPS + c4);
      }
    if (PART != 1)
      for (int i = gl; i < 32; i += G) {
        const int pl = i >> 4, r = (i >> 1) & 7, c4 = (i & 1) * 4;
        uint8_t* rp = pl ? rec_v : rec_u;
        *reinterpret_cast<uint32_t*>(rp + (size_t)(my * 8 + r) * uv_stride + mx * 8 + c4) =
            *reinterpret_cast<const uint32_t*>(S.out + (pl ? V_OFF : U_OFF) + r * BPS + c4);
      }
    if (gl == 0) {
      const uint32_t nzuv = nzuv_all;
      const bool i16 = !use_i4;
      const int dcflag = nz_dc > 0;
      const bool skip = mb_skip;
      if (PART != 2) {
        hdr[0] = use_i4 ? 1 : 0;
        hdr[1] = (uint8_t)(i16 ? best16 : 0);
        hdr[3] = (uint8_t)segment;
        hdr[5] = (uint8_t)(i16 ? nz_dc : 0);
        hdr[6] = 0; hdr[7] = 0;
      }
      if (PART != 1) hdr[2] = (uint8_t)best_uv;
      if (PART == 0) hdr[4] = skip ? 1 : 0;  // split by plane: serial_merge_kernel sets it once both halves are there
      // NZ context words (bottom row / right column flags); a plane-split pass writes its half into its own array
      const uint32_t out_t = ((nzy_flags >> 12) & 0xf) | (((nzuv >> 2) & 3) << 4) | (((nzuv >> 6) & 3) << 6);
      const uint32_t yl = ((nzy_flags >> 3) & 1) | (((nzy_flags >> 7) & 1) << 1) | (((nzy_flags >> 11) & 1) << 2) | (((nzy_flags >> 15) & 1) << 3);
      const uint32_t ul = ((nzuv >> 1) & 1) | (((nzuv >> 3) & 1) << 1);
      const uint32_t vl = ((nzuv >> 5) & 1) | (((nzuv >> 7) & 1) << 1);
      const uint32_t out_l = yl | (ul << 4) | (vl << 6);
      const int tdc = i16 ? dcflag : top_nz_dc, ldc = i16 ? dcflag : left_nz_dc;
      ctxw[mb_idx] = pack_ctx(out_t, out_l, tdc, ldc);
      if ((FAST || SERIAL) && PART != 2) P.ctx2[(size_t)img * nmb + mb_idx] = trial_modes;
    }
  }
}

// Joins the halves a plane-split serial pass left behind for the macroblocks [mb_begin, mb_end) of every image that is not
// parked: NZ context word = luma bits of ctx | chroma bits of ctx_uv, skip flag from the nz counts of both planes.
struct SerialMergeParams {
  uint8_t* hdr; uint32_t* ctx; const uint32_t* ctx_uv; const ImageParams* img;
  int n_images, nmb, mb_begin, mb_end;
};
__global__ void __launch_bounds__(256) serial_merge_kernel(const SerialMergeParams P) {
  const int per = P.mb_end - P.mb_begin;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)per * P.n_images) return;
  const int img = (int)(t / per), mb = P.mb_begin + (int)(t % per);
  if (P.img[img].seg[0].flags & 0x100) return;
  uint8_t* h = P.hdr + ((size_t)img * P.nmb + mb) * 48;
  bool any = false;
  for (int b = 0; b < 24; ++b) any |= h[24 + b] > 0;
  h[4] = (!any && (h[0] == 1 || h[5] == 0)) ? 1 : 0;
  const size_t k = (size_t)img * P.nmb + mb;
  P.ctx[k] = (P.ctx[k] & 0x00030f0fu) | (P.ctx_uv[k] & 0x0000f0f0u);
}

#define WG_STAGE_TABLES(NT)                                                                                                        \
  __shared__ __align__(16) uint16_t s_lc[LC_SIZE];                                                                                 \
  __shared__ __align__(16) uint16_t s_lfc[2048];                                                                                   \
  __shared__ __align__(16) uint16_t s_i4cost[1000];                                                                                \
  __shared__ __align__(16) uint16_t s_eob[EOB_SIZE];                                                                               \
  extern __shared__ __align__(16) unsigned char s_dyn[];                                                                           \
  MBShared* s_mb = reinterpret_cast<MBShared*>(s_dyn);                                                                             \
  for (int i = threadIdx.x; i < LC_SIZE / 8; i += (NT)) reinterpret_cast<uint4*>(s_lc)[i] = reinterpret_cast<const uint4*>(P.lc)[i];       \
  for (int i = threadIdx.x; i < 2048 / 8; i += (NT)) reinterpret_cast<uint4*>(s_lfc)[i] = reinterpret_cast<const uint4*>(P.lfc)[i];        \
  for (int i = threadIdx.x; i < 1000 / 8; i += (NT)) reinterpret_cast<uint4*>(s_i4cost)[i] = reinterpret_cast<const uint4*>(P.i4_costs)[i]; \
  for (int i = threadIdx.x; i < EOB_SIZE / 8; i += (NT)) reinterpret_cast<uint4*>(s_eob)[i] = reinterpret_cast<const uint4*>(P.eob)[i];    \
  __syncthreads();                                                                                                                 \
  CostTabs T;                                                                                                                      \
  T.lc = s_lc; T.eob = s_eob; T.lfc = s_lfc; T.lfc_hi = s_lfc

// Method < 3 (the reference's non-RD decisions, serial-path semantics): same wavefront, lighter body.
template <int G, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) encode_fast_wave_kernel(const EncKernelParams P, int wave) {
  constexpr int MPW = 32 / G;
  WG_STAGE_TABLES(WARPS * 32);
  const int warp = threadIdx.x >> 5;
  encode_mb_group<G, true>(P, wave, ((long long)blockIdx.x * WARPS + warp) * MPW, s_mb + warp * MPW, T, s_i4cost);
}

// collectAllStats (encode_proba.go:171-313) for the serial path's probability refreshes (encode_frame.go:113-117): token
// statistics over the WHOLE per-macroblock array as it stands -- macroblocks of this pass above the refresh point, data of
// the previous pass (or the zero state) below it.  One thread per macroblock; its NZ contexts are rebuilt from the
// neighbours' headers exactly as the reference's raster walk would carry them: a skipped neighbour leaves zero flags, and the
// WHT-DC context comes from the nearest I16 macroblock above / to the left (I4 macroblocks do not touch it).
// hdr [48] = mb_type, i16, uv, segment, skip, nz_dc, 0, 0, modes[16], nz_y[16], nz_uv[8]; stats must be zeroed before.
struct AllStatsParams {
  const uint8_t* hdr;      // [n][nmb][48]
  const int16_t* coeffs;   // [n][nmb][400]
  unsigned int* stats;     // [n][STATS_SIZE]
  int n_images, mb_w, mb_h;
  int cut;                 // macroblocks with raster index >= cut come from the PREVIOUS pass (hdr_prev / coeffs_prev) or, when
                           // those are null, count in their zero state (not yet encoded in the first pass: I16, not skipped, no
                           // coefficients -- encode_frame.go:35-57 over a fresh mbInfo); nmb = none
  const uint8_t* hdr_prev;
  const int16_t* coeffs_prev;
};
__global__ void __launch_bounds__(128) collect_all_stats_kernel(const AllStatsParams P) {
  const int nmb = P.mb_w * P.mb_h;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ uint8_t s_zero[48];  // the zero-state header
  if (threadIdx.x < 48) s_zero[threadIdx.x] = 0;
  __syncthreads();
  if (gid >= (long long)nmb * P.n_images) return;
  const int img = (int)(gid / nmb), idx = (int)(gid % nmb), mx = idx % P.mb_w, my = idx / P.mb_w;
  const uint8_t* H = P.hdr + (size_t)img * nmb * 48;
  const uint8_t* HP = P.hdr_prev ? P.hdr_prev + (size_t)img * nmb * 48 : nullptr;
  auto hdr_of = [&](int i) -> const uint8_t* { return i >= P.cut ? (HP ? HP + (size_t)i * 48 : s_zero) : H + (size_t)i * 48; };
  const uint8_t* h = hdr_of(idx);
  if (h[4]) return;  // skipped: contributes nothing
  const int16_t* c = ((idx >= P.cut && P.coeffs_prev) ? P.coeffs_prev : P.coeffs) + ((size_t)img * nmb + idx) * 400;
  unsigned int* st = P.stats + (size_t)img * STATS_SIZE;
  auto y_flag = [](const uint8_t* n, int b) -> uint32_t { return n[24 + b] > (n[0] == 0 ? 1 : 0); };  // l = nz > first
  uint32_t top = 0, left = 0;
  if (my > 0) {
    const uint8_t* n = hdr_of(idx - P.mb_w);
    if (!n[4]) top = y_flag(n, 12) | (y_flag(n, 13) << 1) | (y_flag(n, 14) << 2) | (y_flag(n, 15) << 3) | ((uint32_t)(n[42] > 0) << 4) |
                     ((uint32_t)(n[43] > 0) << 5) | ((uint32_t)(n[46] > 0) << 6) | ((uint32_t)(n[47] > 0) << 7);
  }
  if (mx > 0) {
    const uint8_t* n = hdr_of(idx - 1);
    if (!n[4]) left = y_flag(n, 3) | (y_flag(n, 7) << 1) | (y_flag(n, 11) << 2) | (y_flag(n, 15) << 3) | ((uint32_t)(n[41] > 0) << 4) |
                      ((uint32_t)(n[43] > 0) << 5) | ((uint32_t)(n[45] > 0) << 6) | ((uint32_t)(n[47] > 0) << 7);
  }
  int first = 0, type = 3;
  if (h[0] == 0) {
    int top_dc = 0, left_dc = 0;
    for (int y = my - 1; y >= 0; --y) {
      const uint8_t* n = hdr_of(y * P.mb_w + mx);
      if (n[0] == 0) { top_dc = n[4] ? 0 : (n[5] > 0); break; }
    }
    for (int x = mx - 1; x >= 0; --x) {
      const uint8_t* n = hdr_of(my * P.mb_w + x);
      if (n[0] == 0) { left_dc = n[4] ? 0 : (n[5] > 0); break; }
    }
    stat_block_dev(c + 384, (int)h[5], 1, 0, top_dc + left_dc, st, true);
    first = 1; type = 0;
  }
  uint32_t tnz = top & 0x0f, lnz = left & 0x0f;
  for (int y = 0; y < 4; ++y) {
    uint32_t l = lnz & 1;
    for (int x = 0; x < 4; ++x) {
      const int b = y * 4 + x, nz = h[24 + b];
      stat_block_dev(c + b * 16, nz, type, first, (int)(l + (tnz & 1)), st, false);
      l = nz > first;
      tnz = (tnz >> 1) | (l << 7);
    }
    tnz >>= 4;
    lnz = (lnz >> 1) | (l << 7);
  }
  for (int ch = 0; ch < 4; ch += 2) {
    tnz = (top >> (4 + ch)) & 0x0f;
    lnz = (left >> (4 + ch)) & 0x0f;
    for (int y = 0; y < 2; ++y) {
      uint32_t l = lnz & 1;
      for (int x = 0; x < 2; ++x) {
        const int k = (ch / 2) * 4 + y * 2 + x, nz = h[40 + k];
        stat_block_dev(c + (16 + k) * 16, nz, 2, 0, (int)(l + (tnz & 1)), st, false);
        l = nz > 0;
        tnz = (tnz >> 1) | (l << 3);
      }
      tnz >>= 2;
      lnz = (lnz >> 1) | (l << 5);
    }
  }
}

// Token statistics of the row-parallel path (collectMBStats, encode_parallel.go:1606-1707; collectCoeffStats,
// encode_proba.go:10-113) after the waves: one (macroblock, block) pair per thread, contexts exactly as the mode search
// held them (neighbour flags from the NZ context words it wrote, in-macroblock flags from the header's nz counts; a skipped
// macroblock contributes nothing).  Counts go to a per-CTA shared-memory histogram first: a CTA covers macroblocks of one
// image, so the 2112 counters of an image see one global atomic per CTA and counter instead of one per token.
struct MBStatsParams {
  const uint8_t* hdr;      // [n][nmb][48]
  const int16_t* coeffs;   // [n][nmb][400]
  const uint32_t* ctxw;    // [n][nmb]
  unsigned int* stats;     // [n][STATS_SIZE], zeroed before
  int mb_w, mb_h, mbs_per_cta;
};
__global__ void __launch_bounds__(256) mb_stats_kernel(const MBStatsParams P) {
  __shared__ unsigned int s_hist[STATS_SIZE];
  for (int i = threadIdx.x; i < STATS_SIZE; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int nmb = P.mb_w * P.mb_h, img = blockIdx.y;
  const int first_mb = blockIdx.x * P.mbs_per_cta, count = min(P.mbs_per_cta, nmb - first_mb);
  const uint8_t* H = P.hdr + (size_t)img * nmb * 48;
  const uint32_t* CW = P.ctxw + (size_t)img * nmb;
  for (int t = threadIdx.x; t < count * 25; t += blockDim.x) {
    const int mb = first_mb + t / 25, b = t % 25;
    const uint8_t* h = H + (size_t)mb * 48;
    if (h[4]) continue;
    const bool use_i4 = h[0] == 1;
    const int my = mb / P.mb_w, mx = mb - my * P.mb_w;
    uint32_t top_nz = 0, left_nz = 0;
    int top_dc = 0, left_dc = 0;
    if (my > 0) { const uint32_t cw = CW[mb - P.mb_w]; top_nz = cw & 0xff; top_dc = (cw >> 16) & 1; }
    if (mx > 0) { const uint32_t cw = CW[mb - 1]; left_nz = (cw >> 8) & 0xff; left_dc = (cw >> 17) & 1; }
    const int16_t* c = P.coeffs + ((size_t)img * nmb + mb) * 400;
    if (b == 24) {
      if (!use_i4) stat_block_dev(c + 384, (int)h[5], 1, 0, min(top_dc + left_dc, 2), s_hist, true);
    } else if (b < 16) {
      const int bx = b & 3, by = b >> 2;
      const int l = bx > 0 ? (h[24 + b - 1] > 0) : ((left_nz >> by) & 1);
      const int tt = by > 0 ? (h[24 + b - 4] > 0) : ((top_nz >> bx) & 1);
      stat_block_dev(c + b * 16, (int)h[24 + b], use_i4 ? 3 : 0, use_i4 ? 0 : 1, l + tt, s_hist, false);
    } else {
      const int k = b - 16, ch = k >> 2, bx = k & 1, by = (k >> 1) & 1;
      const uint32_t tn = (top_nz >> (4 + 2 * ch)) & 3, ln = (left_nz >> (4 + 2 * ch)) & 3;
      const int l = bx > 0 ? (h[24 + b - 1] > 0) : ((ln >> by) & 1);
      const int tt = by > 0 ? (h[24 + b - 2] > 0) : ((tn >> bx) & 1);
      stat_block_dev(c + b * 16, (int)h[24 + b], 2, 0, l + tt, s_hist, false);
    }
  }
  __syncthreads();
  unsigned int* st = P.stats + (size_t)img * STATS_SIZE;
  for (int i = threadIdx.x; i < STATS_SIZE; i += blockDim.x)
    if (s_hist[i]) atomicAdd(st + i, s_hist[i]);
}

// Serial RD path (Method >= 3 where the reference does not go row-parallel): macroblocks in raster order, one launch per
// macroblock index over the whole batch -- the reference's leftDerr carries from the end of one row into the next, so
// rows cannot overlap.  Only offered where no mid-stream probability refresh can occur (<= 96 macroblocks).
template <int G, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) encode_serial_kernel(const EncKernelParams P, int mb_index) {
  constexpr int MPW = 32 / G;
  WG_STAGE_TABLES(WARPS * 32);
  const int warp = threadIdx.x >> 5;
  encode_mb_group<G, false, true>(P, mb_index, ((long long)blockIdx.x * WARPS + warp) * MPW, s_mb + warp * MPW, T, s_i4cost);
}

// Serial RD path with mid-stream probability refreshes (encode_frame.go:35-57): the RD costs follow each image's own
// probability state, so every macroblock group works from ITS image's folded cost tables (rebuilt by the host at each
// refresh, P.lc_img / P.eob_img), staged into shared memory next to the group's work buffers -- the Viterbi reads them a
// dozen times per coefficient position.  One warp per CTA: up to 4 macroblock groups = 4 images = 4 x 13.4 KB of tables;
// macroblocks that share a warp run their data-dependent control flow one after the other, so the host picks P.serial_gpw = 1
// group per warp while the batch leaves SMs to spare.
// The path is split by plane, for frames long enough to refresh their probabilities (config: 3840x2160, rate
// control): inside a refresh segment the luma decisions of a macroblock depend on its left / top / top-right neighbours only,
// so luma runs as x + 2y waves over the segment (tasks = (image, row) pairs of the wave that fall inside [mb_begin, mb_end));
// what forces raster order is the chroma DC error diffusion (leftDerr carries from the end of a row into the next one,
// encode_frame.go:529-566), and chroma depends on nothing in luma -- so chroma runs beside the waves as ONE launch per
// segment in which a group of lanes walks its image's macroblocks in raster order.  A segment thus costs ~mb_w + 2 * rows
// dependent luma steps instead of mb_w * rows whole-macroblock steps.
template <int G>
__global__ void __launch_bounds__(32, 2) encode_serial_luma_wave_kernel(const EncKernelParams P, int wave) {
  const int gpw = P.serial_gpw;
  WG_STAGE_TABLES(32);
  uint16_t* s_tabs = reinterpret_cast<uint16_t*>(s_mb + 32 / G);
  const int y_lo = max(0, (wave - (P.mb_w - 1) + 1) >> 1), y_hi = min(P.mb_h - 1, wave >> 1);
  const int rows = max(y_hi - y_lo + 1, 1);
  const long long task_base = (long long)blockIdx.x * gpw, total = (long long)rows * P.n_images;
  for (int g = 0; g < gpw; ++g) {  // each group's image has its own probability state, hence its own folded cost tables
    if (task_base + g >= total) break;
    const long long img = (task_base + g) / rows;
    uint4* dst = reinterpret_cast<uint4*>(s_tabs + (size_t)g * (LC_SIZE + EOB_SIZE));
    const uint4* a = reinterpret_cast<const uint4*>(P.lc_img + (size_t)img * LC_SIZE);
    const uint4* b = reinterpret_cast<const uint4*>(P.eob_img + (size_t)img * EOB_SIZE);
    for (int i = threadIdx.x; i < LC_SIZE / 8; i += 32) dst[i] = a[i];
    for (int i = threadIdx.x; i < EOB_SIZE / 8; i += 32) dst[LC_SIZE / 8 + i] = b[i];
  }
  __syncwarp();
  CostTabs Tg = T;
  const int g = min((int)(threadIdx.x & 31) / G, gpw - 1);
  Tg.lc = s_tabs + (size_t)g * (LC_SIZE + EOB_SIZE);
  Tg.eob = Tg.lc + LC_SIZE;
  encode_mb_group<G, false, true, 1>(P, wave, task_base, s_mb, Tg, s_i4cost);
}
template <int G>
__global__ void __launch_bounds__(32, 2) encode_serial_chroma_chain_kernel(const EncKernelParams P) {
  const int gpw = P.serial_gpw;
  WG_STAGE_TABLES(32);
  uint16_t* s_tabs = reinterpret_cast<uint16_t*>(s_mb + 32 / G);
  const long long task_base = (long long)blockIdx.x * gpw;
  for (int g = 0; g < gpw; ++g) {
    const long long img = task_base + g;
    if (img >= P.n_images) break;
    uint4* dst = reinterpret_cast<uint4*>(s_tabs + (size_t)g * (LC_SIZE + EOB_SIZE));
    const uint4* a = reinterpret_cast<const uint4*>(P.lc_img + (size_t)img * LC_SIZE);
    const uint4* b = reinterpret_cast<const uint4*>(P.eob_img + (size_t)img * EOB_SIZE);
    for (int i = threadIdx.x; i < LC_SIZE / 8; i += 32) dst[i] = a[i];
    for (int i = threadIdx.x; i < EOB_SIZE / 8; i += 32) dst[LC_SIZE / 8 + i] = b[i];
  }
  __syncwarp();
  CostTabs Tg = T;
  const int g = min((int)(threadIdx.x & 31) / G, gpw - 1);
  Tg.lc = s_tabs + (size_t)g * (LC_SIZE + EOB_SIZE);
  Tg.eob = Tg.lc + LC_SIZE;
#pragma unroll 1
  for (int mb = P.mb_begin; mb < P.mb_end; ++mb) {
    encode_mb_group<G, false, true, 2>(P, mb, task_base, s_mb, Tg, s_i4cost);
    __syncwarp();  // the next macroblock reads (past L1) the chroma borders, context half and diffusion state written here
  }
}

// ------------------------------------------------------------------------------------------------
// RGBA import (internal/lossy/encode.go:671-942 for *image.RGBA / *image.NRGBA sources):
//   Y = RGBToY per pixel (internal/dsp/yuv.go:151); U,V from the gamma-domain 2x2 average
//   (AccumulateRGBA yuv.go:486, LinearToGamma :237, ConvertRGBA32ToUV :553), alpha-weighted when any
//   alpha in the 2x2 block is < 255; edges replicated out to the 16-pixel macroblock grid.
// One thread = 4x2 source pixels: 2 x 128-bit loads, 2 x 32-bit Y stores, 16-bit U and V stores.
struct ImportParams {
  const uint8_t* rgba; size_t image_stride; int stride;  // device RGBA, `stride` bytes per row (multiple of 16)
  int n, width, height, pad_w, pad_h, has_alpha;
  uint8_t* y; uint8_t* u; uint8_t* v; size_t y_plane, uv_plane;
  const uint16_t* gamma_to_linear;  // [256]  (yuv.go:193)
  const uint16_t* linear_to_gamma;  // [34]   (yuv.go:205)
  // dithering (Preprocessing&2): the reference seeds VP8Random identically for every image and draws in a fixed order
  // (all padded luma samples in raster order, then U,V per chroma sample), so the rounding terms depend only on the
  // padded size and the amplitude: one host-built table per batch, shared by all images.  null = fixed rounding.
  const uint16_t* dither_y;   // [pad_h][pad_w]      RandomBits(rg, 16)
  const uint32_t* dither_uv;  // [pad_h/2][pad_w/2][2] RandomBits(rg, 18) for U then V
};
__device__ __forceinline__ int clip_uv(int uv, int rounding) {  // VP8ClipUV (yuv.go:138); rounding = 1<<17 unless dithering
  uv = (uv + rounding + (128 << 18)) >> 18;
  return min(max(uv, 0), 255);
}
__device__ __forceinline__ int lin2gamma(uint32_t v, const uint16_t* l2g) {  // yuv.go:237, shift 0
  const int tab_pos = min((int)(v >> 9), 31);
  const int x = v & 511;
  const int yv = l2g[tab_pos + 1] * x + l2g[tab_pos] * (512 - x);
  return (yv + 64) >> 7;
}
// RGBToY (yuv.go:151) as two byte dot products: 16839 = 65 * 256 + 199, 33059 = 129 * 256 + 35, 6420 = 25 * 256 + 20, so
// Y = (256 * dot(rgb, {65, 129, 25}) + dot(rgb, {199, 35, 20}) + rounding + (16 << 16)) >> 16 -- the same integer, 4 instructions.
__device__ __forceinline__ uint32_t rgb_to_y_px(uint32_t p, uint32_t rounding) {
  const uint32_t lo = __dp4a(p, 0x001423c7u, rounding + (16u << 16));
  return (__dp4a(p, 0x00198141u, 0u) * 256u + lo) >> 16;
}
// blockDim = (64, 4): a thread takes 4 x 2 source pixels at quad column blockIdx.x * 64 + threadIdx.x, row pair
// blockIdx.y * 4 + threadIdx.y of image blockIdx.z -- no index divisions.
template <bool ALPHA, bool DITHER>
__global__ void __launch_bounds__(256) import_rgba_kernel(const ImportParams P) {
  __shared__ uint16_t s_g2l[256];
  __shared__ uint16_t s_l2g[34];
  const int tid = threadIdx.y * 64 + threadIdx.x;
  s_g2l[tid] = P.gamma_to_linear[tid];
  if (tid < 34) s_l2g[tid] = P.linear_to_gamma[tid];
  __syncthreads();
  const int qw = P.pad_w >> 2, qh = P.pad_h >> 1;
  const int cx = blockIdx.x * 64 + threadIdx.x, cy = blockIdx.y * 4 + threadIdx.y, img = blockIdx.z;
  if (cx >= qw || cy >= qh) return;
  const int x0 = cx * 4, y0 = cy * 2;
  const uint8_t* base = P.rgba + (size_t)img * P.image_stride;
  uint32_t px[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint8_t* row = base + (size_t)min(y0 + r, P.height - 1) * P.stride;
    if (x0 + 3 < P.width) {
      const uint4 q = *reinterpret_cast<const uint4*>(row + 4 * x0);
      px[r][0] = q.x; px[r][1] = q.y; px[r][2] = q.z; px[r][3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) px[r][i] = *reinterpret_cast<const uint32_t*>(row + 4 * min(x0 + i, P.width - 1));
    }
  }
  uint8_t* yp = P.y + (size_t)img * P.y_plane;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    uint32_t yv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)  // RGBToYRounding (yuv.go:159) with the drawn term, else the fixed 1 << 15
      yv[i] = rgb_to_y_px(px[r][i], DITHER ? (uint32_t)P.dither_y[(size_t)(y0 + r) * P.pad_w + x0 + i] : (1u << 15));
    *reinterpret_cast<uint32_t*>(yp + (size_t)(y0 + r) * P.pad_w + x0) =
        __byte_perm(__byte_perm(yv[0], yv[1], 0x0040), __byte_perm(yv[2], yv[3], 0x0040), 0x5410);
  }
  uint32_t uu = 0, vv = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t p0 = px[0][2 * k], p1 = px[0][2 * k + 1], p2 = px[1][2 * k], p3 = px[1][2 * k + 1];
    uint32_t a0 = 255, a1 = 255, a2 = 255, a3 = 255;
    if (ALPHA) { a0 = p0 >> 24; a1 = p1 >> 24; a2 = p2 >> 24; a3 = p3 >> 24; }
    const uint32_t ta = a0 + a1 + a2 + a3;
    int c[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int sh = 8 * ch;
      const uint32_t l0 = s_g2l[(p0 >> sh) & 0xff], l1 = s_g2l[(p1 >> sh) & 0xff];
      const uint32_t l2 = s_g2l[(p2 >> sh) & 0xff], l3 = s_g2l[(p3 >> sh) & 0xff];
      if (!ALPHA || ta == 4 * 255 || ta == 0) {
        c[ch] = lin2gamma(l0 + l1 + l2 + l3, s_l2g);
      } else {  // LinearToGammaWeighted (yuv.go:466); kInvAlpha[a] = (1<<19)/a
        const uint32_t sum = a0 * l0 + a1 * l1 + a2 * l2 + a3 * l3;
        c[ch] = lin2gamma((sum * ((1u << 19) / ta)) >> (19 - 2), s_l2g);
      }
      c[ch] &= 0xffff;
    }
    int ru = 1 << 17, rv = 1 << 17;
    if (DITHER) {
      const size_t di = ((size_t)cy * (P.pad_w >> 1) + (x0 >> 1) + k) * 2;
      ru = (int)P.dither_uv[di]; rv = (int)P.dither_uv[di + 1];
    }
    uu |= (uint32_t)clip_uv(-9719 * c[0] - 19081 * c[1] + 28800 * c[2], ru) << (8 * k);
    vv |= (uint32_t)clip_uv(28800 * c[0] - 24116 * c[1] - 4684 * c[2], rv) << (8 * k);
  }
  const size_t uvo = (size_t)img * P.uv_plane + (size_t)cy * (P.pad_w >> 1) + (x0 >> 1);
  *reinterpret_cast<uint16_t*>(P.u + uvo) = (uint16_t)uu;
  *reinterpret_cast<uint16_t*>(P.v + uvo) = (uint16_t)vv;
}

// ------------------------------------------------------------------------------------------------
// Analysis (internal/lossy/encode_analysis.go:245-728): per macroblock, DC and TM 16x16 predictions built
// from SOURCE pixels, 16 forward transforms each, histogram of min(|c|>>3, 31), alpha = 510*last/max;
// chroma DC-only.  mixed = 255 - ((3*luma + uv + 2) >> 2).  16 lanes per macroblock, one 4x4 block per lane.
struct AnalysisParams {
  const uint8_t* y; const uint8_t* u; const uint8_t* v; size_t y_plane, uv_plane;
  int n, mb_w, mb_h;
  int width, height;  // the luma analysis replicates the last real column / row itself (encode_analysis.go:412-424): with dithering
                      // the padded samples of the plane are NOT replicas (each drew its own rounding term)
  uint8_t* alpha;     // [n][nmb] mixed alpha
  uint8_t* uv_alpha;  // [n][nmb] chroma alpha (host sums it for dq_uv_ac)
};
__device__ __forceinline__ int histo_alpha16(int* hist, int gl) {  // alpha from a 32-bin shared histogram, 16 lanes
  const int c0 = hist[gl], c1 = hist[gl + 16];
  int mx = max(c0, c1);
  int last = c1 > 0 ? gl + 16 : (c0 > 0 ? gl : -1);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o, 16));
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o, 16));
  }
  if (last < 0) last = 1;
  int alpha = 0;
  // 2 * 255 * last / mx: numerator < 2^14 and denominator <= 512, so the IEEE single quotient truncates to the exact integer one
  // (an exact quotient is representable; otherwise it lies >= 1 / mx from the integers, the rounding error is < 2^-10 of that)
  if (mx > 1) alpha = (int)__fdiv_rn((float)(2 * 255 * last), (float)mx);
  return min(alpha, 255);
}
__device__ __forceinline__ void histo_add16(const int* c, int* hist) {
#pragma unroll
  for (int k = 0; k < 16; ++k) atomicAdd(&hist[min(abs(c[k]) >> 3, 31)], 1);
}
__global__ void __launch_bounds__(128) analysis_kernel(const AnalysisParams P) {
  __shared__ int s_hist[8][32];
  const int lane = threadIdx.x & 31, gl = lane & 15;
  const int grp = threadIdx.x >> 4;  // 8 macroblocks per CTA
  const int nmb = P.mb_w * P.mb_h;
  // grid = (ceil(mb_w / 8), mb_h, n): no index divisions
  const bool active = (int)blockIdx.x * 8 + grp < P.mb_w;
  const int img = active ? (int)blockIdx.z : 0;
  const int my = active ? (int)blockIdx.y : 0, mx = active ? (int)blockIdx.x * 8 + grp : 0;
  const int mb = my * P.mb_w + mx;
  const int ys = P.mb_w * 16, uvs = P.mb_w * 8;
  const uint8_t* yp = P.y + (size_t)img * P.y_plane + (size_t)my * 16 * ys + mx * 16;
  int* hist = s_hist[grp];
  const int bx = gl & 3, by = gl >> 2;
  // source coordinates clamped to the picture (relative to this macroblock's origin)
  const int xmax = P.width - 1 - mx * 16, ymax = P.height - 1 - my * 16;
  auto cx = [&](int i) { return min(i, xmax); };
  auto cy = [&](int j) { return min(j, ymax); };
  int src[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint8_t* rowp = yp + (size_t)cy(by * 4 + j) * ys;
#pragma unroll
    for (int i = 0; i < 4; ++i) src[4 * j + i] = rowp[cx(bx * 4 + i)];
  }
  // DC value from source neighbours (encode_analysis.go:455-500)
  int s = 0;
  if (my > 0) s += yp[-(ptrdiff_t)ys + cx(gl)];
  if (mx > 0) s += yp[(size_t)cy(gl) * ys - 1];
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
  const int count = (my > 0 ? 16 : 0) + (mx > 0 ? 16 : 0);
  const int dc_val = count == 32 ? (s + 16) >> 5 : (count == 16 ? (s + 8) >> 4 : 128);  // (s + count / 2) / count
  int best_alpha = 256;
  int pred[16], c[16];
  {
    hist[gl] = 0; hist[gl + 16] = 0;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) pred[i] = dc_val;
    ftransform(src, pred, c);
    histo_add16(c, hist);
    __syncwarp();
    best_alpha = min(best_alpha, histo_alpha16(hist, gl));
    __syncwarp();
  }
  if (mx > 0 && my > 0) {  // uniform per 16-lane group; groups of a warp may diverge here, shuffles use width 16
    hist[gl] = 0; hist[gl + 16] = 0;
    const uint8_t* topp = yp - (ptrdiff_t)ys;
    const int tl = topp[-1];
    const int t[4] = {(int)topp[cx(bx * 4)], (int)topp[cx(bx * 4 + 1)], (int)topp[cx(bx * 4 + 2)], (int)topp[cx(bx * 4 + 3)]};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = yp[(size_t)cy(by * 4 + j) * ys - 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) pred[4 * j + i] = clip8(t[i] + l - tl);
    }
    ftransform(src, pred, c);
  }
  __syncwarp();
  if (mx > 0 && my > 0) histo_add16(c, hist);
  __syncwarp();
  {
    const int a = histo_alpha16(hist, gl);  // for edge macroblocks this re-reads the DC histogram: same value
    best_alpha = min(best_alpha, a);
  }
  __syncwarp();
  const int luma_alpha = min(best_alpha, 255);
  // chroma: lanes 0-7 read U neighbours, 8-15 V neighbours; lanes 0-3 transform U blocks, 4-7 V blocks
  const int pl = gl >> 3, k = gl & 7;
  const uint8_t* cp = (pl ? P.v : P.u) + (size_t)img * P.uv_plane + (size_t)my * 8 * uvs + mx * 8;
  int cs = 0;
  if (my > 0) cs += cp[-(ptrdiff_t)uvs + k];
  if (mx > 0) cs += cp[(size_t)k * uvs - 1];
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o, 8);
  const int ccount = (my > 0 ? 8 : 0) + (mx > 0 ? 8 : 0);
  const int cdc = ccount == 16 ? (cs + 8) >> 4 : (ccount == 8 ? (cs + 4) >> 3 : 128);  // (cs + ccount / 2) / ccount
  const int dc_u = __shfl_sync(0xffffffffu, cdc, 0, 16), dc_v = __shfl_sync(0xffffffffu, cdc, 8, 16);
  hist[gl] = 0; hist[gl + 16] = 0;
  __syncwarp();
  if (gl < 8) {
    const int bpl = gl >> 2, bb = gl & 3;
    const uint8_t* bp = (bpl ? P.v : P.u) + (size_t)img * P.uv_plane + (size_t)(my * 8 + (bb >> 1) * 4) * uvs + mx * 8 + (bb & 1) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(bp + (size_t)j * uvs);
      src[4 * j] = w & 0xff; src[4 * j + 1] = (w >> 8) & 0xff; src[4 * j + 2] = (w >> 16) & 0xff; src[4 * j + 3] = w >> 24;
    }
    const int dv = bpl ? dc_v : dc_u;
#pragma unroll
    for (int i = 0; i < 16; ++i) pred[i] = dv;
    ftransform(src, pred, c);
    histo_add16(c, hist);
  }
  __syncwarp();
  const int uv_alpha = histo_alpha16(hist, gl);
  if (active && gl == 0) {
    int mixed = 255 - ((3 * luma_alpha + uv_alpha + 2) >> 2);
    mixed = min(max(mixed, 0), 255);
    P.alpha[(size_t)img * nmb + mb] = (uint8_t)mixed;
    P.uv_alpha[(size_t)img * nmb + mb] = (uint8_t)uv_alpha;
  }
}

}  // namespace wg
