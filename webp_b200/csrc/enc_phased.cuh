// Phase-synchronous mode search of the row-parallel RD path (Method >= 3, encode_parallel.go:168-1496).
//
// One CTA owns M macroblocks of a wave (x + 2y = wave, any image of the batch) and walks them through the reference's
// per-macroblock pipeline in lock step.  Every phase is a flat list of INDEPENDENT tasks -- (macroblock, mode, 4x4 block),
// (macroblock, sub-block, candidate) ... -- dealt densely over all threads of the CTA, so a warp runs 32 tasks of the same
// kind (same predictor, same loop shape) instead of 4 macroblocks x 8 lanes with 3 lanes busy; phases are separated by
// __syncthreads() and exchange data only through shared memory.  Consequences:
//   * lanes are busy in the expensive phases (the <=3 RD candidates of the I4 search with their trellis: 6 per macroblock
//     and step, because the 16 sub-blocks run as an anti-diagonal wavefront bx + 2*by -> 10 dependent steps instead of 16);
//   * the warps of a CTA are in the same few KB of code at the same time (the old kernel's 131 KB body with 12 warps at 12
//     places was instruction-fetch bound);
//   * no warp shuffles and no per-thread state across phases: the very same functions run on the CPU as
//     `for (tid...) phase(tid)` (the hostcheck test harness, threads in shuffled order), which checks the schedule and every
//     barrier against the oracle without a GPU.
// The I4 early exit of tryI4ModesRDParallel (encode_parallel.go:830: running score >= the I16 score, or > 15000 header
// bits) is evaluated on partial sums in wavefront order: rates, distortions and header costs are non-negative, so any
// subset sum that trips the test implies the raster-order prefix test trips too, and an exit only ever means "I16 wins"
// (every I4 side effect is then overwritten, encode_parallel.go:572-592) -- same decisions, same bytes.
#pragma once
#include <string.h>
#include "enc_common.cuh"

namespace wg {

#ifdef __CUDA_ARCH__
#define WG_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define WG_ATOMIC_OR(p, v) atomicOr((p), (v))
#define WG_POPC(x) __popc(x)
#else
#define WG_ATOMIC_ADD(p, v) (*(p) += (v))
#define WG_ATOMIC_OR(p, v) (*(p) |= (v))
#define WG_POPC(x) __builtin_popcount(x)
#endif

WG_HD void ph_cp16(void* d, const void* s) {
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
#else
  memcpy(d, s, 16);
#endif
}
WG_HD void ph_cp8(void* d, const void* s) {
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(s);
#else
  memcpy(d, s, 8);
#endif
}
WG_HD uint32_t ph_ld32(const uint8_t* p) {
#ifdef __CUDA_ARCH__
  return *reinterpret_cast<const uint32_t*>(p);
#else
  uint32_t v; memcpy(&v, p, 4); return v;
#endif
}
WG_HD void ph_st32(uint8_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint32_t*>(p) = v;
#else
  memcpy(p, &v, 4);
#endif
}
template <int STRIDE>
WG_HD void ph_load4x4(const uint8_t* p, int* d) {  // p 4-byte aligned
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t w = ph_ld32(p + j * STRIDE);
    d[4 * j + 0] = w & 0xff; d[4 * j + 1] = (w >> 8) & 0xff; d[4 * j + 2] = (w >> 16) & 0xff; d[4 * j + 3] = w >> 24;
  }
}
template <int STRIDE>
WG_HD void ph_store4x4(uint8_t* p, const int* d) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    ph_st32(p + j * STRIDE, (uint32_t)d[4 * j] | ((uint32_t)d[4 * j + 1] << 8) | ((uint32_t)d[4 * j + 2] << 16) | ((uint32_t)d[4 * j + 3] << 24));
}
WG_HD void ph_store_lev(int16_t* lev, const int* q) {  // lev 16-byte aligned
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = ((uint32_t)q[2 * i] & 0xffffu) | ((uint32_t)q[2 * i + 1] << 16);
  ph_cp16(lev, w); ph_cp16(lev + 8, w + 4);
}
WG_HD void ph_load_lev(const int16_t* lev, int* q) {  // lev 16-byte aligned
  uint32_t w[8];
  ph_cp16(w, lev); ph_cp16(w + 4, lev + 8);
#pragma unroll
  for (int i = 0; i < 8; ++i) { q[2 * i] = (int)(int16_t)(w[i] & 0xffffu); q[2 * i + 1] = (int)(int16_t)(w[i] >> 16); }
}
// source block b of the compact `in` buffer: luma 0..15 (stride 16), chroma 16..23 (U at 256, V at 320, stride 8)
WG_HD void ph_load_src(const uint8_t* in, int b, int* d) {
  if (b < 16) ph_load4x4<16>(in + (b >> 2) * 64 + (b & 3) * 4, d);
  else { const int k = b - 16; ph_load4x4<8>(in + 256 + (k >> 2) * 64 + ((k >> 1) & 1) * 32 + (k & 1) * 4, d); }
}
// block b -> offset of its top-left sample in the BPS work buffer
WG_HD int ph_block_off(int b) {
  if (b < 16) return Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
  const int k = b - 16;
  return ((k & 4) ? V_OFF : U_OFF) + ((k >> 1) & 1) * 4 * BPS + (k & 1) * 4;
}
// 16x16 / 8x8 intra prediction of ONE 4x4 block straight from the borders of the work buffer (predict_lossy.go:27-181):
// mode 0 = DC with the border variant already folded into dcv, 1 = TM, 2 = VE, 3 = HE.  `plane` = Y_OFF / U_OFF / V_OFF,
// (x4, y4) = position of the block inside the plane.
WG_HD void ph_pred_square_block(const uint8_t* buf, int plane, int x4, int y4, int mode, int dcv, int* p) {
  const uint8_t* d = buf + plane;
  if (mode == 1) {
    const int tl = d[-BPS - 1];
    const uint32_t tw = ph_ld32(d - BPS + x4);
    const int t0 = tw & 0xff, t1 = (tw >> 8) & 0xff, t2 = (tw >> 16) & 0xff, t3 = tw >> 24;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int base = d[-1 + (y4 + j) * BPS] - tl;
      p[4 * j + 0] = clip8(base + t0); p[4 * j + 1] = clip8(base + t1); p[4 * j + 2] = clip8(base + t2); p[4 * j + 3] = clip8(base + t3);
    }
  } else if (mode == 2) {
    const uint32_t tw = ph_ld32(d - BPS + x4);
#pragma unroll
    for (int j = 0; j < 4; ++j) { p[4 * j + 0] = tw & 0xff; p[4 * j + 1] = (tw >> 8) & 0xff; p[4 * j + 2] = (tw >> 16) & 0xff; p[4 * j + 3] = tw >> 24; }
  } else if (mode == 3) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int l = d[-1 + (y4 + j) * BPS]; p[4 * j + 0] = l; p[4 * j + 1] = l; p[4 * j + 2] = l; p[4 * j + 3] = l; }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = dcv;
  }
}
WG_HD bool ph_mode_allowed(int mode, int mx, int my) {  // VE needs a top row, HE a left column, TM both (encode_parallel.go:640-651)
  return !((mode == 2 && my == 0) || (mode == 3 && mx == 0) || (mode == 1 && (mx == 0 || my == 0)));
}

// ---- shared memory of one macroblock slot
struct PhAcc { int rate16[4], disto16[4], td16[4], anyac16[4], rate_uv[4], disto_uv[4], ac_uv[4]; };
struct PhSearch {  // I16 + chroma search, two modes per pass
  int16_t lev[2][24][16];
  int dc[2][16];
  int16_t dcrec[2][16];
  uint8_t nz[2][24];
  PhAcc acc;  // per mode, across both passes
};
struct alignas(16) PhI4Slot {  // one 4x4 sub-block in flight
  int16_t lev[3][16];   // 16-byte aligned rows (copied with 128-bit moves)
  uint8_t rec[3][16];
  uint8_t pred[10][16];
  unsigned long long score[3];
  int disto[3], rate[3];
  int sse[10];
  int tsrc;
  uint8_t smode[10];
  uint8_t nz[3], cmode[3];
  uint8_t ncand, K, b, valid, top_mode, left_mode, nz_ctx, pad;
  uint8_t bank_pad[16];  // slot stride = 112 words = 16 (mod 32 banks), see PhMB::bank_pad
};
enum { PH_OUT2 = 576 };
struct PhI4 {
  int16_t lev4[16][16];  // levels of the committed sub-blocks: same place as PhFinal::lev[0..15]
  uint8_t out2[PH_OUT2]; // luma trial buffer (BPS layout, borders included)
  PhI4Slot slot[2];
};
struct PhFinal {
  int16_t lev[25][16];   // MBEncInfo.Coeffs layout (encode.go:250): 16 Y, 4 U, 4 V, WHT
  int dc[16];
  int16_t dcrec[16];
  uint8_t nz[24];
};
struct alignas(16) PhMB {
  union alignas(16) { PhSearch s; PhI4 q; PhFinal f; } u;
  alignas(16) uint8_t in[384];       // source macroblock: Y 16x16 (stride 16), U 8x8 at 256, V 8x8 at 320 (stride 8)
  alignas(16) uint8_t out[YUV_SIZE]; // the reference's yuvOut work buffer (constants.go:66-75): borders, final prediction, reconstruction
  alignas(16) uint8_t hdr[48];
  int active, img, mx, my, segment;
  int top_nz, left_nz, top_nz_dc, left_nz_dc;
  uint8_t top_modes[4], left_modes[4];
  int dcv[3];
  int src_flat, best16, best_uv, use_i4, nz_dc;
  unsigned long long score16;
  int tot_rate, tot_disto, tot_hdr;
  uint32_t nzmask, modes_lo, modes_hi;
  // Lanes of a warp work on the same field of different macroblocks (and slots), so the strides decide the bank conflicts:
  // a macroblock stride of 4 (mod 8) words and a slot stride of 16 spread 16 macroblocks x 2 slots over 8 bank groups
  // (4-way worst case; a stride that is a multiple of 32 words -- the natural size was -- puts all 32 lanes on one bank).
  uint8_t bank_pad[16];
};
static_assert((sizeof(PhMB) / 4) % 8 == 4 && (sizeof(PhI4Slot) / 4) % 32 == 16, "bank spreading strides");

static_assert(sizeof(PhI4Slot) % 16 == 0 && offsetof(PhI4, out2) % 16 == 0 && offsetof(PhI4, slot) % 16 == 0, "128-bit copies");
static_assert(offsetof(PhMB, in) % 16 == 0 && offsetof(PhMB, out) % 16 == 0 && offsetof(PhMB, hdr) % 16 == 0 && sizeof(PhMB) % 16 == 0, "128-bit copies");
static_assert(sizeof(PhFinal) <= sizeof(PhI4) && sizeof(PhSearch) <= sizeof(PhI4), "the I4 scratch is the largest member");

// wave geometry: macroblock slot `task` of wave `wave` -> (img, mx, my)
WG_HD bool ph_task_coords(const EncKernelParams& P, int wave, long long task, int* img, int* mx, int* my) {
  const int y_lo = max(0, (wave - (P.mb_w - 1) + 1) >> 1), y_hi = min(P.mb_h - 1, wave >> 1);
  const int rows = y_hi - y_lo + 1;
  if (rows <= 0 || task >= (long long)rows * P.n_images) return false;
  const int slot = (int)(task / rows);
  *img = P.img_order ? P.img_order[slot] : slot;
  *my = y_lo + (int)(task % rows);
  *mx = wave - 2 * *my;
  return true;
}

// ---- phase: import the source macroblock (importBlock, encode_iterator.go:145), prediction borders and neighbour
// contexts (encode_parallel.go:431-495).  L = NT / M lanes per macroblock.
template <int M, int NT>
WG_HD void ph_load(const EncKernelParams& P, PhMB* mbs, int wave, long long task_base, int tid) {
  constexpr int L = NT / M;
  const int m = tid / L, gl = tid % L;
  if (m >= M) return;
  PhMB& S = mbs[m];
  int img = 0, mx = 0, my = 0;
  const bool active = ph_task_coords(P, wave, task_base + m, &img, &mx, &my);
  const int nmb = P.mb_w * P.mb_h, mb_idx = my * P.mb_w + mx;
  if (gl == 0) { S.active = active; S.img = img; S.mx = mx; S.my = my; }
  if (!active) return;
  const int y_stride = P.mb_w * 16, uv_stride = P.mb_w * 8;
  const uint8_t* src_y = P.src_y + (size_t)img * P.y_plane;
  const uint8_t* src_u = P.src_u + (size_t)img * P.uv_plane;
  const uint8_t* src_v = P.src_v + (size_t)img * P.uv_plane;
  const uint8_t* rec_y = P.rec_y + (size_t)img * P.y_plane;
  const uint8_t* rec_u = P.rec_u + (size_t)img * P.uv_plane;
  const uint8_t* rec_v = P.rec_v + (size_t)img * P.uv_plane;
  const int x0 = mx * 16, y0 = my * 16;
  const int ww = min(16, P.width - x0), hh = min(16, P.height - y0);
  if (ww == 16 && hh == 16) {
    for (int i = gl; i < 32; i += L) {
      if (i < 16) ph_cp16(S.in + i * 16, src_y + (size_t)(y0 + i) * y_stride + x0);
      else { const int k = i - 16; ph_cp8(S.in + 256 + k * 8, ((k >> 3) ? src_v : src_u) + (size_t)(my * 8 + (k & 7)) * uv_stride + mx * 8); }
    }
  } else {  // partial macroblock: replicate the last valid sample / row inside the block
    for (int i = gl; i < 256; i += L) {
      const int r = i >> 4, c = i & 15;
      S.in[i] = src_y[(size_t)(y0 + min(r, hh - 1)) * y_stride + x0 + min(c, ww - 1)];
    }
    const int uvw = (ww + 1) >> 1, uvh = (hh + 1) >> 1;
    for (int i = gl; i < 128; i += L) {
      const int pl = i >> 6, r = min((i >> 3) & 7, uvh - 1), c = min(i & 7, uvw - 1);
      const uint8_t* sp = pl ? src_v : src_u;
      S.in[256 + i] = sp[(size_t)(my * 8 + r) * uv_stride + mx * 8 + c];
    }
  }
  uint8_t* o = S.out;
  for (int i = gl; i < 20; i += L) {  // top row 16 + top-right 4, the latter replicated under rows 3, 7, 11 (encode_parallel.go:461-495)
    int v = 127;
    if (my > 0) v = rec_y[(size_t)(y0 - 1) * y_stride + ((i < 16 || mx < P.mb_w - 1) ? x0 + i : x0 + 15)];
    o[Y_OFF - BPS + i] = (uint8_t)v;
    if (i >= 16) { o[Y_OFF - BPS + i + 4 * BPS] = (uint8_t)v; o[Y_OFF - BPS + i + 8 * BPS] = (uint8_t)v; o[Y_OFF - BPS + i + 12 * BPS] = (uint8_t)v; }
  }
  for (int j = gl; j < 16; j += L) o[Y_OFF - 1 + j * BPS] = mx > 0 ? rec_y[(size_t)(y0 + j) * y_stride + x0 - 1] : 129;
  for (int i = gl; i < 16; i += L) {
    const int pl = i >> 3, c = i & 7;
    const uint8_t* rp = pl ? rec_v : rec_u;
    const int off = pl ? V_OFF : U_OFF;
    o[off - BPS + c] = my > 0 ? rp[(size_t)(my * 8 - 1) * uv_stride + mx * 8 + c] : 127;
    o[off - 1 + c * BPS] = mx > 0 ? rp[(size_t)(my * 8 + c) * uv_stride + mx * 8 - 1] : 129;
  }
  if (gl == (1 % L)) {
    const bool both = mx > 0 && my > 0;
    o[Y_OFF - BPS - 1] = both ? rec_y[(size_t)(y0 - 1) * y_stride + x0 - 1] : (my > 0 ? 129 : 127);
    o[U_OFF - BPS - 1] = both ? rec_u[(size_t)(my * 8 - 1) * uv_stride + mx * 8 - 1] : (my > 0 ? 129 : 127);
    o[V_OFF - BPS - 1] = both ? rec_v[(size_t)(my * 8 - 1) * uv_stride + mx * 8 - 1] : (my > 0 ? 129 : 127);
  }
  if (gl == (2 % L)) {  // neighbour NZ contexts and 4x4 modes (updateNZContextParallel, encode_parallel.go:341-428)
    const uint32_t* ctxw = P.ctx + (size_t)img * nmb;
    const uint8_t* hdr = P.out_hdr + ((size_t)img * nmb + mb_idx) * 48;
    int top_nz = 0, left_nz = 0, top_dc = 0, left_dc = 0;
    uint8_t tm[4] = {0, 0, 0, 0}, lm[4] = {0, 0, 0, 0};
    if (my > 0) {
      const uint32_t cw = ctxw[mb_idx - P.mb_w];
      top_nz = cw & 0xff; top_dc = (cw >> 16) & 1;
      const uint8_t* th = hdr - (size_t)P.mb_w * 48;
      if (th[0] == 1) { tm[0] = th[8 + 12]; tm[1] = th[8 + 13]; tm[2] = th[8 + 14]; tm[3] = th[8 + 15]; }
    }
    if (mx > 0) {
      const uint32_t cw = ctxw[mb_idx - 1];
      left_nz = (cw >> 8) & 0xff; left_dc = (cw >> 17) & 1;
      const uint8_t* lh = hdr - 48;
      if (lh[0] == 1) { lm[0] = lh[8 + 3]; lm[1] = lh[8 + 7]; lm[2] = lh[8 + 11]; lm[3] = lh[8 + 15]; }
    }
    S.top_nz = top_nz; S.left_nz = left_nz; S.top_nz_dc = top_dc; S.left_nz_dc = left_dc;
    for (int i = 0; i < 4; ++i) { S.top_modes[i] = tm[i]; S.left_modes[i] = lm[i]; }
    S.segment = P.segment[(size_t)img * nmb + mb_idx];
  }
}

// ---- phase: DC prediction values (border variants of checkMode folded in), isFlatSource16, accumulators
template <int M, int NT>
WG_HD void ph_prep(PhMB* mbs, int tid) {
  for (int t = tid; t < M * 4; t += NT) {
    PhMB& S = mbs[t >> 2];
    const int k = t & 3;
    if (!S.active) continue;
    if (k < 3) {
      const int size = k == 0 ? 16 : 8, shift = k == 0 ? 4 : 3;
      const uint8_t* d = S.out + (k == 0 ? Y_OFF : (k == 1 ? U_OFF : V_OFF));
      int s = 0, v = 128;
      if (S.my > 0) for (int i = 0; i < size; ++i) s += d[i - BPS];
      if (S.mx > 0) for (int i = 0; i < size; ++i) s += d[-1 + i * BPS];
      if (S.mx > 0 && S.my > 0) v = (s + size) >> (shift + 1);
      else if (S.mx > 0 || S.my > 0) v = (s + (size >> 1)) >> shift;
      S.dcv[k] = v;
    } else {
      const uint32_t v0 = (uint32_t)S.in[0] * 0x01010101u;  // isFlatSource16 (encode_analysis.go:358)
      int ok = 1;
      for (int i = 0; i < 64; ++i) ok &= (ph_ld32(S.in + 4 * i) == v0);
      S.src_flat = ok;
      int* a = reinterpret_cast<int*>(&S.u.s.acc);
      for (int i = 0; i < (int)(sizeof(PhAcc) / 4); ++i) a[i] = 0;
    }
  }
}

// task index -> (mode slot, macroblock, block) of the two-mode search passes: luma tasks first, chroma after, so that a
// warp holds one kind of task and one mode
template <int M>
WG_HD void ph_search_task(int t, int* m2, int* m, int* b) {
  if (t < 2 * M * 16) { *m2 = t / (M * 16); const int r = t % (M * 16); *m = r >> 4; *b = r & 15; }
  else { const int u = t - 2 * M * 16; *m2 = u / (M * 8); const int r = u % (M * 8); *m = r >> 3; *b = 16 + (r & 7); }
}
WG_HD void ph_pred_of_block(const PhMB& S, int b, int mode, int* p) {
  if (b < 16) ph_pred_square_block(S.out, Y_OFF, (b & 3) * 4, (b >> 2) * 4, mode, S.dcv[0], p);
  else { const int k = b - 16; ph_pred_square_block(S.out, (k & 4) ? V_OFF : U_OFF, (k & 1) * 4, ((k >> 1) & 1) * 4, mode, S.dcv[1 + (k >> 2)], p); }
}

// ---- I16 (pickBestI16ModeRDParallel, encode_parallel.go:624-735) and chroma (pickBestUVModeRDParallel, :1030-1116)
// searches, modes 2*pass and 2*pass+1.  A: transform + quantise every block (chroma also reconstructs: no WHT in the way).
template <int M, int NT>
WG_HD void ph_search_a(const EncKernelParams& P, PhMB* mbs, int pass, int tid) {
  for (int t = tid; t < 2 * M * 24; t += NT) {
    int m2, m, b;
    ph_search_task<M>(t, &m2, &m, &b);
    PhMB& S = mbs[m];
    const int mode = pass * 2 + m2;
    if (!S.active || !ph_mode_allowed(mode, S.mx, S.my)) continue;
    const SegParams& seg = P.img[S.img].seg[S.segment];
    PhSearch& Q = S.u.s;
    int s[16], p[16], c[16];
    ph_load_src(S.in, b, s);
    ph_pred_of_block(S, b, mode, p);
    ftransform(s, p, c);
    int16_t* lev = Q.lev[m2][b];
    int q[16];
    if (b < 16) {
      Q.dc[m2][b] = c[0];
      Q.nz[m2][b] = (uint8_t)quantize_block(c, q, seg.y1, 1);
      ph_store_lev(lev, q);
    } else {
      Q.nz[m2][b] = (uint8_t)quantize_block(c, q, seg.uv, 0);
      ph_store_lev(lev, q);
      int dq[16], r[16], ac = 0;
#pragma unroll
      for (int i = 1; i < 16; ++i) ac += (q[i] != 0);
      dequant_block(q, dq, seg.uv);
      itransform(p, dq, r);
      WG_ATOMIC_ADD(&Q.acc.disto_uv[mode], sse16(s, r));
      WG_ATOMIC_ADD(&Q.acc.ac_uv[mode], ac);
    }
  }
}
// B: the WHT of the sixteen luma DCs per mode: quantise, cost, reconstruct (encode_parallel.go:668-690)
template <int M, int NT>
WG_HD void ph_search_b(const EncKernelParams& P, PhMB* mbs, const CostTabs& T, int pass, int tid) {
  for (int t = tid; t < 2 * M; t += NT) {
    const int m2 = t / M;
    PhMB& S = mbs[t % M];
    const int mode = pass * 2 + m2;
    if (!S.active || !ph_mode_allowed(mode, S.mx, S.my)) continue;
    const SegParams& seg = P.img[S.img].seg[S.segment];
    PhSearch& Q = S.u.s;
    int d[16], w[16], q[16], dq[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = Q.dc[m2][i];
    fwht(d, w);
    const int nz_dc = quantize_block(w, q, seg.y2, 0);
    Q.acc.rate16[mode] = kModeFixedCost16(mode) + token_cost(q, nz_dc, 1, min(S.top_nz_dc + S.left_nz_dc, 2), 0, T);
    dequant_block(q, dq, seg.y2);
    iwht(dq, d);
    ph_store_lev(Q.dcrec[m2], d);
  }
}
// C: token costs with the in-macroblock NZ contexts, luma reconstruction + SSE + TDisto
template <int M, int NT>
WG_HD void ph_search_c(const EncKernelParams& P, PhMB* mbs, const CostTabs& T, int pass, int tid) {
  for (int t = tid; t < 2 * M * 24; t += NT) {
    int m2, m, b;
    ph_search_task<M>(t, &m2, &m, &b);
    PhMB& S = mbs[m];
    const int mode = pass * 2 + m2;
    if (!S.active || !ph_mode_allowed(mode, S.mx, S.my)) continue;
    const SegParams& seg = P.img[S.img].seg[S.segment];
    PhSearch& Q = S.u.s;
    if (b < 16) {
      const int bx = b & 3, by = b >> 2;
      const int l = bx > 0 ? (Q.nz[m2][b - 1] > 0) : ((S.left_nz >> by) & 1);
      const int tt = by > 0 ? (Q.nz[m2][b - 4] > 0) : ((S.top_nz >> bx) & 1);
      int q[16], dq[16], p[16], r[16], s[16];
      ph_load_lev(Q.lev[m2][b], q);
      const int rate = token_cost(q, Q.nz[m2][b], 0, l + tt, 1, T);
      dequant_block(q, dq, seg.y1);
      dq[0] = Q.dcrec[m2][b];
      ph_pred_of_block(S, b, mode, p);
      itransform(p, dq, r);
      ph_load_src(S.in, b, s);
      WG_ATOMIC_ADD(&Q.acc.rate16[mode], rate);
      WG_ATOMIC_ADD(&Q.acc.disto16[mode], sse16(s, r));
      if (seg.tlambda_sd > 0) WG_ATOMIC_ADD(&Q.acc.td16[mode], tdisto4x4(s, r));
      if (Q.nz[m2][b] > 0) WG_ATOMIC_ADD(&Q.acc.anyac16[mode], 1);
    } else {
      const int k = b - 16, ch = k >> 2, bx = k & 1, by = (k >> 1) & 1;
      const int tn = (S.top_nz >> (4 + 2 * ch)) & 3, ln = (S.left_nz >> (4 + 2 * ch)) & 3;
      const int l = bx > 0 ? (Q.nz[m2][b - 1] > 0) : ((ln >> by) & 1);
      const int tt = by > 0 ? (Q.nz[m2][b - 2] > 0) : ((tn >> bx) & 1);
      int q[16];
      ph_load_lev(Q.lev[m2][b], q);
      WG_ATOMIC_ADD(&Q.acc.rate_uv[mode], token_cost(q, Q.nz[m2][b], 2, l + tt, 0, T));
    }
  }
}
// decisions of both searches (strict '<' in mode order, as the reference's loops)
template <int M, int NT>
WG_HD void ph_decide(const EncKernelParams& P, PhMB* mbs, int tid) {
  for (int t = tid; t < 2 * M; t += NT) {
    PhMB& S = mbs[t % M];
    if (!S.active) continue;
    const SegParams& seg = P.img[S.img].seg[S.segment];
    const PhAcc& A = S.u.s.acc;
    unsigned long long best = ~0ull;
    if (t < M) {
      int best16 = 0, rate16 = 0, disto16 = 0;
      for (int mode = 0; mode < 4; ++mode) {
        if (!ph_mode_allowed(mode, S.mx, S.my)) continue;
        int disto = A.disto16[mode];
        if (seg.tlambda_sd > 0) disto += (seg.tlambda_sd * A.td16[mode] + 128) >> 8;
        if (S.src_flat && A.anyac16[mode] == 0) disto *= 2;
        const unsigned long long score = rd_score(disto, A.rate16[mode], seg.lambda_i16);
        if (score < best) { best = score; best16 = mode; rate16 = A.rate16[mode]; disto16 = disto; }
      }
      S.best16 = best16;
      S.score16 = rd_score(disto16, rate16, seg.lambda_mode);
    } else {
      int best_uv = 0;
      for (int mode = 0; mode < 4; ++mode) {
        if (!ph_mode_allowed(mode, S.mx, S.my)) continue;
        int rate = A.rate_uv[mode] + kModeFixedCostUV(mode);
        if (mode > 0 && A.ac_uv[mode] <= 2) rate += 140 * 8;
        const unsigned long long score = rd_score(A.disto_uv[mode], rate, seg.lambda_uv);
        if (score < best) { best = score; best_uv = mode; }
      }
      S.best_uv = best_uv;
    }
  }
}

// ---- I4 search (tryI4ModesRDParallel, encode_parallel.go:738-1027) as an anti-diagonal wavefront over the sub-blocks
template <int M, int NT>
WG_HD void ph_i4_init(PhMB* mbs, int tid) {
  for (int t = tid; t < M * 37; t += NT) {
    PhMB& S = mbs[t / 37];
    const int i = t % 37;
    if (!S.active) continue;
    if (i < 36) ph_cp16(S.u.q.out2 + 16 * i, S.out + 16 * i);
    else { S.tot_rate = 0; S.tot_disto = 0; S.tot_hdr = 0; S.nzmask = 0; S.modes_lo = 0; S.modes_hi = 0; }
  }
}
// The exit test of tryI4ModesRDParallel (encode_parallel.go:830) on the totals committed so far (see the file comment)
WG_HD bool ph_i4_alive(const EncKernelParams& P, const PhMB& S) {
  return S.active && !(rd_score(S.tot_disto, S.tot_rate + 211, P.img[S.img].seg[S.segment].lambda_mode) >= S.score16 || S.tot_hdr > 15000);
}
// sub-block of wavefront step `step` in slot j (bx + 2*by == step), or -1
WG_HD int ph_i4_block(int step, int j) {
  const int by_lo = max(0, (step - 2) >> 1), by_hi = min(3, step >> 1);
  const int by = by_lo + j;
  if (by > by_hi) return -1;
  const int bx = step - 2 * by;
  if (bx < 0 || bx > 3) return -1;
  return by * 4 + bx;
}
WG_HD int ph_get_mode(const PhMB& S, int k) { return (k < 8) ? (S.modes_lo >> (4 * k)) & 15 : (S.modes_hi >> (4 * (k - 8))) & 15; }
// P1: prediction SSE of every eligible mode (the pre-screen of pickBestI4ModeRDParallel, encode_parallel.go:955-966)
template <int M, int NT>
WG_HD void ph_i4_prescreen(const EncKernelParams& P, PhMB* mbs, int step, int tid) {
  for (int t = tid; t < 10 * 2 * M; t += NT) {
    const int mode = t / (2 * M), r = t % (2 * M), j = r & 1;
    PhMB& S = mbs[r >> 1];
    PhI4Slot& W = S.u.q.slot[j];
    const int b = ph_i4_block(step, j);
    const bool valid = b >= 0 && ph_i4_alive(P, S);
    if (!valid) { if (mode == 0) W.valid = 0; continue; }
    const int bx = b & 3, by = b >> 2;
    const bool has_top = S.my > 0 || by > 0, has_left = S.mx > 0 || bx > 0;
    uint32_t elig = 0;
#pragma unroll
    for (int k = 0; k < 10; ++k)
      if (!((!has_top && needs_top4(k)) || (!has_left && needs_left4(k)))) elig |= 1u << k;
    if (mode == 0) {
      W.valid = 1; W.b = (uint8_t)b; W.ncand = (uint8_t)WG_POPC(elig);
      W.top_mode = (uint8_t)(by == 0 ? S.top_modes[bx] : ph_get_mode(S, b - 4));
      W.left_mode = (uint8_t)(bx == 0 ? S.left_modes[by] : ph_get_mode(S, b - 1));
      const int l = bx > 0 ? ((S.nzmask >> (b - 1)) & 1) : ((S.left_nz >> by) & 1);
      const int tt = by > 0 ? ((S.nzmask >> (b - 4)) & 1) : ((S.top_nz >> bx) & 1);
      W.nz_ctx = (uint8_t)(l + tt);
    }
    if (!((elig >> mode) & 1)) continue;
    const uint8_t* pb = S.u.q.out2 + Y_OFF + by * 4 * BPS + bx * 4;
    int e[13], s[16], p[16];
    e[0] = pb[-BPS - 1];
    {
      const uint32_t a = ph_ld32(pb - BPS), c = ph_ld32(pb - BPS + 4);
      e[1] = a & 0xff; e[2] = (a >> 8) & 0xff; e[3] = (a >> 16) & 0xff; e[4] = a >> 24;
      e[5] = c & 0xff; e[6] = (c >> 8) & 0xff; e[7] = (c >> 16) & 0xff; e[8] = c >> 24;
    }
    e[9] = pb[-1]; e[10] = pb[-1 + BPS]; e[11] = pb[-1 + 2 * BPS]; e[12] = pb[-1 + 3 * BPS];
    ph_load_src(S.in, b, s);
    pred4(mode, e, p);
    const int pos = WG_POPC(elig & ((1u << mode) - 1));
    W.sse[pos] = sse16(s, p);
    W.smode[pos] = (uint8_t)mode;
    ph_store4x4<4>(W.pred[mode], p);
  }
}
// P2: the reference's selection sort of the first K entries, literally (encode_parallel.go:969-983); source half of TDisto
template <int M, int NT>
WG_HD void ph_i4_sort(const EncKernelParams& P, PhMB* mbs, int tid) {
  for (int t = tid; t < 2 * 2 * M; t += NT) {
    const int what = t / (2 * M), r = t % (2 * M);
    PhMB& S = mbs[r >> 1];
    PhI4Slot& W = S.u.q.slot[r & 1];
    if (!S.active || !W.valid) continue;  // W.valid was written in P1 for every slot of an active macroblock
    if (what == 0) {
      const int n_cand = W.ncand, K = min(P.max_i4_modes, n_cand);
      W.K = (uint8_t)K;
      int sse[10], md[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) { sse[i] = W.sse[i]; md[i] = W.smode[i]; }  // entries >= n_cand are never looked at
#pragma unroll
      for (int i = 0; i < 3; ++i) {  // K <= 3 rounds (getMaxI4RDModes)
        if (i < K) {
          int mi = i, mv = sse[i], mm = md[i];
#pragma unroll
          for (int jj = i + 1; jj < 10; ++jj)
            if (jj < n_cand && sse[jj] < mv) { mv = sse[jj]; mm = md[jj]; mi = jj; }
          const int ts = sse[i], tm = md[i];
#pragma unroll
          for (int jj = i + 1; jj < 10; ++jj)
            if (jj == mi) { sse[jj] = ts; md[jj] = tm; }
          sse[i] = mv; md[i] = mm;
        }
      }
      W.smode[0] = (uint8_t)md[0]; W.smode[1] = (uint8_t)md[1]; W.smode[2] = (uint8_t)md[2];
    } else if (P.img[S.img].seg[S.segment].tlambda_sd > 0) {
      int s[16];
      ph_load_src(S.in, W.b, s);
      W.tsrc = ttransform(s);
    }
  }
}
// P3: full RD of the K candidates (encode_parallel.go:985-1027): transform, trellis / quantise, reconstruct, SSE + TDisto,
// flatness penalty, token cost, mode cost
template <int M, int NT>
WG_HD void ph_i4_rd(const EncKernelParams& P, PhMB* mbs, const CostTabs& T, const uint16_t* i4cost, int tid) {
  for (int t = tid; t < 3 * 2 * M; t += NT) {
    const int k = t / (2 * M), r_ = t % (2 * M);
    PhMB& S = mbs[r_ >> 1];
    PhI4Slot& W = S.u.q.slot[r_ & 1];
    if (!S.active || !W.valid || k >= W.K) continue;
    const SegParams& seg = P.img[S.img].seg[S.segment];
    const int mode = W.smode[k], nz_ctx = W.nz_ctx;
    int s[16], p[16], c[16], q[16], dq[16], r[16];
    ph_load_src(S.in, W.b, s);
    ph_load4x4<4>(W.pred[mode], p);
    ftransform(s, p, c);
    int16_t* lev = W.lev[k];
    int nz;
    if (P.method >= 4) {
      ph_store_lev(lev, c);
      nz = trellis_block_v3(lev, seg.y1, 0, 3, nz_ctx, seg.tlambda_i4, T);
      ph_load_lev(lev, q);
    } else {
      nz = quantize_block(c, q, seg.y1, 0);
      ph_store_lev(lev, q);
    }
    dequant_block(q, dq, seg.y1);
    itransform(p, dq, r);
    int disto = sse16(s, r);
    if (seg.tlambda_sd > 0) disto += (seg.tlambda_sd * (abs(ttransform(r) - W.tsrc) >> 5) + 128) >> 8;
    int rate = 0;
    if (mode > 0) {  // isFlat(levels, 1, 3) (encode_analysis.go:374)
      int cnt = 0;
#pragma unroll
      for (int i = 1; i < 16; ++i) cnt += (q[i] != 0);
      if (cnt <= 3) rate = 140;
    }
    rate += token_cost(q, nz, 3, nz_ctx, 0, T);
    rate += i4cost[(W.top_mode * 10 + W.left_mode) * 10 + mode];
    W.score[k] = rd_score(disto, rate, seg.lambda_i4);
    W.disto[k] = disto; W.rate[k] = rate; W.nz[k] = (uint8_t)nz; W.cmode[k] = (uint8_t)mode;
    ph_store4x4<4>(W.rec[k], r);
  }
}
// P4: pick the winner of each sub-block in flight and commit it (modes, NZ bit, totals, levels, reconstruction).  One lane
// per sub-block; the two sub-blocks of a macroblock add into the same totals (integer sums: order does not matter).
template <int M, int NT>
WG_HD void ph_i4_commit(PhMB* mbs, const uint16_t* i4cost, int tid) {
  for (int t = tid; t < 2 * M; t += NT) {
    PhMB& S = mbs[t >> 1];
    PhI4Slot& W = S.u.q.slot[t & 1];
    if (!S.active || !W.valid) continue;
    const int K = W.K;
    unsigned long long sc[3];
    int di[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { sc[k] = W.score[k]; di[k] = W.disto[k]; }
    unsigned long long bs = ~0ull;
    int best = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < K && !(256ull * (unsigned long long)di[k] >= bs) && sc[k] < bs) { bs = sc[k]; best = k; }
    }
    const int b = W.b, bm = W.cmode[best];
    if (b < 8) WG_ATOMIC_OR(&S.modes_lo, (uint32_t)bm << (4 * b)); else WG_ATOMIC_OR(&S.modes_hi, (uint32_t)bm << (4 * (b - 8)));
    WG_ATOMIC_ADD(&S.tot_rate, W.rate[best]);
    WG_ATOMIC_ADD(&S.tot_disto, W.disto[best]);
    WG_ATOMIC_ADD(&S.tot_hdr, (int)i4cost[(W.top_mode * 10 + W.left_mode) * 10 + bm]);
    if (W.nz[best] > 0) WG_ATOMIC_OR(&S.nzmask, 1u << b);
    S.hdr[24 + b] = W.nz[best];
    ph_cp16(S.u.q.lev4[b], W.lev[best]);
    ph_cp16(S.u.q.lev4[b] + 8, W.lev[best] + 8);
    uint8_t* o = S.u.q.out2 + Y_OFF + (b >> 2) * 4 * BPS + (b & 3) * 4;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) ph_st32(o + rr * BPS, ph_ld32(W.rec[best] + 4 * rr));
  }
}
// ---- final residuals and reconstruction (encode_parallel.go:1164-1407)
// F0: the decision (encode_parallel.go:572-592); an I4 macroblock takes the trial buffer as its reconstruction
template <int M, int NT>
WG_HD void ph_final_decide(const EncKernelParams& P, PhMB* mbs, int tid) {
  for (int t = tid; t < M * 16; t += NT) {
    PhMB& S = mbs[t >> 4];
    const int i = t & 15;
    if (!S.active) continue;
    const bool use_i4 = ph_i4_alive(P, S);  // on the complete totals: score4 < score16 and <= 15000 header bits
    if (i == 0) S.use_i4 = use_i4;
    if (use_i4) {
      ph_cp8(S.out + Y_OFF + i * BPS, S.u.q.out2 + Y_OFF + i * BPS);
      ph_cp8(S.out + Y_OFF + i * BPS + 8, S.u.q.out2 + Y_OFF + i * BPS + 8);
      S.hdr[8 + i] = (uint8_t)ph_get_mode(S, i);
    } else {
      S.hdr[8 + i] = 0;
    }
  }
}
// F1: I16 macroblocks: forward transform against the winning prediction (raw coefficients for the trellis / quantiser);
// chroma of every macroblock: transform, quantise, reconstruct with the winning mode
template <int M, int NT>
WG_HD void ph_final_transform(const EncKernelParams& P, PhMB* mbs, int tid) {
  for (int t = tid; t < M * 24; t += NT) {
    int m, b;
    if (t < M * 16) { m = t >> 4; b = t & 15; } else { const int u = t - M * 16; m = u >> 3; b = 16 + (u & 7); }
    PhMB& S = mbs[m];
    if (!S.active || (b < 16 && S.use_i4)) continue;
    const SegParams& seg = P.img[S.img].seg[S.segment];
    PhFinal& F = S.u.f;
    int s[16], p[16], c[16];
    ph_load_src(S.in, b, s);
    ph_pred_of_block(S, b, b < 16 ? S.best16 : S.best_uv, p);
    ftransform(s, p, c);
    int16_t* lev = F.lev[b];
    if (b < 16) {
      F.dc[b] = c[0];
      c[0] = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) lev[i] = (int16_t)c[i];
    } else {
      int q[16], dq[16], r[16];
      const int nz = quantize_block(c, q, seg.uv, 0);
      ph_store_lev(lev, q);
      F.nz[b] = (uint8_t)nz;
      S.hdr[24 + b] = (uint8_t)nz;
      dequant_block(q, dq, seg.uv);
      itransform(p, dq, r);
      ph_store4x4<BPS>(S.out + ph_block_off(b), r);
    }
  }
}
// F2: I16 AC levels.  With trellis the NZ contexts chain the blocks: anti-diagonal d of the 4x4 block grid per call
// (7 calls); without, one call (d < 0) quantises all blocks.  The WHT block rides along with d <= 0.
template <int M, int NT>
WG_HD void ph_final_i16_levels(const EncKernelParams& P, PhMB* mbs, const CostTabs& T, int d, int tid) {
  const int per_mb = d < 0 ? 16 : 4;
  for (int t = tid; t < M * (per_mb + 1); t += NT) {
    const int k = t / M;
    PhMB& S = mbs[t % M];
    if (!S.active || S.use_i4) continue;
    const SegParams& seg = P.img[S.img].seg[S.segment];
    PhFinal& F = S.u.f;
    if (k == per_mb) {  // WHT of the DCs (encode_parallel.go:1200-1225)
      if (d > 0) continue;
      int dd[16], w[16], q[16], dq[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) dd[i] = F.dc[i];
      fwht(dd, w);
      S.nz_dc = quantize_block(w, q, seg.y2, 0);
      ph_store_lev(F.lev[24], q);
      dequant_block(q, dq, seg.y2);
      iwht(dq, dd);
      ph_store_lev(F.dcrec, dd);
      continue;
    }
    int b;
    if (d < 0) b = k;
    else {
      const int by_lo = max(0, d - 3), by_hi = min(3, d), by = by_lo + k;
      if (by > by_hi) continue;
      b = by * 4 + (d - by);
    }
    int nz;
    if (d < 0) {
      int c[16], q[16];
      ph_load_lev(F.lev[b], c);
      nz = quantize_block(c, q, seg.y1, 1);
      ph_store_lev(F.lev[b], q);
    }
    else {
      const int bx = b & 3, by = b >> 2;
      const int l = bx > 0 ? (F.nz[b - 1] > 0) : ((S.left_nz >> by) & 1);
      const int tt = by > 0 ? (F.nz[b - 4] > 0) : ((S.top_nz >> bx) & 1);
      nz = trellis_block_v3(F.lev[b], seg.y1, 1, 0, l + tt, seg.tlambda_i16, T);
    }
    F.nz[b] = (uint8_t)nz;
    S.hdr[24 + b] = (uint8_t)nz;
  }
}
// F3: I16 reconstruction; header bytes and the NZ context word (encode_parallel.go:341-428, 1410-1496)
template <int M, int NT>
WG_HD void ph_final_recon(const EncKernelParams& P, PhMB* mbs, int tid) {
  for (int t = tid; t < M * 17; t += NT) {
    const int k = t / M;
    PhMB& S = mbs[t % M];
    if (!S.active) continue;
    PhFinal& F = S.u.f;
    if (k < 16) {
      if (S.use_i4) { F.lev[24][k] = 0; continue; }  // no WHT block on I4 macroblocks
      const SegParams& seg = P.img[S.img].seg[S.segment];
      int q[16], dq[16], p[16], r[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) q[i] = F.lev[k][i];
      dequant_block(q, dq, seg.y1);
      dq[0] = F.dcrec[k];
      ph_pred_of_block(S, k, S.best16, p);
      itransform(p, dq, r);
      ph_store4x4<BPS>(S.out + ph_block_off(k), r);
      continue;
    }
    const bool i16 = !S.use_i4;
    uint32_t nzy = 0, nzuv = 0;
    if (i16) { for (int b = 0; b < 16; ++b) nzy |= (uint32_t)(F.nz[b] > 0) << b; } else nzy = S.nzmask;
    for (int b = 0; b < 8; ++b) nzuv |= (uint32_t)(F.nz[16 + b] > 0) << b;
    const int nz_dc = i16 ? S.nz_dc : 0;
    const bool skip = nzy == 0 && nz_dc == 0 && nzuv == 0;
    S.hdr[0] = i16 ? 0 : 1; S.hdr[1] = (uint8_t)(i16 ? S.best16 : 0); S.hdr[2] = (uint8_t)S.best_uv; S.hdr[3] = (uint8_t)S.segment;
    S.hdr[4] = skip ? 1 : 0; S.hdr[5] = (uint8_t)nz_dc; S.hdr[6] = 0; S.hdr[7] = 0;
    const uint32_t out_t = ((nzy >> 12) & 0xf) | (((nzuv >> 2) & 3) << 4) | (((nzuv >> 6) & 3) << 6);
    const uint32_t yl = ((nzy >> 3) & 1) | (((nzy >> 7) & 1) << 1) | (((nzy >> 11) & 1) << 2) | (((nzy >> 15) & 1) << 3);
    const uint32_t ul = ((nzuv >> 1) & 1) | (((nzuv >> 3) & 1) << 1);
    const uint32_t vl = ((nzuv >> 5) & 1) | (((nzuv >> 7) & 1) << 1);
    const int dcflag = nz_dc > 0;
    const int tdc = i16 ? dcflag : S.top_nz_dc, ldc = i16 ? dcflag : S.left_nz_dc;
    const int nmb = P.mb_w * P.mb_h;
    P.ctx[(size_t)S.img * nmb + S.my * P.mb_w + S.mx] = pack_ctx(out_t, yl | (ul << 4) | (vl << 6), tdc, ldc);
  }
}
// F4: export -- reconstruction planes, the 48-byte header and the 400 levels, all as 8 / 16-byte stores
template <int M, int NT>
WG_HD void ph_export(const EncKernelParams& P, PhMB* mbs, int tid) {
  const int nmb = P.mb_w * P.mb_h;
  for (int t = tid; t < M * 85; t += NT) {
    PhMB& S = mbs[t / 85];
    const int i = t % 85;
    if (!S.active) continue;
    const int mb_idx = S.my * P.mb_w + S.mx;
    if (i < 16) {
      uint8_t* d = P.rec_y + (size_t)S.img * P.y_plane + (size_t)(S.my * 16 + i) * (P.mb_w * 16) + S.mx * 16;
      ph_cp8(d, S.out + Y_OFF + i * BPS);
      ph_cp8(d + 8, S.out + Y_OFF + i * BPS + 8);
    } else if (i < 32) {
      const int pl = (i >> 3) & 1, r = i & 7;
      uint8_t* d = (pl ? P.rec_v : P.rec_u) + (size_t)S.img * P.uv_plane + (size_t)(S.my * 8 + r) * (P.mb_w * 8) + S.mx * 8;
      ph_cp8(d, S.out + (pl ? V_OFF : U_OFF) + r * BPS);
    } else if (i < 35) {
      ph_cp16(P.out_hdr + ((size_t)S.img * nmb + mb_idx) * 48 + (i - 32) * 16, S.hdr + (i - 32) * 16);
    } else {
      ph_cp16(P.out_coeffs + ((size_t)S.img * nmb + mb_idx) * 400 + (i - 35) * 8, &S.u.f.lev[0][0] + (i - 35) * 8);
    }
  }
}

// The whole per-CTA schedule.  WG_PH(stmt) runs `stmt` for every thread of the CTA and ends with a barrier.
#if defined(__CUDA_ARCH__) && defined(WG_PHASE_CLOCK)
// profiling build only (tools/phase_clock.py): CTA 0 timestamps every phase boundary
#define WG_PH(stmt) { stmt; } __syncthreads(); if (P.phase_clock && (int)blockIdx.x == P.clock_cta && threadIdx.x == 0) P.phase_clock[ph_i_++] = clock64()
#elif defined(__CUDA_ARCH__)
#define WG_PH(stmt) { stmt; } __syncthreads()
#else
#define WG_PH(stmt) for (int k_ = 0; k_ < NT; ++k_) { const int tid = order ? order[k_] : k_; stmt; }
#endif
template <int M, int NT>
WG_HD void ph_run_cta(const EncKernelParams& P, PhMB* mbs, const CostTabs& T, const uint16_t* i4cost, int wave, long long task_base,
                      int tid, const int* order) {
  (void)tid; (void)order;
#if defined(__CUDA_ARCH__) && defined(WG_PHASE_CLOCK)
  int ph_i_ = 0;
  if (P.phase_clock && (int)blockIdx.x == P.clock_cta && threadIdx.x == 0) P.phase_clock[ph_i_++] = clock64();
#endif
  WG_PH((ph_load<M, NT>(P, mbs, wave, task_base, tid)));
  WG_PH((ph_prep<M, NT>(mbs, tid)));
  for (int pass = 0; pass < 2; ++pass) {
    WG_PH((ph_search_a<M, NT>(P, mbs, pass, tid)));
    WG_PH((ph_search_b<M, NT>(P, mbs, T, pass, tid)));
    WG_PH((ph_search_c<M, NT>(P, mbs, T, pass, tid)));
  }
  WG_PH((ph_decide<M, NT>(P, mbs, tid)));
  WG_PH((ph_i4_init<M, NT>(mbs, tid)));
  for (int step = 0; step < 10; ++step) {
    WG_PH((ph_i4_prescreen<M, NT>(P, mbs, step, tid)));
    WG_PH((ph_i4_sort<M, NT>(P, mbs, tid)));
    WG_PH((ph_i4_rd<M, NT>(P, mbs, T, i4cost, tid)));
    WG_PH((ph_i4_commit<M, NT>(mbs, i4cost, tid)));
  }
  WG_PH((ph_final_decide<M, NT>(P, mbs, tid)));
  WG_PH((ph_final_transform<M, NT>(P, mbs, tid)));
  if (P.method >= 4) {
    for (int d = 0; d < 7; ++d) { WG_PH((ph_final_i16_levels<M, NT>(P, mbs, T, d, tid))); }
  } else {
    WG_PH((ph_final_i16_levels<M, NT>(P, mbs, T, -1, tid)));
  }
  WG_PH((ph_final_recon<M, NT>(P, mbs, tid)));
  WG_PH((ph_export<M, NT>(P, mbs, tid)));
}

#ifdef __CUDACC__
// One wave per launch; grid = ceil(macroblocks of the wave / M).  Cost tables staged once per CTA.
template <int M, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) encode_phased_kernel(const EncKernelParams P, int wave) {
  __shared__ __align__(16) uint16_t s_lc[LC_SIZE];
  __shared__ __align__(16) uint16_t s_lfc[LFC_NEAR];
  __shared__ __align__(16) uint16_t s_i4cost[1000];
  __shared__ __align__(16) uint16_t s_eob[EOB_SIZE];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  PhMB* mbs = reinterpret_cast<PhMB*>(s_dyn);
  for (int i = threadIdx.x; i < LC_SIZE / 8; i += NT) reinterpret_cast<uint4*>(s_lc)[i] = reinterpret_cast<const uint4*>(P.lc)[i];
  for (int i = threadIdx.x; i < LFC_NEAR / 8; i += NT) reinterpret_cast<uint4*>(s_lfc)[i] = reinterpret_cast<const uint4*>(P.lfc)[i];
  for (int i = threadIdx.x; i < 1000 / 8; i += NT) reinterpret_cast<uint4*>(s_i4cost)[i] = reinterpret_cast<const uint4*>(P.i4_costs)[i];
  for (int i = threadIdx.x; i < EOB_SIZE / 8; i += NT) reinterpret_cast<uint4*>(s_eob)[i] = reinterpret_cast<const uint4*>(P.eob)[i];
  CostTabs T;
  T.lc = s_lc; T.eob = s_eob; T.lfc = s_lfc; T.lfc_hi = P.lfc;
  ph_run_cta<M, NT>(P, mbs, T, s_i4cost, wave, (long long)blockIdx.x * M, (int)threadIdx.x, nullptr);
}
#endif

}  // namespace wg
