// Definitions shared by the mode-search kernels (enc_kernels.cuh: wavefront kernels with G lanes per macroblock, used for the
// Method < 3 and serial RD paths; enc_phased.cuh: the phase-synchronous kernel of the row-parallel RD path) and by the CPU
// harness that runs the phased kernel's code (the hostcheck test harness).  Plain C++ when compiled without nvcc.
#pragma once
#include <stddef.h>
#include "vp8_dev.cuh"

namespace wg {

enum { BPS = 32, YUV_SIZE = BPS * 17 + BPS * 9, Y_OFF = BPS + 8, U_OFF = Y_OFF + BPS * 16 + BPS, V_OFF = U_OFF + 16 };

struct ImageParams {  // per image, written by the host after segmentation
  SegParams seg[4];
};

struct EncKernelParams {
  const uint8_t* src_y; const uint8_t* src_u; const uint8_t* src_v;  // padded source planes [n][..]
  uint8_t* rec_y; uint8_t* rec_u; uint8_t* rec_v;                    // padded reconstruction planes
  const uint8_t* segment;       // [n][nmb]
  const ImageParams* img;       // [n]
  uint32_t* ctx;                // [n][nmb] packed NZ context (see pack_ctx)
  int8_t* top_derr;             // [n][mb_w][2][2] serial RD path: DC error diffusion state (enc.topDerr)
  int8_t* left_derr;            // [n][2][2] (enc.leftDerr)
  const uint16_t* lc_img;       // [n][LC_SIZE] per-image folded level costs (serial path with probability refreshes), or null
  const uint16_t* eob_img;      // [n][EOB_SIZE]
  int serial_gpw;               // serial RD kernels with per-image tables: macroblock groups (images) per warp actually used, 1..32/G (0 = all)
  // Serial RD path split by plane (refresh route): luma runs as left/top/top-right waves over the macroblocks [mb_begin, mb_end) of a
  // refresh segment (serial_wave = 1: `wave` is x + 2y and tasks are (image, row) pairs), chroma as one raster-order chain per image
  // (its DC error diffusion carries leftDerr from macroblock to macroblock, encode_frame.go:529-566).  ctx_uv receives the chroma half
  // of the NZ context words; serial_merge_kernel joins the halves and sets the skip flag.
  int serial_wave, mb_begin, mb_end;
  uint32_t* ctx_uv;             // [n][nmb]
  uint32_t* ctx2;               // [n][nmb] Method < 3 / serial RD: trial 4x4 modes of the bottom row / right column (mode-cost context)
#ifdef WG_PHASE_CLOCK
  unsigned long long* phase_clock;  // profiling build: CTA `clock_cta` timestamps its phase boundaries
  int clock_cta;
#endif
  unsigned int* stats;          // [n][4][8][3][11][2] token statistics (ProbaStats, encode_proba.go), zeroed before the waves
  uint8_t* out_hdr;             // [n][nmb][48]: mb_type,i16,uv,segment,skip,nz_dc,0,0, modes[16], nz[24]
  int16_t* out_coeffs;          // [n][nmb][400]
  const uint16_t* i4_costs;     // [10][10][10]
  const uint16_t* lc; const uint16_t* eob; const uint16_t* lfc;  // folded cost tables (vp8_dev.cuh CostTabs), global
  int n_images, width, height, mb_w, mb_h;
  int method, max_i4_modes;
  size_t y_plane, uv_plane;     // bytes per image plane
  // Row-parallel path: image slot s of a wave's task list holds image img_order[s] (null: the identity).  The host sorts the
  // batch by expected cost, most expensive images first, so that a wave's longest CTAs start first instead of last.
  const int* img_order;
};

// ctx word: bits 0-7 out_t (4 Y, 2 U, 2 V), 8-15 out_l, 16 top-DC carry, 17 left-DC carry
WG_HD uint32_t pack_ctx(uint32_t out_t, uint32_t out_l, int top_dc, int left_dc) {
  return (out_t & 0xff) | ((out_l & 0xff) << 8) | ((uint32_t)top_dc << 16) | ((uint32_t)left_dc << 17);
}

// checkMode (internal/lossy/decode_frame.go:6)
WG_HD int check_mode(int mx, int my, int mode) {
  if (mode == 0) {
    if (mx == 0) return my == 0 ? 6 : 5;
    if (my == 0) return 4;
  }
  return mode;
}
WG_HD bool needs_top4(int m) { return m == 1 || m == 2 || m == 4 || m == 5 || m == 6 || m == 7 || m == 8; }
WG_HD bool needs_left4(int m) { return m == 1 || m == 3 || m == 4 || m == 8 || m == 9; }

// modeFixedCost16 / modeFixedCostUV (internal/lossy/encode_analysis.go:1481,1485)
WG_HD int kModeFixedCost16(int m) { return m == 0 ? 663 : (m == 2 ? 872 : 919); }
WG_HD int kModeFixedCostUV(int m) { return m == 0 ? 302 : (m == 1 ? 984 : (m == 2 ? 439 : 642)); }

WG_HD unsigned long long rd_score(int disto, int rate, int lambda) {
  return (unsigned long long)(long long)rate * (unsigned long long)(long long)lambda + 256ull * (unsigned long long)(long long)disto;
}

enum { STATS_SIZE = 4 * 8 * 3 * 11 * 2 };

}  // namespace wg
