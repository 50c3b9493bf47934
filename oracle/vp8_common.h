// ORACLE -- TEST INFRASTRUCTURE ONLY.
// CPU restatement of the deepteams/webp VP8 lossy pixel pipeline (pure-Go scalar code is
// the arithmetic spec).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg may build, link or call anything under
// oracle/.  The product (webp_b200/) never includes these files.
//
// Parity pinning: decoder/upsampler/import are pinned against libwebp 1.6.0 (Pillow) and
// the reference's two decode fixtures (tests/data/*.webp, reference testdata/).  The encoder
// *decision* path is pinned only by restating the cited source (no Go toolchain in this
// image; the reference holds no bitstream golden) -- "parity unpinned" for encoder bytes,
// mitigated by libwebp decoding every oracle stream to exactly the oracle's reconstruction.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

namespace orc {

#include "vp8_tables.inc"

// internal/lossy/constants.go:66-75
enum { BPS = 32, YUV_SIZE = BPS * 17 + BPS * 9, Y_OFF = BPS * 1 + 8, U_OFF = Y_OFF + BPS * 16 + BPS,
       V_OFF = U_OFF + 16 };

// internal/lossy/constants.go:6-36
enum { B_DC_PRED = 0, B_TM_PRED, B_VE_PRED, B_HE_PRED, B_RD_PRED, B_VR_PRED, B_LD_PRED, B_VL_PRED,
       B_HD_PRED, B_HU_PRED, NUM_BMODES };
enum { DC_PRED = 0, TM_PRED = 1, V_PRED = 2, H_PRED = 3, NUM_PRED_MODES = 4 };
enum { B_DC_PRED_NOTOP = 4, B_DC_PRED_NOLEFT = 5, B_DC_PRED_NOTOPLEFT = 6 };
enum { NUM_MB_SEGMENTS = 4, NUM_TYPES = 4, NUM_BANDS = 8, NUM_CTX = 3, NUM_PROBAS = 11 };

// internal/lossy/constants.go:79-88
static const uint8_t kBands[17] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0};
static const uint8_t kZigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
// internal/lossy/constants.go:258-262 (tree for the ten 4x4 modes; leaves are -mode)
static const int8_t kYModesIntra4[18] = {-B_DC_PRED, 1, -B_TM_PRED, 2, -B_VE_PRED, 3, 4, 6, -B_HE_PRED, 5,
                                         -B_RD_PRED, -B_VR_PRED, -B_LD_PRED, 7, -B_VL_PRED, 8, -B_HD_PRED,
                                         -B_HU_PRED};
// internal/lossy/constants.go:265-270
static const uint8_t kCat3[] = {173, 148, 140, 0};
static const uint8_t kCat4[] = {176, 155, 140, 135, 0};
static const uint8_t kCat5[] = {180, 157, 141, 134, 130, 0};
static const uint8_t kCat6[] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0};
static const uint8_t* const kCat3456[4] = {kCat3, kCat4, kCat5, kCat6};

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// [type][band][ctx][proba] coefficient probabilities (internal/lossy/proba.go:9-41)
struct Proba {
  uint8_t segments[3];
  uint8_t bands[NUM_TYPES][NUM_BANDS][NUM_CTX][NUM_PROBAS];
};
static inline void reset_proba(Proba* p) {
  p->segments[0] = p->segments[1] = p->segments[2] = 255;
  memcpy(p->bands, kCoeffsProba0, sizeof(p->bands));
}
static inline int bit_cost(int bit, uint8_t prob) {  // internal/dsp/cost.go:43
  return bit == 0 ? kEntropyCost[prob] : kEntropyCost[255 - prob];
}

}  // namespace orc
