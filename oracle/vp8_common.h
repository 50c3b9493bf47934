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

// Integer-operation accounting ("the oracle must carry exact per-stage op counters, compile-time switch").
// With -DORC_COUNT_OPS every primitive on the encoder's mode-search path counts its invocations (or, where the work is data
// dependent, its inner units: trellis positions and transitions actually evaluated, coefficients the cost walk visits,
// coefficients quantised).  kOpWeight is the number of scalar integer operations (add / sub / mul / shift / compare /
// select / abs / clip; loads, stores and address arithmetic not counted) of ONE unit in the reference's formulation, counted
// by hand from the cited Go source once; bench.py reports sum(count * weight) / time against the integer-issue peak.
enum OpStage { OP_FTRANSFORM, OP_ITRANSFORM, OP_FWHT, OP_IWHT, OP_QUANT_COEFF, OP_DEQUANT_BLOCK, OP_SSE4X4, OP_TTRANSFORM, OP_PRED4,
               OP_PRED_TM_PIXEL, OP_PRED_DC_SUM, OP_TRELLIS_PRESCAN_COEFF, OP_TRELLIS_POS, OP_TRELLIS_TRANS, OP_TRELLIS_TERMINAL,
               OP_TRELLIS_BACKTRACK, OP_TOKEN_COEFF, OP_MODE_SCORE, OP_STAGES };
static const char* const kOpName[OP_STAGES] = {"ftransform", "itransform", "fwht", "iwht", "quantize_coeff", "dequant_block", "sse4x4",
                                               "ttransform", "pred4", "pred_tm_pixel", "pred_dc_sum", "trellis_prescan_coeff", "trellis_pos",
                                               "trellis_transition", "trellis_terminal", "trellis_backtrack", "token_cost_coeff", "mode_score"};
// transforms.go:371 (2 x 4 butterflies of 22 ops), :265 (18 + 35 per row/column), :500 / :223 (WHT), encode_quant.go:16 per
// coefficient (sign, abs, +sharpen, clamp, mul, +bias, shift, min, sign, nz compare + max), :81 (16 mul), ssim.go:188 (16 x
// sub/mul/add), :266 (32 + 80), predict_lossy.go:185-424 (~30 averaged over the ten modes), TM: add, sub, clip(2) per pixel, DC:
// one add per border sample, encode_trellis.go: pre-scan 6 per coefficient, 30 per position (level / threshold / two distortion
// deltas), 8 per (valid previous context, candidate level) transition (rate sum, mul, two adds, compare, select), 4 per terminal
// check, 4 per backtrack step, encode_quant.go:170 6 per coefficient walked, RDScore 3.
static const unsigned kOpWeight[OP_STAGES] = {176, 212, 80, 84, 11, 16, 48, 112, 30, 4, 1, 6, 30, 8, 4, 4, 6, 3};
#ifdef ORC_COUNT_OPS
inline thread_local unsigned long long g_op_count[OP_STAGES] = {0};
#define ORC_COUNT(stage, n) (orc::g_op_count[orc::stage] += (unsigned long long)(n))
#else
#define ORC_COUNT(stage, n) ((void)0)
#endif

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
