/* webpgpu.h -- C ABI of the B200 (sm_100a) VP8 lossy pixel pipeline.
 *
 * Drop-in boundary for deepteams/webp: these are the entry points a `//go:build cuda && cgo`
 * file in internal/dsp and internal/lossy binds (see INTEGRATION.md for the cgo stubs).
 * Plain pointers and sizes only; no exceptions cross the boundary; no caller pointer is
 * retained after a call returns (cgo rule).  Every function returns 0 on success or a
 * negative wgpu_status; wgpu_last_error(ctx) gives the message.  One ctx == one GPU; calls on
 * one ctx are serialised by the library, different ctxs are independent.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   wgpu_encode_batch        lossy.NewEncoder + (*VP8Encoder).EncodeFrame   internal/lossy/encode.go:452,1324
 *                            (called from encodeLossyWithAlpha, encode.go:472-546) + writeRIFFSimple encode.go:968
 *   wgpu_decode_batch        lossy.DecodeFrame (+ buildYCbCr / buildNRGBA)    internal/lossy/decode.go:209, webp.go:351,379
 *   wgpu_import_rgba         (*VP8Encoder).importImage                       internal/lossy/encode.go:671
 *   wgpu_cleanup_transparent cleanupTransparentAreaLossy                      encode.go:788
 *   wgpu_upsample_nrgba      buildNRGBA / dsp.UpsampleLinePairNRGBA           webp.go:379, internal/dsp/upsample.go:130
 *   wgpu_plane_metrics       dsp.SSE / SSIMGet / SSIMGetClipped / PSNRFromSSE internal/dsp/ssim.go:116-181
 *   wgpu_dsp_*_batch         the per-block operator surface                  internal/dsp/dsp.go:12-37,
 *                            FTransformDirect/ITransformDirect/SSE4x4Direct/TDisto4x4/PredLuma4Direct,
 *                            lossy.QuantizeCoeffs, lossy.TrellisQuantizeBlock  internal/lossy/encode_quant.go:16, encode_trellis.go:23
 */
#ifndef WEBPGPU_H_
#define WEBPGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wgpu_ctx wgpu_ctx;

typedef enum {
  WGPU_OK = 0,
  WGPU_ERR_INVALID = -1,     /* bad argument / option out of range (validateConfig, encode.go:259) */
  WGPU_ERR_UNSUPPORTED = -2, /* configuration routed to a reference path not built yet (serial path) */
  WGPU_ERR_CUDA = -3,        /* CUDA runtime failure; there is NO CPU fallback */
  WGPU_ERR_NOMEM = -4,
  WGPU_ERR_BITSTREAM = -5,   /* malformed VP8/RIFF input */
  WGPU_ERR_TOO_SMALL = -6    /* caller output buffer too small */
} wgpu_status;

/* lossy.EncodeConfig (internal/lossy/encode.go:46-63) after the EncoderOptions mapping of
 * encode.go:478-528; defaults = lossy.DefaultConfig (encode.go:66-86). */
typedef struct {
  int quality;          /* 0..100 */
  int method;           /* 0..6; >=3: row-parallel path semantics (serial RD semantics when height <= 48 or a target is set: <= 96 macroblocks); <3: statLoop + serial encodeFrame semantics */
  int sns_strength;     /* 0..100, default 50 */
  int filter_strength;  /* 0..100, default 60 */
  int filter_sharpness; /* 0..7 */
  int filter_type;      /* 0 simple, 1 strong */
  int partitions;       /* 0..3 (log2) */
  int segments;         /* 1..4 */
  int preprocessing;    /* bit0: segment smoothing */
  int has_alpha;        /* 0: opaque input (alpha bytes ignored) */
  int passes;           /* EncodeConfig.Pass, 1..10 (statLoop iterations on the Method < 3 path); 0 = 1 */
  int dither_amp;       /* VP8Random.amp = int(256 * EncodeConfig.Dithering), 0..256 (encode.go:563-567, dsp/random.go:39); 0 = off */
  int target_size;      /* EncodeConfig.TargetSize in bytes; > 0: size search (doSearch, internal/lossy/encode.go:1338) */
  float target_psnr;    /* EncodeConfig.TargetPSNR in dB; > 0: the reference's "PSNR search" (it measures 99.0 dB every pass) */
  int qmin, qmax;       /* EncodeConfig.QMin / QMax after resolveQMax (encode.go:305): quality clamp of the search; qmax <= 0 means 100 */
  int use_sharp_yuv;    /* EncoderOptions.UseSharpYUV (encode.go:62, 531-535): source planes from sharpyuv.Convert (WebP matrix, sRGB transfer) instead of importImage; has_alpha / dither_amp then have no effect on the planes */
} wgpu_enc_options;

void wgpu_enc_options_default(wgpu_enc_options* o, int quality);

/* ---- context ------------------------------------------------------------------------- */
int wgpu_ctx_create(int device_ordinal, wgpu_ctx** out);
void wgpu_ctx_destroy(wgpu_ctx* ctx);
const char* wgpu_last_error(const wgpu_ctx* ctx); /* ctx may be NULL: last create error */
int wgpu_sync(wgpu_ctx* ctx);
/* host worker threads used for bitstream serialisation / parsing (0 = hardware concurrency) */
int wgpu_set_host_threads(wgpu_ctx* ctx, int n);
/* pinned host staging (device image-batch allocator; internal/pool analogue) */
void* wgpu_host_alloc(wgpu_ctx* ctx, size_t bytes);
void wgpu_host_free(wgpu_ctx* ctx, void* p);
/* Device image-batch allocator -- replaces the bucketed byte pools of internal/pool/pool.go:14-71 (Get / Put, bucketIndex) and
 * the encoder / decoder object pools of internal/lossy/encode.go:391 (resetForReuse) for the GPU side.  A context's device and
 * pinned buffers are grow-only and sized in classes: wgpu_pool_bucket(bytes) = the reference's seven buckets up to 1 MiB, above
 * that the next sixteenth of the enclosing power of two.  wgpu_ctx_mem_info reports the bytes a context holds and how many
 * buffers are live; wgpu_ctx_trim returns the working buffers to the driver (constant tables stay) and forgets whatever picture
 * state they held -- the next call re-reserves.  Results never depend on what a buffer held before (tests/test_gpu_codec.py). */
size_t wgpu_pool_bucket(size_t bytes);
int wgpu_ctx_mem_info(wgpu_ctx* ctx, size_t* device_bytes, size_t* pinned_bytes, int* buffers);
int wgpu_ctx_trim(wgpu_ctx* ctx);

/* ---- encoder: whole job --------------------------------------------------------------- */
/* Encode n same-size RGBA images (4 B/px, `stride` bytes per row, image i at rgba + i*image_stride).
 * out receives n RIFF/WebP files, file i at out + i*out_stride, length out_sizes[i].
 * Output bytes are identical to the reference's webp.Encode on its parallel path
 * (Method>=3, single pass, GOMAXPROCS>1, height>48). */
int wgpu_encode_batch(wgpu_ctx* ctx, const uint8_t* rgba, int n, int width, int height, int stride,
                      size_t image_stride, const wgpu_enc_options* opt, uint8_t* out, size_t out_stride,
                      size_t* out_sizes);

/* Staged form of the same job (used by bench.py to separate H2D / device / host time):
 *   upload  : host RGBA -> HBM (async on the ctx stream)
 *   device  : import + analysis + segmentation + wavefront mode search, results stay in HBM
 *   finish  : D2H of per-MB modes/levels + host token/probability/bool coding -> RIFF bytes */
int wgpu_enc_upload(wgpu_ctx* ctx, const uint8_t* rgba, int n, int width, int height, int stride, size_t image_stride);
int wgpu_enc_device(wgpu_ctx* ctx, const wgpu_enc_options* opt);
int wgpu_enc_finish(wgpu_ctx* ctx, uint8_t* out, size_t out_stride, size_t* out_sizes);

/* Finer-grained form for a host that keeps the reference's own segmentation and serialiser (the Go shim):
 *   analyze : import + per-macroblock analysis; alphas [n][nmb] (mixed alpha, encode_analysis.go:245) and
 *             uv_alpha_sum [n] (sum of per-MB chroma alphas) are copied back
 *   search  : the host supplies what assignSegments / setSegmentParams / setupSegment produced -- per image four
 *             wgpu_segment (SegmentInfo subset, internal/lossy/encode.go:278-323) and the per-MB segment map --
 *             and the wavefront mode search runs; fetch per-MB results with wgpu_enc_fetch.
 * wgpu_enc_device == analyze + the library's own restatement of that host step + search. */
typedef struct {
  int quant, iquant, bias, dc_quant, dc_iquant, dc_bias; /* SegmentQuant: Quant, IQuant, Bias, DCQuant, DCIQuant, DCBias */
  int16_t sharpen[16];
} wgpu_seg_quant;
typedef struct {
  wgpu_seg_quant y1, y2, uv;
  int lambda_i4, lambda_i16, lambda_uv, lambda_mode, tlambda_i4, tlambda_i16, tlambda_sd, reserved; /* reserved: 0 */
} wgpu_segment;
/* setupSegment (internal/lossy/encode.go:1084): quantiser matrices, biases, sharpening and lambdas of one segment from
 * its quantiser index; provided so hosts and tests can build wgpu_segment without restating the tables. */
int wgpu_setup_segment(int quant_index, int dq_uv_dc, int dq_uv_ac, int method, int sns_strength, wgpu_segment* out);
int wgpu_enc_analyze(wgpu_ctx* ctx, const wgpu_enc_options* opt, uint8_t* alphas, int64_t* uv_alpha_sum);
int wgpu_enc_search(wgpu_ctx* ctx, const wgpu_segment* segments /* [n][4] */, const uint8_t* segment_map /* [n][nmb] */);

/* Debug/parity taps of the last wgpu_enc_device call (any pointer may be NULL).  Layouts as the
 * oracle's taps: mb_hdr [nmb][8] = {mb_type,i16_mode,uv_mode,segment,skip,nz_dc,0,0}, mb_modes [nmb][16],
 * mb_nz [nmb][24], mb_coeffs [nmb][400] int16, planes padded to 16*mb_w x 16*mb_h (chroma half). */
int wgpu_enc_fetch(wgpu_ctx* ctx, int image, uint8_t* mb_hdr, uint8_t* mb_modes, uint8_t* mb_nz, int16_t* mb_coeffs,
                   uint8_t* recon_y, uint8_t* recon_u, uint8_t* recon_v, uint8_t* src_y, uint8_t* src_u,
                   uint8_t* src_v, uint8_t* alphas);

/* ---- decoder: whole job --------------------------------------------------------------- */
/* Header probe (webp.DecodeConfig): dimensions of a RIFF/WebP or raw VP8 lossy stream. */
int wgpu_decode_info(const uint8_t* data, size_t len, int* width, int* height);
/* Decode n lossy streams of identical dimensions.  Host parses headers/modes/tokens (bool
 * decoder), the GPU reconstructs (predict + inverse transform, wavefront), loop-filters and
 * optionally converts.  Planes are written with stride 16*mb_w (luma) / 8*mb_w (chroma),
 * plane i at y + i*y_plane_stride etc.  nrgba (optional) gets width*height*4 bytes per image
 * with buildNRGBA semantics (webp.go:379) and A=255. */
int wgpu_decode_batch(wgpu_ctx* ctx, const uint8_t* const* streams, const size_t* lens, int n, uint8_t* y,
                      uint8_t* u, uint8_t* v, size_t y_plane_stride, size_t uv_plane_stride, uint8_t* nrgba,
                      size_t nrgba_image_stride);

/* Staged form of the same job (bench.py separates host parse + H2D / device / D2H):
 *   parse  : host boolean decoding of every stream into per-macroblock coefficients + modes, then H2D
 *   device : reconstruction waves, loop-filter waves, optional fancy upsampling; results stay in HBM
 *   fetch  : D2H of planes and/or NRGBA (any pointer may be NULL) */
int wgpu_dec_parse(wgpu_ctx* ctx, const uint8_t* const* streams, const size_t* lens, int n, int* width, int* height);
int wgpu_dec_device(wgpu_ctx* ctx, int want_nrgba);
int wgpu_dec_fetch(wgpu_ctx* ctx, uint8_t* y, uint8_t* u, uint8_t* v, size_t y_plane_stride, size_t uv_plane_stride,
                   uint8_t* nrgba, size_t nrgba_image_stride);

/* Pre-parsed entry: the Go decoder keeps parseHeaders / parseIntraModeRow / decodeMB (internal/lossy/decode.go:245-438,
 * decode_tree.go:35, decode_mb.go:111-313) and hands over what reconstructRow + doFilter + buildNRGBA consume, for whole frames instead
 * of one macroblock row.  wgpu_mb_data is MBData (decode.go:122-131: Coeffs dequantised with the WHT applied, the two 2-bit-per-block
 * transform codes NonZeroY / NonZeroUV, IsI4x4, IModes, UVMode, Skip, Segment) followed by that macroblock's FInfo (decode.go:107-112) as
 * precomputeFilterStrengths left it.  filter_type[i] = dec.filterType of image i (0 none, 1 simple, 2 complex, decode.go:399).
 * Runs the reconstruction, loop-filter and (want_nrgba) upsampling kernels; wgpu_dec_fetch then returns planes / NRGBA. */
typedef struct wgpu_mb_data {
  int16_t coeffs[384];
  uint32_t non_zero_y, non_zero_uv;
  uint8_t imodes[16];
  uint8_t is_i4x4, uv_mode, skip, segment;
  uint8_t f_limit, f_ilevel, f_inner, hev_thresh;
} wgpu_mb_data;
int wgpu_dec_reconstruct(wgpu_ctx* ctx, int n, int width, int height, const wgpu_mb_data* mbs /* [n][mb_h][mb_w] */,
                         const uint8_t* filter_type /* [n] */, int want_nrgba);

/* cleanupTransparentAreaLossy (encode.go:788-890) on n NRGBA images (non-premultiplied, 4 bytes per pixel): colours under
 * fully transparent pixels are smoothed / flattened per 8x8 block before the lossy encode of an image with alpha. */
int wgpu_cleanup_transparent(wgpu_ctx* ctx, const uint8_t* nrgba, int n, int width, int height, int stride, size_t image_stride,
                             uint8_t* out /* [n][height][4 * width] */);
/* wgpu_import_rgba: has_alpha bit 0 = alpha-weighted chroma, bit 1 = SharpYUV planes (sharpyuv/sharpyuv.go:40 Convert +
 * importYCbCr internal/lossy/encode.go:544), bits 8..16 = dithering amplitude (0 = fixed rounding). */
/* ---- stage-level entry points (host buffers in/out) ------------------------------------ */
int wgpu_import_rgba(wgpu_ctx* ctx, const uint8_t* rgba, int n, int width, int height, int stride,
                     size_t image_stride, int has_alpha, uint8_t* y, uint8_t* u, uint8_t* v);
int wgpu_upsample_nrgba(wgpu_ctx* ctx, int n, int width, int height, const uint8_t* y, int y_stride,
                        const uint8_t* u, const uint8_t* v, int uv_stride, size_t y_plane_stride,
                        size_t uv_plane_stride, const uint8_t* alpha, uint8_t* nrgba);
/* per plane pair: sse[i] (u64) and ssim_sum[i] (f64, sum over all pixels of SSIMGet/SSIMGetClipped) */
int wgpu_plane_metrics(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int width, int height, int stride,
                       size_t plane_stride, uint64_t* sse, double* ssim_sum);
double wgpu_psnr_from_sse(uint64_t sse, uint64_t count);

/* ---- per-block dsp operator surface, batched (n blocks, dense 4x4 tiles of 16 bytes) ---- */
int wgpu_dsp_ftransform_batch(wgpu_ctx* ctx, int n, const uint8_t* src, const uint8_t* ref, int16_t* out);
int wgpu_dsp_itransform_batch(wgpu_ctx* ctx, int n, const uint8_t* ref, const int16_t* in, uint8_t* dst);
int wgpu_dsp_fwht_batch(wgpu_ctx* ctx, int n, const int16_t* in, int16_t* out);
int wgpu_dsp_iwht_batch(wgpu_ctx* ctx, int n, const int16_t* in, int16_t* out);
int wgpu_dsp_sse4x4_batch(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int32_t* out);
int wgpu_dsp_tdisto4x4_batch(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int32_t* out);
/* ctx13[i] = {tl, t0..t7, l0..l3}; out[(i*10+mode)*16..] for the ten B_* modes */
int wgpu_dsp_pred4_batch(wgpu_ctx* ctx, int n, const uint8_t* ctx13, uint8_t* out);
/* PredLuma16Direct / PredChroma8Direct (predict_lossy.go:27-181): size 16 or 8; ctx[i] = {tl, top[size], left[size]};
 * out[((i*7+mode)*size + row)*size + col] for modes DC, TM, VE, HE, DC-noTop, DC-noLeft, DC-noTopLeft */
int wgpu_dsp_pred_square_batch(wgpu_ctx* ctx, int n, int size, const uint8_t* ctx_px, uint8_t* out);
/* quantizer derived as initSegmentQuant(dc_q, ac_q, type) (+ Y1 sharpening when sharpen!=0) */
int wgpu_dsp_quantize_batch(wgpu_ctx* ctx, int n, const int16_t* in, int dc_q, int ac_q, int type, int sharpen,
                            int first, int16_t* out, int32_t* nz);
int wgpu_dsp_trellis_batch(wgpu_ctx* ctx, int n, const int16_t* in, int dc_q, int ac_q, int qtype, int sharpen,
                           int first, int ctx_type, const int32_t* ctx0, int lambda, int16_t* out, int32_t* nz);
int wgpu_dsp_token_cost_batch(wgpu_ctx* ctx, int n, const int16_t* levels, const int32_t* nz, int ctx_type,
                              const int32_t* ctx0, int first, int32_t* out);

/* The rest of the dsp function surface (internal/dsp/dsp.go:12-37 and the *Direct / exported functions), batched:
 *   sse16x16 (ssim.go:220), tDisto16x16Go (:327): a, b [n][256] (16x16, stride 16) -> out [n]
 *   DequantCoeffs / dequantCoeffsGo (internal/lossy/encode_quant.go:81): in [n][16] -> out [n][16]
 *   FTransform2 (dsp.go:14): n pairs of adjacent blocks, src / ref [n][2][16] -> out [n][2][16]
 *   decoder transforms (transforms.go:37-216): kind 0 transformOne, 1 transformDC, 2 transformAC3 on 4x4 blocks
 *     (in [n][16], ref / dst [n][16]); kind 3 transformUV, 4 transformDCUV on 8x8 tiles (in [n][4][16], ref / dst [n][64], stride 8)
 *   loop-filter set (filter.go:93-242) on 24x24 tiles holding the block at (4, 4): kind 0 SimpleVFilter16, 1 SimpleHFilter16,
 *     2 SimpleVFilter16i, 3 SimpleHFilter16i, 4 VFilter16, 5 HFilter16, 6 VFilter16i, 7 HFilter16i, 8 VFilter8, 9 HFilter8,
 *     10 VFilter8i, 11 HFilter8i (8..11: one chroma plane per tile); tiles_in / tiles_out [n][24][24]
 *   UpsampleLinePair (upsample.go:45, channels = 3) / UpsampleLinePairNRGBA (:130, channels = 4) on n line pairs:
 *     top_y / bot_y [n][width] (bot_y NULL = last row of an odd height), *_u / *_v [n][(width + 1) / 2], alpha rows optional,
 *     top_dst / bot_dst [n][width][channels] */
/* VP8BitWriter (internal/bitio/writer_bool.go:58-150: PutBit per token, Flush, Finish) over n flat token arrays, the way the
 * encoder codes its token partitions on the device (chunk-parallel).  tokens = the arrays back to back, bit | prob << 8 per token
 * (encode_token.go:20 Token{Bit, Prob}), totals[i] tokens each; partition i goes to out + i * out_stride, sizes[i] bytes
 * (out_stride >= totals[i] / 8 + 4).  *rounds (may be NULL) receives how many state-relaxation rounds the batch needed. */
int wgpu_dsp_boolcode_batch(wgpu_ctx* ctx, int n, const uint16_t* tokens, const unsigned long long* totals, uint8_t* out,
                            size_t out_stride, unsigned int* sizes, int* rounds);
int wgpu_dsp_sse16x16_batch(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int32_t* out);
int wgpu_dsp_tdisto16x16_batch(wgpu_ctx* ctx, int n, const uint8_t* a, const uint8_t* b, int32_t* out);
int wgpu_dsp_dequant_batch(wgpu_ctx* ctx, int n, const int16_t* in, int dc_q, int ac_q, int16_t* out);
int wgpu_dsp_ftransform2_batch(wgpu_ctx* ctx, int n, const uint8_t* src, const uint8_t* ref, int16_t* out);
int wgpu_dsp_dec_transform_batch(wgpu_ctx* ctx, int n, int kind, const int16_t* in, const uint8_t* ref, uint8_t* dst);
int wgpu_dsp_filter_batch(wgpu_ctx* ctx, int n, int kind, const uint8_t* tiles_in, int thresh, int ithresh, int hev_thresh,
                          uint8_t* tiles_out);
int wgpu_dsp_upsample_line_pair_batch(wgpu_ctx* ctx, int n, int width, const uint8_t* top_y, const uint8_t* bot_y,
                                      const uint8_t* top_u, const uint8_t* top_v, const uint8_t* bot_u, const uint8_t* bot_v,
                                      const uint8_t* alpha_top, const uint8_t* alpha_bot, int channels, uint8_t* top_dst,
                                      uint8_t* bot_dst);

/* Token statistics for a host that keeps the serial path's probability refresh to itself (refreshProbas -> collectAllStats,
 * encode_frame.go:113, encode_proba.go:171): after wgpu_enc_device / wgpu_enc_search, the counts over the per-macroblock array of
 * every image as it stands with macroblocks of raster index >= upto_mb still in their zero state (not yet encoded: I16, not skipped,
 * no coefficients) -- upto_mb = mb_w * mb_h is collectAllStats over the finished frame.  stats [n][4][8][3][11][2] uint32. */
int wgpu_enc_stats(wgpu_ctx* ctx, int upto_mb, uint32_t* stats);

/* ---- measurement helpers (device timing on the library's own stream) -------------------- */
int wgpu_timer_begin(wgpu_ctx* ctx);           /* records a CUDA event on the ctx stream */
int wgpu_timer_end(wgpu_ctx* ctx, float* ms);   /* records + synchronises, returns elapsed ms */
/* Number of kernels this library launched on ctx since creation (for bench.py gpu_launches). */
uint64_t wgpu_launch_count(const wgpu_ctx* ctx);
/* Bytes this context has copied host->device and device->host (every cudaMemcpy*Async it issued) since the last call with
 * reset != 0.  bench.py reports the per-step figures of the e2e leg from it. */
int wgpu_transfer_bytes(wgpu_ctx* ctx, uint64_t* h2d, uint64_t* d2h, int reset);
/* Timed device-only repetitions of one stage over data already uploaded by wgpu_enc_upload /
 * decoded data; used for roofline numbers.  stage: 0 import, 1 analysis, 2 mode search (all waves), 3 SSE + SSIM of the
 * source luma planes against the reconstruction (dsp/ssim.go:12-181; 5 = the SSE alone), 4 fancy upsampling of the batch last decoded on this
 * context (wgpu_dec_parse + wgpu_dec_device) to NRGBA. */
int wgpu_enc_stage_time(wgpu_ctx* ctx, const wgpu_enc_options* opt, int stage, int reps, float* ms_per_rep);

#ifdef __cplusplus
}
#endif
#endif /* WEBPGPU_H_ */
