"""Batched twins of the reference's per-block operator surface (internal/dsp/dsp.go:12-37 and the *Direct
functions), plus the plane-level stages.  Blocks are dense 4x4 tiles (16 values, raster order).
"""
import ctypes as C

import numpy as np

from . import native


def _ctx(ctx):
    return ctx or native.default_context()


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def FTransformBatch(src, ref, ctx=None):
    """dsp.FTransformDirect (internal/dsp/transforms.go:371) on n blocks: uint8 [n][16] x2 -> int16 [n][16]."""
    ctx = _ctx(ctx); src = _c(src, np.uint8); ref = _c(ref, np.uint8)
    n = src.shape[0]; out = np.empty((n, 16), np.int16)
    ctx.check(native.lib().wgpu_dsp_ftransform_batch(ctx.handle, n, src.ctypes.data, ref.ctypes.data, out.ctypes.data))
    return out


def ITransformBatch(ref, coeffs, ctx=None):
    """dsp.ITransformDirect (transforms.go:265): ref uint8 [n][16], coeffs int16 [n][16] -> uint8 [n][16]."""
    ctx = _ctx(ctx); ref = _c(ref, np.uint8); coeffs = _c(coeffs, np.int16)
    n = ref.shape[0]; out = np.empty((n, 16), np.uint8)
    ctx.check(native.lib().wgpu_dsp_itransform_batch(ctx.handle, n, ref.ctypes.data, coeffs.ctypes.data, out.ctypes.data))
    return out


def FTransformWHTBatch(dcs, ctx=None):
    """dsp.FTransformWHT (transforms.go:500)."""
    ctx = _ctx(ctx); dcs = _c(dcs, np.int16); n = dcs.shape[0]; out = np.empty((n, 16), np.int16)
    ctx.check(native.lib().wgpu_dsp_fwht_batch(ctx.handle, n, dcs.ctypes.data, out.ctypes.data))
    return out


def TransformWHTBatch(coeffs, ctx=None):
    """dsp.TransformWHT (transforms.go:223); out[b] = DC of block b."""
    ctx = _ctx(ctx); coeffs = _c(coeffs, np.int16); n = coeffs.shape[0]; out = np.empty((n, 16), np.int16)
    ctx.check(native.lib().wgpu_dsp_iwht_batch(ctx.handle, n, coeffs.ctypes.data, out.ctypes.data))
    return out


def SSE4x4Batch(a, b, ctx=None):
    """dsp.SSE4x4Direct (ssim.go:188)."""
    ctx = _ctx(ctx); a = _c(a, np.uint8); b = _c(b, np.uint8); n = a.shape[0]; out = np.empty(n, np.int32)
    ctx.check(native.lib().wgpu_dsp_sse4x4_batch(ctx.handle, n, a.ctypes.data, b.ctypes.data, out.ctypes.data))
    return out


def TDisto4x4Batch(a, b, ctx=None):
    """dsp.TDisto4x4 (ssim.go:315)."""
    ctx = _ctx(ctx); a = _c(a, np.uint8); b = _c(b, np.uint8); n = a.shape[0]; out = np.empty(n, np.int32)
    ctx.check(native.lib().wgpu_dsp_tdisto4x4_batch(ctx.handle, n, a.ctypes.data, b.ctypes.data, out.ctypes.data))
    return out


def PredLuma4Batch(ctx13, ctx=None):
    """dsp.PredLuma4Direct (predict_lossy.go:185-451) for all ten modes: ctx13 uint8 [n][13] = {tl, t0..t7, l0..l3}
    -> uint8 [n][10][16]."""
    ctx = _ctx(ctx); c = _c(ctx13, np.uint8); n = c.shape[0]; out = np.empty((n, 10, 16), np.uint8)
    ctx.check(native.lib().wgpu_dsp_pred4_batch(ctx.handle, n, c.ctypes.data, out.ctypes.data))
    return out


def PredSquareBatch(ctx_px, size, ctx=None):
    """dsp.PredLuma16Direct (size 16) / PredChroma8Direct (size 8) (predict_lossy.go:27-181) for the seven modes:
    ctx_px uint8 [n][1 + 2*size] = {tl, top[size], left[size]} -> uint8 [n][7][size][size]."""
    ctx = _ctx(ctx); c = _c(ctx_px, np.uint8); n = c.shape[0]; out = np.empty((n, 7, size, size), np.uint8)
    ctx.check(native.lib().wgpu_dsp_pred_square_batch(ctx.handle, n, size, c.ctypes.data, out.ctypes.data))
    return out


def QuantizeCoeffsBatch(coeffs, dc_q, ac_q, qtype, sharpen, first, ctx=None):
    """lossy.QuantizeCoeffs (internal/lossy/encode_quant.go:16) -> (levels int16 [n][16], nz int32 [n])."""
    ctx = _ctx(ctx); c = _c(coeffs, np.int16); n = c.shape[0]
    out = np.empty((n, 16), np.int16); nz = np.empty(n, np.int32)
    ctx.check(native.lib().wgpu_dsp_quantize_batch(ctx.handle, n, c.ctypes.data, dc_q, ac_q, qtype, int(sharpen), first,
                                                  out.ctypes.data, nz.ctypes.data))
    return out, nz


def TrellisQuantizeBlockBatch(coeffs, dc_q, ac_q, qtype, sharpen, first, ctx_type, ctx0, lam, ctx=None):
    """lossy.TrellisQuantizeBlock (internal/lossy/encode_trellis.go:23) with the default coefficient probabilities."""
    ctx = _ctx(ctx); c = _c(coeffs, np.int16); c0 = _c(ctx0, np.int32); n = c.shape[0]
    out = np.empty((n, 16), np.int16); nz = np.empty(n, np.int32)
    ctx.check(native.lib().wgpu_dsp_trellis_batch(ctx.handle, n, c.ctypes.data, dc_q, ac_q, qtype, int(sharpen), first, ctx_type,
                                                 c0.ctypes.data, lam, out.ctypes.data, nz.ctypes.data))
    return out, nz


def TokenCostForCoeffsBatch(levels, nz, ctx_type, ctx0, first, ctx=None):
    """lossy.TokenCostForCoeffs (internal/lossy/encode_quant.go:170) with the default coefficient probabilities."""
    ctx = _ctx(ctx); lv = _c(levels, np.int16); z = _c(nz, np.int32); c0 = _c(ctx0, np.int32); n = lv.shape[0]
    out = np.empty(n, np.int32)
    ctx.check(native.lib().wgpu_dsp_token_cost_batch(ctx.handle, n, lv.ctypes.data, z.ctypes.data, ctx_type, c0.ctypes.data, first,
                                                    out.ctypes.data))
    return out


def ImportRGBA(rgba, has_alpha=False, ctx=None, dither_amp=0, sharp_yuv=False):
    """(*VP8Encoder).importImage (internal/lossy/encode.go:671): uint8 [n][h][w][4] -> padded Y, U, V planes.  sharp_yuv: the planes
    NewEncoderFromYUV gets under UseSharpYUV instead (sharpyuv.Convert, sharpyuv/sharpyuv.go:40, + importYCbCr, encode.go:544)."""
    ctx = _ctx(ctx); a = _c(rgba, np.uint8); n, h, w = a.shape[:3]
    mbw, mbh = (w + 15) >> 4, (h + 15) >> 4
    y = np.empty((n, mbh * 16, mbw * 16), np.uint8); u = np.empty((n, mbh * 8, mbw * 8), np.uint8); v = np.empty_like(u)
    ctx.check(native.lib().wgpu_import_rgba(ctx.handle, a.ctypes.data, n, w, h, w * 4, w * h * 4, int(bool(has_alpha)) | (2 if sharp_yuv else 0) | (int(dither_amp) << 8), y.ctypes.data,
                                           u.ctypes.data, v.ctypes.data))
    return y, u, v


def CleanupTransparentArea(nrgba, ctx=None):
    """cleanupTransparentAreaLossy (encode.go:788): uint8 NRGBA [n][h][w][4] -> same shape, colours under alpha == 0 smoothed per 8x8 block."""
    ctx = _ctx(ctx); a = _c(nrgba, np.uint8); n, h, w = a.shape[:3]
    out = np.empty_like(a)
    ctx.check(native.lib().wgpu_cleanup_transparent(ctx.handle, a.ctypes.data, n, w, h, w * 4, w * h * 4, out.ctypes.data))
    return out


def UpsampleNRGBA(y, u, v, width, height, alpha=None, ctx=None):
    """buildNRGBA (webp.go:379) / dsp.UpsampleLinePairNRGBA (internal/dsp/upsample.go:130): planes [n][H][S] -> NRGBA."""
    ctx = _ctx(ctx); y = _c(y, np.uint8); u = _c(u, np.uint8); v = _c(v, np.uint8)
    n = y.shape[0]; out = np.empty((n, height, width, 4), np.uint8)
    al = _c(alpha, np.uint8) if alpha is not None else None
    ctx.check(native.lib().wgpu_upsample_nrgba(ctx.handle, n, width, height, y.ctypes.data, y.shape[2], u.ctypes.data, v.ctypes.data,
                                              u.shape[2], y[0].nbytes, u[0].nbytes, al.ctypes.data if al is not None else None,
                                              out.ctypes.data))
    return out


def PlaneMetrics(a, b, ctx=None):
    """dsp.SSE (ssim.go:172) and the sum over all pixels of SSIMGet / SSIMGetClipped (ssim.go:116,132): planes [n][h][w]."""
    ctx = _ctx(ctx); a = _c(a, np.uint8); b = _c(b, np.uint8); n, h, w = a.shape
    sse = np.zeros(n, np.uint64); ssim = np.zeros(n, np.float64)
    ctx.check(native.lib().wgpu_plane_metrics(ctx.handle, n, a.ctypes.data, b.ctypes.data, w, h, w, a[0].nbytes, sse.ctypes.data,
                                             ssim.ctypes.data))
    return sse, ssim


def PSNRFromSSE(sse, count):
    """dsp.PSNRFromSSE (ssim.go:163)."""
    return float(native.lib().wgpu_psnr_from_sse(int(sse), int(count)))


def SSE16x16Batch(a, b, ctx=None):
    """dsp.SSE16x16Direct (ssim.go:220): uint8 [n][16][16] x2 -> int32 [n]."""
    ctx = _ctx(ctx); a = _c(a, np.uint8); b = _c(b, np.uint8); n = a.shape[0]; out = np.empty(n, np.int32)
    ctx.check(native.lib().wgpu_dsp_sse16x16_batch(ctx.handle, n, a.ctypes.data, b.ctypes.data, out.ctypes.data))
    return out


def TDisto16x16Batch(a, b, ctx=None):
    """dsp.TDisto16x16 (ssim.go:327)."""
    ctx = _ctx(ctx); a = _c(a, np.uint8); b = _c(b, np.uint8); n = a.shape[0]; out = np.empty(n, np.int32)
    ctx.check(native.lib().wgpu_dsp_tdisto16x16_batch(ctx.handle, n, a.ctypes.data, b.ctypes.data, out.ctypes.data))
    return out


def DequantCoeffsBatch(levels, dc_q, ac_q, ctx=None):
    """lossy.DequantCoeffs (internal/lossy/encode_quant.go:81): int16 [n][16] -> int16 [n][16] (int16 truncation as the reference)."""
    ctx = _ctx(ctx); levels = _c(levels, np.int16); n = levels.shape[0]; out = np.empty((n, 16), np.int16)
    ctx.check(native.lib().wgpu_dsp_dequant_batch(ctx.handle, n, levels.ctypes.data, int(dc_q), int(ac_q), out.ctypes.data))
    return out


def FTransform2Batch(src, ref, ctx=None):
    """dsp.FTransform2 (dsp.go:14): two adjacent blocks per call: uint8 [n][2][16] x2 -> int16 [n][2][16]."""
    ctx = _ctx(ctx); src = _c(src, np.uint8); ref = _c(ref, np.uint8); n = src.shape[0]; out = np.empty((n, 2, 16), np.int16)
    ctx.check(native.lib().wgpu_dsp_ftransform2_batch(ctx.handle, n, src.ctypes.data, ref.ctypes.data, out.ctypes.data))
    return out


DEC_TRANSFORMS = {"Transform": 0, "TransformDC": 1, "TransformAC3": 2, "TransformUV": 3, "TransformDCUV": 4}


def DecTransformBatch(kind, coeffs, ref, ctx=None):
    """dsp.Transform / TransformDC / TransformAC3 (4x4 blocks: coeffs int16 [n][16], ref uint8 [n][16]) and TransformUV /
    TransformDCUV (8x8 tiles: coeffs [n][4][16], ref [n][8][8]) -- transforms.go:37-216.  Returns ref + inverse transform."""
    ctx = _ctx(ctx); coeffs = _c(coeffs, np.int16); ref = _c(ref, np.uint8); n = ref.shape[0]; out = np.empty_like(ref)
    ctx.check(native.lib().wgpu_dsp_dec_transform_batch(ctx.handle, n, DEC_TRANSFORMS[kind], coeffs.ctypes.data, ref.ctypes.data, out.ctypes.data))
    return out


FILTERS = ["SimpleVFilter16", "SimpleHFilter16", "SimpleVFilter16i", "SimpleHFilter16i", "VFilter16", "HFilter16", "VFilter16i", "HFilter16i",
           "VFilter8", "HFilter8", "VFilter8i", "HFilter8i"]


def FilterBatch(kind, tiles, thresh, ithresh=0, hev_thresh=0, ctx=None):
    """The loop-filter set of internal/dsp/filter.go:93-242 on uint8 tiles [n][24][24] holding the block at (4, 4)."""
    ctx = _ctx(ctx); tiles = _c(tiles, np.uint8); n = tiles.shape[0]; out = np.empty_like(tiles)
    ctx.check(native.lib().wgpu_dsp_filter_batch(ctx.handle, n, FILTERS.index(kind), tiles.ctypes.data, int(thresh), int(ithresh), int(hev_thresh),
                                                 out.ctypes.data))
    return out


def UpsampleLinePairBatch(top_y, bot_y, top_u, top_v, bot_u, bot_v, channels=3, alpha_top=None, alpha_bot=None, ctx=None):
    """dsp.UpsampleLinePair (upsample.go:45, channels=3) / UpsampleLinePairNRGBA (:130, channels=4) on n line pairs [n][width]."""
    ctx = _ctx(ctx); top_y = _c(top_y, np.uint8); n, width = top_y.shape
    arrs = [None if a is None else _c(a, np.uint8) for a in (bot_y, top_u, top_v, bot_u, bot_v, alpha_top, alpha_bot)]
    ptr = [None if a is None else a.ctypes.data for a in arrs]
    td = np.empty((n, width, channels), np.uint8); bd = np.empty((n, width, channels), np.uint8)
    ctx.check(native.lib().wgpu_dsp_upsample_line_pair_batch(ctx.handle, n, width, top_y.ctypes.data, ptr[0], ptr[1], ptr[2], ptr[3], ptr[4],
                                                             ptr[5], ptr[6], channels, td.ctypes.data, bd.ctypes.data))
    return td, (bd if bot_y is not None else None)


def BoolCodeBatch(token_arrays, ctx=None):
    """bitio VP8BitWriter (writer_bool.go:58-150) over n flat token arrays (uint16 bit | prob << 8, encode_token.go:20): PutBit per
    token + Finish, coded chunk-parallel on the device.  Returns (list of coded partitions, relaxation rounds used)."""
    ctx = _ctx(ctx)
    arrs = [_c(a, np.uint16) for a in token_arrays]
    n = len(arrs)
    totals = np.array([len(a) for a in arrs], np.uint64)
    flat = np.concatenate(arrs + [np.zeros(1, np.uint16)])
    stride = int(totals.max()) + 64
    out = np.zeros((n, stride), np.uint8); sizes = np.zeros(n, np.uint32)
    rounds = C.c_int(0)
    ctx.check(native.lib().wgpu_dsp_boolcode_batch(ctx.handle, n, flat.ctypes.data, totals.ctypes.data, out.ctypes.data, stride, sizes.ctypes.data,
                                                   C.byref(rounds)))
    return [out[i, :int(sizes[i])].copy() for i in range(n)], int(rounds.value)
