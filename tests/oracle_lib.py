"""ctypes bindings to the CPU oracle (oracle/_build/liboracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class OrcEncCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "quality", "method", "sns_strength", "filter_strength", "filter_sharpness", "filter_type",
        "partitions", "segments", "preprocessing", "has_alpha", "passes", "dither_amp", "target_size")] + [
        ("target_psnr", C.c_float), ("qmin", C.c_int), ("qmax", C.c_int), ("use_sharp_yuv", C.c_int)]


def default_cfg(quality=75, method=4, **kw):
    """lossy.DefaultConfig (internal/lossy/encode.go:66) + EncoderOptions mapping (encode.go:478-528)."""
    c = OrcEncCfg(quality=quality, method=method, sns_strength=50, filter_strength=60, filter_sharpness=0,
                  filter_type=1, partitions=0, segments=4, preprocessing=0, has_alpha=0, passes=1, dither_amp=0,
                  target_size=0, target_psnr=0.0, qmin=0, qmax=100, use_sharp_yuv=0)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_encode.restype = C.c_long
        _LIB.orc_encode_batch.restype = C.c_long
        _LIB.orc_plane_sse.restype = C.c_uint64
        _LIB.orc_plane_ssim.restype = C.c_double
        _LIB.orc_psnr_from_sse.restype = C.c_double
        _LIB.orc_psnr_from_sse.argtypes = [C.c_uint64, C.c_uint64]
        _LIB.orc_level_fixed_costs.restype = C.POINTER(C.c_uint16)
        _LIB.orc_entropy_cost.restype = C.POINTER(C.c_uint16)
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def encode(rgba, cfg=None, taps=False):
    """rgba: uint8 [h][w][4].  Returns bytes (RIFF) or (bytes, dict of taps)."""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    cfg = cfg or default_cfg()
    mbw, mbh = (w + 15) >> 4, (h + 15) >> 4
    nmb = mbw * mbh
    out = np.zeros(w * h * 2 + 65536, np.uint8)
    t = None
    if taps:
        t = dict(
            mb_hdr=np.zeros((nmb, 8), np.uint8), mb_modes=np.zeros((nmb, 16), np.uint8),
            mb_nz=np.zeros((nmb, 24), np.uint8), mb_coeffs=np.zeros((nmb, 400), np.int16),
            recon_y=np.zeros((mbh * 16, mbw * 16), np.uint8), recon_u=np.zeros((mbh * 8, mbw * 8), np.uint8),
            recon_v=np.zeros((mbh * 8, mbw * 8), np.uint8), src_y=np.zeros((mbh * 16, mbw * 16), np.uint8),
            src_u=np.zeros((mbh * 8, mbw * 8), np.uint8), src_v=np.zeros((mbh * 8, mbw * 8), np.uint8),
            alphas=np.zeros(nmb, np.uint8), seg=np.zeros((4, 8), np.int32))
    args = [_p(t[k]) if t else None for k in ("mb_hdr", "mb_modes", "mb_nz", "mb_coeffs", "recon_y", "recon_u",
                                               "recon_v", "src_y", "src_u", "src_v", "alphas", "seg")]
    n = lib().orc_encode(_p(rgba), C.c_int(rgba.strides[0]), w, h, C.byref(cfg), _p(out), C.c_long(out.size), *args)
    if n < 0:
        raise RuntimeError("orc_encode failed: %d" % n)
    data = out[:n].tobytes()
    return (data, t) if taps else data


_OPS_LIB = None


def encode_ops(rgba, cfg=None):
    """Integer operations of one encode, counted by the oracle built with its per-stage counters (oracle/vp8_common.h OpStage,
    liboracle_ops.so): {stage name: units x operations per unit}.  SURVEY.md 8(d): the numerator of the integer-issue roofline."""
    global _OPS_LIB
    if _OPS_LIB is None:
        path = os.path.join(ROOT, "oracle", "_build", "liboracle_ops.so")
        if not os.path.exists(path):
            build()
        _OPS_LIB = C.CDLL(path)
        _OPS_LIB.orc_op_name.restype = C.c_char_p
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    cfg = cfg or default_cfg()
    cnt = (C.c_ulonglong * 64)()
    wt = (C.c_uint * 64)()
    n = _OPS_LIB.orc_encode_ops(_p(rgba), C.c_int(rgba.strides[0]), w, h, C.byref(cfg), cnt, wt)
    if n < 0:
        raise RuntimeError("orc_encode_ops failed: %d" % n)
    return {_OPS_LIB.orc_op_name(i).decode(): int(cnt[i]) * int(wt[i]) for i in range(n)}


def encode_tokens(rgba, cfg=None):
    """(tokens uint16 [n] = bit | prob << 8, coded token partition bytes) of one encode: the input and the expected output of the
    device boolean coder (oracle/capi.cc orc_encode_tokens)."""
    L = lib()
    L.orc_encode_tokens.restype = C.c_long
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    cfg = cfg or default_cfg()
    plen = C.c_long(0)
    cap = w * h * 4 + (1 << 16)
    toks = np.empty(cap, np.uint16)
    part = np.empty(cap, np.uint8)
    n = L.orc_encode_tokens(_p(rgba), C.c_int(rgba.strides[0]), w, h, C.byref(cfg), toks.ctypes.data_as(C.c_void_p), C.c_long(cap),
                            part.ctypes.data_as(C.c_void_p), C.c_long(cap), C.byref(plen))
    if n < 0 or n > cap or plen.value > cap:
        raise RuntimeError("orc_encode_tokens failed: %d" % n)
    return toks[:n].copy(), part[:plen.value].copy()


def boolcode(tokens):
    """VP8BitWriter over a flat token array (uint16 bit | prob << 8): the coded bytes (oracle/capi.cc orc_boolcode)."""
    L = lib()
    L.orc_boolcode.restype = C.c_long
    t = np.ascontiguousarray(tokens, dtype=np.uint16)
    out = np.empty(len(t) + 64, np.uint8)
    m = L.orc_boolcode(t.ctypes.data_as(C.c_void_p), C.c_ulonglong(len(t)), out.ctypes.data_as(C.c_void_p), C.c_long(len(out)))
    if m < 0:
        raise RuntimeError("orc_boolcode failed")
    return out[:m].copy()


def adversarial_tokens(rng, n, mode):
    """Token streams that stress the boolean coder beyond what an encoder emits: 0 bits follow the probabilities, 1 improbable bits
    and extreme probabilities (carries, big shifts, probability 0), 2 long runs of near-certain zeros (few shifts), 3 all ones against
    tiny probabilities (0xff runs), 4 near-certain zeros only (range states merge slowly: many relaxation rounds), 5 uniform."""
    if mode == 0:
        p = rng.integers(1, 256, n); b = (rng.random(n) * 256 >= p)
    elif mode == 1:
        p = rng.choice([1, 2, 254, 255, 128, 0], n); b = rng.integers(0, 2, n)
    elif mode == 2:
        p = np.full(n, 255); b = rng.random(n) < 0.002
    elif mode == 3:
        p = rng.integers(1, 4, n); b = np.ones(n)
    elif mode == 4:
        p = rng.integers(250, 256, n); b = np.zeros(n)
    else:
        p = rng.integers(0, 256, n); b = rng.integers(0, 2, n)
    return (np.asarray(b).astype(np.uint16) | (np.asarray(p).astype(np.uint16) << 8)).astype(np.uint16)


def encode_batch(rgba_batch, cfg=None, threads=1):
    """rgba_batch uint8 [n][h][w][4]; returns total compressed bytes (timing leg)."""
    rgba_batch = np.ascontiguousarray(rgba_batch, dtype=np.uint8)
    n, h, w = rgba_batch.shape[:3]
    cfg = cfg or default_cfg()
    sizes = np.zeros(n, np.int64)
    r = lib().orc_encode_batch(_p(rgba_batch), n, w * 4, w, h, C.byref(cfg), threads, _p(sizes))
    if r < 0:
        raise RuntimeError("orc_encode_batch failed")
    return r, sizes


def decode_info(data):
    w, h, mbw, mbh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = lib().orc_decode_info(data, C.c_long(len(data)), C.byref(w), C.byref(h), C.byref(mbw), C.byref(mbh))
    if rc:
        raise RuntimeError("orc_decode_info: %d" % rc)
    return w.value, h.value, mbw.value, mbh.value


def decode(data, filter=True, taps=False, libwebp_inner_rule=False):
    """Returns (w, h, Y, U, V) padded planes (and per-MB parsed data if taps)."""
    w, h, mbw, mbh = decode_info(data)
    y = np.zeros((mbh * 16, mbw * 16), np.uint8)
    u = np.zeros((mbh * 8, mbw * 8), np.uint8)
    v = np.zeros((mbh * 8, mbw * 8), np.uint8)
    t = None
    nmb = mbw * mbh
    if taps:
        t = dict(coeffs=np.zeros((nmb, 384), np.int16), meta=np.zeros((nmb, 24), np.uint8),
                 nz=np.zeros((nmb, 2), np.uint32), hdr=np.zeros(4, np.int32))
    rc = lib().orc_decode(data, C.c_long(len(data)), (int(bool(filter)) | (2 if libwebp_inner_rule else 0)), _p(y), _p(u), _p(v),
                          _p(t["coeffs"]) if t else None, _p(t["meta"]) if t else None,
                          _p(t["nz"]) if t else None, _p(t["hdr"]) if t else None)
    if rc:
        raise RuntimeError("orc_decode: %d" % rc)
    return (w, h, y, u, v, t) if taps else (w, h, y, u, v)


def import_rgba(rgba, has_alpha=False, dither_amp=0):
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    mbw, mbh = (w + 15) >> 4, (h + 15) >> 4
    y = np.zeros((mbh * 16, mbw * 16), np.uint8)
    u = np.zeros((mbh * 8, mbw * 8), np.uint8)
    v = np.zeros((mbh * 8, mbw * 8), np.uint8)
    lib().orc_import_rgba(_p(rgba), C.c_int(rgba.strides[0]), w, h, int(has_alpha) | (int(dither_amp) << 8), _p(y), _p(u), _p(v))
    return y, u, v


def cleanup_transparent(nrgba):
    """cleanupTransparentAreaLossy (encode.go:788) on a copy of an NRGBA image [h][w][4]."""
    out = np.ascontiguousarray(nrgba, dtype=np.uint8).copy()
    h, w = out.shape[:2]
    lib().orc_cleanup_transparent(_p(out), C.c_int(out.strides[0]), w, h)
    return out


def build_nrgba(w, h, y, u, v, alpha=None):
    out = np.zeros((h, w, 4), np.uint8)
    lib().orc_build_nrgba(w, h, _p(y), C.c_int(y.strides[0]), _p(u), _p(v), C.c_int(u.strides[0]),
                          _p(alpha) if alpha is not None else None, _p(out))
    return out


def plane_sse(a, b):
    h, w = a.shape
    return int(lib().orc_plane_sse(_p(a), C.c_int(a.strides[0]), _p(b), C.c_int(b.strides[0]), w, h))


def plane_ssim(a, b):
    h, w = a.shape
    return float(lib().orc_plane_ssim(_p(a), C.c_int(a.strides[0]), _p(b), C.c_int(b.strides[0]), w, h))


def plane_ssim_map(a, b):
    h, w = a.shape
    out = np.zeros((h, w), np.float64)
    lib().orc_plane_ssim_map(_p(a), C.c_int(a.strides[0]), _p(b), C.c_int(b.strides[0]), w, h, _p(out))
    return out


# ------------------------------------------------------------------ synthetic images
import sys as _sys
if ROOT not in _sys.path:
    _sys.path.insert(0, ROOT)
from webp_b200.synth import synth_image  # noqa: E402,F401  (shared generator; pure numpy, no native code)


def sharp_yuv(rgba):
    """sharpyuv.Convert (WebP matrix, sRGB transfer) on RGBA [h][w][4]: tight Y, U, V planes and the refinement passes run."""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    y = np.zeros((h, w), np.uint8)
    u = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
    v = np.zeros_like(u)
    iters = lib().orc_sharp_yuv(_p(rgba), C.c_int(rgba.strides[0]), w, h, _p(y), _p(u), _p(v))
    return y, u, v, iters
