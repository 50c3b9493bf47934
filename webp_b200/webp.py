"""Host-side mirror of the reference's public API for the VP8 lossy path, over the C ABI.

Names, argument meaning and error behaviour follow the Go package (paths relative to the reference):
  Encode / EncoderOptions / DefaultOptions / OptionsForPreset / validateConfig     encode.go:42-334,424
  Decode / DecodeConfig / buildYCbCr / buildNRGBA                                  webp.go:88-140,351-450
Only the lossy pixel path is behind this boundary (BASELINE.json north_star); Lossless and alpha-plane options are
rejected with the same "webp: ..." error style instead of being emulated.  Metadata (ICC / EXIF / XMP) puts the VP8X
container of mux.py around the same bitstream; animations are animation.py.
EncodeBatch / DecodeBatch are the additions a GPU backend needs: one image cannot fill a B200.
"""
import ctypes as C
import io
import math
from dataclasses import dataclass, field

import numpy as np

from . import native

MaxDimension = 16383
PresetDefault, PresetPicture, PresetPhoto, PresetDrawing, PresetIcon, PresetText = range(6)


@dataclass
class EncoderOptions:
    """encode.go:42-187.  Negative values are the reference's "use the libwebp default" sentinels."""
    Lossless: bool = False
    Quality: float = 0.0
    Method: int = 0
    Preset: int = PresetDefault
    UseSharpYUV: bool = False
    Exact: bool = False
    TargetSize: int = 0
    TargetPSNR: float = 0.0
    Preprocessing: int = 0
    SNSStrength: int = 0
    FilterStrength: int = 0
    FilterSharpness: int = 0
    FilterType: int = 0
    Partitions: int = 0
    Segments: int = 0
    Pass: int = 0
    EmulateJpegSize: bool = False
    QMin: int = 0
    QMax: int = 0
    AlphaCompression: int = 0
    AlphaFiltering: int = 0
    AlphaQuality: int = 0
    ICC: bytes = b""
    EXIF: bytes = b""
    XMP: bytes = b""


Options = EncoderOptions


def DefaultOptions():
    """encode.go:196-214."""
    return EncoderOptions(Quality=75, Lossless=False, Method=4, SNSStrength=-1, FilterStrength=-1, FilterSharpness=0,
                          FilterType=-1, Partitions=0, Segments=-1, Pass=-1, QMin=0, QMax=-1, AlphaCompression=-1,
                          AlphaFiltering=-1, AlphaQuality=-1)


def OptionsForPreset(preset, quality):
    """encode.go:218-252."""
    o = DefaultOptions()
    o.Quality = quality
    o.Preset = preset
    if preset == PresetPicture:
        o.SNSStrength, o.FilterSharpness, o.FilterStrength = 80, 4, 35
        o.Preprocessing &= ~2
    elif preset == PresetPhoto:
        o.SNSStrength, o.FilterSharpness, o.FilterStrength = 80, 3, 30
        o.Preprocessing |= 2
    elif preset == PresetDrawing:
        o.SNSStrength, o.FilterSharpness, o.FilterStrength = 25, 6, 10
    elif preset == PresetIcon:
        o.SNSStrength, o.FilterStrength = 0, 0
        o.Preprocessing &= ~2
    elif preset == PresetText:
        o.SNSStrength, o.FilterStrength, o.Segments = 0, 0, 2
        o.Preprocessing &= ~2
    return o


def _resolve_qmax(v):
    return 100 if v < 0 else v


def validateConfig(o):
    """encode.go:259-334; returns an error string or None."""
    q = float(o.Quality)
    if q < 0 or q > 100 or math.isnan(q) or math.isinf(q):
        return "webp: invalid Quality %.2f (must be 0-100, finite)" % q
    if o.Method < 0 or o.Method > 6:
        return "webp: invalid Method %d (must be 0-6)" % o.Method
    if o.TargetSize < 0:
        return "webp: invalid TargetSize %d (must be >= 0)" % o.TargetSize
    if o.TargetPSNR < 0 or math.isnan(o.TargetPSNR) or math.isinf(o.TargetPSNR):
        return "webp: invalid TargetPSNR %.2f (must be >= 0, finite)" % o.TargetPSNR
    if o.Preprocessing < 0 or o.Preprocessing > 3:
        return "webp: invalid Preprocessing %d (must be 0-3)" % o.Preprocessing
    if o.Preset < PresetDefault or o.Preset > PresetText:
        return "webp: invalid Preset %d" % o.Preset
    if o.SNSStrength > 100:
        return "webp: invalid SNSStrength %d (must be 0-100 or negative sentinel)" % o.SNSStrength
    if o.FilterStrength > 100:
        return "webp: invalid FilterStrength %d (must be 0-100 or negative sentinel)" % o.FilterStrength
    if o.FilterSharpness < 0 or o.FilterSharpness > 7:
        return "webp: invalid FilterSharpness %d (must be 0-7)" % o.FilterSharpness
    if o.FilterType > 1:
        return "webp: invalid FilterType %d (must be 0 or 1, or negative sentinel)" % o.FilterType
    if o.Partitions < 0 or o.Partitions > 3:
        return "webp: invalid Partitions %d (must be 0-3)" % o.Partitions
    if o.Segments > 4:
        return "webp: invalid Segments %d (must be 1-4 or 0/-1 for default)" % o.Segments
    if o.Pass > 10:
        return "webp: invalid Pass %d (must be 1-10 or 0/-1 for default)" % o.Pass
    qmax = _resolve_qmax(o.QMax)
    if o.QMin < 0 or qmax > 100 or o.QMin > qmax:
        return "webp: invalid QMin/QMax %d/%d (must be 0-100, QMin <= QMax)" % (o.QMin, o.QMax)
    if o.AlphaCompression > 1:
        return "webp: invalid AlphaCompression %d (must be 0 or 1)" % o.AlphaCompression
    if o.AlphaFiltering > 2:
        return "webp: invalid AlphaFiltering %d (must be 0, 1 or 2)" % o.AlphaFiltering
    if o.AlphaQuality > 100:
        return "webp: invalid AlphaQuality %d (must be 0-100)" % o.AlphaQuality
    return None


def lossy_config(o, has_alpha=False):
    """EncoderOptions -> lossy.EncodeConfig (encode.go:478-528 over lossy.DefaultConfig, internal/lossy/encode.go:66-86)."""
    c = native.EncOptions(quality=int(o.Quality), method=o.Method, sns_strength=50, filter_strength=60, filter_sharpness=0,
                          filter_type=1, partitions=0, segments=4, preprocessing=0, has_alpha=int(has_alpha), passes=1, dither_amp=0,
                          target_size=0, target_psnr=0.0, qmin=0, qmax=100, use_sharp_yuv=int(bool(o.UseSharpYUV)))  # encode.go:531
    if o.SNSStrength >= 0:
        c.sns_strength = o.SNSStrength
    if o.FilterStrength >= 0:
        c.filter_strength = o.FilterStrength
    c.filter_sharpness = o.FilterSharpness
    if o.FilterType >= 0:
        c.filter_type = o.FilterType
    c.partitions = o.Partitions
    if o.Segments > 0:
        c.segments = o.Segments
    c.preprocessing = o.Preprocessing
    if o.Pass > 0:
        c.passes = o.Pass
    if o.TargetSize > 0:
        c.target_size = int(o.TargetSize)
    if o.TargetPSNR > 0:
        c.target_psnr = float(np.float32(o.TargetPSNR))
    c.qmin = int(o.QMin)
    c.qmax = 100 if o.QMax < 0 else int(o.QMax)  # resolveQMax (encode.go:305-306; -1 is the "unset" sentinel)
    if o.Preprocessing & 2:
        # encode.go:563-567 in float32: x = Quality/100; dithering = 1.0 + (0.5 - 1.0) * x^4; dsp.InitRandom: amp = int(256 * dithering)
        x = np.float32(o.Quality) / np.float32(100.0)
        x2 = np.float32(x * x)
        d = np.float32(1.0) + np.float32(np.float32(np.float32(-0.5) * x2) * x2)
        c.dither_amp = 0 if d < 0 else (256 if d > 1 else int(np.float32(256.0) * d))
    return c


class WebPError(ValueError):
    pass


def _unsupported(o):
    if o.Lossless:
        return "webp: Lossless (VP8L) is outside the GPU lossy path"
    return None


def _as_rgba_batch(imgs):
    a = np.asarray(imgs)
    if a.dtype != np.uint8 or a.ndim != 4 or a.shape[-1] not in (3, 4):
        raise WebPError("webp: images must be uint8 [n][h][w][3|4]")
    if a.shape[-1] == 3:
        a = np.concatenate([a, np.full(a.shape[:-1] + (1,), 255, np.uint8)], axis=-1)
    return np.ascontiguousarray(a)


def EncodeBatch(imgs, opts=None, ctx=None):
    """Encode n same-size RGB(A) images; returns a list of WebP files (bytes).  Batch twin of Encode."""
    if imgs is None:
        raise WebPError("webp: nil image")
    o = opts or DefaultOptions()
    err = validateConfig(o) or _unsupported(o)
    if err:
        raise WebPError(err)
    a = _as_rgba_batch(imgs)
    n, h, w = a.shape[:3]
    if w <= 0 or h <= 0:
        raise WebPError("webp: invalid image dimensions %dx%d" % (w, h))
    if w > MaxDimension or h > MaxDimension:
        raise WebPError("webp: image dimension %dx%d exceeds maximum %d" % (w, h, MaxDimension))
    if not bool((a[..., 3] == 255).all()):
        raise WebPError("webp: non-opaque alpha needs the ALPH (VP8L) chunk, which is outside the GPU lossy path")
    ctx = ctx or native.default_context()
    cfg = lossy_config(o, has_alpha=False)
    cap = w * h * 2 + 65536
    out = np.empty((n, cap), np.uint8)
    sizes = np.zeros(n, np.uint64)
    ctx.check(native.lib().wgpu_encode_batch(ctx.handle, a.ctypes.data, n, w, h, w * 4, w * h * 4, C.byref(cfg), out.ctypes.data,
                                            cap, sizes.ctypes.data))
    files = [out[i, :int(sizes[i])].tobytes() for i in range(n)]
    if o.ICC or o.EXIF or o.XMP:  # writeRIFF (encode.go:955): metadata asks for the VP8X container around the same bitstream
        from . import mux
        files = [mux.writeRIFFExtended(mux.FourCCVP8, mux.riff_payload(f), None, w, h, bytes(o.ICC), bytes(o.EXIF), bytes(o.XMP)) for f in files]
    return files


def Encode(w, img, opts=None, ctx=None):
    """webp.Encode (encode.go:424): writes one WebP file to the binary writer w."""
    if w is None:
        raise WebPError("webp: nil writer")
    if img is None:
        raise WebPError("webp: nil image")
    w.write(EncodeBatch(np.asarray(img)[None], opts, ctx)[0])


@dataclass
class Config:
    Width: int
    Height: int
    ColorModel: str = "YCbCr"


@dataclass
class YCbCr:
    """*image.YCbCr with 4:2:0 subsampling, as webp.Decode returns for lossy images without alpha (webp.go:351)."""
    Y: np.ndarray
    Cb: np.ndarray
    Cr: np.ndarray
    Rect: tuple = field(default=(0, 0, 0, 0))

    @property
    def YStride(self):
        return self.Y.shape[1]

    @property
    def CStride(self):
        return self.Cb.shape[1]


def _read_all(r):
    if r is None:
        raise WebPError("webp: nil reader")
    return r if isinstance(r, (bytes, bytearray, memoryview)) else r.read()


def DecodeConfig(r):
    """webp.DecodeConfig (webp.go:101)."""
    data = bytes(_read_all(r))
    w, h = C.c_int(), C.c_int()
    rc = native.lib().wgpu_decode_info(data, len(data), C.byref(w), C.byref(h))
    if rc == native.ERR_UNSUPPORTED:
        raise WebPError("webp: ALPH chunk / alpha or animation flag: outside the GPU lossy path")
    if rc != native.OK:
        raise WebPError("webp: invalid VP8 lossy stream")
    return Config(w.value, h.value)


def DecodeBatch(streams, nrgba=False, ctx=None):
    """Decode n same-size lossy WebP files.  Returns a list of YCbCr, or (list, NRGBA uint8 [n][h][w][4]) when
    nrgba=True (buildNRGBA semantics, A = 255)."""
    streams = [bytes(_read_all(s)) for s in streams]
    if not streams:
        return []
    cfg = DecodeConfig(streams[0])
    w, h = cfg.Width, cfg.Height
    mbw, mbh = (w + 15) >> 4, (h + 15) >> 4
    n = len(streams)
    ctx = ctx or native.default_context()
    y = np.empty((n, mbh * 16, mbw * 16), np.uint8)
    u = np.empty((n, mbh * 8, mbw * 8), np.uint8)
    v = np.empty((n, mbh * 8, mbw * 8), np.uint8)
    rgba = np.empty((n, h, w, 4), np.uint8) if nrgba else None
    ptrs = (C.c_char_p * n)(*streams)
    lens = (C.c_size_t * n)(*[len(s) for s in streams])
    ctx.check(native.lib().wgpu_decode_batch(ctx.handle, ptrs, lens, n, y.ctypes.data, u.ctypes.data, v.ctypes.data,
                                            y[0].nbytes, u[0].nbytes, rgba.ctypes.data if nrgba else None, w * h * 4))
    cw, chh = (w + 1) // 2, (h + 1) // 2
    imgs = [YCbCr(y[i, :h, :w].copy(), u[i, :chh, :cw].copy(), v[i, :chh, :cw].copy(), (0, 0, w, h)) for i in range(n)]
    return (imgs, rgba) if nrgba else imgs


def Decode(r, ctx=None):
    """webp.Decode (webp.go:88) for a lossy image without alpha: returns the 4:2:0 planes (no RGB conversion)."""
    return DecodeBatch([_read_all(r)], ctx=ctx)[0]


def decode_padded(streams, nrgba=False, ctx=None):
    """Parity-test helper: macroblock-padded planes exactly as lossy.DecodeFrame hands them out (decode.go:209)."""
    streams = [bytes(s) for s in streams]
    cfg = DecodeConfig(streams[0])
    w, h = cfg.Width, cfg.Height
    mbw, mbh = (w + 15) >> 4, (h + 15) >> 4
    n = len(streams)
    ctx = ctx or native.default_context()
    y = np.empty((n, mbh * 16, mbw * 16), np.uint8)
    u = np.empty((n, mbh * 8, mbw * 8), np.uint8)
    v = np.empty((n, mbh * 8, mbw * 8), np.uint8)
    rgba = np.empty((n, h, w, 4), np.uint8) if nrgba else None
    ptrs = (C.c_char_p * n)(*streams)
    lens = (C.c_size_t * n)(*[len(s) for s in streams])
    ctx.check(native.lib().wgpu_decode_batch(ctx.handle, ptrs, lens, n, y.ctypes.data, u.ctypes.data, v.ctypes.data,
                                            y[0].nbytes, u[0].nbytes, rgba.ctypes.data if nrgba else None, w * h * 4))
    return w, h, y, u, v, rgba
