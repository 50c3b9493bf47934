"""Independent cross-check: libwebp 1.6.0 as bundled with Pillow, driven through ctypes.

VP8 decoding is normative, so libwebp's WebPDecodeYUV / WebPDecodeRGBA pin the oracle's decoder,
loop filter and fancy upsampler; WebPPictureImportRGBA pins the RGB->YUV import.  (It cannot pin
encoder decisions: the reference's analysis / trellis are its own variants)
Optional: tests skip when the library is not present.
"""
import ctypes as C
import glob
import os

import numpy as np

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        try:
            import PIL
        except ImportError:
            return None
        cands = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libwebp-*.so*"))
        if not cands:
            return None
        for dep in glob.glob(os.path.join(os.path.dirname(cands[0]), "libsharpyuv-*.so*")):
            C.CDLL(dep, mode=C.RTLD_GLOBAL)  # dependency of libwebp, same private directory
        _LIB = C.CDLL(cands[0])
        _LIB.WebPDecodeYUV.restype = C.POINTER(C.c_uint8)
        _LIB.WebPDecodeRGBA.restype = C.POINTER(C.c_uint8)
        _LIB.WebPFree.argtypes = [C.c_void_p]
    return _LIB


def decode_yuv(data):
    L = lib()
    w, h, stride, uv_stride = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    u, v = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint8)()
    y = L.WebPDecodeYUV(data, C.c_size_t(len(data)), C.byref(w), C.byref(h), C.byref(u), C.byref(v),
                        C.byref(stride), C.byref(uv_stride))
    if not y:
        raise RuntimeError("WebPDecodeYUV failed")
    W, H = w.value, h.value
    ya = np.ctypeslib.as_array(y, shape=(H, stride.value))[:, :W].copy()
    uh, uw = (H + 1) // 2, (W + 1) // 2
    ua = np.ctypeslib.as_array(u, shape=(uh, uv_stride.value))[:, :uw].copy()
    va = np.ctypeslib.as_array(v, shape=(uh, uv_stride.value))[:, :uw].copy()
    L.WebPFree(y)
    return W, H, ya, ua, va


def decode_rgba(data):
    L = lib()
    w, h = C.c_int(), C.c_int()
    p = L.WebPDecodeRGBA(data, C.c_size_t(len(data)), C.byref(w), C.byref(h))
    if not p:
        raise RuntimeError("WebPDecodeRGBA failed")
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
    L.WebPFree(p)
    return a


class _WebPPicture(C.Structure):
    # libwebp encode.h WebPPicture (ABI 0x020f); only the leading fields are accessed.
    _fields_ = [("use_argb", C.c_int), ("colorspace", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("y", C.POINTER(C.c_uint8)), ("u", C.POINTER(C.c_uint8)), ("v", C.POINTER(C.c_uint8)),
                ("y_stride", C.c_int), ("uv_stride", C.c_int), ("a", C.POINTER(C.c_uint8)), ("a_stride", C.c_int),
                ("pad1", C.c_uint32 * 2), ("argb", C.POINTER(C.c_uint32)), ("argb_stride", C.c_int),
                ("pad2", C.c_uint32 * 3), ("tail", C.c_uint8 * 256)]


def import_rgba_yuv(rgba):
    """RGBA -> YUV420 through WebPPictureImportRGBA with use_argb=0 (libwebp picture_csp_enc.c)."""
    L = lib()
    pic = _WebPPicture()
    if not L.WebPPictureInitInternal(C.byref(pic), 0x020f):
        raise RuntimeError("WebPPictureInit failed (ABI mismatch)")
    h, w = rgba.shape[:2]
    pic.use_argb = 0
    pic.width, pic.height = w, h
    rgba = np.ascontiguousarray(rgba)
    if not L.WebPPictureImportRGBA(C.byref(pic), rgba.ctypes.data_as(C.c_void_p), C.c_int(rgba.strides[0])):
        raise RuntimeError("WebPPictureImportRGBA failed")
    y = np.ctypeslib.as_array(pic.y, shape=(h, pic.y_stride))[:, :w].copy()
    uh, uw = (h + 1) // 2, (w + 1) // 2
    u = np.ctypeslib.as_array(pic.u, shape=(uh, pic.uv_stride))[:, :uw].copy()
    v = np.ctypeslib.as_array(pic.v, shape=(uh, pic.uv_stride))[:, :uw].copy()
    L.WebPPictureFree(C.byref(pic))
    return y, u, v
