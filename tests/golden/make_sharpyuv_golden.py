"""Generates sharpyuv_libsharpyuv.json: SHA-256 of the Y/U/V planes libsharpyuv (the C library the reference's
testc/sharpyuv suite compares sharpyuv.Convert with) produces for seeded inputs, WebP matrix, sRGB transfer, 8 bits.
Run in the build container: python tests/golden/make_sharpyuv_golden.py   (needs Pillow's bundled libsharpyuv)."""
import ctypes as C, glob, hashlib, json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from webp_b200.synth import synth_image

CASES = [(32, 32, 0), (33, 17, 1), (1, 1, 2), (2, 1, 3), (1, 2, 3), (5, 7, 4), (128, 96, 5), (130, 71, 6), (256, 192, 7), (640, 480, 8),
         (100, 1, 9), (1, 100, 10), (64, 64, -1), (97, 33, -2), (2, 2, -3)]


def case_image(w, h, idx):
    if idx >= 0:
        return synth_image(w, h, idx)
    if idx == -3:  # TestConvertSharp_2x2 (sharpyuv/sharpyuv_test.go:369)
        return np.array([[[255, 0, 0, 255], [0, 255, 0, 255]], [[0, 0, 255, 255], [255, 255, 0, 255]]], np.uint8)
    rng = np.random.default_rng(100 - idx)
    return rng.integers(0, 256, (h, w, 4), dtype=np.uint8) if idx == -1 else (rng.integers(0, 2, (h, w, 4), dtype=np.uint8) * 255)


def libsharpyuv():
    import PIL
    so = glob.glob(os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs", "libsharpyuv*"))
    if not so:
        return None
    S = C.CDLL(so[0])
    S.SharpYuvGetConversionMatrix.restype = C.c_void_p
    S.SharpYuvInit(None)
    return S


def convert(S, rgba):
    h, w = rgba.shape[:2]
    rgb = np.ascontiguousarray(rgba[..., :3])
    y = np.zeros((h, w), np.uint8); u = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8); v = np.zeros_like(u)
    p = rgb.ctypes.data
    ok = S.SharpYuvConvert(C.c_void_p(p), C.c_void_p(p + 1), C.c_void_p(p + 2), 3, w * 3, 8, y.ctypes.data_as(C.c_void_p), w,
                           u.ctypes.data_as(C.c_void_p), u.shape[1], v.ctypes.data_as(C.c_void_p), v.shape[1], 8, w, h,
                           C.c_void_p(S.SharpYuvGetConversionMatrix(0)))  # kSharpYuvMatrixWebp
    assert ok == 1
    return y, u, v


def digest(y, u, v):
    return hashlib.sha256(y.tobytes() + u.tobytes() + v.tobytes()).hexdigest()


if __name__ == "__main__":
    S = libsharpyuv()
    out = {"library": "libsharpyuv %#x (Pillow wheel)" % S.SharpYuvGetVersion(), "cases": []}
    for (w, h, idx) in CASES:
        out["cases"].append({"w": w, "h": h, "image": idx, "sha256": digest(*convert(S, case_image(w, h, idx)))})
    json.dump(out, open(os.path.join(HERE, "sharpyuv_libsharpyuv.json"), "w"), indent=1)
    print("wrote", len(out["cases"]), "cases from", out["library"])
