import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import webp_b200
from webp_b200 import native, dsp
sys.path.insert(0, os.path.join(ROOT, "tests"))
ctx = native.Context(0)
def run(w, h, idxs, **kw):
    o = webp_b200.DefaultOptions()
    for k, v in kw.items(): setattr(o, k, v)
    imgs = np.stack([O.synth_image(w, h, i) for i in idxs])
    cfg = webp_b200.webp.lossy_config(o)
    ocfg = O.default_cfg(**{f: getattr(cfg, f) for f, _ in cfg._fields_})
    y, u, v = dsp.ImportRGBA(imgs, False, ctx, dither_amp=cfg.dither_amp)
    for k in range(len(idxs)):
        ey, eu, ev = O.import_rgba(imgs[k], False, dither_amp=cfg.dither_amp)
        print("import", idxs[k], "amp", cfg.dither_amp, np.array_equal(y[k], ey), np.array_equal(u[k], eu), np.array_equal(v[k], ev))
    files = webp_b200.EncodeBatch(imgs, o, ctx)
    from test_gpu_codec import _fetch
    for k in range(len(idxs)):
        exp, t = O.encode(imgs[k], ocfg, taps=True)
        g = _fetch(ctx, k, w, h)
        msg = []
        for name, a, b in (("alphas", g["alphas"], t["alphas"]), ("segment", g["mb_hdr"][:, 3], t["mb_hdr"][:, 3]), ("type", g["mb_hdr"][:, 0], t["mb_hdr"][:, 0]),
                           ("modes", g["mb_modes"], t["mb_modes"]), ("coeffs", g["mb_coeffs"], t["mb_coeffs"]), ("src_y", g["src_y"], t["src_y"])):
            if not np.array_equal(a, b):
                idx = np.argwhere(a != b)
                msg.append("%s: %d diffs first %s" % (name, len(idx), idx[0].tolist()))
        print("image", idxs[k], "bytes equal", files[k] == exp, len(files[k]), len(exp), msg, "seg quant", t["seg"][:, 0].tolist())
os.environ["WGPU_DEVICE_CODER"] = "1"; os.environ["WGPU_DEVICE_PARSER"] = "1"
try:
    run(9, 116, [2, 10, 4], Quality=0, Method=6, TargetPSNR=30.0, SNSStrength=30, FilterStrength=0, FilterSharpness=3, Segments=4, Pass=2, QMin=10, QMax=80)
except Exception as e:
    print("ERR", e)
try:
    imgs = np.stack([O.synth_image(9, 116, i) for i in (2, 10, 4)])
    f = webp_b200.EncodeBatch(imgs, webp_b200.DefaultOptions(), ctx)
    print("plain encode ok", [len(x) for x in f])
    print(webp_b200.webp.decode_padded(f, nrgba=True, ctx=ctx)[0:2])
except Exception as e:
    print("ERR2", e)
