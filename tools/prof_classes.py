"""Mode-search device time per content class of the synthetic generator (0 gradient, 1 gradient + noise + rectangles, 2 noisy):
python tools/prof_classes.py [n] [w] [h]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from webp_b200 import native
from webp_b200.synth import synth_image
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
L = native.lib(); ctx = native.Context(0)
opt = native.EncOptions(); L.wgpu_enc_options_default(opt, 75)
ms = C.c_float()
for kind in (0, 1, 2):
    base = np.stack([synth_image(w, h, 3 * i + kind, kind=kind) for i in range(8)])
    imgs = np.concatenate([base] * ((n + 7) // 8))[:n]
    ctx.check(L.wgpu_enc_upload(ctx.handle, imgs.ctypes.data, n, w, h, w * 4, w * h * 4))
    ctx.check(L.wgpu_enc_device(ctx.handle, C.byref(opt)))
    ctx.check(L.wgpu_enc_stage_time(ctx.handle, C.byref(opt), 2, 2, C.byref(ms)))
    print("class %d: mode_search %.2f ms per %d images -> %.0f Mpix/s" % (kind, ms.value, n, n * w * h / ms.value / 1e3))
