"""Small encode (+ optional decode) job for ncu captures: python tools/prof_encode.py [n] [w] [h] [reps] [decode]
STAGE_TIME=1 prints per-stage device times; SHARP=1 encodes with UseSharpYUV (stage "import" is then the three SharpYUV kernels)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from webp_b200 import native
from webp_b200.synth import synth_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = int(sys.argv[2]) if len(sys.argv) > 2 else 512
h = int(sys.argv[3]) if len(sys.argv) > 3 else 512
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dec = len(sys.argv) > 5
L = native.lib(); ctx = native.Context(0)
imgs = synth_batch(n, w, h, distinct=min(n, 12))
cap = w * h + 65536
out = np.empty((n, cap), np.uint8); sizes = np.zeros(n, np.uint64)
opt = native.EncOptions(); L.wgpu_enc_options_default(opt, 75)
opt.use_sharp_yuv = 1 if os.environ.get("SHARP") else 0
for _ in range(reps):
    ctx.check(L.wgpu_encode_batch(ctx.handle, imgs.ctypes.data, n, w, h, w * 4, w * h * 4, C.byref(opt), out.ctypes.data, cap, sizes.ctypes.data))
ms = C.c_float()
if os.environ.get("STAGE_TIME"):
    for sid, name in ((0, "import"), (1, "analysis"), (2, "mode_search")):
        ctx.check(L.wgpu_enc_stage_time(ctx.handle, C.byref(opt), sid, 2, C.byref(ms)))
        print("stage %s: %.3f ms -> %.1f Mpix/s" % (name, ms.value, n * w * h / ms.value / 1e3))
print("encoded", n, "images", w, h, "bytes", int(sizes.sum()), "launches", ctx.launch_count())
if dec:
    files = [out[i, :int(sizes[i])].tobytes() for i in range(n)]
    ptrs = (C.c_char_p * n)(*files); lens = (C.c_size_t * n)(*[len(f) for f in files])
    rgba = np.empty((n, h, w, 4), np.uint8)
    ctx.check(L.wgpu_decode_batch(ctx.handle, ptrs, lens, n, None, None, None, 0, 0, rgba.ctypes.data, w * h * 4))
    print("decoded; launches", ctx.launch_count())
