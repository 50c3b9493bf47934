"""Device time and achieved HBM bandwidth of the streaming kernels on the bench workload (CUDA events on the library's stream):
import (5.5 B/px), analysis (1.5 B/px), SSE+SSIM (2 B/px), fancy upsampling to NRGBA (5.5 B/px).  SURVEY.md 8(d) byte counts.
  python tools/prof_stream.py [n] [w] [h] [reps]"""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from webp_b200 import native
from webp_b200.synth import synth_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
L = native.lib(); ctx = native.Context(0)
imgs = synth_batch(n, w, h, distinct=min(n, 12))
cap = w * h + 65536
out = np.empty((n, cap), np.uint8); sizes = np.zeros(n, np.uint64)
opt = native.EncOptions(); L.wgpu_enc_options_default(opt, 75)
ctx.check(L.wgpu_encode_batch(ctx.handle, imgs.ctypes.data, n, w, h, w * 4, w * h * 4, C.byref(opt), out.ctypes.data, cap, sizes.ctypes.data))
files = [out[i, :int(sizes[i])].tobytes() for i in range(n)]
ptrs = (C.c_char_p * n)(*files); lens = (C.c_size_t * n)(*[len(f) for f in files])
ctx.check(L.wgpu_dec_parse(ctx.handle, ptrs, lens, n, None, None)); ctx.check(L.wgpu_dec_device(ctx.handle, 1)); ctx.check(L.wgpu_sync(ctx.handle))
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
ms = C.c_float()
mbw, mbh = (w + 15) // 16, (h + 15) // 16
for sid, name, bpp, px in ((0, "import_rgba_kernel", 5.5, n * w * h), (1, "analysis_kernel", 1.5, n * mbw * mbh * 256), (3, "ssim_sep_kernel SSE+SSIM", 2.0, n * mbw * mbh * 256), (5, "sse_kernel (PSNR only)", 2.0, n * mbw * mbh * 256),
                           (4, "upsample_nrgba_kernel", 5.5, n * w * h)):
    ctx.check(L.wgpu_enc_stage_time(ctx.handle, C.byref(opt), sid, 1, C.byref(ms)))  # warm-up: first-use allocations stay outside the timing
    ctx.check(L.wgpu_enc_stage_time(ctx.handle, C.byref(opt), sid, reps, C.byref(ms)))
    gbs = bpp * px / (ms.value * 1e-3) / 1e9
    print("%-26s %.3f ms  %.0f GB/s algorithmic = %.1f %% of the measured %.0f GB/s  (%.1f Gpix/s)" % (name, ms.value, gbs, 100 * gbs / peak, peak, px / ms.value / 1e6))
if os.environ.get("SSIM_REPEAT"):  # steadiness of the SSE+SSIM stage: SSIM_REPEAT=8 python tools/prof_stream.py
    for k in range(int(os.environ["SSIM_REPEAT"])):
        ctx.check(L.wgpu_enc_stage_time(ctx.handle, C.byref(opt), 3, 4, C.byref(ms)))
        print("ssim repeat %d: %.3f ms" % (k, ms.value))
