"""Decode timings on the bench streams: python tools/prof_decode.py [n] [w] [h]
prints wall time of wgpu_dec_parse + sync (GPU macroblock parser), device time of wgpu_dec_device and one-shot wgpu_decode_batch."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from webp_b200 import native
from webp_b200.synth import synth_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
L = native.lib(); ctx = native.Context(0)
imgs = synth_batch(n, w, h, distinct=min(n, 24))
cap = w * h
out = np.empty((n, cap), np.uint8); sizes = np.zeros(n, np.uint64)
opt = native.EncOptions(); L.wgpu_enc_options_default(opt, 75)
ctx.check(L.wgpu_encode_batch(ctx.handle, imgs.ctypes.data, n, w, h, w * 4, w * h * 4, C.byref(opt), out.ctypes.data, cap, sizes.ctypes.data))
files = [out[i, :int(sizes[i])].tobytes() for i in range(n)]
ptrs = (C.c_char_p * n)(*files); lens = (C.c_size_t * n)(*[len(f) for f in files])
h_rgba = L.wgpu_host_alloc(ctx.handle, n * w * h * 4)
for rep in range(3):
    t0 = time.perf_counter()
    ctx.check(L.wgpu_dec_parse(ctx.handle, ptrs, lens, n, None, None)); ctx.check(L.wgpu_sync(ctx.handle))
    t1 = time.perf_counter()
    ms = C.c_float()
    ctx.check(L.wgpu_timer_begin(ctx.handle)); ctx.check(L.wgpu_dec_device(ctx.handle, 1)); ctx.check(L.wgpu_timer_end(ctx.handle, C.byref(ms)))
    t2 = time.perf_counter()
    ctx.check(L.wgpu_dec_fetch(ctx.handle, None, None, None, 0, 0, h_rgba, w * h * 4))
    t3 = time.perf_counter()
    print("parse+sync %.1f ms   device %.2f ms   fetch %.1f ms (%.1f GB/s)" % ((t1 - t0) * 1e3, ms.value, (t3 - t2) * 1e3, n * w * h * 4 / (t3 - t2) / 1e9))
t0 = time.perf_counter()
ctx.check(L.wgpu_decode_batch(ctx.handle, ptrs, lens, n, None, None, None, 0, 0, h_rgba, w * h * 4))
print("one-shot wgpu_decode_batch %.1f ms -> %.0f Mpix/s" % ((time.perf_counter() - t0) * 1e3, n * w * h / (time.perf_counter() - t0) / 1e6))
