"""How the GPU macroblock parser (dec_parse_kernel, one chain per image) behaves with k contexts in flight:
python tools/prof_decode_conc.py [n] [w] [h]   -- wall time of wgpu_dec_parse + sync per context, k = 1, 2, 4, 8 at once."""
import ctypes as C, os, sys, threading, time
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
print("stream bytes: min %d  mean %d  max %d" % (min(map(len, files)), sum(map(len, files)) // n, max(map(len, files))))
ptrs = (C.c_char_p * n)(*files); lens = (C.c_size_t * n)(*[len(f) for f in files])
ctxs = [ctx] + [native.Context(0) for _ in range(7)]
for c in ctxs:  # buffers reserved, kernels loaded
    c.check(L.wgpu_dec_parse(c.handle, ptrs, lens, n, None, None)); c.check(L.wgpu_sync(c.handle))
for k in (1, 2, 4, 8):
    times = [0.0] * k
    bar = threading.Barrier(k)

    def run(i):
        c = ctxs[i]
        bar.wait()
        t0 = time.perf_counter()
        for _ in range(2):
            c.check(L.wgpu_dec_parse(c.handle, ptrs, lens, n, None, None)); c.check(L.wgpu_sync(c.handle))
        times[i] = (time.perf_counter() - t0) / 2
    ths = [threading.Thread(target=run, args=(i,)) for i in range(k)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    wall = time.perf_counter() - t0
    print("k = %d contexts: parse+sync per batch %s ms; %d batches in %.0f ms -> %.2f Gpix/s" % (
        k, " ".join("%.0f" % (x * 1e3) for x in times), 2 * k, wall * 1e3, 2 * k * n * w * h / wall / 1e9))
