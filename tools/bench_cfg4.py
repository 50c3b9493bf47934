"""BASELINE configs[3]: 3840x2160 lossy encode, Method 6, TargetPSNR multi-pass.  Times the public batch call end to end
(host RGBA in, WebP files out) and the CPU oracle on the same images.  python tools/bench_cfg4.py [batch] [w] [h]"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import webp_b200
from webp_b200 import native
from webp_b200.synth import synth_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
h = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
ctx = native.Context(0)
imgs = synth_batch(n, w, h, distinct=min(n, 8))
o = webp_b200.DefaultOptions()
o.Method = 6
o.TargetPSNR = 42.0
webp_b200.EncodeBatch(imgs[:2], o, ctx)  # warm-up
l0 = ctx.launch_count()
t0 = time.perf_counter()
files = webp_b200.EncodeBatch(imgs, o, ctx)
dt = time.perf_counter() - t0
print("GPU: %d images %dx%d in %.2f s -> %.1f Mpix/s, %d launches, %.1f KB/file" % (n, w, h, dt, n * w * h / dt / 1e6, ctx.launch_count() - l0,
                                                                                  sum(len(f) for f in files) / n / 1e3))
if os.environ.get("WITH_ORACLE"):
    import oracle_lib as O
    k = min(n, int(os.environ["WITH_ORACLE"]))
    cfg = O.default_cfg(method=6, target_psnr=42.0)
    t0 = time.perf_counter()
    exp = [O.encode(imgs[i], cfg) for i in range(k)]
    dt = time.perf_counter() - t0
    print("oracle: %d images in %.2f s on one core -> %.1f Mpix/s per core; identical: %s" % (k, dt, k * w * h / dt / 1e6, all(exp[i] == files[i] for i in range(k))))
