"""Small workload over every encoder route, the device parser and the boolean-coder twin, for compute-sanitizer:
  compute-sanitizer --tool memcheck python tools/sanitize_workload.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import oracle_lib as O, webp_b200
from webp_b200 import native, dsp
ctx = native.Context(0)
imgs = np.stack([O.synth_image(200, 150, i) for i in range(5)])
o = webp_b200.DefaultOptions()
f = webp_b200.EncodeBatch(imgs, o, ctx)
assert all(f[k] == O.encode(imgs[k]) for k in range(5))
big = np.stack([O.synth_image(640, 480, i) for i in (2, 5)])
f2 = webp_b200.EncodeBatch(big, o, ctx)
assert all(f2[k] == O.encode(big[k]) for k in range(2))
o2 = webp_b200.DefaultOptions(); o2.Method = 2; o2.Quality = 80
f3 = webp_b200.EncodeBatch(imgs, o2, ctx)
o3 = webp_b200.DefaultOptions(); o3.TargetSize = 3000
f4 = webp_b200.EncodeBatch(imgs, o3, ctx)
os.environ["WGPU_DEVICE_PARSER"] = "1"
w, h, y, u, v, rgba = webp_b200.webp.decode_padded(f2, nrgba=True, ctx=ctx)
rng = np.random.default_rng(1)
streams = [O.adversarial_tokens(rng, n, m) for n, m in ((0, 0), (5, 1), (9000, 1), (30000, 3), (20000, 4))]
got, r = dsp.BoolCodeBatch(streams, ctx)
assert all(np.array_equal(g, O.boolcode(t)) for g, t in zip(got, streams))
print("sanitizer workload ok", r)
