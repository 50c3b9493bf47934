"""BASELINE configs[4]: thumbnail workload -- N synthetic 256x256 images, lossy q80 method 2 (serial-path semantics),
batched on one GPU.  Prints one JSON line: e2e Mpix/s (host RGBA -> WebP files), device ms, CPU oracle on all host cores.
  python tools/bench_thumbs.py [n_images=10000] [batch=2500] [contexts=3]"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from webp_b200 import native
from webp_b200.synth import synth_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
W = H = 256
import threading
NCTX = int(sys.argv[3]) if len(sys.argv) > 3 else 3  # contexts in flight: upload / kernels / coder / D2H of consecutive batches overlap
L = native.lib()
opt = native.EncOptions(); L.wgpu_enc_options_default(opt, 80); opt.method = 2
cap = W * H
class Wk:
    def __init__(self, first):
        self.ctx = native.Context(0)
        self.h_in = L.wgpu_host_alloc(self.ctx.handle, B * W * H * 4)
        self.h_out = L.wgpu_host_alloc(self.ctx.handle, B * cap)
        self.imgs = np.ctypeslib.as_array(C.cast(self.h_in, C.POINTER(C.c_uint8)), shape=(B, H, W, 4))
        self.imgs[:] = synth_batch(B, W, H, distinct=48, first_index=first)
        self.sizes = np.zeros(B, np.uint64)
    def step(self):
        self.ctx.check(L.wgpu_encode_batch(self.ctx.handle, self.h_in, B, W, H, W * 4, W * H * 4, C.byref(opt), self.h_out, cap, self.sizes.ctypes.data))
wks = [Wk(0) for _ in range(NCTX)]
for wk in wks:
    for _ in range(2): wk.step()
ctx, h_in, imgs, sizes = wks[0].ctx, wks[0].h_in, wks[0].imgs, wks[0].sizes
steps = (N + B - 1) // B
ms = C.c_float()
counter, lock = iter(range(steps)), threading.Lock()
def run(wk):
    while True:
        with lock:
            if next(counter, None) is None: return
        wk.step()
ths = [threading.Thread(target=run, args=(wk,)) for wk in wks]
t0 = time.perf_counter()
for t in ths: t.start()
for t in ths: t.join()
dt = time.perf_counter() - t0
ctx.check(L.wgpu_enc_upload(ctx.handle, h_in, B, W, H, W * 4, W * H * 4)); ctx.check(L.wgpu_sync(ctx.handle))
ctx.check(L.wgpu_timer_begin(ctx.handle)); ctx.check(L.wgpu_enc_device(ctx.handle, C.byref(opt))); ctx.check(L.wgpu_timer_end(ctx.handle, C.byref(ms)))
import oracle_lib
cores = os.cpu_count() or 1
sample = np.ascontiguousarray(imgs[:max(256, 16 * cores)])
t1 = time.perf_counter(); oracle_lib.encode_batch(sample, oracle_lib.default_cfg(quality=80, method=2), threads=cores); cdt = time.perf_counter() - t1
print(json.dumps({"workload": "%d synthetic 256x256 images, lossy q80 method 2, batches of %d (BASELINE configs[4])" % (steps * B, B),
                  "contexts": NCTX, "e2e_mpix_s": steps * B * W * H / dt / 1e6, "e2e_images_s": steps * B / dt, "device_ms_per_batch": ms.value,
                  "device_mpix_s": B * W * H / ms.value / 1e3, "compressed_bytes_per_batch": int(sizes.sum()),
                  "cpu_oracle_mpix_s": len(sample) * W * H / cdt / 1e6, "cpu_cores": cores}))
