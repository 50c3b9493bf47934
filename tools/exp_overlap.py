"""Experiment: does the device boolean coder (finish stage of context A) overlap the mode-search waves (device stage of
context B)?  Prints the duration of each alone and when run concurrently.  python tools/exp_overlap.py [n] [w] [h]"""
import ctypes as C, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from webp_b200 import native
from webp_b200.synth import synth_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
L = native.lib()
imgs = synth_batch(n, w, h, distinct=min(n, 12))
cap = w * h + 65536
opt = native.EncOptions(); L.wgpu_enc_options_default(opt, 75)


class W:
    def __init__(self):
        self.ctx = native.Context(0)
        self.h_in = L.wgpu_host_alloc(self.ctx.handle, imgs.nbytes)
        self.h_out = L.wgpu_host_alloc(self.ctx.handle, n * cap)
        C.memmove(self.h_in, imgs.ctypes.data, imgs.nbytes)
        self.sizes = np.zeros(n, np.uint64)

    def device(self):
        t = time.perf_counter()
        self.ctx.check(L.wgpu_enc_upload(self.ctx.handle, self.h_in, n, w, h, w * 4, w * h * 4))
        self.ctx.check(L.wgpu_enc_device(self.ctx.handle, C.byref(opt)))
        self.ctx.check(L.wgpu_sync(self.ctx.handle))
        return (time.perf_counter() - t) * 1e3

    def finish(self):
        t = time.perf_counter()
        self.ctx.check(L.wgpu_enc_finish(self.ctx.handle, self.h_out, cap, self.sizes.ctypes.data))
        return (time.perf_counter() - t) * 1e3


a, b = W(), W()
for k in range(2):
    a.device(); a.finish(); b.device(); b.finish()
print("alone: device %.1f ms, finish %.1f ms" % (b.device(), (a.device(), a.finish())[1]))
for rep in range(3):
    a.device()
    res = {}
    ta = threading.Thread(target=lambda: res.__setitem__("finish", a.finish()))
    tb = threading.Thread(target=lambda: res.__setitem__("device", b.device()))
    t0 = time.perf_counter()
    ta.start(); tb.start(); ta.join(); tb.join()
    print("concurrent: finish(A) %.1f ms, device(B) %.1f ms, both done after %.1f ms" % (res["finish"], res["device"], (time.perf_counter() - t0) * 1e3))
    b.finish()
