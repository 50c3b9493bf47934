#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): lossy encode & decode Mpix/s, 1536x1024 q75 m4, batch of 256 images per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic images:
  value : encode throughput with the RGBA batch already resident in HBM (import + analysis + host segment plan +
          wavefront mode search; per-macroblock modes/levels left in HBM), CUDA events on the library's stream
  e2e   : the same batch through the reference-facing call wgpu_encode_batch (webp.Encode's batch twin): pinned
          host RGBA in, H2D, kernels, D2H of modes/levels, host token/bool coding, WebP files out
  decode: the streams produced above through wgpu_dec_* / wgpu_decode_batch (host parse, recon + loop filter +
          fancy upsampling on the GPU, NRGBA back to the host)
Images shard across ranks with no collective (weak scaling: every rank encodes its own batch).
--impl reference times the oracle (a C++ port of the reference's CPU path; no Go toolchain in this image) with all
host threads on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1536, 1024
ALG_BYTES_PER_PX = {"mode_search": 6.44, "import": 5.5, "analysis": 1.5, "recon": 4.66, "filter": 3.0, "upsample": 5.5}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


METRIC = "lossy encode & decode Mpix/s (1536x1024 q75 m4), bit-exact"
WORKLOAD = "synthetic 1536x1024 RGBA lossy encode q75 method 4, batch of %d images per GPU (BASELINE configs[1])"


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference encoder, one image per thread on all host cores, bounded sample."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from webp_b200.synth import synth_batch
    cores = os.cpu_count() or 1
    sample = max(cores, 8)  # images per step: >= one per thread, ~10-20 s per step on this workload
    imgs = synth_batch(sample, W, H, distinct=min(sample, 24))
    for _ in range(args.warmup if args.warmup < 2 else 1):
        oracle_lib.encode_batch(imgs[:cores], threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_lib.encode_batch(imgs, threads=cores)
    dt = time.perf_counter() - t0
    v = sample * W * H * args.steps / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": WORKLOAD % args.batch, "sample_images_per_step": sample,
                       "note": "the reference's row-parallel encoder (Method >= 3) restated in C++ (oracle/), all host cores"},
            "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port",
                             "sample": "%d images of the 1536x1024 q75 m4 workload per step, one image per thread (C++ oracle, -O2)" % sample},
            "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step (BASELINE configs[1]: 256)")
    ap.add_argument("--e2e-workers", type=int, default=5, help="contexts per GPU used by the e2e leg: upload, device and finish stages of consecutive batches overlap")
    ap.add_argument("--value-contexts", type=int, default=2, help="contexts whose device-resident steps run side by side in the `value` leg")
    ap.add_argument("--gpu-slots", type=int, default=2, help="how many contexts may have their mode-search waves on the GPU at once")
    ap.add_argument("--finish-slots", type=int, default=0, help="how many contexts may be in their finish stage at once (default: 3 when the token partitions are coded on the GPU, else 1)")
    ap.add_argument("--decode-workers", type=int, default=0, help="contexts per GPU for the decode e2e leg (default: 8 with the GPU macroblock parser, whose ~0.35 s latency per batch they hide; else the encode workers)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads per context (default: cores / ranks on this node)")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    from webp_b200 import native
    from webp_b200.synth import synth_batch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    L = native.lib()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = args.host_threads or max(1, (os.cpu_count() or 1) // max(1, local_world))
    ctx = native.Context(local, host_threads=host_threads)
    n, K = args.batch, args.steps
    px_step = n * W * H
    # pinned host staging: input batch and output files
    in_bytes = n * W * H * 4
    cap = W * H  # per-file capacity (bytes)
    opt = native.EncOptions()
    L.wgpu_enc_options_default(opt, 75)

    # the library codes the token partitions on the GPU for batches of at least 32 images (webpgpu.cu device_coder_wanted)
    env_coder = os.environ.get("WGPU_DEVICE_CODER", "")
    device_coder = (env_coder != "0") if env_coder else n >= 32
    finish_slots = args.finish_slots or (3 if device_coder else 1)
    upload_stage, gpu_stage, host_stage = threading.Lock(), threading.BoundedSemaphore(max(1, args.gpu_slots)), threading.BoundedSemaphore(finish_slots)

    class Worker:
        """One wgpu_ctx + its own pinned input/output staging.  Two workers per GPU let the host-side entropy coding of
        one batch overlap the GPU mode search of the next (contexts are independent by the ABI contract)."""

        def __init__(self, c, first_index):
            self.ctx = c
            self.h_in = L.wgpu_host_alloc(c.handle, in_bytes)
            self.h_out = L.wgpu_host_alloc(c.handle, n * cap)
            if not self.h_in or not self.h_out:
                raise SystemExit("pinned allocation failed")
            self.imgs = np.ctypeslib.as_array(C.cast(self.h_in, C.POINTER(C.c_uint8)), shape=(n, H, W, 4))
            self.imgs[:] = synth_batch(n, W, H, distinct=24, first_index=first_index)
            self.out = np.ctypeslib.as_array(C.cast(self.h_out, C.POINTER(C.c_uint8)), shape=(n, cap))
            self.sizes = np.zeros(n, np.uint64)

        def encode_e2e(self):
            """wgpu_encode_batch spelled as its three public stages so that two workers pipeline: one batch is in its GPU
            stage (H2D + kernels) while the previous one is in its host stage (D2H + token/bool coding)."""
            h = self.ctx.handle
            t = [time.perf_counter()]
            with upload_stage:  # H2D of this batch rides under the kernels of the batch before it
                t.append(time.perf_counter())
                self.ctx.check(L.wgpu_enc_upload(h, self.h_in, n, W, H, W * 4, W * H * 4))
                self.ctx.check(L.wgpu_sync(h))
            t.append(time.perf_counter())
            with gpu_stage:
                t.append(time.perf_counter())
                self.ctx.check(L.wgpu_enc_device(h, C.byref(opt)))
                self.ctx.check(L.wgpu_sync(h))
            t.append(time.perf_counter())
            with host_stage:
                t.append(time.perf_counter())
                self.ctx.check(L.wgpu_enc_finish(h, self.h_out, cap, self.sizes.ctypes.data))
            t.append(time.perf_counter())
            if os.environ.get("BENCH_TRACE"):  # wait-upload, upload, wait-gpu, device, wait-finish, finish (ms)
                sys.stderr.write("[bench] stages ms: " + " ".join("%.0f" % ((b - a) * 1e3) for a, b in zip(t, t[1:])) + "\n")

        def free(self):
            if self.h_in:
                L.wgpu_host_free(self.ctx.handle, self.h_in)
                L.wgpu_host_free(self.ctx.handle, self.h_out)
                self.h_in = self.h_out = None

    w0 = Worker(ctx, rank * 24)
    imgs, out, sizes, h_in = w0.imgs, w0.out, w0.sizes, w0.h_in
    workers = [w0]
    if args.e2e_workers > 1:
        workers += [Worker(native.Context(local, host_threads=host_threads), rank * 24) for _ in range(args.e2e_workers - 1)]
    for wk in workers:
        for _ in range(max(args.warmup, 3) if wk is w0 else 1):
            wk.encode_e2e()
    clocks = ClockSampler(local)
    clocks.start()
    # ---- value: device-resident encode (inputs already in HBM)
    # K steps (batches) in all, dealt to `--value-contexts` contexts whose streams run side by side: the 222-wave sequence of
    # one batch leaves the GPU partly empty at its narrow ends (66 % mean occupancy of the macroblock slots), a second
    # sequence fills them.  Timed with CUDA events on each context's stream from a common start; the slowest one counts.
    vws = workers[:max(1, min(args.value_contexts, len(workers)))]
    for wk in vws:
        wk.ctx.check(L.wgpu_enc_upload(wk.ctx.handle, wk.h_in, n, W, H, W * 4, W * H * 4))
        wk.ctx.check(L.wgpu_sync(wk.ctx.handle))
    launches0 = [wk.ctx.launch_count() for wk in vws]
    ms = C.c_float()
    shares = [K // len(vws) + (1 if i < K % len(vws) else 0) for i in range(len(vws))]
    v_ms = [0.0] * len(vws)
    go = threading.Barrier(len(vws))

    def value_run(i):
        wk, m = vws[i], C.c_float()
        go.wait()
        wk.ctx.check(L.wgpu_timer_begin(wk.ctx.handle))
        for _ in range(shares[i]):
            wk.ctx.check(L.wgpu_enc_device(wk.ctx.handle, C.byref(opt)))
        wk.ctx.check(L.wgpu_timer_end(wk.ctx.handle, C.byref(m)))
        v_ms[i] = m.value
    barrier()
    vths = [threading.Thread(target=value_run, args=(i,)) for i in range(len(vws))]
    for t in vths:
        t.start()
    for t in vths:
        t.join()
    barrier()
    dev_ms = max_over_ranks(max(v_ms))
    launches = sum(wk.ctx.launch_count() - l0 for wk, l0 in zip(vws, launches0))
    value = px_step * K * world / (dev_ms * 1e-3) / 1e6
    # ---- e2e: host RGBA -> WebP files through the public batch call; K batches in total, dealt to the workers
    for wk in workers:
        wk.ctx.transfer_bytes(reset=True)
    barrier()
    t0 = time.perf_counter()
    if len(workers) == 1:
        for _ in range(K):
            w0.encode_e2e()
    else:
        counter = iter(range(K))
        lock = threading.Lock()

        def run(wk):
            while True:
                with lock:
                    if next(counter, None) is None:
                        return
                wk.encode_e2e()
        ths = [threading.Thread(target=run, args=(wk,)) for wk in workers]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = px_step * K * world / e2e_s / 1e6
    nmb = ((W + 15) // 16) * ((H + 15) // 16)
    # bytes actually copied inside the timed region, counted by the library at every cudaMemcpy*Async it issues
    xfer = [wk.ctx.transfer_bytes() for wk in workers]
    h2d = sum(x[0] for x in xfer) // K  # RGBA + segment map + per-image segment parameters (+ token bases)
    d2h = sum(x[1] for x in xfer) // K  # analysis alphas + per-MB headers + probabilities + coded partitions (or tokens)
    # ---- per-kernel device times for the roofline (CUDA events on the library's stream)
    stage_ms = {}
    for name, sid, reps in (("import", 0, 5), ("analysis", 1, 5), ("mode_search", 2, max(1, min(K, 3)))):
        ctx.check(L.wgpu_enc_stage_time(ctx.handle, C.byref(opt), sid, reps, C.byref(ms)))
        stage_ms[name] = ms.value
    peak, peak_kind = measured_peak_gbs()
    dom = "mode_search"
    ach = ALG_BYTES_PER_PX[dom] * px_step / (stage_ms[dom] * 1e-3) / 1e9
    waves = (W // 16) + 2 * (H // 16 - 1)
    roofline = {"bound": "hbm", "kernel": "encode_wave_kernel (%d wave launches per step, timed together)" % waves, "achieved": ach,
                "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (wave 110 of 222: 12288 macroblocks = 3.15 Mpix, 20.3 MB
                # algorithmic) from the ncu --set full capture in profiles/r1_mode_search_h_end_of_round.md
                "traffic": 115.5e6, "traffic_launch": "wave 110 of 222 (12288 macroblocks, 20.3 MB algorithmic)",
                "note": "bound by dependent-instruction latency x the 222-step wavefront chain (issue slots 45 % busy), not by HBM; see DESIGN.md 5",
                "stages_ms": stage_ms,
                "stages_gbs": {k: ALG_BYTES_PER_PX[k] * px_step / (v * 1e-3) / 1e9 for k, v in stage_ms.items()}}
    result = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": K,
              "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "u8/int32", "data": "synthetic",
              "config": {"workload": WORKLOAD % n,
                         "batch_per_gpu": n, "l2": "inputs (%.0f MB RGBA per step) exceed the 126 MB L2" % (in_bytes / 1e6),
                         "parallelism": "images sharded across %d GPU(s), no collective" % world,
                         "batches_in_flight": len(vws)},
              "e2e": {"value": e2e, "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / K * 1e3,
                      "compressed_bytes_per_step": int(sizes.sum()), "workers_per_gpu": len(workers), "host_threads_per_worker": host_threads,
                      "token_partition_coder": "gpu" if device_coder else "host", "finish_slots": finish_slots, "gpu_slots": args.gpu_slots,
                      "api": "wgpu_enc_upload + wgpu_enc_device + wgpu_enc_finish (== wgpu_encode_batch: pinned host RGBA in, WebP files out), %d batches dealt to %d contexts whose upload / device / finish stages overlap" % (K, len(workers))},
              "gpu_launches": int(launches), "roofline": roofline}
    # ---- decode of the streams just produced (BASELINE configs[2])
    if not args.no_decode:
        files = [out[i, :int(sizes[i])].tobytes() for i in range(n)]
        ptrs = (C.c_char_p * n)(*files)
        lens = (C.c_size_t * n)(*[len(f) for f in files])
        h_rgba = L.wgpu_host_alloc(ctx.handle, in_bytes)
        for _ in range(2):
            ctx.check(L.wgpu_decode_batch(ctx.handle, ptrs, lens, n, None, None, None, 0, 0, h_rgba, W * H * 4))
        ctx.check(L.wgpu_dec_parse(ctx.handle, ptrs, lens, n, None, None))
        ctx.check(L.wgpu_sync(ctx.handle))
        l0 = ctx.launch_count()
        barrier()
        ctx.check(L.wgpu_timer_begin(ctx.handle))
        for _ in range(K):
            ctx.check(L.wgpu_dec_device(ctx.handle, 1))
        ctx.check(L.wgpu_timer_end(ctx.handle, C.byref(ms)))
        barrier()
        dms = max_over_ranks(ms.value)
        dl = ctx.launch_count() - l0
        # e2e: the staged public calls (wgpu_dec_parse -> wgpu_dec_device -> wgpu_dec_fetch == wgpu_decode_batch), K batches dealt to
        # the same contexts as the encode leg so the host parse of one batch overlaps the GPU + D2H stage of the previous one
        env_parser = os.environ.get("WGPU_DEVICE_PARSER", "")
        device_parser = (env_parser != "0") if env_parser else n >= 32  # webpgpu.cu device_parser_wanted
        # the encode staging of the extra contexts is no longer needed: give the pinned memory back before the decode buffers
        for wk in workers[1:]:
            wk.free()
        n_dec = args.decode_workers or (8 if device_parser else len(workers))
        dctxs = [wk.ctx for wk in workers] + [native.Context(local, host_threads=host_threads) for _ in range(max(0, n_dec - len(workers)))]
        dctxs = dctxs[:max(1, n_dec)]
        dec_bufs = [h_rgba] + [L.wgpu_host_alloc(c.handle, in_bytes) for c in dctxs[1:]]
        if not all(dec_bufs):
            raise SystemExit("pinned allocation failed")

        class DW:
            def __init__(self, c):
                self.ctx = c
        dworkers = [DW(c) for c in dctxs]
        for wk, buf in zip(dworkers[1:], dec_bufs[1:]):
            wk.ctx.check(L.wgpu_decode_batch(wk.ctx.handle, ptrs, lens, n, None, None, None, 0, 0, buf, W * H * 4))

        def decode_e2e(wk, buf):
            h = wk.ctx.handle
            if device_parser:
                # macroblocks are parsed on the GPU, one warp per image: a ~200 ms latency chain that occupies 3 % of the
                # warp slots, so the contexts only take turns for the D2H of the finished batch
                wk.ctx.check(L.wgpu_dec_parse(h, ptrs, lens, n, None, None))
                wk.ctx.check(L.wgpu_dec_device(h, 1))
                wk.ctx.check(L.wgpu_sync(h))
                with gpu_stage:
                    wk.ctx.check(L.wgpu_dec_fetch(h, None, None, None, 0, 0, buf, W * H * 4))
                return
            with host_stage:
                wk.ctx.check(L.wgpu_dec_parse(h, ptrs, lens, n, None, None))
            with gpu_stage:
                wk.ctx.check(L.wgpu_dec_device(h, 1))
                wk.ctx.check(L.wgpu_dec_fetch(h, None, None, None, 0, 0, buf, W * H * 4))
        for wk in dworkers:
            wk.ctx.transfer_bytes(reset=True)
        barrier()
        t0 = time.perf_counter()
        if len(dworkers) == 1:
            for _ in range(K):
                decode_e2e(w0, h_rgba)
        else:
            dcounter = iter(range(K))
            dlock = threading.Lock()

            def drun(wk, buf):
                while True:
                    with dlock:
                        if next(dcounter, None) is None:
                            return
                    decode_e2e(wk, buf)
            ths = [threading.Thread(target=drun, args=(wk, buf)) for wk, buf in zip(dworkers, dec_bufs)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        barrier()
        ds = max_over_ranks(time.perf_counter() - t0)
        dxfer = [wk.ctx.transfer_bytes() for wk in dworkers]
        for wk, buf in zip(dworkers[1:], dec_bufs[1:]):
            L.wgpu_host_free(wk.ctx.handle, buf)
        result["decode"] = {"value": px_step * K * world / (dms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": dms / K, "gpu_launches": int(dl),
                            "e2e": {"value": px_step * K * world / ds / 1e6, "unit": "Mpix/s", "ms_per_step": ds / K * 1e3,
                                    "h2d_bytes_per_step": sum(x[0] for x in dxfer) // K, "d2h_bytes_per_step": sum(x[1] for x in dxfer) // K,
                                    "macroblock_parser": "gpu" if device_parser else "host", "workers_per_gpu": len(dworkers)},
                            "config": {"workload": "decode of the %d streams above -> recon + loop filter + fancy upsampling to NRGBA (BASELINE configs[2])" % n},
                            "roofline": {"bound": "hbm", "kernel": "recon_wave + filter_wave + upsample_nrgba (whole device step)",
                                         "achieved": (ALG_BYTES_PER_PX["recon"] + ALG_BYTES_PER_PX["filter"] + ALG_BYTES_PER_PX["upsample"]) * px_step / (dms / K * 1e-3) / 1e9,
                                         "peak": peak, "unit": "GB/s"}}
        result["decode"]["roofline"]["frac"] = result["decode"]["roofline"]["achieved"] / peak
        L.wgpu_host_free(ctx.handle, h_rgba)
    result["clocks"] = clocks.stop()
    # ---- CPU baseline beside it: the oracle port on this box's host cores, bounded sample, rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        cores = os.cpu_count() or 1
        sample = max(cores, 8)
        sub = np.ascontiguousarray(imgs[:sample] if sample <= n else synth_batch(sample, W, H))
        t0 = time.perf_counter()
        oracle_lib.encode_batch(sub, threads=cores)
        dt = time.perf_counter() - t0
        result["cpu_baseline"] = {"value": sample * W * H / dt / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
                                  "sample": "%d images of the same batch, one image per thread, %d threads (C++ oracle -O2; the Go reference cannot be built here)" % (sample, cores)}
    for wk in workers:
        wk.free()
    if rank == 0:
        print(json.dumps(result), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
