#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): lossy encode & decode Mpix/s, 1536x1024 q75 m4, batch of 256 images per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|4|5]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic images:
  value : encode throughput with the RGBA batch already resident in HBM: wgpu_enc_device + wgpu_enc_finish, i.e. import +
          analysis + host segment plan + mode search + token generation + boolean coding of both partitions on the GPU and
          the WebP files laid out in pinned host memory; CUDA events on the library's stream around the whole sequence
  e2e   : the same batch through the reference-facing call wgpu_encode_batch (webp.Encode's batch twin), spelled as its
          three public stages: pinned host RGBA in, H2D, kernels, coded partitions back, WebP files out
  decode: the streams produced above through wgpu_dec_* / wgpu_decode_batch (parse, recon + loop filter + fancy
          upsampling on the GPU, NRGBA back to the host)                                  [config 2 only]
After the timed regions the outputs of EVERY context that took part are compared with the oracle (files byte for byte,
decoded planes and NRGBA); a mismatch makes the run fail.  "parity_checked" counts the comparisons.
--config picks the workload (numbers as in VERDICT/BASELINE.md, 1-based): 2 = BASELINE configs[1] (default, + its decode
configs[2]), 4 = configs[3] 3840x2160 Method 6 TargetPSNR, 5 = configs[4] thumbnails 256x256 q80 Method 2.
Images shard across ranks with no collective (weak scaling: every rank encodes its own batches).
--impl reference times the oracle (a C++ port of the reference's CPU path; no Go toolchain in this image) with all host
threads on a bounded sample of the same workload.
"""
import argparse
import concurrent.futures as cf
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

ALG_BYTES_PER_PX = {"mode_search": 6.44, "import": 5.5, "analysis": 1.5, "recon": 4.66, "filter": 3.0, "upsample": 5.5}
METRIC = "lossy encode & decode Mpix/s (1536x1024 q75 m4), bit-exact"
CONFIGS = {
    2: dict(w=1536, h=1024, quality=75, method=4, target_psnr=0.0, batch=256, distinct=24, steps=12, e2e_workers=3, value_contexts=2, decode=True,
            metric=METRIC, workload="synthetic 1536x1024 RGBA lossy encode q75 method 4, batch of %d images per GPU (BASELINE configs[1])"),
    4: dict(w=3840, h=2160, quality=75, method=6, target_psnr=42.0, batch=128, distinct=8, steps=3, e2e_workers=2, value_contexts=1, decode=False,
            metric="lossy encode Mpix/s (3840x2160 m6 TargetPSNR 42), bit-exact",
            workload="synthetic 3840x2160 RGBA lossy encode method 6, TargetPSNR 42 (three serial RD passes), batch of %d images per GPU (BASELINE configs[3])"),
    5: dict(w=256, h=256, quality=80, method=2, target_psnr=0.0, batch=2500, distinct=48, steps=8, e2e_workers=3, value_contexts=2, decode=False,
            metric="lossy encode Mpix/s (thumbnails 256x256 q80 m2), bit-exact",
            workload="thumbnails: synthetic 256x256 RGBA lossy encode q80 method 2, batches of %d images per GPU (BASELINE configs[4])"),
}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def committed_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed ncu --set full
    summary (profiles/r2_mode_search_traffic.json, written beside the .md summary of the same capture); None if absent."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_mode_search_traffic.json")))
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"]), t.get("launch")
    except Exception:
        return None, None


def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_cores(local, local_world):
    """Multi-rank runs: pin this rank (and every pinned buffer it first-touches afterwards) to its share of the host cores on the
    NUMA node its GPU hangs off, so that the D2H of one rank's NRGBA does not cross the socket interconnect and the ranks do not
    migrate over each other.  Falls back to an even split of the allowed cores when sysfs shows no NUMA placement."""
    info = {"mode": "none"}
    if local_world <= 1 or not hasattr(os, "sched_setaffinity"):
        return info
    allowed = sorted(os.sched_getaffinity(0))
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus  # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        info["pci"] = bus
        info["numa_node"] = node
    except Exception:
        node = -1
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except Exception:
        nodes = []
    info["numa_nodes"] = len(nodes)
    mine = None
    if node >= 0 and len(nodes) > 1:
        try:
            node_cpus = [c for c in _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read()) if c in allowed]
            # the ranks whose GPUs share this node split its cores; without knowing the others' placement assume an even spread
            per_node = max(1, local_world // len(nodes))
            slot = local % per_node
            share = max(1, len(node_cpus) // per_node)
            mine = node_cpus[slot * share:(slot + 1) * share] or node_cpus
            info["mode"] = "numa-local"
        except Exception:
            mine = None
    if mine is None:
        share = max(1, len(allowed) // local_world)
        mine = allowed[local * share:(local + 1) * share] or allowed
        info["mode"] = "even-split"
    try:
        os.sched_setaffinity(0, mine)
        info["cpus"] = len(mine)
    except Exception as e:
        info = {"mode": "none", "error": str(e)}
    return info


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


def oracle_cfg(cfg):
    import oracle_lib
    kw = dict(quality=cfg["quality"], method=cfg["method"])
    if cfg["target_psnr"] > 0:
        kw["target_psnr"] = cfg["target_psnr"]
    return oracle_lib.default_cfg(**kw)


def run_reference(args, cfg, rank):
    """CPU arm: the oracle port of the reference encoder, one image per thread on all host cores, bounded sample."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from webp_b200.synth import synth_batch
    W, H = cfg["w"], cfg["h"]
    cores = os.cpu_count() or 1
    # images per step: at least one per thread, sized for ~5-20 s per step on this workload
    sample = {2: max(cores, 8), 4: max(cores, 2), 5: max(16 * cores, 256)}[args.config]
    imgs = synth_batch(sample, W, H, distinct=min(sample, cfg["distinct"]))
    oc = oracle_cfg(cfg)
    for _ in range(1 if args.warmup else 0):
        oracle_lib.encode_batch(imgs[:min(sample, cores)], oc, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_lib.encode_batch(imgs, oc, threads=cores)
    dt = time.perf_counter() - t0
    v = sample * W * H * args.steps / dt / 1e6
    line = {"impl": "reference", "metric": cfg["metric"], "value": v, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": cfg["workload"] % args.batch, "sample_images_per_step": sample,
                       "note": "a step is a bounded sample of the workload (%d images, not %d): the rate is per pixel, so it compares; the "
                               "encoder is the reference's CPU path restated in C++ (oracle/, -O2), NOT the Go build -- no Go toolchain in this image" % (sample, args.batch)},
            "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port",
                             "sample": "%d images of the workload per step, one image per thread (C++ oracle, -O2)" % sample},
            "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 12 for config 2, 3 for config 4, 8 for config 5)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="workload: 2 = 1536x1024 q75 m4 (headline), 4 = 3840x2160 m6 TargetPSNR, 5 = 256x256 q80 m2 thumbnails")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the config's)")
    ap.add_argument("--e2e-workers", type=int, default=0, help="contexts per GPU used by the e2e leg: upload, device and finish stages of consecutive batches overlap")
    ap.add_argument("--value-contexts", type=int, default=0, help="contexts whose device-resident steps run side by side in the `value` leg")
    ap.add_argument("--gpu-slots", type=int, default=2, help="how many contexts may have their mode-search waves on the GPU at once")
    ap.add_argument("--finish-slots", type=int, default=0, help="how many contexts may be in their finish stage at once (default: 3 when the token partitions are coded on the GPU, else 1)")
    ap.add_argument("--decode-workers", type=int, default=0, help="contexts per GPU for the decode e2e leg (default: 8 with the GPU macroblock parser, whose latency per batch they hide; else the encode workers)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads per context (default: cores / ranks on this node)")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the benchmarked outputs (profiling runs only)")
    ap.add_argument("--no-affinity", action="store_true", help="multi-rank runs: do not pin the rank to the host cores next to its GPU")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.steps = args.steps or cfg["steps"]
    args.batch = args.batch or cfg["batch"]
    args.e2e_workers = args.e2e_workers or cfg["e2e_workers"]
    args.value_contexts = args.value_contexts or cfg["value_contexts"]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, cfg, rank)

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

    W, H = cfg["w"], cfg["h"]
    L = native.lib()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    affinity = {"mode": "none"} if args.no_affinity else bind_host_cores(local, local_world)  # before any pinned allocation
    host_threads = args.host_threads or max(1, (os.cpu_count() or 1) // max(1, local_world))
    ctx = native.Context(local, host_threads=host_threads)
    n, K = args.batch, args.steps
    px_step = n * W * H
    in_bytes = n * W * H * 4
    cap = W * H  # per-file capacity (bytes)
    opt = native.EncOptions()
    L.wgpu_enc_options_default(opt, cfg["quality"])
    opt.method = cfg["method"]
    if cfg["target_psnr"] > 0:
        opt.target_psnr = cfg["target_psnr"]
    do_search = cfg["target_psnr"] > 0

    # the library codes the token partitions on the GPU (webpgpu.cu device_coder_wanted; WGPU_DEVICE_CODER=0 forces the host coder)
    device_coder = os.environ.get("WGPU_DEVICE_CODER", "") != "0" and not do_search
    finish_slots = args.finish_slots or (3 if device_coder else 1)
    upload_stage, gpu_stage, host_stage = threading.Lock(), threading.BoundedSemaphore(max(1, args.gpu_slots)), threading.BoundedSemaphore(finish_slots)
    first_index = rank * cfg["distinct"]

    class Worker:
        """One wgpu_ctx + its own pinned input/output staging (contexts are independent by the ABI contract)."""

        def __init__(self, c):
            self.ctx = c
            self.h_in = L.wgpu_host_alloc(c.handle, in_bytes)
            self.h_out = L.wgpu_host_alloc(c.handle, n * cap)
            if not self.h_in or not self.h_out:
                raise SystemExit("pinned allocation failed")
            self.imgs = np.ctypeslib.as_array(C.cast(self.h_in, C.POINTER(C.c_uint8)), shape=(n, H, W, 4))
            self.imgs[:] = synth_batch(n, W, H, distinct=cfg["distinct"], first_index=first_index)
            self.out = np.ctypeslib.as_array(C.cast(self.h_out, C.POINTER(C.c_uint8)), shape=(n, cap))
            self.sizes = np.zeros(n, np.uint64)
            self.batches = 0

        def encode_e2e(self):
            """wgpu_encode_batch spelled as its three public stages so that the contexts pipeline."""
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
            self.batches += 1
            if os.environ.get("BENCH_TRACE"):  # wait-upload, upload, wait-gpu, device, wait-finish, finish (ms)
                sys.stderr.write("[bench] stages ms: " + " ".join("%.0f" % ((b - a) * 1e3) for a, b in zip(t, t[1:])) + "\n")

        def free(self):
            if self.h_in:
                L.wgpu_host_free(self.ctx.handle, self.h_in)
                L.wgpu_host_free(self.ctx.handle, self.h_out)
                self.h_in = self.h_out = None

    w0 = Worker(ctx)
    imgs, out, sizes, h_in = w0.imgs, w0.out, w0.sizes, w0.h_in
    workers = [w0]
    if args.e2e_workers > 1:
        workers += [Worker(native.Context(local, host_threads=host_threads)) for _ in range(args.e2e_workers - 1)]
    n_warm = max(args.warmup, 3) if not do_search else max(1, min(args.warmup, 3))
    for wk in workers:
        for _ in range(n_warm if wk is w0 else 1):
            wk.encode_e2e()
    clocks = ClockSampler(local)
    clocks.start()
    # ---- value: device-resident encode (inputs already in HBM)
    # K steps (batches) in all, dealt to `--value-contexts` contexts whose streams run side by side: the wave sequence of one
    # batch leaves the GPU partly empty at its narrow ends, a second sequence fills them.  Timed with CUDA events on each
    # context's stream from a common start; the slowest one counts.
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
            # the whole job from HBM-resident RGBA to WebP files: mode search, token generation and boolean coding of both
            # partitions on the device, the coded partitions laid out as files in (pinned) host memory
            wk.ctx.check(L.wgpu_enc_device(wk.ctx.handle, C.byref(opt)))
            wk.ctx.check(L.wgpu_enc_finish(wk.ctx.handle, wk.h_out, cap, wk.sizes.ctypes.data))
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
        wk.batches = 0
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
    # the one-shot public call, one context, nothing overlapped: what a caller of wgpu_encode_batch alone sees
    t1 = time.perf_counter()
    ctx.check(L.wgpu_encode_batch(ctx.handle, h_in, n, W, H, W * 4, W * H * 4, C.byref(opt), w0.h_out, cap, sizes.ctypes.data))
    one_shot_s = time.perf_counter() - t1
    # bytes actually copied inside the timed region, counted by the library at every cudaMemcpy*Async it issues
    xfer = [wk.ctx.transfer_bytes() for wk in workers]
    h2d = sum(x[0] for x in xfer) // (K + 1)
    d2h = sum(x[1] for x in xfer) // (K + 1)
    # ---- per-kernel device times for the roofline (CUDA events on the library's stream)
    stage_ms = {}
    for name, sid, reps in (("import", 0, 5), ("analysis", 1, 5), ("mode_search", 2, max(1, min(K, 3)))):
        if do_search and name == "mode_search":
            continue  # rate control: the pass loop is host driven; the value leg is its device time
        ctx.check(L.wgpu_enc_stage_time(ctx.handle, C.byref(opt), sid, reps, C.byref(ms)))
        stage_ms[name] = ms.value
    clock_rec = clocks.stop() if not cfg["decode"] or args.no_decode else None
    peak, peak_kind = measured_peak_gbs()
    result = {"metric": cfg["metric"], "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": K,
              "warmup": n_warm, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "u8/int32", "data": "synthetic",
              "config": {"workload": cfg["workload"] % n,
                         "batch_per_gpu": n, "l2": "inputs (%.0f MB RGBA per step) exceed the 126 MB L2" % (in_bytes / 1e6),
                         "parallelism": "images sharded across %d GPU(s), no collective" % world,
                         "batches_in_flight": len(vws),
                         "value_region": "wgpu_enc_device + wgpu_enc_finish per step (RGBA resident in HBM -> WebP files in pinned host memory)"},
              "e2e": {"value": e2e, "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / K * 1e3,
                      "compressed_bytes_per_step": int(sizes.sum()), "workers_per_gpu": len(workers), "host_threads_per_worker": host_threads,
                      "token_partition_coder": "gpu" if device_coder else "host", "finish_slots": finish_slots, "gpu_slots": args.gpu_slots,
                      "one_shot": {"value": px_step / one_shot_s / 1e6, "unit": "Mpix/s", "ms": one_shot_s * 1e3,
                                   "api": "one wgpu_encode_batch call on one context, nothing overlapped"},
                      "api": "wgpu_enc_upload + wgpu_enc_device + wgpu_enc_finish (== wgpu_encode_batch: pinned host RGBA in, WebP files out), %d batches dealt to %d contexts whose upload / device / finish stages overlap" % (K, len(workers))},
              "gpu_launches": int(launches)}
    # ---- decode of the streams just produced (BASELINE configs[2])
    dec_check = None
    if cfg["decode"] and not args.no_decode:
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
        # several contexts so that parse, device step and D2H of consecutive batches overlap
        env_parser = os.environ.get("WGPU_DEVICE_PARSER", "")
        device_parser = (env_parser != "0") if env_parser else n >= 32  # webpgpu.cu device_parser_wanted
        for wk in workers[1:]:  # the encode staging of the extra contexts is no longer needed
            wk.keep = [(int(i), wk.out[i, :int(wk.sizes[i])].tobytes(), wk.imgs[i].copy()) for i in _pick(wk.sizes)] if wk.batches else []
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
                self.batches = 0
        dworkers = [DW(c) for c in dctxs]
        for wk, buf in zip(dworkers[1:], dec_bufs[1:]):
            wk.ctx.check(L.wgpu_decode_batch(wk.ctx.handle, ptrs, lens, n, None, None, None, 0, 0, buf, W * H * 4))

        def decode_e2e(wk, buf):
            h = wk.ctx.handle
            wk.batches += 1
            if device_parser:
                # macroblocks are parsed on the GPU, one warp per image: a latency chain that occupies a few percent of the
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
                decode_e2e(dworkers[0], h_rgba)
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
        # the same leg at the lossy.DecodeFrame boundary (internal/lossy/decode.go:107-131: Y, U, V planes out, 1.5 B/px instead
        # of 4): what the e2e figure becomes when the D2H volume, the multi-GPU limiter (tools/d2h_scale_probe.py), shrinks
        mbw_, mbh_ = (W + 15) >> 4, (H + 15) >> 4
        ysz, uvsz = mbw_ * 16 * mbh_ * 16, mbw_ * 8 * mbh_ * 8

        def decode_planes_e2e(wk, buf):
            h = wk.ctx.handle
            wk.ctx.check(L.wgpu_dec_parse(h, ptrs, lens, n, None, None))
            wk.ctx.check(L.wgpu_dec_device(h, 0))
            wk.ctx.check(L.wgpu_sync(h))
            with gpu_stage:
                wk.ctx.check(L.wgpu_dec_fetch(h, buf, buf + n * ysz, buf + n * (ysz + uvsz), ysz, uvsz, None, 0))
        pcounter = iter(range(K))
        plock = threading.Lock()

        def prun(wk, buf):
            while True:
                with plock:
                    if next(pcounter, None) is None:
                        return
                decode_planes_e2e(wk, buf)
        barrier()
        t0 = time.perf_counter()
        ths = [threading.Thread(target=prun, args=(wk, buf)) for wk, buf in zip(dworkers, dec_bufs)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        barrier()
        dps = max_over_ranks(time.perf_counter() - t0)
        # leave every context with a full NRGBA decode behind it again (the parity check below reads the buffers)
        for wk, buf in zip(dworkers, dec_bufs):
            if wk.batches or wk is dworkers[0]:
                wk.ctx.check(L.wgpu_decode_batch(wk.ctx.handle, ptrs, lens, n, None, None, None, 0, 0, buf, W * H * 4))
        t1 = time.perf_counter()
        ctx.check(L.wgpu_decode_batch(ctx.handle, ptrs, lens, n, None, None, None, 0, 0, h_rgba, W * H * 4))
        dec_one_shot_s = time.perf_counter() - t1
        # the same call with the macroblocks parsed by the host's threads (WGPU_DEVICE_PARSER=0, read per call): one context alone on
        # a host with many cores finishes sooner that way; the GPU parser is the default because it needs no host cores and scales
        # with the number of contexts in flight
        saved_parser = os.environ.get("WGPU_DEVICE_PARSER")
        os.environ["WGPU_DEVICE_PARSER"] = "0"
        ctx.check(L.wgpu_decode_batch(ctx.handle, ptrs, lens, n, None, None, None, 0, 0, h_rgba, W * H * 4))  # warm-up: staging buffers
        t1 = time.perf_counter()
        ctx.check(L.wgpu_decode_batch(ctx.handle, ptrs, lens, n, None, None, None, 0, 0, h_rgba, W * H * 4))
        dec_one_shot_host_s = time.perf_counter() - t1
        if saved_parser is None:
            os.environ.pop("WGPU_DEVICE_PARSER", None)
        else:
            os.environ["WGPU_DEVICE_PARSER"] = saved_parser
        ctx.check(L.wgpu_decode_batch(ctx.handle, ptrs, lens, n, None, None, None, 0, 0, h_rgba, W * H * 4))  # back on the default route
        clock_rec = clocks.stop()
        # what the oracle comparison below looks at: NRGBA of every decode context that ran, planes of the first one
        mbw, mbh = (W + 15) >> 4, (H + 15) >> 4
        pick = _pick(sizes)
        dec_check = {"files": files, "pick": pick, "nrgba": [], "planes": None}
        for wk, buf in zip(dworkers, dec_bufs):
            if wk.batches or wk is dworkers[0]:
                a = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), shape=(n, H, W, 4))
                dec_check["nrgba"].append([a[i].copy() for i in pick])
        py = np.empty((n, mbh * 16, mbw * 16), np.uint8); pu = np.empty((n, mbh * 8, mbw * 8), np.uint8); pv = np.empty_like(pu)
        ctx.check(L.wgpu_dec_fetch(ctx.handle, py.ctypes.data, pu.ctypes.data, pv.ctypes.data, py[0].nbytes, pu[0].nbytes, None, 0))
        dec_check["planes"] = [(py[i].copy(), pu[i].copy(), pv[i].copy()) for i in pick]
        del py, pu, pv
        for wk, buf in zip(dworkers[1:], dec_bufs[1:]):
            L.wgpu_host_free(wk.ctx.handle, buf)
        dec_gbs = (ALG_BYTES_PER_PX["recon"] + ALG_BYTES_PER_PX["filter"] + ALG_BYTES_PER_PX["upsample"]) * px_step / (dms / K * 1e-3) / 1e9
        result["decode"] = {"value": px_step * K * world / (dms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": dms / K, "gpu_launches": int(dl),
                            "e2e": {"value": px_step * K * world / ds / 1e6, "unit": "Mpix/s", "ms_per_step": ds / K * 1e3,
                                    "h2d_bytes_per_step": sum(x[0] for x in dxfer) // K, "d2h_bytes_per_step": sum(x[1] for x in dxfer) // K,
                                    "macroblock_parser": "gpu" if device_parser else "host", "workers_per_gpu": len(dworkers),
                                    "planes_out": {"value": px_step * K * world / dps / 1e6, "unit": "Mpix/s", "ms_per_step": dps / K * 1e3,
                                                   "d2h_bytes_per_step": n * (ysz + 2 * uvsz),
                                                   "api": "same stages, Y/U/V planes out (the lossy.DecodeFrame boundary) instead of NRGBA"},
                                    "one_shot": {"value": px_step / dec_one_shot_s / 1e6, "unit": "Mpix/s", "ms": dec_one_shot_s * 1e3,
                                                 "api": "one wgpu_decode_batch call on one context, nothing overlapped",
                                                 "host_parser": {"value": px_step / dec_one_shot_host_s / 1e6, "unit": "Mpix/s", "ms": dec_one_shot_host_s * 1e3,
                                                                 "threads": host_threads, "how": "WGPU_DEVICE_PARSER=0"}}},
                            "config": {"workload": "decode of the %d streams above -> recon + loop filter + fancy upsampling to NRGBA (BASELINE configs[2])" % n},
                            "roofline": {"bound": "hbm", "kernel": "recon_wave + filter_wave + upsample_nrgba (whole device step)",
                                         "achieved": dec_gbs, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": dec_gbs / peak}}
        L.wgpu_host_free(ctx.handle, h_rgba)
    result["clocks"] = clock_rec
    if world > 1:
        result["host_affinity"] = affinity
    # ---- roofline of the dominant kernel (the mode search): integer issue, SURVEY.md 8(d).  Operations are COUNTED by the oracle
    # built with its per-stage counters (oracle/vp8_common.h OpStage) over the distinct images of this rank's batch, weighted by how
    # often each one occurs in it; peak = SMs x 128 lanes x the SM clock observed during the timed region.
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oc = oracle_cfg(cfg)
    sm_mhz = (clock_rec or {}).get("sm_mhz") or 1965.0
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    peak_ops = sm_count * 128 * sm_mhz * 1e6
    nmb = ((W + 15) // 16) * ((H + 15) // 16)
    ops_step, ops_by_stage = None, None
    if not do_search or os.environ.get("BENCH_COUNT_OPS"):
        d = min(n, cfg["distinct"])
        mult = [len(range(i, n, d)) for i in range(d)]
        with cf.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 1, d)) as ex:
            counted = list(ex.map(lambda i: oracle_lib.encode_ops(imgs[i], oc), range(d)))
        ops_by_stage = {}
        for m, c in zip(mult, counted):
            for k, v in c.items():
                ops_by_stage[k] = ops_by_stage.get(k, 0) + m * v
        ops_step = float(sum(ops_by_stage.values()))
    traffic, traffic_launch = committed_traffic() if args.config == 2 else (None, None)
    t_two = dev_ms / K * 1e-3                      # per step with `batches_in_flight` sequences side by side (the value leg; import + analysis inside)
    t_one = stage_ms.get("mode_search", 0.0) * 1e-3  # one batch's wave sequence alone
    roofline = {"bound": "int_issue", "kernel": "encode_phased_kernel (one launch per wave, %d per step)" % ((W + 15) // 16 + 2 * ((H + 15) // 16 - 1))
                if cfg["method"] >= 3 and not do_search else ("encode_serial_luma_wave_kernel + encode_serial_chroma_chain_kernel (serial RD passes, split by plane)" if do_search else "encode_fast_wave_kernel"),
                "unit": "Tiop/s", "peak": peak_ops / 1e12, "peak_kind": "%d SMs x 128 lanes x %.0f MHz observed" % (sm_count, sm_mhz),
                "ops_per_macroblock": ops_step / (n * nmb) if ops_step else None,
                "ops_by_stage_per_macroblock": {k: round(v / (n * nmb), 1) for k, v in ops_by_stage.items()} if ops_by_stage else None,
                "traffic": traffic, "traffic_launch": traffic_launch,
                "algorithmic_bytes_per_step": ALG_BYTES_PER_PX["mode_search"] * px_step,
                "stages_ms": stage_ms,
                "stages_gbs": {k: ALG_BYTES_PER_PX[k] * px_step / (v * 1e-3) / 1e9 for k, v in stage_ms.items()},
                "hbm_peak_gbs": peak, "hbm_peak_kind": peak_kind}
    if ops_step and t_one > 0:
        # the kernel's own time: CUDA events around the wave launches of one batch (wgpu_enc_stage_time), nothing else on the GPU
        roofline["achieved"] = ops_step / t_one / 1e12
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
        roofline["kernel_ms_per_step"] = t_one * 1e3
        roofline["avg_launch_us"] = t_one * 1e6 / ((W + 15) // 16 + 2 * ((H + 15) // 16 - 1))
        # for scale: the same operations over the whole step of the `value` leg (every other kernel, the coder and the file layout inside)
        roofline["value_leg"] = {"achieved": ops_step / t_two / 1e12, "frac": ops_step / t_two / peak_ops, "ms_per_step": t_two * 1e3,
                                 "batches_in_flight": len(vws)}
    else:
        roofline["achieved"] = roofline["frac"] = None
        roofline["note"] = "operations not counted for the rate-controlled workload by default (BENCH_COUNT_OPS=1 counts them: three CPU passes per image)"
    result["roofline"] = roofline
    # ---- parity of what was benchmarked: files of every encode context, decoded planes / NRGBA of the decode contexts
    checked = 0
    if not args.no_parity:
        jobs = []
        w0.keep = [(int(i), out[i, :int(sizes[i])].tobytes(), imgs[i].copy()) for i in _pick(sizes)]
        for wi, wk in enumerate(workers):
            if not hasattr(wk, "keep"):
                wk.keep = [(int(i), wk.out[i, :int(wk.sizes[i])].tobytes(), wk.imgs[i].copy()) for i in _pick(wk.sizes)] if wk.batches or wk is w0 else []
            jobs += [(wi, i, f, im) for (i, f, im) in wk.keep]
        with cf.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 1, max(1, len(jobs)))) as ex:
            exp = list(ex.map(lambda j: oracle_lib.encode(j[3], oc), jobs))
        for (wi, i, f, _), e in zip(jobs, exp):
            if f != e:
                raise SystemExit("PARITY FAILURE: context %d image %d: %d bytes vs oracle %d bytes" % (wi, i, len(f), len(e)))
            checked += 1
        if dec_check:
            for k, i in enumerate(dec_check["pick"]):
                _, _, ey, eu, ev = oracle_lib.decode(dec_check["files"][i])
                gy, gu, gv = dec_check["planes"][k]
                if not (np.array_equal(gy, ey) and np.array_equal(gu, eu) and np.array_equal(gv, ev)):
                    raise SystemExit("PARITY FAILURE: decoded planes of image %d differ from the oracle" % i)
                en = oracle_lib.build_nrgba(W, H, ey, eu, ev)
                for ci, per_ctx in enumerate(dec_check["nrgba"]):
                    if not np.array_equal(per_ctx[k], en):
                        raise SystemExit("PARITY FAILURE: NRGBA of image %d (decode context %d) differs from the oracle" % (i, ci))
                    checked += 1
                checked += 1
    result["parity_checked"] = checked
    # ---- CPU baseline beside it: the oracle port on this box's host cores, bounded sample, rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = {2: max(cores, 8), 4: max(cores, 2), 5: max(16 * cores, 256)}[args.config]
        sub = np.ascontiguousarray(imgs[:sample] if sample <= n else synth_batch(sample, W, H, distinct=cfg["distinct"]))
        t0 = time.perf_counter()
        oracle_lib.encode_batch(sub, oc, threads=cores)
        dt = time.perf_counter() - t0
        result["cpu_baseline"] = {"value": sample * W * H / dt / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
                                  "sample": "%d images of the same batch, one image per thread, %d threads (C++ oracle -O2; the Go reference cannot be built here)" % (sample, cores)}
    for wk in workers:
        wk.free()
    if rank == 0:
        print(json.dumps(result), flush=True)
    if dist:
        dist.destroy_process_group()


def _pick(sizes):
    """Images of a batch the parity check looks at: first, last, the one with the longest stream, the median-sized one."""
    s = np.asarray(sizes).astype(np.int64)
    order = np.argsort(s, kind="stable")
    return sorted({0, len(s) - 1, int(order[-1]), int(order[len(s) // 2])})


if __name__ == "__main__":
    main()
