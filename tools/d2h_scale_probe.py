"""Host-side limiter of the multi-GPU decode e2e leg: aggregate pinned D2H bandwidth of all ranks at once, with and without the
NUMA-local core binding bench.py applies (launch under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/d2h_scale_probe.py
Rank 0 prints the topology (nvidia-smi topo -m, NUMA nodes) and one line per phase."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NBYTES = 256 * 1536 * 1024 * 4  # one decode batch of NRGBA


def phase(name, solo=False):
    dev = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")
    host = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
    host.fill_(1)  # first touch under the current affinity
    up = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")
    for direction in ("d2h", "h2d"):
        for who in ((0,) if solo else (None,)):
            active = who is None or rank == who
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            if active:
                for _ in range(6):
                    if direction == "d2h":
                        host.copy_(dev, non_blocking=True)
                    else:
                        up.copy_(host, non_blocking=True)
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ranks = 1 if solo else world
            if rank == 0:
                print("%-34s %s: %d rank(s) x 6 x %.2f GB in %.3f s -> %.1f GB/s aggregate, %.1f GB/s per GPU" %
                      (name, direction, ranks, NBYTES / 1e9, t.item(), ranks * 6 * NBYTES / t.item() / 1e9, 6 * NBYTES / t.item() / 1e9), flush=True)
    del dev, host, up


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
            if cmd[0] == "lscpu":
                out = "\n".join(l for l in out.splitlines() if any(k in l for k in ("NUMA", "Socket", "Core", "CPU(s):", "Model name")))
            print(out, flush=True)
        except Exception as e:
            print(cmd, "failed:", e)
    print("allowed cores:", len(os.sched_getaffinity(0)), flush=True)
phase("one rank alone, no binding", solo=True)
phase("all ranks, no binding")
info = bench.bind_host_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
gathered = [None] * world
dist.all_gather_object(gathered, info)
if rank == 0:
    print("binding:", gathered, flush=True)
phase("all ranks, bound next to the GPU")
dist.barrier()
dist.destroy_process_group()
