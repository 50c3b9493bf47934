"""CPU suite: the N>1 path (images sharded across ranks, no data-path collective; barrier + max-over-ranks timing)
exercised with world_size 2 on gloo."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import numpy as np, torch, torch.distributed as dist
    import oracle_lib
    from webp_b200.synth import synth_batch
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    per_rank = 3
    # bench.py's sharding rule: rank r works on images [r*per_rank, (r+1)*per_rank) of the global synthetic sequence
    mine = synth_batch(per_rank, 64, 64, distinct=per_rank, first_index=rank * per_rank)
    sizes = [len(oracle_lib.encode(im)) for im in mine]
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)   # stand-in for this rank's elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, sizes)
    if rank == 0:
        print(json.dumps({"max_t": t.item(), "sizes": gathered}))
    dist.destroy_process_group()
""") % (ROOT, ROOT)


def test_world2_sharding_and_max_reduce(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["max_t"] == 2.0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from webp_b200.synth import synth_batch
    whole = synth_batch(6, 64, 64, distinct=6, first_index=0)
    expect = [len(oracle_lib.encode(im)) for im in whole]
    assert r["sizes"][0] + r["sizes"][1] == expect  # disjoint shards, together the whole job, no exchange needed
