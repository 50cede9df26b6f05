"""World-size-2 gloo test (CPU) of the data-parallel plumbing: flat-arena all-reduce in buckets, parameter
broadcast, seeded task schedule, sample sharding, and the `--impl reference` rank gating of bench.py."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import magic_b200
from magic_b200 import parallel
rank, world, local = parallel.init_distributed()
assert world == 2 and dist.get_backend() == "gloo"
torch.manual_seed(100 + rank)
flat = torch.randn(10007)
mine = flat.clone()
other = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(other, mine)
ar = parallel.FlatAllReduce(flat, bucket_bytes=4096 * 4)
assert len(ar.bounds) == 3 and ar.bounds[0][1] == 10007 and ar.bounds[-1][0] == 0
ar()
assert torch.allclose(flat, (other[0] + other[1]) / 2, atol=1e-6)
p = torch.full((64,), float(rank))
parallel.broadcast_flat(p, 0)
assert float(p.abs().max()) == 0.0
sched = parallel.task_schedule(1234, 50, ["mlm", "sap", "cfp"], [1, 1, 1])
objs = [None, None]
dist.all_gather_object(objs, sched)
assert objs[0] == objs[1] and set(sched) == {"mlm", "sap", "cfp"}
idx = parallel.shard_indices(101, rank, world, seed=3)
both = [None, None]
dist.all_gather_object(both, idx)
assert len(both[0]) == len(both[1]) == 51 and set(both[0]) | set(both[1]) == set(range(101))
dist.barrier()
sys.stdout.write("rank " + str(rank) + " ok\n")
sys.stdout.flush()
''' % ROOT


def test_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2 and "rank 0" in r.stdout and "rank 1" in r.stdout, r.stdout


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29732", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    import json
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
