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
# overlapped exchange: the stage plan + StageSync's issue / rest / wait on a fake two-group arena (CPU, gloo)
from types import SimpleNamespace
names = ["bert.embeddings.word_embeddings.weight", "bert.lang_encoder.layer.0.attention.self.query.weight",
         "bert.lang_encoder.layer.1.attention.self.query.weight", "bert.local_encoder.vp_pos_embeddings.0.weight",
         "bert.local_encoder.encoder.crossattention.0.attention.self.query.weight",
         "bert.global_encoder.gmap_step_embeddings.weight",
         "bert.global_encoder.encoder.crossattention.0.attention.self.query.weight", "bert.txt_emb_w.weight",
         "global_sap_head.net.0.weight"]
entries, off = [], 0
for i, n in enumerate(names):
    k = 37 + 11 * i
    entries.append((n, None, off, k))
    off += (k + 7) // 8 * 8
n_decay = off
entries.append(("bert.lang_encoder.layer.1.attention.self.query.bias", None, off, 5))
off += 8
torch.manual_seed(200 + rank)
arena = SimpleNamespace(entries=entries, n_decay=n_decay, total=off, flat_g=torch.randn(off))
before = [torch.zeros(off) for _ in range(world)]
dist.all_gather(before, arena.flat_g.clone())
sy = parallel.StageSync(arena, 2, bucket_bytes=64 * 4)
st0 = [entries[i][2] for i in (4, 6, 7, 8)]
assert [lo for lo, hi in sy.stages[0]] == [entries[4][2], entries[6][2]] and sy.stages[1] == [(entries[2][2], entries[3][2])]
cover = sorted([r for st in sy.stages for r in st] + sy.rest)
assert parallel.text_stage_cuts(9) == [6, 3] and parallel.text_stage_cuts(6) == [4, 2] and parallel.text_stage_cuts(1) == []
assert cover[0][0] == 0 and cover[-1][1] == off and all(a[1] == b[0] for a, b in zip(cover[:-1], cover[1:]))
sy.begin("eager")
sy.expect(0, "txt"); sy.expect(0, "g_in"); sy.expect(1, "txt_mid")
fired = []
sy._ready = lambda st: (fired.append(st), sy._issue(st))
sy.fire(0, "txt"); assert fired == []
sy.fire(0, "g_in"); assert fired == [0]
sy.fire(0, "g_in"); assert fired == [0]          # a stage starts once
sy.issue_rest(sy.fired)                           # stage 1 never completed: its range travels with the rest
sy.wait()
assert torch.allclose(arena.flat_g, (before[0] + before[1]) / 2, atol=1e-6)
# validation metrics are reduced like the reference does (train_r2r_magic.py:456-458: sum(all_gather(x)))
from magic_b200.train_loop import all_gather
assert all_gather(rank * 1.5 + 1) == [1.0, 2.5] and all_gather({"n": rank}) == [{"n": 0}, {"n": 1}]
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
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--workload",
                        "magic_s_pretrain_b64"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    import json
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
