"""torchrun --nproc-per-node 2 scripts/overlap_diag.py : which arena ranges disagree across ranks after the overlapped
gradient exchange (before the optimizer)?  Eager and graph mode, with and without helper streams."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import magic_b200  # noqa: E402
from magic_b200 import ops  # noqa: E402
from magic_b200.graph_index import flatten_batch  # noqa: E402
from magic_b200.parallel import init_distributed  # noqa: E402
from magic_b200.train_step import PretrainStepper  # noqa: E402

rank, world, local = init_distributed()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
name = sys.argv[1] if len(sys.argv) > 1 else "magic_s_pretrain_b64"
w = dict(bench.WORKLOADS[name])
w["B"] = min(w["B"], 16)
pools = {t: [flatten_batch(b, device=dev) for b in bench.make_pool(t, 2, w, 77 + rank * 100 + (0 if t == "mlm" else 50))]
         for t in ("mlm", "sap")}


def check(st, tag):
    """Compare every parameter's gradient across the two ranks."""
    bad = []
    for arena, nm in ((st.arena, "student"), (st.t_arena if st.co_update else None, "teacher")):
        if arena is None:
            continue
        mine = arena.flat_g.clone()
        other = mine.clone()
        dist.broadcast(other, 0)
        diff = (mine - other).abs()
        if rank == 1:
            for n, p, o, k in arena.entries:
                d = float(diff[o:o + k].max())
                if d > 0:
                    bad.append((nm, n, d, float(mine[o:o + k].abs().max())))
    out = [None, None]
    dist.all_gather_object(out, bad)
    if rank == 0:
        b = out[1]
        print(f"  [{tag}] parameters whose gradient differs across ranks after the exchange: {len(b)}")
        for x in b[:12]:
            print("      ", x)


def run(overlap, graphs, side, branch, steps=3):
    cfg_s, cfg_t = bench.make_cfgs(w, 0.1)
    torch.manual_seed(1)
    student = magic_b200.GlocalTextPathCMTPreTraining(cfg_s).to(dev).train().set_compute_dtype(torch.bfloat16)
    teacher = None
    if cfg_t is not None:
        torch.manual_seed(0)
        teacher = magic_b200.GlocalTextPathCMTPreTraining(cfg_t).to(dev).set_compute_dtype(torch.bfloat16)
        teacher = teacher.train() if w.get("co_update") else teacher.eval()
    st = PretrainStepper(student, teacher, use_graphs=graphs, co_update=bool(w.get("co_update")), overlap=overlap,
                         rw_generator=torch.Generator().manual_seed(5), side_stream=side, branch_streams=branch)
    ops.set_seed(dev, 999)
    orig = st.opt.apply
    state = {"i": 0}

    def apply():
        torch.cuda.synchronize()
        check(st, f"overlap={overlap} graphs={graphs} side={side} branch={branch} step {state['i']} "
                  f"fired={[sy.fired for sy in st.syncs]}")
        state["i"] += 1
        orig()

    if not graphs or world > 1:
        st.opt.apply = apply
    for i in range(steps):
        task = "mlm" if i % 2 == 0 else "sap"
        st.step(task, pools[task][(i // 2) % 2])
    torch.cuda.synchronize()
    for a in (st.arena, st.t_arena):
        if a is not None:
            a.release()


for graphs in (False, True):
    for side, branch in ((True, True), (False, False)):
        for overlap in (False, True):
            run(overlap, graphs, side, branch)
dist.barrier()
dist.destroy_process_group()
