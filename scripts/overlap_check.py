"""torchrun --nproc-per-node 2 scripts/overlap_check.py [workload]
Data-parallel step with the gradient exchange overlapped with backward (parallel.StageSync) vs the plain
exchange-after-backward path: same seeds, same batches, N steps each, eager and CUDA-graph mode.  At world size 2 the
all-reduce of two values is order independent: after the exchange every rank must hold BIT-IDENTICAL gradients (checked
right before each optimizer step) and, with the deterministic grad-norm reduction, bit-identical parameters after the
steps.  The two paths are compared with each other within the run-to-run noise of the fp32 atomics in backward."""
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


def run(overlap, graphs, steps=4):
    cfg_s, cfg_t = bench.make_cfgs(w, 0.1)
    torch.manual_seed(1)
    student = magic_b200.GlocalTextPathCMTPreTraining(cfg_s).to(dev).train().set_compute_dtype(torch.bfloat16)
    teacher = None
    if cfg_t is not None:
        torch.manual_seed(0)
        teacher = magic_b200.GlocalTextPathCMTPreTraining(cfg_t).to(dev).set_compute_dtype(torch.bfloat16)
        teacher = teacher.train() if w.get("co_update") else teacher.eval()
    st = PretrainStepper(student, teacher, use_graphs=graphs, co_update=bool(w.get("co_update")), overlap=overlap,
                         rw_generator=torch.Generator().manual_seed(5), tasks=("mlm", "sap"))
    ops.set_seed(dev, 999)
    orig, bad = st.opt.apply, []

    def apply(task=None):  # the exchange has been waited for: compare the gradient arenas across the ranks
        for a in (st.arena, st.t_arena if st.co_update else None):
            if a is not None:
                other = a.flat_g.clone()
                dist.broadcast(other, 0)
                bad.append(float((other - a.flat_g).abs().max()))
        orig(task)

    if not graphs:  # (graph mode: the warm-up steps before a capture run without the exchange)
        st.opt.apply = apply
    for i in range(steps):
        task = "mlm" if i % 2 == 0 else "sap"
        out = st.step(task, pools[task][(i // 2) % 2])
    torch.cuda.synchronize()
    res = [st.arena.flat_p.clone()] + ([st.t_arena.flat_p.clone()] if st.co_update else [])
    desc = st.exchange_description()
    assert all(b == 0.0 for b in bad), f"gradients differ across ranks after the exchange: {bad}"
    for a in (st.arena, st.t_arena):
        if a is not None:
            a.release()
    return res, out.clone(), desc


for graphs in (False, True):
    base, out0, d0 = run(False, graphs)
    base2, _, _ = run(False, graphs)          # the plain path against itself: the noise floor of this comparison
    over, out1, d1 = run("force", graphs)
    md = max(float((a - b).abs().max()) for a, b in zip(base, over))
    noise = max(float((a - b).abs().max()) for a, b in zip(base, base2))
    gathered = [None] * world
    dist.all_gather_object(gathered, (md, noise))
    # replicas stay in sync: every rank holds the same parameters
    insync = []
    for t in over:
        ref = t.clone()
        dist.broadcast(ref, 0)
        insync.append(bool(torch.equal(ref, t)))
    sync_all = [None] * world
    dist.all_gather_object(sync_all, all(insync))
    if rank == 0:
        print(f"[{name}] graphs={graphs}: max |param diff| overlapped vs plain {max(g[0] for g in gathered):.3e} "
              f"(plain vs plain, run to run: {max(g[1] for g in gathered):.3e}); replicas bit-identical after the steps: "
              f"{all(sync_all)}; loss {out0.tolist()} vs {out1.tolist()}")
        print("   ", d1)
    assert all(sync_all), "ranks diverged"
dist.barrier()
if rank == 0:
    print("overlap_check done")
dist.destroy_process_group()
