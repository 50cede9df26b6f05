"""Diagnostic: per-step loss / parameter drift of stepper variants against the plain eager single-stream step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import magic_b200
from magic_b200.graph_index import batch_to_device
from magic_b200.train_step import PretrainStepper
import test_stepper_gpu as T


def run(teacher, n_steps, prefetch, sync, **kw):
    s, t = T.models(teacher)
    st = PretrainStepper(s, t, lr=1e-3, rw_generator=torch.Generator().manual_seed(7), **kw)
    pools = T.host_pool()
    out = []
    for i in range(n_steps):
        task = "mlm" if i % 2 == 0 else "sap"
        hb = pools[task][(i // 2) % len(pools[task])]
        if prefetch:
            pinned = {k: (v.pin_memory() if torch.is_tensor(v) else
                          ({kk: (vv.pin_memory() if torch.is_tensor(vv) else vv) for kk, vv in v.items()}
                           if k == magic_b200.INDEX_KEY else v)) for k, v in hb.items()}
            b = st.prefetch(task, pinned)
        else:
            b = batch_to_device(hb, "cuda")
        l = st.step(task, b).clone()
        if sync:
            torch.cuda.synchronize()
        out.append((l, st.arena.flat_p.clone()))
    torch.cuda.synchronize()
    return out


base = run(False, 6, False, True, use_graphs=False, side_stream=False, branch_streams=False)
variants = {
    "eager+streams": dict(prefetch=False, kw=dict(use_graphs=False, side_stream=True, branch_streams=True)),
    "graph": dict(prefetch=False, kw=dict(use_graphs=True, side_stream=True, branch_streams=True)),
    "graph nostreams": dict(prefetch=False, kw=dict(use_graphs=True, side_stream=False, branch_streams=False)),
    "graph+prefetch": dict(prefetch=True, kw=dict(use_graphs=True, side_stream=True, branch_streams=True)),
    "graph+prefetch nostreams": dict(prefetch=True, kw=dict(use_graphs=True, side_stream=False, branch_streams=False)),
    "eager+prefetch": dict(prefetch=True, kw=dict(use_graphs=False, side_stream=False, branch_streams=False)),
}
for name, v in variants.items():
    for sync in (True, False):
        for rep in range(2):
            r = run(False, 6, v["prefetch"], sync, **v["kw"])
            dl = [abs(a[0][0].item() - b[0][0].item()) / abs(b[0][0].item()) for a, b in zip(r, base)]
            dp = [((a[1] - b[1]).norm() / b[1].norm()).item() for a, b in zip(r, base)]
            print(f"{name:26s} sync={int(sync)} rep{rep} dloss " + " ".join(f"{x:.1e}" for x in dl) + "  dparam " +
                  " ".join(f"{x:.1e}" for x in dp), flush=True)
