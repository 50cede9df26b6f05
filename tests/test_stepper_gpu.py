"""GPU tests of the training step driver (train_step.PretrainStepper): CUDA-graph replay, concurrent stream
branches, the weight-gradient side stream and the host->device prefetch must all reproduce the plain eager,
single-stream step on the same inputs (fp32 mode, dropout off; fp32 atomics reorder sums, hence 2e-5)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

import magic_b200  # noqa: E402
from magic_b200 import synth  # noqa: E402
from magic_b200.config import make_config  # noqa: E402
from magic_b200.graph_index import prepare_batch, batch_to_device, pad_batch  # noqa: E402
from magic_b200.train_step import PretrainStepper  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def models(teacher, dtype=torch.float32, tasks=("mlm", "sap")):
    cfg_s = make_config(128, role="student", teacher_hidden_size=256 if teacher else None, hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, pretrain_tasks=tasks)
    torch.manual_seed(1)
    s = magic_b200.GlocalTextPathCMTPreTraining(cfg_s).to(DEV).train().set_compute_dtype(dtype)
    t = None
    if teacher:
        cfg_t = make_config(256, role="teacher", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                            pretrain_tasks=tasks)
        torch.manual_seed(0)
        t = magic_b200.GlocalTextPathCMTPreTraining(cfg_t).to(DEV).eval().set_compute_dtype(dtype)
    return s, t


def host_pool(n=3, B=4, tasks=("mlm", "sap")):
    pools = {}
    for ti, task in enumerate(tasks):
        bs = [prepare_batch(synth.make_batch(task, B, seed=50 + i + 100 * ti)) for i in range(n)]
        K = magic_b200.INDEX_KEY
        rcap = (max(b["traj_view_img_fts"].shape[0] for b in bs) + 7) // 8 * 8
        mcap = (max(b[K]["mlm_rows"].numel() for b in bs) + 63) // 64 * 64 if task == "mlm" else None
        ecap = (max(b[K]["entries"].numel() for b in bs) + 255) // 256 * 256
        scap = (max(b[K]["src_ids"].numel() for b in bs) + 255) // 256 * 256
        pools[task] = [pad_batch(b, rcap, mcap, ecap, scap) for b in bs]
    return pools


def run(teacher, n_steps=4, prefetch=False, lookahead=False, order=None, **kw):
    """`lookahead`: announce the next batch to step() (frozen-teacher pipelining: its teacher forward runs under the
    current step's backward).  `order`: task sequence (default MLM / SAP alternating)."""
    order = order or ["mlm" if i % 2 == 0 else "sap" for i in range(n_steps)]
    tasks = tuple(t_ for t_ in ("mlm", "sap", "cfp") if t_ in order or t_ != "cfp")
    s, t = models(teacher, tasks=tasks)
    g = torch.Generator().manual_seed(7)
    st = PretrainStepper(s, t, lr=1e-3, rw_generator=g, **kw)
    pools = host_pool(tasks=tasks)
    seen = {t_: 0 for t_ in tasks}

    def stage(task):
        hb = pools[task][seen[task] % len(pools[task])]
        seen[task] += 1
        if prefetch:
            pinned = {k: (v.pin_memory() if torch.is_tensor(v) else
                          ({kk: (vv.pin_memory() if torch.is_tensor(vv) else vv) for kk, vv in v.items()}
                           if k == magic_b200.INDEX_KEY else v)) for k, v in hb.items()}
            return st.prefetch(task, pinned)
        return batch_to_device(hb, DEV)

    losses = []
    nxt = stage(order[0])
    for i, task in enumerate(order):
        b = nxt
        nxt = stage(order[i + 1]) if i + 1 < len(order) else None
        if lookahead and nxt is not None:
            torch.cuda.synchronize()  # a batch announced as `next` must be resident
            losses.append(st.step(task, b, next=(order[i + 1], nxt)).clone())
        else:
            losses.append(st.step(task, b).clone())
    torch.cuda.synchronize()
    run.last_stepper = st
    return torch.stack(losses).cpu(), st.arena.flat_p.clone().cpu()


@pytest.mark.parametrize("teacher", [False, True])
def test_graph_streams_prefetch_match_plain_eager(teacher):
    base_l, base_p = run(teacher, use_graphs=False, side_stream=False, branch_streams=False)
    assert torch.isfinite(base_l).all()
    if teacher:
        assert (base_l[:, 2] > 0).all()  # the KD term is live
    for kw in (dict(use_graphs=False, side_stream=True, branch_streams=True),
               dict(use_graphs=True, side_stream=True, branch_streams=True),
               dict(use_graphs=True, side_stream=True, branch_streams=True, prefetch=True),
               dict(use_graphs=False, side_stream=False, branch_streams=False, prefetch=True)):
        l, p = run(teacher, **kw)
        assert rel(l, base_l) < 2e-5, (kw, l, base_l)
        assert rel(p, base_p) < 2e-5, kw


def test_teacher_pipelining_matches_serial_order():
    """Frozen teacher + graphs: the teacher forward of the announced next batch runs on its own stream under the
    current step's backward.  Same losses and parameters as the plain eager step -- alternating tasks (teacher output
    buffers of the two tasks are disjoint) and a sequence with repeated tasks (the next teacher forward must wait
    for the step that still reads those buffers), with resident batches and with prefetched ones."""
    for order in (None, ["sap", "sap", "mlm", "mlm", "mlm", "sap"]):
        n = 6
        base_l, base_p = run(True, n_steps=n, order=order, use_graphs=False, side_stream=False, branch_streams=False)
        for kw in (dict(lookahead=True), dict(lookahead=True, prefetch=True), dict(lookahead=False),
                   dict(lookahead=True, pipeline_teacher=False)):
            l, p = run(True, n_steps=n, order=order, use_graphs=True, **kw)
            assert rel(l, base_l) < 2e-5, (order, kw, l, base_l)
            assert rel(p, base_p) < 2e-5, (order, kw)


def test_graph_replay_follows_per_step_mkrw_draw():
    """The MKRW weights live in device memory read by the captured kernels: two replays of the same graph on the
    same batch with different draws must give different KD totals, equal to the eager values for those draws."""
    s, t = models(True)
    pools = host_pool(n=1)
    b = batch_to_device(pools["sap"][0], DEV)
    outs = {}
    for graphs in (False, True):
        s2, t2 = copy.deepcopy(s), copy.deepcopy(t)
        st = PretrainStepper(s2, t2, lr=0.0, weight_decay=0.0, rw_generator=torch.Generator().manual_seed(3),
                             use_graphs=graphs)
        outs[graphs] = torch.stack([st.step("sap", b).clone() for _ in range(3)]).cpu()
    assert rel(outs[True], outs[False]) < 2e-5
    kd = outs[True][:, 2]
    assert (kd[0] - kd[1]).abs() > 1e-4 * kd[0].abs() and (kd[1] - kd[2]).abs() > 1e-4 * kd[1].abs()


def test_icod_co_update_steps_both_models_and_graph_replay_matches_eager():
    """ICoD (co_update=True): one step moves the parameters of BOTH models; the CUDA-graph replay of the two-model
    step (teacher forward with grad, two optimizers) reproduces the eager step."""
    res = {}
    for graphs in (False, True):
        s, t = models(True)
        t.train()
        p0_s, p0_t = None, None
        st = PretrainStepper(s, t, lr=1e-3, rw_generator=torch.Generator().manual_seed(11), use_graphs=graphs,
                             side_stream=graphs, branch_streams=graphs, co_update=True)
        p0_s, p0_t = st.arena.flat_p.clone(), st.t_arena.flat_p.clone()
        pools = host_pool(n=2)
        outs = []
        for i in range(4):
            task = "mlm" if i % 2 == 0 else "sap"
            outs.append(st.step(task, batch_to_device(pools[task][(i // 2) % 2], DEV)).clone())
        torch.cuda.synchronize()
        outs = torch.stack(outs).cpu()
        assert outs.shape == (4, 6) and torch.isfinite(outs).all()
        assert (outs[:, 2] > 0).all() and (outs[:, 5] > 0).all()  # both KD directions are live
        assert (st.arena.flat_p - p0_s).abs().max() > 0 and (st.t_arena.flat_p - p0_t).abs().max() > 0
        res[graphs] = (outs, st.arena.flat_p.clone().cpu(), st.t_arena.flat_p.clone().cpu())
    for a, b in zip(res[True], res[False]):
        assert rel(a, b) < 2e-5


def test_tasks_without_makd_take_the_supervised_step_beside_a_teacher():
    """The reference's own task list is [mlm, sap, cfp] with distillation on (config/r2r_magic_pretrain.json:49-53,
    :62-64).  MAKD is defined for the MLM and SAP steps; a cfp step of the same run is the plain
    supervised step -- KD term exactly 0, no teacher forward -- and graph replay + teacher pipelining across such a
    step reproduce the eager order.  `tasks=` switches the task-aware AdamW on, as the loop does."""
    order = ["cfp", "mlm", "cfp", "sap", "mlm", "cfp"]
    tasks = ("mlm", "sap", "cfp")
    base_l, base_p = run(True, order=order, use_graphs=False, side_stream=False, branch_streams=False, tasks=tasks)
    assert torch.isfinite(base_l).all()
    is_cfp = torch.tensor([t == "cfp" for t in order])
    assert (base_l[is_cfp, 2] == 0).all() and (base_l[~is_cfp, 2] > 0).all()
    assert torch.allclose(base_l[is_cfp, 0], base_l[is_cfp, 1], rtol=1e-6)  # total == supervised mean
    assert run.last_stepper.last_named_losses() is None            # the last step (cfp) carried no KD terms
    l, p = run(True, order=order, use_graphs=True, side_stream=True, branch_streams=True, lookahead=True, tasks=tasks)
    assert rel(l, base_l) < 2e-5, (l, base_l)
    assert rel(p, base_p) < 2e-5
