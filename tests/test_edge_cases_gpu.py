"""Edge shapes of the hot path against the fp32 oracle: a batch of ONE sample, single-step trajectories (the graph
is [stop] + one visited node + its candidates), 4-token instructions (2-token ones inside the batch), a graph capped
at 4 nodes, and all of them at once -- forward + MAKD distillation losses + backward, in fp32 mode (FFMA kernels,
1e-4) and then in bf16 mode (tcgen05 / mma.sync kernels, 2e-2) on the same weights.  These are the ragged / minimum
sizes of the reference's batch schema (pretrain_src/data/tasks.py:110-166, 392-451); the maximum sizes (L = 160,
G = 50, T = 12, B = 128) are in test_model_gpu.py / test_bf16_parity_gpu.py.

`test_edge_batches_on_the_oracle` is the CPU half: the oracle and the host index builder accept the same batches and
keep the SAP invariants of SURVEY.md 8(c)(iv)."""
import copy

import pytest
import torch

import magic_b200
from magic_b200 import synth
from magic_b200.graph_index import batch_to_device, prepare_batch
from oracle import magic_oracle as O

CASES = {
    "one_sample": dict(B=1),
    "single_step_paths": dict(B=4, T_max=1),
    "four_token_text": dict(B=4, L=4),
    "four_node_graph": dict(B=4, T_max=2, G_max=4),
    "all_minimal": dict(B=1, T_max=1, L=4),
}
RW = [1.3, 0.6, 1.1, 0.9, 1.1]


def make(task, case, seed=5):
    kw = dict(CASES[case])
    return synth.make_batch(task, kw.pop("B"), seed=seed, **kw)


def obatch(b):
    return {k: v for k, v in b.items() if k != magic_b200.INDEX_KEY}


def oracle_pair(dev, seed):
    cfg_t = O.make_config(256, role="teacher", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    cfg_s = O.make_config(128, role="student", teacher_hidden_size=256, hidden_dropout_prob=0.0,
                          attention_probs_dropout_prob=0.0)
    torch.manual_seed(seed)
    t, s = O.GlocalTextPathCMTPreTraining(cfg_t), O.GlocalTextPathCMTPreTraining(cfg_s)
    g = torch.Generator().manual_seed(seed + 100)
    for m in (t, s):
        for n, p in m.named_parameters():
            if n.endswith("bias") or "LayerNorm" in n or "norm" in n:
                p.data.add_(torch.randn(p.shape, generator=g) * 0.05)
            elif "sprel_linear.weight" in n:
                p.data.fill_(-0.07)
    return cfg_t, cfg_s, t.to(dev).eval(), s.to(dev).train()


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("task", ["sap", "mlm"])
def test_edge_batches_on_the_oracle(task, case):
    b = make(task, case)
    prepare_batch(b)
    idx = b[magic_b200.INDEX_KEY]
    assert idx is not None
    _, _, teacher, student = oracle_pair("cpu", 31)
    tot, sup, kd, named, s_out, _ = O.distill_step_loss(student, teacher, obatch(b), task, torch.tensor(RW))
    assert torch.isfinite(tot) and sup.item() > 0 and kd.item() >= 0
    assert all(torch.isfinite(v) and v.item() >= 0 for v in named.values())
    if task == "sap":
        g, f = s_out["global_logits"], s_out["fused_logits"]
        assert torch.equal(torch.isinf(g), torch.isinf(f))  # fused == -inf exactly where global == -inf
        assert not torch.isinf(g[:, 0]).any()               # [stop] is never masked
        lens = b["gmap_lens"]
        for i in range(g.shape[0]):
            assert torch.isinf(g[i, int(lens[i]):]).all()   # padding is masked
    tot.backward()
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in student.parameters())


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("task", ["sap", "mlm"])
def test_edge_batches_cuda_vs_oracle(task, case):
    from magic_b200 import makd
    dev = "cuda"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg_t, cfg_s, t_o, s_o = oracle_pair(dev, 31)
    t_p = magic_b200.GlocalTextPathCMTPreTraining.from_pretrained(None, config=copy.copy(cfg_t),
                                                                  state_dict=t_o.state_dict()).to(dev).eval()
    s_p = magic_b200.GlocalTextPathCMTPreTraining.from_pretrained(None, config=copy.copy(cfg_s),
                                                                  state_dict=s_o.state_dict()).to(dev).train()
    b = batch_to_device(prepare_batch(make(task, case)), dev)
    tot_o, sup_o, kd_o, L_o, so, _ = O.distill_step_loss(s_o, t_o, obatch(b), task, torch.tensor(RW, device=dev))
    tot_o.backward()
    go = {n: p.grad for n, p in s_o.named_parameters()}
    gmax = max(g.norm().item() for g in go.values() if g is not None)
    for dtype, tol, gtol in ((torch.float32, 1e-4, 2e-3), (torch.bfloat16, 2e-2, 0.35)):
        t_p.set_compute_dtype(dtype)
        s_p.set_compute_dtype(dtype)
        s_p.zero_grad(set_to_none=True)
        mix, res, sp, _ = makd.distill_step_loss(s_p, t_p, b, task, RW)
        mix[0].backward()
        for got, want, what in ((mix[0], tot_o, "total"), (mix[1], sup_o, "supervised"), (mix[2], kd_o, "kd")):
            assert abs(got.item() - want.item()) <= tol * abs(want.item()) + 1e-7, (dtype, what, got.item(), want.item())
        named = makd.named_losses(res)
        for k, v in L_o.items():
            assert abs(named[k] - v.item()) <= tol * abs(v.item()) + 1e-6, (dtype, k, named[k], v.item())
        assert _rel(sp["loss"], so["loss"]) <= tol, dtype
        if task == "sap":
            for k in ("global_logits", "local_logits", "fused_logits"):
                assert torch.equal(torch.isinf(sp[k]), torch.isinf(so[k])), (dtype, k)  # masks bit-exact
                if dtype == torch.float32:
                    assert torch.equal(sp[k].argmax(1), so[k].argmax(1)), k              # argmax bit-exact
        elif dtype == torch.float32:
            assert torch.equal(sp["logits"].argmax(1), so["logits"].argmax(1))
        bad = []
        for n, p in s_p.named_parameters():
            ref = go[n]
            if ref is None or ref.norm() < 1e-4 * gmax:
                continue
            assert p.grad is not None, (dtype, n)
            r = _rel(p.grad, ref)
            if r > gtol:
                bad.append((n, round(r, 4), ref.norm().item()))
        assert not bad, (dtype, bad[:8])
