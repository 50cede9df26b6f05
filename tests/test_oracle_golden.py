"""The oracle is frozen: on seeded inputs/weights it must reproduce the committed golden vectors (CPU);
on the GPU the CUDA path must match the same golden vectors through the public module API."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import gen_model_golden as G  # noqa: E402

GOLD = os.path.join(HERE, "golden", "model_golden.pt")


def _close(a, b, tol):
    a, b = a.float().cpu(), b.float().cpu()
    fin = torch.isfinite(b)
    assert torch.equal(fin, torch.isfinite(a))
    return ((a[fin] - b[fin]).norm() / (b[fin].norm() + 1e-20)).item() <= tol


@pytest.mark.parametrize("task", ["sap", "mlm"])
def test_oracle_reproduces_golden(task):
    gold = torch.load(GOLD)[task]
    _, _, teacher, student = G.build()
    rec = G.run(task, teacher, student)
    assert _close(rec["total"], gold["total"], 1e-5) and _close(rec["loss"], gold["loss"], 1e-5)
    for k, v in gold["named"].items():
        assert abs(rec["named"][k] - v) <= 1e-5 * abs(v) + 1e-9, k
    if task == "sap":
        assert torch.equal(rec["fused_logits"].argmax(1), gold["fused_logits"].argmax(1))
        assert _close(rec["fused_logits"], gold["fused_logits"], 1e-5)
    else:
        assert torch.equal(rec["logits_argmax"], gold["logits_argmax"])


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["sap", "mlm"])
def test_cuda_path_matches_golden(task):
    import magic_b200
    from magic_b200 import makd, synth
    from magic_b200.graph_index import prepare_batch, batch_to_device
    gold = torch.load(GOLD)[task]
    cfg_t, cfg_s, teacher, student = G.build()
    dev = "cuda"
    t_p = magic_b200.GlocalTextPathCMTPreTraining.from_pretrained(None, config=cfg_t, state_dict=teacher.state_dict()).to(dev).eval()
    s_p = magic_b200.GlocalTextPathCMTPreTraining.from_pretrained(None, config=cfg_s, state_dict=student.state_dict()).to(dev).train()
    b = batch_to_device(prepare_batch(synth.make_batch(task, 4, seed=2026)), dev)
    mix, res, s_out, t_out = makd.distill_step_loss(s_p, t_p, b, task, G.RW)
    assert _close(mix[0], gold["total"], 1e-4) and _close(mix[1], gold["sup"], 1e-4) and _close(mix[2], gold["kd"], 1e-4)
    assert _close(s_out["loss"], gold["loss"], 1e-4)
    named = makd.named_losses(res)
    for k, v in gold["named"].items():
        assert abs(named[k] - v) <= 1e-4 * abs(v) + 1e-7, (k, named[k], v)
    if task == "sap":
        for k in ("global_logits", "local_logits", "fused_logits"):
            assert torch.equal(torch.isinf(s_out[k]).cpu(), torch.isinf(gold[k])), k
            assert torch.equal(s_out[k].argmax(1).cpu(), gold[k].argmax(1)), k
            assert _close(s_out[k], gold[k], 1e-4), k
    else:
        assert torch.equal(s_out["logits"].argmax(1).cpu(), gold["logits_argmax"])
        assert _close(s_out["logits"][:, :64], gold["logits_head"], 1e-4)


def test_oracle_icod_graphs_are_disjoint_and_s2t_vanishes_on_matched_models():
    """ICoD restatement (oracle.icod_step_loss; agent.py:553-556, 1019-1022, 1136-1149), CPU:
    (i) every cross-model target is detached, so the small model's loss gives no gradient to the large model and
        vice versa (which is why one backward over the sum equals the reference's two backward calls);
    (ii) with kd_alpha = t_kd_alpha = 1 and the predict ability only, both directions compare the same two logit
        sets with swapped roles: KL(t||s) and KL(s||t) are both >= 0 and vanish together when the logits agree."""
    from oracle import magic_oracle as O
    from magic_b200 import synth
    _, _, teacher, student = G.build()
    teacher.train()
    b = synth.make_batch("sap", 4, seed=2027)
    rw = torch.tensor(G.RW)
    tot_s, tot_t, Ls, Lt, s_out, t_out = O.icod_step_loss(student, teacher, b, "sap", rw, rw)
    assert set(Ls) == set(Lt) and all(float(v) >= 0 for v in list(Ls.values()) + list(Lt.values()))
    tot_t.backward(retain_graph=True)
    assert all(p.grad is None or float(p.grad.abs().max()) == 0 for p in student.parameters())
    assert any(p.grad is not None and float(p.grad.abs().max()) > 0 for p in teacher.parameters())
    for p in teacher.parameters():
        p.grad = None
    tot_s.backward()
    assert all(p.grad is None or float(p.grad.abs().max()) == 0 for p in teacher.parameters())
    assert any(p.grad is not None and float(p.grad.abs().max()) > 0 for p in student.parameters())
    # (ii) identical logits -> the predict KL is exactly 0 in both roles
    same = dict(s_out)
    k = dict(kdl_tasks=("predict",))
    z1 = O.makd_losses(student, s_out, same, "sap", rw, None, k, role="t2s")["predict_loss"]
    z2 = O.makd_losses(student, same, s_out, "sap", rw, None, k, role="s2t")["predict_loss"]
    assert abs(float(z1)) < 1e-7 and abs(float(z2)) < 1e-7
