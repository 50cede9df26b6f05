"""CPU tests of the loop library's host logic (train_loop.py): RunningMeter and ModelSaver semantics of the reference
(pretrain_src/utils/logger.py:66-94, utils/save.py:23-74), and the seeded MetaLoader."""
import math
import os

import torch
import torch.nn as nn

import magic_b200  # noqa: F401
from magic_b200.train_loop import MetaLoader, ModelSaver, RunningMeter

REF = "/root/reference/pretrain_src"


def test_running_meter_matches_reference_semantics():
    m = RunningMeter("loss/mlm/total_loss")
    assert m.val == 0 and m.name == "loss/mlm/total_loss"
    m(2.0)
    assert m.val == 2.0                      # first value is taken as is
    m(4.0)
    assert abs(m.val - (4.0 * 0.01 + 2.0 * 0.99)) < 1e-12
    before = m.val
    m(float("nan"))
    assert m.val == before                   # NaN updates are dropped (logger.py:80-81)
    assert str(m).startswith("loss/mlm/total_loss: ")
    if os.path.exists(os.path.join(REF, "utils", "logger.py")):  # live check against the reference class
        src = open(os.path.join(REF, "utils", "logger.py")).read()
        i = src.index("class RunningMeter")
        ns = {"math": math}
        exec(src[i:], ns)
        r = ns["RunningMeter"]("x")
        ours = RunningMeter("x")
        for v in (1.0, 3.0, float("nan"), -2.0, 0.5):
            r(v), ours(v)
            assert r.val == ours.val


class _Opt:
    def state_dict(self):
        return {"step": 3, "x": torch.ones(2)}


def test_model_saver_layout(tmp_path):
    model = nn.Sequential(nn.Linear(3, 2))
    wrapped = nn.Module()
    wrapped.module = model       # a DDP-like wrapper: keys gain the `module.` prefix
    saver = ModelSaver(str(tmp_path))
    saver.save(wrapped, 7, _Opt())
    saver.save_latest(wrapped, 8, _Opt())
    saver.save_latest(wrapped, 9, _Opt(), is_max=True)
    files = sorted(os.listdir(tmp_path))
    assert files == ["model_step_7.pt", "model_step_best.pt", "model_step_latest.pt", "train_state_7.pt",
                     "train_state_best_9.pt", "train_state_latest.pt"]
    sd = torch.load(tmp_path / "model_step_7.pt")
    assert set(sd) == {"0.weight", "0.bias"} and all(v.device.type == "cpu" for v in sd.values())
    assert torch.equal(sd["0.weight"], model[0].weight.detach())
    st = torch.load(tmp_path / "train_state_7.pt")
    assert st["step"] == 7 and st["optimizer"]["step"] == 3


def test_meta_loader_is_seeded_and_cycles():
    loaders = {"mlm": [1, 2], "sap": [10], "cfp": [100, 200, 300]}
    a = [x for x in MetaLoader(loaders, [1, 1, 1], seed=5, num_steps=30)]
    b = [x for x in MetaLoader(loaders, [1, 1, 1], seed=5, num_steps=30)]
    assert a == b and len(a) == 30 and {t for t, _ in a} == {"mlm", "sap", "cfp"}
    mlm = [v for t, v in a if t == "mlm"]
    assert mlm[:4] == [1, 2, 1, 2][:len(mlm[:4])]          # a task's loader restarts when exhausted
    only = [t for t, _ in MetaLoader(loaders, [1, 0, 0], seed=1, num_steps=10)]
    assert set(only) == {"mlm"}


def test_inactive_in_task_rules():
    """model.inactive_in_task: which parameters a step of a task leaves without gradient (the optimizer skips them like
    the reference's `p.grad is None`); the GPU test checks the same rule against the real backward."""
    from magic_b200.model import inactive_in_task as ina
    assert not ina("mlm", "mlm_head.predictions.transform.dense.weight") and ina("sap", "mlm_head.predictions.bias")
    assert not ina("sap", "global_sap_head.net.0.weight") and ina("mlm", "sap_fuse_linear.net.3.bias")
    assert ina("mlm", "bert.global_encoder.sprel_linear.weight") and not ina("sap", "bert.global_encoder.sprel_linear.weight")
    assert ina("mrc", "bert.global_encoder.encoder.crossattention.0.attention.self.query.weight")
    assert not ina("mrc", "bert.local_encoder.encoder.crossattention.0.attention.self.query.weight")
    for t in ("mlm", "sap", "mrc", "cfp"):
        assert not ina(t, "bert.lang_encoder.layer.0.attention.self.query.weight")
        assert not ina(t, "bert.embeddings.word_embeddings.weight")
        assert ina(t, "bert.txt_emb_w.weight", kd=False)            # KD projections only train under a teacher
    assert not ina("mlm", "bert.vp_txt_w.weight", kd=True) and ina("sap", "bert.vp_txt_w.weight", kd=True)
    assert not ina("sap", "bert.global_cross_w.weight", kd=True) and ina("mlm", "bert.local_cross_w.bias", kd=True)
    assert not ina("sap", "bert.kdl_img_w.weight", kd=True) and not ina("mlm", "bert.kdl_txt_weight", kd=True)


def test_task_tables_of_the_task_aware_adamw():
    """optim.task_tables: per-task segment tables of magic_adamw_seg from an arena layout (pure host logic)."""
    from magic_b200.model import inactive_in_task
    from magic_b200.optim import task_tables
    names = ["bert.embeddings.word_embeddings.weight", "bert.lang_encoder.layer.0.attention.self.query.weight",
             "bert.global_encoder.sprel_linear.weight", "bert.txt_emb_w.weight", "bert.vp_txt_w.weight",
             "bert.global_cross_w.weight", "mlm_head.predictions.transform.dense.weight", "global_sap_head.net.0.weight",
             # no-decay group
             "bert.lang_encoder.layer.0.attention.self.query.bias", "mlm_head.predictions.bias", "global_sap_head.net.0.bias"]
    entries, off = [], 0
    for i, n in enumerate(names):
        k = 10 + 3 * i
        entries.append((n, off, k))
        off += (k + 7) // 8 * 8
        if i == 7:
            n_decay = off
    slots, tabs = task_tables(entries, n_decay, off, ["mlm", "sap"], lambda t, n: inactive_in_task(t, n, kd=True))
    assert slots == [frozenset({"mlm", "sap"}), frozenset({"mlm"}), frozenset({"sap"})]
    o = {n: e[1] for n, e in zip(names, entries)}
    (b_m, c_m), (b_m2, c_m2) = tabs["mlm"]
    # decay group in an MLM step: trunk (slot 0) | sprel_linear (untouched) | txt_emb_w (0) | vp_txt_w (mlm-only slot 1) |
    # global_cross_w (untouched) | mlm head (slot 1) | sap head (untouched)
    assert c_m == [0, -1, 0, 1, -1, 1, -1]
    assert b_m == [0, o[names[2]], o[names[3]], o[names[4]], o[names[5]], o[names[6]], o[names[7]], n_decay]
    assert c_m2 == [0, 1, -1] and b_m2 == [0, o[names[9]] - n_decay, o[names[10]] - n_decay, off - n_decay]
    (b_s, c_s), (_, c_s2) = tabs["sap"]
    # SAP step: trunk | sprel_linear (sap-only: slot 2) | txt_emb_w | vp_txt_w (untouched) | global_cross_w (2) | mlm head | sap head
    assert c_s == [0, 2, 0, -1, 2, -1, 2] and c_s2 == [0, -1, 2]
    assert b_s[0] == 0 and b_s[-1] == n_decay and b_s == sorted(b_s)
    # one task only: every parameter it touches is slot 0, the rest is skipped
    slots1, tabs1 = task_tables(entries, n_decay, off, ["sap"], lambda t, n: inactive_in_task(t, n, kd=False))
    assert slots1 == [frozenset({"sap"})] and set(tabs1["sap"][0][1]) == {0, -1}
