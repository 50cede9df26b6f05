"""GPU test of the loop library (train_loop.train): a short distillation run on synthetic loaders -- the loss goes
down, the meters the reference declares (train_r2r_magic.py:380-390) are fed, validation returns the reference's
metric keys (:412-587), and a checkpoint written by ModelSaver resumes bit-exactly (weights + AdamW state)."""
import copy
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

import magic_b200  # noqa: E402
from magic_b200 import synth  # noqa: E402
from magic_b200.config import make_config  # noqa: E402
from magic_b200.graph_index import batch_to_device, pad_batch, prepare_batch  # noqa: E402
from magic_b200.train_loop import MetaLoader, ModelSaver, train, validate  # noqa: E402
from magic_b200.train_step import PretrainStepper  # noqa: E402

DEV = "cuda"
TASKS = ("mlm", "sap", "cfp")


def _models():
    cfg_s = make_config(128, role="student", teacher_hidden_size=256, hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, pretrain_tasks=TASKS)
    cfg_t = make_config(256, role="teacher", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                        pretrain_tasks=TASKS)
    torch.manual_seed(1)
    s = magic_b200.GlocalTextPathCMTPreTraining(cfg_s).to(DEV).train()
    torch.manual_seed(0)
    t = magic_b200.GlocalTextPathCMTPreTraining(cfg_t).to(DEV).eval()
    return s, t


def _loaders(n=2, B=8):
    K = magic_b200.INDEX_KEY
    out = {}
    for ti, task in enumerate(("mlm", "sap")):
        bs = [prepare_batch(synth.make_batch(task, B, seed=300 + 10 * ti + i)) for i in range(n)]
        rcap = (max(b["traj_view_img_fts"].shape[0] for b in bs) + 7) // 8 * 8
        mcap = (max(b[K]["mlm_rows"].numel() for b in bs) + 63) // 64 * 64 if task == "mlm" else None
        ecap = (max(b[K]["entries"].numel() for b in bs) + 255) // 256 * 256
        scap = (max(b[K]["src_ids"].numel() for b in bs) + 255) // 256 * 256
        out[task] = [pad_batch(b, rcap, mcap, ecap, scap) for b in bs]
    return out


def _val_loaders():
    return {task: [batch_to_device(prepare_batch(synth.make_batch(task, 6, seed=900 + i)), DEV) for i in range(2)]
            for task in TASKS}


def test_short_run_feeds_meters_validates_and_resumes(tmp_path):
    s, t = _models()
    opts = SimpleNamespace(learning_rate=2e-3, warmup_steps=2, num_train_steps=24, log_steps=6, valid_steps=12,
                           output_dir=str(tmp_path))
    stepper = PretrainStepper(s, t, lr=opts.learning_rate, use_graphs=True, rw_generator=torch.Generator().manual_seed(3))
    logs = []
    saver = ModelSaver(str(tmp_path / "ckpts"))
    res = train(opts, stepper, MetaLoader(_loaders(), [1, 1], seed=4, num_steps=opts.num_train_steps),
                val_dataloaders=_val_loaders(), val2_dataloaders={"sap": _val_loaders()["sap"]}, model_saver=saver,
                log=logs.append)
    assert res["global_step"] == 24
    steps = [l for l in logs if "step" in l]
    assert [l["step"] for l in steps] == [6, 12, 18, 24]
    keys = set(steps[-1])
    for task in ("mlm", "sap"):
        for k in ("total_loss", "supervised_loss", "kdl_loss", "txt", "img", "global", "local", "predict"):
            assert f"loss/{task}/{k}" in keys, (task, k)
    # 24 steps on 2 + 2 batches at lr 2e-3: the total loss of both tasks falls
    assert steps[-1]["loss/mlm/total_loss"] < steps[0]["loss/mlm/total_loss"]
    assert steps[-1]["loss/sap/total_loss"] < steps[0]["loss/sap/total_loss"]
    assert all(torch.isfinite(torch.tensor(float(v))) for v in steps[-1].values())
    vals = [l for l in logs if any(k.startswith("val_seen") for k in l)]
    assert len(vals) == 2
    for k in ("val_seen_mlm_loss", "val_seen_mlm_acc", "val_seen_mlm_tok_per_s", "val_seen_sap_gloss", "val_seen_sap_facc",
              "val_seen_cfp_gloss", "val_seen_cfp_facc"):
        assert k in vals[-1], k
    assert 0.0 <= vals[-1]["val_seen_sap_facc"] <= 1.0 and vals[-1]["val_seen_mlm_loss"] > 0
    assert (tmp_path / "ckpts" / "model_step_latest.pt").exists() and (tmp_path / "ckpts" / "train_state_latest.pt").exists()
    assert res["best_unseen_facc"] >= 0

    # resume: a fresh model + optimizer loaded from the checkpoint takes the same next step as the original
    sd = torch.load(tmp_path / "ckpts" / "model_step_latest.pt")
    assert set(sd) == set(s.state_dict())
    ts = torch.load(tmp_path / "ckpts" / "train_state_latest.pt")
    assert ts["step"] == 24
    s2, t2 = _models()
    stepper2 = PretrainStepper(s2, t2, lr=opts.learning_rate, use_graphs=False,
                               rw_generator=torch.Generator().manual_seed(9))
    s2.load_state_dict(sd)            # in place, AFTER the arena exists: the bf16/fp32 arena must follow (sync_lowp)
    stepper2.opt.load_state_dict(ts["optimizer"])
    for a, b in zip(s.parameters(), s2.parameters()):
        assert torch.equal(a, b)
    assert torch.equal(stepper.opt.m, stepper2.opt.m) and stepper2.opt.step_count == stepper.opt.step_count
    b = batch_to_device(_loaders()["sap"][0], DEV)
    stepper.use_graphs = False
    stepper.pipeline_teacher = False
    stepper.rw_generator = torch.Generator().manual_seed(9)
    o1 = stepper.step("sap", b, lr=1e-3).clone()
    o2 = stepper2.step("sap", b, lr=1e-3).clone()
    assert torch.allclose(o1, o2, rtol=2e-5, atol=1e-6), (o1, o2)
    d = max(float((a - c).abs().max()) for a, c in zip(s.parameters(), s2.parameters()))
    assert d < 1e-5, d


def test_validate_matches_reference_metric_definitions():
    """validate_sap / validate_mlm against the formulas of train_r2r_magic.py:440-533 evaluated with torch on the
    model's own outputs."""
    import torch.nn.functional as F
    s, _ = _models()
    loaders = _val_loaders()
    out = validate(s, {"sap": loaders["sap"], "mlm": loaders["mlm"]}, setname="_unseen")
    s.eval()
    g = l = f = n = 0.0
    gc = 0
    with torch.no_grad():
        for b in loaders["sap"]:
            o = s(b, task="sap", compute_loss=False)
            g += F.cross_entropy(o["global_logits"], o["global_act_labels"], reduction="sum").item()
            f += F.cross_entropy(o["fused_logits"], o["global_act_labels"], reduction="sum").item()
            gc += int((o["global_logits"].argmax(1) == o["global_act_labels"]).sum())
            n += len(o["global_act_labels"])
        ml = mw = 0.0
        for b in loaders["mlm"]:
            sc = s(b, task="mlm", compute_loss=False)["predict"]
            lab = b["txt_labels"][b["txt_labels"] != -1]
            ml += F.cross_entropy(sc, lab, reduction="sum").item()
            mw += lab.numel()
    assert abs(out["val_unseen_sap_gloss"] - g / n) < 1e-4 * abs(g / n)
    assert abs(out["val_unseen_sap_floss"] - f / n) < 1e-4 * abs(f / n)
    assert abs(out["val_unseen_sap_gacc"] - gc / n) < 1e-9
    assert abs(out["val_unseen_mlm_loss"] - ml / mw) < 1e-4 * abs(ml / mw)
    assert s.training is False  # validate() restores the mode it found
