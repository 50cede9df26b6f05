"""Boundary hardening: drive the drop-in model through the REFERENCE's own construction code --
the real `config/r2r_magic_model_config.json`, the teacher / student config derivation of
`train_r2r_magic.py:122-160` and the METER checkpoint remap of `:185-208`, both EXECUTED FROM THE REFERENCE FILE
(source lines extracted at test time, nothing restated), then `set_dropout` (`utils/misc.py:19-25`) and the
`no_decay` grouping of `optim/misc.py:12-37`.  Also pins `FusedAdamW`'s arithmetic inputs: `optim.get_lr_sched`
against the fixture produced by the reference's `optim/sched.py` (tests/golden/gen_adamw_golden.py).

CPU-only (the nn.Module tree holds parameters; no kernels run).  Needs /root/reference, which exists in the build
container only: skipped elsewhere."""
import copy
import os
import re
import textwrap
from types import SimpleNamespace

import pytest
import torch

import magic_b200
from magic_b200 import optim as MO
from magic_b200.arena import NO_DECAY

REF = "/root/reference/pretrain_src"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adamw_ref.pt")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_r2r_magic.py")),
                               reason="/root/reference is only present in the build container")


class EasyDict(dict):
    """easydict is not installed here; attribute access over a dict is all the reference block uses."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _block(src, first, last):
    """Lines of `src` from the one containing `first` to the one containing `last`, dedented."""
    lines = src.split("\n")
    i = next(k for k, ln in enumerate(lines) if first in ln)
    j = next(k for k, ln in enumerate(lines) if last in ln and k >= i)
    return textwrap.dedent("\n".join(lines[i:j + 1]))


def _reference_configs():
    import json
    from transformers import PretrainedConfig
    src = open(os.path.join(REF, "train_r2r_magic.py")).read()
    opts_json = json.load(open(os.path.join(REF, "config", "r2r_magic_pretrain.json")))
    opts = SimpleNamespace(train_datasets=opts_json["train_datasets"], kdl=opts_json["kdl"], cuda_first_device="cuda:0")
    model_config = PretrainedConfig.from_json_file(os.path.join(REF, "config", "r2r_magic_model_config.json"))
    ns = dict(opts=opts, model_config=model_config, EasyDict=EasyDict, copy=copy)
    exec(_block(src, "model_config.pretrain_tasks = []", "model_config.cuda_first_device = opts.cuda_first_device"), ns)
    exec(_block(src, "kdl_cfg = EasyDict(opts.kdl)", "student_model_config.kdl = kdl_cfg"), ns)
    return src, ns["teacher_model_config"], ns["student_model_config"]


def _synthetic_meter(model, n_text_layers):
    """A METER-shaped state dict (`text_transformer.*`, `cross_modal_image_layers.*`, plus foreign keys) whose tensors
    are recognisable: built by renaming the model's own keys BACKWARDS through the reference remap."""
    sd = {}
    for k, v in model.state_dict().items():
        t = torch.full_like(v, float(len(sd) % 97) / 97.0)
        if k.startswith("bert.embeddings."):
            sd["text_transformer." + k[len("bert."):]] = t
        elif k.startswith("bert.lang_encoder.layer."):
            parts = k.split(".")
            parts[3] = str(2 * int(parts[3]))  # jump_init_txt: METER layer 2k -> ours k (train_r2r_magic.py:195-201)
            sd["text_transformer.encoder." + ".".join(parts[2:])] = t
        elif k.startswith("bert.global_encoder.encoder.crossattention."):
            sd["cross_modal_image_layers." + k[len("bert.global_encoder.encoder.crossattention."):]] = t
    for i in range(n_text_layers):  # the odd METER text layers the jump skips, and unrelated towers
        sd[f"text_transformer.encoder.layer.{2 * i + 1}.attention.self.query.weight"] = torch.zeros(2, 2)
    sd["vit_model.visual.conv1.weight"] = torch.zeros(3, 3)
    sd["cross_modal_text_layers.0.attention.self.query.weight"] = torch.zeros(2, 2)
    return sd


@needs_ref
def test_reference_config_derivation_and_meter_remap_drive_the_model():
    src, cfg_t, cfg_s = _reference_configs()
    # train_r2r_magic.py:142-143,156-157,154,160
    assert (cfg_t.hidden_size, cfg_t.num_attention_heads, cfg_t.intermediate_size, cfg_t.role) == (256, 4, 1024, "teacher")
    assert (cfg_s.hidden_size, cfg_s.num_attention_heads, cfg_s.intermediate_size, cfg_s.role) == (128, 2, 512, "student")
    assert cfg_s.teacher_hidden_size == 256 and cfg_s.kd and cfg_s.kdl.kd_temperature == 2
    assert cfg_s.pretrain_tasks == {"mlm", "sap", "cfp"}
    # the remap loop (first occurrence = the teacher's, :188-208): from the `for` line to the line before `del tmp`
    lines = src.split("\n")
    i = next(k for k, ln in enumerate(lines) if "for param_name, param in tmp.items():" in ln)
    j = next(k for k, ln in enumerate(lines) if "del tmp" in ln and k > i)
    loop = textwrap.dedent("\n".join(lines[i:j]))
    for cfg, role in ((cfg_t, "teacher"), (cfg_s, "student")):
        blank = magic_b200.GlocalTextPathCMTPreTraining(copy.copy(cfg))
        meter = _synthetic_meter(blank, cfg.num_l_layers)
        ns = dict(tmp=meter, teacher_checkpoint={}, model_config=cfg)
        exec(loop, ns)
        ckpt = ns["teacher_checkpoint"]
        model = magic_b200.GlocalTextPathCMTPreTraining.from_pretrained(
            pretrained_model_name_or_path=None, config=copy.copy(cfg), state_dict=ckpt)  # :260-262, :275-277
        info = model.load_info
        produced = [k for k in ckpt if k.startswith("bert.")]
        # every key the remap produces under `bert.` is consumed, except the odd METER text layers the jump leaves
        # with their original (out of range or colliding) index and that carry foreign shapes here
        leftovers = [k for k in produced if k not in info["loaded"]]
        assert all(re.match(r"bert\.lang_encoder\.layer\.\d+\.attention\.self\.query\.weight", k) for k in leftovers), \
            leftovers[:5]
        assert any(k.startswith("bert.local_encoder.encoder.crossattention.") for k in info["loaded"])
        assert any(k.startswith("bert.global_encoder.encoder.crossattention.") for k in info["loaded"])
        assert any(k.startswith("bert.lang_encoder.layer.%d." % (cfg.num_l_layers - 1)) for k in info["loaded"])
        assert "bert.embeddings.word_embeddings.weight" in info["loaded"]
        # local and global encoders are initialised from the SAME METER layers (:204-206)
        sd = model.state_dict()
        a = sd["bert.local_encoder.encoder.crossattention.0.attention.self.query.weight"]
        b = sd["bert.global_encoder.encoder.crossattention.0.attention.self.query.weight"]
        assert torch.equal(a, b) and torch.equal(a, meter["cross_modal_image_layers.0.attention.self.query.weight"])
        # foreign keys are tolerated and reported, never loaded
        assert "vit_model.visual.conv1.weight" in info["unexpected"]
        # heads follow pretrain_tasks (:104-107): mlm + sap + cfp, no mrc classifier
        assert hasattr(model, "mlm_head") and hasattr(model, "global_sap_head") and hasattr(model, "cfp_txt_proj")
        assert not hasattr(model, "image_classifier")
        # KD projection heads exist on the student only, sized student -> teacher (agent_base.py:330)
        has = hasattr(model.bert, "txt_emb_w")
        assert has == (role == "student")
        if has:
            assert tuple(model.bert.txt_emb_w.weight.shape) == (256, 128)
        # .train() / .eval() as :263-266, :278-279
        model.eval() if role == "teacher" else model.train()


@needs_ref
def test_set_dropout_and_no_decay_grouping_from_the_reference():
    import sys
    _, cfg_t, cfg_s = _reference_configs()
    model = magic_b200.GlocalTextPathCMTPreTraining(copy.copy(cfg_s))
    misc_src = open(os.path.join(REF, "utils", "misc.py")).read()
    ns = dict(torch=torch, LOGGER=SimpleNamespace(info=lambda *a, **k: None))
    exec(_block(misc_src, "def set_dropout(model, drop_p):", "LOGGER.info(f'{name} set to {drop_p}')"), ns)
    ns["set_dropout"](model, 0.37)
    drops = [m for m in model.modules() if isinstance(m, torch.nn.Dropout)]
    assert len(drops) > 20 and all(m.p == 0.37 for m in drops)
    sys.path.insert(0, REF)
    try:
        from optim.misc import build_optimizer
    finally:
        sys.path.remove(REF)
    opts = SimpleNamespace(optim="adamw", learning_rate=5e-5, betas=[0.9, 0.98], weight_decay=0.01)
    opt = build_optimizer(model, opts)
    ref_nodecay = {id(p) for p in opt.param_groups[1]["params"]}
    ours = {id(p) for n, p in model.named_parameters() if any(nd in n for nd in NO_DECAY)}
    assert ref_nodecay == ours and len(ours) > 50
    assert opt.param_groups[0]["weight_decay"] == 0.01 and opt.defaults["eps"] == 1e-6
    # every parameter is in exactly one group (the tied decoder weight appears once)
    n_all = len(list(model.parameters()))
    assert len(opt.param_groups[0]["params"]) + len(opt.param_groups[1]["params"]) == n_all


def test_lr_schedule_matches_reference_fixture():
    gold = torch.load(GOLD)
    opts = SimpleNamespace(**gold["opts"])
    for step, lr in gold["sched"]:
        assert MO.get_lr_sched(step, opts) == lr, (step, lr)


@needs_ref
def test_lr_schedule_matches_live_reference():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_sched", os.path.join(REF, "optim", "sched.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    opts = SimpleNamespace(learning_rate=5e-5, warmup_steps=10000, num_train_steps=200000)
    for s in (0, 1, 9999, 10000, 10001, 150000, 200000, 200001):
        assert MO.get_lr_sched(s, opts) == m.get_lr_sched(s, opts)
