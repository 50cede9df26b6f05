"""Row a14 pinned against the reference's own code: `GMapNavAgent.compute_kd_losses` (map_nav_src/r2r/agent.py:546-719).

tests/golden/makd_agent_ref.pt holds seeded KD outputs of a small and a large model, the small model's up-projections
and the 10 named losses the reference's SOURCE produced for both roles (t2s / s2t), the three ability-weight branches
(none / RW / learned_weight), with and without MKTD sample weights, reductions 'mean' and 'sum'
(tests/golden/gen_makd_agent_golden.py).  CPU: the oracle's `makd_losses` reproduces every case, and -- when
/root/reference is mounted -- so does a fresh run of the reference source.  GPU: the product's fused path
(`makd.compute_kd_losses`: up-projection GEMMs + one makd_mse launch + one makd_kl launch) reproduces every case in
fp32 mode to 1e-4 and its gradients match the oracle's autograd."""
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import gen_makd_agent_golden as G  # noqa: E402

from oracle import magic_oracle as O  # noqa: E402

GOLD = torch.load(os.path.join(HERE, "golden", "makd_agent_ref.pt"))
IDS = ["-".join(str(v) for v in r["case"].values()) for r in GOLD["cases"]]
ORDER = ("txt", "img", "global", "local", "predict")  # positions of softmax_weights, agent.py:588,621,685,688,715


def flat(o, dev="cpu", grad=False):
    """agent-style outputs (nav_outs nested, 'nav_logits') -> the pretraining output names of SURVEY.md 8(b)."""
    d = {k: v for k, v in o.items() if k not in ("nav_outs", "nav_logits")}
    d.update(o["nav_outs"])
    d["logits"] = o["nav_logits"]
    d = {k: v.to(dev).clone() for k, v in d.items()}
    if grad:
        for v in d.values():
            v.requires_grad_(True)
    return d


def models(dev="cpu"):
    st = GOLD["state"]
    small, large = types.SimpleNamespace(), types.SimpleNamespace()
    for n in G.PROJ:
        lin = nn.Linear(G.HS, G.HT)
        lin.load_state_dict(st["small"][n])
        setattr(small, n, lin.to(dev))
    for n in G.LEARNED:
        setattr(small, n, nn.Parameter(st["small_learned"][n].clone().to(dev)))
        setattr(large, n, nn.Parameter(st["large_learned"][n].clone().to(dev)))
    return types.SimpleNamespace(bert=small), types.SimpleNamespace(bert=large)


def kdl_of(case):
    return dict(kdl_adaptive_ability_weight=case["kind"] != "plain",
                kdl_adaptive_ability_weight_type=case["kind"] if case["kind"] != "plain" else "RW",
                kd_loss_type=case["loss_type"], kd_temperature=2.0)


def call(fn, case, small, large, s_out, t_out, rw, dev="cpu"):
    """Both implementations share the signature (small model, learner's outputs, target's outputs, task, rw, MKTD
    weights of the TARGET side, kdl, role) -- agent.py:1022 / :1024."""
    s_w = GOLD["s_w"].to(dev) if case["weighted"] else None
    t_w = GOLD["t_w"].to(dev) if case["weighted"] else None
    if case["role"] == "t2s":
        return fn(small, s_out, t_out, "sap", rw, t_w, kdl_of(case), role="t2s")
    return fn(small, t_out, s_out, "sap", rw, s_w, kdl_of(case), role="s2t", weight_owner=large)


@pytest.mark.parametrize("rec", GOLD["cases"], ids=IDS)
def test_oracle_makd_matches_the_reference_agent(rec):
    small, large = models()
    with torch.no_grad():
        L = call(O.makd_losses, rec["case"], small, large, flat(GOLD["s_out"]), flat(GOLD["t_out"]), GOLD["rw"])
    assert set(L) == set(rec["named"]) == set(G.NAMES)
    for k, v in rec["named"].items():
        assert abs(float(L[k]) - v) <= 1e-5 * abs(v) + 1e-8, (k, float(L[k]), v)


@pytest.mark.skipif(not os.path.exists(G.REF_AGENT), reason="reference tree not mounted")
def test_fixture_is_what_the_reference_source_produces():
    kd = G.load(G.REF_KD, "ref_kd_fin_live")
    fn = G.reference_method(kd)
    s_out, t_out, s_w, t_w, rw, small, large = G.build_inputs()
    assert torch.equal(rw, GOLD["rw"]) and torch.equal(s_out["txt_embeds"], GOLD["s_out"]["txt_embeds"])
    with torch.no_grad():
        for rec in GOLD["cases"]:
            live = G.run_reference(fn, kd, rec["case"], s_out, t_out, s_w, t_w, rw, small, large)
            for k, v in rec["named"].items():
                assert abs(live[k] - v) <= 1e-6 * abs(v) + 1e-9, (rec["case"], k)


def product_outputs(o, dev):
    """The product's KD outputs carry the attention maps as per-layer lists ([self | cross] pairs for the x-layers)."""
    d = flat(o, dev, grad=True)
    Lt = d["txt_embeds"].shape[1]

    def split(name, key, pair):
        a = d.pop(key)
        if pair:
            n = a.shape[2]
            d[name] = [(a[:, i, :, :n].contiguous(), a[:, i, :, n:].contiguous()) for i in range(a.shape[1])]
        else:
            d[name] = [a[:, i].contiguous() for i in range(a.shape[1])]
        return a

    roots = dict(txt_attns=split("txt_attn_list", "txt_attns", False), img_attns=split("img_attn_list", "img_attns", False),
                 gmap_attns=split("gmap_attn_list", "gmap_attns", True), vp_attns=split("vp_attn_list", "vp_attns", True))
    assert Lt == G.L
    return d, roots


@pytest.mark.gpu
@pytest.mark.parametrize("rec", GOLD["cases"], ids=IDS)
def test_cuda_makd_matches_the_reference_agent(rec):
    from magic_b200 import makd
    dev, case = "cuda", rec["case"]
    small, large = models(dev)
    rw = [float(x) for x in GOLD["rw"]]
    s_p, s_roots = product_outputs(GOLD["s_out"], dev)
    t_p, t_roots = product_outputs(GOLD["t_out"], dev)
    res = call(makd.compute_kd_losses, case, small, large, s_p, t_p, rw, dev)
    named = makd.named_losses(res)
    for k, v in rec["named"].items():
        assert abs(named[k] - v) <= 1e-4 * abs(v) + 1e-7, (k, named[k], v)
    # gradients of the fused path vs autograd through the oracle on the same device
    total = res["mse_total"] + res["kl"]
    total.backward()
    small_o, large_o = models(dev)
    s_o, t_o = flat(GOLD["s_out"], dev, grad=True), flat(GOLD["t_out"], dev, grad=True)
    L = call(O.makd_losses, case, small_o, large_o, s_o, t_o, torch.tensor(rw, device=dev), dev)
    sum(L.values()).backward()
    learner_p, learner_o = (s_p, s_o) if case["role"] == "t2s" else (t_p, t_o)
    learner_roots = s_roots if case["role"] == "t2s" else t_roots

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-20)).item()

    for k in ("txt_embeds", "pano_embeds", "pano_fused_embeds", "gmap_embeds", "vp_embeds", "logits"):
        assert rel(learner_p[k].grad, learner_o[k].grad) < 2e-3, k
    for k, root in learner_roots.items():
        assert rel(root.grad, learner_o[k].grad) < 2e-3, k
    if case["role"] == "t2s":  # the up-projections learn only in role t2s (their s2t outputs are detached targets)
        for n in G.PROJ:
            for pn in ("weight", "bias"):
                assert rel(getattr(getattr(small.bert, n), pn).grad, getattr(getattr(small_o.bert, n), pn).grad) < 2e-3, n
    if case["kind"] == "learned_weight":
        owner_p, owner_o = (small, small_o) if case["role"] == "t2s" else (large, large_o)
        for n in G.LEARNED:
            assert rel(getattr(owner_p.bert, n).grad, getattr(owner_o.bert, n).grad) < 2e-3, n


@pytest.mark.skipif(not os.path.exists(G.REF_AGENT), reason="reference tree not mounted")
def test_reference_function_runs_on_our_models():
    """Drop-in level 1 of INTEGRATION.md 1b: the reference agent's own `compute_kd_losses` source, unmodified, on a
    pair of OUR models wrapped like the agent wraps them (`self.vln_bert.vln_bert`): the KD heads and learned weights
    it reaches for (`s_model.txt_emb_w`, `s_model.kdl_txt_weight`, agent.py:552-568, 585) resolve on our model, and
    the result equals the oracle composition on the same tensors (CPU: the heads are plain nn.Linear modules)."""
    import magic_b200
    from magic_b200.config import make_config
    kd = G.load(G.REF_KD, "ref_kd_fin_dropin")
    fn = G.reference_method(kd)
    kdl = dict(kdl_adaptive_ability_weight=True, kdl_adaptive_ability_weight_type="learned_weight")
    cfg_s = make_config(128, role="student", teacher_hidden_size=256, pretrain_tasks=("sap",), kdl=kdl)
    cfg_t = make_config(256, role="teacher", pretrain_tasks=("sap",))
    torch.manual_seed(0)
    small = magic_b200.GlocalTextPathCMTPreTraining(cfg_s)
    large = magic_b200.GlocalTextPathCMTPreTraining(cfg_t)
    assert small.txt_emb_w is small.bert.txt_emb_w and small.kdl_txt_weight is small.bert.kdl_txt_weight
    with pytest.raises(AttributeError):
        small.no_such_head
    with pytest.raises(AttributeError):
        large.txt_emb_w                      # the large model owns no up-projections
    g = torch.Generator().manual_seed(5)
    B, L, V, Gn, VP = 3, 9, 6, 5, 7

    def outs(h, n_l):
        def r(*s):
            return torch.randn(*s, generator=g)
        return dict(txt_embeds=r(B, L, h), txt_attns=torch.softmax(r(B, n_l, L, L), -1), pano_embeds=r(B, V, h),
                    pano_fused_embeds=r(B, h), img_attns=torch.softmax(r(B, 2, V, V), -1),
                    nav_outs=dict(gmap_embeds=r(B, Gn, h), gmap_attns=torch.softmax(r(B, 3, Gn, Gn + L), -1),
                                  vp_embeds=r(B, VP, h), vp_attns=torch.softmax(r(B, 3, VP, VP + L), -1)),
                    nav_logits=r(B, Gn) * 2, sample_weights=torch.rand(B, generator=g))

    s_out, t_out = outs(128, 6), outs(256, 6)
    args = types.SimpleNamespace(
        kd_loss_type="mean", kd_ability_types=["txt", "img", "local", "global", "action"], train_kdl_noFeat=False,
        train_kdl_noAttn=False, train_kdl_noLogit=False, kdl_temperature=2.0, kdl_adaptive_ability_weight=True,
        kdl_adaptive_ability_weight_type="learned_weight", kdl_logit_loss="kd", ignoreid=-100, kdl_dkd_alpha=1.0,
        kdl_dkd_beta=1.0)
    agent = types.SimpleNamespace(args=args, vln_bert=types.SimpleNamespace(vln_bert=small),
                                  teacher_vln_bert=types.SimpleNamespace(vln_bert=large),
                                  kdl_feat_loss=kd.mse_loss, kdl_attn_loss=kd.mse_loss, kdl_logit_loss=kd.kd_loss)
    with torch.no_grad():
        ref = fn(agent, 0, s_out, t_out, {n: 0. for n in G.NAMES}, torch.zeros(B, dtype=torch.long), role="t2s")
        mine = O.makd_losses(small, flat(s_out), flat(t_out), "sap", None, t_out["sample_weights"],
                             dict(kdl, kd_loss_type="mean", kd_temperature=2.0), role="t2s")
    for k in G.NAMES:
        assert abs(float(ref[k]) - float(mine[k])) <= 1e-5 * abs(float(ref[k])) + 1e-8, k
