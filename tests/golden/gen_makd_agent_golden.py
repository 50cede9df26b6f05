"""Generates tests/golden/makd_agent_ref.pt by EXECUTING THE REFERENCE's own MAKD aggregation in the build container:
the SOURCE of `GMapNavAgent.compute_kd_losses` (/root/reference/map_nav_src/r2r/agent.py:546-719) is compiled as a
method of a stub agent (the module itself cannot be imported: it needs MatterSim, line_profiler and the absent
models/ package) and bound to the reference's own `mse_loss` / `kd_loss` (map_nav_src/utils/kd_loss.py, the way
agent_base.py:156-169 binds them).  Every case runs both roles -- 't2s' and the ICoD role 's2t' -- on the same
seeded tensors: no adaptive ability weight, 'RW' (= 'grad') and 'learned_weight', with and without MKTD sample
weights, reductions 'mean' (the pretraining flavour, pretrain_src/optim/kd_loss.py) and 'sum'.

Run:  python tests/golden/gen_makd_agent_golden.py      (needs /root/reference; the output is committed)"""
import importlib.util
import os
import re
import textwrap
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF_AGENT = "/root/reference/map_nav_src/r2r/agent.py"
REF_KD = "/root/reference/map_nav_src/utils/kd_loss.py"
OUT = os.path.join(HERE, "makd_agent_ref.pt")

B, L, V, G, VP = 4, 10, 7, 5, 8
HS, HT = 16, 32
NL_S, NL_T, NP, NX = 3, 5, 2, 3
NAMES = ("txt_emb_loss", "txt_attn_loss", "img_emb_loss", "avg_img_emb_loss", "img_attn_loss", "local_emb_loss",
         "local_attn_loss", "global_emb_loss", "global_attn_loss", "predict_loss")  # agent.py:824-835
PROJ = ("txt_emb_w", "kdl_img_w", "kdl_avg_img_w", "global_cross_w", "local_cross_w")
LEARNED = ("kdl_txt_weight", "kdl_img_weight", "kdl_global_weight", "kdl_local_weight", "kdl_predict_weight")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def reference_method(kd):
    src = open(REF_AGENT).read()
    m = re.search(r"^    def compute_kd_losses\(.*?(?=^    # @profile|^    def )", src, re.S | re.M)
    ns = dict(torch=torch, F=F, nn=nn)
    exec(textwrap.dedent(m.group(0)), ns)
    return ns["compute_kd_losses"]


def outputs(h, n_l, g, probs=True):
    """One model's KD outputs for a single navigation step (one panorama per sample), the keys compute_kd_losses
    reads: agent.py:560-704."""
    def r(*s):
        return torch.randn(*s, generator=g)

    def attn(*s):
        return torch.softmax(r(*s), -1) if probs else r(*s)

    logits = r(B, G) * 3
    return dict(txt_embeds=r(B, L, h), txt_attns=attn(B, n_l, L, L), pano_embeds=r(B, V, h),
                pano_fused_embeds=r(B, h), img_attns=attn(B, NP, V, V),
                nav_outs=dict(gmap_embeds=r(B, G, h), gmap_attns=attn(B, NX, G, G + L), vp_embeds=r(B, VP, h),
                              vp_attns=attn(B, NX, VP, VP + L)),
                nav_logits=logits)


def small_model(g):
    """The SMALL model's KD heads (it owns the up-projections HS -> HT, agent_base.py:330) + its learned weights."""
    m = types.SimpleNamespace()
    for n in PROJ:
        lin = nn.Linear(HS, HT)
        with torch.no_grad():
            lin.weight.copy_(torch.randn(HT, HS, generator=g) * 0.3)
            lin.bias.copy_(torch.randn(HT, generator=g) * 0.1)
        setattr(m, n, lin)
    for i, n in enumerate(LEARNED):
        setattr(m, n, nn.Parameter(torch.tensor([0.3 * i - 0.4])))
    return m


def large_model():
    m = types.SimpleNamespace()
    for i, n in enumerate(LEARNED):
        setattr(m, n, nn.Parameter(torch.tensor([0.7 - 0.25 * i])))
    return m


def build_inputs():
    g = torch.Generator().manual_seed(20261018)
    s_out, t_out = outputs(HS, NL_S, g), outputs(HT, NL_T, g)
    for o in (s_out, t_out):  # masked nodes: -inf in both models' logits (+ one only in the teacher's)
        o["nav_logits"][:, 3] = float("-inf")
    t_out["nav_logits"][1, 4] = float("-inf")
    s_w, t_w = torch.rand(B, generator=g), torch.rand(B, generator=g)
    rw = torch.softmax(torch.randn(5, generator=g) / 4.0, 0) * 5  # agent.py:866-869
    small = small_model(g)
    return s_out, t_out, s_w, t_w, rw, small, large_model()


def cases():
    for role in ("t2s", "s2t"):
        for kind in ("plain", "RW", "learned_weight"):
            for weighted in (True, False):
                for loss_type in (("mean", "sum") if role == "t2s" else ("mean",)):  # s2t is always 'mean' (:557)
                    yield dict(role=role, kind=kind, weighted=weighted, loss_type=loss_type)


def run_reference(fn, kd, case, s_out, t_out, s_w, t_w, rw, small, large):
    args = types.SimpleNamespace(
        kd_loss_type=case["loss_type"], kd_ability_types=["txt", "img", "local", "global", "action"],
        train_kdl_noFeat=False, train_kdl_noAttn=False, train_kdl_noLogit=False, kdl_temperature=2.0,
        kdl_adaptive_ability_weight=case["kind"] != "plain",
        kdl_adaptive_ability_weight_type=case["kind"] if case["kind"] != "plain" else "RW",
        kdl_logit_loss="kd", ignoreid=-100, kdl_dkd_alpha=1.0, kdl_dkd_beta=1.0)
    agent = types.SimpleNamespace(args=args, vln_bert=types.SimpleNamespace(vln_bert=small),
                                  teacher_vln_bert=types.SimpleNamespace(vln_bert=large),
                                  kdl_feat_loss=kd.mse_loss, kdl_attn_loss=kd.mse_loss, kdl_logit_loss=kd.kd_loss)
    so, to = dict(s_out), dict(t_out)
    so["sample_weights"] = s_w if case["weighted"] else None  # agent.py:1009-1011
    to["sample_weights"] = t_w if case["weighted"] else None  # agent.py:1019-1020
    zero = {n: 0. for n in NAMES}
    nav_targets = torch.zeros(B, dtype=torch.long)
    if case["role"] == "t2s":   # agent.py:1022
        out = fn(agent, 0, so, to, zero, nav_targets, role="t2s", softmax_weights=rw)
    else:                       # agent.py:1024
        out = fn(agent, 0, to, so, zero, nav_targets, role="s2t", softmax_weights=rw)
    return {n: float(v) for n, v in out.items()}


def main():
    kd = load(REF_KD, "ref_kd_fin")
    fn = reference_method(kd)
    s_out, t_out, s_w, t_w, rw, small, large = build_inputs()
    rec = []
    with torch.no_grad():
        for case in cases():
            rec.append(dict(case=case, named=run_reference(fn, kd, case, s_out, t_out, s_w, t_w, rw, small, large)))
    state = dict(small={n: {k: v.detach().clone() for k, v in getattr(small, n).state_dict().items()} for n in PROJ},
                 small_learned={n: getattr(small, n).detach().clone() for n in LEARNED},
                 large_learned={n: getattr(large, n).detach().clone() for n in LEARNED})
    torch.save(dict(s_out=s_out, t_out=t_out, s_w=s_w, t_w=t_w, rw=rw, state=state, cases=rec,
                    dims=dict(B=B, L=L, V=V, G=G, VP=VP, HS=HS, HT=HT)), OUT)
    print("wrote", OUT, len(rec), "cases")
    for r in rec[:3]:
        print(r)


if __name__ == "__main__":
    main()
