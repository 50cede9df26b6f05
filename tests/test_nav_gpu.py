"""GPU parity of the fine-tune / inference path (nav.VLNBert: language / panorama / navigation modes over an online
GraphMap with the [MEM] slot) against the fp32 oracle (oracle/nav_oracle.py) on synthetic episodes.
Tolerances: masks and argmax actions bit-exact; embeddings / logits 1e-4 relative in fp32 mode, 2e-2 in bf16 mode."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import magic_b200  # noqa: E402
from magic_b200 import nav, nav_synth  # noqa: E402
from oracle import magic_oracle as O  # noqa: E402
from oracle import nav_oracle as NO  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def build(h, seed=0, kd=True, **kw):
    cfg = O.make_config(h, role="student", teacher_hidden_size=256 if kd else None, hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, pretrain_tasks=("sap",), **kw)
    torch.manual_seed(seed)
    oracle = NO.VLNBert(copy.copy(cfg))
    g = torch.Generator().manual_seed(seed + 100)
    for n, p in oracle.named_parameters():
        if n.endswith("bias") or "LayerNorm" in n or "norm" in n:
            p.data.add_(torch.randn(p.shape, generator=g) * 0.05)
        elif "sprel_linear.weight" in n:
            p.data.fill_(-0.07)
    oracle = oracle.to(DEV).eval()
    prod = nav.VLNBert(copy.copy(cfg)).to(DEV).eval()
    missing, unexpected = prod.load_state_dict(oracle.state_dict(), strict=False)
    assert not unexpected and not [k for k in missing if "kdl_" not in k], (missing, unexpected)
    return oracle, prod


def to_dev(d):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in d.items()}


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("B,h", [(1, 128), (4, 128), (16, 128), (3, 768)])
def test_rollout_matches_oracle(dtype, tol, B, h):
    oracle, prod = build(h, num_l_layers=2 if h == 768 else 6, num_x_layers=2 if h == 768 else 3)
    prod.set_compute_dtype(dtype)
    rng = np.random.RandomState(7)
    worlds = [nav_synth.NavWorld(n=22, seed=40 + b) for b in range(B)]
    cur = [int(rng.randint(0, 22)) for _ in range(B)]
    obs = [w.observe(c, instr=nav_synth.make_instr(rng, 40)) for w, c in zip(worlds, cur)]
    gmaps = [nav.GraphMap(ob["viewpoint"]) for ob in obs]
    for gm, ob in zip(gmaps, obs):
        gm.update_graph(ob)
    lang = nav.language_inputs(obs, DEV)
    with torch.no_grad():
        t_o, a_o = oracle("language", lang)
        t_p, a_p = prod("language", lang)
    assert rel(t_p, t_o) < tol and rel(a_p, a_o) < tol
    last = None
    for t in range(5):
        for gm, ob in zip(gmaps, obs):
            gm.node_step_ids[ob["viewpoint"]] = t + 1
        pin = nav.panorama_inputs(obs, DEV)
        with torch.no_grad():
            pe_o, pm_o, pf_o, pa_o = oracle("panorama", pin)
            pe_p, pm_p, pf_p, pa_p = prod("panorama", pin)
        assert torch.equal(pm_o, pm_p)
        valid = pm_o[..., None].expand_as(pe_o)
        assert rel(pe_p.float()[valid], pe_o[valid]) < tol and rel(pf_p, pf_o) < tol and rel(pa_p, pa_o) < tol
        # the graph is fed the ORACLE's embeddings so both models see identical navigation inputs at every step
        for i, (gm, ob) in enumerate(zip(gmaps, obs)):
            gm.update_node_embed(ob["viewpoint"], pf_o[i], rewrite=True)
            for j, c in enumerate(pin["cand_vpids"][i]):
                if not gm.graph.visited(c):
                    gm.update_node_embed(c, pe_o[i, j])
        nin = nav.nav_gmap_inputs(obs, gmaps, last)
        nin.update(nav.nav_vp_inputs_mem(obs, gmaps, pe_o, pin["cand_vpids"], pin["view_lens"], pin["nav_types"], last))
        nin.update(txt_embeds=t_o, txt_masks=lang["txt_masks"])
        nin = to_dev(nin)
        with torch.no_grad():
            no = oracle("navigation", nin)
            npd = prod("navigation", nin)
        for k in ("global_logits", "local_logits", "fused_logits"):
            assert torch.equal(torch.isinf(npd[k]), torch.isinf(no[k])), (t, k)   # masks bit-exact
            fin = ~torch.isinf(no[k])
            assert rel(npd[k][fin], no[k][fin]) < tol, (t, k)
        if dtype == torch.float32:
            assert torch.equal(npd["fused_logits"].argmax(1), no["fused_logits"].argmax(1))  # actions bit-exact
        gm_valid = nin["gmap_masks"].clone()
        gm_valid[:, 1] = True  # the [MEM] query row is computed (it only is no key)
        for k, m in (("gmap_embeds", gm_valid), ("vp_embeds", nin["vp_masks"])):
            mm = m[..., None].expand_as(no[k])
            assert rel(npd[k].float()[mm], no[k][mm]) < tol, (t, k)
        assert rel(npd["cls_embeds"], no["cls_embeds"]) < tol
        for k, m in (("gmap_attns", gm_valid), ("vp_attns", nin["vp_masks"])):
            mm = m[:, None, :, None].expand_as(no[k])
            assert rel(npd[k][mm], no[k][mm]) < tol, (t, k)
        # the [MEM] key column of the graph self-attention map is exactly zero (agent.py:228)
        G = nin["gmap_masks"].shape[1]
        assert float(npd["gmap_attns"][:, :, :, 1].abs().max()) == 0.0 and G > 2
        last = no["cls_embeds"]
        # act on the oracle's decision: move to the chosen node (or stay on [stop])
        act = no["fused_logits"].argmax(1).tolist()
        nxt = []
        for i, (w, ob) in enumerate(zip(worlds, obs)):
            vp = nin["gmap_vpids"][i][act[i]]
            j = w.index(vp) if vp is not None else w.index(ob["viewpoint"])
            if vp is None:  # [stop]: keep exploring anyway so the graphs keep growing
                j = int(np.nonzero(w.adj[j])[0][t % int(w.adj[j].sum())])
            nxt.append(w.observe(j, heading=0.4 * (t + 1), elevation=0.05 * i, instr=ob["instr_encoding"]))
        obs = nxt
        for gm, ob in zip(gmaps, obs):
            gm.update_graph(ob)


def test_checkpoint_from_pretraining_loads():
    """A pretraining checkpoint (GlocalTextPathCMTPreTraining.state_dict) drives the navigation wrapper unchanged."""
    cfg = O.make_config(128, pretrain_tasks=("mlm", "sap"), hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(0)
    pre = magic_b200.GlocalTextPathCMTPreTraining(cfg).to(DEV).eval()
    wrapped = nav.VLNBert.from_pretraining(pre)
    fresh = nav.VLNBert(copy.copy(cfg)).to(DEV).eval()
    fresh.vln_bert.load_state_dict(pre.state_dict(), strict=False)
    rng = np.random.RandomState(0)
    w = nav_synth.NavWorld(n=12, seed=1)
    lang = nav.language_inputs([w.observe(0, instr=nav_synth.make_instr(rng, 20))], DEV)
    with torch.no_grad():
        a, _ = wrapped("language", lang)
        b, _ = fresh("language", lang)
    assert torch.equal(a, b)
    with pytest.raises(NotImplementedError):
        wrapped("nope", lang)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_graph_replayed_decision_step_equals_eager(dtype, tol):
    """nav.NavStepper (panorama + navigation modes as CUDA graphs over capacity-padded buffers) == the eager modes, for
    graphs that grow from step to step."""
    B = 4
    _, prod = build(128)
    prod.set_compute_dtype(dtype)
    stepper = nav.NavStepper(prod, B, G=48, Lt=48)
    rng = np.random.RandomState(11)
    worlds = [nav_synth.NavWorld(n=22, seed=70 + b) for b in range(B)]
    obs = [w.observe(int(rng.randint(0, 22)), instr=nav_synth.make_instr(rng, 40)) for w in worlds]
    gmaps = [nav.GraphMap(ob["viewpoint"]) for ob in obs]
    for gm, ob in zip(gmaps, obs):
        gm.update_graph(ob)
    lang = nav.language_inputs(obs, DEV)
    with torch.no_grad():
        txt, _ = prod("language", lang)
    last = None
    for t in range(4):
        for gm, ob in zip(gmaps, obs):
            gm.node_step_ids[ob["viewpoint"]] = t + 1
        pin = nav.panorama_inputs(obs, DEV)
        with torch.no_grad():
            pe, pm, pf, pa = prod("panorama", pin)
        pe2, pm2, pf2, pa2 = stepper.panorama(pin)
        assert torch.equal(pm, pm2) and rel(pe2, pe) < tol and rel(pf2, pf) < tol and rel(pa2, pa) < tol
        for i, (gm, ob) in enumerate(zip(gmaps, obs)):
            gm.update_node_embed(ob["viewpoint"], pf[i], rewrite=True)
            for j, c in enumerate(pin["cand_vpids"][i]):
                if not gm.graph.visited(c):
                    gm.update_node_embed(c, pe[i, j])
        nin = nav.nav_gmap_inputs(obs, gmaps, last)
        nin.update(nav.nav_vp_inputs_mem(obs, gmaps, pe, pin["cand_vpids"], pin["view_lens"], pin["nav_types"], last))
        nin.update(txt_embeds=txt, txt_masks=lang["txt_masks"], txt_lens=lang["txt_lens"])
        with torch.no_grad():
            a = prod("navigation", nin)
        b = stepper.navigation(nin)
        for k in ("global_logits", "local_logits", "fused_logits"):
            assert torch.equal(torch.isinf(a[k]), torch.isinf(b[k])), (t, k)
            fin = ~torch.isinf(a[k])
            assert rel(b[k][fin], a[k][fin]) < tol, (t, k)
        assert torch.equal(a["fused_logits"].argmax(1), b["fused_logits"].argmax(1)) or dtype == torch.bfloat16
        valid = nin["gmap_masks"].clone()
        valid[:, 1] = True
        mm = valid[..., None].expand_as(a["gmap_embeds"])
        assert rel(b["gmap_embeds"].float()[mm], a["gmap_embeds"].float()[mm]) < tol
        assert rel(b["vp_embeds"], a["vp_embeds"]) < tol and rel(b["cls_embeds"], a["cls_embeds"]) < tol
        m4 = valid[:, None, :, None].expand_as(a["gmap_attns"])
        assert a["gmap_attns"].shape == b["gmap_attns"].shape and rel(b["gmap_attns"][m4], a["gmap_attns"][m4]) < tol
        assert rel(b["vp_attns"], a["vp_attns"]) < tol
        last = a["cls_embeds"].clone()
        act = a["fused_logits"].argmax(1).tolist()
        nxt = []
        for i, (w, ob) in enumerate(zip(worlds, obs)):
            vp = nin["gmap_vpids"][i][act[i]]
            nb = np.nonzero(w.adj[w.index(ob["viewpoint"])])[0]
            j = w.index(vp) if vp is not None else int(nb[t % len(nb)])
            nxt.append(w.observe(j, heading=0.4 * (t + 1), instr=ob["instr_encoding"]))
        obs = nxt
        for gm, ob in zip(gmaps, obs):
            gm.update_graph(ob)
