"""Host half of the fine-tune / inference path (nav.py): the online graph, the position features and the
collators -- checked against brute force and against the REFERENCE's own collator source
(map_nav_src/r2r/agent.py:108-173, :175-245, :289-328, executed from /root/reference when it is mounted)."""
import os
import re
import textwrap
import types

import numpy as np
import pytest
import torch

import magic_b200  # noqa: F401
from magic_b200 import nav, nav_synth

REF_AGENT = "/root/reference/map_nav_src/r2r/agent.py"


def _walk(world, start, steps, seed):
    rng = np.random.RandomState(seed)
    path = [start]
    for _ in range(steps):
        nb = np.nonzero(world.adj[path[-1]])[0]
        path.append(int(rng.choice(nb)))
    return path


def test_floyd_graph_matches_brute_force():
    world = nav_synth.NavWorld(n=20, seed=3)
    gmap = nav.GraphMap(world.ids[0])
    seen_edges = {}
    for i in _walk(world, 0, 9, seed=1):
        ob = world.observe(i)
        gmap.update_graph(ob)
        for cc in ob["candidate"]:
            j = world.index(cc["viewpointId"])
            seen_edges[(i, j)] = seen_edges[(j, i)] = float(np.linalg.norm(world.pos[i] - world.pos[j]))
        # brute force: Floyd-Warshall over the nodes seen so far, intermediate nodes restricted to VISITED ones
        names = list(gmap.node_positions)
        idx = {v: k for k, v in enumerate(names)}
        n = len(names)
        D = np.full((n, n), np.inf)
        np.fill_diagonal(D, 0)
        for (a, b), w in seen_edges.items():
            D[idx[world.ids[a]], idx[world.ids[b]]] = w
        for k in [idx[v] for v in names if gmap.graph.visited(v)]:
            D = np.minimum(D, D[:, k, None] + D[None, k, :])
        cur = ob["viewpoint"]
        for v in names:
            assert gmap.graph.distance(cur, v) == pytest.approx(D[idx[cur], idx[v]], rel=1e-12)
            p = gmap.graph.path(cur, v)
            if v != cur:
                assert p[-1] == v
                hops = [cur] + p
                assert sum(seen_edges[(world.index(a), world.index(b))] for a, b in zip(hops[:-1], hops[1:])) == \
                    pytest.approx(D[idx[cur], idx[v]], rel=1e-9)


def test_pos_fts_follow_the_pretraining_arithmetic():
    """GraphMap.get_pos_fts == dataset.py:553-575 arithmetic (restated with numpy on the same graph)."""
    world = nav_synth.NavWorld(n=16, seed=5)
    gmap = nav.GraphMap(world.ids[2])
    for i in _walk(world, 2, 5, seed=2):
        ob = world.observe(i, heading=0.7, elevation=-0.2)
        gmap.update_graph(ob)
    vpids = [None, None] + list(gmap.node_positions)
    f = gmap.get_pos_fts(ob["viewpoint"], vpids, 0.7, -0.2)
    assert f.shape == (len(vpids), 7) and f.dtype == np.float32
    assert np.all(f[:2, [0, 2, 4, 5, 6]] == 0) and np.all(f[:2, [1, 3]] == 1)  # sin 0 / cos 0, zero distances
    a = np.asarray(ob["position"])
    for k, vp in enumerate(vpids[2:], 2):
        b = gmap.node_positions[vp]
        dx, dy, dz = b - a
        xy, xyz = max(np.hypot(dx, dy), 1e-8), max(np.sqrt(dx * dx + dy * dy + dz * dz), 1e-8)
        h = np.arcsin(dx / xy)
        if b[1] < a[1]:
            h = np.pi - h
        h, e = np.float32(h - 0.7), np.float32(np.arcsin(dz / xyz) + 0.2)
        want = np.array([np.sin(h), np.cos(h), np.sin(e), np.cos(e), xyz / 30, gmap.graph.distance(ob["viewpoint"], vp) / 30,
                         len(gmap.graph.path(ob["viewpoint"], vp)) / 10], dtype=np.float32)
        assert np.allclose(f[k], want, rtol=1e-6, atol=1e-7)


def test_node_embeds_running_mean_and_rewrite():
    gmap = nav.GraphMap("a")
    e1, e2, e3 = torch.arange(4.0), torch.ones(4), torch.full((4,), 5.0)
    gmap.update_node_embed("x", e1)
    gmap.update_node_embed("x", e2)
    assert torch.allclose(gmap.get_node_embed("x"), (e1 + e2) / 2)
    gmap.update_node_embed("x", e3, rewrite=True)
    assert torch.equal(gmap.get_node_embed("x"), e3)
    gmap.update_node_embed("y", e2, teacher=True)
    gmap.update_node_embed("x", e1, teacher=True)
    assert torch.equal(gmap.node_embeds(["y", "x"], teacher=True), torch.stack([e2, e1]))
    for k in range(80):  # slab growth keeps earlier rows
        gmap.update_node_embed(f"n{k}", torch.full((4,), float(k)))
    assert torch.equal(gmap.get_node_embed("x"), e3) and gmap.get_node_embed("n79")[0] == 79


def _reference_methods(names):
    """The reference collators' SOURCE, compiled as methods of a stub agent (the module itself cannot be imported:
    it needs MatterSim, line_profiler and the absent models/ package)."""
    src = open(REF_AGENT).read()
    out = {}
    for name in names:
        m = re.search(rf"^    def {name}\(.*?(?=^    def )", src, re.S | re.M)
        out[name] = textwrap.dedent(m.group(0))
    import sys
    sys.path.insert(0, "/root/reference/map_nav_src")
    try:
        from utils.ops import pad_tensors, gen_seq_masks
    finally:
        sys.path.pop(0)

    def pad_tensors_wgrad(tensors, lens=None):  # models/ops.py is absent; DUET-lineage helper: zero-pad + stack
        n = max(t.size(0) for t in tensors)
        return torch.stack([torch.cat([t, t.new_zeros(n - t.size(0), *t.shape[1:])], 0) for t in tensors], 0)

    from torch.nn.utils.rnn import pad_sequence
    ns = dict(np=np, torch=torch, pad_tensors=pad_tensors, gen_seq_masks=gen_seq_masks, pad_sequence=pad_sequence,
              pad_tensors_wgrad=pad_tensors_wgrad)
    for name, code in out.items():
        exec(code, ns)
    stub = types.SimpleNamespace(args=types.SimpleNamespace(act_visited_nodes=False, enc_full_graph=True,
                                                            image_feat_size=768))
    return {name: types.MethodType(ns[name], stub) for name in names}


@pytest.mark.skipif(not os.path.exists(REF_AGENT), reason="reference tree not mounted")
def test_collators_match_the_reference_source(monkeypatch):
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    ref = _reference_methods(["_panorama_feature_variable_do", "_nav_gmap_variable", "_nav_vp_variable_mem"])
    B, h = 3, 8
    worlds = [nav_synth.NavWorld(n=18, seed=10 + b) for b in range(B)]
    walks = [_walk(w, b, 4, seed=20 + b) for b, w in enumerate(worlds)]
    gmaps = [nav.GraphMap(w.ids[walk[0]]) for w, walk in zip(worlds, walks)]
    g = torch.Generator().manual_seed(0)
    last = None
    for t in range(4):
        obs = [w.observe(walk[t], heading=0.3 * t, elevation=0.1 * b) for b, (w, walk) in enumerate(zip(worlds, walks))]
        for gm, ob in zip(gmaps, obs):
            gm.update_graph(ob)
            gm.node_step_ids[ob["viewpoint"]] = t + 1
        mine_p = nav.panorama_inputs(obs, "cpu")
        ref_p = ref["_panorama_feature_variable_do"](obs)
        for k in ("view_img_fts", "loc_fts", "nav_types", "view_lens"):
            assert torch.equal(mine_p[k], ref_p[k]), k
        assert mine_p["cand_vpids"] == ref_p["cand_vpids"]
        pano = torch.randn(B, mine_p["view_img_fts"].shape[1], h, generator=g)
        fused = torch.randn(B, h, generator=g)
        for i, (gm, ob) in enumerate(zip(gmaps, obs)):  # agent.py:905-924
            gm.update_node_embed(ob["viewpoint"], fused[i], rewrite=True)
            for j, c in enumerate(mine_p["cand_vpids"][i]):
                if not gm.graph.visited(c):
                    gm.update_node_embed(c, pano[i, j])
        mine_g = nav.nav_gmap_inputs(obs, gmaps, last)
        ref_g = ref["_nav_gmap_variable"](obs, gmaps, last)
        for k in ("gmap_step_ids", "gmap_pos_fts", "gmap_visited_masks", "gmap_pair_dists", "gmap_masks"):
            assert torch.equal(mine_g[k], ref_g[k]), (t, k)
        assert torch.allclose(mine_g["gmap_img_embeds"], ref_g["gmap_img_embeds"], rtol=1e-6, atol=1e-7)
        assert mine_g["gmap_vpids"] == ref_g["gmap_vpids"] and mine_g["no_vp_left"] == ref_g["no_vp_left"]
        mine_v = nav.nav_vp_inputs_mem(obs, gmaps, pano, mine_p["cand_vpids"], mine_p["view_lens"], mine_p["nav_types"], last)
        ref_v = ref["_nav_vp_variable_mem"](obs, gmaps, pano, ref_p["cand_vpids"], ref_p["view_lens"], ref_p["nav_types"],
                                            last)
        for k in ("vp_img_embeds", "vp_pos_fts", "vp_masks", "vp_nav_masks"):
            assert torch.equal(mine_v[k], ref_v[k]), (t, k)
        assert mine_v["vp_cand_vpids"] == ref_v["vp_cand_vpids"]
        last = torch.randn(B, h, generator=g)


def test_nav_index_tables():
    """Integer fusion tables == the string loops of the oracle, on a hand-made case with a back-tracking candidate."""
    from oracle.nav_oracle import fuse_logits
    gv = [[None, None, "a", "b", "c", "d"], [None, None, "p", "q"]]
    vis = torch.tensor([[0, 1, 1, 1, 0, 0], [0, 1, 1, 0, 0, 0]], dtype=torch.bool)
    gm = torch.tensor([[1, 0, 1, 1, 1, 1], [1, 0, 1, 1, 0, 0]], dtype=torch.bool)
    vc = [[None, None, "a", "c", "c"], [None, None, "q"]]  # b0: 'a' leads back (visited); 'c' twice: the last one wins
    nm = torch.tensor([[1, 0, 1, 1, 1], [1, 0, 1, 0, 0]], dtype=torch.bool)
    ix = nav.nav_index({"gmap_vpids": gv, "vp_cand_vpids": vc, "gmap_visited_masks": vis, "gmap_masks": gm, "vp_nav_masks": nm})
    assert ix["g_valid"].tolist() == [[1, 0, 0, 0, 1, 1], [1, 0, 0, 1, 0, 0]]
    assert ix["node2cand"].tolist() == [[-1, -1, -1, -1, 4, -1], [-1, -1, -1, 2, -1, -1]]
    assert ix["bw_mask"].tolist() == [[0, 0, 1, 0, 0], [0, 0, 0, 0, 0]]
    g = torch.randn(2, 6).masked_fill(vis | ~gm, float("-inf"))
    l = torch.randn(2, 5).masked_fill(~nm, float("-inf"))
    want = fuse_logits(g, l, gv, vis, vc, 2)
    got = g.clone()
    got[:, 0] += l[:, 0]
    for b in range(2):
        bw = sum(l[b, j] for j in range(5) if ix["bw_mask"][b, j])
        for n in range(1, 6):
            if ix["g_valid"][b, n]:
                c = int(ix["node2cand"][b, n])
                got[b, n] += l[b, c] if c >= 0 else bw
    assert torch.equal(torch.isinf(got), torch.isinf(want))
    assert torch.allclose(got[~torch.isinf(got)], want[~torch.isinf(want)])


def test_world_tables_and_graphworld_on_cpu():
    """nav_synth.world_tables (Floyd-Warshall distances / hop counts) against networkx, and featurizer.GraphWorld's table
    layout (built on the CPU here; the kernels that read it are covered by tests/test_featurize_graph_gpu.py)."""
    import networkx as nx
    from magic_b200.featurizer import GraphWorld
    w = nav_synth.NavWorld(n=30, seed=4)
    pos, dist, hops, cands, view_ang = nav_synth.world_tables(w)
    G = nx.Graph()
    for i in range(w.n):
        for j in np.nonzero(w.adj[i])[0]:
            G.add_edge(i, int(j), weight=float(np.linalg.norm(w.pos[i] - w.pos[j])))
    d = dict(nx.all_pairs_dijkstra_path_length(G))
    for i in range(w.n):
        for j in range(w.n):
            assert dist[i, j] == pytest.approx(d[i][j], rel=1e-5, abs=1e-6)
            assert dist[i, j] == 0 or hops[i, j] >= 1
    assert np.array_equal(hops, hops.T) and view_ang.shape == (36, 2)
    world = GraphWorld(pos, dist, hops, cands, view_ang, device="cpu")
    assert world.N == 30 and world.C == max(len(c) for c in cands) == world.max_cands
    for i, lst in enumerate(cands):
        assert int(world.n_cand[i]) == len(lst)
        assert world.cand_vp[i, :len(lst)].tolist() == [c[0] for c in lst] and (world.cand_vp[i, len(lst):] == -1).all()
        assert world.cand_view[i, :len(lst)].tolist() == [c[1] for c in lst]
    assert world.pos.dtype == torch.float64 and world.dist.dtype == torch.float32 and world.hops.dtype == torch.int32


def test_checkpoint_layouts_the_wrapper_accepts():
    """nav.remap_agent_keys / VLNBert.load_state_dict: our own layout, a DDP-prefixed one, the reference's fine-tune
    layout ([INFERRED]: no `.bert` level under `vln_bert`, agent_base.py:328-330) and a bare pretraining checkpoint
    (train_r2r_magic.py:189-208 key tree) all load into the same parameters; unknown keys are still reported."""
    import copy
    from magic_b200.config import make_config
    cfg = make_config(128, role="student", teacher_hidden_size=256, pretrain_tasks=("sap",))
    torch.manual_seed(3)
    src = nav.VLNBert(copy.copy(cfg))
    own = src.state_dict()
    assert any(k.startswith("vln_bert.bert.txt_emb_w.") for k in own) and "vln_bert.global_sap_head.net.0.weight" in own
    layouts = {
        "own": dict(own),
        "ddp": {"module." + k: v for k, v in own.items()},
        "reference_finetune": {k.replace("vln_bert.bert.", "vln_bert.", 1): v for k, v in own.items()},
        "pretraining": {k[len("vln_bert."):]: v for k, v in own.items()},
    }
    assert "vln_bert.txt_emb_w.weight" in layouts["reference_finetune"] and "bert.txt_emb_w.weight" in layouts["pretraining"]
    for name, sd in layouts.items():
        torch.manual_seed(4)
        dst = nav.VLNBert(copy.copy(cfg))
        res = dst.load_state_dict(sd)  # strict
        assert not res.missing_keys and not res.unexpected_keys, name
        for k, v in dst.state_dict().items():
            assert torch.equal(v, own[k]), (name, k)
    dst = nav.VLNBert(copy.copy(cfg))
    res = dst.load_state_dict(dict(own, **{"vln_bert.not_a_module.weight": torch.zeros(1)}), strict=False)
    assert res.unexpected_keys == ["vln_bert.not_a_module.weight"]
    with pytest.raises(RuntimeError):
        dst.load_state_dict({k: v for k, v in own.items() if "global_sap_head" not in k})  # strict: missing keys raise


@pytest.mark.skipif(not os.path.exists("/root/reference/map_nav_src/r2r/parser.py"), reason="reference tree not mounted")
def test_agent_args_build_both_roles(monkeypatch):
    """`VLNBert(self.args, role=...)` exactly as the agent calls it (agent.py:36-38), with the namespace the reference's
    own parser produces for the flags of scripts/run_r2r_kdl_valid.sh (MAGIC-B student 384 <- teacher 768)."""
    import sys
    import warnings
    src = open("/root/reference/map_nav_src/r2r/parser.py").read()
    ns = {}
    exec(src, ns)
    ns["postprocess_args"] = lambda a: a  # the path bookkeeping creates directories; the model flags are untouched by it
    flags = ("--mode valid --tokenizer roberta --enc_full_graph --graph_sprels --fusion dynamic --num_l_layers 6 "
             "--num_x_layers 3 --num_pano_layers 2 --angle_feat_size 4 --dropout 0.1 --adaptive_pano_fusion --train_kdl "
             "--kdl_temperature 2 --teacher_hidden_size 768 --teacher_num_l_layers 6 --teacher_num_pano_layers 2 "
             "--teacher_num_x_layers 3 --teacher_mlp_ratio 4 --student_num_l_layers 6 --student_num_x_layers 3 "
             "--student_num_pano_layers 2 --student_hidden_size 384 --student_mlp_ratio 4 --kdl_adaptive_ability_weight "
             "--kdl_adaptive_ability_weight_type RW --rw_temp 4 --do_back_txt").split()
    monkeypatch.setattr(sys, "argv", ["main_nav.py"] + flags)
    args = ns["parse_args"]()
    args.image_feat_size = 768  # set by postprocess_args (parser.py:218)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        student = nav.VLNBert(args, role="student")
        teacher = nav.VLNBert(args, role="teacher")
    assert any("causal-intervention" in str(x.message) for x in w)  # --do_back_txt is outside this path: said, not silent
    cs, ct = student.config, teacher.config
    assert (cs.hidden_size, cs.num_attention_heads, cs.intermediate_size) == (384, 6, 1536)
    assert (ct.hidden_size, ct.num_attention_heads, ct.intermediate_size) == (768, 12, 3072)
    assert (cs.num_l_layers, cs.num_x_layers, cs.num_pano_layers) == (6, 3, 2) and cs.vocab_size == 50265
    assert cs.role == "student" and ct.role == "teacher" and student.want_attn and teacher.want_attn
    # the student owns the up-projections to the teacher's width; the teacher has none (agent.py:550-568)
    assert student.vln_bert.txt_emb_w.weight.shape == (768, 384) and not hasattr(teacher.vln_bert, "txt_emb_w")
    assert student.vln_bert.bert.embeddings.word_embeddings.weight.shape == (50265, 384)
