"""GPU batch featuriser, graph half (featurizer.GraphFeaturizer -> magic_featurize_graph) against golden vectors made
by the REFERENCE's own loader code (tests/golden/gen_featurizer_golden.py: dataset.py get_input + tasks.py SapDataset /
sap_collate on a synthetic world).  Integer tensors and index tables bit-exact; float features within 1e-6 (numpy
evaluates sin / cos of fp32 angles in fp32, the kernel in fp64 and rounds)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

import magic_b200  # noqa: E402
from magic_b200.featurizer import GraphFeaturizer, GraphWorld, attach_text  # noqa: E402
from magic_b200.graph_index import INDEX_KEY  # noqa: E402

DEV = "cuda"
GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# "" = 18-viewpoint scans, paths of up to 9 steps; "_long" = 40-viewpoint scans, paths of up to 21 viewpoints (the
# loader's TRAIN_MAX_STEP + 1), graphs of 30 nodes, many revisits
FIXTURES = ["", "_long"]


def load(tag=""):
    g = torch.load(os.path.join(GOLD_DIR, f"featurizer_graph{tag}.pt"), weights_only=False)
    w = g["world"]
    positions = {s: {v: w["pos"][s][v] for v in w["nodes"][s]} for s in w["nodes"]}
    world, rows = GraphWorld.from_tables(positions, w["dist"], w["paths_len"], w["cands"], w["view_ang"], DEV)
    return g, world, rows


@pytest.mark.parametrize("tag", FIXTURES)
@pytest.mark.parametrize("case", ["plain", "correct_heading"])
@pytest.mark.parametrize("slack", [0, 3])
def test_featuriser_matches_reference_loader(case, slack, tag):
    g, world, rows = load(tag)
    c = g["cases"][case]
    ref, rix = c["batch"], c["index"]
    B = len(c["paths"])
    G = ref["gmap_step_ids"].shape[1] + slack
    R = sum(len(p) for p in c["paths"])
    n_ent = rix["entries"].numel()
    feat = GraphFeaturizer(world, B, Tmax=max(len(p) for p in c["paths"]) + slack, G=G, R_cap=R + slack,
                           E_cap=n_ent + 5 * slack, S_cap=n_ent + 7 * slack, correct_heading=c["correct_heading"])
    paths = [[rows[f"{s}_{v}"] for v in p] for s, p in zip(c["scans"], c["paths"])]
    nxt = [-1 if v is None else rows[f"{s}_{v}"] for s, v in zip(c["scans"], c["next_vp"])]
    prev = [None if v is None else rows[f"{s}_{v}"] for s, v in zip(c["scans"], c["prev_vp"])]
    for rep in range(2):  # the second call reuses every buffer
        out = feat(paths, c["headings"], nxt, prev)
        feat.check()
    ix = out[INDEX_KEY]
    G0 = ref["gmap_step_ids"].shape[1]
    cpu = lambda t: t.detach().cpu()
    # ---- integer tensors: bit-exact ------------------------------------------------------------------
    assert torch.equal(cpu(out["gmap_lens"]), ref["gmap_lens"])
    assert torch.equal(cpu(out["gmap_step_ids"])[:, :G0], ref["gmap_step_ids"])
    assert torch.equal(cpu(out["gmap_visited_masks"])[:, :G0], ref["gmap_visited_masks"])
    assert int(cpu(out["gmap_step_ids"])[:, G0:].abs().sum()) == 0
    assert torch.equal(cpu(out["traj_nav_types"])[:R], ref["traj_nav_types"])
    assert torch.equal(cpu(out["traj_vp_view_lens"])[:R], ref["traj_vp_view_lens"])
    assert torch.equal(cpu(out["global_act_labels"]), ref["global_act_labels"])
    assert torch.equal(cpu(out["local_act_labels"]), ref["local_act_labels"])
    # graph nodes = the reference's gmap_vpids (strings) as world rows
    node_vp = cpu(out["gmap_node_vp"])
    for b in range(B):
        want = [-1] + [rows[f"{c['scans'][b]}_{v}"] for v in ref["gmap_vpids"][b][1:]]
        assert node_vp[b, :len(want)].tolist() == want and (node_vp[b, len(want):] == -1).all()
    # candidate-first view order: the candidates' view indices, then the unused views ascending
    perm = cpu(out["traj_view_perm"])
    r = 0
    for b in range(B):
        for t, vp in enumerate(c["paths"][b]):
            cand = g["world"]["cands"][f"{c['scans'][b]}_{vp}"]
            views = [v[0] for v in cand.values()]
            want = views + [i for i in range(36) if i not in set(views)]
            assert perm[r].tolist() == want, (b, t)
            assert int(out["traj_vp_index"][r]) == rows[f"{c['scans'][b]}_{vp}"]
            r += 1
    if slack:  # padded panoramas: one zero view, referenced by nothing
        assert (perm[R:] == -1).all() and (cpu(out["traj_vp_view_lens"])[R:] == 1).all()
    # ---- float features ------------------------------------------------------------------------------
    def close(a, b, name):
        assert a.shape == b.shape, (name, a.shape, b.shape)
        assert float((a - b).abs().max()) <= 1e-6, (name, float((a - b).abs().max()))
    close(cpu(out["gmap_pos_fts"])[:, :G0], ref["gmap_pos_fts"], "gmap_pos_fts")
    close(cpu(out["gmap_pair_dists"])[:, :G0, :G0], ref["gmap_pair_dists"], "gmap_pair_dists")
    assert torch.equal(cpu(out["gmap_pair_dists"])[:, :G0, :G0], ref["gmap_pair_dists"])  # a gather: exact
    close(cpu(out["vp_pos_fts"]), ref["vp_pos_fts"], "vp_pos_fts")
    close(cpu(out["traj_loc_fts"])[:R], ref["traj_loc_fts"], "traj_loc_fts")
    # ---- the model's index tables == graph_index.build_index on the reference batch --------------------
    n = rix["entries"].numel()
    assert torch.equal(cpu(ix["entries"])[:n], rix["entries"]) and int(cpu(ix["entries"])[n:].abs().sum()) == 0
    if slack == 0:
        assert torch.equal(cpu(ix["node_ptr"]), rix["node_ptr"])
    else:  # wider graph capacity: node n of sample b moves to b * G + n
        got = cpu(ix["node_ptr"])
        for b in range(B):
            assert torch.equal(got[b * G:b * G + G0] - got[b * G], rix["node_ptr"][b * G0:b * G0 + G0] - rix["node_ptr"][b * G0])
        assert int(got[-1]) == n
    ns = rix["src_ids"].numel()
    assert ns == n  # every source feeds exactly one node
    assert torch.equal(cpu(ix["src_ids"])[:ns], rix["src_ids"])
    assert torch.equal(cpu(ix["src_ptr"])[:ns + 1], rix["src_ptr"]) and (cpu(ix["src_ptr"])[ns:] == ns).all()
    assert torch.equal(cpu(ix["src_w"])[:n], rix["src_w"])
    if slack == 0:
        assert torch.equal(cpu(ix["src_nodes"])[:n], rix["src_nodes"])
    for k in ("g_valid", "node2cand"):
        assert torch.equal(cpu(ix[k])[:, :G0], rix[k]), k
    for k in ("l_valid", "bw_mask"):
        assert torch.equal(cpu(ix[k]), rix[k]), k
    assert torch.equal(cpu(ix["vp_gather"]), rix["vp_gather"])
    assert torch.equal(cpu(ix["last_rows"]), rix["last_rows"])
    assert torch.equal(cpu(ix["key_lens_gmap"]), rix["key_lens_gmap"])
    assert torch.equal(cpu(ix["key_lens_vp"]), rix["key_lens_vp"])
    assert torch.equal(cpu(ix["key_lens_pano"])[:R], rix["key_lens_pano"])


def test_model_runs_on_a_device_built_batch():
    """A SAP forward + backward on a batch whose graph half never existed on the host equals the same step on the
    host-built batch (reference loader output + graph_index.build_index)."""
    from magic_b200 import synth
    from magic_b200.config import make_config
    from magic_b200.featurizer import FeatureStore
    from magic_b200.graph_index import batch_to_device
    g, world, rows = load()
    c = g["cases"]["plain"]
    ref, rix = c["batch"], c["index"]
    B = len(c["paths"])
    G = ref["gmap_step_ids"].shape[1]
    feat = GraphFeaturizer(world, B, Tmax=max(len(p) for p in c["paths"]), G=G, R_cap=sum(len(p) for p in c["paths"]),
                           E_cap=rix["entries"].numel(), S_cap=rix["src_ids"].numel())
    paths = [[rows[f"{s}_{v}"] for v in p] for s, p in zip(c["scans"], c["paths"])]
    nxt = [-1 if v is None else rows[f"{s}_{v}"] for s, v in zip(c["scans"], c["next_vp"])]
    dev_b = feat(paths, c["headings"], nxt)
    feat.check()
    attach_text(dev_b, ref["txt_ids"], ref["txt_lens"])
    cfg = make_config(128, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(0)
    model = magic_b200.GlocalTextPathCMTPreTraining(cfg).to(DEV).train()
    store = FeatureStore(synth.make_store(world.N, seed=3, dtype=torch.float32), DEV, dtype=torch.float32).attach(model)
    host_b = dict(ref)
    host_b["traj_vp_index"] = dev_b["traj_vp_index"].cpu().clone()
    host_b["traj_view_perm"] = dev_b["traj_view_perm"].cpu().clone()
    host_b[INDEX_KEY] = dict(rix)
    host_b = batch_to_device(host_b, DEV)
    outs = []
    for b in (host_b, dev_b):
        model.zero_grad(set_to_none=True)
        o = model(b, "sap", True)
        o["loss"][torch.isfinite(o["loss"])].sum().backward()  # (random paths can label an already visited node: inf)
        outs.append((o["loss"].detach().clone(), o["fused_logits"].detach().clone(),
                     model.bert.global_encoder.gmap_pos_embeddings[0].weight.grad.detach().clone()))
    (l0, f0, g0), (l1, f1, g1) = outs
    assert torch.equal(torch.isinf(f0), torch.isinf(f1)) and torch.equal(f0.argmax(1), f1.argmax(1))
    assert torch.allclose(l0, l1, rtol=1e-5, atol=1e-6)
    assert torch.allclose(g0, g1, rtol=1e-3, atol=1e-5)  # (1-ulp differences of the position features)
    assert store.N == world.N
