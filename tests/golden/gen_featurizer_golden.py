"""Golden vectors for the GPU batch featuriser's graph half (tests/test_featurize_graph_gpu.py).

Runs the REFERENCE's own loader code -- pretrain_src/data/dataset.py `R2RTextPathData.get_input` (get_cur_angle,
get_traj_pano_fts, get_gmap_inputs, get_gmap_pos_fts, get_vp_pos_fts, get_act_labels), data/tasks.py `SapDataset`
item conversion and `sap_collate` -- on a synthetic world (random connectivity graphs, random candidate tables),
and stores the world, the sampled paths and every produced batch tensor.  The model-side index tables are added by
graph_index.build_index on those reference batches (our host implementation, itself checked against the oracle's
string loops in tests/test_graph_index.py).

Run in the build container (the reference tree is not available on the GPU box):
    python tests/golden/gen_featurizer_golden.py        -> tests/golden/featurizer_graph.pt
"""
import importlib.util
import math
import os
import sys
import types

import networkx as nx
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/pretrain_src"


def ref_modules():
    for stub in ("jsonlines", "h5py", "nltk", "lmdb", "msgpack_numpy", "pynvml"):
        if stub not in sys.modules:
            try:
                __import__(stub)
            except Exception:
                m = types.ModuleType(stub)
                m.patch = lambda *a, **k: None
                sys.modules[stub] = m
    utils = types.ModuleType("utils")
    logger = types.ModuleType("utils.logger")
    logger.LOGGER = types.SimpleNamespace(info=lambda *a, **k: None)
    sys.modules.setdefault("utils", utils)
    sys.modules.setdefault("utils.logger", logger)
    pkg = types.ModuleType("refdata")
    pkg.__path__ = [os.path.join(REF, "data")]
    sys.modules["refdata"] = pkg
    out = {}
    for name in ("common", "dataset", "tasks"):
        spec = importlib.util.spec_from_file_location(f"refdata.{name}", os.path.join(REF, "data", f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refdata.{name}"] = m
        spec.loader.exec_module(m)
        out[name] = m
    return out


def make_world(rng, n_scans=3, n_per=18):
    graphs, cands = {}, {}
    for s in range(n_scans):
        scan = f"scan{s}"
        G = nx.Graph()
        pos = np.concatenate([rng.uniform(-10, 10, (n_per, 2)), rng.uniform(-1.5, 1.5, (n_per, 1))], 1)
        ids = [f"v{s}_{i:02d}" for i in range(n_per)]
        d = np.linalg.norm(pos[:, None] - pos[None], axis=-1)
        edges = set()
        for i in range(n_per):
            for j in np.argsort(d[i])[1:1 + rng.randint(2, 5)]:
                edges.add((min(i, int(j)), max(i, int(j))))
        order = np.argsort(pos[:, 0])
        for a, b in zip(order[:-1], order[1:]):
            edges.add((min(int(a), int(b)), max(int(a), int(b))))
        for a, b in sorted(edges):
            G.add_edge(ids[a], ids[b], weight=float(d[a, b]))
        nx.set_node_attributes(G, values={ids[i]: pos[i] for i in range(n_per)}, name="position")
        graphs[scan] = G
        for i in range(n_per):
            nb = list(G.neighbors(ids[i]))
            rng.shuffle(nb)
            views = rng.permutation(36)[:len(nb)]
            cands[f"{scan}_{ids[i]}"] = {v: [int(p), 0.0, float(rng.uniform(-0.25, 0.25)), float(rng.uniform(-0.2, 0.2))]
                                         for v, p in zip(nb, views)}
    return graphs, cands


def sample_paths(rng, graphs, n, t_max):
    items = []
    for k in range(n):
        scan = sorted(graphs)[k % len(graphs)]
        G = graphs[scan]
        nodes = list(G.nodes)
        path = [nodes[rng.randint(len(nodes))]]
        for _ in range(int(rng.randint(1, t_max))):
            nb = list(G.neighbors(path[-1]))
            path.append(nb[rng.randint(len(nb))])  # revisits happen: the ordered-dict corner cases
        L = int(rng.randint(8, 24))
        items.append(dict(instr_id=f"i{k}", scan=scan, path=path, heading=float(rng.uniform(0, 6.28)),
                          instr_encoding=[0] + rng.randint(3, 50000, L - 2).tolist() + [2]))
    return items


def main(tag="", n_per=18, n_items=12, t_max=9, seed=2024):
    mods = ref_modules()
    ds, tasks, common = mods["dataset"], mods["tasks"], mods["common"]
    rng = np.random.RandomState(seed)
    graphs, cands = make_world(rng, n_per=n_per)
    sd = {s: dict(nx.all_pairs_dijkstra_path_length(G)) for s, G in graphs.items()}
    sp = {s: dict(nx.all_pairs_dijkstra_path(G)) for s, G in graphs.items()}
    feats = {}
    cases = {}
    for name, correct in (("plain", False), ("correct_heading", True)):
        db = ds.R2RTextPathData.__new__(ds.R2RTextPathData)
        db.graphs, db.shortest_distances, db.shortest_paths, db.scanvp_cands = graphs, sd, sp, cands
        db.all_point_rel_angles = [common.get_view_rel_angles(baseViewId=i) for i in range(36)]
        db.angle_feat_size, db.image_feat_size, db.max_txt_len = 4, 16, 100
        db.act_visited_node, db.z_dicts = False, None
        db.args = types.SimpleNamespace(correct_heading=correct)

        def get_feat(scan, vp, type="hdf5"):
            key = f"{scan}_{vp}"
            if key not in feats:
                feats[key] = np.random.RandomState(abs(hash(key)) % (2 ** 31)).randn(36, 16 + 8).astype(np.float32)
            return feats[key]
        db.get_scanvp_feature = get_feat
        db.data = sample_paths(rng, graphs, n_items, t_max)
        sap = tasks.SapDataset.__new__(tasks.SapDataset)
        sap.nav_db = db
        sap.end_vp_pos_ratio = 0.2
        np.random.seed(7)
        ends = []
        orig = db.get_input

        rec = {}
        orig_angle, orig_labels = db.get_cur_angle, db.get_act_labels

        def spy_angle(scan, path, start_heading):   # the path BEFORE the TRAIN_MAX_STEP truncation (dataset.py:660-665)
            rec["prev"] = path[-2] if len(path) > 1 else None
            return orig_angle(scan, path, start_heading)

        def spy_labels(end_vp, end_idx, item, gmap_vpids, traj_cand_vpids):
            rec["end_idx"] = end_idx
            return orig_labels(end_vp, end_idx, item, gmap_vpids, traj_cand_vpids)
        db.get_cur_angle, db.get_act_labels = spy_angle, spy_labels

        def spy(idx, end_vp_type, **kw):
            out = orig(idx, end_vp_type, **kw)
            ends.append((idx, rec["end_idx"], rec["prev"]))
            return out
        db.get_input = spy
        samples = [sap[i] for i in range(len(db.data))]
        batch = tasks.sap_collate([dict(s) for s in samples])
        # what the device featuriser receives: the truncated path (as taken by get_input) and the next gt viewpoint
        paths, nxt, prevs = [], [], []
        for (idx, end_idx, prev), s in zip(ends, samples):
            item = db.data[idx]
            paths.append(list(s["traj_vpids"]))
            full = item["path"]
            # R2R get_act_labels (dataset.py:622-640) decides "stop" by VALUE: a prefix that ends on the final viewpoint
            # (the path revisits it) is labelled stop as well
            nxt.append(None if s["traj_vpids"][-1] == full[-1] else full[end_idx + 1])
            # the heading comes from the TRUE predecessor of the end viewpoint; a path longer than TRAIN_MAX_STEP is cut
            # to its first 20 viewpoints + the end viewpoint afterwards, so traj_vpids[-2] is not it
            prevs.append(prev)
        from magic_b200.graph_index import build_index
        index = build_index(batch)
        keep = {k: v for k, v in batch.items() if torch.is_tensor(v) and k != "traj_view_img_fts"}
        keep.update(traj_step_lens=batch["traj_step_lens"], gmap_vpids=batch["gmap_vpids"],
                    traj_cand_vpids=batch["traj_cand_vpids"], traj_vpids=batch["traj_vpids"])
        cases[name] = dict(batch=keep, index={k: v for k, v in index.items()}, paths=paths, next_vp=nxt, prev_vp=prevs,
                           scans=[db.data[e[0]]["scan"] for e in ends], headings=[db.data[e[0]]["heading"] for e in ends],
                           correct_heading=correct)
    world = dict(
        nodes={s: list(G.nodes) for s, G in graphs.items()},
        pos={s: {v: np.asarray(G.nodes[v]["position"], dtype=np.float64) for v in G.nodes} for s, G in graphs.items()},
        dist=sd, paths_len={s: {a: {b: len(p) for b, p in d.items()} for a, d in sp[s].items()} for s in sp},
        cands=cands, view_ang=common.get_view_rel_angles(baseViewId=12) if False else
        [common.get_view_rel_angles(baseViewId=i) for i in range(36)][12])
    out = os.path.join(ROOT, "tests", "golden", f"featurizer_graph{tag}.pt")
    torch.save(dict(world=world, cases=cases, numpy=np.__version__), out)
    print("wrote", out, os.path.getsize(out), "bytes;", {k: len(v["paths"]) for k, v in cases.items()})


if __name__ == "__main__":
    main()
    # long-horizon case (RxR-like): 40-viewpoint scans, paths of up to 21 viewpoints (TRAIN_MAX_STEP = 20 + the end
    # viewpoint, dataset.py:663-665), graphs of 40+ nodes, many revisits
    main(tag="_long", n_per=40, n_items=16, t_max=22, seed=77)
