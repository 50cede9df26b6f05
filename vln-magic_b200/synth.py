"""Synthetic MAGIC pretraining batches (SURVEY.md 8(d) "Synthetic inputs").

Host-side only.  Produces per-sample dicts in the format of the reference task datasets
(`MlmDataset.__getitem__` pretrain_src/data/tasks.py:67-108, `SapDataset.__getitem__` :352-390) and
collates them into the batch schema of `mlm_collate` (:110-166) / `sap_collate` (:392-451):
padded tensors + python lists of viewpoint-id strings.  The graph bookkeeping restates
`get_gmap_inputs` / `get_vp_pos_fts` (pretrain_src/data/dataset.py:513-586).

No dependency on the reference or the oracle: this is what bench.py and the GPU tests feed the model
on the GPU box (where /root/reference does not exist).  tests/test_synth_collate.py checks the collate
against the reference's own collate functions whenever /root/reference is present.
"""
import math
import random

import numpy as np
import torch

VOCAB_RANGE = (1996, 29611)  # tasks.py:59
MASK_ID = 50264
N_VIEWS = 36
IMG_DIM = 768


def random_word(tokens, vocab_range, mask, rng):
    """BERT 15% masking, 80/10/10 (restates tasks.py:11-52 with an explicit `random.Random`)."""
    out_tok, out_lab = [], []
    for tok in tokens:
        p = rng.random()
        if p < 0.15:
            p /= 0.15
            if p < 0.8:
                out_tok.append(mask)
            elif p < 0.9:
                out_tok.append(rng.choice(list(range(*vocab_range))))
            else:
                out_tok.append(tok)
            out_lab.append(tok)
        else:
            out_tok.append(tok)
            out_lab.append(-1)
    if all(o == -1 for o in out_lab):
        out_lab[0] = tokens[0]
        out_tok[0] = mask
    return out_tok, out_lab


def _angle_fts(h, e):
    return np.stack([np.sin(h), np.cos(h), np.sin(e), np.cos(e)], 1).astype(np.float32)


def _make_graph(b, T, G_max, force_full, rs):
    """Path + per-step candidate lists.  Returns (path, cands[t] list of vp-id strings)."""
    path = [f"s{b}_p{t}" for t in range(T)]
    fixed = [(1 if t < T - 1 else 0) + (1 if t > 0 else 0) for t in range(T)]
    K = [int(rs.randint(max(2, fixed[t]), 7)) for t in range(T)]  # K in {2..6} candidates per step
    cap = sum(6 - fixed[t] for t in range(T))
    pool_cap = max(0, G_max - 1 - T)
    F = min(pool_cap, cap) if force_full else int(rs.randint(0, min(pool_cap, cap) + 1))
    while sum(K[t] - fixed[t] for t in range(T)) < F:
        t = int(rs.randint(0, T))
        if K[t] < 6:
            K[t] += 1
    slots = [(t, i) for t in range(T) for i in range(K[t] - fixed[t])]
    new_slots = set(rs.choice(len(slots), size=F, replace=False).tolist()) if F > 0 else set()
    pool = [f"s{b}_f{i}" for i in range(F)]
    used, si, cands = 0, 0, []
    for t in range(T):
        c = []
        if t < T - 1:
            c.append(path[t + 1])
        if t > 0:
            c.append(path[t - 1])
        for _ in range(K[t] - fixed[t]):
            if si in new_slots:
                c.append(pool[used])
                used += 1
            else:  # re-observe a frontier node already seen from an earlier step (or drop the slot)
                choices = [p for p in pool[:used] if p not in c]
                if choices:
                    c.append(choices[int(rs.randint(0, len(choices)))])
            si += 1
        perm = rs.permutation(len(c))
        cands.append([c[i] for i in perm])
    return path, cands


def _gmap(path, cands):
    """dataset.py:513-549 restated: node order [None] + visited (first visit) + frontier (first seen)."""
    visited, unvisited = {}, {}
    for t, vp in enumerate(path):
        visited[vp] = t + 1
        if vp in unvisited:
            del unvisited[vp]
        for nv in cands[t]:
            if nv not in visited:
                unvisited[nv] = 0
    vpids = [None] + list(visited.keys()) + list(unvisited.keys())
    step_ids = [0] + list(visited.values()) + list(unvisited.values())
    vmask = [0] + [1] * len(visited) + [0] * len(unvisited)
    return vpids, step_ids, vmask


def make_store(n_panos=512, seed=77, dtype=torch.bfloat16):
    """A synthetic panorama feature database [n_panos, 36, 768] (unit-scale like CLIP ViT-B/16 features), rounded to
    `dtype` so host-side gathers and the device store agree bit for bit."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_panos, N_VIEWS, IMG_DIM, generator=g).to(dtype)


def make_samples(task, B, L=80, T_max=5, G_max=20, seed=1234, with_labels=None, obj_dim=0, max_objects=6, store=None):
    """task 'og' (object grounding, data/tasks.py:455-501): every panorama additionally carries 0..max_objects object
    tokens after its 36 views (token order [cand_views, noncand_views, objs], data/dataset.py:447,494-508)."""
    rs = np.random.RandomState(seed)
    rng = random.Random(seed)
    samples = []
    lens = rs.randint(L // 2, L + 1, size=B)
    lens[int(rs.randint(0, B))] = L
    Ts = rs.randint(2, T_max + 1, size=B) if T_max >= 2 else np.ones(B, dtype=np.int64)
    full_idx = int(rs.randint(0, B))
    Ts[full_idx] = T_max
    for b in range(B):
        n = int(lens[b])
        toks = [0] + rs.randint(3, 50000, size=n - 2).tolist() + [2]
        T = int(Ts[b])
        path, cands = _make_graph(b, T, G_max, b == full_idx, rs)
        out = {}
        if task == "mlm":
            ids, labs = random_word(toks, VOCAB_RANGE, MASK_ID, rng)
            out["txt_ids"] = torch.LongTensor(ids)
            out["txt_labels"] = torch.LongTensor(labs)
        else:
            out["txt_ids"] = torch.LongTensor(toks)
        if store is None:
            out["traj_view_img_fts"] = [torch.from_numpy(rs.randn(N_VIEWS, IMG_DIM).astype(np.float32))
                                        for _ in range(T)]
        else:
            # featurizer.py: every step names a panorama of the store and the order of its views -- the candidate
            # views first (in candidate order), then the remaining views ascending (dataset.py:742-756)
            vps = rs.randint(0, store.shape[0], size=T)
            perms = []
            for t in range(T):
                cv = rs.choice(N_VIEWS, size=len(cands[t]), replace=False)
                rest = np.setdiff1d(np.arange(N_VIEWS), cv)
                perms.append(np.concatenate([cv, rest]).astype(np.int32))
            out["traj_vp_index"] = [int(v) for v in vps]
            out["traj_view_perm"] = [torch.from_numpy(p_) for p_ in perms]
            out["traj_view_img_fts"] = [store[int(v)][torch.from_numpy(p_).long()].float() for v, p_ in zip(vps, perms)]
        loc, nav = [], []
        n_objs = [int(rs.randint(0, max_objects + 1)) for _ in range(T)] if task == "og" else [0] * T
        if task == "og" and b == full_idx:
            n_objs[-1] = max_objects
        for t in range(T):
            nt = N_VIEWS + n_objs[t]
            ang = _angle_fts(rs.uniform(-math.pi, math.pi, nt), rs.uniform(-math.pi / 6, math.pi / 6, nt))
            box = np.ones((nt, 3), np.float32)
            if n_objs[t]:
                box[N_VIEWS:] = rs.uniform(0.05, 1.0, (n_objs[t], 3)).astype(np.float32)  # dataset.py:489-491
            loc.append(torch.from_numpy(np.concatenate([ang, box], 1)))
            nav.append(torch.LongTensor([1] * len(cands[t]) + [0] * (N_VIEWS - len(cands[t])) + [2] * n_objs[t]))
        if task == "og":
            if obj_dim <= 0:
                raise ValueError("task 'og' needs obj_dim > 0 (config.obj_feat_size)")
            out["traj_obj_img_fts"] = [torch.from_numpy(rs.randn(n, obj_dim).astype(np.float32)) for n in n_objs]
            # dataset.py:487,493: the object-name ids, [0] for a panorama without objects
            out["traj_reverie_obj_names"] = [torch.from_numpy(rs.randint(1, 400, size=n).astype(np.int64)) if n
                                             else torch.zeros(1, dtype=torch.int64) for n in n_objs]
            # get_obj_label (dataset.py:307-320): index among the last panorama's objects, -100 if not present
            out["obj_labels"] = int(rs.randint(0, n_objs[-1])) if (n_objs[-1] and rs.rand() > 0.1) else -100
        if task == "mrc":
            # MrcDataset.__getitem__ (data/tasks.py:220-223): mask >= 1 view of the last panorama, zero its feature,
            # soft labels over image_prob_size classes for every view (no [stop])
            m = rs.rand(N_VIEWS) < 0.15
            if not m.any():
                m[int(rs.randint(0, N_VIEWS))] = True
            mt = torch.from_numpy(m)
            out["traj_view_img_fts"][-1] = out["traj_view_img_fts"][-1].masked_fill(mt[:, None], 0)
            out["vp_view_mrc_masks"] = mt
            pr = torch.softmax(torch.from_numpy(rs.randn(N_VIEWS, 1000).astype(np.float32)) * 2, -1)
            out["vp_view_probs"] = pr
        out["traj_loc_fts"] = loc
        out["traj_nav_types"] = nav
        out["traj_cand_vpids"] = cands
        out["traj_vpids"] = path
        vpids, step_ids, vmask = _gmap(path, cands)
        G = len(vpids)
        out["gmap_vpids"] = vpids
        out["gmap_step_ids"] = torch.LongTensor(step_ids)
        out["gmap_visited_masks"] = torch.BoolTensor(vmask)
        pos = np.concatenate([_angle_fts(rs.uniform(-math.pi, math.pi, G), rs.uniform(-0.5, 0.5, G)),
                              rs.uniform(0, 1, (G, 3)).astype(np.float32)], 1)
        pos[0] = [0, 1, 0, 1, 0, 0, 0]  # [stop] row, dataset.py:556-558
        out["gmap_pos_fts"] = torch.from_numpy(pos.astype(np.float32))
        d = rs.uniform(0, 30, (G, G)).astype(np.float32)
        d = np.triu(d, 1)
        d = d + d.T
        d[0, :] = 0
        d[:, 0] = 0
        out["gmap_pair_dists"] = torch.from_numpy(d)
        K = len(cands[-1])
        vp = np.zeros((N_VIEWS + n_objs[-1] + 1, 14), np.float32)  # vp_ft_len = tokens of the last panorama
        start = np.concatenate([_angle_fts(rs.uniform(-math.pi, math.pi, 1), rs.uniform(-0.5, 0.5, 1)),
                                rs.uniform(0, 1, (1, 3)).astype(np.float32)], 1)
        vp[:, :7] = start
        vp[1:K + 1, 7:] = np.concatenate([_angle_fts(rs.uniform(-math.pi, math.pi, K), rs.uniform(-0.5, 0.5, K)),
                                          rs.uniform(0, 1, (K, 3)).astype(np.float32)], 1)
        out["vp_pos_fts"] = torch.from_numpy(vp)
        out["vp_angles"] = rs.uniform(-math.pi, math.pi, (N_VIEWS, 2)).astype(np.float32)
        if task in ("sap", "cfp") if with_labels is None else with_labels:
            visited = set(path)
            opts = [(0, 0)] + [(vpids.index(c), j + 1) for j, c in enumerate(cands[-1]) if c not in visited]
            g, l = opts[int(rs.randint(0, len(opts)))]
            out["global_act_labels"] = g
            out["local_act_labels"] = l
        samples.append(out)
    return samples


def pad_tensors(tensors):
    """common.py:9-24 semantics."""
    m = max(t.size(0) for t in tensors)
    out = torch.zeros(len(tensors), m, *tensors[0].shape[1:], dtype=tensors[0].dtype)
    for i, t in enumerate(tensors):
        out[i, :t.size(0)] = t
    return out


def _pad_1d(seqs, value):
    m = max(len(s) for s in seqs)
    out = torch.full((len(seqs), m), value, dtype=seqs[0].dtype)
    for i, s in enumerate(seqs):
        out[i, :len(s)] = s
    return out


def collate(samples):
    """Same output schema as the reference mlm_collate / sap_collate (tasks.py:110-166, :392-451)."""
    batch = {k: [x[k] for x in samples] for k in samples[0].keys()}
    batch["txt_lens"] = torch.LongTensor([len(x) for x in batch["txt_ids"]])
    batch["txt_ids"] = _pad_1d(batch["txt_ids"], 0)
    if "txt_labels" in batch:
        batch["txt_labels"] = _pad_1d(batch["txt_labels"], -1)
    batch["traj_step_lens"] = [len(x) for x in batch["traj_view_img_fts"]]
    batch["traj_vp_view_lens"] = torch.LongTensor(sum([[len(y) for y in x] for x in batch["traj_view_img_fts"]], []))
    if "traj_obj_img_fts" in batch:  # og_collate, data/tasks.py:515-524, 557
        batch["traj_vp_obj_lens"] = torch.LongTensor(sum([[len(y) for y in x] for x in batch["traj_obj_img_fts"]], []))
        batch["traj_obj_img_fts"] = pad_tensors(sum(batch["traj_obj_img_fts"], []))
        batch["traj_reverie_obj_names"] = pad_tensors(sum(batch["traj_reverie_obj_names"], []))
        batch["obj_labels"] = torch.LongTensor(batch["obj_labels"])
    batch["traj_view_img_fts"] = pad_tensors(sum(batch["traj_view_img_fts"], []))
    if "traj_vp_index" in batch:  # compact view keys (featurizer.py); not part of the reference schema
        batch["traj_vp_index"] = torch.LongTensor(sum(batch["traj_vp_index"], []))
        batch["traj_view_perm"] = torch.stack(sum(batch["traj_view_perm"], []), 0)
    batch["traj_loc_fts"] = pad_tensors(sum(batch["traj_loc_fts"], []))
    batch["traj_nav_types"] = _pad_1d(sum(batch["traj_nav_types"], []), 0)
    batch["traj_reverie_loc_fts"] = None
    batch["gmap_lens"] = torch.LongTensor([len(x) for x in batch["gmap_step_ids"]])
    batch["gmap_step_ids"] = _pad_1d(batch["gmap_step_ids"], 0)
    batch["gmap_visited_masks"] = _pad_1d(batch["gmap_visited_masks"], False)
    batch["gmap_pos_fts"] = pad_tensors(batch["gmap_pos_fts"])
    G = int(batch["gmap_lens"].max())
    d = torch.zeros(len(samples), G, G)
    for i, x in enumerate(batch["gmap_pair_dists"]):
        d[i, :x.shape[0], :x.shape[1]] = x
    batch["gmap_pair_dists"] = d
    # literal reference behaviour (tasks.py:153): len(x[-1]) of a [Vp,14] tensor == 14; models derive the real
    # local length from traj_vp_view_lens (see graph_index.build_index)
    batch["vp_lens"] = torch.LongTensor([len(x[-1]) for x in batch["vp_pos_fts"]])
    batch["vp_pos_fts"] = pad_tensors(batch["vp_pos_fts"])
    if "vp_view_mrc_masks" in batch:  # mrc_collate, data/tasks.py:296-298
        batch["vp_view_mrc_masks"] = _pad_1d(batch["vp_view_mrc_masks"], False)
        batch["vp_view_probs"] = pad_tensors(batch["vp_view_probs"])
    if "global_act_labels" in batch:
        batch["local_act_labels"] = torch.LongTensor(batch["local_act_labels"])
        batch["global_act_labels"] = torch.LongTensor(batch["global_act_labels"])
    return batch


def make_batch(task, B, L=80, T_max=5, G_max=20, seed=1234, obj_dim=0, max_objects=6, store=None):
    return collate(make_samples(task, B, L, T_max, G_max, seed, obj_dim=obj_dim, max_objects=max_objects, store=store))


def batch_to(batch, device, non_blocking=False):
    return {k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v) for k, v in batch.items()}
