"""Fine-tune / inference path: the model half of the reference's navigation agent.

The reference agent (`map_nav_src/r2r/agent.py`) drives a `VLNBert` in three modes over an online `GraphMap`:

    txt_embeds, txt_attns                              = vln_bert('language',   language_inputs)   agent.py:797
    pano_embeds, pano_masks, pano_fused_embeds, attns  = vln_bert('panorama',   pano_inputs)       agent.py:885
    nav_outs (dict: gmap_embeds, vp_embeds, *_logits,
              cls_embeds, gmap_attns, vp_attns)        = vln_bert('navigation', nav_inputs)        agent.py:964-967

`models/model.py` (VLNBert) and `models/graph_utils.py` (GraphMap) are imported at agent.py:29-30 but are NOT in the
reference tree, so -- exactly like the pretraining model -- this file restates them from their call sites:

  * `VLNBert`      the three modes on the SAME hand-written kernels as pretraining (ops.py -> libmagic_b200), at
                   batch 1-16; parameters live in a `GlocalTextPathCMTPreTraining`, so a pretraining checkpoint
                   loads unchanged (agent_base.py:298-359 reconciles prefixes the same way)
  * `GraphMap`     online topological map: incremental all-pairs shortest paths (dense matrices, vectorised
                   relaxation), running-mean node embeddings kept ON THE DEVICE, position features of
                   pretrain_src/data/dataset.py:553-586 (same arithmetic as pretraining, pinned by that file)
  * collators      `language_inputs`, `panorama_inputs`, `nav_gmap_inputs`, `nav_vp_inputs_mem` produce the dicts of
                   agent.py:60-98, :108-173, :175-245, :289-328 (keys, [stop] / [MEM] layout, masks); a CPU test runs
                   the reference's own collator source against them

[DECISION]s (the files are absent upstream): with the [MEM] slot (`enc_full_graph`, agent.py:197-200) the local ->
global logit fusion skips BOTH special slots (j >= 2; a literal j > 0 would add the masked -inf [MEM] logit to every
unvisited node); `cls_embeds` = global [MEM] output + local [MEM] output (the two branches are summed, as in the MLM
text branch).
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .model import GlocalTextPathCMTPreTraining, _Ctx, _cross_encoder, stack_attns

MAX_DIST = 30.0  # pretrain_src/data/dataset.py:21-22
MAX_STEP = 10.0


# ---------------------------------------------------------------------------------------------------
# online graph
# ---------------------------------------------------------------------------------------------------
class FloydGraph:
    """All-pairs shortest paths of a growing undirected graph.  Interface from the call sites agent.py:192,220
    (`visited`, `distance`) and the DUET-lineage map (`add_edge`, `update`, `path`).  Dense [cap, cap] distance and
    next-hop matrices; `update(k)` relaxes every pair through k with one vectorised minimum."""

    def __init__(self, cap=64):
        self._ids, self._names = {}, []
        self._d = np.full((cap, cap), np.inf, dtype=np.float64)
        self._nxt = np.full((cap, cap), -1, dtype=np.int64)
        self._visited = set()

    def _idx(self, k):
        i = self._ids.get(k)
        if i is None:
            i = self._ids[k] = len(self._names)
            self._names.append(k)
            if i >= self._d.shape[0]:
                cap = 2 * self._d.shape[0]
                d = np.full((cap, cap), np.inf)
                n = np.full((cap, cap), -1, dtype=np.int64)
                d[:i, :i], n[:i, :i] = self._d[:i, :i], self._nxt[:i, :i]
                self._d, self._nxt = d, n
            self._d[i, i] = 0.0
        return i

    def visited(self, k):
        return k in self._visited

    def distance(self, x, y):
        if x == y:
            return 0.0
        return float(self._d[self._idx(x), self._idx(y)])

    def add_edge(self, x, y, dis):
        i, j = self._idx(x), self._idx(y)
        if dis < self._d[i, j]:
            self._d[i, j] = self._d[j, i] = dis
            self._nxt[i, j], self._nxt[j, i] = j, i

    def update(self, k):
        """k becomes a visited node: every pair may now route through it."""
        n, kk = len(self._names), self._idx(k)
        d = self._d[:n, :n]
        via = d[:, kk, None] + d[None, kk, :]
        better = via < d
        if better.any():
            d[better] = via[better]
            nx = self._nxt[:n, :n]
            nx[better] = np.broadcast_to(nx[:, kk, None], (n, n))[better]
        self._visited.add(k)

    def path(self, x, y):
        """Nodes after x up to and including y ([] when x == y), like the lineage's recursive path()."""
        if x == y:
            return []
        i, j = self._idx(x), self._idx(y)
        out = []
        while i != j:
            i = int(self._nxt[i, j])
            if i < 0:
                raise KeyError(f"no path {x} -> {y}")
            out.append(self._names[i])
        return out


def rel_pos_fts(a, b, base_heading=0.0, base_elevation=0.0):
    """heading / elevation / distance of b seen from a -- the arithmetic of pretrain_src/data/common.py:142-160
    (the simulator's x-y axes are transposed: heading = asin(dx / |xy|), mirrored when b lies behind)."""
    dx, dy, dz = b[0] - a[0], b[1] - a[1], b[2] - a[2]
    xy = max(math.sqrt(dx * dx + dy * dy), 1e-8)
    xyz = max(math.sqrt(dx * dx + dy * dy + dz * dz), 1e-8)
    heading = math.asin(dx / xy)
    if b[1] < a[1]:
        heading = math.pi - heading
    return heading - base_heading, math.asin(dz / xyz) - base_elevation, xyz


def angle_fts(headings, elevations, angle_feat_size=4):
    """[sin h, cos h, sin e, cos e] x (size / 4) -- map_nav_src/utils/data.py:176-182."""
    f = np.stack([np.sin(headings), np.cos(headings), np.sin(elevations), np.cos(elevations)], 1).astype(np.float32)
    return np.concatenate([f] * (angle_feat_size // 4), 1) if angle_feat_size > 4 else f


class GraphMap:
    """One episode's topological map (agent.py:755-757 builds one per observation and calls `update_graph(ob)`).

    Node embeddings: visited nodes hold the fused panorama embedding of the visit (`rewrite=True`), unvisited nodes
    the running mean of every partial view that saw them (agent.py:905-924).  Sums and counts live in device
    slabs (student and teacher separately), so `node_embeds(vpids)` is one indexed read -- no per-node python
    stacking on the step path."""

    def __init__(self, start_vp, cap=64):
        self.start_vp = start_vp
        self.node_positions = {}
        self.graph = FloydGraph(cap)
        self.node_step_ids = {}
        self.node_stop_scores = {}
        self._slot = {}
        self._sum = {False: None, True: None}
        self._cnt = {False: None, True: None}
        self._cap = cap

    # -- topology --------------------------------------------------------------------------------
    def update_graph(self, ob):
        """`ob`: dict with viewpoint, position, candidate = [{viewpointId, position}] (env.py observation schema)."""
        cur = ob["viewpoint"]
        self.node_positions[cur] = np.asarray(ob["position"], dtype=np.float64)
        for cc in ob["candidate"]:
            v = cc["viewpointId"]
            self.node_positions[v] = np.asarray(cc["position"], dtype=np.float64)
            dist = float(np.linalg.norm(self.node_positions[cur] - self.node_positions[v]))
            self.graph.add_edge(cur, v, dist)
        self.graph.update(cur)

    # -- embeddings ------------------------------------------------------------------------------
    def _slab(self, teacher, like):
        if self._sum[teacher] is None:
            self._sum[teacher] = torch.zeros(self._cap, like.shape[-1], dtype=torch.float32, device=like.device)
            self._cnt[teacher] = torch.zeros(self._cap, dtype=torch.float32, device=like.device)
        return self._sum[teacher], self._cnt[teacher]

    def _row(self, vp):
        r = self._slot.get(vp)
        if r is None:
            r = self._slot[vp] = len(self._slot)
            if r >= self._cap:
                self._cap *= 2
                for t in (False, True):
                    if self._sum[t] is not None:
                        s = torch.zeros(self._cap, self._sum[t].shape[1], dtype=torch.float32, device=self._sum[t].device)
                        c = torch.zeros(self._cap, dtype=torch.float32, device=s.device)
                        s[:r], c[:r] = self._sum[t][:r], self._cnt[t][:r]
                        self._sum[t], self._cnt[t] = s, c
        return r

    def update_node_embed(self, vp, embed, rewrite=False, teacher=False):
        s, c = self._slab(teacher, embed)
        r = self._row(vp)
        if r >= s.shape[0]:
            s, c = self._slab(teacher, embed)
        e = embed.detach().float()
        if rewrite:
            s[r], c[r] = e, 1.0
        else:
            s[r] += e
            c[r] += 1.0

    def get_node_embed(self, vp, teacher=False):
        s, c = self._sum[teacher], self._cnt[teacher]
        r = self._slot[vp]
        return s[r] / c[r]

    def node_embeds(self, vpids, teacher=False):
        """[len(vpids), h] means in one gather (the batched form of get_node_embed)."""
        s, c = self._sum[teacher], self._cnt[teacher]
        rows = torch.as_tensor([self._slot[v] for v in vpids], dtype=torch.int64, device=s.device)
        return s[rows] / c[rows, None]

    # -- position features -----------------------------------------------------------------------
    def get_pos_fts(self, cur_vp, gmap_vpids, cur_heading, cur_elevation, angle_feat_size=4):
        """[N, 7] = [sin h, cos h, sin e, cos e, line_dist / 30, shortest_dist / 30, shortest_steps / 10]; a None id
        ([stop] / [MEM]) gives a zero angle and zero distances (dataset.py:553-575)."""
        ang = np.zeros((len(gmap_vpids), 2), dtype=np.float32)
        dist = np.zeros((len(gmap_vpids), 3), dtype=np.float32)
        a = self.node_positions[cur_vp]
        for i, vp in enumerate(gmap_vpids):
            if vp is None:
                continue
            h, e, d = rel_pos_fts(a, self.node_positions[vp], cur_heading, cur_elevation)
            ang[i] = (h, e)
            dist[i] = (d / MAX_DIST, self.graph.distance(cur_vp, vp) / MAX_DIST,
                       len(self.graph.path(cur_vp, vp)) / MAX_STEP)
        return np.concatenate([angle_fts(ang[:, 0], ang[:, 1], angle_feat_size), dist], 1)


# ---------------------------------------------------------------------------------------------------
# collators (dict keys and layouts of agent.py)
# ---------------------------------------------------------------------------------------------------
def _pad(ts, dtype=None):
    n = max(t.shape[0] for t in ts)
    out = torch.zeros(len(ts), n, *ts[0].shape[1:], dtype=dtype or ts[0].dtype, device=ts[0].device)
    for i, t in enumerate(ts):
        out[i, :t.shape[0]] = t
    return out


def seq_masks(lens, max_len=None):
    lens = torch.as_tensor(lens)
    max_len = int(lens.max()) if max_len is None else max_len
    return torch.arange(max_len, device=lens.device)[None, :] < lens[:, None]


def language_inputs(obs, device):
    """agent.py:60-98 (`_language_variable`, without the intervention dictionaries): padded ids + masks."""
    lens = [len(ob["instr_encoding"]) for ob in obs]
    ids = torch.zeros(len(obs), max(lens), dtype=torch.int64)
    for i, ob in enumerate(obs):
        ids[i, :lens[i]] = torch.as_tensor(ob["instr_encoding"], dtype=torch.int64)
    return {"txt_ids": ids.to(device), "txt_masks": seq_masks(lens).to(device),
            "txt_lens": torch.as_tensor(lens, dtype=torch.int64, device=device)}


def panorama_inputs(obs, device, image_feat_size=768):
    """agent.py:108-173: candidate views first (nav type 1), then the remaining views of the 36; loc = angle
    features + a [1, 1, 1] box."""
    img, loc, nav, cand, lens = [], [], [], [], []
    for ob in obs:
        f_img, f_ang, types, ids, used = [], [], [], [], set()
        for cc in ob["candidate"]:
            f_img.append(cc["feature"][:image_feat_size])
            f_ang.append(cc["feature"][image_feat_size:])
            types.append(1)
            ids.append(cc["viewpointId"])
            used.add(cc["pointId"])
        for k, x in enumerate(ob["feature"]):
            if k not in used:
                f_img.append(x[:image_feat_size])
                f_ang.append(x[image_feat_size:])
        types.extend([0] * (36 - len(used)))
        f_img, f_ang = np.stack(f_img, 0), np.stack(f_ang, 0)
        box = np.ones((len(f_img), 3), dtype=np.float32)
        img.append(torch.from_numpy(f_img.astype(np.float32)))
        loc.append(torch.from_numpy(np.concatenate([f_ang, box], 1).astype(np.float32)))
        nav.append(torch.as_tensor(types, dtype=torch.int64))
        cand.append(ids)
        lens.append(len(f_img))
    return {"view_img_fts": _pad(img).to(device), "loc_fts": _pad(loc).to(device), "nav_types": _pad(nav).to(device),
            "view_lens": torch.as_tensor(lens, dtype=torch.int64, device=device), "cand_vpids": cand}


def nav_gmap_inputs(obs, gmaps, last_embeds=None, teacher=False, hidden=None, device=None):
    """agent.py:175-245 with `enc_full_graph` (node order [stop], [MEM], visited, unvisited): step ids, running-mean
    node embeddings ([stop] = 0, [MEM] = last step's cls_embeds or 0), position features, pair distances, masks
    ([MEM] can never be attended or chosen: gmap_masks[:, 1] = False)."""
    vpids_b, lens, embeds, steps, pos, pair, vis, none_left = [], [], [], [], [], [], [], []
    for i, gmap in enumerate(gmaps):
        visited = [k for k in gmap.node_positions if gmap.graph.visited(k)]
        unvisited = [k for k in gmap.node_positions if not gmap.graph.visited(k)]
        none_left.append(len(unvisited) == 0)
        vpids = [None, None] + visited + unvisited
        nodes = gmap.node_embeds(vpids[2:], teacher)
        mem = torch.zeros_like(nodes[0]) if last_embeds is None else last_embeds[i].detach().float()
        embeds.append(torch.cat([torch.zeros_like(nodes[:1]), mem[None], nodes], 0))
        steps.append(torch.as_tensor([gmap.node_step_ids.get(vp, 0) for vp in vpids], dtype=torch.int64))
        pos.append(torch.from_numpy(gmap.get_pos_fts(obs[i]["viewpoint"], vpids, obs[i]["heading"], obs[i]["elevation"])))
        n = len(vpids)
        d = np.zeros((n, n), dtype=np.float32)
        for a in range(2, n):
            for b in range(a + 1, n):
                d[a, b] = d[b, a] = gmap.graph.distance(vpids[a], vpids[b])
        pair.append(torch.from_numpy(d))
        vis.append(torch.as_tensor([0, 1] + [1] * len(visited) + [0] * len(unvisited), dtype=torch.bool))
        vpids_b.append(vpids)
        lens.append(n)
    device = device or embeds[0].device
    G = max(lens)
    masks = seq_masks(lens, G)
    masks[:, 1] = False
    dists = torch.zeros(len(obs), G, G)
    for i, d in enumerate(pair):
        dists[i, :lens[i], :lens[i]] = d
    return {"gmap_vpids": vpids_b, "gmap_img_embeds": _pad(embeds).to(device), "gmap_step_ids": _pad(steps).to(device),
            "gmap_pos_fts": _pad(pos).to(device), "gmap_visited_masks": _pad(vis).to(device),
            "gmap_pair_dists": dists.to(device), "gmap_masks": masks.to(device), "no_vp_left": none_left,
            "gmap_lens": torch.as_tensor(lens, dtype=torch.int64, device=device)}


def nav_vp_inputs_mem(obs, gmaps, pano_embeds, cand_vpids, view_lens, nav_types, last_embeds=None):
    """agent.py:289-328: local tokens = [stop], [MEM], views; position features = start-relative (7) + candidate-
    relative (7); only [stop] and candidate views are valid actions."""
    B = len(obs)
    mem = torch.zeros_like(pano_embeds[:, :1]) if last_embeds is None else last_embeds.to(pano_embeds.dtype).unsqueeze(1)
    emb = torch.cat([torch.zeros_like(pano_embeds[:, :1]), mem, pano_embeds], 1)
    Vp = emb.shape[1]
    pos = np.zeros((B, Vp, 14), dtype=np.float32)
    for i, gmap in enumerate(gmaps):
        ob = obs[i]
        cf = gmap.get_pos_fts(ob["viewpoint"], cand_vpids[i], ob["heading"], ob["elevation"])
        sf = gmap.get_pos_fts(ob["viewpoint"], [gmap.start_vp], ob["heading"], ob["elevation"])
        pos[i, :, :7] = sf
        pos[i, 2:2 + len(cf), 7:] = cf
    dev = pano_embeds.device
    ones, zeros = torch.ones(B, 1, dtype=torch.bool, device=dev), torch.zeros(B, 1, dtype=torch.bool, device=dev)
    return {"vp_img_embeds": emb, "vp_pos_fts": torch.from_numpy(pos).to(dev),
            "vp_masks": seq_masks(view_lens + 2, Vp), "vp_nav_masks": torch.cat([ones, zeros, nav_types == 1], 1),
            "vp_cand_vpids": [[None, None] + list(x) for x in cand_vpids], "vp_lens": view_lens + 2}


def nav_index(nav_inputs, n_special=2):
    """Integer tables of the local -> global logit fusion for `ops.sap_fuse` (same kernel as pretraining's SAP head):
    g_valid = reachable, unvisited nodes; l_valid = [stop] + candidate views; node2cand[b, n] = local slot whose logit
    joins node n; bw_mask = candidates that lead back to visited nodes (their logits sum into every other unvisited
    node).  Exact integer logic on the viewpoint-id strings; built on the host once per step."""
    gv, vc = nav_inputs["gmap_vpids"], nav_inputs["vp_cand_vpids"]
    B = len(gv)
    G, Vp = nav_inputs["gmap_masks"].shape[1], nav_inputs["vp_nav_masks"].shape[1]
    vis = nav_inputs["gmap_visited_masks"].cpu().numpy().astype(bool)
    gm = nav_inputs["gmap_masks"].cpu().numpy().astype(bool)
    g_valid = (gm & ~vis).astype(np.uint8)
    l_valid = nav_inputs["vp_nav_masks"].cpu().numpy().astype(np.uint8)
    node2cand = np.full((B, G), -1, dtype=np.int32)
    bw_mask = np.zeros((B, Vp), dtype=np.uint8)
    for b in range(B):
        visited = set(vp for vp, m in zip(gv[b], vis[b].tolist()) if m)
        tmp = {}
        for j, c in enumerate(vc[b]):
            if j < n_special:
                continue
            if c in visited:
                bw_mask[b, j] = 1
            else:
                tmp[c] = j  # the last candidate with this id wins (dict overwrite)
        for n, vp in enumerate(gv[b]):
            if n > 0 and vp not in visited and vp in tmp:
                node2cand[b, n] = tmp[vp]
    dev = nav_inputs["gmap_masks"].device
    t = lambda a: torch.from_numpy(a).to(dev)
    return {"g_valid": t(g_valid), "l_valid": t(l_valid), "node2cand": t(node2cand), "bw_mask": t(bw_mask)}


# ---------------------------------------------------------------------------------------------------
# model
_TOKENIZER_VOCAB = {"bert": (30522, 2, 512), "roberta": (50265, 1, 514), "xlm": (250002, 1, 514)}


def config_from_agent_args(args, role="student"):
    """The model config of one role from the fine-tune agent's argparse namespace, the object the agent hands to
    `VLNBert(self.args, role=...)` (agent.py:36-38; flags: r2r/parser.py:56-66, 118, 173-193, scripts/run_r2r_kdl_valid.sh).
    `<role>_hidden_size / _num_l_layers / _num_pano_layers / _num_x_layers / _mlp_ratio` give the sizes the way
    train_r2r_magic.py:141-160 derives them for pretraining (intermediate = hidden * mlp_ratio, heads = hidden / 64);
    `train_kdl` switches the KD outputs on and gives the student its up-projections to `teacher_hidden_size`.
    [INFERRED] (models/model.py is absent upstream): dropout = `--dropout` for hidden and attention probabilities;
    vocabulary / type / position sizes follow `--tokenizer`; R2R / RxR carry no object tokens (obj_feat_size 0).
    The causal-intervention dictionaries (`do_back_*`, `do_front_*`) are outside this path and ignored."""
    from .config import make_config
    pre = "teacher" if role == "teacher" else "student"

    def g(name, default=None):
        return getattr(args, f"{pre}_{name}", getattr(args, name, default))

    kd = bool(getattr(args, "train_kdl", False))
    vocab, types, positions = _TOKENIZER_VOCAB[getattr(args, "tokenizer", "roberta")]
    if any(getattr(args, k, False) for k in ("do_back_img", "do_back_txt", "do_front_img", "do_front_his", "do_front_txt")):
        import warnings
        warnings.warn("magic_b200.nav.VLNBert: the causal-intervention inputs (do_back_* / do_front_*) are not part of "
                      "this path and are ignored")
    kdl = None
    if kd and pre == "student":
        kdl = dict(kdl_adaptive_ability_weight=bool(getattr(args, "kdl_adaptive_ability_weight", False)),
                   kdl_adaptive_ability_weight_type=getattr(args, "kdl_adaptive_ability_weight_type", "RW"))
    over = dict(kdl=kdl) if kdl is not None else {}
    p = float(getattr(args, "dropout", 0.1))
    return make_config(
        int(g("hidden_size")), int(g("num_l_layers", 6)), int(g("num_x_layers", 3)), int(g("num_pano_layers", 2)),
        mlp_ratio=g("mlp_ratio", 4), role=pre, teacher_hidden_size=int(args.teacher_hidden_size) if kd and pre == "student" else None,
        pretrain_tasks=("sap",), hidden_dropout_prob=p, attention_probs_dropout_prob=p,
        image_feat_size=int(getattr(args, "image_feat_size", 768)), angle_feat_size=int(getattr(args, "angle_feat_size", 4)),
        obj_feat_size=0, graph_sprels=bool(getattr(args, "graph_sprels", True)),
        glocal_fuse=getattr(args, "fusion", "dynamic") == "dynamic",
        adaptive_pano_fusion=bool(getattr(args, "adaptive_pano_fusion", True)),
        cfp_temperature=float(getattr(args, "cfp_temperature", 1.0)), vocab_size=vocab, type_vocab_size=types,
        max_position_embeddings=positions, kd=kd, **over)


def remap_agent_keys(state_dict, own_keys):
    """Checkpoint keys of the other layouts this model meets -> ours (`vln_bert.bert.<encoders, KD heads>`,
    `vln_bert.<task heads>`).  A key that already matches stays; otherwise
      * a DDP `module.` prefix is dropped (agent_base.py:336-339);
      * `vln_bert.X` becomes `vln_bert.bert.X` -- [INFERRED] the layout of the reference's fine-tune checkpoints: its
        agent reaches the KD heads and learned weights directly on the wrapped model (`vln_bert.vln_bert.txt_emb_w`,
        agent.py:552-568; `k.split('.')[1] in ['txt_emb_w', ...]`, agent_base.py:328-330), i.e. without a `.bert` level;
      * a pretraining key (`bert.X`, `global_sap_head.X`; train_r2r_magic.py:189-208) gains the `vln_bert.` prefix.
    Keys that match nothing are passed through for `load_state_dict` to report."""
    own = set(own_keys)
    out = type(state_dict)() if isinstance(state_dict, dict) else {}
    if hasattr(state_dict, "_metadata"):
        out._metadata = state_dict._metadata  # torch's per-module version records travel with the OrderedDict
    for k, v in state_dict.items():
        if k not in own:
            k2 = k[7:] if k.startswith("module.") else k
            if k2 in own:
                k = k2
            elif k2.startswith("vln_bert.") and "vln_bert.bert." + k2[9:] in own:
                k = "vln_bert.bert." + k2[9:]
            elif "vln_bert." + k2 in own:
                k = "vln_bert." + k2
            elif "vln_bert.bert." + k2 in own:
                k = "vln_bert.bert." + k2
        out[k] = v
    return out


# ---------------------------------------------------------------------------------------------------
class VLNBert(nn.Module):
    """`VLNBert(config, role)`; `forward(mode, batch)` with mode in {'language', 'panorama', 'navigation'}
    (agent.py:36-39, :797, :885, :964).  Holds a `GlocalTextPathCMTPreTraining` (attribute `vln_bert`, so checkpoint
    keys read `vln_bert.bert.*` / `vln_bert.global_sap_head.*`); all math runs in libmagic_b200."""

    def __init__(self, config, role="student"):
        super().__init__()
        if not hasattr(config, "hidden_size") and hasattr(config, "student_hidden_size"):
            config = config_from_agent_args(config, role)  # the agent's `VLNBert(self.args, role=...)`, agent.py:36-38
        if not hasattr(config, "pretrain_tasks"):
            config.pretrain_tasks = ("sap",)
        config.role = role
        self.config = config
        self.vln_bert = GlocalTextPathCMTPreTraining(config)
        self.want_attn = bool(getattr(config, "kd", False))

    @classmethod
    def from_pretraining(cls, model, role=None):
        """Wrap an existing pretraining model (shares its parameters, arena and compute dtype)."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.config = model.config
        self.vln_bert = model
        self.want_attn = bool(getattr(model.config, "kd", False))
        return self

    def set_compute_dtype(self, dtype):
        self.vln_bert.set_compute_dtype(dtype)
        return self

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Also accepts the agent-level layouts `remap_agent_keys` lists (DDP prefix, the reference's fine-tune layout,
        a bare pretraining checkpoint)."""
        return super().load_state_dict(remap_agent_keys(state_dict, self.state_dict().keys()), strict=strict, **kw)

    def _fc(self):
        m = self.vln_bert
        base = 0 if getattr(self.config, "role", "student") == "student" else 1 << 20
        return _Ctx(m.compute_dtype, self.training, self.want_attn, base)

    def forward(self, mode, batch):
        arena = getattr(self.vln_bert, "_magic_arena", None)
        if arena is not None and not torch.cuda.is_current_stream_capturing():
            arena.sync_lowp()
        if mode == "language":
            return self.forward_language(batch)
        if mode == "panorama":
            return self.forward_panorama(batch)
        if mode == "navigation":
            return self.forward_navigation(batch)
        raise NotImplementedError(f"wrong mode: {mode}")

    @staticmethod
    def _lens(batch, key_lens, key_masks):
        if batch.get(key_lens) is not None:
            return batch[key_lens].to(torch.int32)
        return batch[key_masks].sum(1).to(torch.int32)

    def forward_language(self, batch):
        fc = self._fc()
        B, L = batch["txt_ids"].shape
        ix = {"key_lens_txt": self._lens(batch, "txt_lens", "txt_masks")}
        x, attns = self.vln_bert.bert.forward_text(batch, ix, fc)
        return x.view(B, L, -1), stack_attns(attns) if self.want_attn else None

    def forward_panorama(self, batch):
        fc = self._fc()
        lens = batch["view_lens"]
        pb = {"traj_view_img_fts": batch["view_img_fts"], "traj_loc_fts": batch["loc_fts"],
              "traj_nav_types": batch["nav_types"], "traj_vp_view_lens": lens}
        ix = {"key_lens_pano": lens.to(torch.int32)}
        pano, fused, attns = self.vln_bert.bert.forward_pano(pb, ix, fc)
        return pano, seq_masks(lens, pano.shape[1]), fused, stack_attns(attns) if self.want_attn else None

    @staticmethod
    def mem_slot(batch):
        """1 when the batch uses the [stop], [MEM], ... layout (`enc_full_graph`), else -1."""
        if batch.get("mem_slot") is not None:
            return int(batch["mem_slot"])
        vp = batch.get("gmap_vpids")
        return 1 if (vp is not None and len(vp[0]) > 1 and vp[0][1] is None) else -1

    def forward_navigation(self, batch, index=None):
        m, fc = self.vln_bert, self._fc()
        bert = m.bert
        ge, le = bert.global_encoder, bert.local_encoder
        txt = batch["txt_embeds"]
        B, Lt, h = txt.shape
        G, Vp = batch["gmap_step_ids"].shape[1], batch["vp_pos_fts"].shape[1]
        txt2 = txt.reshape(B * Lt, h)
        if txt2.dtype != fc.dtype:
            txt2 = txt2.to(fc.dtype)
        k_txt = self._lens(batch, "txt_lens", "txt_masks")
        # gmap_masks is a prefix mask with ONE hole: the [MEM] slot (agent.py:228) lies inside the valid prefix but is
        # never a key.  The kernels take the prefix as a key length and the hole as `key_skip` (magic_attn_set_key_skip)
        mem = self.mem_slot(batch)
        k_g = self._lens(batch, "gmap_lens", "gmap_masks")
        if batch.get("gmap_lens") is None and mem >= 0:
            k_g = k_g + 1  # the masked-out [MEM] slot still occupies a position
        k_v = self._lens(batch, "vp_lens", "vp_masks")
        pe = ge.gmap_pos_embeddings
        g_in = ops.posfuse(batch["gmap_img_embeds"].reshape(B * G, h).to(fc.dtype), batch["gmap_step_ids"].reshape(-1),
                           ge.gmap_step_embeddings.weight, None, batch["gmap_pos_fts"].reshape(B * G, -1), pe[0].weight,
                           pe[0].bias, pe[1].weight, pe[1].bias, pe[1].eps, fc.dtype)
        pv = le.vp_pos_embeddings
        v_in = ops.posfuse(batch["vp_img_embeds"].reshape(B * Vp, h).to(fc.dtype), None, None, None,
                           batch["vp_pos_fts"].reshape(B * Vp, -1), pv[0].weight, pv[0].bias, pv[1].weight, pv[1].bias,
                           pv[1].eps, fc.dtype)
        dists = batch["gmap_pair_dists"] if ge.sprel_linear is not None else None
        (v, v_attn), (g, g_attn) = ops.run_branches(
            lambda: _cross_encoder(le.encoder, v_in, txt2, B, Vp, Lt, k_v, k_txt, fc),
            lambda: _cross_encoder(ge.encoder, g_in, txt2, B, G, Lt, k_g, k_txt, fc, dists, ge.sprel_linear, key_skip=mem))
        g3, v3 = g.view(B, G, h), v.view(B, Vp, h)
        ix = index if index is not None else nav_index(batch, n_special=2 if mem >= 0 else 1)
        ix = dict(ix)
        ix["stop_rows_g"] = torch.arange(B, dtype=torch.int64, device=txt.device) * G
        ix["stop_rows_v"] = torch.arange(B, dtype=torch.int64, device=txt.device) * Vp
        gl, ll, fl = m.sap_logits({"gmap_embeds": g3, "vp_embeds": v3}, ix)
        out = {"gmap_embeds": g3, "vp_embeds": v3, "global_logits": gl, "local_logits": ll, "fused_logits": fl,
               "cls_embeds": ops.add(g3[:, max(mem, 0)].contiguous(), v3[:, max(mem, 0)].contiguous())}
        if self.want_attn:
            out["gmap_attns"], out["vp_attns"] = stack_attns(g_attn), stack_attns(v_attn)
        return out


# ---------------------------------------------------------------------------------------------------
# graph-replayed decision step
# ---------------------------------------------------------------------------------------------------
class NavStepper:
    """The panorama and navigation modes of a `VLNBert` as two CUDA graphs over fixed-capacity input buffers.

    At batch 1-16 a navigation decision is ~190 kernels of a few microseconds each: launched eagerly from Python the
    host is the bottleneck (3.6 ms per navigation forward of MAGIC-S on B200), replayed as a graph the device time is
    what remains.  Inputs of any size up to the capacities (`G` graph nodes, `Lt` text tokens, 36 views + [stop] +
    [MEM]) are padded into the static buffers -- padded graph nodes / text tokens lie beyond the key lengths and get
    -inf logits, exactly like the padding the agent's own collators produce (agent.py:225-245) -- and the outputs are
    sliced back.  Inference only (no autograd through a replay)."""

    V, VP = 36, 38

    def __init__(self, model, B, G=64, Lt=80):
        self.model, self.B, self.G, self.Lt = model, B, G, Lt
        m = model.vln_bert
        dev = next(m.parameters()).device
        h = m.config.hidden_size
        F = m.config.image_feat_size
        z = lambda shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=dev)
        self.pano_in = {"view_img_fts": z((B, self.V, F)), "loc_fts": z((B, self.V, 7)),
                        "nav_types": z((B, self.V), torch.int64), "view_lens": torch.full((B,), self.V, dtype=torch.int64, device=dev)}
        self.nav_in = {"txt_embeds": z((B, Lt, h)), "txt_lens": torch.ones(B, dtype=torch.int64, device=dev),
                       "gmap_img_embeds": z((B, G, h)), "gmap_step_ids": z((B, G), torch.int64),
                       "gmap_pos_fts": z((B, G, 7)), "gmap_pair_dists": z((B, G, G)),
                       "gmap_lens": torch.full((B,), 2, dtype=torch.int64, device=dev),
                       "vp_img_embeds": z((B, self.VP, h)), "vp_pos_fts": z((B, self.VP, 14)),
                       "vp_lens": torch.full((B,), self.VP, dtype=torch.int64, device=dev), "mem_slot": 1}
        self.nav_ix = {"g_valid": z((B, G), torch.uint8), "l_valid": z((B, self.VP), torch.uint8),
                       "node2cand": torch.full((B, G), -1, dtype=torch.int32, device=dev), "bw_mask": z((B, self.VP), torch.uint8)}
        self.graphs, self.outs = {}, {}
        self.stream = torch.cuda.Stream()

    def _capture(self, key, fn):
        s = self.stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):  # warm-up: lazy initialisation (function attributes, tensor maps) outside the capture
                fn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                out = fn()
        torch.cuda.current_stream().wait_stream(s)
        self.graphs[key], self.outs[key] = g, out

    def _replay(self, key):
        s = self.stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.graphs[key].replay()
        torch.cuda.current_stream().wait_stream(s)
        return self.outs[key]

    @staticmethod
    def _fill(dst, src):
        """dst[:src.shape] = src, rest zero (shapes differ only by padding)."""
        if tuple(dst.shape) == tuple(src.shape):
            dst.copy_(src, non_blocking=True)
            return
        dst.zero_()
        dst[tuple(slice(0, n) for n in src.shape)].copy_(src, non_blocking=True)

    def panorama(self, pano_inputs):
        """== model('panorama', pano_inputs) for B panoramas of 36 views."""
        for k in ("view_img_fts", "loc_fts", "nav_types", "view_lens"):
            self._fill(self.pano_in[k], pano_inputs[k])
        if "pano" not in self.graphs:
            self._capture("pano", lambda: self.model.forward_panorama(self.pano_in))
        return self._replay("pano")

    def navigation(self, nav_inputs, index=None):
        """== model('navigation', nav_inputs) with the graph padded to G nodes and the text to Lt tokens; the returned
        tensors are sliced back to the input's own sizes."""
        B, G0 = nav_inputs["gmap_step_ids"].shape
        Lt0 = nav_inputs["txt_embeds"].shape[1]
        if B != self.B or G0 > self.G or Lt0 > self.Lt or nav_inputs["vp_pos_fts"].shape[1] != self.VP:
            raise ValueError(f"navigation inputs (B={B}, G={G0}, Lt={Lt0}) exceed the stepper's capacities")
        ni = dict(nav_inputs)
        if ni.get("txt_lens") is None:
            ni["txt_lens"] = ni["txt_masks"].sum(1)
        if ni.get("gmap_lens") is None:
            ni["gmap_lens"] = ni["gmap_masks"].sum(1) + 1
        if ni.get("vp_lens") is None:
            ni["vp_lens"] = ni["vp_masks"].sum(1)
        for k in ("txt_embeds", "txt_lens", "gmap_img_embeds", "gmap_step_ids", "gmap_pos_fts", "gmap_pair_dists",
                  "gmap_lens", "vp_img_embeds", "vp_pos_fts", "vp_lens"):
            self._fill(self.nav_in[k], ni[k].to(self.nav_in[k].dtype))
        ix = index if index is not None else nav_index(nav_inputs, n_special=2)
        self.nav_ix["node2cand"].fill_(-1)
        for k in ("g_valid", "l_valid", "bw_mask"):
            self._fill(self.nav_ix[k], ix[k])
        self.nav_ix["node2cand"][:, :G0].copy_(ix["node2cand"], non_blocking=True)
        if "nav" not in self.graphs:
            self._capture("nav", lambda: self.model.forward_navigation(self.nav_in, index=self.nav_ix))
        o = self._replay("nav")
        out = {"gmap_embeds": o["gmap_embeds"][:, :G0], "vp_embeds": o["vp_embeds"], "cls_embeds": o["cls_embeds"],
               "global_logits": o["global_logits"][:, :G0], "fused_logits": o["fused_logits"][:, :G0],
               "local_logits": o["local_logits"]}
        if "gmap_attns" in o:  # [B, layers, G, G + Lt] -> the input's own graph / text sizes
            ga, va = o["gmap_attns"], o["vp_attns"]
            out["gmap_attns"] = torch.cat([ga[:, :, :G0, :G0], ga[:, :, :G0, self.G:self.G + Lt0]], -1)
            out["vp_attns"] = torch.cat([va[:, :, :, :self.VP], va[:, :, :, self.VP:self.VP + Lt0]], -1)
        return out
