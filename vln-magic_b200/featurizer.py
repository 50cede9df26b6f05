"""GPU batch featuriser, feature half.

The reference reads every panorama's 36 x 768 CLIP features from a host-side cache per sample, collates them
(pretrain_src/data/dataset.py:210-244, 742-756; data/tasks.py:121-133) and copies ~28 MB of fp32 per step to the GPU.
Here the features of all panoramas live in HBM once (R2R: 10 567 panoramas x 36 x 768 bf16 = 0.58 GB of the 180 GB);
a batch carries, per trajectory step, the panorama's row in the store and the order of its views (`traj_vp_index`
[sum T] int64, `traj_view_perm` [sum T, 36] int32; -1 = padded view), and `traj_view_img_fts` is gathered on the
device in the model's compute dtype -- no fp32 staging, no cast kernel, 50x fewer host->device bytes.  The all-pairs
shortest-distance matrix (dataset.py:545-549 reads it per sample) can be resident too: `gather_pair_dists`.

The view ORDER stays the reference's: candidate views first (in candidate order), then the remaining views in
ascending view index (dataset.py:742-756): it is computed on the host with the rest of the (tiny) geometry."""
import torch

from ._lib import call, dt, ptr, stream

VIEW_KEYS = ("traj_vp_index", "traj_view_perm")


class FeatureStore:
    def __init__(self, view_fts, device="cuda", dtype=torch.bfloat16, pair_dists=None):
        """view_fts: [N, V, D] (any float dtype, CPU or GPU); pair_dists: optional [N, N] fp32."""
        if view_fts.dim() != 3 or view_fts.shape[2] % 8 != 0:
            raise ValueError("view_fts must be [N, V, D] with D a multiple of 8")
        self.view_fts = view_fts.to(device=device, dtype=dtype).contiguous()
        self.N, self.V, self.D = self.view_fts.shape
        self.pair_dists = pair_dists.to(device=device, dtype=torch.float32).contiguous() if pair_dists is not None \
            else None

    @property
    def nbytes(self):
        return self.view_fts.numel() * self.view_fts.element_size()

    def gather_views(self, vp_index, view_perm, out_dtype=None, out=None):
        """-> [R, V, D] in `out_dtype` (default: the store's): row r = panorama vp_index[r] with its views in the order
        view_perm[r] (-1 -> a zero row)."""
        R = vp_index.shape[0]
        out_dtype = out_dtype or self.view_fts.dtype
        if out is None:
            out = torch.empty(R, self.V, self.D, dtype=out_dtype, device=self.view_fts.device)
        vp = vp_index.contiguous()
        pm = view_perm.contiguous()
        if vp.dtype != torch.int64 or pm.dtype != torch.int32 or tuple(pm.shape) != (R, self.V):
            raise ValueError("vp_index must be int64 [R] and view_perm int32 [R, V]")
        call("magic_gather_views", ptr(self.view_fts), dt(self.view_fts), self.N, ptr(vp), ptr(pm), ptr(out), dt(out), R,
             self.V, self.D, stream())
        return out

    def gather_pair_dists(self, node_vp):
        """node_vp [B, G] int64 store rows of the graph nodes (-1 for [stop] and padding) -> gmap_pair_dists [B, G, G]."""
        if self.pair_dists is None:
            raise ValueError("this store holds no distance matrix")
        B, G = node_vp.shape
        out = torch.empty(B, G, G, dtype=torch.float32, device=node_vp.device)
        call("magic_gather_pair_dists", ptr(self.pair_dists), self.N, ptr(node_vp.contiguous()), ptr(out), B, G, stream())
        return out

    def attach(self, *models):
        """Models read `traj_view_img_fts` from this store whenever a batch carries the compact view keys instead."""
        for m in models:
            if m is not None:
                m.feature_store = self
                if hasattr(m, "bert"):
                    m.bert.feature_store = self
        return self


def compact_batch(batch):
    """Drop the materialised features of a batch that also carries the compact view keys (what travels host->device)."""
    if not all(k in batch for k in VIEW_KEYS):
        raise ValueError("batch has no traj_vp_index / traj_view_perm")
    out = {k: v for k, v in batch.items() if k != "traj_view_img_fts"}
    return out


# ---------------------------------------------------------------------------------------------------
# graph half: the batch tensors that derive from the navigation graph, built on the device
# ---------------------------------------------------------------------------------------------------
import ctypes  # noqa: E402

import numpy as np  # noqa: E402

from . import _lib as L  # noqa: E402
from .graph_index import INDEX_KEY  # noqa: E402


class MagicFeatArgs(ctypes.Structure):
    """Mirror of `MagicFeatArgs` in include/magic_b200.h (field order matters)."""
    _PTRS = ("pos dist hops cand_vp cand_view cand_ang n_cand view_ang path path_len start_heading next_vp prev_vp row0 "
             "traj_vp_index traj_view_perm traj_loc_fts traj_nav_types traj_vp_view_lens gmap_node_vp gmap_step_ids "
             "gmap_visited_masks gmap_lens gmap_pos_fts gmap_pair_dists vp_pos_fts global_act_labels local_act_labels "
             "node_ptr entries src_ids src_ptr src_nodes src_w n_src g_valid l_valid node2cand bw_mask vp_gather "
             "key_lens_gmap key_lens_vp last_rows slab_entries slab_nodes slab_rank slab_ptr slab_total slab_nvis "
             "status").split()
    _INTS = "N C B Tmax G Vp R R_cap E_s E_cap S_cap correct_heading".split()
    _fields_ = [(n, ctypes.c_void_p) for n in _PTRS] + [(n, ctypes.c_int) for n in _INTS]


class GraphWorld:
    """The navigation world, resident on the device: positions, all-pairs shortest distances / hop counts and the
    candidate table of every viewpoint (in the reference's `scanvp_cands` dict order).

    `cands[i]` = ordered list of (next viewpoint row, view index, heading offset, elevation offset)."""

    def __init__(self, pos, dist, hops, cands, view_ang, device="cuda"):
        N = len(cands)
        C = max(1, max(len(c) for c in cands))
        cv = np.full((N, C), -1, dtype=np.int32)
        cw = np.zeros((N, C), dtype=np.int32)
        ca = np.zeros((N, C, 2), dtype=np.float32)
        nc = np.zeros(N, dtype=np.int32)
        for i, lst in enumerate(cands):
            nc[i] = len(lst)
            for j, (nx, view, dh, de) in enumerate(lst):
                cv[i, j], cw[i, j], ca[i, j] = nx, view, (dh, de)
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(device)
        self.N, self.C, self.device = N, C, torch.device(device)
        self.pos = t(np.asarray(pos, dtype=np.float64), torch.float64)
        self.dist = t(np.asarray(dist, dtype=np.float32), torch.float32)
        self.hops = t(np.asarray(hops, dtype=np.int32), torch.int32)
        self.cand_vp, self.cand_view, self.cand_ang, self.n_cand = t(cv, torch.int32), t(cw, torch.int32), \
            t(ca, torch.float32), t(nc, torch.int32)
        self.view_ang = t(np.asarray(view_ang, dtype=np.float32), torch.float32)
        self.max_cands = int(nc.max())

    @classmethod
    def from_tables(cls, positions, distances, path_lens, scanvp_cands, view_ang, device="cuda"):
        """positions[scan][vp] = xyz; distances[scan][a][b]; path_lens[scan][a][b] = len(shortest path a -> b);
        scanvp_cands['scan_vp'][next_vp] = [view, ?, heading, elevation] (dict order = candidate order).
        -> (world, {'scan_vp': row})."""
        rows = {}
        for scan in sorted(positions):
            for vp in positions[scan]:
                rows[f"{scan}_{vp}"] = len(rows)
        N = len(rows)
        pos = np.zeros((N, 3), dtype=np.float64)
        dist = np.zeros((N, N), dtype=np.float32)
        hops = np.zeros((N, N), dtype=np.int32)
        cands = [[] for _ in range(N)]
        for scan in positions:
            for vp in positions[scan]:
                i = rows[f"{scan}_{vp}"]
                pos[i] = positions[scan][vp]
                for other, d in distances[scan][vp].items():
                    j = rows[f"{scan}_{other}"]
                    dist[i, j] = d
                    hops[i, j] = path_lens[scan][vp][other] - 1
                for nx, v in scanvp_cands.get(f"{scan}_{vp}", {}).items():
                    cands[i].append((rows[f"{scan}_{nx}"], int(v[0]), float(v[2]), float(v[3])))
        return cls(pos, dist, hops, cands, view_ang, device), rows

    @classmethod
    def from_reference_tables(cls, graphs, shortest_distances, shortest_paths, scanvp_cands, view_ang, device="cuda"):
        """From the reference loader's own structures (data/common.py:111-139 `load_nav_graphs`, dataset.py:160-175):
        networkx graphs with node 'position', `shortest_distances[scan][a][b]`, `shortest_paths[scan][a][b]`."""
        positions = {s: {v: G.nodes[v]["position"] for v in G.nodes} for s, G in graphs.items()}
        lens = {s: {a: {b: len(p) for b, p in d.items()} for a, d in shortest_paths[s].items()} for s in shortest_paths}
        return cls.from_tables(positions, shortest_distances, lens, scanvp_cands, view_ang, device)


class GraphFeaturizer:
    """One call per batch: paths (viewpoint rows) in, the graph-derived batch tensors and the model's index tables out,
    all in device memory at fixed capacities (one CUDA graph per task keeps replaying).  Output buffers are allocated
    once and reused by every call."""

    def __init__(self, world, B, Tmax, G, R_cap=None, E_cap=None, S_cap=None, Vp=37, correct_heading=False):
        self.world, self.B, self.Tmax, self.G, self.Vp = world, B, Tmax, G, Vp
        dev = world.device
        self.R_cap = R_cap or B * Tmax
        self.E_s = Tmax * (world.max_cands + 1)
        self.E_cap = E_cap or B * self.E_s
        self.S_cap = S_cap or self.E_cap
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        R = self.R_cap
        self.out = dict(
            traj_vp_index=z(R, torch.int64), traj_view_perm=z((R, 36), torch.int32), traj_loc_fts=z((R, 36, 7), torch.float32),
            traj_nav_types=z((R, 36), torch.int64), traj_vp_view_lens=z(R, torch.int64),
            gmap_node_vp=z((B, G), torch.int64), gmap_step_ids=z((B, G), torch.int64),
            gmap_visited_masks=z((B, G), torch.uint8), gmap_lens=z(B, torch.int64), gmap_pos_fts=z((B, G, 7), torch.float32),
            gmap_pair_dists=z((B, G, G), torch.float32), vp_pos_fts=z((B, Vp, 14), torch.float32),
            global_act_labels=z(B, torch.int64), local_act_labels=z(B, torch.int64))
        self.idx = dict(
            node_ptr=z(B * G + 1, torch.int32), entries=z(self.E_cap, torch.int32), src_ids=z(self.S_cap, torch.int32),
            src_ptr=z(self.S_cap + 1, torch.int32), src_nodes=z(self.E_cap, torch.int32), src_w=z(self.E_cap, torch.float32),
            g_valid=z((B, G), torch.uint8), l_valid=z((B, Vp), torch.uint8), node2cand=z((B, G), torch.int32),
            bw_mask=z((B, Vp), torch.uint8), vp_gather=z(B * Vp, torch.int64), key_lens_gmap=z(B, torch.int32),
            key_lens_vp=z(B, torch.int32), last_rows=z(B, torch.int64))
        self.n_src_dev = z(1, torch.int32)
        self.slab = dict(slab_entries=z(B * self.E_s, torch.int32), slab_nodes=z(B * self.E_s, torch.int32),
                         slab_rank=z(B * self.E_s, torch.int32), slab_ptr=z(B * (G + 1), torch.int32),
                         slab_total=z(B, torch.int32), slab_nvis=z(B, torch.int32))
        self.status = z(1, torch.int32)
        self.inp = dict(path=z((B, Tmax), torch.int32), path_len=z(B, torch.int32), start_heading=z(B, torch.float32),
                        next_vp=z(B, torch.int32), prev_vp=z(B, torch.int32), row0=z(B, torch.int32))
        from .ops import PinnedRing
        self.ring = PinnedRing(B * (Tmax + 5), dtype=torch.int32, n=4)
        self.wire = z((B, Tmax + 5), torch.int32)
        self.correct_heading = int(bool(correct_heading))
        self.stop_rows_g = torch.arange(B, dtype=torch.int64, device=dev) * G
        self.stop_rows_v = torch.arange(B, dtype=torch.int64, device=dev) * Vp

    def __call__(self, paths, start_headings, next_vps=None, prev_vps=None):
        """paths: B lists of viewpoint rows (len <= Tmax); next_vps: ground-truth next row per sample, -1 = stop
        (None: no labels); prev_vps: the viewpoint each agent came from, when it is not paths[b][-2] (the reference loader
        cuts a path longer than TRAIN_MAX_STEP to its first 20 viewpoints + the end one AFTER reading the heading off the
        true predecessor, dataset.py:658-665); None / -1: paths[b][-2].
        -> batch dict (device tensors, views of the featuriser's buffers)."""
        B, Tmax, w = self.B, self.Tmax, self.world
        if len(paths) != B:
            raise ValueError(f"expected {B} paths")
        lens = [len(p) for p in paths]
        if max(lens) > Tmax or min(lens) < 1:
            raise ValueError("path length outside [1, Tmax]")
        R = sum(lens)
        if R > self.R_cap:
            raise ValueError(f"panorama capacity {self.R_cap} < {R}")
        # the ~B*(Tmax+4) integers a batch costs on the wire: one pinned staging buffer, one H2D copy
        h = torch.zeros(B, Tmax + 5, dtype=torch.int32)
        row0 = 0
        for b, p in enumerate(paths):
            h[b, :lens[b]] = torch.as_tensor(p, dtype=torch.int32)
            h[b, Tmax] = lens[b]
            h[b, Tmax + 1] = row0
            h[b, Tmax + 2] = (-2 if next_vps is None else int(next_vps[b]))
            h[b, Tmax + 4] = (-1 if prev_vps is None or prev_vps[b] is None else int(prev_vps[b]))
            row0 += lens[b]
        h[:, Tmax + 3] = torch.as_tensor(np.asarray(start_headings, dtype=np.float32)).view(torch.int32)
        d = self.ring.upload(h, self.wire.view(-1)).view(B, Tmax + 5)
        self.inp["path"].copy_(d[:, :Tmax])
        self.inp["path_len"].copy_(d[:, Tmax])
        self.inp["row0"].copy_(d[:, Tmax + 1])
        self.inp["next_vp"].copy_(d[:, Tmax + 2])
        self.inp["prev_vp"].copy_(d[:, Tmax + 4])
        self.inp["start_heading"].copy_(d[:, Tmax + 3].contiguous().view(torch.float32))
        a = MagicFeatArgs()
        for k in ("pos", "dist", "hops", "cand_vp", "cand_view", "cand_ang", "n_cand", "view_ang"):
            setattr(a, k, getattr(w, k).data_ptr())
        for grp in (self.inp, self.out, self.idx, self.slab):
            for k, t in grp.items():
                setattr(a, k, t.data_ptr())
        if next_vps is None:
            a.next_vp = None
        a.n_src, a.status = self.n_src_dev.data_ptr(), self.status.data_ptr()
        a.N, a.C, a.B, a.Tmax, a.G, a.Vp, a.R, a.R_cap = w.N, w.C, B, Tmax, self.G, self.Vp, R, self.R_cap
        a.E_s, a.E_cap, a.S_cap, a.correct_heading = self.E_s, self.E_cap, self.S_cap, self.correct_heading
        L.call("magic_featurize_graph", ctypes.byref(a), L.stream())
        batch = dict(self.out)
        batch["gmap_visited_masks"] = self.out["gmap_visited_masks"].view(torch.bool)
        batch["traj_step_lens"] = lens
        ix = dict(self.idx)
        ix.update(stop_rows_g=self.stop_rows_g, stop_rows_v=self.stop_rows_v,
                  key_lens_pano=self.out["traj_vp_view_lens"].to(torch.int32), n_nodes=B * self.G, n_src=self.S_cap)
        if self.R_cap > R:  # KD row weights of padded panoramas (graph_index.pad_batch)
            ix["pano_row_scale"] = torch.cat([torch.full((R,), self.R_cap / R, dtype=torch.float32, device=w.device),
                                              torch.zeros(self.R_cap - R, dtype=torch.float32, device=w.device)])
        batch[INDEX_KEY] = ix
        return batch

    def check(self):
        """Raises if the last calls overflowed a capacity (reads one int back: call it off the step's critical path)."""
        code = int(self.status.item())
        if code:
            self.status.zero_()
            raise L.MagicError({1: "more than 256 distinct viewpoints in a sample", 2: "graph larger than G",
                                3: "two candidates share a view (more than 36 tokens in a panorama)",
                                4: "per-sample entry capacity E_s too small", 5: "entry capacity E_cap / S_cap too small"}
                               .get(code, f"featuriser error {code}"))


def attach_text(batch, txt_ids, txt_lens, txt_labels=None, device=None):
    """Text half of a batch (host-side: tokenisation and MLM masking stay in the loader): ids / lengths (+ labels) and
    the text-dependent index tables of graph_index.build_index."""
    device = device or batch["gmap_lens"].device
    B, Lt = txt_ids.shape
    batch["txt_ids"], batch["txt_lens"] = txt_ids.to(device), txt_lens.to(device)
    ix = batch[INDEX_KEY]
    ix["key_lens_txt"] = txt_lens.to(torch.int32).to(device)
    ix["cls_rows_txt"] = (torch.arange(B, dtype=torch.int64) * Lt).to(device)
    ix["arange_b"] = torch.arange(B, dtype=torch.int64, device=device)
    if txt_labels is not None:
        batch["txt_labels"] = txt_labels.to(device)
        sel = txt_labels != -1
        pos = sel.reshape(-1).nonzero()[:, 0]
        ix["mlm_rows"] = pos.to(torch.int64).to(device)
        ix["mlm_labels"] = txt_labels[sel].to(torch.int64).to(device)
        ix["mlm_row_sample"] = (pos // Lt).to(torch.int64).to(device)
        ix["mlm_inv_count"] = (1.0 / sel.sum(1).clamp(min=1).to(torch.float32)).to(device)
    return batch
