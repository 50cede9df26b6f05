"""Host-side conversion of the batch's viewpoint-id STRING lists into integer index tables (once per
batch, on the CPU copy of the batch -- ideally inside the data loader, before the H2D move).

Spec: SURVEY.md A.4; semantics follow the reference data pipeline
(pretrain_src/data/dataset.py:513-549 node order / visited masks, :742-756 candidate-first view order)
and the DUET-lineage model loops the reference's missing `pretrain_goat.py` runs on strings
(gmap feature aggregation; global/local logit fusion).  Everything here is exact integer logic.
"""
import numpy as np
import torch

INDEX_KEY = "_magic_index"


def _cpu(t):
    return t.detach().cpu() if torch.is_tensor(t) else t


def build_index(batch):
    """-> dict of CPU tensors (int32 / uint8 / int64 / float32) + python ints."""
    steps = batch["traj_step_lens"]
    B = len(steps)
    # views per panorama (compact batches carry the view order instead of the features: featurizer.py)
    Vv = batch["traj_view_img_fts"].shape[1] if batch.get("traj_view_img_fts") is not None \
        else batch["traj_view_perm"].shape[1]
    # tokens per panorama: the views, followed by the panorama's object tokens when the batch carries objects
    # (og_collate, data/tasks.py:503-559; token order dataset.py:447)
    V = batch["traj_loc_fts"].shape[1]
    G = batch["gmap_step_ids"].shape[1]
    Vp = batch["vp_pos_fts"].shape[1]
    gmap_lens = _cpu(batch["gmap_lens"]).tolist()
    visited_masks = _cpu(batch["gmap_visited_masks"]).numpy().astype(bool)
    view_lens = _cpu(batch["traj_vp_view_lens"])
    obj_lens = _cpu(batch["traj_vp_obj_lens"]) if batch.get("traj_vp_obj_lens") is not None else None
    tok_lens = view_lens if obj_lens is None else view_lens + obj_lens
    nav_types = _cpu(batch["traj_nav_types"]).numpy()

    node_ptr, entries = [0], []
    last_rows = []
    row0 = 0
    for b in range(B):
        T = steps[b]
        visited_row, cand_tokens = {}, {}
        for t in range(T):
            visited_row[batch["traj_vpids"][b][t]] = row0 + t  # last occurrence wins (dict overwrite)
            for j, c in enumerate(batch["traj_cand_vpids"][b][t]):
                cand_tokens.setdefault(c, []).append((row0 + t) * V + j)
        vps = batch["gmap_vpids"][b]
        for n in range(G):
            if 1 <= n < len(vps):
                v = vps[n]
                if v in visited_row:
                    entries.append(-(visited_row[v] + 1))
                else:
                    entries.extend(cand_tokens[v])
            node_ptr.append(len(entries))
        row0 += T
        last_rows.append(row0 - 1)
    n_rows = row0
    # valid local tokens = views of the last step + [stop] (DUET lineage; the reference collate's 'vp_lens' key is
    # the constant 14 = len(x[-1]) of a [Vp,14] tensor, pretrain_src/data/tasks.py:153, and is not a length)
    vp_lens_t = tok_lens[torch.as_tensor(last_rows, dtype=torch.int64)] + 1
    vp_lens = vp_lens_t.tolist()

    # reverse CSR (unique source -> nodes, weights 1/count(node)) for the deterministic backward
    node_of_entry = np.repeat(np.arange(B * G), np.diff(np.asarray(node_ptr)))
    cnt = np.diff(np.asarray(node_ptr))
    ent = np.asarray(entries, dtype=np.int64)
    order = np.argsort(ent, kind="stable")
    ent_sorted = ent[order]
    src_ids, starts = np.unique(ent_sorted, return_index=True)
    src_ptr = np.append(starts, len(ent_sorted))
    src_nodes = node_of_entry[order]
    src_w = (1.0 / cnt[src_nodes]).astype(np.float32)

    # local branch: [stop] + the last step's views
    vp_gather = np.full((B, Vp), -1, dtype=np.int64)
    l_valid = np.zeros((B, Vp), dtype=np.uint8)
    for b in range(B):
        n = min(Vp - 1, V)
        vp_gather[b, 1:1 + n] = last_rows[b] * V + np.arange(n)
        l_valid[b, 0] = 1
        nt = nav_types[last_rows[b]]
        for j in range(1, min(Vp, vp_lens[b])):
            if j - 1 < V and nt[j - 1] == 1:
                l_valid[b, j] = 1

    # SAP masks + local->global scatter table
    g_valid = np.zeros((B, G), dtype=np.uint8)
    node2cand = np.full((B, G), -1, dtype=np.int32)
    bw_mask = np.zeros((B, Vp), dtype=np.uint8)
    for b in range(B):
        vps = batch["gmap_vpids"][b]
        vis = visited_masks[b]
        for n in range(min(G, gmap_lens[b])):
            if not vis[n]:
                g_valid[b, n] = 1
        visited_nodes = set(vp for vp, m in zip(vps, vis.tolist()) if m)
        tmp = {}
        for j, c in enumerate(batch["traj_cand_vpids"][b][-1]):
            if j + 1 >= Vp:
                break
            if c in visited_nodes:
                bw_mask[b, j + 1] = 1
            else:
                tmp[c] = j + 1  # last candidate with this id wins (dict overwrite)
        for n, vp in enumerate(vps):
            if n > 0 and vp not in visited_nodes and vp in tmp:
                node2cand[b, n] = tmp[vp]

    idx = dict(
        node_ptr=torch.from_numpy(np.asarray(node_ptr, dtype=np.int32)),
        entries=torch.from_numpy(np.asarray(entries, dtype=np.int32).reshape(-1)),
        src_ids=torch.from_numpy(src_ids.astype(np.int32)), src_ptr=torch.from_numpy(src_ptr.astype(np.int32)),
        src_nodes=torch.from_numpy(src_nodes.astype(np.int32)), src_w=torch.from_numpy(src_w),
        vp_gather=torch.from_numpy(vp_gather.reshape(-1)), l_valid=torch.from_numpy(l_valid),
        g_valid=torch.from_numpy(g_valid), node2cand=torch.from_numpy(node2cand), bw_mask=torch.from_numpy(bw_mask),
        last_rows=torch.from_numpy(np.asarray(last_rows, dtype=np.int64)),
        stop_rows_g=torch.arange(B, dtype=torch.int64) * G, stop_rows_v=torch.arange(B, dtype=torch.int64) * Vp,
        key_lens_txt=_cpu(batch["txt_lens"]).to(torch.int32), key_lens_gmap=_cpu(batch["gmap_lens"]).to(torch.int32),
        key_lens_vp=vp_lens_t.to(torch.int32), key_lens_pano=tok_lens.to(torch.int32),
        n_nodes=B * G, n_src=int(len(src_ids)),
    )
    if obj_lens is not None:
        # panorama token t of row r comes from view t (t < view_len) or object t - view_len; -1 = padding (zero row)
        R, O = n_rows, batch["traj_obj_img_fts"].shape[1]
        tt = np.arange(V)[None, :]
        vl, ol = view_lens.numpy()[:, None], obj_lens.numpy()[:, None]
        rr = np.arange(R)[:, None]
        idx["pano_view_src"] = torch.from_numpy(np.where(tt < vl, rr * Vv + tt, -1).astype(np.int64).reshape(-1))
        idx["pano_obj_src"] = torch.from_numpy(
            np.where((tt >= vl) & (tt < vl + ol), rr * O + (tt - vl), -1).astype(np.int64).reshape(-1))
        idx["pano_lens"] = tok_lens.to(torch.int64)
        # OG head: the object tokens of the LAST panorama inside the local sequence ([stop] + views + objects)
        last = np.asarray(last_rows)
        kk = np.arange(O)[None, :]
        lv, lo = view_lens.numpy()[last][:, None], obj_lens.numpy()[last][:, None]
        ok = kk < lo
        idx["og_rows"] = torch.from_numpy(
            np.where(ok, np.arange(B)[:, None] * Vp + 1 + lv + kk, -1).astype(np.int64).reshape(-1))
        idx["og_valid"] = torch.from_numpy(ok.astype(np.uint8))
    Lt = batch["txt_ids"].shape[1]
    idx["cls_rows_txt"] = torch.arange(B, dtype=torch.int64) * Lt
    idx["arange_b"] = torch.arange(B, dtype=torch.int64)
    if "vp_view_mrc_masks" in batch:
        # MRC: masked views of the last-step panorama, row-major (b, view) order = boolean-mask order of
        # `vp_view_probs[mask]` (train_r2r_magic.py:483; data/tasks.py:183-187)
        m = _cpu(batch["vp_view_mrc_masks"]).bool()
        V = Vv
        bb, jj = m.nonzero(as_tuple=True)
        last = torch.from_numpy(np.asarray(last_rows, dtype=np.int64))
        idx["mrc_rows"] = (bb * Vp + 1 + jj).to(torch.int64)          # rows of vp_embeds [B*Vp, h]
        idx["mrc_tgt_rows"] = (bb * m.shape[1] + jj).to(torch.int64)  # rows of vp_view_probs [B*V, C]
        idx["mrc_fts_rows"] = (last[bb] * V + jj).to(torch.int64)     # rows of traj_view_img_fts [R*V, F]
    if "txt_labels" in batch:
        lab = _cpu(batch["txt_labels"])
        sel = (lab != -1)
        pos = sel.reshape(-1).nonzero()[:, 0]          # row-major (b, position) order
        idx["mlm_rows"] = pos.to(torch.int64)
        idx["mlm_labels"] = lab[sel].to(torch.int64)
        idx["mlm_row_sample"] = (pos // lab.shape[1]).to(torch.int64)
        counts = sel.sum(1).clamp(min=1).to(torch.float32)
        idx["mlm_inv_count"] = (1.0 / counts)
    return idx


def _pad1(t, n, value=0):
    if t.numel() >= n:
        return t
    return torch.cat([t, torch.full((n - t.numel(),), value, dtype=t.dtype)])


def pad_batch(batch, n_panos=None, n_masked=None, n_entries=None, n_sources=None):
    """Pad a prepared CPU batch to fixed capacities so every batch of a task has the same tensor shapes
    (one CUDA graph per task).  Padded panoramas have a single all-zero view and are referenced by no index
    table; padded masked-token rows gather a zero row, carry label -1 (ignored) and are excluded from the
    supervised mean through `mlm_inv_n`.  Padding must not change the distillation objective either:
    `pano_row_scale` / `mlm_row_scale` are per-row KD weights (0 on padding, capacity / real count elsewhere)."""
    ix = batch[INDEX_KEY]
    if n_panos is not None:
        R = batch["traj_vp_view_lens"].shape[0]
        if n_panos < R:
            raise ValueError(f"pano capacity {n_panos} < {R}")
        if n_panos > R and batch.get("traj_obj_img_fts") is not None:
            raise NotImplementedError("pad_batch: object-grounding batches are not padded (og is not an R2R / RxR task)")
        if n_panos > R:
            def padr(t, value=0):
                pad = torch.full((n_panos - R, *t.shape[1:]), value, dtype=t.dtype)
                return torch.cat([t, pad], 0)
            for k in ("traj_view_img_fts", "traj_loc_fts", "traj_nav_types", "traj_vp_index"):
                if batch.get(k) is not None:
                    batch[k] = padr(batch[k])
            if batch.get("traj_view_perm") is not None:  # padded panoramas: every view masked (-1 -> zero row)
                batch["traj_view_perm"] = padr(batch["traj_view_perm"], -1)
            batch["traj_vp_view_lens"] = padr(batch["traj_vp_view_lens"], 1)
            ix["key_lens_pano"] = padr(ix["key_lens_pano"], 1)
        # KD row weights of the panorama tensors: 0 on padded panoramas, n_panos / R on real ones, so that a mean over
        # the padded tensor equals the mean over the real rows (makd.compute_kd_losses)
        ix["pano_row_scale"] = torch.cat([torch.full((R,), n_panos / max(R, 1), dtype=torch.float32),
                                          torch.zeros(n_panos - R, dtype=torch.float32)])
    if n_entries is not None:
        if ix["entries"].numel() > n_entries:
            raise ValueError("gmap entry capacity too small")
        ix["entries"] = _pad1(ix["entries"], n_entries)
        ix["src_nodes"] = _pad1(ix["src_nodes"], n_entries)
        ix["src_w"] = _pad1(ix["src_w"], n_entries)
    if n_sources is not None:
        ns = ix["src_ids"].numel()
        if ns > n_sources:
            raise ValueError("gmap source capacity too small")
        last = int(ix["src_ptr"][-1]) if ix["src_ptr"].numel() else 0
        ix["src_ids"] = _pad1(ix["src_ids"], n_sources)
        ix["src_ptr"] = _pad1(ix["src_ptr"], n_sources + 1, last)  # empty ranges -> skipped by the kernel
        ix["n_src"] = n_sources
    if "mlm_rows" in ix:
        n = ix["mlm_rows"].numel()
        ix["mlm_inv_n"] = torch.tensor([1.0 / max(n, 1)], dtype=torch.float32)
        if n_masked is not None:
            if n_masked < n:
                raise ValueError(f"masked-token capacity {n_masked} < {n}")
            if n_masked > n:
                pad = n_masked - n
                ix["mlm_rows"] = torch.cat([ix["mlm_rows"], torch.full((pad,), -1, dtype=torch.int64)])
                ix["mlm_labels"] = torch.cat([ix["mlm_labels"], torch.full((pad,), -1, dtype=torch.int64)])
                ix["mlm_row_sample"] = torch.cat([ix["mlm_row_sample"], torch.zeros(pad, dtype=torch.int64)])
            # KD row weights of the [n_masked, vocab] logits: padded rows 0, real rows n_masked / n
            ix["mlm_row_scale"] = torch.cat([torch.full((n,), n_masked / max(n, 1), dtype=torch.float32),
                                             torch.zeros(n_masked - n, dtype=torch.float32)])
    return batch


def prepare_batch(batch):
    """Attach the index tables to a (CPU) collate batch.  Call before moving the batch to the GPU."""
    if INDEX_KEY not in batch:
        batch[INDEX_KEY] = build_index(batch)
    return batch


def index_to(idx, device, non_blocking=False):
    return {k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v) for k, v in idx.items()}


def batch_to_device(batch, device, non_blocking=False):
    out = {}
    for k, v in batch.items():
        if k == INDEX_KEY:
            out[k] = index_to(v, device, non_blocking)
        elif torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=non_blocking)
        else:
            out[k] = v
    return out


# ---------------------------------------------------------------------------------------------------
# flat batches: every tensor of a batch (and of its index tables) as a view into ONE byte buffer, so a batch moves
# host -> device (or into a CUDA graph's static inputs) with a single copy instead of ~35 small ones
# ---------------------------------------------------------------------------------------------------
FLAT_KEY = "_magic_flat"
_ALIGN = 256


def _tensor_items(batch):
    for k in sorted(batch.keys()):
        v = batch[k]
        if k == FLAT_KEY:
            continue
        if k == INDEX_KEY:
            for kk in sorted(v.keys()):
                if torch.is_tensor(v[kk]):
                    yield (k, kk), v[kk]
        elif torch.is_tensor(v):
            yield (k, None), v


def flat_layout(batch):
    """-> ([(key, subkey, offset, shape, dtype)], total bytes); deterministic in the batch's shape signature."""
    lay, off = [], 0
    for (k, kk), t in _tensor_items(batch):
        lay.append((k, kk, off, tuple(t.shape), t.dtype))
        off += (t.numel() * t.element_size() + _ALIGN - 1) // _ALIGN * _ALIGN
    return lay, max(off, _ALIGN)


def _views(batch, flat, lay):
    out = {k: v for k, v in batch.items() if not torch.is_tensor(v) and k != INDEX_KEY}
    if INDEX_KEY in batch:
        out[INDEX_KEY] = {kk: vv for kk, vv in batch[INDEX_KEY].items() if not torch.is_tensor(vv)}
    for k, kk, off, shape, dtype in lay:
        n = 1
        for s in shape:
            n *= s
        nb = n * torch.empty(0, dtype=dtype).element_size()
        t = flat[off:off + nb].view(dtype).view(shape)
        if kk is None:
            out[k] = t
        else:
            out[k][kk] = t
    out[FLAT_KEY] = flat
    return out


def flatten_batch(batch, device=None, pin=False):
    """Repack a batch so that all its tensors are views into one uint8 buffer (`batch[FLAT_KEY]`), on the host
    (optionally pinned) or on `device`.  A batch that is already flat moves with one copy."""
    lay, total = flat_layout(batch)
    src_flat = batch.get(FLAT_KEY)
    if src_flat is None or src_flat.numel() != total:
        src_flat = torch.empty(total, dtype=torch.uint8)
        host = _views(batch, src_flat, lay)
        for (k, kk), t in _tensor_items(batch):
            dst = host[k] if kk is None else host[k][kk]
            dst.copy_(t)
    else:
        host = batch
    if device is None:
        if pin and not src_flat.is_pinned():
            return _views(host, src_flat.pin_memory(), lay)
        return host
    return _views(host, src_flat.to(device, non_blocking=True), lay)


def copy_batch_(dst, src):
    """dst <- src for two batches of the same signature; one copy when both are flat."""
    fd, fs = dst.get(FLAT_KEY), src.get(FLAT_KEY)
    if fd is not None and fs is not None and fd.numel() == fs.numel():
        fd.copy_(fs, non_blocking=True)
        return 1
    n = 0
    for (k, kk), t in _tensor_items(src):
        (dst[k] if kk is None else dst[k][kk]).copy_(t, non_blocking=True)
        n += 1
    return n


def alloc_like(batch, device):
    """An uninitialised flat device batch with the layout of `batch` (python-side members shared by reference)."""
    lay, total = flat_layout(batch)
    return _views(batch, torch.empty(total, dtype=torch.uint8, device=device), lay)
