"""Host logic, exact integer work: the index tables built from the viewpoint-id strings must reproduce the
oracle's python/string loops bit-exactly.  The device kernels are emulated here with plain
indexing so the check runs without a GPU; the kernels themselves are checked in the -m gpu tests."""
import pytest
import torch

import magic_b200  # noqa: F401
from magic_b200 import synth
from magic_b200.graph_index import build_index, pad_batch, prepare_batch, INDEX_KEY
from oracle import magic_oracle as O


def emulate_gmap_aggregate(tokens, fused, ix, h):
    out = torch.zeros(ix["n_nodes"], h)
    ptr, ent = ix["node_ptr"].tolist(), ix["entries"].tolist()
    flat = tokens.reshape(-1, h)
    for n in range(ix["n_nodes"]):
        e = ent[ptr[n]:ptr[n + 1]]
        if not e:
            continue
        rows = [flat[s] if s >= 0 else fused[-(s + 1)] for s in e]
        out[n] = rows[0] if len(rows) == 1 else torch.stack(rows, 0).sum(0) / len(rows)
    return out


def emulate_sap_fuse(gl_raw, ll_raw, ix):
    g_valid, l_valid = ix["g_valid"].bool(), ix["l_valid"].bool()
    gl = torch.where(g_valid, gl_raw, torch.full_like(gl_raw, float("-inf")))
    ll = torch.where(l_valid, ll_raw, torch.full_like(ll_raw, float("-inf")))
    fl = gl.clone()
    B, G = gl.shape
    for b in range(B):
        bw = torch.zeros(())
        for j in range(1, ll.shape[1]):
            if ix["bw_mask"][b, j]:
                bw = bw + ll[b, j]
        fl[b, 0] = fl[b, 0] + ll[b, 0]
        for n in range(1, G):
            if g_valid[b, n]:
                c = int(ix["node2cand"][b, n])
                fl[b, n] = fl[b, n] + (ll[b, c] if c >= 0 else bw)
    return gl, ll, fl


@pytest.mark.parametrize("seed,L,T,G", [(1, 80, 5, 20), (2, 80, 5, 20), (3, 160, 12, 50)])
def test_index_tables_reproduce_string_loops(seed, L, T, G):
    b = synth.make_batch("sap", 8, L=L, T_max=T, G_max=G, seed=seed)
    ix = build_index(b)
    h = 16
    g = torch.Generator().manual_seed(seed)
    R = sum(b["traj_step_lens"])
    pano, fused = torch.randn(R, 36, h, generator=g), torch.randn(R, h, generator=g)
    ref = O.aggregate_gmap_features(pano, fused, b)
    got = emulate_gmap_aggregate(pano, fused, ix, h).view(ref.shape)
    assert torch.equal(got, ref)
    # reverse CSR is the exact adjoint of the forward table
    w = torch.zeros(ix["n_nodes"], R * 36 + R)
    ptr, ent = ix["node_ptr"].tolist(), ix["entries"].tolist()
    for n in range(ix["n_nodes"]):
        e = ent[ptr[n]:ptr[n + 1]]
        for s in e:
            w[n, s if s >= 0 else R * 36 + (-(s + 1))] += 1.0 / len(e)
    w2 = torch.zeros_like(w)
    sp = ix["src_ptr"].tolist()
    for i, s in enumerate(ix["src_ids"].tolist()):
        for e in range(sp[i], sp[i + 1]):
            w2[int(ix["src_nodes"][e]), s if s >= 0 else R * 36 + (-(s + 1))] += float(ix["src_w"][e])
    assert torch.allclose(w, w2)
    # SAP masks and local->global fusion
    B, Gp = b["gmap_step_ids"].shape
    gl_raw, ll_raw = torch.randn(B, Gp, generator=g), torch.randn(B, 37, generator=g)
    gmask = O.gen_seq_masks(b["gmap_lens"], Gp)
    gl_ref = gl_raw.masked_fill(b["gmap_visited_masks"], float("-inf")).masked_fill(~gmask, float("-inf"))
    rows = O.last_step_rows(b)
    nav = torch.cat([torch.ones(B, 1, dtype=torch.bool), b["traj_nav_types"][rows] == 1], 1) & O.gen_seq_masks(O.vp_lens_of(b), 37)
    ll_ref = ll_raw.masked_fill(~nav, float("-inf"))
    fl_ref = O.fuse_sap_logits(gl_ref, ll_ref, b)
    gl, ll, fl = emulate_sap_fuse(gl_raw, ll_raw, ix)
    assert torch.equal(gl, gl_ref) and torch.equal(ll, ll_ref) and torch.equal(fl, fl_ref)
    assert torch.equal(fl.argmax(1), fl_ref.argmax(1))
    # local-branch gather: [stop] + last-step views
    vp = ix["vp_gather"].view(B, 37)
    assert (vp[:, 0] == -1).all()
    for i in range(B):
        assert vp[i, 1:].tolist() == [rows[i] * 36 + j for j in range(36)]


def test_mlm_row_order_and_padding():
    b = prepare_batch(synth.make_batch("mlm", 8, seed=4))
    ix = b[INDEX_KEY]
    lab = b["txt_labels"]
    assert torch.equal(ix["mlm_labels"], lab[lab != -1])                 # train_r2r_magic.py:450-452 order
    assert torch.equal(lab.reshape(-1)[ix["mlm_rows"]], ix["mlm_labels"])
    n = ix["mlm_rows"].numel()
    R = b["traj_view_img_fts"].shape[0]
    pad_batch(b, n_panos=R + 5, n_masked=n + 7, n_entries=ix["entries"].numel() + 10, n_sources=ix["src_ids"].numel() + 3)
    assert b["traj_view_img_fts"].shape[0] == R + 5 and int(b["traj_vp_view_lens"][-1]) == 1
    assert ix["mlm_rows"].numel() == n + 7 and int(ix["mlm_rows"][-1]) == -1 and int(ix["mlm_labels"][-1]) == -1
    assert abs(float(ix["mlm_inv_n"]) - 1.0 / n) < 1e-9
    sp = ix["src_ptr"]
    assert sp.numel() == ix["src_ids"].numel() + 1 and int(sp[-1]) == int(sp[-4])    # padded sources are empty


def _collate_with_index(samples):
    return prepare_batch(synth.collate(samples))


def test_index_tables_are_built_in_loader_workers():
    """the collate function attaches the index tables in the DataLoader worker, so the training
    process never walks the viewpoint-id strings.  A torch DataLoader with worker processes returns batches whose
    tables equal the ones built in the main process."""
    import torch.utils.data as tud
    samples = synth.make_samples("sap", 12, seed=21)
    want = [_collate_with_index([dict(s) for s in samples[i:i + 4]]) for i in range(0, 12, 4)]
    try:
        got = list(tud.DataLoader([dict(s) for s in samples], batch_size=4, shuffle=False, num_workers=2,
                                  collate_fn=_collate_with_index, timeout=60))
    except RuntimeError as e:  # a worker that failed to start on a loaded box is not what this test is about
        pytest.skip(f"DataLoader workers unavailable: {e}")
    assert len(got) == 3
    K = magic_b200.INDEX_KEY
    for a, b in zip(want, got):
        assert set(a[K]) == set(b[K])
        for k, v in a[K].items():
            assert torch.equal(v, b[K][k]) if torch.is_tensor(v) else v == b[K][k], k
        assert torch.equal(a["gmap_pair_dists"], b["gmap_pair_dists"]) and a["gmap_vpids"] == b["gmap_vpids"]
