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

from ._lib import BF16, F32, call, dt, ptr, stream

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
