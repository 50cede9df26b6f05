"""GPU batch featuriser, feature half (featurizer.py / csrc/featurize.cu): the gathers are exact copies (index work:
bit-exact), and a model fed a COMPACT batch (panorama rows + view orders, features gathered from the device store)
produces exactly what it produces on the materialised batch the reference collate would have built."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import magic_b200  # noqa: E402
from magic_b200 import synth  # noqa: E402
from magic_b200.config import make_config  # noqa: E402
from magic_b200.featurizer import FeatureStore, compact_batch  # noqa: E402
from magic_b200.graph_index import batch_to_device, pad_batch, prepare_batch  # noqa: E402

DEV = "cuda"


@pytest.mark.parametrize("sdt,odt", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32),
                                     (torch.bfloat16, torch.float32), (torch.float32, torch.bfloat16)])
def test_gather_views_is_an_exact_copy(sdt, odt):
    g = torch.Generator().manual_seed(3)
    fts = torch.randn(37, 36, 768, generator=g).to(sdt)
    store = FeatureStore(fts, DEV, dtype=sdt)
    R = 29
    vp = torch.randint(0, 37, (R,), generator=g)
    perm = torch.stack([torch.randperm(36, generator=g) for _ in range(R)]).int()
    perm[3, 30:] = -1      # padded views
    perm[7] = -1           # a padded panorama
    out = store.gather_views(vp.to(DEV), perm.to(DEV), out_dtype=odt)
    ref = fts[vp[:, None], perm.long().clamp(min=0)].to(odt)
    ref[perm < 0] = 0
    assert out.dtype == odt and torch.equal(out.cpu(), ref)


def test_gather_pair_dists():
    g = torch.Generator().manual_seed(4)
    N = 50
    d = torch.rand(N, N, generator=g) * 30
    d = (d + d.T) / 2
    d.fill_diagonal_(0)
    store = FeatureStore(torch.zeros(N, 36, 8), DEV, pair_dists=d)
    node_vp = torch.randint(0, N, (6, 20), generator=g)
    node_vp[:, 0] = -1      # [stop]
    node_vp[2, 15:] = -1    # padding
    out = store.gather_pair_dists(node_vp.to(DEV)).cpu()
    ok = (node_vp >= 0)
    ref = d[node_vp.clamp(min=0)[:, :, None], node_vp.clamp(min=0)[:, None, :]] * (ok[:, :, None] & ok[:, None, :])
    assert torch.equal(out, ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("task", ["sap", "mlm", "mrc"])
def test_model_on_compact_batch_equals_materialised_batch(task, dtype):
    cfg = make_config(128, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, pretrain_tasks=("mlm", "sap", "mrc"))
    torch.manual_seed(0)
    model = magic_b200.GlocalTextPathCMTPreTraining(cfg).to(DEV).train().set_compute_dtype(dtype)
    cpu_store = synth.make_store(96, seed=5, dtype=dtype)
    FeatureStore(cpu_store, DEV, dtype=dtype).attach(model)
    full = synth.make_batch(task, 6, seed=31, store=cpu_store)
    comp = compact_batch(synth.make_batch(task, 6, seed=31, store=cpu_store))
    assert "traj_view_img_fts" not in comp
    K = magic_b200.INDEX_KEY
    outs = []
    for b in (full, comp):
        prepare_batch(b)
        R = b["traj_vp_view_lens"].shape[0]
        b = pad_batch(b, R + 3, b[K]["mlm_rows"].numel() + 5 if task == "mlm" else None, b[K]["entries"].numel() + 7,
                      b[K]["src_ids"].numel() + 7)
        model.zero_grad(set_to_none=True)
        o = model(batch_to_device(b, DEV), task, True)
        o["loss"].float().sum().backward()
        outs.append((o["loss"].detach().float().clone(), o["pano_embeds"].detach().float().clone(),
                     model.bert.img_embeddings.img_linear.weight.grad.detach().clone()))
    nbytes = lambda b: sum(v.numel() * v.element_size() for v in b.values() if torch.is_tensor(v))
    if task != "mrc":  # (MRC also ships its [B, 36, 1000] soft labels)
        assert nbytes(comp) * 10 < nbytes(full)  # what travels host -> device shrinks by more than 10x
    (l0, p0, g0), (l1, p1, g1) = outs
    assert torch.equal(l0, l1) and torch.equal(p0, p1)
    assert torch.allclose(g0, g1, rtol=1e-5, atol=1e-7)  # (split-K / atomics order in the weight gradient)
