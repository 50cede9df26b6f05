"""Host logic: the synthetic batch generator's collate must reproduce the reference collate functions
(pretrain_src/data/tasks.py:110-166 mlm_collate, :263-324 mrc_collate, :392-451 sap_collate, :503-559 og_collate,
:618-677 cfp_collate) key by key, bit-exactly."""
import os
import sys
import types

import pytest
import torch

import magic_b200  # noqa: F401
from magic_b200 import synth

REF = "/root/reference/pretrain_src"


def _ref_tasks():
    """Import the reference data/tasks.py (+ data/common.py) without its absent third-party deps."""
    import importlib.util
    for stub in ("networkx", "pynvml"):
        if stub not in sys.modules:
            try:
                __import__(stub)
            except Exception:
                sys.modules[stub] = types.ModuleType(stub)
    pkg = types.ModuleType("refdata")
    pkg.__path__ = [os.path.join(REF, "data")]
    sys.modules["refdata"] = pkg
    for name in ("common", "tasks"):
        spec = importlib.util.spec_from_file_location(f"refdata.{name}", os.path.join(REF, "data", f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refdata.{name}"] = m
        spec.loader.exec_module(m)
    return sys.modules["refdata.tasks"]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference not mounted")
@pytest.mark.parametrize("task", ["mlm", "sap", "mrc", "cfp", "og"])
def test_collate_matches_reference(task):
    tasks = _ref_tasks()
    samples = synth.make_samples(task, 6, seed=5, obj_dim=32 if task == "og" else 0)
    ours = synth.collate([dict(s) for s in samples])
    ref_samples = [dict(s) for s in samples]
    if task == "cfp":  # CfpDataset adds a per-sample config echo (data/tasks.py:606); not a model input
        for s_ in ref_samples:
            s_["extra_heads"] = False
    ref = getattr(tasks, task + "_collate")(ref_samples)
    ref.pop("extra_heads", None)
    assert set(ours.keys()) == set(ref.keys()), set(ours.keys()) ^ set(ref.keys())
    for k, v in ref.items():
        if torch.is_tensor(v):
            assert v.dtype == ours[k].dtype and v.shape == ours[k].shape, k
            assert torch.equal(v, ours[k]), k
        elif k == "vp_angles":
            assert all((a == b).all() for a, b in zip(v, ours[k]))
        else:
            assert v == ours[k], k


@pytest.mark.parametrize("task,L,T,G", [("mlm", 80, 5, 20), ("sap", 80, 5, 20), ("sap", 160, 12, 50)])
def test_schema_and_invariants(task, L, T, G):
    b = synth.make_batch(task, 8, L=L, T_max=T, G_max=G, seed=3)
    B = 8
    assert b["txt_ids"].shape == (B, L) and b["txt_ids"].dtype == torch.int64
    assert int(b["txt_lens"].max()) == L and int(b["txt_lens"].min()) >= L // 2
    R = sum(b["traj_step_lens"])
    assert b["traj_view_img_fts"].shape == (R, 36, 768) and b["traj_loc_fts"].shape == (R, 36, 7)
    assert b["gmap_step_ids"].shape[1] == G == int(b["gmap_lens"].max())       # one sample fills the graph budget
    assert max(b["traj_step_lens"]) == T
    d = b["gmap_pair_dists"]
    assert torch.equal(d, d.transpose(1, 2)) and float(d[:, 0].abs().max()) == 0 and float(d[:, :, 0].abs().max()) == 0
    for i in range(B):
        n = int(b["gmap_lens"][i])
        assert len(b["gmap_vpids"][i]) == n and b["gmap_vpids"][i][0] is None
        assert not bool(b["gmap_visited_masks"][i, 0])
        assert int(b["gmap_visited_masks"][i].sum()) == b["traj_step_lens"][i]
    if task == "mlm":
        assert (b["txt_labels"] != -1).any(1).all()
    else:
        for i in range(B):
            g, l = int(b["global_act_labels"][i]), int(b["local_act_labels"][i])
            assert not bool(b["gmap_visited_masks"][i, g]) and g < int(b["gmap_lens"][i])
            assert l == 0 or b["traj_nav_types"][sum(b["traj_step_lens"][:i + 1]) - 1, l - 1] == 1
