"""Pins oracle/kd_loss_oracle.py against the REFERENCE's kd_loss.py: via the committed golden file produced by
importing the reference (tests/golden/gen_kd_golden.py) and, when /root/reference is present, live."""
import importlib.util
import os

import pytest
import torch

from oracle import kd_loss_oracle as KO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kd_loss_ref.pt")


def _check(rec, pre_fn, fin_fn):
    a = rec["args"]
    if rec["kind"] == "mse":
        got = KO.mse_loss(a["s"], a["t"], a["w"])
        assert torch.allclose(got, pre_fn, rtol=1e-6, atol=0), (got, pre_fn)
        for lt in ("sum", "mean"):
            exp = fin_fn[lt]
            if isinstance(exp, str):
                with pytest.raises(ValueError):
                    KO.mse_loss(a["s"], a["t"], a["w"], loss_type=lt)
            else:
                assert torch.allclose(KO.mse_loss(a["s"], a["t"], a["w"], loss_type=lt), exp, rtol=1e-6)
    elif rec["kind"] == "kd":
        got = KO.kd_loss(a["s"], a["t"], temperature=a["T"], t_sample_weights=a["w"])
        assert torch.allclose(got, pre_fn, rtol=1e-5, atol=1e-8), (got, pre_fn)
        for lt in ("sum", "mean"):
            assert torch.allclose(KO.kd_loss(a["s"], a["t"], temperature=a["T"], t_sample_weights=a["w"], loss_type=lt),
                                  fin_fn[lt], rtol=1e-5, atol=1e-8)
    elif rec["kind"] == "exp":
        assert torch.equal(KO.exponential_decay(a["x"], a["rate"]), pre_fn)
    else:
        assert torch.equal(KO.invert_normalized_losses(a["x"]), pre_fn)


def test_oracle_matches_golden_from_reference():
    recs = torch.load(GOLD)
    assert len(recs) >= 10
    for rec in recs:
        _check(rec, rec["pretrain"], {lt: rec.get("finetune_" + lt) for lt in ("sum", "mean")})


@pytest.mark.skipif(not os.path.exists("/root/reference/pretrain_src/optim/kd_loss.py"), reason="reference not mounted")
def test_oracle_matches_live_reference():
    def load(path, name):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    pre = load("/root/reference/pretrain_src/optim/kd_loss.py", "ref_kd_pre_live")
    g = torch.Generator().manual_seed(7)
    for _ in range(5):
        B, C = 5, 33
        s, t = torch.randn(B, C, generator=g) * 2, torch.randn(B, C, generator=g) * 2
        s[:, :3] = float("-inf")
        t[:, :3] = float("-inf")
        w = torch.rand(B, generator=g)
        for ww in (None, w):
            assert torch.allclose(KO.kd_loss(s, t, temperature=2, t_sample_weights=ww),
                                  pre.kd_loss(s, t, temperature=2, t_sample_weights=ww), rtol=1e-5, atol=1e-8)
        x, y = torch.randn(B, 4, 9, generator=g), torch.randn(B, 4, 9, generator=g)
        for ww in (None, w, w[:2]):
            assert torch.allclose(KO.mse_loss(x, y, ww), pre.mse_loss(x, y, ww), rtol=1e-6)
