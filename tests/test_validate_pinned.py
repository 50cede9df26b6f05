"""The validation hooks pinned against the reference's own code: tests/golden/validate_ref.pt
holds canned `compute_loss=False` model outputs and the metrics the SOURCE of `validate_mlm / validate_mrc /
validate_sap / validate_cfp` (pretrain_src/train_r2r_magic.py:440-577) computed from them
(tests/golden/gen_validate_golden.py).  GPU: `train_loop.validate` on the same replayed outputs gives the same
numbers under the reference's metric names (losses 1e-5 relative + 1e-6, accuracies exact)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import gen_validate_golden as G  # noqa: E402

GOLD = torch.load(os.path.join(HERE, "golden", "validate_ref.pt"))


@pytest.mark.skipif(not os.path.exists(G.REF), reason="reference tree not mounted")
def test_fixture_is_what_the_reference_source_produces():
    data = G.canned()
    assert torch.equal(data["sap"][0][1]["global_logits"], GOLD["data"]["sap"][0][1]["global_logits"])
    live = G.run_reference(data)
    for task, d in GOLD["ref"].items():
        assert set(live[task]) == set(d)
        for k, v in d.items():
            assert abs(live[task][k] - v) <= 1e-9 * abs(v) + 1e-12, (task, k)


def test_fixture_covers_the_reference_metric_names():
    assert set(GOLD["ref"]["mlm"]) == {"loss", "acc"} and set(GOLD["ref"]["mrc"]) == {"loss", "acc"}
    for task in ("sap", "cfp"):
        assert set(GOLD["ref"][task]) == {"gloss", "lloss", "floss", "gacc", "lacc", "facc"}


class _Model(torch.nn.Module):
    """`validate` needs .eval() / .train() / .training; the outputs come from one Replay per task."""

    def __init__(self, replays):
        super().__init__()
        self.replays = replays

    def forward(self, batch, task=None, compute_loss=True):
        return self.replays[task](batch, task=task, compute_loss=compute_loss)


@pytest.mark.gpu
def test_validate_hooks_match_the_reference_functions():
    from magic_b200.train_loop import validate
    replays = {t: G.Replay(GOLD["data"][t], "cuda") for t in ("mlm", "mrc", "sap", "cfp")}
    model = _Model(replays).train()
    loaders = {t: r.loader() for t, r in replays.items()}
    out, best, flag = validate(model, loaders, setname="_unseen", max_metrix=0.0, tem=GOLD["temperature"])
    assert model.training
    for task, d in GOLD["ref"].items():
        for k, v in d.items():
            got = out[f"val_unseen_{task}_{k}"]
            if k.endswith("acc"):
                assert got == v, (task, k, got, v)
            else:
                assert abs(got - v) <= 1e-5 * abs(v) + 1e-6, (task, k, got, v)
        rate = [k for k in out if k.startswith(f"val_unseen_{task}_") and k.endswith("_per_s")]
        assert len(rate) == 1 and out[rate[0]] > 0
    # train_r2r_magic.py:423-426: the best unseen fused SAP accuracy is tracked with >=
    assert flag and best == GOLD["ref"]["sap"]["facc"]
