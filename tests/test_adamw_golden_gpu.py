"""FusedAdamW (clip + AdamW + LR schedule on the flat arena, csrc/optim.cu) against the fixture produced by the
REFERENCE's own `optim/adamw.py` + `optim/misc.py` + `optim/sched.py` (tests/golden/gen_adamw_golden.py): six steps,
two of them clipped at grad_norm 5.0, warm-up-linear learning rate.  fp32, 1e-6 relative per parameter."""
import os
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

from magic_b200 import optim as MO  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adamw_ref.pt")


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(24, 40)
        self.LayerNorm = nn.LayerNorm(40)
        self.emb = nn.Embedding(17, 8)
        self.out = nn.Linear(40, 3, bias=False)


@pytest.mark.parametrize("lowp", [False, True])
def test_fused_adamw_matches_reference_fixture(lowp):
    gold = torch.load(GOLD)
    model = Tiny().cuda()
    model.load_state_dict(gold["init"])
    opts = SimpleNamespace(**gold["opts"])
    opt = MO.build_optimizer(model, opts, lowp=lowp)
    arena = opt.arena
    assert sorted(n for n, p, o, k in arena.entries if o >= arena.n_decay) == sorted(gold["no_decay"])
    params = dict(model.named_parameters())
    for rec in gold["steps"]:
        opt.zero_grad()
        for n, g in rec["grads"].items():
            params[n]._magic_grad.copy_(g)
        opt.step(MO.get_lr_sched(rec["step"], opts))
        assert abs(opt.grad_norm() - rec["grad_norm"]) <= 1e-5 * rec["grad_norm"]
        for n, ref in rec["params"].items():
            got = params[n].detach().cpu()
            err = ((got - ref).norm() / ref.norm()).item()
            assert err < 1e-6, (rec["step"], n, err)
        if lowp:  # the bf16 shadow (tensor-core operand) follows the fp32 parameters
            for n, p in params.items():
                assert torch.equal(p._magic_lowp, p.detach().to(torch.bfloat16)), n


def test_shadow_follows_load_state_dict():
    """an in-place parameter write after the arena exists (checkpoint resume) must reach the bf16
    shadow before the next forward / step."""
    gold = torch.load(GOLD)
    model = Tiny().cuda()
    opts = SimpleNamespace(**gold["opts"])
    opt = MO.build_optimizer(model, opts, lowp=True)
    arena = opt.arena
    assert not arena.sync_lowp()
    model.load_state_dict(gold["init"])
    assert arena.sync_lowp()          # noticed through the parameters' version counters
    for n, p in model.named_parameters():
        assert torch.equal(p._magic_lowp.cpu(), gold["init"][n].to(torch.bfloat16)), n
    assert not arena.sync_lowp()
    moved = nn.Parameter(torch.zeros_like(model.out.weight))  # storage swapped: check() re-attaches and refreshes
    model.out.weight.data = moved.data
    assert arena.sync_lowp()
    assert model.out.weight.data_ptr() == arena.flat_p.data_ptr() + \
        [o for n, p, o, k in arena.entries if n == "out.weight"][0] * 4
    assert float(model.out.weight._magic_lowp.float().abs().max()) == 0.0


class TwoHeads(nn.Module):
    def __init__(self):
        super().__init__()
        self.trunk = nn.Linear(24, 40)
        self.LayerNorm = nn.LayerNorm(40)
        self.head_a = nn.Linear(40, 5)
        self.head_b = nn.Linear(40, 7)
        self.never = nn.Linear(8, 8)


@pytest.mark.parametrize("lowp", [False, True])
def test_task_inactive_parameters_follow_the_reference(lowp):
    """Multi-task loop: the other task's head has `grad is None` in the reference -- skipped entirely, with its own
    step count (adamw.py:66-67, :86).  FusedAdamW.configure_tasks + magic_adamw_seg reproduce that on the flat arena
    (fixture: tests/golden/gen_adamw_tasks_golden.py, the reference's own optimizer files)."""
    gold = torch.load(os.path.join(os.path.dirname(GOLD), "adamw_tasks_ref.pt"))
    model = TwoHeads().cuda()
    model.load_state_dict(gold["init"])
    opts = SimpleNamespace(**gold["opts"])
    opt = MO.build_optimizer(model, opts, lowp=lowp)
    inactive = lambda task, name: name.startswith("never.") or name.startswith("head_b." if task == "a" else "head_a.")
    opt.configure_tasks(["a", "b"], inactive)
    assert len(opt.slots) == 3 and opt.slots[0] == frozenset("ab")
    params = dict(model.named_parameters())
    for rec in gold["steps"]:
        opt.zero_grad()
        for n, g in rec["grads"].items():
            params[n]._magic_grad.copy_(g)
        opt.step(MO.get_lr_sched(rec["step"], opts), task=rec["task"])
        assert abs(opt.grad_norm() - rec["grad_norm"]) <= 1e-5 * rec["grad_norm"]
        for n, ref in rec["params"].items():
            got = params[n].detach().cpu()
            err = ((got - ref).norm() / ref.norm()).item()
            assert err < 1e-6, (rec["step"], rec["task"], n, err)
        for n in ("never.weight", "never.bias"):   # untouched by every task: bit-identical to the initial values
            assert torch.equal(params[n].detach().cpu(), gold["init"][n])
        if lowp:
            for n, p in params.items():
                assert torch.equal(p._magic_lowp, p.detach().to(torch.bfloat16)), n
    # without the task tables every element would have been decayed: the fixture tells the two apart
    a_w = gold["steps"][1]["params"]["head_a.weight"]
    assert torch.equal(a_w, gold["steps"][0]["params"]["head_a.weight"])   # step 2 is task b: head_a untouched
