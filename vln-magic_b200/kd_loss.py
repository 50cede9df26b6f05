"""CUDA-backed drop-in for the reference KD loss primitives, same names / argument meaning / error
behaviour as pretrain_src/optim/kd_loss.py:5-54 (and the fine-tune variant map_nav_src/utils/kd_loss.py:6-66
when `loss_type` is passed).  All arithmetic runs in the fused MAKD kernels (csrc/makd.cu)."""
import torch

from . import ops
from ._lib import call, ptr, stream


def mse_loss(s_inputs, t_inputs, t_sample_weights=None, loss_type=None, **kwargs):
    w = t_sample_weights
    if w is not None and s_inputs.shape[0] != w.shape[0]:
        if loss_type is not None:  # map_nav_src/utils/kd_loss.py:16-17
            raise ValueError("Shape mismatch between sample weights and inputs")
        w = None  # pretrain_src/optim/kd_loss.py:15-16: silently unweighted
    if loss_type not in (None, "mean", "sum"):
        raise ValueError("Unsupported loss_type. Choose 'sum' or 'mean'.")
    scale = 1.0 if loss_type == "sum" else 1.0 / max(s_inputs.numel(), 1)
    t = t_inputs if t_inputs.shape == s_inputs.shape else t_inputs.expand_as(s_inputs)
    per_seg, _ = ops.makd_mse([(s_inputs, t.detach() if not t.requires_grad else t, w, scale)])
    return per_seg[0]


def kd_loss(student_logits, teacher_logits, temperature=1, epsilon=1e-6, t_sample_weights=None, loss_type=None,
            **kwargs):
    R, C = student_logits.shape
    t2 = float(temperature) ** 2
    if t_sample_weights is None:
        scale = t2 if loss_type == "sum" else t2 / (R * C)  # nn.KLDivLoss('mean') averages over B*C, kd_loss.py:29
    else:
        scale = t2 if loss_type == "sum" else t2 / R       # per-row sum, weight, mean over rows, kd_loss.py:31-40
    return ops.makd_kl(student_logits, teacher_logits, temperature, t_sample_weights, scale)


def exponential_decay(t_sample_losses, decay_rate=0.1):
    x = t_sample_losses.detach().contiguous().float()
    out = torch.empty_like(x)
    call("magic_exp_decay", ptr(x), ptr(out), x.numel(), float(decay_rate), stream())
    return out


def invert_normalized_losses(t_sample_losses, **kwargs):
    x = t_sample_losses.detach().contiguous().float()
    out = torch.empty_like(x)
    call("magic_invert_norm", ptr(x), ptr(out), x.numel(), stream())
    return out
