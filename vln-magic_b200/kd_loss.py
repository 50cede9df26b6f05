"""CUDA-backed drop-in for the reference KD loss primitives, same names / argument meaning / error
behaviour as pretrain_src/optim/kd_loss.py:5-54 (and the fine-tune variant map_nav_src/utils/kd_loss.py:6-66
when `loss_type` is passed; `nav_mse_loss` / `nav_kd_loss` are that file's functions WITH its default
`loss_type='sum'`).  All arithmetic runs in the fused MAKD kernels (csrc/makd.cu).

Deviation, by design: the teacher argument never receives a gradient (every call site of the reference passes a
detached teacher tensor, agent.py:566-704); passing a teacher tensor that requires grad raises instead of silently
returning no gradient."""
import torch

from . import ops
from ._lib import call, ptr, stream


def _no_teacher_grad(t):
    if t.requires_grad and torch.is_grad_enabled():
        raise ValueError("the teacher argument of the fused KD losses is treated as a constant: pass t.detach() "
                         "(as every reference call site does, agent.py:566-704)")
    return t


def mse_loss(s_inputs, t_inputs, t_sample_weights=None, loss_type=None, **kwargs):
    t_inputs = _no_teacher_grad(t_inputs)
    w = t_sample_weights
    if w is not None and s_inputs.shape[0] != w.shape[0]:
        if loss_type is not None:  # map_nav_src/utils/kd_loss.py:16-17
            raise ValueError("Shape mismatch between sample weights and inputs")
        w = None  # pretrain_src/optim/kd_loss.py:15-16: silently unweighted
    if loss_type not in (None, "mean", "sum"):
        raise ValueError("Unsupported loss_type. Choose 'sum' or 'mean'.")
    scale = 1.0 if loss_type == "sum" else 1.0 / max(s_inputs.numel(), 1)
    t = t_inputs if t_inputs.shape == s_inputs.shape else t_inputs.expand_as(s_inputs)
    per_seg, _ = ops.makd_mse([(s_inputs, t.detach(), w, scale)])
    return per_seg[0]


def kd_loss(student_logits, teacher_logits, temperature=1, epsilon=1e-6, t_sample_weights=None, loss_type=None,
            **kwargs):
    teacher_logits = _no_teacher_grad(teacher_logits)
    if loss_type not in (None, "mean", "sum"):
        raise ValueError("Unsupported loss_type. Choose 'sum' or 'mean'.")
    R, C = student_logits.shape
    t2 = float(temperature) ** 2
    if t_sample_weights is None:
        scale = t2 if loss_type == "sum" else t2 / (R * C)  # nn.KLDivLoss('mean') averages over B*C, kd_loss.py:29
    else:
        scale = t2 if loss_type == "sum" else t2 / R       # per-row sum, weight, mean over rows, kd_loss.py:31-40
    return ops.makd_kl(student_logits, teacher_logits, temperature, t_sample_weights, scale)


def nav_mse_loss(s_inputs, t_inputs, t_sample_weights=None, loss_type="sum", **kwargs):
    """map_nav_src/utils/kd_loss.py:6-25 (default reduction 'sum', raises on a weight / batch mismatch)."""
    return mse_loss(s_inputs, t_inputs, t_sample_weights, loss_type=loss_type, **kwargs)


def nav_kd_loss(student_logits, teacher_logits, temperature=1, epsilon=1e-6, t_sample_weights=None, loss_type="sum",
                **kwargs):
    """map_nav_src/utils/kd_loss.py:27-54 (default reduction 'sum')."""
    return kd_loss(student_logits, teacher_logits, temperature, epsilon, t_sample_weights, loss_type=loss_type,
                   **kwargs)


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
