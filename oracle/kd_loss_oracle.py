"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the reference KD loss primitives.

Follows /root/reference/pretrain_src/optim/kd_loss.py:5-54 (the `mean` reductions used in pretraining)
and /root/reference/map_nav_src/utils/kd_loss.py:6-66 (the fine-tune variant with loss_type in
{sum, mean} that raises on a weight/shape mismatch).  PINNED: tests/test_oracle_kd_pinned.py compares
every function here against tests/golden/kd_loss_ref.pt, which tests/golden/gen_kd_golden.py produced by
importing the reference file itself in the build container (and re-checks against the live reference
file whenever /root/reference is present).
"""
import torch

NEG_FILL = -1e6  # kd_loss.py:21-22: -inf logits are replaced by -1e6 before the softmax


def _weighted(per_elem, w, strict):
    if w is None:
        return per_elem
    if per_elem.shape[0] == w.shape[0]:  # kd_loss.py:11-13
        return per_elem * w.view(-1, *([1] * (per_elem.dim() - 1)))
    if strict:  # map_nav_src/utils/kd_loss.py:16-17
        raise ValueError("Shape mismatch between sample weights and inputs")
    return per_elem  # kd_loss.py:15-16: silently unweighted


def mse_loss(s_inputs, t_inputs, t_sample_weights=None, loss_type=None, **kwargs):
    """loss_type None -> pretrain semantics (mean, silent fallback); 'sum'/'mean' -> fine-tune semantics."""
    sq = (s_inputs - t_inputs) ** 2
    sq = _weighted(sq, t_sample_weights, strict=loss_type is not None)
    if loss_type in (None, "mean"):
        return sq.mean()
    if loss_type == "sum":
        return sq.sum()
    raise ValueError("Unsupported loss_type. Choose 'sum' or 'mean'.")


def _kl_pointwise(logq, p):
    # ATen kl_div (log_target=False): xlogy(p, p) - p * logq ; p == 0 contributes exactly 0
    return torch.xlogy(p, p) - p * logq


def kd_loss(student_logits, teacher_logits, temperature=1, epsilon=1e-6, t_sample_weights=None,
            loss_type=None, **kwargs):
    s = torch.where(student_logits == float("-inf"), torch.full_like(student_logits, NEG_FILL), student_logits)
    t = torch.where(teacher_logits == float("-inf"), torch.full_like(teacher_logits, NEG_FILL), teacher_logits)
    p = torch.softmax(t / temperature, dim=1)
    logq = torch.log_softmax(s / temperature, dim=1)
    kl = _kl_pointwise(logq, p)
    scale = temperature ** 2
    if t_sample_weights is None:
        # nn.KLDivLoss(reduction='mean') averages over ALL elements (B*C), kd_loss.py:29
        red = kl.sum() if loss_type == "sum" else kl.mean()
        return red * scale
    per_row = kl.sum(1) * t_sample_weights.view(-1)  # kd_loss.py:31-40
    red = per_row.sum() if loss_type == "sum" else per_row.mean()
    return red * scale


def exponential_decay(t_sample_losses, decay_rate=0.1):
    return torch.exp(-decay_rate * t_sample_losses)  # kd_loss.py:43-44


def invert_normalized_losses(t_sample_losses, **kwargs):
    lo, hi = torch.min(t_sample_losses), torch.max(t_sample_losses)  # kd_loss.py:46-54
    return 1 - (t_sample_losses - lo) / (hi - lo)
