"""Flat parameter arena: every parameter of a model lives in ONE contiguous fp32 buffer (and so do the
gradients, the AdamW moments and the optional bf16 shadow used by the tensor-core GEMMs).

Layout = [weight-decay group | no-decay group], each in module order, which keeps query/key/value
weights (and their biases) back to back so the packed QKV projection is a single GEMM with no copies.
The grouping rule is the reference's (pretrain_src/optim/misc.py:13-22: names containing 'bias',
'LayerNorm.bias' or 'LayerNorm.weight' get no weight decay).  Parameters stay ordinary nn.Parameters
(state_dict / checkpoints unchanged); their `.data` and `.grad` are views into the arena.
"""
import torch

from ._lib import F32, BF16, call, ptr, stream

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")
ALIGN = 8  # elements: keeps the bf16 shadow 16-byte aligned (TMA) and fp32 32-byte aligned


class ParamArena:
    def __init__(self, model, lowp=False, requires_grad_only=True, with_grads=True):
        named = [(n, p) for n, p in model.named_parameters() if (p.requires_grad or not requires_grad_only)]
        if not named:
            raise ValueError("model has no parameters")
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("ParamArena needs the model on a CUDA device")
        decay = [(n, p) for n, p in named if not any(nd in n for nd in NO_DECAY)]
        nodecay = [(n, p) for n, p in named if any(nd in n for nd in NO_DECAY)]
        self.entries = []
        off = 0
        for group in (decay, nodecay):
            for n, p in group:
                self.entries.append((n, p, off, p.numel()))
                off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            if group is decay:
                self.n_decay = off
        self.total = off
        self.device = dev
        self.flat_p = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.total, dtype=torch.float32, device=dev) if with_grads else None
        self.flat_lowp = torch.zeros(self.total, dtype=torch.bfloat16, device=dev) if lowp else None
        for n, p, o, k in self.entries:
            view = self.flat_p[o:o + k].view(p.shape)
            view.copy_(p.data)
            p.data = view
            if with_grads:
                g = self.flat_g[o:o + k].view(p.shape)
                p._magic_grad = g
                p.grad = g
            if lowp:
                p._magic_lowp = self.flat_lowp[o:o + k].view(p.shape)
        self.model = model
        model._magic_arena = self  # the model's forward calls sync_lowp() through this (stale-shadow guard)
        self._versions = None
        if lowp:
            self.refresh_lowp()

    def zero_grad(self):
        if self.flat_g is not None:
            self.flat_g.zero_()

    def _version_sum(self):
        return sum(p._version for _, p, _, _ in self.entries)

    def refresh_lowp(self):
        """Recompute the bf16 shadow (the weight operand of every tensor-core GEMM) from the fp32 parameters.
        The fused optimizer keeps it current; ANY other in-place parameter write (load_state_dict, manual
        re-initialisation, an EMA copy) must be followed by this -- `sync_lowp()` does it automatically."""
        if self.flat_lowp is not None:
            call("magic_cast", ptr(self.flat_p), F32, ptr(self.flat_lowp), BF16, self.total, stream())
        self._versions = self._version_sum()

    def sync_lowp(self):
        """Refresh the shadow if a parameter was written through torch since the last refresh (autograd version
        counters: `load_state_dict`, `p.copy_()`, `p.data = ...` re-attached by check()).  The fused AdamW writes
        through raw pointers and refreshes the shadow itself, so the training loop never triggers this.
        Called at the start of every model forward and every PretrainStepper.step (a ~20 us host loop)."""
        if self.flat_lowp is None:
            return False
        moved = self.check()
        if moved or self._versions != self._version_sum():
            self.refresh_lowp()
            return True
        return False

    def check(self):
        """Re-attach if something (e.g. .to(), load_state_dict on a new storage) detached the views.
        -> True if any parameter had to be copied back into the arena."""
        moved = False
        for n, p, o, k in self.entries:
            if p.data.data_ptr() != self.flat_p.data_ptr() + o * 4:
                view = self.flat_p[o:o + k].view(p.shape)
                view.copy_(p.data)
                p.data = view
                moved = True
            if self.flat_g is not None and (p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + o * 4):
                p.grad = p._magic_grad
        if moved and self.flat_lowp is not None:
            self.refresh_lowp()
        return moved

    def release(self):
        if getattr(self.model, "_magic_arena", None) is self:
            del self.model._magic_arena
        for n, p, o, k in self.entries:
            p.data = p.data.clone()
            p.grad = None
            for a in ("_magic_grad", "_magic_lowp"):
                if hasattr(p, a):
                    delattr(p, a)
