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
        if lowp:
            self.refresh_lowp()

    def zero_grad(self):
        if self.flat_g is not None:
            self.flat_g.zero_()

    def refresh_lowp(self):
        if self.flat_lowp is not None:
            call("magic_cast", ptr(self.flat_p), F32, ptr(self.flat_lowp), BF16, self.total, stream())

    def check(self):
        """Re-attach if something (e.g. .to(), load_state_dict on a new storage) detached the views."""
        for n, p, o, k in self.entries:
            if p.data.data_ptr() != self.flat_p.data_ptr() + o * 4:
                view = self.flat_p[o:o + k].view(p.shape)
                view.copy_(p.data)
                p.data = view
            if self.flat_g is not None and (p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + o * 4):
                p.grad = p._magic_grad

    def release(self):
        for n, p, o, k in self.entries:
            p.data = p.data.clone()
            p.grad = None
            for a in ("_magic_grad", "_magic_lowp"):
                if hasattr(p, a):
                    delattr(p, a)
