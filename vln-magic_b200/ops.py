"""Autograd wrappers around the C ABI (include/magic_b200.h).  Every op is a hand-written sm_100a kernel;
torch only provides device memory, streams and the autograd tape.

Parameter gradients: if a parameter carries a `_magic_grad` tensor (a view into the flat gradient arena,
see arena.py) the kernels ACCUMULATE straight into it and autograd receives None for that parameter;
otherwise a fresh zero buffer is filled and returned to autograd (works under stock DDP).
A parameter may carry `_magic_lowp` (bf16 shadow) which the GEMMs use as their weight operand.
"""
import math

import torch

from . import _lib as L
from ._lib import call, dt, ptr, stream

_SEED = {}
_DROP_ON = True


def seed_tensor(device):
    """Device-resident 64-bit dropout seed (bumped once per step by `bump_seed`)."""
    key = str(device)
    if key not in _SEED:
        _SEED[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return _SEED[key]


def set_seed(device, value):
    seed_tensor(device).fill_(int(value))


def bump_seed(device):
    seed_tensor(device).add_(0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF)


class PinnedRing:
    """Host -> device upload of a few per-step scalars without a host sync and without the overwrite race of a
    single pinned buffer: the host may run several steps ahead of the GPU, so each upload uses the next of `n`
    pinned slots and a slot is only rewritten after the copy that last read it has completed."""

    def __init__(self, numel, dtype=torch.float32, n=8):
        self.bufs = [torch.zeros(numel, dtype=dtype).pin_memory() for _ in range(n)]
        self.events = [None] * n
        self.i = 0

    def upload(self, values, dst):
        k = self.i % len(self.bufs)
        self.i += 1
        if self.events[k] is not None:
            self.events[k].synchronize()
        buf = self.bufs[k]
        buf.copy_(torch.as_tensor(values, dtype=buf.dtype).reshape(-1))
        dst.copy_(buf, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev
        return dst


_salt = [0]


def next_salt():
    _salt[0] += 1
    return _salt[0]


def reset_salt():
    _salt[0] = 0


def _sink(p, shape=None):
    """-> (buffer to accumulate into, value to hand back to autograd, beta for GEMM accumulation)."""
    g = getattr(p, "_magic_grad", None)
    if g is not None:
        return g, None, 1.0
    buf = torch.zeros(p.shape if shape is None else shape, dtype=torch.float32, device=p.device)
    return buf, buf, 0.0


def _wop(w):
    lw = getattr(w, "_magic_lowp", None)
    return lw if lw is not None else w


def gemm(A, sam, sak, B, sbk, sbn, C, M, N, K, bias=None, act=0, pre_out=None, dact_pre=None, residual=None,
         alpha=1.0, beta=0.0, drop_p=0.0, salt=0, seed=None, allow_tc=1, ldc=None):
    call("magic_gemm", ptr(A), dt(A), sam, sak, ptr(B), dt(B), sbk, sbn, ptr(C), dt(C),
         C.stride(0) if ldc is None else ldc, M, N, K,
         ptr(bias), act, ptr(pre_out), ptr(dact_pre), dt(dact_pre) if dact_pre is not None else 0,
         dact_pre.stride(0) if dact_pre is not None else 0, ptr(residual),
         residual.stride(0) if residual is not None else 0, alpha, beta, drop_p, salt, ptr(seed), allow_tc, stream())


def _lin_fwd(x2, w, bias, out, act=0, pre_out=None, residual=None, drop_p=0.0, salt=0, seed=None):
    M, K = x2.shape
    N = w.shape[0]
    wo = _wop(w)
    gemm(x2, K, 1, wo, 1, K, out, M, N, K, bias=bias, act=act, pre_out=pre_out, residual=residual, drop_p=drop_p,
         salt=salt, seed=seed)


def _rows2d(t):
    """2-D view usable as a GEMM operand without a copy when the rows are unit-stride (padded ld allowed)."""
    if not (t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]):
        t = t.reshape(-1, t.shape[-1]).contiguous()
    if t.dtype == torch.bfloat16 and t.shape[1] >= 4096 and t.stride(0) % 8 != 0:
        # a vocabulary-wide gradient that autograd re-materialised densely (CE + KD both feed the MLM logits, the
        # engine sums them into a fresh [rows, 50265] tensor): re-pad the rows to 16 B so the tcgen05 path takes it
        # (the FFMA fallback costs 4.4 ms per GEMM at the teacher-distillation sizes, the copy 40 us)
        pad = torch.empty(t.shape[0], (t.shape[1] + 7) // 8 * 8, dtype=t.dtype, device=t.device)[:, :t.shape[1]]
        pad.copy_(t)
        t = pad
    return t


def _lin_dgrad(dy2, w, dx, act=0, dact_pre=None, drop_p=0.0, salt=0, seed=None, residual=None):
    M, N = dy2.shape
    K = w.shape[1]
    wo = _wop(w)
    gemm(dy2, dy2.stride(0), 1, wo, K, 1, dx, M, K, N, act=act, dact_pre=dact_pre, drop_p=drop_p, salt=salt,
         seed=seed, residual=residual)


class GradLink:
    """Side channel between the two consumers of a sublayer input x in `y = LN(x + f(x))`: LayerNorm's backward
    parks the residual-branch gradient here (and reports None to autograd), the first linear of f picks it up and
    adds it in the epilogue of its dgrad GEMM, so x receives ONE gradient and autograd launches no `add` kernel.
    LayerNorm's backward always runs before f's first linear (f's output feeds the LayerNorm)."""
    __slots__ = ("dres",)

    def __init__(self):
        self.dres = None

    def take(self, like):
        d, self.dres = self.dres, None
        if d is not None and (d.dtype != like.dtype or d.shape != like.shape or not d.is_contiguous()):
            d = d.to(like.dtype).reshape(like.shape).contiguous()
        return d


# ---------------------------------------------------------------------------------------------------
# weight-gradient side stream: dW / db are only needed by the optimizer, so when every parameter
# accumulates straight into the gradient arena the wgrad GEMMs and bias column-sums are issued on a second
# stream and overlap the dgrad chain (under CUDA-graph capture they become a parallel branch of the graph).
# Operands are kept alive until `join_side_stream()` so the caching allocator cannot recycle them early.
# ---------------------------------------------------------------------------------------------------
_SIDE = {"on": False, "stream": None, "hold": [], "used": False}


def enable_side_stream(flag=True):
    _SIDE["on"] = bool(flag)


def _side_stream(*sinks):
    """-> the side stream (already ordered after the current stream) or None."""
    if not _SIDE["on"] or any(getattr(p, "_magic_grad", None) is None for p in sinks if p is not None):
        return None
    if _SIDE["stream"] is None:
        _SIDE["stream"] = torch.cuda.Stream(priority=getattr(torch.cuda.current_stream(), "priority", 0))
    _SIDE["stream"].wait_stream(torch.cuda.current_stream())
    _SIDE["used"] = True
    return _SIDE["stream"]


def join_side_stream():
    """Order the current stream after all side-stream work issued so far (call before the optimizer)."""
    if _SIDE["used"]:
        torch.cuda.current_stream().wait_stream(_SIDE["stream"])
        _SIDE["used"] = False
    _SIDE["hold"].clear()


# ---------------------------------------------------------------------------------------------------
# independent sub-networks as concurrent stream branches
# ---------------------------------------------------------------------------------------------------
_BRANCH = {"on": True, "stream": None}


def enable_branch_streams(flag=True):
    _BRANCH["on"] = bool(flag)


def _record(obj, stream):
    """Tell the caching allocator that tensors produced on the branch stream are consumed on `stream`."""
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, (tuple, list)):
        for o in obj:
            _record(o, stream)
    elif isinstance(obj, dict):
        for o in obj.values():
            _record(o, stream)


_ATTN_STREAMS = {}


def _attn_stream(main):
    """A second stream per launching stream for the key-major half of the attention backward."""
    k = main.cuda_stream
    if k not in _ATTN_STREAMS:
        _ATTN_STREAMS[k] = torch.cuda.Stream(priority=getattr(main, "priority", 0))
    return _ATTN_STREAMS[k]


def run_branches(side_fn, main_fn):
    """Run two independent closures concurrently: `side_fn` on a second stream, `main_fn` on the current one,
    then join.  Returns (side_result, main_result).  The call order (side first) is fixed, so stateful host-side
    counters (dropout salts) do not depend on whether branching is enabled."""
    if not _BRANCH["on"]:
        a = side_fn()
        return a, main_fn()
    main = torch.cuda.current_stream()
    # one side stream per LAUNCHING stream, so branches nest (teacher || student, and inside each text || panorama)
    # without sharing a stream and picking up each other's work as false dependencies
    pool = _BRANCH.setdefault("pool", {})
    side = pool.get(main.cuda_stream)
    if side is None:
        # (a branch inherits the priority of the stream it forks from: the training graph is captured on a high-
        # priority stream, the frozen teacher's on a normal one -- train_step.PretrainStepper)
        side = pool[main.cuda_stream] = torch.cuda.Stream(priority=getattr(main, "priority", 0))
    side.wait_stream(main)
    with torch.cuda.stream(side):
        a = side_fn()
    b = main_fn()
    main.wait_stream(side)
    _record(a, main)
    return a, b


def gemm_wgrad(dy2, x2, gw, gb, M, N, K, beta, dw_ld=None):
    """gw[N,K] = beta*gw + dy^T x ; gb[N] += colsum(dy) -- one launch (bias gradient rides the wgrad GEMM)."""
    call("magic_gemm_wgrad", ptr(dy2), dt(dy2), dy2.stride(0), ptr(x2), dt(x2), x2.stride(0), ptr(gw),
         K if dw_ld is None else dw_ld, ptr(gb), M, N, K, beta, 1, stream())


def _lin_wgrad(dy2, x2, w, bias):
    """dW[N,K] (+)= dy^T x ; db (+)= colsum(dy).  Returns the autograd values."""
    M, N = dy2.shape
    K = x2.shape[1]
    side = _side_stream(w, bias)
    if side is not None:
        _SIDE["hold"].extend((dy2, x2))
        with torch.cuda.stream(side):
            gemm_wgrad(dy2, x2, w._magic_grad, bias._magic_grad if bias is not None else None, M, N, K, 1.0)
        return None, None
    gw, rw, beta = _sink(w)
    gb, rb = None, None
    if bias is not None:
        gb, rb, _ = _sink(bias)
    gemm_wgrad(dy2, x2, gw, gb, M, N, K, beta)
    return rw, rb


class LinearFn(torch.autograd.Function):
    """y = drop(act(x W^T + b)) (+ residual)."""

    @staticmethod
    def forward(ctx, x, w, bias, act, residual, drop_p, salt, pad_out=False, link=None):
        ctx.link = link
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        M, N = x2.shape[0], w.shape[0]
        if pad_out and N % 8 != 0:
            # leading dimension padded to 16 bytes so the gradient of this output is a legal TMA operand
            out = torch.empty(M, (N + 7) // 8 * 8, dtype=x.dtype, device=x.device)[:, :N]
        else:
            out = torch.empty(M, N, dtype=x.dtype, device=x.device)
        # the pre-activation copy is only needed by backward: a no-grad forward (frozen teacher, validation) skips the
        # second output tile of the epilogue and its HBM write
        pre = torch.empty_like(out) if (act != 0 and any(ctx.needs_input_grad)) else None
        seed = seed_tensor(x.device) if drop_p > 0 else None
        res2 = residual.reshape(-1, N).contiguous() if residual is not None else None
        _lin_fwd(x2, w, bias, out, act, pre, res2, drop_p, salt, seed)
        ctx.save_for_backward(x2, w, bias, pre)
        ctx.meta = (act, drop_p, salt, x.shape, residual is not None)
        return out if x.dim() == 2 else out.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, bias, pre = ctx.saved_tensors
        act, drop_p, salt, xshape, has_res = ctx.meta
        dy2 = _rows2d(dy) if (act == 0 and drop_p == 0) else dy.reshape(-1, dy.shape[-1]).contiguous()
        dres = dy if has_res else None
        if act != 0 or drop_p > 0:
            dz = torch.empty_like(dy2)
            seed = seed_tensor(dy.device) if drop_p > 0 else None
            call("magic_act_bwd", ptr(dy2), ptr(pre) if pre is not None else ptr(dy2), ptr(dz), dz.numel(), act,
                 dt(dz), drop_p, salt, ptr(seed), stream())
        else:
            dz = dy2
        dx = None
        gres = ctx.link.take(x2) if ctx.link is not None else None
        if ctx.needs_input_grad[0]:
            if dz.shape[1] >= 4096 and x2.dtype != torch.float32:
                # vocabulary-sized reduction with a tiny output (MLM decoder dgrad): fp32 C so the GEMM may split K
                dx32 = torch.empty(x2.shape, dtype=torch.float32, device=x2.device)
                _lin_dgrad(dz, w, dx32)
                dx = dx32.to(x2.dtype)
                if gres is not None:
                    dx.add_(gres)
            else:
                dx = torch.empty_like(x2)
                _lin_dgrad(dz, w, dx, residual=gres)
            dx = dx.view(xshape)
        rw, rb = _lin_wgrad(dz, x2, w, bias)
        return dx, rw, rb, None, dres, None, None, None, None


def linear(x, w, bias=None, act=0, residual=None, drop_p=0.0, salt=0, pad_out=False, link=None):
    return LinearFn.apply(x, w, bias, act, residual, drop_p if _DROP_ON else 0.0, salt, pad_out, link)


def _adjacent(ts):
    """True if the tensors are laid out back to back in memory (flat-arena packing)."""
    if any(t is None for t in ts):
        return False
    for a, b in zip(ts[:-1], ts[1:]):
        if a.dtype != b.dtype or b.data_ptr() != a.data_ptr() + a.numel() * a.element_size():
            return False
    return True


class PackedLinearFn(torch.autograd.Function):
    """[y_0 | y_1 | ...] = x [W_0; W_1; ...]^T + [b_0 | b_1 | ...]  into ONE [M, sum N_i] buffer (packed QKV / KV).
    If the weights (and their grad sinks) are adjacent in the arena this is a single GEMM each way."""

    @staticmethod
    def forward(ctx, x, n, link, *wb):
        ctx.link = link
        ws, bs = wb[:n], wb[n:]
        K = x.shape[-1]
        x2 = x.reshape(-1, K).contiguous()
        M = x2.shape[0]
        Ns = [w.shape[0] for w in ws]
        Nt = sum(Ns)
        out = torch.empty(M, Nt, dtype=x.dtype, device=x.device)
        wops = [_wop(w) for w in ws]
        if _adjacent(wops) and _adjacent(list(bs)):
            gemm(x2, K, 1, wops[0], 1, K, out, M, Nt, K, bias=bs[0])
        else:
            off = 0
            for w, b, Ni in zip(wops, bs, Ns):
                gemm(x2, K, 1, w, 1, K, out[:, off:off + Ni], M, Ni, K, bias=b, ldc=Nt)
                off += Ni
        ctx.save_for_backward(x2, *ws, *bs)
        ctx.meta = (n, Ns, x.shape)
        return out

    @staticmethod
    def backward(ctx, dy):
        n, Ns, xshape = ctx.meta
        saved = ctx.saved_tensors
        x2, ws, bs = saved[0], saved[1:1 + n], saved[1 + n:]
        M, K = x2.shape
        Nt = sum(Ns)
        dy2 = dy.reshape(M, Nt).contiguous()
        wops = [_wop(w) for w in ws]
        dx = None
        gres = ctx.link.take(x2) if ctx.link is not None else None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x2)
            if _adjacent(wops):
                gemm(dy2, Nt, 1, wops[0], K, 1, dx, M, K, Nt, residual=gres)
            else:
                off = 0
                for i, (w, Ni) in enumerate(zip(wops, Ns)):
                    gemm(dy2[:, off:off + Ni], Nt, 1, w, K, 1, dx, M, K, Ni, beta=0.0 if i == 0 else 1.0)
                    off += Ni
                if gres is not None:
                    dx.add_(gres)
            dx = dx.view(xshape)
        gws = [getattr(w, "_magic_grad", None) for w in ws]
        gbs = [getattr(b, "_magic_grad", None) for b in bs]
        rws, rbs = [None] * n, [None] * n
        if _adjacent(gws) and _adjacent(gbs):
            side = _side_stream()
            if side is not None:
                _SIDE["hold"].extend((dy2, x2))
            with torch.cuda.stream(side if side is not None else torch.cuda.current_stream()):
                gemm_wgrad(dy2, x2, gws[0], gbs[0], M, Nt, K, 1.0)
        else:
            off = 0
            for i, (w, b, Ni) in enumerate(zip(ws, bs, Ns)):
                gw, rws[i], beta = _sink(w)
                dv = dy2[:, off:off + Ni]
                gb, rbs[i], _ = _sink(b)
                gemm_wgrad(dv, x2, gw, gb, M, Ni, K, beta)
                off += Ni
        return (dx, None, None, *rws, *rbs)


def packed_linear(x, ws, bs, link=None):
    return PackedLinearFn.apply(x, len(ws), link, *ws, *bs)


class FFNFn(torch.autograd.Function):
    """y = drop_inner(act(x W1^T + b1)) W2^T + b2 (+ residual); the activation derivative is fused into the
    epilogue of the dgrad GEMM."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, act, residual, drop_p, salt, drop_out_p, salt_out, link=None):
        ctx.link = link
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        M = x2.shape[0]
        I, N = w1.shape[0], w2.shape[0]
        hmid = torch.empty(M, I, dtype=x.dtype, device=x.device)
        pre = torch.empty_like(hmid) if any(ctx.needs_input_grad) else None  # backward only (see LinearFn)
        seed = seed_tensor(x.device) if (drop_p > 0 or drop_out_p > 0) else None
        _lin_fwd(x2, w1, b1, hmid, act, pre, None, drop_p, salt, seed)
        out = torch.empty(M, N, dtype=x.dtype, device=x.device)
        res2 = residual.reshape(-1, N).contiguous() if residual is not None else None
        _lin_fwd(hmid, w2, b2, out, residual=res2, drop_p=drop_out_p, salt=salt_out, seed=seed)
        ctx.save_for_backward(x2, w1, b1, w2, b2, pre, hmid)
        ctx.meta = (act, drop_p, salt, x.shape, residual is not None, drop_out_p, salt_out)
        return out.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w1, b1, w2, b2, pre, hmid = ctx.saved_tensors
        act, drop_p, salt, xshape, has_res, drop_out_p, salt_out = ctx.meta
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        seed = seed_tensor(dy.device) if (drop_p > 0 or drop_out_p > 0) else None
        if drop_out_p > 0:
            dyd = torch.empty_like(dy2)
            call("magic_act_bwd", ptr(dy2), ptr(dy2), ptr(dyd), dyd.numel(), 0, dt(dyd), drop_out_p, salt_out,
                 ptr(seed), stream())
            dy2 = dyd
        dz = torch.empty_like(pre)
        _lin_dgrad(dy2, w2, dz, act=act, dact_pre=pre, drop_p=drop_p, salt=salt, seed=seed)
        rw2, rb2 = _lin_wgrad(dy2, hmid, w2, b2)
        dx = None
        gres = ctx.link.take(x2) if ctx.link is not None else None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x2)
            _lin_dgrad(dz, w1, dx, residual=gres)
            dx = dx.view(xshape)
        rw1, rb1 = _lin_wgrad(dz, x2, w1, b1)
        return dx, rw1, rb1, rw2, rb2, None, (dy if has_res else None), None, None, None, None, None


def ffn(x, w1, b1, w2, b2, act=L.ACT_GELU, residual=None, drop_p=0.0, salt=0, drop_out_p=0.0, salt_out=0, link=None):
    if not _DROP_ON:
        drop_p = drop_out_p = 0.0
    return FFNFn.apply(x, w1, b1, w2, b2, act, residual, drop_p, salt, drop_out_p, salt_out, link)


class LayerNormFn(torch.autograd.Function):
    """y = drop_out(LN(drop_in(x) + res))."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps, p_in, salt_in, p_out, salt_out, link=None):
        ctx.link = link
        h = x.shape[-1]
        x2 = x.reshape(-1, h).contiguous()
        r2 = res.reshape(-1, h).contiguous() if res is not None else None
        M = x2.shape[0]
        y = torch.empty_like(x2)
        stats = torch.empty(M, 2, dtype=torch.float32, device=x.device)
        seed = seed_tensor(x.device) if (p_in > 0 or p_out > 0) else None
        call("magic_ln_fwd", ptr(x2), ptr(r2), ptr(gamma), ptr(beta), ptr(y), ptr(stats), M, h, eps, dt(x2), p_in,
             salt_in, p_out, salt_out, ptr(seed), stream())
        ctx.save_for_backward(x2, r2, gamma, beta, stats)
        ctx.meta = (p_in, salt_in, p_out, salt_out, x.shape)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, r2, gamma, beta, stats = ctx.saved_tensors
        p_in, salt_in, p_out, salt_out, shape = ctx.meta
        h = x2.shape[1]
        dy2 = dy.reshape(-1, h).contiguous()
        dx = torch.empty_like(x2)
        dres = torch.empty_like(x2) if (r2 is not None and p_in > 0) else None
        gg, rg, _ = _sink(gamma)
        gb, rb, _ = _sink(beta)
        seed = seed_tensor(dy.device) if (p_in > 0 or p_out > 0) else None
        call("magic_ln_bwd", ptr(dy2), ptr(x2), ptr(r2), ptr(gamma), ptr(stats), ptr(dx), ptr(dres), ptr(gg), ptr(gb),
             x2.shape[0], h, dt(x2), p_in, salt_in, p_out, salt_out, ptr(seed), stream())
        dxv = dx.view(shape)
        dresv = None
        if r2 is not None:
            dresv = dres.view(shape) if dres is not None else dxv
            if ctx.link is not None and ctx.needs_input_grad[1]:
                ctx.link.dres = dres if dres is not None else dx  # consumed by the dgrad GEMM of the sublayer's first linear
                dresv = None
        return dxv, dresv, rg, rb, None, None, None, None, None, None


def layer_norm(x, gamma, beta, eps, res=None, p_in=0.0, salt_in=0, p_out=0.0, salt_out=0, link=None):
    if not _DROP_ON:
        p_in = p_out = 0.0
    return LayerNormFn.apply(x, res, gamma, beta, eps, p_in, salt_in, p_out, salt_out, link)


class EmbedLNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, word, pos, typ, gamma, beta, eps, out_dtype, p_out, salt):
        B, Lx = ids.shape
        h = word.shape[1]
        ids = ids.contiguous()
        y = torch.empty(B, Lx, h, dtype=out_dtype, device=ids.device)
        stats = torch.empty(B * Lx, 2, dtype=torch.float32, device=ids.device)
        seed = seed_tensor(ids.device) if p_out > 0 else None
        call("magic_embed_ln_fwd", ptr(ids), ptr(word), ptr(pos), ptr(typ), ptr(gamma), ptr(beta), ptr(y), ptr(stats),
             B * Lx, Lx, h, eps, dt(y), p_out, salt, ptr(seed), stream())
        ctx.save_for_backward(ids, word, pos, typ, gamma, beta, stats)
        ctx.meta = (p_out, salt)
        return y

    @staticmethod
    def backward(ctx, dy):
        ids, word, pos, typ, gamma, beta, stats = ctx.saved_tensors
        p_out, salt = ctx.meta
        B, Lx = ids.shape
        h = word.shape[1]
        dy = dy.contiguous()
        gw, rw, _ = _sink(word)
        gp, rp, _ = _sink(pos)
        gt, rt, _ = _sink(typ)
        gg, rg, _ = _sink(gamma)
        gb, rb, _ = _sink(beta)
        seed = seed_tensor(dy.device) if p_out > 0 else None
        call("magic_embed_ln_bwd", ptr(dy), ptr(ids), ptr(word), ptr(pos), ptr(typ), ptr(gamma), ptr(stats), ptr(gw),
             ptr(gp), ptr(gt), ptr(gg), ptr(gb), B * Lx, Lx, h, dt(dy), p_out, salt, ptr(seed), stream())
        return None, rw, rp, rt, rg, rb, None, None, None, None


def embed_ln(ids, word, pos, typ, gamma, beta, eps, out_dtype, p_out=0.0, salt=0):
    return EmbedLNFn.apply(ids, word, pos, typ, gamma, beta, eps, out_dtype, p_out if _DROP_ON else 0.0, salt)


class PosFuseFn(torch.autograd.Function):
    """y = xin + emb[idx] + cst + LN(W f + b)."""

    @staticmethod
    def forward(ctx, xin, idx, emb, cst, f, W, b, gamma, beta, eps, out_dtype):
        K = f.shape[-1]
        h = W.shape[0]
        f2 = f.reshape(-1, K).contiguous().float()
        M = f2.shape[0]
        x2 = xin.reshape(-1, h).contiguous() if xin is not None else None
        i2 = idx.reshape(-1).contiguous() if idx is not None else None
        y = torch.empty(M, h, dtype=out_dtype, device=f.device)
        stats = torch.empty(M, 2, dtype=torch.float32, device=f.device)
        call("magic_posfuse_fwd", ptr(x2), ptr(i2), ptr(emb), ptr(cst), ptr(f2), ptr(W), ptr(b), ptr(gamma), ptr(beta),
             ptr(y), ptr(stats), M, h, K, eps, dt(y), stream())
        ctx.save_for_backward(i2, emb, cst, f2, W, b, gamma, beta, stats)
        ctx.has_x = xin is not None
        ctx.oshape = (*f.shape[:-1], h)
        return y.view(ctx.oshape)

    @staticmethod
    def backward(ctx, dy):
        i2, emb, cst, f2, W, b, gamma, beta, stats = ctx.saved_tensors
        h, K = W.shape
        dy2 = dy.reshape(-1, h).contiguous()
        ge, re_, _ = _sink(emb) if emb is not None else (None, None, 0)
        gc, rc, _ = _sink(cst) if cst is not None else (None, None, 0)
        gW, rW, _ = _sink(W)
        gb, rb, _ = _sink(b)
        gg, rg, _ = _sink(gamma)
        gbt, rbt, _ = _sink(beta)
        call("magic_posfuse_bwd", ptr(dy2), ptr(i2), ptr(f2), ptr(W), ptr(b), ptr(gamma), ptr(stats), ptr(ge), ptr(gc),
             ptr(gW), ptr(gb), ptr(gg), ptr(gbt), dy2.shape[0], h, K, dt(dy2), stream())
        return (dy if ctx.has_x else None), None, re_, rc, None, rW, rb, rg, rbt, None, None


def posfuse(xin, idx, emb, cst, f, W, b, gamma, beta, eps, out_dtype):
    return PosFuseFn.apply(xin, idx, emb, cst, f, W, b, gamma, beta, eps, out_dtype)


class GatherRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx):
        h = src.shape[-1]
        s2 = src.reshape(-1, h).contiguous()
        idx = idx.contiguous()
        out = torch.empty(idx.numel(), h, dtype=src.dtype, device=src.device)
        call("magic_gather_rows", ptr(s2), ptr(idx), ptr(out), idx.numel(), h, dt(s2), stream())
        ctx.save_for_backward(idx)
        ctx.sshape = src.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        h = dout.shape[-1]
        dout = dout.contiguous()
        n_src = 1
        for s in ctx.sshape[:-1]:
            n_src *= s
        dsrc = torch.empty(ctx.sshape, dtype=dout.dtype, device=dout.device)
        call("magic_scatter_rows", ptr(dout), ptr(idx), ptr(dsrc), idx.numel(), n_src, h, dt(dout), stream())
        return dsrc, None


def gather_rows(src, idx):
    return GatherRowsFn.apply(src, idx)


class PanoFuseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, lens):
        R, V, h = x.shape
        x = x.contiguous()
        fused = torch.empty(R, h, dtype=x.dtype, device=x.device)
        probs = torch.empty(R, V, dtype=torch.float32, device=x.device)
        wv = w.reshape(-1) if w is not None else None
        call("magic_pano_fuse_fwd", ptr(x), ptr(wv), ptr(bias), ptr(lens), ptr(fused), ptr(probs), R, V, h, dt(x),
             stream())
        ctx.save_for_backward(x, w, bias, lens, probs)
        return fused

    @staticmethod
    def backward(ctx, dfused):
        x, w, bias, lens, probs = ctx.saved_tensors
        R, V, h = x.shape
        dfused = dfused.contiguous()
        dx = torch.empty_like(x)
        gw = rw = gb = rb = None
        if w is not None:
            gw, rw, _ = _sink(w)
            gb, rb, _ = _sink(bias)
        call("magic_pano_fuse_bwd", ptr(dfused), ptr(x), ptr(w), ptr(lens), ptr(probs), ptr(dx), ptr(gw), ptr(gb), R, V,
             h, dt(x), stream())
        return dx, rw, rb, None


def pano_fuse(x, w, bias, lens):
    return PanoFuseFn.apply(x, w, bias, lens)


class RowDotFn(torch.autograd.Function):
    """y[m] = x[m,:] . w + b  (fp32 output)."""

    @staticmethod
    def forward(ctx, x, w, bias):
        h = x.shape[-1]
        x2 = x.reshape(-1, h).contiguous()
        y = torch.empty(x2.shape[0], dtype=torch.float32, device=x.device)
        if w.numel() != h:
            raise L.MagicError("rowdot: weight must have h elements")
        call("magic_rowdot_fwd", ptr(x2), ptr(w), ptr(bias), ptr(y), x2.shape[0], h, dt(x2), stream())
        ctx.save_for_backward(x2, w, bias)
        ctx.xshape = x.shape
        return y.view(x.shape[:-1])

    @staticmethod
    def backward(ctx, dy):
        x2, w, bias = ctx.saved_tensors
        h = x2.shape[1]
        dy = dy.reshape(-1).contiguous().float()
        dx = torch.empty_like(x2)
        gw, rw, _ = _sink(w)
        gb, rb, _ = _sink(bias) if bias is not None else (None, None, 0)
        call("magic_rowdot_bwd", ptr(dy), ptr(x2), ptr(w), ptr(dx), ptr(gw), ptr(gb), x2.shape[0], h, dt(x2), stream())
        return dx.view(ctx.xshape), rw, rb


def rowdot(x, w, bias):
    return RowDotFn.apply(x, w, bias)


class AttentionFn(torch.autograd.Function):
    """Fused attention on packed projections.  `qsrc` [B*Lq, *] holds Q at column q_off; `kvsrc` [B*Lk, *]
    holds K at k_off and V at v_off (qsrc may be the same tensor as kvsrc: packed QKV)."""

    @staticmethod
    def forward(ctx, qsrc, kvsrc, q_off, k_off, v_off, B, H, Lq, Lk, key_lens, dists, sprel_w, sprel_b, need_pbar,
                drop_p, salt, key_skip=-1):
        same = kvsrc is None
        kv = qsrc if same else kvsrc
        hd = H * 64
        dev = qsrc.device
        out = torch.empty(B * Lq, hd, dtype=qsrc.dtype, device=dev)
        lse = torch.empty(B, H, Lq, dtype=torch.float32, device=dev)
        pbar = torch.empty(B, Lq, Lk, dtype=torch.float32, device=dev) if need_pbar else None
        esz = qsrc.element_size()
        seed = seed_tensor(dev) if drop_p > 0 else None
        scale = 1.0 / math.sqrt(64.0)
        sw = sprel_w.reshape(-1) if sprel_w is not None else None
        sb = sprel_b.reshape(-1) if sprel_b is not None else None
        with _key_skip(key_skip):
            call("magic_attn_fwd", qsrc.data_ptr() + q_off * esz, kv.data_ptr() + k_off * esz,
                 kv.data_ptr() + v_off * esz, qsrc.stride(0), kv.stride(0), kv.stride(0), ptr(out), ptr(lse), ptr(pbar),
                 Lq * Lk, Lk, B, H, Lq, Lk, ptr(key_lens), ptr(dists), ptr(sw), ptr(sb), scale, dt(qsrc), drop_p, salt,
                 ptr(seed), stream())
        ctx.key_skip = key_skip
        ctx.save_for_backward(qsrc, kvsrc, key_lens, dists, sprel_w, sprel_b, lse, out)
        ctx.meta = (q_off, k_off, v_off, B, H, Lq, Lk, drop_p, salt, scale, need_pbar)
        if need_pbar:
            return out, pbar
        return out, None

    @staticmethod
    def backward(ctx, dout, dpbar):
        qsrc, kvsrc, key_lens, dists, sprel_w, sprel_b, lse, out = ctx.saved_tensors
        q_off, k_off, v_off, B, H, Lq, Lk, drop_p, salt, scale, need_pbar = ctx.meta
        same = kvsrc is None
        kv = qsrc if same else kvsrc
        dev = qsrc.device
        esz = qsrc.element_size()
        dout = dout.contiguous()
        if dpbar is not None:
            dpbar = dpbar.contiguous().float()
        delta = torch.empty(B, H, Lq, dtype=torch.float32, device=dev)
        # every column of the packed buffers is written by the kernel (q | k | v blocks)
        dqsrc = torch.empty_like(qsrc)
        dkv = dqsrc if same else torch.empty_like(kvsrc)
        gs = rs_w = rs_b = None
        if dists is not None:
            gs = torch.zeros(2, dtype=torch.float32, device=dev)
        seed = seed_tensor(dev) if drop_p > 0 else None
        sw = sprel_w.reshape(-1) if sprel_w is not None else None
        sb = sprel_b.reshape(-1) if sprel_b is not None else None
        done = False
        if ctx.key_skip >= 0:
            L.load().magic_attn_set_key_skip(int(ctx.key_skip))  # (reset after the launches below)
        if (dpbar is None and qsrc.dtype == torch.bfloat16 and _BRANCH["on"] and (L._PROFILE is None or L._PROFILE_GRAPH)
                and B * H * Lq <= 24576):  # latency-bound sizes only: at H = 12 the halves fill the GPU on their own
            # no KD-map gradient: the key-major kernel (dK, dV) recomputes delta = dO . O itself, so it does not depend
            # on the query-major kernel (dQ) and the two run concurrently -- a forked branch of the captured graph
            main = torch.cuda.current_stream()
            side = _attn_stream(main)
            side.wait_stream(main)

            def part(which):
                return L.call_rc("magic_attn_bwd_part", qsrc.data_ptr() + q_off * esz, kv.data_ptr() + k_off * esz,
                                 kv.data_ptr() + v_off * esz, qsrc.stride(0), kv.stride(0), kv.stride(0), ptr(dout),
                                 ptr(out), ptr(lse), ptr(delta), dqsrc.data_ptr() + q_off * esz,
                                 dkv.data_ptr() + k_off * esz, dkv.data_ptr() + v_off * esz, dqsrc.stride(0),
                                 dkv.stride(0), dkv.stride(0), ptr(gs), B, H, Lq, Lk, ptr(key_lens), ptr(dists),
                                 ptr(sw), ptr(sb), scale, dt(qsrc), drop_p, salt, ptr(seed), which, stream())

            with torch.cuda.stream(side):
                rc = part(2)
            if rc == 0:
                if part(1) != 0:
                    raise L.MagicError("magic_attn_bwd_part: the query-major half declined a shape the key-major took")
                main.wait_stream(side)
                for t in (qsrc, kv, dout, out, lse, dqsrc, dkv, key_lens, dists, sw, sb, seed):
                    if t is not None:
                        t.record_stream(side)
                done = True
        if not done:
            call("magic_attn_bwd", qsrc.data_ptr() + q_off * esz, kv.data_ptr() + k_off * esz,
                 kv.data_ptr() + v_off * esz, qsrc.stride(0), kv.stride(0), kv.stride(0), ptr(dout), ptr(lse), ptr(dpbar),
                 Lq * Lk, Lk, ptr(delta), dqsrc.data_ptr() + q_off * esz, dkv.data_ptr() + k_off * esz,
                 dkv.data_ptr() + v_off * esz, dqsrc.stride(0), dkv.stride(0), dkv.stride(0), ptr(gs), B, H, Lq, Lk,
                 ptr(key_lens), ptr(dists), ptr(sw), ptr(sb), scale, dt(qsrc), drop_p, salt, ptr(seed), stream())
        if ctx.key_skip >= 0:
            L.load().magic_attn_set_key_skip(-1)
        if dists is not None:
            gw = getattr(sprel_w, "_magic_grad", None)
            if gw is not None:
                gw.view(-1).add_(gs[0:1])
                sprel_b._magic_grad.view(-1).add_(gs[1:2])
            else:
                rs_w = gs[0:1].view(sprel_w.shape)
                rs_b = gs[1:2].view(sprel_b.shape)
        return (dqsrc, None if same else dkv, None, None, None, None, None, None, None, None, None, rs_w, rs_b, None,
                None, None, None)


class _key_skip:
    """Scope of `magic_attn_set_key_skip`: the launches inside never attend key `idx` (-1: no-op)."""

    def __init__(self, idx):
        self.idx = int(idx)

    def __enter__(self):
        if self.idx >= 0:
            L.load().magic_attn_set_key_skip(self.idx)

    def __exit__(self, *a):
        if self.idx >= 0:
            L.load().magic_attn_set_key_skip(-1)
        return False


def attention(qsrc, kvsrc, q_off, k_off, v_off, B, H, Lq, Lk, key_lens=None, dists=None, sprel_w=None, sprel_b=None,
              need_pbar=False, drop_p=0.0, salt=0, key_skip=-1):
    """`key_skip`: index of one key inside the valid prefix that is never attended (the navigation graph's [MEM]
    slot, agent.py:228); -1 = none."""
    return AttentionFn.apply(qsrc, kvsrc, q_off, k_off, v_off, B, H, Lq, Lk, key_lens, dists, sprel_w, sprel_b,
                             need_pbar, drop_p if _DROP_ON else 0.0, salt, key_skip)


class GmapAggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, fused, gi):
        h = tokens.shape[-1]
        t2 = tokens.reshape(-1, h).contiguous()
        f2 = fused.contiguous()
        n_nodes = gi["n_nodes"]
        out = torch.empty(n_nodes, h, dtype=tokens.dtype, device=tokens.device)
        call("magic_gmap_aggregate_fwd", ptr(t2), ptr(f2), ptr(gi["node_ptr"]), ptr(gi["entries"]), ptr(out), n_nodes,
             h, dt(t2), stream())
        ctx.gi = gi
        ctx.shapes = (tokens.shape, fused.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        gi = ctx.gi
        tshape, fshape = ctx.shapes
        dout = dout.contiguous()
        h = dout.shape[-1]
        dt_ = torch.empty(tshape, dtype=dout.dtype, device=dout.device)
        df = torch.empty(fshape, dtype=dout.dtype, device=dout.device)
        call("magic_gmap_aggregate_bwd", ptr(dout), ptr(gi["src_ids"]), ptr(gi["src_ptr"]), ptr(gi["src_nodes"]),
             ptr(gi["src_w"]), ptr(dt_), dt_.numel() // h, ptr(df), df.numel() // h, gi["n_src"], h, dt(dout), stream())
        return dt_, df, None


def gmap_aggregate(tokens, fused, gi):
    return GmapAggFn.apply(tokens, fused, gi)


class SapFuseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g_raw, l_raw, gate_raw, si):
        B, G = g_raw.shape
        Vp = l_raw.shape[1]
        g_raw, l_raw = g_raw.contiguous().float(), l_raw.contiguous().float()
        gate = gate_raw.contiguous().float() if gate_raw is not None else None
        gl, ll, fl = torch.empty_like(g_raw), torch.empty_like(l_raw), torch.empty_like(g_raw)
        call("magic_sap_fuse_fwd", ptr(g_raw), ptr(l_raw), ptr(gate), ptr(si["g_valid"]), ptr(si["l_valid"]),
             ptr(si["node2cand"]), ptr(si["bw_mask"]), ptr(gl), ptr(ll), ptr(fl), B, G, Vp, stream())
        ctx.save_for_backward(g_raw, l_raw, gate)
        ctx.si = si
        return gl, ll, fl

    @staticmethod
    def backward(ctx, dgl, dll, dfl):
        g_raw, l_raw, gate = ctx.saved_tensors
        si = ctx.si
        B, G = g_raw.shape
        Vp = l_raw.shape[1]
        z = lambda d, ref: torch.zeros_like(ref) if d is None else d.contiguous().float()
        dgl, dll, dfl = z(dgl, g_raw), z(dll, l_raw), z(dfl, g_raw)
        dg, dl = torch.empty_like(g_raw), torch.empty_like(l_raw)
        dgate = torch.empty(B, dtype=torch.float32, device=g_raw.device) if gate is not None else None
        call("magic_sap_fuse_bwd", ptr(dgl), ptr(dll), ptr(dfl), ptr(g_raw), ptr(l_raw), ptr(gate), ptr(si["g_valid"]),
             ptr(si["l_valid"]), ptr(si["node2cand"]), ptr(si["bw_mask"]), ptr(dg), ptr(dl), ptr(dgate), B, G, Vp,
             stream())
        return dg, dl, dgate, None


def sap_fuse(g_raw, l_raw, gate_raw, si):
    return SapFuseFn.apply(g_raw, l_raw, gate_raw, si)


class CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        R, C = logits.shape
        logits = _rows2d(logits)
        labels = labels.contiguous()
        loss = torch.empty(R, dtype=torch.float32, device=logits.device)
        lse = torch.empty(R, dtype=torch.float32, device=logits.device)
        call("magic_ce_fwd", ptr(logits), ptr(labels), ptr(loss), ptr(lse), R, C, logits.stride(0), ignore_index,
             dt(logits), stream())
        ctx.save_for_backward(logits, labels, lse)
        ctx.ignore_index = ignore_index
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits, labels, lse = ctx.saved_tensors
        R, C = logits.shape
        dloss = dloss.contiguous().float()
        dlogits = torch.empty_strided(logits.shape, logits.stride(), dtype=logits.dtype, device=logits.device)
        call("magic_ce_bwd", ptr(logits), ptr(labels), ptr(lse), ptr(dloss), ptr(dlogits), R, C, logits.stride(0),
             ctx.ignore_index, dt(logits), stream())
        return dlogits, None, None


def cross_entropy(logits, labels, ignore_index=-100):
    return CrossEntropyFn.apply(logits, labels, ignore_index)


# ---------------------------------------------------------------------------------------------------
# MAKD fused losses
# ---------------------------------------------------------------------------------------------------
def _mk_segs(pairs, with_ds):
    arr = (L.MagicMseSeg * len(pairs))()
    for i, p in enumerate(pairs):
        s, t = p["s"], p["t"]
        arr[i].s, arr[i].t = s.data_ptr(), t.data_ptr()
        arr[i].ds = p["ds"].data_ptr() if with_ds else None
        arr[i].w = p["w"].data_ptr() if p["w"] is not None else None
        arr[i].scale_dev = p["sdev"].data_ptr() if p.get("sdev") is not None else None
        arr[i].rows, arr[i].inner = p["rows"], p["inner"]
        arr[i].s_rs, arr[i].t_rs = p["s_rs"], p["t_rs"]
        arr[i].scale = p["scale"]
        arr[i].s_dt, arr[i].t_dt = dt(s), dt(t)
    return arr


class MakdMseFn(torch.autograd.Function):
    """One launch for all MSE segments.  Returns (per-segment losses [n], total scalar).  `aw_dev` (optional device
    tensor [5]) holds the ability multipliers the segments read through `scale_dev`; when it requires grad (learned
    ability weights, agent.py:585) backward also returns d total / d aw_dev."""

    @staticmethod
    def forward(ctx, meta, aw_dev, *tensors):
        n = len(meta)
        ss, ts, ws = tensors[:n], tensors[n:2 * n], tensors[2 * n:3 * n]
        pairs = []
        for i in range(n):
            s, t = ss[i].contiguous(), ts[i].contiguous()
            rows = s.shape[0]
            inner = s.numel() // max(rows, 1)
            w = ws[i]
            if w is not None:
                w = w.contiguous().float()
            pairs.append(dict(s=s, t=t, w=w, rows=rows, inner=inner, s_rs=inner, t_rs=inner, scale=meta[i]["scale"],
                              sdev=meta[i].get("sdev")))
        loss = torch.empty(L.MAKD_MAX_SEGS + 1, dtype=torch.float32, device=ss[0].device)
        segs = _mk_segs(pairs, False)
        call("magic_makd_mse_fwd", segs, n, ptr(loss), stream())
        ctx.pairs = pairs
        ctx.ability = [m.get("ability") for m in meta]
        ctx.aw = (aw_dev.detach(), loss) if (aw_dev is not None and aw_dev.requires_grad) else None
        return loss[:n], loss[L.MAKD_MAX_SEGS]

    @staticmethod
    def backward(ctx, dseg, dtot):
        pairs = ctx.pairs
        n = len(pairs)
        dseg = dseg.contiguous().float() if dseg is not None else None
        dtot = dtot.reshape(1).contiguous().float() if dtot is not None else None
        outs = []
        for p in pairs:
            p["ds"] = torch.empty_like(p["s"])
            outs.append(p["ds"])
        segs = _mk_segs(pairs, True)
        call("magic_makd_mse_bwd", segs, n, ptr(dseg), ptr(dtot), stream())
        g_aw = None
        if ctx.aw is not None:  # learned ability weights: d/d aw[a] = sum over the ability's segments of (raw loss)
            aw, loss = ctx.aw
            idx = torch.tensor(ctx.ability, dtype=torch.int64, device=aw.device)
            up = (dseg if dseg is not None else 0) + (dtot if dtot is not None else 0)
            g_aw = torch.zeros_like(aw).index_add_(0, idx, up * loss[:n] / aw[idx])
        ctx.pairs = None
        return (None, g_aw, *outs, *([None] * (2 * n)))


def makd_mse(pairs, aw_dev=None, ability=None):
    """pairs: list of (s, t, w_or_None, scale[, scale_dev]). `scale_dev` is an optional 1-element DEVICE tensor
    multiplied into the host `scale` (device-resident ability weight; a slice of `aw_dev`, `ability[i]` = its index).
    Returns (per-segment losses [n], their sum)."""
    meta = [dict(scale=float(p[3]), sdev=(p[4] if len(p) > 4 else None),
                 ability=(ability[i] if ability is not None else None)) for i, p in enumerate(pairs)]
    return MakdMseFn.apply(meta, aw_dev, *[p[0] for p in pairs], *[p[1] for p in pairs], *[p[2] for p in pairs])


class LossMixFn(torch.autograd.Function):
    """total = alpha*(mse_total + kl) + (1-alpha)*mean(sup).  Returns [total, sup_mean, kd] (grad flows via [0])."""

    @staticmethod
    def forward(ctx, mse_total, kl, sup, alpha, inv_n):
        sup = sup.contiguous().float()
        out = torch.empty(3, dtype=torch.float32, device=sup.device)
        call("magic_loss_mix_fwd", ptr(mse_total), ptr(kl), ptr(sup), sup.numel(), alpha, ptr(inv_n), ptr(out),
             stream())
        ctx.meta = (mse_total is not None, kl is not None, sup.numel(), alpha)
        ctx.inv_n = inv_n
        return out

    @staticmethod
    def backward(ctx, dout):
        has_mse, has_kl, n, alpha = ctx.meta
        g = dout[0:1].contiguous()
        d_mse = torch.empty((), dtype=torch.float32, device=dout.device) if has_mse else None
        d_kl = torch.empty((), dtype=torch.float32, device=dout.device) if has_kl else None
        d_sup = torch.empty(n, dtype=torch.float32, device=dout.device)
        call("magic_loss_mix_bwd", ptr(g), n, alpha, ptr(ctx.inv_n), ptr(d_mse), ptr(d_kl), ptr(d_sup), stream())
        return d_mse, d_kl, d_sup, None, None


def loss_mix(mse_total, kl, sup, alpha, inv_n=None):
    return LossMixFn.apply(mse_total, kl, sup, float(alpha), inv_n)


class MakdKlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, t, temperature, w, scale, sdev, aw_dev):
        R, C = s.shape
        s, t = _rows2d(s), _rows2d(t)
        if t.dtype != s.dtype:
            t = t.to(s.dtype)
        if t.stride(0) != s.stride(0):
            s, t = s.contiguous(), t.contiguous()
        if w is not None:
            w = w.contiguous().float()
        stats = torch.empty(R, 2, dtype=torch.float32, device=s.device)
        loss = torch.empty(1, dtype=torch.float32, device=s.device)
        call("magic_makd_kl_fwd", ptr(s), ptr(t), R, C, s.stride(0), temperature, ptr(w), scale, ptr(sdev), ptr(stats),
             ptr(loss), dt(s), stream())
        ctx.save_for_backward(s, t, w, stats, sdev)
        ctx.meta = (temperature, scale)
        ctx.aw = (aw_dev.detach(), loss) if (aw_dev is not None and aw_dev.requires_grad) else None
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        s, t, w, stats, sdev = ctx.saved_tensors
        temperature, scale = ctx.meta
        R, C = s.shape
        g = dloss.reshape(1).contiguous().float()
        ds = torch.empty_strided(s.shape, s.stride(), dtype=s.dtype, device=s.device)
        call("magic_makd_kl_bwd", ptr(s), ptr(t), ptr(ds), R, C, s.stride(0), temperature, ptr(w), scale, ptr(sdev),
             ptr(stats), ptr(g), dt(s), stream())
        g_aw = None
        if ctx.aw is not None:  # the predict ability is index 4 of the ability weights
            aw, loss = ctx.aw
            g_aw = torch.zeros_like(aw)
            g_aw[4:5] = g * loss / aw[4:5]
        return ds, None, None, None, None, None, g_aw


def makd_kl(s, t, temperature, w, scale, scale_dev=None, aw_dev=None):
    return MakdKlFn.apply(s, t, float(temperature), w, float(scale), scale_dev, aw_dev)


# ---------------------------------------------------------------------------------------------------
# glue
# ---------------------------------------------------------------------------------------------------
class AddFn(torch.autograd.Function):
    """out = scale * (a + b (+ c)), elementwise, same shape/dtype."""

    @staticmethod
    def forward(ctx, a, b, c, scale):
        a, b = a.contiguous(), b.contiguous()
        c = c.contiguous() if c is not None else None
        out = torch.empty_like(a)
        call("magic_add", ptr(a), ptr(b), ptr(c), ptr(out), a.numel(), scale, dt(a), stream())
        ctx.has_c, ctx.scale = c is not None, scale
        return out

    @staticmethod
    def backward(ctx, d):
        if ctx.scale != 1.0:
            d = d * ctx.scale
        return d, d, (d if ctx.has_c else None), None


def add(a, b, c=None, scale=1.0):
    return AddFn.apply(a, b, c, float(scale))


class MatmulNTFn(torch.autograd.Function):
    """C[M,N] = alpha * A[M,K] B[N,K]^T for two ACTIVATIONS (contrastive similarity matrices)."""

    @staticmethod
    def forward(ctx, a, b, alpha):
        a, b = a.contiguous(), b.contiguous()
        M, K = a.shape
        N = b.shape[0]
        ldc = (N + 7) // 8 * 8  # 16-byte rows so the gradient is a legal tensor-core operand
        out = torch.empty(M, ldc, dtype=a.dtype, device=a.device)[:, :N]
        gemm(a, K, 1, b, 1, K, out, M, N, K, alpha=alpha)
        ctx.save_for_backward(a, b)
        ctx.alpha = alpha
        return out

    @staticmethod
    def backward(ctx, dc):
        a, b = ctx.saved_tensors
        M, K = a.shape
        N = b.shape[0]
        dc = _rows2d(dc)
        da = torch.empty_like(a)
        db = torch.empty_like(b)
        gemm(dc, dc.stride(0), 1, b, K, 1, da, M, K, N, alpha=ctx.alpha)
        gemm(dc, 1, dc.stride(0), a, K, 1, db, N, K, M, alpha=ctx.alpha)
        return da, db, None


def matmul_nt(a, b, alpha=1.0):
    return MatmulNTFn.apply(a, b, float(alpha))


class L2NormFn(torch.autograd.Function):
    """y = x / max(||x||_2, eps) per row (F.normalize)."""

    @staticmethod
    def forward(ctx, x, eps):
        x = x.contiguous()
        R, h = x.shape
        y = torch.empty_like(x)
        inv = torch.empty(R, dtype=torch.float32, device=x.device)
        call("magic_l2norm_fwd", ptr(x), ptr(y), ptr(inv), R, h, eps, dt(x), stream())
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        call("magic_l2norm_bwd", ptr(dy), ptr(y), ptr(inv), ptr(dx), y.shape[0], y.shape[1], dt(y), stream())
        return dx, None


def l2norm(x, eps=1e-12):
    return L2NormFn.apply(x, float(eps))


class SoftCrossEntropyFn(torch.autograd.Function):
    """loss[r] = KL(targets[r] || softmax(logits[r])), targets fp32 soft labels (MRC)."""

    @staticmethod
    def forward(ctx, logits, targets):
        logits = _rows2d(logits)
        targets = targets.contiguous().float()
        R, C = logits.shape
        loss = torch.empty(R, dtype=torch.float32, device=logits.device)
        stats = torch.empty(R, 2, dtype=torch.float32, device=logits.device)
        call("magic_soft_ce_fwd", ptr(logits), ptr(targets), ptr(loss), ptr(stats), R, C, logits.stride(0),
             targets.stride(0), dt(logits), stream())
        ctx.save_for_backward(logits, targets, stats)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits, targets, stats = ctx.saved_tensors
        R, C = logits.shape
        dloss = dloss.contiguous().float()
        dlogits = torch.empty_strided(logits.shape, logits.stride(), dtype=logits.dtype, device=logits.device)
        call("magic_soft_ce_bwd", ptr(logits), ptr(targets), ptr(stats), ptr(dloss), ptr(dlogits), R, C,
             logits.stride(0), targets.stride(0), dt(logits), stream())
        return dlogits, None


def soft_cross_entropy(logits, targets):
    return SoftCrossEntropyFn.apply(logits, targets)


def zero_rows_(x2d, rows):
    """In place: x2d[rows[i], :] = 0 (rows < 0 are skipped).  No autograd (network inputs)."""
    if rows.numel():
        call("magic_zero_rows", ptr(x2d), ptr(rows), rows.numel(), x2d.shape[1], dt(x2d), stream())
    return x2d


class Cat2Fn(torch.autograd.Function):
    """[a | b] along the last dim for 2-D inputs."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        R, ca, cb = a.shape[0], a.shape[1], b.shape[1]
        out = torch.empty(R, ca + cb, dtype=a.dtype, device=a.device)
        esz = a.element_size()
        call("magic_copy2d", ptr(a), ca, out.data_ptr(), ca + cb, R, ca, dt(a), stream())
        call("magic_copy2d", ptr(b), cb, out.data_ptr() + ca * esz, ca + cb, R, cb, dt(a), stream())
        ctx.dims = (R, ca, cb)
        return out

    @staticmethod
    def backward(ctx, d):
        R, ca, cb = ctx.dims
        d = d.contiguous()
        da = torch.empty(R, ca, dtype=d.dtype, device=d.device)
        db = torch.empty(R, cb, dtype=d.dtype, device=d.device)
        esz = d.element_size()
        call("magic_copy2d", d.data_ptr(), ca + cb, ptr(da), ca, R, ca, dt(d), stream())
        call("magic_copy2d", d.data_ptr() + ca * esz, ca + cb, ptr(db), cb, R, cb, dt(d), stream())
        return da, db


def cat2(a, b):
    return Cat2Fn.apply(a, b)


class MaskFillFn(torch.autograd.Function):
    """out = valid ? x : fill (fp32); gradient passes where valid."""

    @staticmethod
    def forward(ctx, x, valid, fill):
        x = x.contiguous().float()
        valid = valid.contiguous()
        out = torch.empty_like(x)
        call("magic_mask_fill", ptr(x), ptr(valid), ptr(out), x.numel(), fill, stream())
        ctx.save_for_backward(valid)
        return out

    @staticmethod
    def backward(ctx, dy):
        (valid,) = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(dy)
        call("magic_mask_fill", ptr(dy), ptr(valid), ptr(dx), dy.numel(), 0.0, stream())
        return dx, None, None


def mask_fill(x, valid, fill=float("-inf")):
    return MaskFillFn.apply(x, valid, float(fill))


def row_weights(src, idx, scale, n):
    """out[i] = (src[idx[i]] or src[i] or 1) * (scale[i] or 1), fp32 [n]; idx < 0 gives 0.  No autograd (KD weights)."""
    dev = (src if src is not None else scale).device
    out = torch.empty(n, dtype=torch.float32, device=dev)
    call("magic_row_weights", ptr(src), ptr(idx), ptr(scale), ptr(out), n, stream())
    return out


def segment_mean(vals, seg, inv_count, n_seg):
    """out[s] = inv_count[s] * sum_{i: seg[i]==s} vals[i]  (no autograd; used for MKTD weights / logging)."""
    out = torch.empty(n_seg, dtype=torch.float32, device=vals.device)
    call("magic_segsum", ptr(vals.contiguous().float()), ptr(seg), ptr(inv_count), ptr(out), vals.numel(), n_seg,
         stream())
    return out
