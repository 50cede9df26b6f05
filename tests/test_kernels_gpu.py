"""GPU parity tests of every C-ABI kernel against plain fp32 PyTorch math (and the oracle's KD functions).
Tolerances: fp32 mode 1e-4 relative (north_star), bf16 mode 2e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import magic_b200  # noqa: E402
from magic_b200 import ops, _lib as L  # noqa: E402
from magic_b200 import kd_loss as KD  # noqa: E402
from oracle import kd_loss_oracle as KO  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def close(a, b, tol, what=""):
    r = rel(a, b)
    assert r <= tol, f"{what}: rel err {r:.3e} > {tol}"


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("M,N,K", [(200, 96, 72), (64, 128, 128), (37, 50, 768), (300, 384, 128)])
def test_linear_fwd_bwd(dtype, tol, M, N, K):
    torch.manual_seed(0)
    x = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) * 0.1
    b = torch.randn(N, device=DEV) * 0.1
    res = torch.randn(M, N, device=DEV)
    for act, fn in [(0, lambda v: v), (1, F.gelu), (2, F.relu)]:
        xr, wr, br, rr = [t.clone().requires_grad_() for t in (x, w, b, res)]
        if dtype == torch.bfloat16:
            xin, rin = xr.to(dtype), rr.to(dtype)
        else:
            xin, rin = xr, rr
        y = ops.linear(xin, wr, br, act=act, residual=rin)
        ref_x = x.to(dtype).float() if dtype == torch.bfloat16 else x
        x2, w2, b2, r2 = [t.clone().requires_grad_() for t in (ref_x, w, b, res)]
        yr = fn(F.linear(x2, w2, b2)) + (r2.to(dtype).float() if dtype == torch.bfloat16 else r2)
        close(y, yr, tol, f"linear fwd act={act}")
        g = torch.randn_like(yr)
        y.backward(g.to(dtype))
        yr.backward(g.to(dtype).float())
        close(xr.grad, x2.grad, tol * 2, f"linear dx act={act}")
        close(wr.grad, w2.grad, tol * 2, f"linear dw act={act}")
        close(br.grad, b2.grad, tol * 2, f"linear db act={act}")
        close(rr.grad, r2.grad, tol * 2, f"linear dres act={act}")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_ffn_and_packed(dtype, tol):
    torch.manual_seed(1)
    M, h, I = 150, 128, 512
    x = torch.randn(M, h, device=DEV)
    w1, b1 = torch.randn(I, h, device=DEV) * 0.05, torch.randn(I, device=DEV) * 0.05
    w2, b2 = torch.randn(h, I, device=DEV) * 0.05, torch.randn(h, device=DEV) * 0.05
    ps = [t.clone().requires_grad_() for t in (x, w1, b1, w2, b2)]
    rs = [t.clone().requires_grad_() for t in (x.to(dtype).float(), w1, b1, w2, b2)]
    y = ops.ffn(ps[0].to(dtype), ps[1], ps[2], ps[3], ps[4])
    yr = F.linear(F.gelu(F.linear(rs[0], rs[1], rs[2])), rs[3], rs[4])
    close(y, yr, tol, "ffn fwd")
    g = torch.randn_like(yr)
    y.backward(g.to(dtype))
    yr.backward(g.to(dtype).float())
    for a, b, n in zip(ps, rs, ["dx", "dw1", "db1", "dw2", "db2"]):
        close(a.grad, b.grad, tol * 3, "ffn " + n)
    # packed qkv
    ws = [torch.randn(h, h, device=DEV) * 0.05 for _ in range(3)]
    bs = [torch.randn(h, device=DEV) * 0.05 for _ in range(3)]
    pw = [t.clone().requires_grad_() for t in ws]
    pb = [t.clone().requires_grad_() for t in bs]
    px = x.clone().requires_grad_()
    out = ops.packed_linear(px.to(dtype), pw, pb)
    rw = [t.clone().requires_grad_() for t in ws]
    rb = [t.clone().requires_grad_() for t in bs]
    rx = x.to(dtype).float().clone().requires_grad_()
    ref = torch.cat([F.linear(rx, a, b) for a, b in zip(rw, rb)], 1)
    close(out, ref, tol, "packed fwd")
    g = torch.randn_like(ref)
    out.backward(g.to(dtype))
    ref.backward(g.to(dtype).float())
    close(px.grad, rx.grad, tol * 3, "packed dx")
    for a, b in zip(pw + pb, rw + rb):
        close(a.grad, b.grad, tol * 3, "packed dparam")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("h", [128, 200, 256, 384, 768])  # 200: scalar kernel; the others: vectorised kernel
def test_layer_norm(dtype, tol, h):
    torch.manual_seed(2)
    M = 333
    x, r = torch.randn(M, h, device=DEV), torch.randn(M, h, device=DEV)
    g, b = torch.rand(h, device=DEV) + 0.5, torch.randn(h, device=DEV)
    xs = [t.clone().requires_grad_() for t in (x, r, g, b)]
    rs = [t.clone().requires_grad_() for t in (x.to(dtype).float(), r.to(dtype).float(), g, b)]
    y = ops.layer_norm(xs[0].to(dtype), xs[2], xs[3], 1e-12, res=xs[1].to(dtype))
    yr = F.layer_norm(rs[0] + rs[1], (h,), rs[2], rs[3], 1e-12)
    close(y, yr, tol, "ln fwd")
    go = torch.randn_like(yr)
    y.backward(go.to(dtype))
    yr.backward(go.to(dtype).float())
    for a, c, n in zip(xs, rs, ["dx", "dres", "dgamma", "dbeta"]):
        close(a.grad, c.grad, tol * 3, "ln " + n)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_embed_ln_posfuse_gather(dtype, tol):
    torch.manual_seed(3)
    B, Lx, h, vocab = 5, 23, 128, 1000
    ids = torch.randint(0, vocab, (B, Lx), device=DEV)
    word, pos, typ = [torch.randn(n, h, device=DEV) * 0.1 for n in (vocab, 64, 1)]
    g, b = torch.rand(h, device=DEV) + 0.5, torch.randn(h, device=DEV)
    ps = [t.clone().requires_grad_() for t in (word, pos, typ, g, b)]
    rs = [t.clone().requires_grad_() for t in (word, pos, typ, g, b)]
    y = ops.embed_ln(ids, ps[0], ps[1], ps[2], ps[3], ps[4], 1e-12, dtype)
    e = rs[0][ids] + rs[1][torch.arange(Lx, device=DEV)][None] + rs[2][0]
    yr = F.layer_norm(e, (h,), rs[3], rs[4], 1e-12)
    close(y, yr, tol, "embed fwd")
    go = torch.randn_like(yr)
    y.backward(go.to(dtype))
    yr.backward(go.to(dtype).float())
    for a, c, n in zip(ps, rs, ["dword", "dpos", "dtype", "dgamma", "dbeta"]):
        close(a.grad, c.grad, tol * 3, "embed " + n)
    # posfuse
    M, K = 77, 7
    xin = torch.randn(M, h, device=DEV)
    idx = torch.randint(0, 3, (M,), device=DEV)
    emb, cst = torch.randn(3, h, device=DEV), torch.randn(1, h, device=DEV)
    f = torch.randn(M, K, device=DEV)
    W, bb = torch.randn(h, K, device=DEV), torch.randn(h, device=DEV)
    ps = [t.clone().requires_grad_() for t in (xin, emb, cst, W, bb, g, b)]
    rs = [t.clone().requires_grad_() for t in (xin.to(dtype).float(), emb, cst, W, bb, g, b)]
    y = ops.posfuse(ps[0].to(dtype), idx, ps[1], ps[2], f, ps[3], ps[4], ps[5], ps[6], 1e-12, dtype)
    yr = rs[0] + rs[1][idx] + rs[2][0] + F.layer_norm(F.linear(f, rs[3], rs[4]), (h,), rs[5], rs[6], 1e-12)
    close(y, yr, tol, "posfuse fwd")
    go = torch.randn_like(yr)
    y.backward(go.to(dtype))
    yr.backward(go.to(dtype).float())
    for a, c, n in zip(ps, rs, ["dxin", "demb", "dcst", "dW", "db", "dgamma", "dbeta"]):
        close(a.grad, c.grad, tol * 3, "posfuse " + n)
    # gather / scatter
    src = torch.randn(40, h, device=DEV).to(dtype).requires_grad_()
    gi = torch.tensor([3, -1, 7, 39, 0, -1, 12], device=DEV)
    out = ops.gather_rows(src, gi)
    ref = torch.where((gi >= 0)[:, None], src.detach()[gi.clamp(min=0)], torch.zeros_like(out))
    assert torch.equal(out, ref)
    out.backward(torch.ones_like(out))
    exp = torch.zeros(40, h, device=DEV)
    exp[gi[gi >= 0]] = 1
    assert torch.equal(src.grad.float(), exp)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_pano_fuse_rowdot(dtype, tol):
    torch.manual_seed(4)
    R, V, h = 9, 36, 128
    x = torch.randn(R, V, h, device=DEV)
    w, b = torch.randn(1, h, device=DEV) * 0.2, torch.randn(1, device=DEV)
    lens = torch.tensor([36, 30, 36, 1, 12, 36, 36, 20, 36], device=DEV)
    ps = [t.clone().requires_grad_() for t in (x, w, b)]
    rs = [t.clone().requires_grad_() for t in (x.to(dtype).float(), w, b)]
    y = ops.pano_fuse(ps[0].to(dtype), ps[1], ps[2], lens)
    mask = torch.arange(V, device=DEV)[None] < lens[:, None]
    s = F.linear(rs[0], rs[1], rs[2]).squeeze(-1).masked_fill(~mask, float("-inf"))
    yr = (torch.softmax(s, -1)[..., None] * rs[0]).sum(1)
    close(y, yr, tol, "pano_fuse fwd")
    go = torch.randn_like(yr)
    y.backward(go.to(dtype))
    yr.backward(go.to(dtype).float())
    close(ps[0].grad, rs[0].grad, tol * 3, "pano_fuse dx")
    close(ps[1].grad, rs[1].grad, tol * 3, "pano_fuse dw")
    # masked-mean variant
    y2 = ops.pano_fuse(x.to(dtype), None, None, lens)
    ref2 = (x.to(dtype).float() * mask[..., None]).sum(1) / lens[:, None]
    close(y2, ref2, tol, "pano mean")
    # rowdot
    xx = torch.randn(50, h, device=DEV)
    wv, bv = torch.randn(h, device=DEV), torch.randn(1, device=DEV)
    ps = [t.clone().requires_grad_() for t in (xx, wv, bv)]
    rs = [t.clone().requires_grad_() for t in (xx.to(dtype).float(), wv, bv)]
    y = ops.rowdot(ps[0].to(dtype), ps[1], ps[2])
    yr = rs[0] @ rs[1] + rs[2]
    close(y, yr, tol, "rowdot fwd")
    go = torch.randn_like(yr)
    y.backward(go)
    yr.backward(go)
    for a, c, n in zip(ps, rs, ["dx", "dw", "db"]):
        close(a.grad, c.grad, tol * 3, "rowdot " + n)


def _ref_attn(q, k, v, lens, dists, sw, sb, H):
    B, Lq, hd = q.shape
    Lk = k.shape[1]
    d = hd // H
    qh, kh, vh = [t.view(B, -1, H, d).transpose(1, 2) for t in (q, k, v)]
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
    if dists is not None:
        s = s + (dists * sw + sb)[:, None]
    if lens is not None:
        m = torch.arange(Lk, device=q.device)[None] < lens[:, None]
        s = s + (~m).float()[:, None, None, :] * -10000.0
    p = torch.softmax(s, -1)
    o = (p @ vh).transpose(1, 2).reshape(B, Lq, hd)
    return o, p.mean(1)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("cfg", [dict(B=3, H=2, Lq=20, Lk=20, sprel=True), dict(B=4, H=2, Lq=37, Lk=80, sprel=False),
                                 dict(B=2, H=12, Lq=80, Lk=80, sprel=False), dict(B=2, H=2, Lq=160, Lk=50, sprel=False),
                                 dict(B=5, H=2, Lq=36, Lk=36, sprel=False), dict(B=2, H=2, Lq=50, Lk=50, sprel=True),
                                 dict(B=2, H=2, Lq=160, Lk=160, sprel=False), dict(B=2, H=4, Lq=80, Lk=37, sprel=False),
                                 dict(B=3, H=2, Lq=65, Lk=33, sprel=False), dict(B=1, H=2, Lq=200, Lk=200, sprel=False),
                                 dict(B=2, H=3, Lq=40, Lk=40, sprel=True), dict(B=3, H=6, Lq=37, Lk=80, sprel=False)])
def test_attention(dtype, tol, cfg):
    torch.manual_seed(5)
    B, H, Lq, Lk = cfg["B"], cfg["H"], cfg["Lq"], cfg["Lk"]
    hd = H * 64
    self_attn = Lq == Lk
    lens = torch.randint(max(1, Lk // 2), Lk + 1, (B,), device=DEV)
    lens[0] = Lk
    dists = torch.rand(B, Lq, Lk, device=DEV) * 30 if cfg["sprel"] else None
    sw = torch.tensor([[-0.05]], device=DEV, requires_grad=True)
    sb = torch.tensor([0.1], device=DEV, requires_grad=True)
    sw2, sb2 = sw.detach().clone().requires_grad_(), sb.detach().clone().requires_grad_()
    if self_attn:
        qkv = (torch.randn(B * Lq, 3 * hd, device=DEV)).to(dtype).requires_grad_()
        o, pbar = ops.attention(qkv, None, 0, hd, 2 * hd, B, H, Lq, Lk, lens.int(), dists, sw if dists is not None else None,
                                sb if dists is not None else None, need_pbar=True)
        ref_in = qkv.detach().float().requires_grad_()
        q, k, v = [ref_in[:, i * hd:(i + 1) * hd].reshape(B, Lq, hd) for i in range(3)]
    else:
        qt = torch.randn(B * Lq, hd, device=DEV).to(dtype).requires_grad_()
        kv = torch.randn(B * Lk, 2 * hd, device=DEV).to(dtype).requires_grad_()
        o, pbar = ops.attention(qt, kv, 0, 0, hd, B, H, Lq, Lk, lens.int(), need_pbar=True)
        rq, rkv = qt.detach().float().requires_grad_(), kv.detach().float().requires_grad_()
        q, k, v = rq.reshape(B, Lq, hd), rkv[:, :hd].reshape(B, Lk, hd), rkv[:, hd:].reshape(B, Lk, hd)
    oref, pref = _ref_attn(q, k, v, lens, dists, sw2 if dists is not None else None, sb2 if dists is not None else None, H)
    close(o.view(B, Lq, hd), oref, tol, "attn out")
    close(pbar, pref, max(tol, 1e-5), "attn pbar")
    go, gp = torch.randn_like(oref), torch.randn_like(pref)
    (o.view(B, Lq, hd).float() * go.to(dtype).float()).sum().backward(retain_graph=True)
    (pbar * gp).sum().backward()
    ((oref * go.to(dtype).float()).sum() + (pref * gp).sum()).backward()
    if self_attn:
        close(qkv.grad, ref_in.grad, tol * 4, "attn dqkv")
    else:
        close(qt.grad, rq.grad, tol * 4, "attn dq")
        close(kv.grad, rkv.grad, tol * 4, "attn dkv")
    if dists is not None:
        close(sw.grad, sw2.grad, tol * 10, "attn dsprel_w")
        # d(bias) is analytically zero (softmax is shift invariant): compare absolutely against the scale of dw
        assert (sb.grad - sb2.grad).abs().max().item() <= 1e-3 * max(1.0, sw2.grad.abs().max().item())


def test_ce_and_sap_fuse():
    torch.manual_seed(6)
    R, C = 40, 50265
    logits = (torch.randn(R, C, device=DEV) * 2).requires_grad_()
    labels = torch.randint(0, C, (R,), device=DEV)
    labels[3] = -1
    loss = ops.cross_entropy(logits, labels, -1)
    ref_l = logits.detach().clone().requires_grad_()
    ref = F.cross_entropy(ref_l, labels, reduction="none", ignore_index=-1)
    close(loss, ref, 1e-5, "ce fwd")
    w = torch.rand(R, device=DEV)
    (loss * w).sum().backward()
    (ref * w).sum().backward()
    close(logits.grad, ref_l.grad, 1e-4, "ce bwd")
    # -inf logits
    lg = torch.randn(6, 20, device=DEV)
    lg[:, 5:9] = float("-inf")
    lab = torch.tensor([0, 1, 2, 3, -100, 10], device=DEV)
    a = lg.clone().requires_grad_()
    b = lg.clone().requires_grad_()
    la = ops.cross_entropy(a, lab, -100)
    lb = F.cross_entropy(b, lab, reduction="none", ignore_index=-100)
    close(la, lb, 1e-5, "ce -inf fwd")
    la.sum().backward()
    lb.sum().backward()
    close(a.grad, b.grad, 1e-5, "ce -inf bwd")


def test_makd_mse_kl_vs_oracle():
    torch.manual_seed(7)
    B = 6
    for dtype, tol in [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)]:
        s1 = torch.randn(B, 80, 256, device=DEV).to(dtype).requires_grad_()
        t1 = torch.randn(B, 80, 256, device=DEV).to(dtype)
        s2 = torch.randn(17, 36, 256, device=DEV).to(dtype).requires_grad_()   # dim0 != B -> unweighted
        t2 = torch.randn(17, 36, 256, device=DEV).to(dtype)
        s3 = torch.rand(B, 33, 33, device=DEV).requires_grad_()                # fp32 attention map, ragged tail
        t3 = torch.rand(B, 33, 33, device=DEV)
        w = torch.rand(B, device=DEV)
        per, tot = ops.makd_mse([(s1, t1, w, 1.3 / s1.numel()), (s2, t2, None, 0.7 / s2.numel()),
                                 (s3, t3, w, 2.0 / s3.numel())])
        r1, r2, r3 = [x.detach().float().requires_grad_() for x in (s1, s2, s3)]
        e1 = KO.mse_loss(r1, t1.float(), w) * 1.3
        e2 = KO.mse_loss(r2, t2.float(), w) * 0.7   # silently unweighted (shape mismatch)
        e3 = KO.mse_loss(r3, t3, w) * 2.0
        close(per, torch.stack([e1, e2, e3]), tol, "makd mse fwd")
        close(tot, e1 + e2 + e3, tol, "makd mse total")
        (tot * 0.5 + per[1] * 2.0).backward()
        ((e1 + e2 + e3) * 0.5 + e2 * 2.0).backward()
        close(s1.grad, r1.grad, tol * 2 + (1e-2 if dtype == torch.bfloat16 else 0), "makd ds1")
        close(s2.grad, r2.grad, tol * 2 + (1e-2 if dtype == torch.bfloat16 else 0), "makd ds2")
        close(s3.grad, r3.grad, 1e-5, "makd ds3")
    # API-compatible functions
    a, b = torch.randn(B, 20, 64, device=DEV), torch.randn(B, 20, 64, device=DEV)
    w = torch.rand(B, device=DEV)
    close(KD.mse_loss(a, b), KO.mse_loss(a, b), 1e-5, "mse_loss")
    close(KD.mse_loss(a, b, w), KO.mse_loss(a, b, w), 1e-5, "mse_loss w")
    close(KD.mse_loss(a, b, w[:3]), KO.mse_loss(a, b, w[:3]), 1e-5, "mse_loss mismatched w")
    close(KD.mse_loss(a, b, w, loss_type="sum"), KO.mse_loss(a, b, w, loss_type="sum"), 1e-5, "mse_loss sum")
    with pytest.raises(ValueError):
        KD.mse_loss(a, b, w[:3], loss_type="sum")
    # KL with -inf masks
    sl = torch.randn(B, 20, device=DEV) * 3
    tl = torch.randn(B, 20, device=DEV) * 3
    sl[:, 7:11] = float("-inf")
    tl[:, 7:11] = float("-inf")
    tl[2, 15] = float("-inf")
    for ww in (None, w):
        for lt in (None, "sum"):
            x = sl.clone().requires_grad_()
            y = sl.clone().requires_grad_()
            got = KD.kd_loss(x, tl, temperature=2, t_sample_weights=ww, loss_type=lt)
            exp = KO.kd_loss(y, tl, temperature=2, t_sample_weights=ww, loss_type=lt)
            assert abs(got.item() - exp.item()) <= 1e-5 * abs(exp.item()) + 1e-6, (got.item(), exp.item())
            got.backward()
            exp.backward()
            assert torch.allclose(x.grad, y.grad, rtol=1e-4, atol=1e-6)
    close(KD.exponential_decay(w * 3, 0.7), KO.exponential_decay(w * 3, 0.7), 1e-6, "exp decay")
    close(KD.invert_normalized_losses(w), KO.invert_normalized_losses(w), 1e-6, "invert norm")
    # vocab-sized KL rows
    sl, tl = torch.randn(50, 50265, device=DEV), torch.randn(50, 50265, device=DEV)
    got = KD.kd_loss(sl, tl, temperature=2)
    exp = KO.kd_loss(sl, tl, temperature=2)
    assert abs(got.item() - exp.item()) <= 1e-4 * abs(exp.item()) + 1e-7


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 3e-2)])
def test_grad_link_matches_autograd_sum(dtype, tol):
    """y = LN(x + f(x)) with ops.GradLink (the residual-branch gradient rides the dgrad GEMM epilogue of f's first
    linear, LayerNorm reports None for the residual) gives the same gradients as letting autograd add the two
    branches -- for the FFN block, the packed-QKV block and a plain linear, and launches no torch add kernel."""
    torch.manual_seed(21)
    M, h, I = 300, 128, 512
    x0 = torch.randn(M, h, device=DEV)
    gamma, beta = torch.rand(h, device=DEV) + 0.5, torch.randn(h, device=DEV) * 0.1
    w1, b1 = torch.randn(I, h, device=DEV) * 0.05, torch.randn(I, device=DEV) * 0.05
    w2, b2 = torch.randn(h, I, device=DEV) * 0.05, torch.randn(h, device=DEV) * 0.05
    wq = [torch.randn(h, h, device=DEV) * 0.05 for _ in range(3)]
    bq = [torch.randn(h, device=DEV) * 0.05 for _ in range(3)]
    g = torch.randn(M, h, device=DEV).to(dtype)

    def run(use_link):
        x = x0.clone().requires_grad_()
        xs = x.to(dtype) * 1.0  # non-leaf sublayer input, as inside the model
        ps = [t.clone().requires_grad_() for t in (w1, b1, w2, b2, gamma, beta)]
        l1 = ops.GradLink() if use_link else None
        f = ops.ffn(xs, ps[0], ps[1], ps[2], ps[3], link=l1)
        y1 = ops.layer_norm(f, ps[4], ps[5], 1e-12, res=xs, link=l1)
        qs = [t.clone().requires_grad_() for t in wq + bq]
        l2 = ops.GradLink() if use_link else None
        qkv = ops.packed_linear(y1, qs[:3], qs[3:], link=l2)
        d = ops.linear(qkv[:, :h].contiguous() + qkv[:, h:2 * h] + qkv[:, 2 * h:], ps[2][:, :h].contiguous(), None)
        y2 = ops.layer_norm(d, ps[4], ps[5], 1e-12, res=y1, link=l2)
        l3 = ops.GradLink() if use_link else None
        q = ops.linear(y2, qs[0], qs[3], link=l3)
        y3 = ops.layer_norm(q, ps[4], ps[5], 1e-12, res=y2, link=l3)
        y3.backward(g)
        assert all(l is None or l.dres is None for l in (l1, l2, l3))  # every parked gradient was consumed
        return [x.grad] + [p.grad for p in ps] + [p.grad for p in qs]

    a, b = run(True), run(False)
    for i, (ga, gb) in enumerate(zip(a, b)):
        close(ga, gb, tol, f"grad_link tensor {i}")


def test_makd_edge_cases():
    """Empty segment, rows shorter than one 16-byte vector, many small rows that make one CTA's chunk range span
    segments, and vocabulary-sized bf16 KL rows with a padded leading dimension (single-pass forward)."""
    torch.manual_seed(11)
    B = 5
    w = torch.rand(B, device=DEV)
    s0, t0 = torch.zeros(0, 64, device=DEV, requires_grad=True), torch.zeros(0, 64, device=DEV)
    s1 = torch.randn(B, 3, device=DEV).requires_grad_()
    t1 = torch.randn(B, 3, device=DEV)
    s2 = torch.randn(3000, 40, device=DEV).bfloat16().requires_grad_()
    t2 = torch.randn(3000, 40, device=DEV).bfloat16()
    s3 = torch.randn(B, 8192 * 2 + 24, device=DEV).bfloat16().requires_grad_()
    t3 = torch.randn(B, 8192 * 2 + 24, device=DEV).bfloat16()
    per, tot = ops.makd_mse([(s0, t0, None, 1.0), (s1, t1, w, 1.0 / s1.numel()), (s2, t2, None, 1.0 / s2.numel()),
                             (s3, t3, w, 1.0 / s3.numel())])
    refs = [torch.zeros((), device=DEV)]
    leaves = []
    for s, t, ww in ((s1, t1, w), (s2, t2, None), (s3, t3, w)):
        r = s.detach().float().requires_grad_()
        leaves.append(r)
        refs.append(KO.mse_loss(r, t.float(), ww))
    close(per, torch.stack(refs), 1e-2, "makd edge fwd")
    assert abs(per[0].item()) == 0.0
    tot.backward()
    sum(refs[1:]).backward()
    close(s1.grad, leaves[0].grad, 1e-5, "makd edge ds1")
    close(s2.grad, leaves[1].grad, 2e-2, "makd edge ds2")
    close(s3.grad, leaves[2].grad, 2e-2, "makd edge ds3")
    # vocabulary-sized rows, bf16, leading dimension padded to 50272, masked columns, per-row weights
    R, C = 24, 50265
    sl = (torch.randn(R, 50272, device=DEV) * 2).bfloat16()[:, :C]
    tl = (torch.randn(R, 50272, device=DEV) * 2).bfloat16()[:, :C]
    sl[:, 100:200] = float("-inf")
    tl[:, 100:200] = float("-inf")
    wr = torch.rand(R, device=DEV)
    x = sl.clone().requires_grad_()  # clone() of a column slice is dense: ops re-pads it
    got = KD.kd_loss(x, tl, temperature=2, t_sample_weights=wr)
    y = sl.float().clone().requires_grad_()
    exp = KO.kd_loss(y, tl.float(), temperature=2, t_sample_weights=wr)
    assert abs(got.item() - exp.item()) <= 2e-3 * abs(exp.item()), (got.item(), exp.item())
    got.backward()
    exp.backward()
    close(x.grad, y.grad, 2e-2, "kd bf16 vocab grad")


def test_dropout_is_consistent_between_fwd_and_bwd():
    """The backward regenerates the forward mask: check d(out)/d(x) numerically along a random direction."""
    torch.manual_seed(8)
    ops.set_seed(DEV, 1234)
    M, h = 64, 128
    x = torch.randn(M, h, device=DEV, dtype=torch.float32, requires_grad=True)
    r = torch.randn(M, h, device=DEV)
    g, b = torch.ones(h, device=DEV), torch.zeros(h, device=DEV)

    def f(xx):
        return ops.layer_norm(xx, g, b, 1e-5, res=r, p_in=0.3, salt_in=11, p_out=0.2, salt_out=12)

    y = f(x)
    frac = (y == 0).float().mean().item()
    assert 0.1 < frac < 0.3, frac
    go = torch.randn_like(y)
    (y * go).sum().backward()
    d = torch.randn_like(x)
    eps = 1e-2
    num = ((f(x.detach() + eps * d) - f(x.detach() - eps * d)) * go).sum() / (2 * eps)
    ana = (x.grad * d).sum()
    assert abs(num.item() - ana.item()) <= 2e-2 * abs(ana.item()) + 1e-3, (num.item(), ana.item())
    # attention-prob dropout + FFN inner dropout
    B, H, Lq = 2, 2, 16
    qkv = torch.randn(B * Lq, 3 * H * 64, device=DEV, requires_grad=True)

    def fa(t):
        return ops.attention(t, None, 0, H * 64, 2 * H * 64, B, H, Lq, Lq, None, drop_p=0.25, salt=5)[0]

    o = fa(qkv)
    go = torch.randn_like(o)
    (o * go).sum().backward()
    d = torch.randn_like(qkv)
    num = ((fa(qkv.detach() + eps * d) - fa(qkv.detach() - eps * d)) * go).sum() / (2 * eps)
    ana = (qkv.grad * d).sum()
    assert abs(num.item() - ana.item()) <= 2e-2 * abs(ana.item()) + 1e-3, (num.item(), ana.item())


@pytest.mark.parametrize("cfg", [dict(B=3, H=2, Lq=80, Lk=80, sprel=False), dict(B=4, H=2, Lq=36, Lk=36, sprel=False),
                                 dict(B=2, H=2, Lq=50, Lk=50, sprel=True), dict(B=2, H=2, Lq=37, Lk=160, sprel=False)])
def test_attention_mma_matches_simt_with_dropout(cfg):
    """bf16 tensor-core attention vs the fp32 SIMT kernels on the same (bf16-representable) inputs with attention
    dropout on: both regenerate the same stateless mask, so outputs and gradients agree to bf16 precision."""
    torch.manual_seed(12)
    ops.set_seed(DEV, 4321)
    B, H, Lq, Lk = cfg["B"], cfg["H"], cfg["Lq"], cfg["Lk"]
    hd = H * 64
    lens = torch.randint(max(1, Lk // 2), Lk + 1, (B,), device=DEV).int()
    dists = torch.rand(B, Lq, Lk, device=DEV) * 30 if cfg["sprel"] else None
    res = {}
    base_q = torch.randn(B * Lq, 3 * hd if Lq == Lk else hd, device=DEV).bfloat16()
    base_kv = torch.randn(B * Lk, 2 * hd, device=DEV).bfloat16()
    go = torch.randn(B * Lq, hd, device=DEV).bfloat16()
    gp = torch.randn(B, Lq, Lk, device=DEV) * 0.1
    for dtype in (torch.float32, torch.bfloat16):
        sw = torch.tensor([[-0.05]], device=DEV, requires_grad=True) if dists is not None else None
        sb = torch.tensor([0.1], device=DEV, requires_grad=True) if dists is not None else None
        q = base_q.to(dtype).requires_grad_()
        if Lq == Lk:
            o, pbar = ops.attention(q, None, 0, hd, 2 * hd, B, H, Lq, Lk, lens, dists, sw, sb, True, 0.2, 9)
            kv = None
        else:
            kv = base_kv.to(dtype).requires_grad_()
            o, pbar = ops.attention(q, kv, 0, 0, hd, B, H, Lq, Lk, lens, None, None, None, True, 0.2, 9)
        ((o.float() * go.float()).sum() + (pbar * gp).sum()).backward()
        res[dtype] = (o.float(), pbar, q.grad.float(), kv.grad.float() if kv is not None else None,
                      sw.grad if sw is not None else None)
    a, b_ = res[torch.float32], res[torch.bfloat16]
    assert (a[0] == 0).float().mean() < 0.05  # dropout is on P, not on the output
    close(b_[0], a[0], 2e-2, "mma attn out")
    close(b_[1], a[1], 1e-2, "mma attn pbar")
    close(b_[2], a[2], 4e-2, "mma attn dq(kv)")
    if a[3] is not None:
        close(b_[3], a[3], 4e-2, "mma attn dkv")
    if a[4] is not None:
        close(b_[4], a[4], 5e-2, "mma attn dsprel_w")


def test_adamw_matches_reference_update_order():
    torch.manual_seed(9)
    n = 10007
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV) * 3
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p0 = p.clone()
    lr, b1, b2, eps, wd, maxn = 5e-5, 0.9, 0.98, 1e-6, 0.01, 5.0
    rm, rv, rp = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), p0.clone()
    for step in range(1, 4):
        ss = torch.zeros(1 + 2048, device=DEV)  # [0] = sum g^2, the rest is the kernel's scratch (MAGIC_SUMSQ_SCRATCH)
        L.call("magic_sumsq", g.data_ptr(), n, ss.data_ptr(), 1, L.stream())
        ss2 = torch.zeros(1 + 2048, device=DEV)
        L.call("magic_sumsq", g.data_ptr(), n, ss2.data_ptr(), 1, L.stream())
        assert torch.equal(ss[0], ss2[0])  # deterministic: no floating-point atomics
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        hyper = torch.tensor([lr, lr * math.sqrt(bc2) / bc1, b1, b2, eps, maxn, 0, 0], device=DEV)
        shadow = torch.empty(n, device=DEV, dtype=torch.bfloat16)
        L.call("magic_adamw", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), shadow.data_ptr(), n,
               hyper.data_ptr(), wd, ss.data_ptr(), L.stream())
        # reference order: pretrain_src/optim/adamw.py:84-110 after clip_grad_norm_(5.0)
        norm = g.norm()
        gc = g * min(1.0, (maxn / (norm + 1e-6)).item())
        rm.mul_(b1).add_(gc, alpha=1 - b1)
        rv.mul_(b2).addcmul_(gc, gc, value=1 - b2)
        denom = rv.sqrt().add_(eps)
        rp.addcdiv_(rm, denom, value=-(lr * math.sqrt(bc2) / bc1))
        rp.add_(rp, alpha=-lr * wd)
        close(p, rp, 5e-6, f"adamw p step {step}")
        assert torch.equal(shadow, p.to(torch.bfloat16))
    close(m, rm, 5e-6, "adamw m")
    close(v, rv, 5e-6, "adamw v")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("cfg", [dict(B=3, H=2, L=20, sprel=True), dict(B=2, H=12, L=37, sprel=True),
                                 dict(B=2, H=2, L=80, sprel=False)])
def test_attention_key_skip(dtype, tol, cfg):
    """`key_skip` (the navigation graph's [MEM] slot, agent.py:228): one key inside the valid prefix is never
    attended -- probability exactly 0, no gradient into its K / V rows -- forward and backward, both kernel families."""
    torch.manual_seed(21)
    B, H, L = cfg["B"], cfg["H"], cfg["L"]
    hd = H * 64
    lens = torch.randint(max(3, L // 2), L + 1, (B,), device=DEV)
    dists = torch.rand(B, L, L, device=DEV) * 30 if cfg["sprel"] else None
    sw = torch.tensor([[-0.05]], device=DEV, requires_grad=True) if dists is not None else None
    sb = torch.tensor([0.1], device=DEV, requires_grad=True) if dists is not None else None
    qkv = torch.randn(B * L, 3 * hd, device=DEV).to(dtype).requires_grad_()
    o, pbar = ops.attention(qkv, None, 0, hd, 2 * hd, B, H, L, L, lens.int(), dists, sw, sb, need_pbar=True, salt=77,
                            key_skip=1)
    ref_in = qkv.detach().float().requires_grad_()
    q, k, v = [ref_in[:, i * hd:(i + 1) * hd].reshape(B, L, hd) for i in range(3)]
    qh, kh, vh = [t.view(B, L, H, 64).transpose(1, 2) for t in (q, k, v)]
    s = qh @ kh.transpose(-1, -2) / 8.0
    if dists is not None:
        s = s + (dists * sw.detach() + sb.detach())[:, None]
    m = torch.arange(L, device=DEV)[None] < lens[:, None]
    m[:, 1] = False
    s = s.masked_fill(~m[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    oref = (p @ vh).transpose(1, 2).reshape(B, L, hd)
    close(o.view(B, L, hd), oref, tol, "key_skip out")
    close(pbar, p.mean(1), max(tol, 1e-5), "key_skip pbar")
    assert float(pbar[:, :, 1].abs().max()) == 0.0
    go = torch.randn_like(oref)
    (o.view(B, L, hd).float() * go.to(dtype).float()).sum().backward()
    (oref * go.to(dtype).float()).sum().backward()
    close(qkv.grad, ref_in.grad, tol * 4, "key_skip dqkv")
    g3 = qkv.grad.float().view(B, L, 3 * hd)
    assert float(g3[:, 1, hd:].abs().max()) == 0.0  # the skipped key's K and V rows get exactly zero gradient
