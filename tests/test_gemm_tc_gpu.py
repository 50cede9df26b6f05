"""GPU tests of the tcgen05 / TMA GEMM through the C ABI (`magic_gemm`): every operand layout the linear layers
use (forward K-major x K-major, dgrad K-major x MN-major, wgrad MN-major x MN-major), every tile width, ragged
M/N/K edges, persistent multi-tile CTAs, split-K with TMA reduce-add, and each fused epilogue mode, against fp32
PyTorch math on the same bf16-rounded operands.  bf16 products are exact in fp32, so with fp32 output only the
accumulation order differs: tolerance 2e-5; bf16 output adds one rounding: 4e-3 (north_star bf16 mode: 2e-2)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import magic_b200  # noqa: E402
from magic_b200 import ops, _lib as L  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def bf(t):
    return t.to(torch.bfloat16)


SHAPES = [
    (128, 64, 64), (128, 128, 128), (5120, 128, 128), (5120, 384, 128), (2368, 512, 128), (640, 128, 512),
    (1000, 200, 72), (333, 136, 200), (77, 64, 1024), (4096, 768, 768), (2560, 3072, 768), (300, 50265 // 16 * 8, 128),
    (19000, 1024, 64),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_fwd_layout_bias(M, N, K, out_dtype):
    torch.manual_seed(M + N + K)
    x, w = bf(torch.randn(M, K, device=DEV)), bf(torch.randn(N, K, device=DEV) * 0.1)
    b = torch.randn(N, device=DEV)
    out = torch.full((M, N), float("nan"), device=DEV, dtype=out_dtype)
    assert L.load().magic_gemm_tc_supported(M, N, K) == 1
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b)
    ref = F.linear(x.float(), w.float(), b)
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref) < (4e-3 if out_dtype == torch.bfloat16 else 2e-5)


@pytest.mark.parametrize("M,N,K", SHAPES[:10])
def test_dgrad_layout(M, N, K):
    """dx[M,K] = dy[M,N] W[N,K]: A K-major, B MN-major."""
    torch.manual_seed(1 + M + N + K)
    dy, w = bf(torch.randn(M, N, device=DEV)), bf(torch.randn(N, K, device=DEV) * 0.1)
    dx = torch.full((M, K), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm(dy, N, 1, w, K, 1, dx, M, K, N)
    ref = dy.float() @ w.float()
    assert rel(dx, ref) < 4e-3


@pytest.mark.parametrize("M,N,K", SHAPES[:10] + [(5120, 128, 50272)])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_wgrad_layout_splitk_reduce(M, N, K, beta):
    """dW[N,K] (+)= dy[M,N]^T x[M,K]: both operands MN-major, fp32 output, reduce over the M tokens."""
    if K > 4096:
        M, N, K = 640, K, 128  # the MLM decoder's weight gradient: [vocab(padded ld), h] over 640 masked rows
    torch.manual_seed(2 + M + N + K)
    dy, x = bf(torch.randn(M, N, device=DEV)), bf(torch.randn(M, K, device=DEV))
    gw0 = torch.randn(N, K, device=DEV)
    gw = gw0.clone()
    ops.gemm(dy, 1, N, x, K, 1, gw, N, K, M, beta=beta)
    ref = dy.float().t() @ x.float() + beta * gw0
    assert rel(gw, ref) < 2e-5


@pytest.mark.parametrize("M,N,K", SHAPES[:10] + [(5120, 128, 50272), (11520, 384, 128), (20480, 3072, 768)])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_wgrad_with_bias_grad_single_launch(M, N, K, beta):
    """magic_gemm_wgrad: dW (+)= dy^T x and db += colsum(dy) from ONE launch (ones-tile UMMA), incl. split-K,
    ragged edges (partial last k-block of tokens, partial last row tile of features) and a padded dy stride."""
    if K > 4096:
        M, N, K = 640, 50265, 128
    torch.manual_seed(5 + M + N + K)
    ldy = (N + 7) // 8 * 8
    dy = bf(torch.randn(M, ldy, device=DEV))[:, :N]
    x = bf(torch.randn(M, K, device=DEV))
    gw0, gb0 = torch.randn(N, K, device=DEV), torch.randn(N, device=DEV)
    gw, gb = gw0.clone(), gb0.clone()
    n0 = L.COUNTERS["calls"]
    ops.gemm_wgrad(dy, x, gw, gb, M, N, K, beta)
    assert L.COUNTERS["calls"] == n0 + 1
    # fp64 reference: at a 20 480-token reduction the fp32 torch matmul's own accumulation error is the size of
    # ours; the bound grows with sqrt(reduction length) (4e-5 at M = 20 480, still inside north_star's 1e-4)
    ref_w = (dy.double().t() @ x.double()).float()
    tol = 2e-5 * max(1.0, (M / 5120) ** 0.5)
    assert rel(gw, ref_w + beta * gw0) < tol
    ref_b = gb0 + dy.float().sum(0)
    assert (gb - ref_b).abs().max().item() < 1e-3 * max(1.0, ref_b.abs().max().item())
    # no bias: pointer may be NULL
    gw2 = gw0.clone()
    ops.gemm_wgrad(dy, x, gw2, None, M, N, K, beta)
    assert rel(gw2, ref_w + beta * gw0) < tol


def test_strided_views_packed_output():
    """Column slices of a packed [M, 3h] buffer as C (ldc > N) and as A (lda > K); padded leading dimension."""
    torch.manual_seed(3)
    M, h = 700, 128
    x = bf(torch.randn(M, h, device=DEV))
    ws = [bf(torch.randn(h, h, device=DEV) * 0.1) for _ in range(3)]
    out = torch.zeros(M, 3 * h, device=DEV, dtype=torch.bfloat16)
    for i, w in enumerate(ws):
        ops.gemm(x, h, 1, w, 1, h, out[:, i * h:(i + 1) * h], M, h, h, ldc=3 * h)
    ref = torch.cat([x.float() @ w.float().t() for w in ws], 1)
    assert rel(out, ref) < 4e-3
    # A = a column slice of the packed buffer
    y = torch.empty(M, h, device=DEV, dtype=torch.bfloat16)
    a = out[:, h:2 * h]
    ops.gemm(a, 3 * h, 1, ws[0], 1, h, y, M, h, h)
    assert rel(y, a.float() @ ws[0].float().t()) < 4e-3
    # padded leading dimension of C (MLM logits: N = 50265 -> ld 50272)
    N = 1001
    ldp = (N + 7) // 8 * 8
    w = bf(torch.randn(N, h, device=DEV) * 0.1)
    bias = torch.randn(N, device=DEV)
    buf = torch.zeros(M, ldp, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, h, 1, w, 1, h, buf[:, :N], M, N, h, bias=bias, ldc=ldp)
    assert rel(buf[:, :N], F.linear(x.float(), w.float(), bias)) < 4e-3
    assert float(buf[:, N:].abs().max()) == 0.0  # the TMA store clips at N: padding untouched


@pytest.mark.parametrize("act,fn", [(1, F.gelu), (2, F.relu)])
@pytest.mark.parametrize("M,N,K", [(5120, 512, 128), (300, 200, 136), (2368, 3072, 768)])
def test_epilogue_act_pre_residual(act, fn, M, N, K):
    torch.manual_seed(4 + M)
    x, w = bf(torch.randn(M, K, device=DEV)), bf(torch.randn(N, K, device=DEV) * 0.1)
    b = torch.randn(N, device=DEV) * 0.5
    res = bf(torch.randn(M, N, device=DEV))
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b, act=act, pre_out=pre, residual=res)
    z = F.linear(x.float(), w.float(), b)
    assert rel(pre, z) < 4e-3
    # the activation is applied to the STORED pre-activation (what backward re-reads)
    assert rel(out, fn(pre.float()) + res.float()) < 4e-3
    # backward fusion: dz = (dh W2) * act'(pre)
    dh = bf(torch.randn(M, K, device=DEV))
    dz = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(dh, K, 1, w, 1, K, dz, M, N, K, act=act, dact_pre=pre)
    p = pre.float().requires_grad_()
    fn(p).backward(torch.ones_like(p))
    assert rel(dz, (dh.float() @ w.float().t()) * p.grad) < 4e-3


def test_dropout_epilogue_matches_between_fwd_and_bwd_and_rate():
    torch.manual_seed(5)
    M, N, K = 2048, 512, 128
    x, w = bf(torch.randn(M, K, device=DEV)), bf(torch.randn(N, K, device=DEV) * 0.1)
    ops.set_seed(torch.device(DEV, 0), 77)
    seed = ops.seed_tensor(torch.device(DEV, 0))
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    pre = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K, act=1, pre_out=pre, drop_p=0.25, salt=9, seed=seed)
    keep = out != 0
    rate = 1.0 - keep.float().mean().item()
    assert abs(rate - 0.25) < 0.01
    assert rel(out[keep], F.gelu(pre)[keep] / 0.75) < 1e-5
    # backward regenerates the same mask from (seed, salt, index)
    dh = bf(torch.ones(M, K, device=DEV))
    dz = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(dh, K, 1, w, 1, K, dz, M, N, K, act=1, dact_pre=pre, drop_p=0.25, salt=9, seed=seed)
    assert bool(((dz != 0) == keep)[pre.abs() > 1e-3].all())


def test_unaligned_shapes_take_the_fp32_kernel():
    """ldc not 16-byte aligned -> the FFMA kernel handles it (no silent corruption)."""
    torch.manual_seed(6)
    M, N, K = 37, 50, 768
    x, w = bf(torch.randn(M, K, device=DEV)), bf(torch.randn(N, K, device=DEV) * 0.1)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K)
    assert rel(out, x.float() @ w.float().t()) < 4e-3


# ---- CTA-pair kernels (cta_group::2, 256 x 256 tiles): shapes the heuristic sends there (>= 37 pair tiles) ---------
PAIR_SHAPES = [(5120, 768, 768), (5120, 3072, 768), (5120, 768, 3072), (11520, 768, 256), (5000, 712, 520),
               (20480, 2304, 768)]


@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
def test_pair_fwd_bias_gelu_residual(M, N, K):
    torch.manual_seed(7 + M + N + K)
    x, w = bf(torch.randn(M, K, device=DEV)), bf(torch.randn(N, K, device=DEV) * 0.05)
    b = torch.randn(N, device=DEV)
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b)
    ref = F.linear(x.float(), w.float(), b)
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref) < 4e-3
    # GELU + pre-activation copy-out, then residual epilogue
    pre = torch.empty_like(out)
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b, act=1, pre_out=pre)
    assert rel(pre, ref) < 4e-3
    assert rel(out, F.gelu(bf(ref).float())) < 6e-3
    res = bf(torch.randn(M, N, device=DEV))
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b, residual=res)
    assert rel(out, ref + res.float()) < 4e-3


@pytest.mark.parametrize("M,N,K", PAIR_SHAPES[:5])
def test_pair_dgrad_mn_major_b_and_dact(M, N, K):
    """dx[M,K] = (dy[M,N] W[N,K]) * gelu'(pre): A K-major, B MN-major, backward-activation epilogue."""
    torch.manual_seed(8 + M + N + K)
    dy, w = bf(torch.randn(M, N, device=DEV)), bf(torch.randn(N, K, device=DEV) * 0.05)
    dx = torch.full((M, K), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm(dy, N, 1, w, K, 1, dx, M, K, N)
    ref = dy.float() @ w.float()
    assert rel(dx, ref) < 4e-3
    pre = bf(torch.randn(M, K, device=DEV))
    ops.gemm(dy, N, 1, w, K, 1, dx, M, K, N, act=1, dact_pre=pre)
    z = pre.float()
    g = 0.5 * (1 + torch.erf(z * 0.70710678)) + z * torch.exp(-0.5 * z * z) * 0.3989422804
    assert rel(dx, ref * g) < 6e-3
