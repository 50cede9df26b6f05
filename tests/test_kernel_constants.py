"""CPU checks of numeric constants baked into the CUDA sources (no GPU, no compilation): the coefficients are parsed
out of csrc/ and the formula is re-evaluated in float32 numpy against the exact function."""
import os
import re

import numpy as np
from scipy.special import erf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = os.path.join(ROOT, "vln-magic_b200", "csrc", "common.cuh")


def _gelu_tail_coeffs():
    src = open(COMMON).read()
    body = src[src.index("float gelu_tail(float a)"):]
    body = body[:body.index("asm(")]
    nums = [float(x[:-1]) for x in re.findall(r"-?\d\.\d+e[+-]\d+f", body)]
    assert len(nums) == 6, nums
    # fmaf(a, c5, c4) then q = fmaf(q, a, c3) ... : highest degree first
    return nums


def _tail32(a):
    c = _gelu_tail_coeffs()
    a = np.minimum(a.astype(np.float32), np.float32(6.0))
    q = a * np.float32(c[0]) + np.float32(c[1])
    for k in c[2:]:
        q = (q * a + np.float32(k)).astype(np.float32)
    return np.exp2(q.astype(np.float64))


def test_one_mufu_gelu_matches_the_exact_gelu():
    """gelu(x) = max(x, 0) - |x| * 2^Q(min(|x|, 6)) with the degree-5 Q of common.cuh (the tcgen05 GEMM's GELU epilogue):
    |error| < 1e-6 everywhere (DESIGN.md 4 quotes 5e-7 for the fit; ex2.approx adds 2 ulp of the tail), i.e. far inside
    bf16 resolution and inside the 1e-4 fp32 parity tolerance."""
    x = np.linspace(-10, 10, 800001)
    g = np.maximum(x, 0) - np.abs(x) * _tail32(np.abs(x))
    ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
    assert np.abs(g - ref).max() < 1e-6
    # Phi(-a) itself (used by the backward epilogue: cdf = 1 - tail for x >= 0): the fit weights the error by |x| (what
    # the GELU value sees), so the tail is loosest at a = 0 (1.3e-5)
    a = np.linspace(0, 6, 60001)
    assert np.abs(_tail32(a) - 0.5 * (1 - erf(a / np.sqrt(2)))).max() < 2e-5


def test_gelu_gradient_formula():
    """gelu'(x) = Phi(x) + x * phi(x) from the same tail and exp2(-x^2 / (2 ln 2)): constants as written in common.cuh."""
    src = open(COMMON).read()
    k = float(re.search(r"\(-(0\.72134752\d+)f \* x \* x\)", src).group(1))
    assert abs(k - 0.5 / np.log(2)) < 1e-9
    inv_sqrt_2pi = float(re.search(r"x \* (0\.39894228\d+)f, g, cdf", src).group(1))
    assert abs(inv_sqrt_2pi - 1 / np.sqrt(2 * np.pi)) < 1e-9
    x = np.linspace(-8, 8, 160001)
    t = _tail32(np.abs(x))
    cdf = np.where(x >= 0, 1 - t, t)
    grad = cdf + x * inv_sqrt_2pi * np.exp2(-k * x * x)
    ref = 0.5 * (1 + erf(x / np.sqrt(2))) + x * np.exp(-0.5 * x * x) / np.sqrt(2 * np.pi)
    assert np.abs(grad - ref).max() < 3e-5
