"""CPU checks of the closed-form approximations the CUDA kernels use (constants restated here; numpy emulates the f32 arithmetic):
the polynomial exp2 of the encoder attention (attention_tcgen05.cu: fa_ex2_poly2) and the erf GELU of the FC1 epilogue
(common.cuh: gelu_erf / gelu_erf2).  They bound the approximation error itself; the kernels are checked end to end in the -m gpu tests."""
import numpy as np
from scipy.special import erf

F = np.float32


def test_polynomial_exp2_of_the_attention_kernel():
    x = np.concatenate([np.linspace(-140, 8, 400001), [-np.inf, -126.0, -125.5, -0.5, 0.0, 0.5, 7.99, 8.0]]).astype(F)
    xc = np.maximum(x, F(-126))
    magic = F(12582912.0)                      # 1.5 * 2^23: the low mantissa bits of t hold round(x)
    t = (xc + magic).astype(F)
    nf = (t - magic).astype(F)
    f = (xc - nf).astype(F)
    assert np.abs(f).max() <= 0.5
    r = (f * F(0.05517146) + F(0.24261086)).astype(F)
    r = (r * f + F(0.69326097)).astype(F)
    r = (r * f + F(0.9999281)).astype(F)
    bits = (r.view(np.int32).astype(np.int64) + ((t.view(np.int32).astype(np.int64) << 23) & 0xFFFFFFFF)) & 0xFFFFFFFF
    got = bits.astype(np.uint32).view(F)
    ref = np.exp2(xc.astype(np.float64))
    rel = np.abs(got / ref - 1)
    live = xc > -120                           # (below: the clamp region, probabilities ~1e-36 that round to nothing)
    assert rel[live].max() < 8e-5, rel[live].max()
    assert np.all(np.isfinite(got)) and got.min() >= 0


def test_erf_gelu_of_the_fc1_epilogue():
    x = np.linspace(-9, 9, 600001).astype(F)
    z = (np.abs(x) * F(0.70710678118654752440)).astype(F)
    t = (F(1) / (F(0.3275911) * z + F(1))).astype(F)
    poly = (F(1.061405429) * t + F(-1.453152027)).astype(F)
    for c in (1.421413741, -0.284496736, 0.254829592):
        poly = (poly * t + F(c)).astype(F)
    e = np.exp2(((z * z).astype(F) * F(-1.4426950408889634)).astype(F)).astype(F)
    erf_abs = ((-poly * t).astype(F) * e + F(1)).astype(F)
    hx = (F(0.5) * x).astype(F)
    got = (hx * np.copysign(erf_abs, x) + hx).astype(F)
    ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    assert np.abs(got - ref).max() < 1e-6, np.abs(got - ref).max()
