"""The conditioning of the reference's Polya-Gamma log-density series (polyagamma.jl:37-91), demonstrated instead of argued.

tests/test_gpu_gibbs.py holds `aug_loglik` to 1e-11 relative where that series is well conditioned (Bernoulli, Categorical:
b = 1) and only to 1e-7 per observation for NegBin / Poisson (b = y + r resp. y + n of the order of 20-40, omega of the order
of b/4).  The reason: the pair terms  prod_n * R_n * exp(-R_n^2 / 8x) * (1 - c_nb exp(-(R_n + 1) / 2x))  change sign and cancel
for x >> 1, so ANY fp64 evaluation of the formula the reference writes down — the reference's own, the oracle's, the CUDA
kernel's — carries an absolute error far above 1e-12.  Here the same formula is evaluated with mpmath at 80 digits and compared
with the fp64 oracle:

  * b = 1 (Bernoulli / Categorical), x in the bulk of PG(1, c): fp64 agrees with the 80-digit value to 1e-12;
  * b = 20 .. 40, x = 5 .. 12 (the NegBin / Poisson regime): fp64 is off by 1e-11 .. 7e-9 ABSOLUTE (a 1e-12 parity bar
    between two fp64 implementations is meaningless there), always below the 1e-7 the GPU test allows per observation;
  * the cancellation itself: the largest pair term exceeds the sum by a factor 1.6e6 at b = 30, x = 10, and by 1e41 at
    x = b = 30 (far tail: the fp64 value is pure rounding noise there — and so is the reference's).
"""
import math

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")


@pytest.fixture(scope="module")
def orc():
    from oracle import orc as o
    o.lib()
    return o


def series_mp(x, b, half_n=100):
    """calc_series (polyagamma.jl:75-91) and the largest |pair term|, at the working precision of mpmath"""
    x, b = mp.mpf(x), mp.mpf(b)
    prods = [mp.mpf(1)]
    for m in range(1, 2 * half_n + 1):
        prods.append(prods[-1] * (1 + (b - 1) / m))
    total, biggest = mp.mpf(0), mp.mpf(0)
    for n in range(0, 2 * half_n + 1, 2):
        Rn = 2 * n + b
        c_nb = ((n + b) / (n + 1)) * (2 / Rn + 1)
        inner = 1 - c_nb * mp.exp((Rn + 1) / (-2 * x))
        term = prods[n] * Rn * mp.exp(Rn * Rn / (-8 * x)) * inner
        total += term
        biggest = max(biggest, abs(term))
    return total, biggest


def logpdf_mp(b, x):
    """logpdf(PolyaGamma(b, 0), x) as polyagamma.jl:37-53 writes it (x >= 1e-2 branch), at the working precision"""
    s, _ = series_mp(x, b)
    ext = (mp.mpf(b) - 1) * mp.log(2) - (mp.log(2 * mp.pi) + 3 * mp.log(mp.mpf(x))) / 2
    return ext + mp.log(s)


def test_b1_series_is_well_conditioned(orc):
    mp.mp.dps = 80
    xs = np.array([0.02, 0.05, 0.1, 0.25, 0.5, 1.0, 2.0])
    got = orc.pg_logpdf(1.0, 0.0, xs)
    for x, g in zip(xs, got):
        want = float(logpdf_mp(1.0, x))
        assert abs(g - want) <= 1e-12 * max(1.0, abs(want)), (x, g, want)


@pytest.mark.parametrize("b,x", [(20.0, 5.0), (20.0, 8.0), (30.0, 8.0), (30.0, 10.0), (40.0, 10.0), (40.0, 12.0)])
def test_large_b_series_cancels_and_fp64_is_only_good_to_1e7(orc, b, x):
    mp.mp.dps = 80
    want = logpdf_mp(b, x)
    got = float(orc.pg_logpdf(b, 0.0, np.array([x]))[0])
    err = abs(got - float(want))
    total, biggest = series_mp(x, b)
    # the fp64 evaluation is within the per-observation tolerance of the GPU test (tests/test_gpu_gibbs.py: 1e-7 * n) ...
    assert err < 1e-7, (b, x, got, float(want), err)
    # ... and the sum is smaller than its largest term by the factor the rounding error is amplified with
    amplification = float(biggest / abs(total))
    assert amplification > 10.0, (b, x, amplification)
    # an fp64 sum of such terms cannot be trusted below ~amplification * eps: the 1e-12 bar is out of reach when this exceeds it
    assert err <= 64 * amplification * np.finfo(float).eps + 1e-15, (b, x, err, amplification)


def test_fp64_error_exceeds_the_1e12_bar_somewhere_in_the_negbin_regime(orc):
    mp.mp.dps = 80
    errs = []
    for b, x in [(20.0, 8.0), (30.0, 10.0), (40.0, 12.0)]:
        errs.append(abs(float(orc.pg_logpdf(b, 0.0, np.array([x]))[0]) - float(logpdf_mp(b, x))))
    assert max(errs) > 1e-10 and max(errs) < 1e-7, errs


def test_cancellation_factor_at_the_negbin_regime():
    mp.mp.dps = 80
    total, biggest = series_mp(10.0, 30.0)
    assert biggest / abs(total) > 1e6
    t_tail, big_tail = series_mp(30.0, 30.0)
    assert big_tail / abs(t_tail) > 1e30          # far tail: no fp64 evaluation of this formula means anything
    # with 80 digits the same sum is stable: 60 and 80 digits agree to 1e-50
    mp.mp.dps = 60
    t60, _ = series_mp(10.0, 30.0)
    assert abs(t60 - total) / abs(total) < mp.mpf(10) ** -50
    assert math.isfinite(float(mp.log(total)))
