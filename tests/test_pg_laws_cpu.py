"""CPU checks of the mathematics behind the general-b Pólya-Gamma sampler (csrc/aug_pgb.cuh) — no GPU, no sampling.

(1) The exact fractional piece: R(x; e) = f(x)/a_0(x) — the J*(e, 0) density over its first series term — must lie in
    [0, 1] and decrease in x for the rejection sampler to be valid.  Checked to 60 digits on a grid of (e, x), and the
    tilted first term is identified with the InverseGaussian(e/z, e^2) proposal.
(2) The certified Gamma convolution for b > 4: the KOLMOGOROV DISTANCE between the law that is sampled (KT explicit
    Gamma terms + a three-cumulant shifted-Gamma tail, KT from the rule in conv_kt) and the exact PG(b, c) law is
    computed from the two characteristic functions by Gil-Pelaez inversion — it is a number, not a sampling estimate —
    and must stay below 3e-6 on the whole (b, c) grid (3e-7 for |c| <= 2.5).  For comparison the reference's own
    200-term truncation (polyagamma.jl:157-164) is biased by b/(2 pi^2 200) in the mean.
(3) The tail-sum polynomials compiled into the kernel (parsed from the header) against Hurwitz-zeta sums.
"""
import os
import re

import mpmath as mp
import numpy as np
import pytest
from scipy import special

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "augmentedgplikelihoods.jl_b200", "csrc", "aug_pgb.cuh")
PI = np.pi


# ------------------------------------------------------------------ (1) fractional piece
def R_series(x, e, dps=60):
    mp.mp.dps = dps
    x, e = mp.mpf(x), mp.mpf(e)
    q = mp.e ** (-2 / x)
    s, cn, n = mp.mpf(0), mp.mpf(1), 0
    while True:
        t = cn * q ** (n * (n + e))
        s += (-1) ** n * t
        if n > 3 and t < mp.mpf(10) ** -(dps - 10):
            return s
        cn = cn * (n + e) / (n + 1) * (2 * n + 2 + e) / (2 * n + e)
        n += 1


def test_fractional_ratio_is_a_probability_and_decreases():
    xs = [0.02, 0.1, 0.3, 0.6, 1, 1.5, 2, 2.5, 2.9, 3.3, 4, 5, 7, 10, 15, 25, 48]
    for e in [1e-6, 0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.999, 1.0]:
        prev = mp.mpf(1)
        for x in xs:
            r = R_series(x, e)
            assert 0 <= r <= 1, (e, x, r)
            assert r <= prev, (e, x)
            prev = r
        # the kernel rejects x > 48 outright: R is then below 1e-3 of the smallest uniform it can draw (2^-53)
        assert R_series(48, e) < mp.mpf(2) ** -63


def test_fractional_ratio_for_e_equal_one_is_jacobis_product():
    # sum (-1)^n (2n+1) q^{n(n+1)} = prod (1 - q^{2m})^3
    for x in [0.5, 1.0, 3.0, 8.0]:
        q = mp.e ** (-2 / mp.mpf(x))
        prod = mp.nprod(lambda m: (1 - q ** (2 * m)) ** 3, [1, mp.inf])
        assert abs(R_series(x, 1) - prod) < mp.mpf(10) ** -40


def test_tilted_first_term_is_the_inverse_gaussian_proposal():
    # a_0(x) exp(-z^2 x/2) = 2^e exp(-e z) IG(x; mean e/z, shape e^2)
    for e, z, x in [(0.5, 0.4, 0.7), (0.3, 2.0, 0.05), (0.9, 0.1, 5.0)]:
        a0 = 2 ** e * e / np.sqrt(2 * PI * x ** 3) * np.exp(-e * e / (2 * x))
        mu, lam = e / z, e * e
        ig = np.sqrt(lam / (2 * PI * x ** 3)) * np.exp(-lam * (x - mu) ** 2 / (2 * mu * mu * x))
        assert a0 * np.exp(-z * z * x / 2) == pytest.approx(2 ** e * np.exp(-e * z) * ig, rel=1e-13)
    # hence the acceptance rate: cosh(z)^-e / (2^e exp(-e z)) = (1 + exp(-2z))^-e >= 2^-e
    for e, z in [(0.5, 0.0), (0.5, 0.5), (1.0, 3.0)]:
        assert np.cosh(z) ** -e / (2 ** e * np.exp(-e * z)) == pytest.approx((1 + np.exp(-2 * z)) ** -e, rel=1e-14)


# ------------------------------------------------------------------ (2) convolution: Kolmogorov distance to the exact law
def conv_kt(absc):                                            # mirrors augb::conv_kt
    return 2 if absc <= 3.0 else min(int(np.ceil(0.4 * absc)) + 1, 32)


def tail_sums(kt, w):
    """T_j = sum_{k > kt} ((k - 1/2)^2 + w)^-j, j = 1..3 (direct sums + Euler-Maclaurin remainder)"""
    N = 200000
    k = np.arange(kt + 1, N + 1, dtype=np.float64)
    d = (k - 0.5) ** 2 + w
    a = float(N)                                              # remainder: integral from N of (u^2 + w)^-j (midpoint rule)
    r1 = (PI / 2 - np.arctan(a / np.sqrt(w))) / np.sqrt(w) if w > 0 else 1 / a
    return np.sum(1 / d) + r1, np.sum(d ** -2.0) + 1 / (3 * a ** 3), np.sum(d ** -3.0) + 1 / (5 * a ** 5)


def log_cosh(u):
    return u + np.log1p(np.exp(-2 * u)) - np.log(2)


def cf_exact(t, b, c):                                        # E exp(i t w), w ~ PG(b, c)
    z = np.sqrt(c * c / 4 - 0.5j * t)
    return np.exp(b * (log_cosh(c / 2 + 0j) - log_cosh(z)))


def cf_sampled(t, b, c):                                      # the law pgb_kernel draws from for b > 4
    s = -1j * t
    w = (c / (2 * PI)) ** 2
    kt = conv_kt(abs(c))
    out = np.ones_like(t, dtype=complex)
    for k in range(1, kt + 1):
        out *= (1 + s / (2 * PI * PI * ((k - 0.5) ** 2 + w))) ** (-b)
    T1, T2, T3 = tail_sums(kt, w)
    theta, shape, loc = T3 / T2, b * T2 ** 3 / T3 ** 2, b * (T1 - T2 * T2 / T3)
    assert loc >= 0 and shape >= 1
    sc = 1 / (2 * PI * PI)
    return out * np.exp(-s * loc * sc) * (1 + s * theta * sc) ** (-shape)


def kolmogorov_distance(b, c, nt=200001, nx=81):
    var = b / 24 if c < 1e-8 else b / (4 * c ** 3) * (np.sinh(c) - c) / np.cosh(c / 2) ** 2
    mean = b / 4 if c < 1e-8 else b / (2 * c) * np.tanh(c / 2)
    sd = np.sqrt(var)
    t = np.linspace(0, 120 / sd, nt)[1:]
    g = (cf_exact(t, b, c) - cf_sampled(t, b, c)) / t
    xs = mean + sd * np.linspace(-5, 9, nx)
    xs = xs[xs > 0]
    dt = t[1] - t[0]
    return max(abs(np.sum(np.imag(np.exp(-1j * t * x) * g)) * dt / PI) for x in xs)


@pytest.mark.parametrize("b", [4.25, 5.5, 10.0, 25.5, 100.0])
def test_convolution_law_is_within_3e_6_of_the_exact_law(b):
    for c in [0.0, 1.0, 2.5, 3.0, 3.01, 5.0, 10.0, 20.0, 40.0]:
        d = kolmogorov_distance(b, c)
        assert d < 3e-6, (b, c, d)
        if c <= 2.5:
            assert d < 4e-7, (b, c, d)


def test_inversion_resolves_a_known_discrepancy():
    """sanity of the method itself: the two-cumulant, 4-term tail of round 1 at b = 0.5 is ~2e-4 away from the exact law"""
    def cf_round1(t, b, c):
        s = -1j * t
        w = (c / (2 * PI)) ** 2
        out = np.ones_like(t, dtype=complex)
        for k in range(1, 5):
            out *= (1 + s / (2 * PI * PI * ((k - 0.5) ** 2 + w))) ** (-b)
        T1, T2, _ = tail_sums(4, w)
        return out * (1 + s * (T2 / T1) / (2 * PI * PI)) ** (-b * T1 * T1 / T2)
    b, c = 0.5, 0.0
    sd, mean = np.sqrt(b / 24), b / 4
    t = np.linspace(0, 400 / sd, 800001)[1:]
    g = (cf_exact(t, b, c) - cf_round1(t, b, c)) / t
    xs = mean + sd * np.linspace(-1.7, 8, 81)
    xs = xs[xs > 0]
    d = max(abs(np.sum(np.imag(np.exp(-1j * t * x) * g)) * (t[1] - t[0]) / PI) for x in xs)
    assert 5e-5 < d < 5e-4, d


# ------------------------------------------------------------------ (3) compiled tail-sum polynomials
def header_table(name):
    src = open(HDR).read()
    m = re.search(r"%s\[\d+\]\s*=\s*\{([^}]*)\}" % name, src)
    return [float(v) for v in m.group(1).replace("\n", " ").split(",") if v.strip()]


def test_compiled_tail_polynomials_match_hurwitz_zeta_sums():
    tabs = {1: header_table("TAIL1"), 2: header_table("TAIL2"), 3: header_table("TAIL3")}
    assert (len(tabs[1]), len(tabs[2]), len(tabs[3])) == (14, 11, 11)
    for j, tol in [(1, 1e-15), (2, 1e-11), (3, 1e-10)]:
        coef = tabs[j]                                          # highest power first (Horner)
        M = len(coef)
        for m in range(M):                                      # (-1)^m binom(j+m-1, m) zeta(2(j+m), 5/2)
            want = (-1) ** m * special.comb(j + m - 1, m) * special.zeta(2 * (j + m), 2.5)
            assert coef[M - 1 - m] == pytest.approx(want, rel=1e-14)
        for x in [0.0, 0.3, 1.0, 2.0]:                          # kernel uses the series for x = |c|/2 <= 2
            w = (x / PI) ** 2
            got = np.polyval(coef, w)
            want = tail_sums(2, w)[j - 1]
            assert got == pytest.approx(want, rel=tol), (j, x)
