"""CPU tests that PIN the oracle (oracle/aug_oracle.cpp).

1. every known-answer assertion the reference's own tests hold for this path
   (SURVEY §4 / §8c): test/SpecialDistributions/polyagamma.jl:27-37,
   test/utils.jl:1-14, test/likelihoods/laplace.jl:6-9;
2. the committed 50-digit mpmath golden vectors (tests/golden/make_golden.py);
3. the two test_auglik invariants (src/TestUtils.jl:107-148) re-expressed on the oracle.
"""
import math

import mpmath as mp
import numpy as np
import pytest

from common import (BERNOULLI, CAT, CAT_BIJ, HETERO, LAPLACE, NEGBIN, POISSON, STUDENTT, golden_arrays,
                    lik_args, load_golden, relerr, synth_inputs)

GOLD = load_golden()


# ---------------------------------------------------------------- reference pins
def test_pg_mean_exact(orc):
    # test/SpecialDistributions/polyagamma.jl:27-28 (exact equality)
    assert orc.pg_mean(1.0, 0.0) == 1 / 4
    assert orc.pg_mean(1.0, 2.0) == math.tanh(1.0) / 4


PG_GRID = [(1, 0.0), (1, 2.0), (3, 0.0), (3, 2.5), (3, 3.2), (1.2, 3.2)]


def _ref_logpdf_mp(b, c, x):
    """ref_logpdf of test/SpecialDistributions/polyagamma.jl:3-19 (4001-term series), at 50 digits."""
    mp.mp.dps = 60
    b, c, x = mp.mpf(b), mp.mpf(c), mp.mpf(x)
    ext = b * mp.log(mp.cosh(c / 2)) - c * c * x / 2 + (b - 1) * mp.log(2) - mp.loggamma(b) - \
        (mp.log(2 * mp.pi) + 3 * mp.log(x)) / 2
    s = mp.mpf(0)
    for n in range(0, 4002):
        t = mp.exp(mp.loggamma(n + b) - mp.loggamma(n + 1) - (2 * n + b) ** 2 / (8 * x) + mp.log(2 * n + b))
        s += t if n % 2 == 0 else -t
        if n > 20 and t < mp.mpf(10) ** -70:
            break
    return float(ext + mp.log(s))


@pytest.mark.parametrize("b,c", PG_GRID)
def test_pg_logpdf_pins(orc, b, c):
    # :33 all(isreal, logpdf.(p, 10 .^ (-7:0.1:7)))
    xs = 10.0 ** np.arange(-7, 7.0001, 0.1)
    lp = orc.pg_logpdf(b, c, xs)
    assert not np.any(np.isnan(lp))
    # :35-36 ref_logpdf ≈ logpdf on 10 .^ (-2.5:0.1:0.5)  (Julia ≈ : rtol sqrt(eps) on the vector norm)
    xs = 10.0 ** np.arange(-2.5, 0.5001, 0.1)
    lp = orc.pg_logpdf(b, c, xs)
    ref = np.array([_ref_logpdf_mp(b, c, x) for x in xs])
    assert np.linalg.norm(lp - ref) <= 1.5e-8 * max(np.linalg.norm(lp), np.linalg.norm(ref))
    assert np.max(np.abs(lp - ref)) < 1e-9


@pytest.mark.parametrize("b,c", PG_GRID)
def test_pg_sampler_mean_pin(orc, b, c):
    # :37 mean(rand(p, 10000)) ≈ mean(p) atol = 1e-2
    x = orc.pg_rand_bc(42, 10000, b, c, float(b).is_integer() and isinstance(b, int))
    assert abs(x.mean() - orc.pg_mean(b, c)) < 1e-2


def test_utils_pins(orc):
    # test/utils.jl:1-14
    L = orc.lib()
    rng = np.random.default_rng(42)
    m, s = rng.standard_normal(), rng.random()
    assert L.orc_second_moment(m, s * s) == pytest.approx(m * m + s * s, rel=1e-15)
    c = m * m + s * s
    assert abs(L.orc_approx_expected_logistic(m, c) - math.exp(m / 2) / math.cosh(c / 2) / 2) < 1e-5
    bigm = 1000.0
    c = bigm + abs(rng.standard_normal())
    assert L.orc_approx_expected_logistic(bigm, c) == 1.0 == L.orc_logistic(bigm)
    assert L.orc_approx_expected_logistic(-800.0, 3.0) == 0.0


def test_laplace_kl_pin():
    # test/likelihoods/laplace.jl:6-9: closed form == KL(InverseGaussian(μ, 2λ) ‖ InverseGamma(1/2, λ))
    mp.mp.dps = 30
    rng = np.random.default_rng(7)
    lam, mu = mp.mpf(rng.random()), mp.mpf(rng.random())

    def logq(x):  # InverseGaussian(μ, 2λ)
        L2 = 2 * lam
        return (mp.log(L2) - mp.log(2 * mp.pi) - 3 * mp.log(x) - L2 * (x - mu) ** 2 / (mu ** 2 * x)) / 2

    def logp(x):  # InverseGamma(1/2, λ)
        return mp.log(lam) / 2 - mp.loggamma(mp.mpf(1) / 2) - mp.mpf(3) / 2 * mp.log(x) - lam / x

    kl = mp.quad(lambda x: mp.exp(logq(x)) * (logq(x) - logp(x)), [0, mu / 10, mu, 10 * mu, mp.inf])
    closed = mp.log(2 * lam) / 2 - mp.log(2 * mp.pi) / 2 - mp.log(lam) / 2 + mp.loggamma(mp.mpf(1) / 2) + lam / mu
    assert abs(kl - closed) < 1e-12 * abs(closed)


def test_pg_moments_closed_form_vs_density(orc):
    # oracle self-validation: integrate exp(logpdf) -> mass 1, mean, variance
    for b, c in [(1, 0.0), (1, 2.0), (3, 2.5), (1.2, 3.2), (0.5, 1.0)]:
        xs = np.concatenate([np.geomspace(1e-4, 1e-2, 400, endpoint=False), np.linspace(1e-2, 12.0, 60000)])
        pdf = np.exp(orc.pg_logpdf(b, c, xs))
        mass = np.trapezoid(pdf, xs)
        mean = np.trapezoid(pdf * xs, xs)
        var = np.trapezoid(pdf * xs * xs, xs) - mean ** 2
        assert abs(mass - 1) < 1e-5
        assert abs(mean - orc.pg_mean(b, c)) < 1e-5
        assert abs(var - orc.pg_var(b, c)) < 1e-5


# ---------------------------------------------------------------- mpmath golden vectors
@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_vs_golden(orc, name):
    case = GOLD[name]
    kind, params, kw = lik_args(case)
    lik = orc.make_lik(kind, *params, **kw)
    y, mu, var = golden_arrays(case)
    want_scalars = kind != CAT
    rc, state, beta, gamma, seq, comp = orc.cavi_step(lik, y, mu, var, want_scalars=want_scalars)
    assert rc == 0
    assert relerr(state[0], case["s0"]) < 1e-14
    if "s1" in case:
        assert relerr(state[1], case["s1"]) < 1e-13
    if "s2" in case:
        assert relerr(state[2], case["s2"]) < 1e-14
    assert relerr(beta, case["beta"]) < 2e-13
    assert relerr(gamma, case["gamma"]) < 2e-13
    if want_scalars:
        for got in (seq, comp):
            assert got[0] == pytest.approx(case["elt"], rel=1e-12)
            assert got[1] == pytest.approx(case["kl"], rel=1e-12, abs=1e-13)
            tot = case.get("eall", case["elt"] + case["kl"])
            assert got[2] == pytest.approx(tot, rel=1e-12)
    if kind in (NEGBIN, POISSON, CAT, CAT_BIJ):
        assert np.array_equal(state[2], y)


def test_cat_nonbij_kl_errors(orc):
    # categorical.jl:165-170
    case = GOLD["cat_K3"]
    kind, params, kw = lik_args(case)
    lik = orc.make_lik(kind, *params, **kw)
    y, mu, var = golden_arrays(case)
    rc, *_ = orc.cavi_step(lik, y, mu, var, want_scalars=True)
    assert rc == -3


# ---------------------------------------------------------------- test_auglik invariants
SPLIT_LIKS = [
    ("bernoulli", BERNOULLI, (), {}),
    ("negbin10", NEGBIN, (10,), dict(r_is_int=True)),
    ("negbin5.5", NEGBIN, (5.5,), {}),
    ("poisson10", POISSON, (10.0,), {}),
    ("laplace1", LAPLACE, (1.0,), {}),
    ("studentt", STUDENTT, (3.0, 1.5), {}),
]


@pytest.mark.parametrize("name,kind,params,kw", SPLIT_LIKS)
def test_full_conditional_omega_invariant(orc, name, kind, params, kw):
    # src/TestUtils.jl:107-116: aug_loglik(Ω) - log p(Ω|y,f) is the same for any Ω
    n = 10
    lik = orc.make_lik(kind, *params, **kw)
    y, mu, var, f = synth_inputs(kind, n, 11, params)
    vals = []
    for seed in (1, 2):
        om, nv = orc.aux_sample(lik, seed, y, f)
        seq, comp = orc.sampled_loglik_terms(lik, y, f, om, nv, True)
        vals.append(comp[5] - orc.full_conditional_logdensity(lik, y, f, om, nv))
    # real-b PG draws use the reference's 200-term truncated Gamma sum, the density does not: same tol
    assert abs(vals[0] - vals[1]) < 1e-5


@pytest.mark.parametrize("name,kind,params,kw", SPLIT_LIKS)
def test_full_conditional_f_invariant(orc, name, kind, params, kw):
    # src/TestUtils.jl:118-131: logtilt + log N(f|0,K) - log N(f|m,S) is the same for any f
    n = 10
    rng = np.random.default_rng(5)
    lik = orc.make_lik(kind, *params, **kw)
    y, mu, var, f = synth_inputs(kind, n, 12, params)
    om, nv = orc.aux_sample(lik, 3, y, f)
    A = rng.random((n, n))
    K = A @ A.T + 1e-6 * np.eye(n)
    beta, gamma = orc.potential_precision(lik, y, f, om, nv)
    S = np.linalg.inv(np.linalg.inv(K) + np.diag(gamma[0]))
    S = (S + S.T) / 2
    m = S @ beta[0]

    def logmvn(x, mean, cov):
        d = x - mean
        sign, ld = np.linalg.slogdet(cov)
        return -0.5 * (d @ np.linalg.solve(cov, d) + ld + n * math.log(2 * math.pi))

    vals = []
    Lc = np.linalg.cholesky(S)
    for _ in range(2):
        fs = np.ascontiguousarray(m + Lc @ rng.standard_normal(n))
        seq, comp = orc.sampled_loglik_terms(lik, y, fs, om, nv, False)
        vals.append(comp[3] + logmvn(fs, np.zeros(n), K) - logmvn(fs, m, S))
    assert abs(vals[0] - vals[1]) < 1e-5 * max(1.0, abs(vals[0]))


def test_fused_equals_separate(orc):
    # src/TestUtils.jl:80-87,162-169: the fused (β, γ) ≈ the separate calls
    for name, case in GOLD.items():
        kind, params, kw = lik_args(case)
        lik = orc.make_lik(kind, *params, **kw)
        y, mu, var = golden_arrays(case)
        state = orc.alloc_state(lik, y.shape[0])
        orc.aux_posterior(lik, y, mu, var, state)
        rc, b1, g1 = orc.expected_potential_precision(lik, y, mu, state)
        rc2, _, b2, g2, _, _ = orc.cavi_step(lik, y, mu, var, want_scalars=False)
        assert rc == 0 and rc2 == 0
        assert np.array_equal(b1, b2) and np.array_equal(g1, g2)
        assert np.all(g1 >= 0)          # :88,171


# ---- the oracle restatements against independent numpy / mpmath evaluations
def test_next_rows_oracle_vs_independent(orc):
    import mpmath as mp
    rng = np.random.default_rng(12)
    n = 40
    y = rng.standard_normal(n)
    mu = rng.standard_normal((2, n))
    var = (0.5 + rng.random((2, n))) ** 2
    f = rng.standard_normal((2, n))
    mp.mp.dps = 40
    tot = mp.mpf(0)
    tots = mp.mpf(0)
    for i in range(n):
        psi = ((mp.mpf(mu[0, i]) - mp.mpf(y[i])) ** 2 + mp.mpf(var[0, i])) / 2
        c = mp.sqrt(mp.mpf(mu[1, i]) ** 2 + mp.mpf(var[1, i]))
        sg = mp.exp(-mp.mpf(mu[1, i]) / 2) * mp.sech(c / 2) / 2
        tot += psi * (1 - sg)
        tots += 1 / (1 + mp.exp(-mp.mpf(f[1, i]))) / 2 * (mp.mpf(y[i]) - mp.mpf(f[0, i])) ** 2
    comp, seq = orc.hetero_lambda_stats(y, mu, var)
    assert comp == pytest.approx(float(tot), rel=1e-14) and seq == pytest.approx(float(tot), rel=1e-13)
    comps, _ = orc.hetero_lambda_stats_sampled(y, f)
    assert comps == pytest.approx(float(tots), rel=1e-14)
    # links
    nl = 4
    lt = rng.normal(0, 0.3, nl + 1)
    ff = rng.standard_normal((6, nl))
    sig = 1 / (1 + np.exp(-np.c_[ff, np.zeros(6)]))
    ref = np.exp(lt) * sig
    ref /= ref.sum(1, keepdims=True)
    rc, got = orc.logisticsoftmax(orc.make_lik(orc.CAT_BIJ, nlatent=nl, logtheta=lt), ff)
    assert rc == 0 and np.allclose(got, ref, rtol=1e-14, atol=0)
    rc, got = orc.logisticsoftmax(orc.make_lik(orc.CAT, nlatent=nl), ff)            # logisticsoftmax(x)
    s4 = sig[:, :nl]
    assert rc == 0 and np.allclose(got, s4 / s4.sum(1, keepdims=True), rtol=1e-14, atol=0)
    m = rng.standard_normal((6, nl))
    c = np.sqrt(m * m + 1.0)
    sg = np.exp(lt[:nl]) * np.exp(m / 2) / np.cosh(c / 2) / 2
    ref = sg / (np.exp(lt[nl]) * 0.5 + sg.sum(1, keepdims=True))
    rc, got = orc.approx_expected_logisticsoftmax(orc.make_lik(orc.CAT_BIJ, nlatent=nl, logtheta=lt), m, c)
    assert rc == 0 and np.allclose(got, ref, rtol=1e-13, atol=0)
