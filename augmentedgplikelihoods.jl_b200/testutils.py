"""test_auglik — the reference's public test battery (src/TestUtils.jl:57-206, `AugmentedGPLikelihoods.TestUtils.test_auglik`)
for device containers: the checks a maintainer runs against a likelihood's augmentation, here against libaugcuda.

Same sections and invariants as the reference:
  Sampling                 init_aux_variables / aux_sample!(rng, Ω, …) / aux_sample; β, γ from the three accessor verbs agree,
                           γ >= 0; logtilt / aug_loglik are finite reals;
    Full conditional Ω     log C = aug_loglik(Ω) − log p(Ω | y, f) is the same for two draws of Ω (atol 1e-5) — for the
                           likelihoods whose full conditional is a Pólya-Gamma with a density the ABI exposes
                           (aug_pg_logpdf: Bernoulli, NegativeBinomial); the others are checked against the oracle in
                           tests/test_gpu_gibbs.py;
    Full conditional f     with a random prior covariance K: S = inv(K⁻¹ + Diagonal(γ)), m = S β, f₁, f₂ ~ N(m, S):
                           logtilt(Ω, f) + log p(f) − log q(f) is the same for both (atol 1e-5);
  Variational Inference    init_aux_posterior / aux_posterior! / aux_posterior; E[β], E[γ] from the three accessor verbs
                           agree, E[γ] >= 0; expected_logtilt, aux_kldivergence, expected_aug_loglik are finite reals
                           (and KL >= 0, which the reference's commented-out optimality check implies).
Raises AssertionError on the first violated check; returns a dict of the scalar values otherwise."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import api as A
from ._lib import BERNOULLI, CAT, CAT_BIJ, HETERO, LAPLACE, NEGBIN, POISSON, STUDENTT


def can_split(lik) -> bool:
    """src/generic.jl: every likelihood but the heteroscedastic one has an explicit prior p(Ω) and tilt."""
    return lik.kind != HETERO


def gen_y(rng: np.random.Generator, lik, f: np.ndarray):
    """y ~ lik(f) (TestUtils.jl:44-56).  f: [n] (scalar latent), [2][n] (heteroscedastic), [n][nl] (categorical)."""
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    k = lik.kind
    if k == BERNOULLI:
        return (rng.random(f.shape[0]) < sig(f)).astype(np.uint8)
    if k == NEGBIN:
        r = float(lik.failures)
        return rng.negative_binomial(r, np.clip(1.0 - sig(f), 1e-12, 1.0)).astype(np.int64)
    if k == POISSON:
        return rng.poisson(lik.lam * sig(f)).astype(np.int64)
    if k == LAPLACE:
        return f + rng.laplace(0.0, lik.beta, f.shape[0])
    if k == STUDENTT:
        return f + lik.sigma * rng.standard_t(lik.nu, f.shape[0])
    if k == HETERO:
        return f[0] + rng.standard_normal(f.shape[1]) / np.sqrt(lik.lam * sig(f[1]))
    nl = lik.nlatent
    K = nl + 1 if lik.bijective else nl
    ff = np.concatenate([f, np.zeros((f.shape[0], 1))], axis=1) if lik.bijective else f
    lt = np.asarray(lik.logtheta, dtype=np.float64)
    p = np.exp(lt)[None, :] * sig(ff)
    p /= p.sum(1, keepdims=True)
    cls = np.array([rng.choice(K, p=row) for row in p])
    y = np.zeros((f.shape[0], nl), dtype=np.uint8)
    rows = np.nonzero(cls < nl)[0]
    y[rows, cls[rows]] = 1
    return y


def _mvn_logpdf(x, mean, cov):
    d = x - mean
    L = np.linalg.cholesky(cov)
    z = np.linalg.solve(L, d)
    return -0.5 * (z @ z) - np.log(np.diag(L)).sum() - 0.5 * len(x) * math.log(2 * math.pi)


def test_auglik(lik, n: int = 10, seed: int = 0, ctx=None, atol: float = 1e-5):
    ctx = ctx or A.default_context()
    rng = np.random.default_rng(seed)
    nl = lik.nlatent
    cat, het = lik.kind in (CAT, CAT_BIJ), lik.kind == HETERO
    shape = (n, nl) if cat else ((2, n) if het else (n,))
    f = rng.standard_normal(shape)
    qmu, qvar = rng.standard_normal(shape), np.ones(shape)              # qf = Normal.(randn(n), 1.0)
    y = gen_y(rng, lik, f)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.tdev)
    host = lambda t: t.detach().cpu().numpy()
    dy, df, qf = dev(y), dev(f), A.Normals(dev(qmu), dev(qvar))
    prng = A.AugPhilox(seed + 1, 0)
    out = {}
    finite = lambda v: isinstance(v, float) and math.isfinite(v)

    # ------------------------------------------------------------------ Sampling
    Om = A.init_aux_variables(prng, lik, n, ctx=ctx)
    assert isinstance(Om, A.AuxSamples) and len(Om) == n
    Om = A.aux_sample_(prng, Om, lik, dy, df, ctx=ctx)
    assert isinstance(Om, A.AuxSamples) and len(Om) == n
    Om_new = A.aux_sample(prng, lik, dy, df, ctx=ctx)
    assert isinstance(Om_new, A.AuxSamples) and len(Om_new) == n
    ff = df if (het or cat) else None
    bs = A.auglik_potential(lik, Om, dy, ff if ff is not None else None, ctx=ctx)
    gs = A.auglik_precision(lik, Om, dy, ff if ff is not None else None, ctx=ctx)
    b2, g2 = A.auglik_potential_and_precision(lik, Om, dy, ff, ctx=ctx)
    assert len(bs) == len(gs) == nl and bs[0].dim() == 1 and gs[0].dim() == 1
    assert all(torch.allclose(a, b) for a, b in zip(bs, b2)) and all(torch.allclose(a, b) for a, b in zip(gs, g2))
    assert all(bool((g >= 0).all()) for g in gs), "the precision must be non-negative"
    if can_split(lik) and lik.kind != CAT:
        out["logtilt"] = A.logtilt(lik, Om, dy, df, ctx=ctx)
        assert finite(out["logtilt"])
    if lik.kind != CAT:                                    # the non-bijective prior NM(1, 1/nl) is improper
        out["aug_loglik"] = A.aug_loglik(lik, Om, dy, df, ctx=ctx)
        assert finite(out["aug_loglik"])
        if can_split(lik):                                 # aug_loglik = logtilt + logdensity(aux_prior(lik, y), Ω)  generic.jl:48-50
            pO = A.aux_prior(lik, dy)
            lp = A.logdensity_def(pO, Om, ctx=ctx)
            assert finite(lp) and len(pO) == n
            assert abs(out["logtilt"] + lp - out["aug_loglik"]) <= 1e-9 * max(1.0, abs(out["aug_loglik"]))
        pc = A.aux_full_conditional(lik, dy, df)           # two more draws through the distribution handle
        Om_a, Om_b = A.tvrand(prng, pc, ctx=ctx), A.tvrand(prng, pc, ctx=ctx)
        assert len(Om_a) == n and not torch.equal(Om_a.omega, Om_b.omega)

    # Full conditional Ω: C = p(f, y) = p(y | Ω, f) p(Ω) / p(Ω | y, f) does not depend on Ω
    if lik.kind in (BERNOULLI, NEGBIN):
        logC = []
        for Om_i in (Om, Om_new):
            lp = 0.0
            w = host(Om_i.omega)
            for i in range(n):
                b = 1.0 if lik.kind == BERNOULLI else float(y[i]) + float(lik.failures)
                lp += float(A.pg_logpdf(b, abs(float(f[i])), dev(w[i:i + 1]), ctx=ctx).item())
            logC.append(A.aug_loglik(lik, Om_i, dy, df, ctx=ctx) - lp)
        assert abs(logC[0] - logC[1]) <= atol, ("full conditional of Ω", logC)
        out["logC_omega"] = logC[0]

    # Full conditional f: q(f) = N(m, S) with S = inv(K⁻¹ + Diagonal(γ)), m = S β
    if not cat:
        Kr = rng.random((n, n))
        K = Kr @ Kr.T + 1e-6 * np.eye(n)
        Kinv = np.linalg.inv(K)
        S = [np.linalg.inv(Kinv + np.diag(host(g))) for g in gs]
        S = [0.5 * (s + s.T) for s in S]
        m = [s @ host(b) for s, b in zip(S, bs)]
        logC = []
        for _ in range(2):
            fj = [rng.multivariate_normal(mj, sj) for mj, sj in zip(m, S)]
            fnew = np.stack(fj) if het else fj[0]
            val = (A.aug_loglik if het else A.logtilt)(lik, Om, dy, dev(fnew), ctx=ctx)
            val += sum(_mvn_logpdf(x, np.zeros(n), K) for x in fj) - sum(_mvn_logpdf(x, mj, sj) for x, mj, sj in zip(fj, m, S))
            logC.append(val)
        if not het:       # the heteroscedastic precision of f depends on g (not a joint Gaussian): the reference skips it too
            assert abs(logC[0] - logC[1]) <= atol * max(1.0, abs(logC[0])), ("full conditional of f", logC)
        out["logC_f"] = logC[0]

    # ------------------------------------------------------------------ Variational inference
    q = A.init_aux_posterior(lik, n, ctx=ctx)
    assert isinstance(q, A.AuxPosterior) and len(q) == n
    q = A.aux_posterior_(q, lik, dy, qf, ctx=ctx)
    assert isinstance(q, A.AuxPosterior)
    q_new = A.aux_posterior(lik, dy, qf, ctx=ctx)
    assert isinstance(q_new, A.AuxPosterior) and len(q_new) == n
    eb = A.expected_auglik_potential(lik, q, dy, qf, ctx=ctx)
    eg = A.expected_auglik_precision(lik, q, dy, qf, ctx=ctx)
    eb2, eg2 = A.expected_auglik_potential_and_precision(lik, q, dy, qf, ctx=ctx)
    assert len(eb) == len(eg) == nl and eb[0].dim() == 1 and eg[0].dim() == 1
    assert all(torch.allclose(a, b) for a, b in zip(eb, eb2)) and all(torch.allclose(a, b) for a, b in zip(eg, eg2))
    assert all(bool((g >= 0).all()) for g in eg), "the expected precision must be non-negative"
    if can_split(lik) and lik.kind != CAT:
        out["expected_logtilt"] = A.expected_logtilt(lik, q, dy, qf, ctx=ctx)
        out["aux_kldivergence"] = A.aux_kldivergence(lik, q, dy, ctx=ctx)
        assert finite(out["expected_logtilt"]) and finite(out["aux_kldivergence"])
        assert out["aux_kldivergence"] >= -1e-9 * max(1.0, abs(out["aux_kldivergence"]))
    if lik.kind != CAT:
        out["expected_aug_loglik"] = A.expected_aug_loglik(lik, q, dy, qf, ctx=ctx)
        assert finite(out["expected_aug_loglik"])
    return out


test_auglik.__test__ = False        # a library function, not a pytest test
