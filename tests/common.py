"""Helpers shared by the CPU (oracle) and GPU (C-ABI) tests."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BERNOULLI, NEGBIN, POISSON, LAPLACE, STUDENTT, HETERO, CAT_BIJ, CAT = range(8)
Y_DTYPE = {0: np.uint8, 1: np.int64, 2: np.int64, 3: np.float64, 4: np.float64, 5: np.float64,
           6: np.uint8, 7: np.uint8}


def load_golden():
    with open(os.path.join(HERE, "golden", "golden_cavi.json")) as fh:
        return json.load(fh)


def golden_arrays(case):
    kind = case["kind"]
    y = np.ascontiguousarray(case["y"], dtype=Y_DTYPE[kind])
    mu = np.ascontiguousarray(case["mu"], dtype=np.float64)
    var = np.ascontiguousarray(case["var"], dtype=np.float64)
    return y, mu, var


def lik_args(case):
    """(kind, params, kwargs) understood by both oracle.orc.make_lik and the product's make_lik."""
    kw = dict(nlatent=case.get("nlatent", 2 if case["kind"] == HETERO else 1),
              r_is_int=bool(case.get("r_is_int", 0)))
    if "logtheta" in case:
        kw["logtheta"] = case["logtheta"]
    return case["kind"], case["params"], kw


def relerr(a, b, floor=1e-300):
    """max |a-b| / max(|b|, floor).  floor > 0 turns the bound into an absolute one for small |b|: used for
    beta = (y - E[n])/2 of the Poisson-type likelihoods, which cancels when E[n] ~ y (a conditioning effect:
    the 1e-12 bar then applies relative to the operands y, E[n] = O(1), not to their difference)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), floor)
    err = np.abs(a - b) / den
    err = np.where((a == b), 0.0, err)
    return float(np.max(err)) if err.size else 0.0


def synth_inputs(kind, n, seed, params=(), nlatent=1):
    """Synthetic inputs of SURVEY §8(d): mu ~ N(0,1), var = (0.5+U)^2, f ~ N(0,1), y from the likelihood."""
    rng = np.random.default_rng(seed)
    if kind in (CAT, CAT_BIJ):
        nl = nlatent
        K = nl + 1 if kind == CAT_BIJ else nl
        mu = rng.standard_normal((n, nl))
        var = (0.5 + rng.random((n, nl))) ** 2
        f = rng.standard_normal((n, nl))
        cls = rng.integers(0, K, n)
        y = np.zeros((n, nl), dtype=np.uint8)
        rows = np.nonzero(cls < nl)[0]
        y[rows, cls[rows]] = 1
        return y, mu, var, f
    if kind == HETERO:
        mu = rng.standard_normal((2, n))
        var = (0.5 + rng.random((2, n))) ** 2
        f = rng.standard_normal((2, n))
        y = f[0] + rng.standard_normal(n) / np.sqrt(params[0] / (1 + np.exp(-f[1])))
        return y, mu, var, f
    mu = rng.standard_normal(n)
    var = (0.5 + rng.random(n)) ** 2
    f = rng.standard_normal(n)
    sig = 1 / (1 + np.exp(-f))
    if kind == BERNOULLI:
        y = (rng.random(n) < sig).astype(np.uint8)
    elif kind == NEGBIN:
        r = params[0]
        y = rng.negative_binomial(r, 1 - np.clip(sig, 1e-9, 1 - 1e-9), n).astype(np.int64)
        y = np.minimum(y, 200)
    elif kind == POISSON:
        y = rng.poisson(params[0] * sig, n).astype(np.int64)
    elif kind == LAPLACE:
        y = f + rng.laplace(0, params[0], n)
    elif kind == STUDENTT:
        y = f + params[1] * rng.standard_t(params[0], n)
    else:
        raise ValueError(kind)
    return y, mu, var, f


def synth_sparse(n, m, seed):
    """Well-conditioned synthetic sparse-GP inputs for SURVEY §8(f) rows 1-2: κ ~ N(0, 1/m) ([n][m], the Julia
    M×N matrix as stored), B = A Aᵀ scaled so that κᵀBκ ≈ 0.3 (symmetric PSD, like K_Z − S at a CAVI fixed point),
    k_tt = κᵀBκ + 0.3 + U/2 (so σ² = k_tt − κᵀBκ lies in (0.3, 0.8) as for a valid SVGP posterior), m ~ N(0, 1)."""
    rng = np.random.default_rng(seed)
    kappa = rng.standard_normal((n, m)) / np.sqrt(m)
    A = rng.standard_normal((m, m))
    B = A @ A.T
    B *= 0.3 * m / np.trace(B)
    B = 0.5 * (B + B.T)
    kdiag = np.einsum("ti,ij,tj->t", kappa, B, kappa) + 0.3 + 0.5 * rng.random(n)
    mvec = rng.standard_normal(m)
    return kappa, mvec, B, kdiag
