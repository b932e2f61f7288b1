#!/usr/bin/env python
"""The Gibbs sampler of the reference's binary-classification example (examples/bernoulli/script.jl:76-87) in its sparse
form on the device.  Per sweep:  f = κᵀu  →  aux_sample!(Ω, lik, y, f)  →  P = K_Z⁻¹ + κ Diagonal(ω) κᵀ, rhs = κ β
(aug_sparse_precision_potential with the sampled ω and β = auglik_potential)  →  u ~ N(S·rhs, S), S = inv(P).

    python examples/sparse_bernoulli_gibbs.py [--n 20000] [--m 16] [--nsamples 300]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import aug_pkg  # noqa: E402
from sparse_bernoulli_cavi import cavi, make_problem  # noqa: E402


def gibbs(A, y, kappa, kdiag, KZ, KZinv, nsamples=300, seed=0):
    ctx = A.default_context()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.tdev)
    lik = A.BernoulliLikelihood()
    n, M = kappa.shape
    rng = np.random.default_rng(seed)
    prng = A.AugPhilox(seed + 1, 0)
    dy, dk, dP0 = dev(y), dev(kappa), dev(KZinv)
    Om = A.init_aux_variables(prng, lik, n)                               # Ω = init_aux_variables(lik, N)      script.jl:90
    u = np.linalg.cholesky(KZ) @ rng.standard_normal(M)
    us = []
    for _ in range(nsamples):
        f = dk @ dev(u)                                                   # f = κᵀu
        A.aux_sample_(prng, Om, lik, dy, f)                               # aux_sample!(Ω, lik, y, f)            script.jl:81
        beta, gamma = A.auglik_potential_and_precision(lik, Om, dy)       # auglik_potential / auglik_precision  :82-83
        P, rhs = A.sparse_precision_potential(dk, gamma[0], beta[0], P0=dP0)
        S = np.linalg.inv(P.cpu().numpy())
        S = 0.5 * (S + S.T)
        mu = S @ rhs.cpu().numpy()
        u = mu + np.linalg.cholesky(S) @ rng.standard_normal(M)           # rand!(MvNormal(μ, Σ), f)             :84
        us.append(u.copy())
    return np.array(us)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20_000)
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--nsamples", type=int, default=300)
    a = ap.parse_args()
    A = aug_pkg.load_package()
    ctx = A.Context(0)
    A.set_default_context(ctx)
    prob = make_problem(a.n, a.m)
    us = gibbs(A, *prob, nsamples=a.nsamples)
    m_cavi, S_cavi, _ = cavi(A, *prob, iters=10, verbose=False)
    burn = a.nsamples // 5
    print("Gibbs posterior mean of u:", np.round(us[burn:].mean(0)[:6], 3))
    print("CAVI  posterior mean of u:", np.round(m_cavi[:6], 3))
    ctx.close()
