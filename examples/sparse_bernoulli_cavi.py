#!/usr/bin/env python
"""The reference's binary-classification example (examples/bernoulli/script.jl) in its sparse form
(docs/src/index.md:154-163) on the device: every CAVI iteration is ONE aug_sparse_cavi_sweep over κ = K_Z⁻¹ K_ZX —
marginals(post_u(x)) → aux_posterior! → E[β], E[γ], ELBO sums → P = K_Z⁻¹ + κ Diagonal(γ) κᵀ, rhs = κ β —
followed by the M×M solve S = inv(P), m = S·rhs, exactly the two lines of `cavi!` (script.jl:35-36).

    python examples/sparse_bernoulli_cavi.py [--n 200000] [--m 64] [--iters 8]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aug_pkg  # noqa: E402


def make_problem(n, m, seed=0, lengthscale=2.0, lik=None, gen_y=None):
    """x sorted in [-10, 10] (script.jl:12), SqExponential kernel with lengthscale 2 (:13), f ~ GP, y ~ Bernoulli(σ(f));
    inducing points on a grid.  Returns y, κ ([n][m], the Julia M×N matrix as stored), k_tt, K_Z, K_Z⁻¹."""
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(-10, 10, n))
    z = np.linspace(-10, 10, m)
    k = lambda a, b: np.exp(-0.5 * (a[:, None] - b[None, :]) ** 2 / lengthscale ** 2)
    KZ = k(z, z) + 1e-6 * np.eye(m)
    KZX = k(z, x)
    L = np.linalg.cholesky(KZ)
    kappa = np.linalg.solve(L.T, np.linalg.solve(L, KZX)).T.copy()       # [n][m]
    KZinv = np.linalg.inv(KZ)
    KZinv = 0.5 * (KZinv + KZinv.T)
    u = L @ rng.standard_normal(m)                                        # a draw of the inducing values
    f = kappa @ u
    if lik is None:
        y = (rng.random(n) < 1.0 / (1.0 + np.exp(-f))).astype(np.uint8)
    else:                                                                 # any scalar-latent likelihood: y ~ lik(f)
        y = gen_y(rng, lik, f)
    kdiag = np.ones(n) + 1e-6
    return y, kappa, kdiag, KZ, KZinv


def kl_mvn(m, S, K, Kinv):
    """KL(N(m, S) ‖ N(0, K)) — kldivergence(u_post.approx.q, u_post.approx.fz), script.jl:69"""
    M = len(m)
    return 0.5 * (np.trace(Kinv @ S) + m @ Kinv @ m - M + np.linalg.slogdet(K)[1] - np.linalg.slogdet(S)[1])


def cavi(A, y, kappa, kdiag, KZ, KZinv, iters=8, verbose=True, lik=None):
    ctx = A.default_context()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.tdev)
    lik = lik or A.BernoulliLikelihood()
    n, M = kappa.shape
    dy, dk, dkd, dP0 = dev(y), dev(kappa), dev(kdiag), dev(KZinv)
    q = A.init_aux_posterior(lik, n)                                      # qΩ = init_aux_posterior(lik, N)   script.jl:43
    m, S = np.zeros(M), KZ.copy()                                         # q(u) = p(u) to start with
    elbos = []
    for it in range(iters):
        B = KZ - S
        P, rhs, scal, _, _ = A.sparse_cavi_sweep_(q, lik, dy, dk, dev(m), dev(0.5 * (B + B.T)), dkd, P0=dP0)
        s = scal.cpu().numpy()
        elbo = s[0] - s[1] - kl_mvn(m, S, KZ, KZinv)                     # aug_elbo of script.jl:65-70 at the CURRENT q(u), qΩ*(q(u))
        elbos.append(elbo)
        S = np.linalg.inv(P.cpu().numpy())                               # S .= inv(K⁻¹ + κ Diagonal(γ) κᵀ)    script.jl:35
        S = 0.5 * (S + S.T)
        m = S @ rhs.cpu().numpy()                                        # m .= S (κ β + K⁻¹ μ₀), μ₀ = 0        script.jl:36
        if verbose:
            print(f"iter {it}: augmented ELBO = {elbo:.6f}")
    return m, S, elbos


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--iters", type=int, default=8)
    a = ap.parse_args()
    A = aug_pkg.load_package()
    ctx = A.Context(0)
    A.set_default_context(ctx)
    prob = make_problem(a.n, a.m)
    m, S, elbos = cavi(A, *prob, iters=a.iters)
    assert all(b >= a_ - 1e-6 * abs(a_) for a_, b in zip(elbos, elbos[1:])), "CAVI must not decrease the ELBO"
    print("posterior mean of u (first 5):", np.round(m[:5], 4))
    ctx.close()
