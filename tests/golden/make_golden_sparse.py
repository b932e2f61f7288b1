#!/usr/bin/env python
"""Golden vectors for the sparse-GP rows, independent of the oracle and of the GPU code:
the documented formulas (docs/src/index.md:154-163, examples/bernoulli/script.jl:29-39) evaluated with mpmath at
50 digits on small seeded inputs, including the Bernoulli path in between (bernoulli.jl:17-65).

    python tests/golden/make_golden_sparse.py   ->  tests/golden/golden_sparse.json
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))


def case(n, m, seed):
    rng = np.random.default_rng(seed)
    kappa = rng.standard_normal((n, m)) / np.sqrt(m)
    A = rng.standard_normal((m, m))
    B = A @ A.T
    B *= 0.3 * m / np.trace(B)
    B = 0.5 * (B + B.T)
    kdiag = np.einsum("ti,ij,tj->t", kappa, B, kappa) + 0.3 + 0.5 * rng.random(n)
    mvec = rng.standard_normal(m)
    y = (rng.random(n) < 0.5).astype(int)
    P0 = rng.standard_normal((m, m))
    P0 = 0.5 * (P0 @ P0.T + (P0 @ P0.T).T)
    r0 = rng.standard_normal(m)
    F = lambda v: mp.mpf(float(v))
    K = [[F(kappa[t, i]) for i in range(m)] for t in range(n)]
    Bm = [[F(B[i, j]) for j in range(m)] for i in range(m)]
    mu = [sum(F(mvec[i]) * K[t][i] for i in range(m)) for t in range(n)]
    var = [F(kdiag[t]) - sum(K[t][i] * Bm[i][j] * K[t][j] for i in range(m) for j in range(m)) for t in range(n)]
    # Bernoulli path: c = sqrt(mu^2 + var), gamma = tanh(c/2)/(2c), beta = y - 1/2,
    # expected_logtilt = sum(-log 2 + beta mu - (mu^2 + var) gamma / 2), KL = sum(logcosh(c/2) - c^2 gamma / 2)
    c = [mp.sqrt(mu[t] ** 2 + var[t]) for t in range(n)]
    gamma = [mp.tanh(c[t] / 2) / (2 * c[t]) for t in range(n)]
    beta = [F(y[t]) - mp.mpf(1) / 2 for t in range(n)]
    elt = sum(-mp.log(2) + beta[t] * mu[t] - (mu[t] ** 2 + var[t]) * gamma[t] / 2 for t in range(n))
    kl = sum(mp.log(mp.cosh(c[t] / 2)) - c[t] ** 2 * gamma[t] / 2 for t in range(n))
    P = [[F(P0[i, j]) + sum(gamma[t] * K[t][i] * K[t][j] for t in range(n)) for j in range(m)] for i in range(m)]
    rhs = [F(r0[i]) + sum(beta[t] * K[t][i] for t in range(n)) for i in range(m)]
    f = lambda v: float(v)
    return dict(n=n, m=m, seed=seed, kappa=kappa.tolist(), B=B.tolist(), kdiag=kdiag.tolist(), mvec=mvec.tolist(),
                y=y.tolist(), P0=P0.tolist(), r0=r0.tolist(), mu=[f(v) for v in mu], var=[f(v) for v in var],
                c=[f(v) for v in c], gamma=[f(v) for v in gamma], beta=[f(v) for v in beta],
                expected_logtilt=f(elt), kl=f(kl), P=[[f(v) for v in row] for row in P], rhs=[f(v) for v in rhs])


if __name__ == "__main__":
    cases = [case(7, 3, 1), case(40, 8, 2), case(33, 20, 3), case(70, 33, 4)]
    with open(os.path.join(HERE, "golden_sparse.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden_sparse.py (mpmath, 50 digits)", "cases": cases}, fh)
    print("wrote", len(cases), "cases")
