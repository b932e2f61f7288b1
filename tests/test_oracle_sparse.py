"""Pins of the oracle's sparse-GP producer / consumer restatements.

The reference ships no code and no test for these steps (they are in the user's loop: examples/bernoulli/script.jl:
29-39, docs/src/index.md:154-163) — parity is unpinned against Julia output.  The pin is an independent evaluation of
the documented formulas: exact rational arithmetic (fractions.Fraction over the binary64 inputs) on small cases, and
numpy longdouble matrix algebra on larger ones."""
from fractions import Fraction

import numpy as np
import pytest

from common import BERNOULLI, synth_inputs, synth_sparse


def _exact(kappa, mvec, B, kdiag, gamma, beta):
    n, m = kappa.shape
    F = Fraction
    mu, var = [], []
    for t in range(n):
        k = [F(float(v)) for v in kappa[t]]
        mu.append(float(sum(F(float(mvec[i])) * k[i] for i in range(m))))
        q = sum(k[i] * F(float(B[i, j])) * k[j] for i in range(m) for j in range(m))
        var.append(float(F(float(kdiag[t])) - q))
    P = np.zeros((m, m))
    for i in range(m):
        for j in range(m):
            P[i, j] = float(sum(F(float(gamma[t])) * F(float(kappa[t, i])) * F(float(kappa[t, j])) for t in range(n)))
    rhs = np.array([float(sum(F(float(beta[t])) * F(float(kappa[t, i])) for t in range(n))) for i in range(m)])
    return np.array(mu), np.array(var), P, rhs


@pytest.mark.parametrize("n,m", [(1, 1), (7, 3), (40, 8), (33, 12)])
def test_oracle_sparse_vs_exact_rational(orc, n, m):
    kappa, mvec, B, kdiag = synth_sparse(n, m, 100 + n)
    rng = np.random.default_rng(n)
    gamma, beta = rng.random(n), rng.standard_normal(n)
    mu, var = orc.sparse_marginals(kappa, mvec, B, kdiag)
    P, rhs = orc.sparse_precision_potential(kappa, gamma, beta)
    emu, evar, eP, erhs = _exact(kappa, mvec, B, kdiag, gamma, beta)
    np.testing.assert_allclose(mu, emu, rtol=0, atol=4e-16 * np.abs(kappa).max() * np.abs(mvec).sum())
    np.testing.assert_allclose(var, evar, rtol=1e-15, atol=1e-16)
    np.testing.assert_allclose(P, eP, rtol=1e-15, atol=1e-18)
    np.testing.assert_allclose(rhs, erhs, rtol=0, atol=1e-15 * np.abs(beta).sum())
    assert np.array_equal(P, P.T)


def test_oracle_sparse_p0_r0_and_sweep_composition(orc):
    n, m = 2000, 24
    kappa, mvec, B, kdiag = synth_sparse(n, m, 7)
    y, _, _, _ = synth_inputs(BERNOULLI, n, 8)
    rng = np.random.default_rng(9)
    P0 = rng.standard_normal((m, m)); P0 = P0 @ P0.T
    r0 = rng.standard_normal(m)
    lik = orc.make_lik(orc.BERNOULLI)
    rc, out = orc.sparse_cavi_sweep(lik, y, kappa, mvec, B, kdiag, P0, r0)
    assert rc == 0
    # longdouble matrix algebra of the documented formulas
    kl = kappa.astype(np.longdouble)
    mu = kl @ mvec.astype(np.longdouble)
    var = kdiag.astype(np.longdouble) - np.einsum("ti,ij,tj->t", kl, B.astype(np.longdouble), kl)
    np.testing.assert_allclose(out["mu"], mu.astype(np.float64), rtol=0, atol=1e-15)
    np.testing.assert_allclose(out["var"], var.astype(np.float64), rtol=1e-14)
    # the path in between is the already-pinned cavi_step on those marginals
    rc2, st, ob, og, seq, comp = orc.cavi_step(lik, y, out["mu"], out["var"])
    assert rc2 == 0
    assert np.array_equal(out["gamma"], og[0]) and np.array_equal(out["beta"], ob[0])
    assert np.array_equal(out["comp"], comp)
    g = out["gamma"].astype(np.longdouble)
    P = (kl.T * g) @ kl + P0
    rhs = kl.T @ out["beta"].astype(np.longdouble) + r0
    np.testing.assert_allclose(out["P"], P.astype(np.float64), rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(out["rhs"], rhs.astype(np.float64), rtol=0, atol=1e-13)
    # S = inv(K_Z⁻¹ + κ Diagonal(γ) κᵀ) is what the user forms next: P must be symmetric positive definite
    assert np.array_equal(out["P"], out["P"].T)
    assert np.all(np.linalg.eigvalsh(out["P"]) > 0)


def test_oracle_sparse_sweep_vs_mpmath_golden(orc):
    """tests/golden/golden_sparse.json (mpmath, 50 digits; make_golden_sparse.py): the whole sweep of the oracle —
    marginals, Bernoulli path, P, rhs, ELBO sums — against an evaluation that shares no code with it."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_sparse.json")) as fh:
        G = json.load(fh)
    for c in G["cases"]:
        A = lambda k, dt=np.float64: np.ascontiguousarray(c[k], dtype=dt)
        rc, o = orc.sparse_cavi_sweep(orc.make_lik(orc.BERNOULLI), A("y", np.uint8), A("kappa"), A("mvec"), A("B"),
                                      A("kdiag"), A("P0"), A("r0"))
        assert rc == 0
        np.testing.assert_allclose(o["mu"], A("mu"), rtol=0, atol=2e-15)
        np.testing.assert_allclose(o["var"], A("var"), rtol=1e-14)
        np.testing.assert_allclose(o["state"][0], A("c"), rtol=1e-14)
        np.testing.assert_allclose(o["gamma"], A("gamma"), rtol=1e-14)
        assert np.array_equal(o["beta"], A("beta"))
        np.testing.assert_allclose(o["P"], A("P"), rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(o["rhs"], A("rhs"), rtol=0, atol=1e-13)
        assert o["comp"][0] == pytest.approx(c["expected_logtilt"], rel=1e-13)
        assert o["comp"][1] == pytest.approx(c["kl"], rel=1e-12)
