"""SURVEY §8(f) rows 1 and 2 on the GPU against the oracle: the sparse-GP producer (marginals of the SVGP posterior,
examples/bernoulli/script.jl:32-33), the consumer (P = P0 + κ Diagonal(γ) κᵀ, rhs = r0 + κβ, docs/src/index.md:
154-163) and the fused one-pass CAVI sweep (producer → aux_posterior! → E[β], E[γ], ELBO sums → consumer).

Tolerances (fp64): per-observation arrays 1e-12 relative; sums over observations 1e-12 relative to the sum of the
absolute values of their terms (P_ij: sqrt(P_ii P_jj) bounds it by Cauchy-Schwarz) — the GPU adds them in a
different (tree / tensor-core) order than the oracle's long-double loop."""
import ctypes as C

import numpy as np
import pytest
import torch

import aug_pkg
from common import BERNOULLI, LAPLACE, NEGBIN, POISSON, STUDENTT, Y_DTYPE, relerr, synth_inputs, synth_sparse
from gpu_common import dev, host, make_lik

pytestmark = pytest.mark.gpu
RTOL = 1e-12

KINDS = [(BERNOULLI, (), {}), (NEGBIN, (10.0,), {"r_is_int": True}), (NEGBIN, (5.5,), {}), (POISSON, (10.0,), {}),
         (LAPLACE, (1.0,), {}), (STUDENTT, (3.0, 1.5), {})]


@pytest.fixture(scope="module")
def A():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    A = aug_pkg.load_package()
    ctx = A.Context(0)
    A.set_default_context(ctx)
    yield A
    A.set_default_context(None)
    ctx.close()


def check_P(P, rhs, oP, orhs, kappa, beta):
    d = np.sqrt(np.abs(np.diag(oP)))
    tolP = RTOL * np.maximum(np.outer(d, d), 1e-300)
    assert np.all(np.abs(P - oP) <= tolP), float(np.max(np.abs(P - oP) / tolP)) * RTOL
    tolr = RTOL * (np.abs(kappa) * np.abs(beta)[:, None]).sum(0)
    assert np.all(np.abs(rhs - orhs) <= tolr + 1e-300)
    assert np.array_equal(P, P.T)                       # mirrored from the lower triangle: exactly symmetric


@pytest.mark.parametrize("n,m", [(1, 1), (5, 8), (100, 16), (4097, 20), (3000, 32), (2049, 33), (5000, 64),
                                 (1500, 100), (4000, 128), (20000, 128), (3000, 129), (5000, 200), (2500, 256)])
def test_marginals_and_consumer_vs_oracle(A, orc, n, m):
    orc.set_threads(8)
    kappa, mvec, B, kdiag = synth_sparse(n, m, 1000 + m)
    rng = np.random.default_rng(n + m)
    gamma, beta = 0.25 * rng.random(n), rng.standard_normal(n)
    P0 = rng.standard_normal((m, m)); P0 = P0 @ P0.T / m; P0 = 0.5 * (P0 + P0.T)
    r0 = rng.standard_normal(m)
    omu, ovar = orc.sparse_marginals(kappa, mvec, B, kdiag)
    oP, orhs = orc.sparse_precision_potential(kappa, gamma, beta, P0, r0)
    q = A.sparse_marginals(dev(kappa), dev(mvec), dev(B), dev(kdiag))
    scale = np.abs(kappa) @ np.abs(mvec)
    assert np.all(np.abs(host(q.mu) - omu) <= RTOL * scale)
    assert relerr(host(q.var), ovar) <= RTOL
    P, rhs = A.sparse_precision_potential(dev(kappa), dev(gamma), dev(beta), dev(P0), dev(r0))
    check_P(host(P) - P0, host(rhs) - r0, oP - P0, orhs - r0, kappa, beta)
    np.testing.assert_allclose(host(P), oP, rtol=1e-11, atol=1e-13)
    orc.set_threads(1)


@pytest.mark.parametrize("kind,params,kw", KINDS)
@pytest.mark.parametrize("n,m", [(3001, 16), (2500, 48), (6000, 128), (3000, 160)])
def test_fused_sweep_vs_oracle(A, orc, kind, params, kw, n, m):
    orc.set_threads(8)
    kappa, mvec, B, kdiag = synth_sparse(n, m, 77 + m)
    y, _, _, _ = synth_inputs(kind, n, 5 + m, params)
    olik = orc.make_lik(kind, *params, **kw)
    rc, o = orc.sparse_cavi_sweep(olik, y, kappa, mvec, B, kdiag)
    assert rc == 0
    lik = make_lik(kind, params, kw)
    qO = A.init_aux_posterior(lik, n)
    P, rhs, scal, qf, bg = A.sparse_cavi_sweep_(qO, lik, dev(y), dev(kappa), dev(mvec), dev(B), dev(kdiag),
                                                want_elbo=True, want_marginals=True, want_potentials=True)
    torch.cuda.synchronize()
    assert relerr(host(qf.var), o["var"]) <= RTOL
    assert np.all(np.abs(host(qf.mu) - o["mu"]) <= RTOL * (np.abs(kappa) @ np.abs(mvec)))
    for i in range(3):
        s = qO._s(i)
        if s is not None and o["state"][i] is not None:
            if o["state"][i].dtype == np.float64:
                assert relerr(host(s), o["state"][i]) <= 4 * RTOL, i
            else:
                assert np.array_equal(host(s), o["state"][i])
    beta, gamma = host(bg[0]), host(bg[1])
    assert relerr(gamma, o["gamma"]) <= 4 * RTOL
    assert relerr(beta, o["beta"], floor=1.0) <= 4 * RTOL        # Poisson-type β = (y − λ̂)/2 cancels (see common.relerr)
    check_P(host(P), host(rhs), o["P"], o["rhs"], kappa, o["beta"])
    s = host(scal)
    for k in range(3):
        assert abs(s[k] - o["comp"][k]) <= 4 * RTOL * max(abs(o["comp"][k]), float(n)), (k, s[k], o["comp"][k])
    # the fused sweep equals the three verbs called one after the other (the reference's call pattern)
    q2 = A.sparse_marginals(dev(kappa), dev(mvec), dev(B), dev(kdiag))
    assert torch.equal(q2.mu, qf.mu) and torch.equal(q2.var, qf.var)
    P2, rhs2 = A.sparse_precision_potential(dev(kappa), bg[1], bg[0])
    assert torch.equal(P2, P) and torch.equal(rhs2, rhs)
    if m > 128:                                            # the cuBLAS composition: same contract, more launches
        assert A.default_context().lib is not None
    orc.set_threads(1)


def test_sweep_is_bit_reproducible_and_optional_outputs(A):
    n, m = 50_000, 64
    kappa, mvec, B, kdiag = synth_sparse(n, m, 3)
    y, _, _, _ = synth_inputs(BERNOULLI, n, 4)
    lik = A.BernoulliLikelihood()
    args = (lik, dev(y), dev(kappa), dev(mvec), dev(B), dev(kdiag))
    P1, r1, s1, _, _ = A.sparse_cavi_sweep_(None, *args)
    P2, r2, s2, _, _ = A.sparse_cavi_sweep_(A.init_aux_posterior(lik, n), *args, want_marginals=True)
    assert torch.equal(P1, P2) and torch.equal(r1, r2) and torch.equal(s1, s2)
    P3, r3, s3, _, _ = A.sparse_cavi_sweep_(None, *args, want_elbo=False)
    assert s3 is None and torch.equal(P1, P3)


def test_unaligned_kappa_and_empty_shard(A, orc):
    n, m = 777, 24
    kappa, mvec, B, kdiag = synth_sparse(n, m, 11)
    rng = np.random.default_rng(12)
    gamma, beta = rng.random(n), rng.standard_normal(n)
    oP, orhs = orc.sparse_precision_potential(kappa, gamma, beta)
    buf = dev(np.r_[0.0, kappa.ravel()])                 # κ at an odd element offset: 8-byte cp.async path
    ctx = A.default_context()
    Pr = torch.zeros(m * m + m, dtype=torch.float64, device="cuda")
    dg, db = dev(gamma), dev(beta)
    A.check(ctx.lib.aug_sparse_precision_potential(ctx.h, n, m, C.c_void_p(buf.data_ptr() + 8),
                                                   C.c_void_p(dg.data_ptr()), C.c_void_p(db.data_ptr()),
                                                   None, None, C.c_void_p(Pr.data_ptr())))
    ctx.sync()
    check_P(host(Pr[: m * m]).reshape(m, m), host(Pr[m * m:]), oP, orhs, kappa, beta)
    # n = 0 (an empty shard): P = P0, rhs = r0
    P0, r0 = dev(np.eye(m)), dev(np.arange(m, dtype=np.float64))
    P, rhs = A.sparse_precision_potential(torch.zeros((0, m), dtype=torch.float64, device="cuda"),
                                          torch.zeros(0, dtype=torch.float64, device="cuda"),
                                          torch.zeros(0, dtype=torch.float64, device="cuda"), P0, r0)
    assert torch.equal(P, P0) and torch.equal(rhs, r0)
    # argument errors mirror the ABI contract
    with pytest.raises(A.AugError):
        A.check(ctx.lib.aug_sparse_precision_potential(ctx.h, 10, 0, C.c_void_p(buf.data_ptr()), None, None, None,
                                                       None, C.c_void_p(Pr.data_ptr())))
    with pytest.raises(ValueError):
        A.sparse_cavi_sweep_(None, A.HeteroscedasticGaussianLikelihood(1.0), dev(np.zeros(4)), dev(np.zeros((4, 2))),
                             dev(np.zeros(2)), dev(np.zeros((2, 2))), dev(np.ones(4)))


def test_dense_precision_potential(A):
    n = 300
    rng = np.random.default_rng(0)
    Kinv = rng.standard_normal((n, n)); Kinv = Kinv @ Kinv.T
    gamma, beta, r0 = rng.random(n), rng.standard_normal(n), rng.standard_normal(n)
    P, rhs = A.dense_precision_potential(dev(Kinv), dev(gamma), dev(beta), dev(r0))
    assert np.array_equal(host(P), Kinv + np.diag(gamma))           # inv(K) + Diagonal(γ)  script.jl:35
    assert np.array_equal(host(rhs), beta + r0)                       # β + K \ mean(fz)       script.jl:36
    K2 = dev(Kinv)
    P2, _ = A.dense_precision_potential(K2, dev(gamma), dev(beta), out=K2)
    assert P2.data_ptr() == K2.data_ptr() and np.array_equal(host(K2), Kinv + np.diag(gamma))


def test_one_cavi_iteration_end_to_end(A, orc):
    """The user's loop of examples/bernoulli/script.jl:29-39 in its sparse form, one iteration on the device:
    sweep → S = inv(K_Z⁻¹ + κ Diagonal(γ) κᵀ), m = S (κ β)  vs the oracle's separate passes + numpy inverse."""
    n, m = 8000, 32
    kappa, mvec, B, kdiag = synth_sparse(n, m, 21)
    y, _, _, _ = synth_inputs(BERNOULLI, n, 22)
    KZinv = np.linalg.inv(B + 0.5 * np.eye(m))
    lik = A.BernoulliLikelihood()
    P, rhs, scal, _, _ = A.sparse_cavi_sweep_(None, lik, dev(y), dev(kappa), dev(mvec), dev(B), dev(kdiag), P0=dev(KZinv))
    S = torch.linalg.inv(P)
    mnew = S @ rhs
    rc, o = orc.sparse_cavi_sweep(orc.make_lik(orc.BERNOULLI), y, kappa, mvec, B, kdiag, np.ascontiguousarray(KZinv), None)
    So = np.linalg.inv(o["P"])
    np.testing.assert_allclose(host(S), So, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(host(mnew), So @ o["rhs"], rtol=1e-9, atol=1e-12)


def test_call_sequence_with_scratch_regrowth_and_second_context(A, orc):
    """Different M (fused kernels and the cuBLAS composition) and n back to back on one context — the scratch buffer is
    re-grown in between — interleaved with path verbs, then the same on a second context with its own stream."""
    orc.set_threads(8)
    lik = A.BernoulliLikelihood()
    olik = orc.make_lik(orc.BERNOULLI)
    ctx2 = A.Context(0)
    try:
        for rep, (n, m) in enumerate([(5000, 128), (70000, 16), (3000, 200), (9000, 64), (200, 8), (4000, 136), (5000, 128)]):
            kappa, mvec, B, kdiag = synth_sparse(n, m, 300 + rep)
            y, mu, var, _ = synth_inputs(BERNOULLI, n, 400 + rep)
            rc, o = orc.sparse_cavi_sweep(olik, y, kappa, mvec, B, kdiag)
            assert rc == 0
            for ctx in (None, ctx2):
                P, rhs, scal, _, _ = A.sparse_cavi_sweep_(None, lik, dev(y), dev(kappa), dev(mvec), dev(B), dev(kdiag), ctx=ctx)
                q = A.init_aux_posterior(lik, n, ctx=ctx)
                A.cavi_step_(q, lik, dev(y), A.Normals(dev(mu), dev(var)), ctx=ctx)      # a path verb in between
                check_P(host(P), host(rhs), o["P"], o["rhs"], kappa, o["beta"])
                assert abs(float(scal[2]) - o["comp"][2]) <= 4 * RTOL * max(abs(o["comp"][2]), float(n))
    finally:
        ctx2.close()
        orc.set_threads(1)


def test_cavi_loop_on_a_sparse_gp_never_decreases_the_elbo(A):
    """examples/sparse_bernoulli_cavi.py = the reference's cavi! loop (examples/bernoulli/script.jl:29-39) in sparse form,
    one aug_sparse_cavi_sweep per iteration.  Coordinate ascent must not decrease the augmented ELBO
    expected_logtilt − aux_kldivergence − KL(q(u) ‖ p(u)) (script.jl:65-70), and it converges in a few iterations."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "sparse_bernoulli_cavi", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples",
                                              "sparse_bernoulli_cavi.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    for n, m in [(20_000, 24), (60_000, 64)]:
        prob = ex.make_problem(n, m, seed=n)
        mu, S, elbos = ex.cavi(A, *prob, iters=10, verbose=False)
        d = np.diff(elbos)
        assert np.all(d >= -1e-7 * np.abs(elbos[:-1])), elbos
        assert d[0] > 0 and abs(d[-1]) <= 1e-6 * abs(elbos[-1]) + 1e-3          # it moved, then converged
        assert np.all(np.linalg.eigvalsh(S) > 0)


def test_gibbs_and_cavi_agree_on_the_sparse_posterior(A):
    """examples/sparse_bernoulli_gibbs.py = gibbs_sample of examples/bernoulli/script.jl:76-87 in sparse form (aux_sample! +
    the consumer verb with the sampled ω).  Both inference schemes of the reference target the same posterior over the
    inducing values: the Gibbs mean (fixed seeds) lies within a few posterior standard deviations / Monte-Carlo errors
    of the CAVI mean, and the Gibbs spread is not smaller than the (mean-field) CAVI spread by more than noise."""
    import importlib.util
    import os
    import sys
    exdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")
    sys.path.insert(0, exdir)
    try:
        spec = importlib.util.spec_from_file_location("sparse_bernoulli_gibbs", os.path.join(exdir, "sparse_bernoulli_gibbs.py"))
        ex = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ex)
    finally:
        sys.path.remove(exdir)
    prob = ex.make_problem(20_000, 12, seed=5)
    us = ex.gibbs(A, *prob, nsamples=400, seed=1)[100:]
    m_cavi, S_cavi, _ = ex.cavi(A, *prob, iters=10, verbose=False)
    sd_cavi = np.sqrt(np.diag(S_cavi))
    sd_gibbs = us.std(0)
    z = (us.mean(0) - m_cavi) / np.maximum(sd_gibbs, sd_cavi)
    assert np.all(np.abs(z) < 1.0), z                       # same posterior mean up to a fraction of its own spread
    assert np.all(sd_gibbs > 0.5 * sd_cavi) and np.all(sd_gibbs < 4.0 * sd_cavi), (sd_gibbs, sd_cavi)


@pytest.mark.parametrize("name", ["negbin_int", "negbin_real", "poisson", "laplace", "studentt"])
def test_cavi_loop_elbo_is_monotone_for_every_scalar_latent_likelihood(A, name):
    """The optimality property the reference's test battery leaves commented out (src/TestUtils.jl:166-190): with the
    auxiliary posterior at its optimum, coordinate ascent on (q(u), qΩ) must never decrease the augmented ELBO
    expected_logtilt − aux_kldivergence − KL(q(u) ‖ p(u)).  That holds only if E[β], E[γ], expected_logtilt and
    aux_kldivergence of a likelihood are mutually consistent — checked here for the other five scalar-latent kinds."""
    import importlib.util
    import os
    lik = {"negbin_int": A.NegativeBinomialLikelihood(10), "negbin_real": A.NegativeBinomialLikelihood(5.5),
           "poisson": A.PoissonLikelihood(10.0), "laplace": A.LaplaceLikelihood(1.0),
           "studentt": A.StudentTLikelihood(3.0, 1.5)}[name]
    spec = importlib.util.spec_from_file_location(
        "sparse_bernoulli_cavi", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples",
                                              "sparse_bernoulli_cavi.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    prob = ex.make_problem(30_000, 32, seed=3, lik=lik, gen_y=A.testutils.gen_y)
    mu, S, elbos = ex.cavi(A, *prob, iters=12, verbose=False, lik=lik)
    d = np.diff(elbos)
    assert np.all(np.isfinite(elbos))
    assert np.all(d >= -1e-7 * np.abs(elbos[:-1])), elbos
    assert d[0] > 0


def test_fused_sweep_vs_mpmath_golden(A):
    """The GPU sweep against tests/golden/golden_sparse.json (mpmath, 50 digits) directly — no oracle in between."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_sparse.json")) as fh:
        G = json.load(fh)
    lik = A.BernoulliLikelihood()
    for c in G["cases"]:
        N = lambda k, dt=np.float64: np.ascontiguousarray(c[k], dtype=dt)
        n = c["n"]
        q = A.init_aux_posterior(lik, n)
        P, rhs, scal, qf, bg = A.sparse_cavi_sweep_(q, lik, dev(N("y", np.uint8)), dev(N("kappa")), dev(N("mvec")), dev(N("B")),
                                                    dev(N("kdiag")), P0=dev(N("P0")), r0=dev(N("r0")),
                                                    want_marginals=True, want_potentials=True)
        assert relerr(host(qf.var), N("var")) <= RTOL and np.all(np.abs(host(qf.mu) - N("mu")) <= 1e-14)
        assert relerr(host(q.c), N("c")) <= RTOL and relerr(host(bg[1]), N("gamma")) <= RTOL
        assert np.array_equal(host(bg[0]), N("beta"))
        np.testing.assert_allclose(host(P), N("P"), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(host(rhs), N("rhs"), rtol=0, atol=1e-12)
        s = host(scal)
        assert s[0] == pytest.approx(c["expected_logtilt"], rel=1e-12) and s[1] == pytest.approx(c["kl"], rel=1e-11)
