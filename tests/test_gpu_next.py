"""SURVEY §8(f) rows 3 and 4 on the GPU against the oracle: the λ statistics of the heteroscedastic likelihood
(examples/heteroscedasticgaussian/script.jl:41-51; docs/src/likelihoods/heteroscedasticgaussian.md:80-84) and the
logistic-softmax links (likelihoods/categorical.jl:1-4, 32-35; utils.jl:17-22).  Tolerance 1e-12 relative."""
import numpy as np
import pytest
import torch

import aug_pkg
from common import CAT, CAT_BIJ, HETERO, relerr, synth_inputs
from gpu_common import dev, host

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("n", [1, 2, 1001, 262_145])
def test_hetero_lambda_stats_vs_oracle(A, orc, n):
    y, mu, var, f = synth_inputs(HETERO, n, 5, (5.0,), 2)
    got = A.hetero_lambda_stats(dev(y), A.Normals(dev(mu), dev(var)))
    comp, seq = orc.hetero_lambda_stats(y, mu, var)
    assert got == pytest.approx(comp, rel=1e-12)
    lik = A.opt_lik(A.HeteroscedasticGaussianLikelihood(3.0), A.Normals(dev(mu), dev(var)), dev(y))
    assert lik.lam == pytest.approx(max(n / (2 * comp), 3.0), rel=1e-12)
    gots = A.hetero_lambda_stats_sampled(dev(y), dev(f))
    comps, _ = orc.hetero_lambda_stats_sampled(y, f)
    assert gots == pytest.approx(comps, rel=1e-12)


def test_hetero_lambda_stats_extreme_and_unaligned(A, orc):
    n = 4097
    y, mu, var, f = synth_inputs(HETERO, n, 6, (5.0,), 2)
    mu[1, :8] = [0.0, 1e-9, -800.0, 800.0, 40.0, -40.0, 1e3, -1e3]      # saturation and out-of-range routes
    var[1, :8] = [0.0, 1e-20, 1.0, 1.0, 1e-4, 1e-4, 1e6, 1e6]
    comp, _ = orc.hetero_lambda_stats(y, mu, var)
    got = A.hetero_lambda_stats(dev(y), A.Normals(dev(mu), dev(var)))
    assert got == pytest.approx(comp, rel=1e-12)
    # views at odd element offsets take the scalar-load kernel
    Y, M, V = dev(np.r_[0.0, y]), dev(np.c_[np.zeros((2, 1)), mu]), dev(np.c_[np.zeros((2, 1)), var])
    M2, V2 = M[:, 1:], V[:, 1:]
    assert not M2.is_contiguous()
    # latent-major with a leading dimension: pass contiguous storage offset by one element through ld
    ctx = A.default_context()
    import ctypes as C
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    A.check(ctx.lib.aug_hetero_lambda_stats(ctx.h, n, C.c_void_p(Y.data_ptr() + 8), C.c_void_p(M.data_ptr() + 8),
                                            C.c_void_p(V.data_ptr() + 8), n + 1, C.c_void_p(out.data_ptr())))
    ctx.sync()
    assert float(out.item()) == pytest.approx(comp, rel=1e-12)


@pytest.mark.parametrize("nl,bij,n", [(1, True, 17), (3, True, 1000), (99, True, 513), (100, False, 300), (5, False, 64)])
def test_logisticsoftmax_vs_oracle(A, orc, nl, bij, n):
    rng = np.random.default_rng(nl)
    K = nl + 1 if bij else nl
    logtheta = rng.normal(0, 0.5, K)
    f = rng.standard_normal((n, nl)) * 3
    f[0, 0] = 800.0                                                    # logistic saturates at 1 (LogExpFunctions bounds)
    f[-1, -1] = -800.0                                                 # ... and at 0
    lik = A.CategoricalLikelihood(list(logtheta), bijective=bij)
    olik = orc.make_lik(CAT_BIJ if bij else CAT, nlatent=nl, logtheta=logtheta)
    rc, ref = orc.logisticsoftmax(olik, f)
    assert rc == 0
    got = host(A.logisticsoftmax(dev(f), lik))
    assert got.shape == (n, K)
    assert relerr(got, ref, floor=1e-300) < 1e-12
    assert np.allclose(got.sum(1), 1.0, atol=1e-14)
    if not bij:                                                        # logisticsoftmax(x), categorical.jl:1-4: logθ = 0
        olik0 = orc.make_lik(CAT, nlatent=nl)
        rc, ref0 = orc.logisticsoftmax(olik0, f[:3])
        g0 = host(A.logisticsoftmax(dev(f[:3])))
        assert relerr(g0, ref0) < 1e-12
        g1 = host(A.logisticsoftmax(dev(f[1])))                        # vector form
        assert relerr(g1, ref0[1]) < 1e-12


@pytest.mark.parametrize("nl,n", [(1, 9), (4, 1000), (99, 257)])
def test_approx_expected_logisticsoftmax_vs_oracle(A, orc, nl, n):
    rng = np.random.default_rng(100 + nl)
    logtheta = rng.normal(0, 0.5, nl + 1)
    mu = rng.standard_normal((n, nl)) * 2
    var = (0.5 + rng.random((n, nl))) ** 2
    c = np.sqrt(mu * mu + var)
    mu[0, 0], c[0, 0] = -800.0, 800.0                                  # saturates to 0 on μ alone (utils.jl:12-13)
    mu[1, 0], c[1, 0] = 40.0, 41.0                                     # ... and to 1
    lik = A.CategoricalLikelihood(list(logtheta), bijective=True)
    olik = orc.make_lik(CAT_BIJ, nlatent=nl, logtheta=logtheta)
    rc, ref = orc.approx_expected_logisticsoftmax(olik, mu, c)
    assert rc == 0
    got = host(A.approx_expected_logisticsoftmax(dev(mu), dev(c), lik))
    assert relerr(got, ref, floor=1e-300) < 1e-12
    with pytest.raises(TypeError):
        A.approx_expected_logisticsoftmax(dev(mu), dev(c), A.CategoricalLikelihood(nl, bijective=False))
