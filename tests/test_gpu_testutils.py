"""The reference's own per-likelihood tests (test/likelihoods/*.jl) are one line each: `test_auglik(lik)`.
Here the same battery (augmentedgplikelihoods.jl_b200/testutils.py mirrors src/TestUtils.jl:57-206) runs against libaugcuda for
every likelihood the reference tests — including the two it skips or does not include (Categorical: `@test_skip`,
HeteroscedasticGaussian: not in runtests.jl)."""
import pytest
import torch

import aug_pkg

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


def liks(A):
    return [
        ("bernoulli", A.BernoulliLikelihood()),                                   # test/likelihoods/bernoulli.jl:2
        ("negbin_int", A.NegativeBinomialLikelihood(10)),                          # negativebinomial.jl:2
        ("negbin_real", A.NegativeBinomialLikelihood(5.5)),                        # negativebinomial.jl:3
        ("poisson", A.PoissonLikelihood(10.0)),                                    # poisson.jl:2
        ("laplace", A.LaplaceLikelihood(1.0)),                                     # laplace.jl:4
        ("studentt", A.StudentTLikelihood(3.0, 1.5)),                              # studentt.jl:6
        ("hetero", A.HeteroscedasticGaussianLikelihood(5.0)),                      # heteroscedasticregression.jl:4
        ("cat_bij", A.CategoricalLikelihood([0.0, 0.0, 0.0], bijective=True)),     # categorical.jl:4-6 (@test_skip upstream)
        ("cat", A.CategoricalLikelihood([0.0, 0.0, 0.0], bijective=False)),        # categorical.jl:13
    ]


@pytest.mark.parametrize("idx", range(9))
@pytest.mark.parametrize("n,seed", [(10, 0), (257, 1)])
def test_auglik_battery(A, idx, n, seed):
    name, lik = liks(A)[idx]
    out = A.testutils.test_auglik(lik, n=n, seed=seed)
    assert isinstance(out, dict)
    if name not in ("cat",):
        assert "expected_aug_loglik" in out and "aug_loglik" in out
