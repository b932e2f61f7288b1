"""AUG_LIK_FAITHFUL_QUIRKS: where the reference's CODE departs from the intended mathematics the library returns the
intended value by default and, with the flag, exactly what the code computes (SURVEY §9 "reference quirks" 4 and 6):

  Q4  logdensity_def(::PolyaGammaNegativeMultinomial, x) sums the PG log-densities over `1:length(x)` with x the
      2-field NamedTuple (ω, n), i.e. over the first two classes only
      (src/SpecialDistributions/polyagammanegativemultinomial.jl:33-39);
  Q6  kldivergence(::NegativeMultinomial, ::NegativeMultinomial) evaluates p * (log(p) - log(q)), which is
      0 * -Inf = NaN when a variational p_j is exactly 0 (src/SpecialDistributions/negativemultinomial.jl:80).

Tested both ways, on the oracle (CPU) and through the C ABI (GPU).
"""
import numpy as np
import pytest

from common import CAT_BIJ, synth_inputs

NL = 5
LOGTHETA = [0.2, -0.1, 0.4, 0.0, 0.3, -0.2]


def _sample_inputs(n=37, seed=3):
    rng = np.random.default_rng(seed)
    y, mu, var, f = synth_inputs(CAT_BIJ, n, seed, (), NL)
    nvar = rng.poisson(0.4, (n, NL)).astype(np.int64)
    omega = rng.gamma(2.0, 0.1, (n, NL)) * ((y + nvar) > 0)       # PG(0, .) is a point mass at 0
    omega = np.where((y + nvar) > 0, np.maximum(omega, 1e-3), 0.0)
    return y, mu, var, f, omega, nvar


def _extra_classes_logpdf(orc, y, omega, nvar):
    """sum over classes j >= 3 (index >= 2) of logpdf(PG(y_j + n_j, 0), ω_j): what the faithful value leaves out"""
    tot = 0.0
    for i in range(y.shape[0]):
        for j in range(2, y.shape[1]):
            b = float(y[i, j] + nvar[i, j])
            if b > 0:
                tot += float(orc.pg_logpdf(b, 0.0, np.array([omega[i, j]]))[0])
    return tot


def test_oracle_q4_two_class_logdensity(orc):
    y, mu, var, f, omega, nvar = _sample_inputs()
    # keep b > 0 everywhere so every class carries a PG log-density term
    nvar = np.maximum(nvar, 1)
    omega = np.maximum(omega, 0.05)
    intended = orc.make_lik(CAT_BIJ, nlatent=NL, logtheta=LOGTHETA)
    faithful = orc.make_lik(CAT_BIJ, nlatent=NL, logtheta=LOGTHETA, faithful_quirks=True)
    _, ci = orc.sampled_loglik_terms(intended, y, f, omega, nvar, True)
    _, cf = orc.sampled_loglik_terms(faithful, y, f, omega, nvar, True)
    assert cf[3] == ci[3]                                            # logtilt is not affected
    extra = _extra_classes_logpdf(orc, y, omega, nvar)
    assert abs(extra) > 1.0
    assert ci[4] - cf[4] == pytest.approx(extra, rel=1e-12)
    assert cf[5] == pytest.approx(cf[3] + cf[4], rel=1e-15)
    # a single latent: `x.ω[2]` is a BoundsError in the reference -> precondition
    one = orc.make_lik(CAT_BIJ, nlatent=1, logtheta=[0.0, 0.0], faithful_quirks=True)
    with pytest.raises(ArithmeticError):
        orc.sampled_loglik_terms(one, y[:, :1].copy(), f[:, :1].copy(), omega[:, :1].copy(), nvar[:, :1].copy(), True)


def test_oracle_q6_nan_at_saturated_p(orc):
    y, mu, var, f, omega, nvar = _sample_inputs(11)
    mu = mu.copy()
    mu[4, 2] = 800.0                                                 # -m < -744.44: approx_expected_logistic == 0 -> p == 0
    intended = orc.make_lik(CAT_BIJ, nlatent=NL, logtheta=LOGTHETA)
    faithful = orc.make_lik(CAT_BIJ, nlatent=NL, logtheta=LOGTHETA, faithful_quirks=True)
    rc, st, b, g, seq, ci = orc.cavi_step(intended, y, mu, var)
    assert rc == 0 and st[1][4, 2] == 0.0 and np.isfinite(ci[1])
    rc, st, b, g, seq, cf = orc.cavi_step(faithful, y, mu, var)
    assert rc == 0 and np.isnan(cf[1]) and np.isnan(cf[2])
    assert cf[0] == ci[0]                                            # expected_logtilt is not affected
    mu[4, 2] = 1.0                                                   # no saturated element: the flag changes nothing
    _, _, _, _, _, ci = orc.cavi_step(intended, y, mu, var)
    _, _, _, _, _, cf = orc.cavi_step(faithful, y, mu, var)
    assert np.array_equal(ci, cf)


@pytest.mark.gpu
def test_gpu_q4_and_q6_match_the_oracle_both_ways(orc):
    import torch
    from gpu_common import dev, host, pkg
    A = pkg()
    y, mu, var, f, omega, nvar = _sample_inputs(257, 9)
    Ω = A.AuxSamples(dev(omega), dev(nvar))
    for quirks in (False, True):
        lik = A.CategoricalLikelihood(LOGTHETA, faithful_quirks=quirks)
        olik = orc.make_lik(CAT_BIJ, nlatent=NL, logtheta=LOGTHETA, faithful_quirks=quirks)
        _, oc = orc.sampled_loglik_terms(olik, y, f, omega, nvar, True)
        assert A.aug_loglik(lik, Ω, dev(y), dev(f)) == pytest.approx(oc[5], rel=1e-11, abs=1e-10)
        assert A.aux_prior(lik, dev(y)).logdensity(Ω) == pytest.approx(oc[4], rel=1e-11, abs=1e-10)
        assert A.logtilt(lik, Ω, dev(y), dev(f)) == pytest.approx(oc[3], rel=1e-12)
    li = A.aug_loglik(A.CategoricalLikelihood(LOGTHETA), Ω, dev(y), dev(f))
    lf = A.aug_loglik(A.CategoricalLikelihood(LOGTHETA, faithful_quirks=True), Ω, dev(y), dev(f))
    assert abs(li - lf) > 1.0                                        # the two conventions really differ
    with pytest.raises(A.AugError) as ei:                            # single latent + quirks: BoundsError in the reference
        one = A.CategoricalLikelihood([0.0, 0.0], faithful_quirks=True)
        A.aug_loglik(one, A.AuxSamples(dev(omega[:, :1].copy()), dev(nvar[:, :1].copy())), dev(y[:, :1].copy()),
                     dev(f[:, :1].copy()))
    assert ei.value.rc == -3
    # Q6 on every CAVI route: tile kernels (large n) and the scalar tails (small n)
    for n in (5, 3000):
        y, mu, var, f = synth_inputs(CAT_BIJ, n, 21, (), NL)
        mu = mu.copy()
        mu[n // 2, 1] = 800.0
        for quirks in (False, True):
            lik = A.CategoricalLikelihood(LOGTHETA, faithful_quirks=quirks)
            q = A.init_aux_posterior(lik, n)
            q, b, g, s = A.cavi_step_(q, lik, dev(y), A.Normals(dev(mu), dev(var)))
            s = host(s)
            olik = orc.make_lik(CAT_BIJ, nlatent=NL, logtheta=LOGTHETA, faithful_quirks=quirks)
            rc, st, ob, og, seq, oc = orc.cavi_step(olik, y, mu, var)
            assert float(q.p[n // 2, 1]) == 0.0
            assert s[0] == pytest.approx(oc[0], rel=1e-12)
            if quirks:
                assert np.isnan(s[1]) and np.isnan(s[2]) and np.isnan(oc[1])
                assert np.isnan(A.aux_kldivergence(lik, q, dev(y), A.Normals(dev(mu), dev(var))))
            else:
                assert s[1] == pytest.approx(oc[1], rel=1e-12) and s[2] == pytest.approx(oc[2], rel=1e-12)
