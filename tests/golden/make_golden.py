"""Generate tests/golden/golden_cavi.json — 50-digit (mpmath) golden vectors.

The reference (Julia) cannot be executed in this image and ships no golden
vectors, so these fixtures are an INDEPENDENT third restatement of the formulas
(Python + mpmath at 50 digits, written from /root/reference/src/likelihoods/*.jl
and src/SpecialDistributions/*.jl; citations inline), rounded to float64 at the
end.  Both the C++ oracle (tests/test_oracle_pins.py, CPU) and the CUDA path
(tests/test_gpu_cavi.py, GPU) are compared against them.

Run:  python tests/golden/make_golden.py      (needs only numpy + mpmath)
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
LO, HI = -744.4400719213812, 36.7368005696771  # LogExpFunctions._logistic_bounds(Float64)
LN2 = mp.log(2)


def M(x):
    return mp.mpf(float(x))


def theta(b, c):  # mean(PolyaGamma(b,c)) polyagamma.jl:25-31
    return b / 4 if c == 0 else b / (2 * c) * mp.tanh(c / 2)


def klpg(b, c):  # polyagamma.jl:99-110
    return b * mp.log(mp.cosh(c / 2)) - c * c * theta(b, c) / 2


def sig_tilde(mu, c):  # approx_expected_logistic utils.jl:11-14
    if float(mu) < LO:
        return mp.mpf(0)
    if float(mu) > HI:
        return mp.mpf(1)
    return mp.exp(mu / 2) / mp.cosh(c / 2) / 2


def kl_poisson(lp, lq):
    if lp == 0:
        return lq
    return lq - lp + lp * mp.log(lp) - lp * mp.log(lq)


def kl_gamma(ap, tp, aq, tq):
    r = tp / tq
    return (ap - aq) * mp.psi(0, ap) - mp.loggamma(ap) + mp.loggamma(aq) - aq * mp.log(r) + ap * (r - 1)


def fl(x):
    return float(x)


def gen_inputs(rng, n, edge=True):
    mu = rng.standard_normal(n)
    var = (0.5 + rng.random(n)) ** 2
    if edge and n >= 8:
        mu[0], var[0] = 0.0, 0.0          # c == 0 branch (polyagamma.jl:26)
        mu[1], var[1] = 1e-9, 1e-18       # tiny c
        mu[2], var[2] = 800.0, 1.0        # -mu < lower bound: sigma~ saturates to 0
        mu[3], var[3] = -40.0, 4.0        # -mu > upper bound: sigma~ saturates to 1
        mu[4], var[4] = 30.0, 2500.0      # large c
        mu[5], var[5] = -1e-3, 1e-6
    return mu, var


def case_bernoulli(rng, n):
    mu, var = gen_inputs(rng, n)
    y = (rng.random(n) < 0.5).astype(np.uint8)
    c, beta, gamma, elt, kl = [], [], [], [], []
    for i in range(n):
        m, v = M(mu[i]), M(var[i])
        s2 = m * m + v
        ci = mp.sqrt(s2)
        sg = 1 if y[i] else -1
        th = theta(mp.mpf(1), ci)
        c.append(fl(ci)); beta.append(sg / 2); gamma.append(fl(th))
        elt.append(-LN2 + (sg * m - s2 * th) / 2)       # bernoulli.jl:62-64
        kl.append(klpg(mp.mpf(1), ci))                   # bernoulli.jl:57
    return dict(kind=0, params=[], y=y.tolist(), mu=mu.tolist(), var=var.tolist(), s0=c, beta=[beta],
                gamma=[gamma], elt=fl(mp.fsum(elt)), kl=fl(mp.fsum(kl)))


def case_negbin(rng, n, r, r_is_int):
    mu, var = gen_inputs(rng, n)
    y = rng.negative_binomial(max(r, 1), 0.5, n).astype(np.int64)
    c, beta, gamma, elt, kl = [], [], [], [], []
    R = M(r)
    for i in range(n):
        m, v, yi = M(mu[i]), M(var[i]), M(y[i])
        s2 = m * m + v
        ci = mp.sqrt(s2)
        th = theta(yi + R, ci)
        const = mp.loggamma(yi + R) - mp.loggamma(yi + 1) - mp.loggamma(R)  # negativebinomial.jl:51-52
        c.append(fl(ci)); beta.append(fl((yi - R) / 2)); gamma.append(fl(th))
        elt.append(const - (yi + R) * LN2 + (m * (yi - R) - s2 * th) / 2)    # :62-64
        kl.append(klpg(yi + R, ci))
    return dict(kind=1, params=[r], r_is_int=int(r_is_int), y=y.tolist(), mu=mu.tolist(), var=var.tolist(),
                s0=c, beta=[beta], gamma=[gamma], elt=fl(mp.fsum(elt)), kl=fl(mp.fsum(kl)))


def case_poisson(rng, n, lam):
    mu, var = gen_inputs(rng, n)
    y = rng.poisson(lam / 2, n).astype(np.int64)
    L = M(lam)
    c, lh, beta, gamma, elt, kl = [], [], [], [], [], []
    for i in range(n):
        m, v, yi = M(mu[i]), M(var[i]), M(y[i])
        s2 = m * m + v
        ci = mp.sqrt(s2)
        g = L * sig_tilde(-m, ci)                                            # poisson.jl:37
        th = theta(yi + g, ci)
        c.append(fl(ci)); lh.append(fl(g)); beta.append(fl((yi - g) / 2)); gamma.append(fl(th))
        elt.append(-(yi + g) * LN2 + ((yi - g) * m - s2 * th) / 2 + yi * mp.log(L) - mp.loggamma(yi + 1))
        kl.append(klpg(yi + g, ci) + kl_poisson(g, L))                       # polyagammapoisson.jl:47-51
    return dict(kind=2, params=[lam], y=y.tolist(), mu=mu.tolist(), var=var.tolist(), s0=c, s1=lh,
                beta=[beta], gamma=[gamma], elt=fl(mp.fsum(elt)), kl=fl(mp.fsum(kl)))


def case_laplace(rng, n, b):
    mu, var = gen_inputs(rng, n, edge=False)
    y = mu + rng.laplace(0, b, n)
    B = M(b)
    lamL = 1 / (2 * B) ** 2
    s0, beta, gamma, elt, kl = [], [], [], [], []
    for i in range(n):
        m, v, yi = M(mu[i]), M(var[i]), M(y[i])
        s2y = (m - yi) ** 2 + v
        mh = 1 / (2 * B * mp.sqrt(s2y))                                      # laplace.jl:48-50
        s0.append(fl(mh)); beta.append(fl(2 * mh * yi)); gamma.append(fl(2 * mh))
        elt.append(mp.loggamma(mp.mpf(1) / 2) - mp.log(mp.sqrt(mp.pi)) - mp.log(2 * B) - s2y * mh)
        kl.append(mp.log(2 * lamL) / 2 - mp.log(2 * mp.pi) / 2 - mp.log(lamL) / 2 +
                  mp.loggamma(mp.mpf(1) / 2) + lamL / mh)                    # laplace.jl:98-104
    return dict(kind=3, params=[b], y=y.tolist(), mu=mu.tolist(), var=var.tolist(), s0=s0, beta=[beta],
                gamma=[gamma], elt=fl(mp.fsum(elt)), kl=fl(mp.fsum(kl)))


def case_studentt(rng, n, nu, sig):
    mu, var = gen_inputs(rng, n, edge=False)
    y = mu + sig * rng.standard_t(nu, n)
    NU, SG = M(nu), M(sig)
    al = (NU + 1) / 2
    s0, beta, gamma, elt, kl = [], [], [], [], []
    for i in range(n):
        m, v, yi = M(mu[i]), M(var[i]), M(y[i])
        s2y = (m - yi) ** 2 + v
        bh = (NU / SG ** 2 + s2y) / 2                                        # studentt.jl:54-56
        th = al / bh
        s0.append(fl(bh)); beta.append(fl(th * yi)); gamma.append(fl(th))
        # logpdf(Normal(y, θ^-1/2), m) - vθ/2   studentt.jl:80-83
        elt.append(-mp.log(2 * mp.pi) / 2 + mp.log(th) / 2 - th * (m - yi) ** 2 / 2 - v * th / 2)
        kl.append(kl_gamma(al, 1 / bh, NU / 2, SG ** 2 / (NU / 2)))          # studentt.jl:91
    return dict(kind=4, params=[nu, sig], y=y.tolist(), mu=mu.tolist(), var=var.tolist(), s0=s0, beta=[beta],
                gamma=[gamma], elt=fl(mp.fsum(elt)), kl=fl(mp.fsum(kl)))


def case_hetero(rng, n, lam):
    muf, varf = gen_inputs(rng, n, edge=False)
    mug, varg = gen_inputs(rng, n)
    y = muf + rng.standard_normal(n)
    L = M(lam)
    C = (mp.log(L) + mp.log(2 / mp.pi)) / 2
    c, lh, psi, bf, bg, gf, gg, tot, a_, k_ = [], [], [], [], [], [], [], [], [], []
    for i in range(n):
        mf, vf, mg, vg, yi = M(muf[i]), M(varf[i]), M(mug[i]), M(varg[i]), M(y[i])
        ps = ((mf - yi) ** 2 + vf) / 2                                       # hetero :42
        ci = mp.sqrt(mg * mg + vg)                                           # :43
        st = sig_tilde(-mg, ci)
        g = L * st * ps                                                      # :44
        th = theta(mp.mpf(1) / 2 + g, ci)
        lsg = L * (1 - st)                                                   # :102
        c.append(fl(ci)); lh.append(fl(g)); psi.append(fl(ps))
        bf.append(fl(yi * lsg / 2)); bg.append(fl((mp.mpf(1) / 2 - g) / 2)); gf.append(fl(lsg)); gg.append(fl(th))
        a = C - (mp.mpf(1) / 2 + g) * LN2 + ((mp.mpf(1) / 2 - g) * mg - (mg * mg + vg) * th) / 2   # :139-140
        k = klpg(mp.mpf(1) / 2 + g, ci) + kl_poisson(g, L / 2 * ((yi - mf) ** 2 + vf))            # :141-143
        a_.append(a); k_.append(k); tot.append(a + k)
    return dict(kind=5, params=[lam], y=y.tolist(), mu=[muf.tolist(), mug.tolist()],
                var=[varf.tolist(), varg.tolist()], s0=c, s1=lh, s2=psi, beta=[bf, bg], gamma=[gf, gg],
                elt=fl(mp.fsum(a_)), kl=fl(mp.fsum(k_)), eall=fl(mp.fsum(tot)))


def case_cat(rng, n, K, bij, logtheta):
    nl = K - 1 if bij else K
    mu = rng.standard_normal((n, nl))
    var = (0.5 + rng.random((n, nl))) ** 2
    mu[0, 0], var[0, 0] = 0.0, 0.0
    mu[1, 1], var[1, 1] = 800.0, 1.0
    cls = rng.integers(0, K, n)
    y = np.zeros((n, nl), dtype=np.uint8)
    for i in range(n):
        if cls[i] < nl:
            y[i, cls[i]] = 1
    lt = [M(t) for t in logtheta]
    if bij:
        D = mp.exp(lt[K - 1]) / 2                                            # categorical.jl:12-14
        denom = D + nl                                                       # :92
        sum_theta = D + mp.fsum(mp.exp(t) for t in lt[:K - 1])               # :18-20
        prior_p = 1 / sum_theta                                              # :155
    else:
        denom = mp.mpf(nl)                                                   # :107
        prior_p = mp.mpf(1) / nl
    p0p = 1 - nl * prior_p
    c = np.zeros((n, nl)); p = np.zeros((n, nl)); beta = np.zeros((nl, n)); gamma = np.zeros((nl, n))
    elt, kl = [], []
    for i in range(n):
        cs, ps = [], []
        for j in range(nl):
            m, v = M(mu[i, j]), M(var[i, j])
            cij = mp.sqrt(m * m + v)
            cs.append(cij); ps.append(sig_tilde(-m, cij) / denom)
        p0 = 1 - mp.fsum(ps)
        e_sum, quad, kpg, knm = mp.mpf(0), mp.mpf(0), mp.mpf(0), mp.mpf(0)
        for j in range(nl):
            m, v, yi = M(mu[i, j]), M(var[i, j]), M(y[i, j])
            nbar = ps[j] / p0                                                # negativemultinomial.jl:54
            th = theta(yi + nbar, cs[j])
            c[i, j] = fl(cs[j]); p[i, j] = fl(ps[j])
            beta[j, i] = fl((yi - nbar) / 2); gamma[j, i] = fl(th)
            e_sum += yi + nbar
            quad += ((yi - nbar) * m - (m * m + v) * th) / 2
            kpg += klpg(yi + nbar, cs[j])
            if ps[j] > 0:
                knm += ps[j] * (mp.log(ps[j]) - mp.log(prior_p))
        elt.append(-e_sum * LN2 + quad)                                      # categorical.jl:176-179
        kl.append(kpg + mp.log(p0) - mp.log(p0p) + knm / p0)                 # nm.jl:78-81
    d = dict(kind=6 if bij else 7, params=[], nlatent=nl, logtheta=list(map(float, logtheta)), y=y.tolist(),
             mu=mu.tolist(), var=var.tolist(), s0=c.tolist(), s1=p.tolist(), beta=beta.tolist(),
             gamma=gamma.tolist())
    if bij:
        d.update(elt=fl(mp.fsum(elt)), kl=fl(mp.fsum(kl)))
    return d


def main():
    rng = np.random.default_rng(20261017)
    cases = {
        "bernoulli": case_bernoulli(rng, 48),
        "negbin_int10": case_negbin(rng, 48, 10, True),          # test/likelihoods/negativebinomial.jl:2
        "negbin_real5.5": case_negbin(rng, 48, 5.5, False),      # :3
        "poisson10": case_poisson(rng, 48, 10.0),                # test/likelihoods/poisson.jl:2
        "laplace1": case_laplace(rng, 48, 1.0),                  # test/likelihoods/laplace.jl:4
        "studentt3_1.5": case_studentt(rng, 48, 3.0, 1.5),       # test/likelihoods/studentt.jl:6
        "hetero5": case_hetero(rng, 48, 5.0),                    # test/likelihoods/heteroscedasticregression.jl:4
        "cat_bij_K3": case_cat(rng, 24, 3, True, [0.0, 0.0, 0.0]),   # test/likelihoods/categorical.jl:2-6
        "cat_bij_K5_theta": case_cat(rng, 24, 5, True, [0.3, -0.2, 0.1, 0.0, 0.5]),
        "cat_K3": case_cat(rng, 24, 3, False, [0.0, 0.0, 0.0]),
        "cat_bij_K100": case_cat(rng, 6, 100, True, [0.0] * 100),
    }
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_cavi.json")
    with open(out, "w") as fh:
        json.dump(cases, fh)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
