#!/usr/bin/env python
"""Small invocations of every kernel family of the path for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aug_pkg  # noqa: E402
from common import BERNOULLI, CAT, CAT_BIJ, HETERO, LAPLACE, NEGBIN, POISSON, STUDENTT, synth_inputs  # noqa: E402
from gpu_common import make_lik  # noqa: E402

A = aug_pkg.load_package()
ctx = A.Context(0)
A.set_default_context(ctx)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
n = int(os.environ.get("SAN_N", "20000"))
if os.environ.get("SAN_CATGIBBS"):
    # the input ring of cat_gibbs_kernel (full / empty mbarriers) only wraps when a CTA gets more than three tiles:
    # SAN_CATGIBBS rows (e.g. 60000 -> ~10 tiles per CTA), odd rows dense (p0 small), even rows sparse
    nn = int(os.environ["SAN_CATGIBBS"])
    rng = np.random.default_rng(5)
    f = rng.standard_normal((nn, 99))
    f[1::2] += 5.0
    y = np.zeros((nn, 99), np.uint8)
    y[np.arange(nn), rng.integers(0, 99, nn)] = 1
    Om = A.aux_sample(A.AugPhilox(3, 0), make_lik(CAT_BIJ, (), dict(nlatent=99)), dev(y), dev(f))
    torch.cuda.synchronize()
    assert ctx.error_flag() == 0
    print("ok catgibbs", nn, float(Om.omega.sum()), int(Om.n.sum()), flush=True)
    ctx.close()
    sys.exit(0)
for kind, params, kw in [(BERNOULLI, (), {}), (NEGBIN, (10,), dict(r_is_int=True)), (NEGBIN, (5.5,), {}), (POISSON, (10.0,), {}),
                         (LAPLACE, (1.0,), {}), (STUDENTT, (3.0, 1.5), {}), (HETERO, (5.0,), dict(nlatent=2)),
                         (CAT_BIJ, (), dict(nlatent=99)), (CAT, (), dict(nlatent=10))]:
    nl = kw.get("nlatent", 1)
    nn = n // 10 if kind in (CAT, CAT_BIJ) else n
    y, mu, var, f = synth_inputs(kind, nn, 1, params, nl)
    lik = make_lik(kind, params, kw)
    q = A.init_aux_posterior(lik, nn)
    q, beta, gamma, scal = A.cavi_step_(q, lik, dev(y), A.Normals(dev(mu), dev(var)), want_elbo=(kind != CAT))
    if kind != CAT:
        A.expected_logtilt(lik, q, dev(y), A.Normals(dev(mu), dev(var)))
    A.expected_auglik_potential_and_precision(lik, q, dev(y), A.Normals(dev(mu), dev(var)))   # from-state verb (staged kernel)
    A.init_aux_variables(A.AugPhilox(4, 0), lik, nn)
    Om = A.aux_sample(A.AugPhilox(3, 0), lik, dev(y), dev(f))
    A.auglik_potential_and_precision(lik, Om, dev(y), dev(f))
    if kind != CAT:
        A.logtilt(lik, Om, dev(y), dev(f))
    torch.cuda.synchronize()
    print("ok", kind, nn, flush=True)
assert ctx.error_flag() == 0
ctx.close()
