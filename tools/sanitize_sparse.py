#!/usr/bin/env python
"""Small sparse-GP sweeps for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python tools/sanitize_sparse.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aug_pkg  # noqa: E402
from common import BERNOULLI, POISSON, synth_inputs, synth_sparse  # noqa: E402

A = aug_pkg.load_package()
ctx = A.Context(0)
A.set_default_context(ctx)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for m, n in [(16, 1500), (24, 700), (64, 900), (128, 40000), (100, 333)]:
    kappa, mvec, B, kdiag = synth_sparse(n, m, m)
    for kind, params, lik in [(BERNOULLI, (), A.BernoulliLikelihood()), (POISSON, (10.0,), A.PoissonLikelihood(10.0))]:
        y, _, _, _ = synth_inputs(kind, n, 3, params)
        q = A.init_aux_posterior(lik, n)
        P, rhs, sc, qf, bg = A.sparse_cavi_sweep_(q, lik, dev(y), dev(kappa), dev(mvec), dev(B), dev(kdiag),
                                                  want_marginals=True, want_potentials=True)
        q2 = A.sparse_marginals(dev(kappa), dev(mvec), dev(B), dev(kdiag))
        P2, r2 = A.sparse_precision_potential(dev(kappa), bg[1], bg[0])
        torch.cuda.synchronize()
        assert torch.isfinite(P).all() and torch.isfinite(sc[:3]).all()
    print("ok", m, n, flush=True)
ctx.close()
