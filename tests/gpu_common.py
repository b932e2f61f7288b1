"""GPU-test helpers: build product likelihood objects from the shared case descriptions."""
import numpy as np
import torch

import aug_pkg
from common import BERNOULLI, CAT, CAT_BIJ, HETERO, LAPLACE, NEGBIN, POISSON, STUDENTT


def pkg():
    return aug_pkg.load_package()


def make_lik(kind, params, kw):
    A = pkg()
    if kind == BERNOULLI:
        return A.BernoulliLikelihood()
    if kind == NEGBIN:
        r = params[0]
        return A.NegativeBinomialLikelihood(int(r) if kw.get("r_is_int") else float(r))
    if kind == POISSON:
        return A.PoissonLikelihood(params[0])
    if kind == LAPLACE:
        return A.LaplaceLikelihood(params[0])
    if kind == STUDENTT:
        return A.StudentTLikelihood(params[0], params[1])
    if kind == HETERO:
        return A.HeteroscedasticGaussianLikelihood(params[0])
    lt = kw.get("logtheta")
    if lt is None:
        K = kw["nlatent"] + 1 if kind == CAT_BIJ else kw["nlatent"]
        lt = [0.0] * K
    return A.CategoricalLikelihood(list(lt), bijective=(kind == CAT_BIJ))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def stack(tup):
    return np.stack([host(t) for t in tup])
