"""ctypes front-end of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product path never
imports this module (see the header of aug_oracle.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BERNOULLI, NEGBIN, POISSON, LAPLACE, STUDENTT, HETERO, CAT_BIJ, CAT = range(8)
Y_DTYPE = {
    BERNOULLI: np.uint8, NEGBIN: np.int64, POISSON: np.int64, LAPLACE: np.float64,
    STUDENTT: np.float64, HETERO: np.float64, CAT_BIJ: np.uint8, CAT: np.uint8,
}


class Lik(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nlatent", C.c_int32), ("r_is_int", C.c_int32),
                ("flags", C.c_int32), ("p", C.c_double * 4), ("logtheta", C.c_void_p)]


def make_lik(kind, *params, nlatent=None, r_is_int=False, logtheta=None, faithful_quirks=False):
    if nlatent is None:
        nlatent = 2 if kind == HETERO else 1
    lik = Lik()
    lik.kind = kind
    lik.nlatent = nlatent
    lik.r_is_int = int(r_is_int)
    lik.flags = 1 if faithful_quirks else 0
    for i, v in enumerate(params):
        lik.p[i] = float(v)
    if logtheta is not None:
        arr = np.ascontiguousarray(logtheta, dtype=np.float64)
        lik._keep = arr
        lik.logtheta = arr.ctypes.data
    return lik


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        d = C.c_double
        for name, nargs in [("orc_second_moment", 2), ("orc_second_moment_y", 3),
                            ("orc_approx_expected_logistic", 2), ("orc_logistic", 1), ("orc_pg_mean", 2),
                            ("orc_pg_kl", 2), ("orc_pg_logpdf", 3), ("orc_digamma", 1), ("orc_normlogcdf", 1),
                            ("orc_mass_texpon", 1), ("orc_pg_var", 2), ("orc_kl_gamma", 4),
                            ("orc_kl_poisson", 2)]:
            f = getattr(L, name)
            f.restype = d
            f.argtypes = [d] * nargs
        L.orc_negbin_logconst.restype = d
        L.orc_negbin_logconst.argtypes = [d, d, C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def set_threads(t):
    lib().orc_set_threads(C.c_int(int(t)))


def max_threads():
    return int(lib().orc_get_max_threads())


def _state_shapes(lik, n):
    k, nl = lik.kind, lik.nlatent
    if k in (CAT, CAT_BIJ):
        return [((n, nl), np.float64), ((n, nl), np.float64), ((n, nl), np.uint8)]
    if k == BERNOULLI or k == LAPLACE or k == STUDENTT:
        return [((n,), np.float64), None, None]
    if k == NEGBIN:
        return [((n,), np.float64), None, ((n,), np.int64)]
    if k == POISSON:
        return [((n,), np.float64), ((n,), np.float64), ((n,), np.int64)]
    if k == HETERO:
        return [((n,), np.float64), ((n,), np.float64), ((n,), np.float64)]
    raise ValueError(k)


def alloc_state(lik, n, with_y_copy=True):
    out = []
    for i, s in enumerate(_state_shapes(lik, n)):
        if s is None or (i == 2 and not with_y_copy and lik.kind != HETERO):
            out.append(None)
        else:
            out.append(np.zeros(s[0], dtype=s[1]))
    return out


def _ld(lik, a):
    return a.shape[-1] if lik.kind == HETERO else 0


def aux_posterior(lik, y, mu, var, state):
    n = y.shape[0]
    rc = lib().orc_aux_posterior(C.byref(lik), C.c_int64(n), _p(y), _p(mu), _p(var), C.c_int64(_ld(lik, mu)),
                                 _p(state[0]), _p(state[1]), _p(state[2]))
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    return state


def expected_potential_precision(lik, y, mu, state):
    n = y.shape[0]
    nl = lik.nlatent
    beta = np.zeros((nl, n))
    gamma = np.zeros((nl, n))
    rc = lib().orc_expected_potential_precision(
        C.byref(lik), C.c_int64(n), _p(y), _p(mu), C.c_int64(_ld(lik, mu) if mu is not None else 0),
        _p(state[0]), _p(state[1]), _p(state[2]), _p(beta), _p(gamma), C.c_int64(n))
    return rc, beta, gamma


def expected_elbo_terms(lik, y, mu, var, state):
    n = y.shape[0]
    seq = np.zeros(8)
    comp = np.zeros(8)
    rc = lib().orc_expected_elbo_terms(C.byref(lik), C.c_int64(n), _p(y), _p(mu), _p(var),
                                       C.c_int64(_ld(lik, mu)), _p(state[0]), _p(state[1]), _p(state[2]),
                                       _p(seq), _p(comp))
    return rc, seq, comp


def cavi_step(lik, y, mu, var, with_y_copy=True, want_scalars=True):
    """Reference call pattern: aux_posterior! then E[β], E[γ] then the ELBO terms (separate passes)."""
    n = y.shape[0]
    nl = lik.nlatent
    state = alloc_state(lik, n, with_y_copy)
    beta = np.zeros((nl, n))
    gamma = np.zeros((nl, n))
    seq = np.zeros(8)
    comp = np.zeros(8)
    rc = lib().orc_cavi_step(C.byref(lik), C.c_int64(n), _p(y), _p(mu), _p(var), C.c_int64(_ld(lik, mu)),
                             _p(state[0]), _p(state[1]), _p(state[2]), _p(beta), _p(gamma), C.c_int64(n),
                             _p(seq) if want_scalars else None, _p(comp) if want_scalars else None)
    return rc, state, beta, gamma, seq, comp


def init_aux_variables(lik, seed, n):
    m = n * lik.nlatent if lik.kind in (CAT, CAT_BIJ) else n
    omega = np.zeros(m)
    nvar = np.zeros(m, dtype=np.int64)
    rc = lib().orc_init_aux_variables(C.byref(lik), C.c_uint64(seed), C.c_int64(n), _p(omega), _p(nvar))
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    return omega, nvar


def aux_sample(lik, seed, y, f):
    n = y.shape[0]
    m = n * lik.nlatent if lik.kind in (CAT, CAT_BIJ) else n
    omega = np.zeros(m)
    nvar = np.zeros(m, dtype=np.int64)
    rc = lib().orc_aux_sample(C.byref(lik), C.c_uint64(seed), C.c_int64(n), _p(y), _p(f),
                              C.c_int64(_ld(lik, f)), _p(omega), _p(nvar))
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    if lik.kind in (CAT, CAT_BIJ):
        omega = omega.reshape(n, lik.nlatent)
        nvar = nvar.reshape(n, lik.nlatent)
    return omega, nvar


def potential_precision(lik, y, f, omega, nvar):
    n = y.shape[0]
    nl = lik.nlatent
    beta = np.zeros((nl, n))
    gamma = np.zeros((nl, n))
    rc = lib().orc_potential_precision(C.byref(lik), C.c_int64(n), _p(y), _p(f),
                                       C.c_int64(_ld(lik, f) if f is not None else 0), _p(omega), _p(nvar),
                                       _p(beta), _p(gamma), C.c_int64(n))
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    return beta, gamma


def sampled_loglik_terms(lik, y, f, omega, nvar, with_prior=True):
    n = y.shape[0]
    seq = np.zeros(8)
    comp = np.zeros(8)
    rc = lib().orc_sampled_loglik_terms(C.byref(lik), C.c_int64(n), _p(y), _p(f), C.c_int64(_ld(lik, f)),
                                        _p(omega), _p(nvar), C.c_int(int(with_prior)), _p(seq), _p(comp))
    if rc == -3:
        raise ArithmeticError("precondition (oracle rc=-3)")
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    return seq, comp


def full_conditional_logdensity(lik, y, f, omega, nvar):
    n = y.shape[0]
    out = C.c_double(0)
    rc = lib().orc_full_conditional_logdensity(C.byref(lik), C.c_int64(n), _p(y), _p(f), C.c_int64(_ld(lik, f)),
                                               _p(omega), _p(nvar), C.byref(out))
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    return out.value


def pg_rand(seed, b, c, b_is_int):
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    out = np.zeros_like(b)
    lib().orc_pg_rand(C.c_uint64(seed), C.c_int64(b.size), _p(b), _p(c), C.c_int(int(b_is_int)), _p(out))
    return out


def pg_rand_bc(seed, n, b, c, b_is_int):
    out = np.zeros(n)
    lib().orc_pg_rand_bc(C.c_uint64(seed), C.c_int64(n), C.c_double(b), C.c_double(c), C.c_int(int(b_is_int)),
                         _p(out))
    return out


def pg_logpdf(b, c, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    lib().orc_pg_logpdf_vec(C.c_int64(x.size), C.c_double(b), C.c_double(c), _p(x), _p(out))
    return out


def pg_mean(b, c):
    return lib().orc_pg_mean(b, c)


def pg_var(b, c):
    return lib().orc_pg_var(b, c)


# ---- SURVEY §8(f) rows 3 and 4 ------------------------------------------------------------------
def hetero_lambda_stats(y, mu, var):
    """(compensated, sequential) value of dot(ψ, 1 .- σ̃g) — opt_lik, examples/heteroscedasticgaussian/script.jl:41-51.
    mu, var: [2][n] latent-major."""
    n = y.shape[0]
    out = np.zeros(2)
    rc = lib().orc_hetero_lambda_stats(C.c_int64(n), _p(np.ascontiguousarray(y, dtype=np.float64)),
                                       _p(np.ascontiguousarray(mu)), _p(np.ascontiguousarray(var)),
                                       C.c_int64(n), _p(out))
    assert rc == 0
    return out[0], out[1]


def hetero_lambda_stats_sampled(y, f):
    n = y.shape[0]
    out = np.zeros(2)
    rc = lib().orc_hetero_lambda_stats_sampled(C.c_int64(n), _p(np.ascontiguousarray(y, dtype=np.float64)),
                                               _p(np.ascontiguousarray(f)), C.c_int64(n), _p(out))
    assert rc == 0
    return out[0], out[1]


def logisticsoftmax(lik, f):
    n, nl = f.shape
    K = nl + 1 if lik.kind == CAT_BIJ else nl
    out = np.zeros((n, K))
    rc = lib().orc_logisticsoftmax(C.byref(lik), C.c_int64(n), _p(np.ascontiguousarray(f)), _p(out))
    return rc, out


def approx_expected_logisticsoftmax(lik, mu, c):
    n, nl = mu.shape
    out = np.zeros((n, nl))
    rc = lib().orc_approx_expected_logisticsoftmax(C.byref(lik), C.c_int64(n), _p(np.ascontiguousarray(mu)),
                                                   _p(np.ascontiguousarray(c)), _p(out))
    return rc, out


# ---- SURVEY §8(f) rows 1 and 2 (sparse-GP producer / consumer) ------------------------------------
def sparse_marginals(kappa, mvec, B, kdiag):
    n, m = kappa.shape
    mu, var = np.zeros(n), np.zeros(n)
    rc = lib().orc_sparse_marginals(C.c_int64(n), C.c_int(m), _p(np.ascontiguousarray(kappa)),
                                    _p(np.ascontiguousarray(mvec)), _p(np.ascontiguousarray(B)),
                                    _p(np.ascontiguousarray(kdiag)), _p(mu), _p(var))
    assert rc == 0
    return mu, var


def sparse_precision_potential(kappa, gamma, beta, P0=None, r0=None):
    n, m = kappa.shape
    Pr = np.zeros(m * m + m)
    rc = lib().orc_sparse_precision_potential(C.c_int64(n), C.c_int(m), _p(np.ascontiguousarray(kappa)),
                                              _p(np.ascontiguousarray(gamma)), _p(np.ascontiguousarray(beta)),
                                              _p(P0), _p(r0), _p(Pr))
    assert rc == 0
    return Pr[: m * m].reshape(m, m), Pr[m * m:]


def sparse_cavi_sweep(lik, y, kappa, mvec, B, kdiag, P0=None, r0=None, with_y_copy=True):
    n, m = kappa.shape
    mu, var = np.zeros(n), np.zeros(n)
    state = alloc_state(lik, n, with_y_copy)
    beta, gamma = np.zeros((1, n)), np.zeros((1, n))
    seq, comp = np.zeros(8), np.zeros(8)
    Pr = np.zeros(m * m + m)
    rc = lib().orc_sparse_cavi_sweep(C.byref(lik), C.c_int64(n), C.c_int(m), _p(y), _p(np.ascontiguousarray(kappa)),
                                     _p(np.ascontiguousarray(mvec)), _p(np.ascontiguousarray(B)),
                                     _p(np.ascontiguousarray(kdiag)), _p(mu), _p(var), _p(state[0]), _p(state[1]),
                                     _p(state[2]), _p(beta), _p(gamma), _p(P0), _p(r0), _p(Pr), _p(seq), _p(comp))
    return rc, dict(mu=mu, var=var, state=state, beta=beta[0], gamma=gamma[0], P=Pr[: m * m].reshape(m, m),
                    rhs=Pr[m * m:], seq=seq, comp=comp)
