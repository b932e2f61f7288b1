"""Host-side mirror of the AugmentedGPLikelihoods.jl interface for the augmentation hot path.

Same verbs, argument order and meaning as the reference (src/AugmentedGPLikelihoods.jl:18-37);
Julia's `!` suffix becomes a trailing underscore.  Arrays are torch CUDA tensors (device memory
and streams are plumbing); every verb is ONE call into libaugcuda.so through the C ABI of
include/augcuda.h.  There is no CPU path: without the shared library or a CUDA device the
calls raise.

Containers (reference semantics in parentheses):
  Normals(mu, var)          qf::AbstractVector{<:Normal}: only mean/var are ever read (utils.jl:1-7).
                            HETERO: mu/var of shape [2, n]; CAT: [n, nl] (categorical.jl:84).
  AuxSamples(omega, n)      Ω::TupleVector with fields ω (and n)  (bernoulli.jl:3-5, poisson.jl:14-18)
  AuxPosterior(lik, fields) qΩ::For whose `only(qΩ.inds)` is the SoA of variational parameters
                            (fields c / y / λ / μ / β / ψ / p exactly as in the reference).
Results of the (expected_)auglik_* verbs are tuples with one length-n vector per latent
(class-major for the Categorical likelihood, categorical.jl:112-136).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib
from ._lib import (BERNOULLI, CAT, CAT_BIJ, HETERO, LAPLACE, NEGBIN, NSCALARS, POISSON, STUDENTT, AugError,
                   AugLik, check)


# ----------------------------------------------------------------------------- likelihoods
class Likelihood:
    kind: int = -1
    nlatent: int = 1

    def _desc(self) -> AugLik:
        raise NotImplementedError

    def _mk(self, *params, r_is_int=False, logtheta=None) -> AugLik:
        d = AugLik()
        d.kind = self.kind
        d.nlatent = self.nlatent
        d.r_is_int = int(r_is_int)
        d.flags = 1 if getattr(self, "faithful_quirks", False) else 0   # AUG_LIK_FAITHFUL_QUIRKS
        for i, v in enumerate(params):
            d.p[i] = float(v)
        if logtheta is not None:
            self._lt = (C.c_double * len(logtheta))(*[float(t) for t in logtheta])
            d.logtheta = C.cast(self._lt, C.c_void_p).value
        return d


class BernoulliLikelihood(Likelihood):
    """BernoulliLikelihood(LogisticLink()) — likelihoods/bernoulli.jl"""
    kind = BERNOULLI

    def _desc(self):
        return self._mk()


class NegativeBinomialLikelihood(Likelihood):
    """NegativeBinomialLikelihood(NBParamFailure(r), LogisticLink()) — likelihoods/negativebinomial.jl.
    An `int` r selects the Int methods (binomial log-constant, integer-b PG sampler)."""
    kind = NEGBIN

    def __init__(self, failures):
        self.failures = failures

    def _desc(self):
        return self._mk(self.failures, r_is_int=isinstance(self.failures, int))


class PoissonLikelihood(Likelihood):
    """PoissonLikelihood(ScaledLogistic(λ)) — likelihoods/poisson.jl"""
    kind = POISSON

    def __init__(self, lam):
        self.lam = float(lam)

    def _desc(self):
        return self._mk(self.lam)


class LaplaceLikelihood(Likelihood):
    """LaplaceLikelihood(β) — likelihoods/laplace.jl:13-17"""
    kind = LAPLACE

    def __init__(self, beta=1.0):
        self.beta = float(beta)

    def _desc(self):
        return self._mk(self.beta)


class StudentTLikelihood(Likelihood):
    """StudentTLikelihood(ν, σ) — likelihoods/studentt.jl:14-21"""
    kind = STUDENTT

    def __init__(self, nu, sigma):
        self.nu, self.sigma = float(nu), float(sigma)

    def _desc(self):
        return self._mk(self.nu, self.sigma)


class HeteroscedasticGaussianLikelihood(Likelihood):
    """HeteroscedasticGaussianLikelihood(InvScaledLogistic(λ)) — likelihoods/heteroscedasticgaussian.jl"""
    kind = HETERO
    nlatent = 2

    def __init__(self, lam):
        self.lam = float(lam)

    def _desc(self):
        return self._mk(self.lam)


class CategoricalLikelihood(Likelihood):
    """CategoricalLikelihood(BijectiveSimplexLink(LogisticSoftMaxLink(logθ))) when bijective (default) or
    CategoricalLikelihood(LogisticSoftMaxLink(logθ)) — likelihoods/categorical.jl:6-47.
    `logtheta` may be an int (number of classes -> zeros, categorical.jl:10)."""

    def __init__(self, logtheta, bijective=True, faithful_quirks=False):
        # faithful_quirks: return what the reference's CODE returns where it departs from the intended formulas
        # (2-class PG log-density sum, NaN for p log p at p = 0) — include/augcuda.h: AUG_LIK_FAITHFUL_QUIRKS
        self.faithful_quirks = bool(faithful_quirks)
        if isinstance(logtheta, int):
            logtheta = [0.0] * logtheta
        self.logtheta = [float(t) for t in logtheta]
        self.bijective = bool(bijective)
        self.kind = CAT_BIJ if bijective else CAT
        self.nlatent = len(self.logtheta) - 1 if bijective else len(self.logtheta)

    def _desc(self):
        return self._mk(logtheta=self.logtheta)


def nlatent(lik: Likelihood) -> int:
    """nlatent(lik) — generic.jl:87, heteroscedasticgaussian.jl:11, categorical.jl:46-47"""
    return lik.nlatent


_Y_DTYPE = {BERNOULLI: torch.uint8, NEGBIN: torch.int64, POISSON: torch.int64, LAPLACE: torch.float64,
            STUDENTT: torch.float64, HETERO: torch.float64, CAT_BIJ: torch.uint8, CAT: torch.uint8}
# state field names per likelihood, in the (s0, s1, s2) order of the C ABI
_FIELDS = {BERNOULLI: ("c", None, None), NEGBIN: ("c", None, "y"), POISSON: ("c", "λ", "y"),
           LAPLACE: ("μ", None, None), STUDENTT: ("β", None, None), HETERO: ("c", "λ", "ψ"),
           CAT_BIJ: ("c", "p", "y"), CAT: ("c", "p", "y")}


def _is_cat(lik):
    return lik.kind in (CAT, CAT_BIJ)


# ----------------------------------------------------------------------------- containers
@dataclass
class Normals:
    mu: torch.Tensor
    var: torch.Tensor

    def __post_init__(self):
        self.mu, self.var = _t(self.mu), _t(self.var)

    def __len__(self):
        return self.mu.shape[-1] if self.mu.dim() == 1 else self.mu.shape[0]


@dataclass
class AuxSamples:
    omega: torch.Tensor
    n: Optional[torch.Tensor] = None

    def __post_init__(self):
        self.omega, self.n = _t(self.omega), _t(self.n)

    @property
    def ω(self):
        return self.omega

    def __len__(self):
        return self.omega.shape[0]


@dataclass
class AuxPosterior:
    lik: Likelihood
    fields: dict = field(default_factory=dict)

    def __getattr__(self, name):
        f = self.__dict__.get("fields", {})
        if name in f:
            return f[name]
        raise AttributeError(name)

    def __len__(self):
        return next(iter(self.fields.values())).shape[0]

    def _s(self, i):
        name = _FIELDS[self.lik.kind][i]
        return self.fields.get(name) if name else None


class AugPhilox:
    """rng argument of the sampling verbs: a counter-based Philox4x32-10 stream (seed, offset).
    Each sampling verb consumes one offset tick (aug_ctx_seed / aug_ctx_get_offset)."""

    def __init__(self, seed=0, offset=0):
        self.seed, self.offset = int(seed), int(offset)


# ----------------------------------------------------------------------------- context
class Context:
    """One per process/GPU: owns the aug_ctx and the CUDA stream every verb runs on."""

    def __init__(self, device: Optional[int] = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("libaugcuda needs a CUDA device (no CPU fallback)")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tdev = torch.device("cuda", self.device)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream()
        h = C.c_void_p()
        check(self.lib.aug_ctx_create(C.byref(h), self.device, C.c_void_p(self.stream.cuda_stream)))
        self.h = h
        self.comm_ready = False     # NCCL communicator attached (dist.init_comm)
        self.p2p_ready = False      # peer-memory mailbox attached (dist.init_p2p)
        self.fused = False          # scalar-producing verbs already return sums over all ranks
        self.deferred = False       # split-phase exchange: scalars are complete after flush() / sync() / the next aux_sample_

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.aug_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seed(self, seed, offset=0):
        check(self.lib.aug_ctx_seed(self.h, C.c_uint64(seed), C.c_uint64(offset)))

    def offset(self):
        o = C.c_uint64()
        check(self.lib.aug_ctx_get_offset(self.h, C.byref(o)))
        return o.value

    def sync(self):
        check(self.lib.aug_ctx_sync(self.h))

    def flush(self):
        """complete a pending split-phase exchange (dist.set_deferred) on the ctx stream"""
        self.enter()
        check(self.lib.aug_comm_flush(self.h))
        self.leave()

    def launches(self):
        n = C.c_uint64()
        check(self.lib.aug_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def sm_count(self):
        n = C.c_int32()
        check(self.lib.aug_ctx_sm_count(self.h, C.byref(n)))
        return n.value

    def error_flag(self):
        f = C.c_uint32()
        check(self.lib.aug_ctx_error_flag(self.h, C.byref(f)))
        return f.value

    # stream hand-off with torch: our stream waits for the caller's pending work and vice versa
    def enter(self):
        self.stream.wait_stream(torch.cuda.current_stream(self.tdev))

    def leave(self):
        torch.cuda.current_stream(self.tdev).wait_stream(self.stream)

    def empty(self, shape, dtype=torch.float64):
        # allocated on the CALLER's current stream (the ctx stream is ordered against it by enter/leave)
        return torch.empty(shape, dtype=dtype, device=self.tdev)


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def set_default_context(ctx: Optional[Context]):
    global _default_ctx
    _default_ctx = ctx


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("device tensor expected")
    if not t.is_contiguous():
        raise ValueError("contiguous tensor expected")
    return C.c_void_p(t.data_ptr())


# ---- host-array dispatch.  The reference's verbs take plain host Vectors (src/generic.jl:1-88); here a verb whose
# array arguments are HOST tensors (torch CPU tensors or numpy arrays) goes through the *_host entry points of the
# C ABI (chunked H2D / kernel / D2H pipeline inside the library) and returns host tensors; device tensors go through
# the device-pointer entry points.  Mixing the two in one call is an error.  Either way the arithmetic runs on the GPU.
def _t(a):
    """numpy array -> torch view of the same memory; tensors pass through"""
    if a is None or torch.is_tensor(a):
        return a
    return torch.as_tensor(a)


def _is_host(*ts) -> bool:
    flags = {bool(t.is_cuda) for t in ts if t is not None}
    if len(flags) > 1:
        raise ValueError("host and device arrays mixed in one call")
    return flags == {False}


def _hptr(t):
    if t is None:
        return None
    if t.is_cuda:
        raise ValueError("host tensor expected")
    if not t.is_contiguous():
        raise ValueError("contiguous tensor expected")
    return C.c_void_p(t.data_ptr())


def _hempty(shape, dtype=torch.float64, pin=True):
    """host output buffer (pinned: the D2H copies of the staging pipeline then run at full PCIe speed)"""
    try:
        return torch.empty(shape, dtype=dtype, pin_memory=pin)
    except RuntimeError:
        return torch.empty(shape, dtype=dtype)


def _hscal():
    return (C.c_double * NSCALARS)()


def _scal_tensor(hs):
    return torch.tensor(list(hs), dtype=torch.float64)


def _check_y(lik, y):
    want = _Y_DTYPE[lik.kind]
    if y.dtype == torch.bool and want == torch.uint8:
        y = y.view(torch.uint8)
    if y.dtype != want:
        raise TypeError(f"y must be {want} for this likelihood, got {y.dtype}")
    return y


def _nobs(lik, y):
    return y.shape[0]


def _ld(lik, t):
    return t.shape[-1] if (lik.kind == HETERO and t is not None) else 0


def _f64(t, name):
    if t.dtype != torch.float64:
        raise TypeError(f"{name} must be float64")
    return t


# ----------------------------------------------------------------------------- variational verbs
def init_aux_posterior(lik: Likelihood, n: int, ctx: Optional[Context] = None, with_y_copy: bool = True,
                       host: bool = False):
    """init_aux_posterior(lik, n) — zero-filled state (bernoulli.jl:7-11 ... categorical.jl:59-70).
    host=True: the state lives in (pinned) host memory and the verbs take the host-buffer path."""
    ctx = ctx or default_context()
    names = _FIELDS[lik.kind]
    shape = (n, lik.nlatent) if _is_cat(lik) else (n,)
    fields = {}
    if not host:
        ctx.enter()
    for i, name in enumerate(names):
        if name is None:
            continue
        if name == "y":
            if not with_y_copy:
                continue
            dt = torch.uint8 if _is_cat(lik) else torch.int64
        else:
            dt = torch.float64
        fields[name] = _hempty(shape, dt) if host else ctx.empty(shape, dt)
    q = AuxPosterior(lik, fields)
    d = lik._desc()
    if host:
        check(ctx.lib.aug_init_aux_posterior_host(ctx.h, C.byref(d), n, _hptr(q._s(0)), _hptr(q._s(1)),
                                                  _hptr(q._s(2))))
        return q
    check(ctx.lib.aug_init_aux_posterior(ctx.h, C.byref(d), n, _ptr(q._s(0)), _ptr(q._s(1)), _ptr(q._s(2))))
    ctx.leave()
    return q


def aux_posterior_(qΩ: AuxPosterior, lik: Likelihood, y, qf: Normals, ctx: Optional[Context] = None):
    """aux_posterior!(qΩ, lik, y, qf) — updates the variational parameters in place, returns qΩ."""
    ctx = ctx or default_context()
    y = _check_y(lik, _t(y))
    d = lik._desc()
    if _is_host(y, qf.mu, qf.var, qΩ._s(0)):
        check(ctx.lib.aug_aux_posterior_host(ctx.h, C.byref(d), _nobs(lik, y), _hptr(y), _hptr(_f64(qf.mu, "mu")),
                                             _hptr(_f64(qf.var, "var")), _ld(lik, qf.mu), _hptr(qΩ._s(0)),
                                             _hptr(qΩ._s(1)), _hptr(qΩ._s(2))))
        return qΩ
    ctx.enter()
    check(ctx.lib.aug_aux_posterior(ctx.h, C.byref(d), _nobs(lik, y), _ptr(y), _ptr(_f64(qf.mu, "mu")),
                                    _ptr(_f64(qf.var, "var")), _ld(lik, qf.mu), _ptr(qΩ._s(0)), _ptr(qΩ._s(1)),
                                    _ptr(qΩ._s(2))))
    ctx.leave()
    return qΩ


def aux_posterior(lik: Likelihood, y, qf: Normals, ctx: Optional[Context] = None):
    """aux_posterior(lik, y, qf) — generic.jl:22-24"""
    ctx = ctx or default_context()
    y = _t(y)
    return aux_posterior_(init_aux_posterior(lik, _nobs(lik, y), ctx, host=not y.is_cuda), lik, y, qf, ctx)


def _alloc_bg(ctx, lik, n, want_beta=True, want_gamma=True):
    b = ctx.empty((lik.nlatent, n)) if want_beta else None
    g = ctx.empty((lik.nlatent, n)) if want_gamma else None
    return b, g


def _tuple(t):
    return tuple(t[i] for i in range(t.shape[0]))


def expected_auglik_potential_and_precision(lik, qΩ: AuxPosterior, y, qf: Optional[Normals] = None,
                                            ctx: Optional[Context] = None, _b=True, _g=True):
    """expected_auglik_potential_and_precision(lik, qΩ, y[, qf]) -> (β tuple, γ tuple)  (a7)"""
    ctx = ctx or default_context()
    y = _check_y(lik, _t(y))
    n = _nobs(lik, y)
    if lik.kind == HETERO and qf is None:
        raise TypeError("the heteroscedastic likelihood needs qf (heteroscedasticgaussian.jl:68-104)")
    d = lik._desc()
    mu = qf.mu if (qf is not None and lik.kind == HETERO) else None
    if _is_host(y, mu, qΩ._s(0)):
        beta = _hempty((lik.nlatent, n)) if _b else None
        gamma = _hempty((lik.nlatent, n)) if _g else None
        check(ctx.lib.aug_expected_potential_precision_host(ctx.h, C.byref(d), n, _hptr(y), _hptr(mu), _ld(lik, mu),
                                                            _hptr(qΩ._s(0)), _hptr(qΩ._s(1)), _hptr(qΩ._s(2)),
                                                            _hptr(beta), _hptr(gamma), n))
        return (_tuple(beta) if _b else None), (_tuple(gamma) if _g else None)
    ctx.enter()
    beta, gamma = _alloc_bg(ctx, lik, n, _b, _g)
    check(ctx.lib.aug_expected_potential_precision(ctx.h, C.byref(d), n, _ptr(y), _ptr(mu), _ld(lik, mu),
                                                   _ptr(qΩ._s(0)), _ptr(qΩ._s(1)), _ptr(qΩ._s(2)), _ptr(beta),
                                                   _ptr(gamma), n))
    ctx.leave()
    return (_tuple(beta) if _b else None), (_tuple(gamma) if _g else None)


def expected_auglik_potential(lik, qΩ, y, qf=None, ctx=None):
    return expected_auglik_potential_and_precision(lik, qΩ, y, qf, ctx, _b=True, _g=False)[0]


def expected_auglik_precision(lik, qΩ, y, qf=None, ctx=None):
    return expected_auglik_potential_and_precision(lik, qΩ, y, qf, ctx, _b=False, _g=True)[1]


def cavi_step_(qΩ: AuxPosterior, lik, y, qf: Normals, want_elbo: bool = True, ctx: Optional[Context] = None,
               out=None, want_beta: bool = True):
    """Fused call pattern of examples/bernoulli/script.jl:29-39: aux_posterior! +
    expected_auglik_potential_and_precision (+ expected_logtilt, aux_kldivergence, expected_aug_loglik).
    Returns (qΩ, β tuple, γ tuple, scalars) with scalars a device tensor of 8 doubles (or None)."""
    ctx = ctx or default_context()
    y = _check_y(lik, _t(y))
    n = _nobs(lik, y)
    d = lik._desc()
    if _is_host(y, qf.mu, qf.var, qΩ._s(0) if qΩ is not None else None):
        # host-buffer path; optional outputs: qΩ=None skips the state, want_beta=False skips β (D2H bytes saved)
        if out is None:
            beta = _hempty((lik.nlatent, n)) if want_beta else None
            gamma = _hempty((lik.nlatent, n))
        else:
            beta, gamma = out[0], out[1]
        hs = _hscal() if want_elbo else None
        s_ = [qΩ._s(i) if qΩ is not None else None for i in range(3)]
        check(ctx.lib.aug_cavi_step_host(ctx.h, C.byref(d), n, _hptr(y), _hptr(_f64(qf.mu, "mu")),
                                         _hptr(_f64(qf.var, "var")), _ld(lik, qf.mu), _hptr(s_[0]), _hptr(s_[1]),
                                         _hptr(s_[2]), _hptr(beta), _hptr(gamma), n, hs))
        return (qΩ, _tuple(beta) if beta is not None else None, _tuple(gamma),
                _scal_tensor(hs) if want_elbo else None)
    ctx.enter()
    if out is None:
        beta, gamma = _alloc_bg(ctx, lik, n)
        scal = ctx.empty((NSCALARS,)) if want_elbo else None
    else:
        beta, gamma, scal = out
    check(ctx.lib.aug_cavi_step(ctx.h, C.byref(d), n, _ptr(y), _ptr(_f64(qf.mu, "mu")), _ptr(_f64(qf.var, "var")),
                                _ld(lik, qf.mu), _ptr(qΩ._s(0)), _ptr(qΩ._s(1)), _ptr(qΩ._s(2)), _ptr(beta),
                                _ptr(gamma), n, _ptr(scal)))
    ctx.leave()
    return qΩ, _tuple(beta), _tuple(gamma), scal


def _elbo_terms(lik, qΩ, y, qf, ctx):
    ctx = ctx or default_context()
    y = _check_y(lik, _t(y))
    d = lik._desc()
    if _is_host(y, qf.mu, qf.var, qΩ._s(0)):
        hs = _hscal()
        check(ctx.lib.aug_expected_elbo_terms_host(ctx.h, C.byref(d), _nobs(lik, y), _hptr(y),
                                                   _hptr(_f64(qf.mu, "mu")), _hptr(_f64(qf.var, "var")),
                                                   _ld(lik, qf.mu), _hptr(qΩ._s(0)), _hptr(qΩ._s(1)),
                                                   _hptr(qΩ._s(2)), hs))
        return _scal_tensor(hs)
    ctx.enter()
    scal = ctx.empty((NSCALARS,))      # the verb writes all 8 slots (its own + zeros): nothing to clear here
    check(ctx.lib.aug_expected_elbo_terms(ctx.h, C.byref(d), _nobs(lik, y), _ptr(y), _ptr(_f64(qf.mu, "mu")),
                                          _ptr(_f64(qf.var, "var")), _ld(lik, qf.mu), _ptr(qΩ._s(0)),
                                          _ptr(qΩ._s(1)), _ptr(qΩ._s(2)), _ptr(scal)))
    ctx.leave()
    return scal


def _reduce(ctx, scal):
    """sum the scalar block over ranks when a communicator is attached"""
    ctx = ctx or default_context()
    if not scal.is_cuda:            # host-buffer verbs are rank-local (include/augcuda.h)
        return scal
    if ctx.fused:                   # exchanged inside the reducing kernel over peer memory
        if ctx.deferred:            # split-phase: the gather has not run yet
            ctx.flush()
        return scal
    if ctx.comm_ready:
        ctx.enter()
        check(ctx.lib.aug_allreduce_scalars(ctx.h, _ptr(scal), NSCALARS))
        ctx.leave()
    return scal


def expected_logtilt(lik, qΩ, y, qf, ctx=None) -> float:
    """expected_logtilt(lik, qΩ, y, qf) — api.jl:219-223 and the per-likelihood methods"""
    return float(_reduce(ctx, _elbo_terms(lik, qΩ, y, qf, ctx))[_lib.S_EXPECTED_LOGTILT].item())


def aux_kldivergence(lik, qΩ, y, qf=None, ctx=None) -> float:
    """aux_kldivergence(lik, qΩ, y) — generic.jl:56-62 (qf is only needed for the heteroscedastic prior)"""
    if qf is None:
        if lik.kind == HETERO:
            raise TypeError("qf is required for the heteroscedastic likelihood")
        n = len(qΩ)
        ctx = ctx or default_context()
        shape = (n, lik.nlatent) if _is_cat(lik) else (n,)
        z = torch.zeros(shape, dtype=torch.float64, device=ctx.tdev if _t(y).is_cuda else "cpu")
        qf = Normals(z, z)
    return float(_reduce(ctx, _elbo_terms(lik, qΩ, y, qf, ctx))[_lib.S_KL].item())


def expected_aug_loglik(lik, qΩ, y, qf, ctx=None) -> float:
    """expected_aug_loglik = expected_logtilt + aux_kldivergence (generic.jl:52-54 — note the plus)"""
    return float(_reduce(ctx, _elbo_terms(lik, qΩ, y, qf, ctx))[_lib.S_EXPECTED_AUGLL].item())


# ----------------------------------------------------------------------------- sampling verbs
def _apply_rng(ctx, rng):
    if rng is not None:
        ctx.seed(rng.seed, rng.offset)


def _store_rng(ctx, rng):
    if rng is not None:
        rng.offset = ctx.offset()


def init_aux_variables(*args, ctx: Optional[Context] = None, i0: int = 0, host: bool = False):
    """init_aux_variables([rng,] lik, n) — generic.jl:32-34 and the per-likelihood methods"""
    rng, (lik, n) = (args[0], args[1:]) if isinstance(args[0], AugPhilox) else (None, args)
    ctx = ctx or default_context()
    _apply_rng(ctx, rng)
    shape = (n, lik.nlatent) if _is_cat(lik) else (n,)
    if host:
        omega = _hempty(shape)
        nv = _hempty(shape, torch.int64) if (lik.kind in (POISSON, HETERO) or _is_cat(lik)) else None
        d = lik._desc()
        check(ctx.lib.aug_init_aux_variables_host(ctx.h, C.byref(d), n, i0, _hptr(omega), _hptr(nv)))
        _store_rng(ctx, rng)
        return AuxSamples(omega, nv)
    ctx.enter()
    omega = ctx.empty(shape)
    nv = ctx.empty(shape, torch.int64) if (lik.kind in (POISSON, HETERO) or _is_cat(lik)) else None
    d = lik._desc()
    check(ctx.lib.aug_init_aux_variables(ctx.h, C.byref(d), n, i0, _ptr(omega), _ptr(nv)))
    ctx.leave()
    _store_rng(ctx, rng)
    return AuxSamples(omega, nv)


def aux_sample_(*args, ctx: Optional[Context] = None, i0: int = 0):
    """aux_sample!([rng,] Ω, lik, y, f) — generic.jl:1-12; returns Ω"""
    rng, (Ω, lik, y, f) = (args[0], args[1:]) if isinstance(args[0], AugPhilox) else (None, args)
    ctx = ctx or default_context()
    _apply_rng(ctx, rng)
    y = _check_y(lik, _t(y))
    f = _t(f)
    d = lik._desc()
    if _is_host(y, f, Ω.omega, Ω.n):
        check(ctx.lib.aug_aux_sample_host(ctx.h, C.byref(d), _nobs(lik, y), i0, _hptr(y), _hptr(_f64(f, "f")),
                                          _ld(lik, f), _hptr(Ω.omega), _hptr(Ω.n)))
        _store_rng(ctx, rng)
        return Ω
    ctx.enter()
    check(ctx.lib.aug_aux_sample(ctx.h, C.byref(d), _nobs(lik, y), i0, _ptr(y), _ptr(_f64(f, "f")), _ld(lik, f),
                                 _ptr(Ω.omega), _ptr(Ω.n)))
    ctx.leave()
    _store_rng(ctx, rng)
    return Ω


def aux_sample(*args, ctx: Optional[Context] = None, i0: int = 0):
    """aux_sample([rng,] lik, y, f) — generic.jl:14-20"""
    rng, (lik, y, f) = (args[0], args[1:]) if isinstance(args[0], AugPhilox) else (None, args)
    ctx = ctx or default_context()
    y = _t(y)
    n = _nobs(lik, y)
    shape = (n, lik.nlatent) if _is_cat(lik) else (n,)
    needs_n = lik.kind in (POISSON, HETERO) or _is_cat(lik)
    if not y.is_cuda:
        omega, nv = _hempty(shape), (_hempty(shape, torch.int64) if needs_n else None)
    else:
        ctx.enter()
        omega = ctx.empty(shape)
        nv = ctx.empty(shape, torch.int64) if needs_n else None
        ctx.leave()
    Ω = AuxSamples(omega, nv)
    return aux_sample_(*(([rng] if rng else []) + [Ω, lik, y, f]), ctx=ctx, i0=i0)


def auglik_potential_and_precision(lik, Ω: AuxSamples, y, f=None, ctx=None, _b=True, _g=True):
    """auglik_potential_and_precision(lik, Ω, y[, f]) (a20)"""
    ctx = ctx or default_context()
    y = _check_y(lik, _t(y))
    n = _nobs(lik, y)
    if lik.kind == HETERO and f is None:
        raise TypeError("the heteroscedastic likelihood needs f (heteroscedasticgaussian.jl:48-66)")
    d = lik._desc()
    ff = _t(f) if lik.kind == HETERO else None
    if _is_host(y, ff, Ω.omega, Ω.n):
        beta = _hempty((lik.nlatent, n)) if _b else None
        gamma = _hempty((lik.nlatent, n)) if _g else None
        check(ctx.lib.aug_potential_precision_host(ctx.h, C.byref(d), n, _hptr(y), _hptr(ff), _ld(lik, ff),
                                                   _hptr(Ω.omega), _hptr(Ω.n), _hptr(beta), _hptr(gamma), n))
        return (_tuple(beta) if _b else None), (_tuple(gamma) if _g else None)
    ctx.enter()
    beta, gamma = _alloc_bg(ctx, lik, n, _b, _g)
    check(ctx.lib.aug_potential_precision(ctx.h, C.byref(d), n, _ptr(y), _ptr(ff), _ld(lik, ff), _ptr(Ω.omega),
                                          _ptr(Ω.n), _ptr(beta), _ptr(gamma), n))
    ctx.leave()
    return (_tuple(beta) if _b else None), (_tuple(gamma) if _g else None)


def auglik_potential(lik, Ω, y, f=None, ctx=None):
    return auglik_potential_and_precision(lik, Ω, y, f, ctx, _b=True, _g=False)[0]


def auglik_precision(lik, Ω, y, f=None, ctx=None):
    return auglik_potential_and_precision(lik, Ω, y, f, ctx, _b=False, _g=True)[1]


def _sampled_terms(lik, Ω, y, f, with_prior, ctx):
    ctx = ctx or default_context()
    y = _check_y(lik, _t(y))
    f = _t(f)
    d = lik._desc()
    if _is_host(y, f, Ω.omega, Ω.n):
        hs = _hscal()
        check(ctx.lib.aug_sampled_loglik_terms_host(ctx.h, C.byref(d), _nobs(lik, y), _hptr(y), _hptr(_f64(f, "f")),
                                                    _ld(lik, f), _hptr(Ω.omega), _hptr(Ω.n), int(with_prior), hs))
        return _scal_tensor(hs)
    ctx.enter()
    scal = ctx.empty((NSCALARS,))      # the verb writes all 8 slots (its own + zeros): nothing to clear here
    check(ctx.lib.aug_sampled_loglik_terms(ctx.h, C.byref(d), _nobs(lik, y), _ptr(y), _ptr(_f64(f, "f")),
                                           _ld(lik, f), _ptr(Ω.omega), _ptr(Ω.n), int(with_prior), _ptr(scal)))
    ctx.leave()
    return _reduce(ctx, scal)


def logtilt(lik, Ω, y, f, ctx=None) -> float:
    """logtilt(lik, Ω, y, f) — generic.jl:40-46"""
    return float(_sampled_terms(lik, Ω, y, f, False, ctx)[_lib.S_LOGTILT].item())


def aug_loglik(lik, Ω, y, f, ctx=None) -> float:
    """aug_loglik(lik, Ω, y, f) = logtilt + logdensity(aux_prior(lik, y), Ω) — generic.jl:48-50"""
    return float(_sampled_terms(lik, Ω, y, f, True, ctx)[_lib.S_AUGLL].item())


# ----------------------------------------------------------------------------- SpecialDistributions
# ----------------------------------------------------------------------------- the two auxiliary-variable laws as objects
class AuxPrior:
    """aux_prior(lik, y) (a10): the prior p(Ω) of the augmentation — PolyaGamma(1, 0) / PolyaGamma(y + r, 0) /
    PolyaGammaPoisson / InverseGamma / Gamma / PolyaGammaNegativeMultinomial per likelihood.  The device
    implementation never materialises per-observation distribution objects: this handle carries (lik, y) and
    evaluates the log-density on the device."""

    def __init__(self, lik, y):
        if lik.kind == HETERO:
            raise TypeError("the heteroscedastic likelihood has no explicit prior / tilt split (can_split == false)")
        self.lik, self.y = lik, y

    def __len__(self):
        return _nobs(self.lik, self.y)

    def logdensity(self, Ω: AuxSamples, ctx=None) -> float:
        """logdensity_def(aux_prior(lik, y), Ω) — generic.jl:49; the prior does not depend on f."""
        ctx = ctx or default_context()
        n = len(self)
        shape = (n, self.lik.nlatent) if _is_cat(self.lik) else (n,)
        f0 = torch.zeros(shape, dtype=torch.float64, device=ctx.tdev if Ω.omega.is_cuda else "cpu")
        return float(_sampled_terms(self.lik, Ω, self.y, f0, True, ctx)[_lib.S_LOGPRIOR].item())


class AuxFullConditional:
    """aux_full_conditional(lik, y, f) (a14): p(Ω | y, f), the law aux_sample! draws from."""

    def __init__(self, lik, y, f):
        self.lik, self.y, self.f = lik, y, f

    def __len__(self):
        return _nobs(self.lik, self.y)

    def rand(self, rng: Optional[AugPhilox] = None, ctx=None, i0: int = 0) -> AuxSamples:
        """tvrand(rng, aux_full_conditional(lik, y, f)) == aux_sample(rng, lik, y, f) — generic.jl:14-20"""
        args = ([rng] if rng is not None else []) + [self.lik, self.y, self.f]
        return aux_sample(*args, ctx=ctx, i0=i0)


def aux_prior(lik, y) -> AuxPrior:
    return AuxPrior(lik, y)


def aux_full_conditional(lik, y, f) -> AuxFullConditional:
    return AuxFullConditional(lik, y, f)


def logdensity_def(dist: AuxPrior, Ω: AuxSamples, ctx=None) -> float:
    return dist.logdensity(Ω, ctx=ctx)


def tvrand(*args, ctx=None, i0: int = 0) -> AuxSamples:
    """tvrand([rng,] dist) for dist = aux_full_conditional(lik, y, f) — TestUtils.jl:108-110"""
    rng, dist = (args[0], args[1]) if isinstance(args[0], AugPhilox) else (None, args[0])
    return dist.rand(rng, ctx=ctx, i0=i0)


def pg_rand(b, c, n=None, b_is_int=None, rng: Optional[AugPhilox] = None, ctx=None, i0=0):
    """rand(PolyaGamma(b, c)[, n]) — SpecialDistributions/polyagamma.jl:121-164"""
    ctx = ctx or default_context()
    _apply_rng(ctx, rng)
    ctx.enter()
    if torch.is_tensor(b):
        out = ctx.empty(b.shape)
        check(ctx.lib.aug_pg_rand(ctx.h, b.numel(), i0, _ptr(b), _ptr(c), int(bool(b_is_int)), _ptr(out)))
    else:
        if b_is_int is None:
            b_is_int = isinstance(b, int)
        out = ctx.empty((int(n),))
        check(ctx.lib.aug_pg_rand_bc(ctx.h, int(n), i0, float(b), float(c), int(b_is_int), _ptr(out)))
    ctx.leave()
    _store_rng(ctx, rng)
    return out


def _ew(fn, ctx, a, b):
    ctx = ctx or default_context()
    ctx.enter()
    out = ctx.empty(a.shape)
    check(fn(ctx.h, a.numel(), _ptr(a), _ptr(b), _ptr(out)))
    ctx.leave()
    return out


def pg_mean(b, c, ctx=None):
    """mean(PolyaGamma(b, c)) — polyagamma.jl:25-31"""
    ctx = ctx or default_context()
    return _ew(ctx.lib.aug_pg_mean, ctx, b, c)


def pg_kldivergence(b, c, ctx=None):
    """kldivergence(PolyaGamma(b, c), PolyaGamma(b, 0)) — polyagamma.jl:99-110"""
    ctx = ctx or default_context()
    return _ew(ctx.lib.aug_pg_kl, ctx, b, c)


def pg_logpdf(b, c, x, ctx=None):
    """logpdf(PolyaGamma(b, c), x) — polyagamma.jl:37-91"""
    ctx = ctx or default_context()
    ctx.enter()
    out = ctx.empty(x.shape)
    check(ctx.lib.aug_pg_logpdf(ctx.h, x.numel(), float(b), float(c), _ptr(x), _ptr(out)))
    ctx.leave()
    return out


def approx_expected_logistic(mu, c, ctx=None):
    """approx_expected_logistic(μ, c) — utils.jl:11-14"""
    ctx = ctx or default_context()
    return _ew(ctx.lib.aug_approx_expected_logistic, ctx, mu, c)


def second_moment(q: Normals, y=None, ctx=None):
    """second_moment(q[, y]) — utils.jl:1-7"""
    ctx = ctx or default_context()
    ctx.enter()
    out = ctx.empty(q.mu.shape)
    check(ctx.lib.aug_second_moment(ctx.h, q.mu.numel(), _ptr(q.mu), _ptr(q.var), _ptr(y), _ptr(out)))
    ctx.leave()
    return out


def fastmath_eval(fn: int, x, ctx=None):
    """diagnostics: csrc/aug_fastmath.cuh functions (0 rcp, 1 rsqrt, 2 exp, 3 log, 4 log on [1,2], 5 sqrt)"""
    ctx = ctx or default_context()
    ctx.enter()
    out = ctx.empty(x.shape)
    check(ctx.lib.aug_fastmath_eval(ctx.h, int(fn), x.numel(), _ptr(x), _ptr(out)))
    ctx.leave()
    return out


# ----------------------------------------------------------------------------- SURVEY §8(f) rows 3 and 4
def hetero_lambda_stats(y, qfg: Normals, ctx=None) -> float:
    """dot(ψ, 1 .- σ̃g) of opt_lik — examples/heteroscedasticgaussian/script.jl:41-51 (summed over all ranks in
    fused multi-GPU mode).  qfg: Normals with mu/var of shape [2][n] (f then g)."""
    ctx = ctx or default_context()
    n = y.numel()
    ctx.enter()
    out = ctx.empty((1,))
    check(ctx.lib.aug_hetero_lambda_stats(ctx.h, n, _ptr(_f64(y, "y")), _ptr(_f64(qfg.mu, "mu")),
                                          _ptr(_f64(qfg.var, "var")), qfg.mu.shape[-1], _ptr(out)))
    ctx.leave()
    return float(out.item())


def opt_lik(lik: HeteroscedasticGaussianLikelihood, qfg: Normals, y, ctx=None, n_total=None):
    """opt_lik(lik, (qf, qg), y): λ = max(N / (2 dot(ψ, 1 .- σ̃g)), λ_old) — script.jl:41-51.
    n_total: the global number of observations when y is a shard (fused multi-GPU mode)."""
    s = hetero_lambda_stats(y, qfg, ctx)
    n = y.numel() if n_total is None else int(n_total)
    return HeteroscedasticGaussianLikelihood(max(n / (2.0 * s), lik.lam))


def hetero_lambda_stats_sampled(y, fg, ctx=None) -> float:
    """Σ σ(gᵢ)/2 (yᵢ − fᵢ)²: rate increment of the Gamma full conditional of λ —
    docs/src/likelihoods/heteroscedasticgaussian.md:80-84.  fg: [2][n]."""
    ctx = ctx or default_context()
    n = y.numel()
    ctx.enter()
    out = ctx.empty((1,))
    check(ctx.lib.aug_hetero_lambda_stats_sampled(ctx.h, n, _ptr(_f64(y, "y")), _ptr(_f64(fg, "f")),
                                                  fg.shape[-1], _ptr(out)))
    ctx.leave()
    return float(out.item())


def logisticsoftmax(f, lik: Optional[CategoricalLikelihood] = None, ctx=None):
    """logisticsoftmax(x) (categorical.jl:1-4) when lik is None, else lik's inverse link applied row-wise:
    LogisticSoftMaxLink(logθ)(f) (categorical.jl:32-35), with the bijective link's appended zero latent.
    f: [n][nl] -> [n][K]."""
    ctx = ctx or default_context()
    f2 = f if f.dim() == 2 else f.reshape(1, -1)
    if lik is None:
        lik = CategoricalLikelihood(f2.shape[1], bijective=False)
    if f2.shape[1] != lik.nlatent:
        raise ValueError("f must have nlatent(lik) columns")
    K = lik.nlatent + 1 if lik.bijective else lik.nlatent
    d = lik._desc()
    ctx.enter()
    out = ctx.empty((f2.shape[0], K))
    check(ctx.lib.aug_logisticsoftmax(ctx.h, C.byref(d), f2.shape[0], _ptr(_f64(f2, "f")), _ptr(out)))
    ctx.leave()
    return out if f.dim() == 2 else out.reshape(-1)


def approx_expected_logisticsoftmax(mu, c, lik: CategoricalLikelihood, ctx=None):
    """approx_expected_logisticsoftmax(μ, c, θ) — utils.jl:17-22, θ = exp.(logθ) of a bijective link.  [n][nl]."""
    ctx = ctx or default_context()
    if not lik.bijective:
        raise TypeError("approx_expected_logisticsoftmax needs the bijective link (θ has one more entry than μ)")
    d = lik._desc()
    ctx.enter()
    out = ctx.empty(mu.shape)
    check(ctx.lib.aug_approx_expected_logisticsoftmax(ctx.h, C.byref(d), mu.shape[0], _ptr(_f64(mu, "mu")),
                                                      _ptr(_f64(c, "c")), _ptr(out)))
    ctx.leave()
    return out


# ----------------------------------------------------------------------------- SURVEY §8(f) rows 1 and 2
def _kappa(kappa):
    """κ = K_Z⁻¹ K_{Z,X}: Julia's column-major M×N matrix = a contiguous [n][m] tensor here."""
    if kappa.dim() != 2:
        raise ValueError("kappa must be [n][m] (the Julia M×N matrix as stored)")
    return _f64(kappa, "kappa")


def sparse_marginals(kappa, mvec, B, kdiag, ctx=None) -> Normals:
    """marginals(post_u(x)) of the SVGP posterior — examples/bernoulli/script.jl:32-33:
    μ_t = κ_tᵀ m, σ²_t = k_tt − κ_tᵀ (K_Z − S) κ_t.  B = K_Z − S ([m][m])."""
    ctx = ctx or default_context()
    kappa = _kappa(kappa)
    n, m = kappa.shape
    ctx.enter()
    mu, var = ctx.empty((n,)), ctx.empty((n,))
    check(ctx.lib.aug_sparse_marginals(ctx.h, n, m, _ptr(kappa), _ptr(_f64(mvec, "mvec")), _ptr(_f64(B, "B")),
                                       _ptr(_f64(kdiag, "kdiag")), _ptr(mu), _ptr(var)))
    ctx.leave()
    return Normals(mu, var)


def sparse_marginals_into_(qf: Normals, j: int, kappa, mvec, B, kdiag, ctx=None) -> Normals:
    """Latent j of a multi-latent qf in place: qf.mu[:, j], qf.var[:, j] for the class-fastest [n][nl] layout of the
    Categorical likelihood (one call per latent GP; κ may be shared), qf.mu[j], qf.var[j] for the latent-major [2][n]
    layout of the heteroscedastic one."""
    ctx = ctx or default_context()
    kappa = _kappa(kappa)
    n, m = kappa.shape
    mu, var = _f64(qf.mu, "mu"), _f64(qf.var, "var")
    if mu.dim() != 2 or not mu.is_contiguous() or not var.is_contiguous():
        raise ValueError("qf must hold contiguous 2-d mu / var")
    if mu.shape[0] == n and mu.shape[1] != n:              # [n][nl]: class fastest
        off, stride = j * 8, mu.shape[1]
    else:                                                  # [nl][n]: latent major
        off, stride = j * n * 8, 1
    ctx.enter()
    check(ctx.lib.aug_sparse_marginals_strided(ctx.h, n, m, _ptr(kappa), _ptr(_f64(mvec, "mvec")), _ptr(_f64(B, "B")),
                                               _ptr(_f64(kdiag, "kdiag")), C.c_void_p(mu.data_ptr() + off),
                                               C.c_void_p(var.data_ptr() + off), stride))
    ctx.leave()
    return qf


def _split_pr(Pr, m):
    return Pr[: m * m].view(m, m), Pr[m * m:]


def _reduce_pr(ctx, Pr):
    if ctx.fused:                   # exchanged inside the verb's finalise launch over peer memory
        return
    if ctx.comm_ready:              # shards of the observation axis: P and rhs are plain sums
        check(ctx.lib.aug_allreduce_scalars(ctx.h, _ptr(Pr), Pr.numel()))


def _prior_terms(ctx, P0, r0):
    """P0 / r0 are given on every rank.  NCCL mode adds them once (rank 0) before the sum over ranks; fused mode adds
    them on every rank after the in-kernel exchange (include/augcuda.h)."""
    if ctx.comm_ready and not ctx.fused and getattr(ctx, "rank", 0) != 0:
        return None, None
    return P0, r0


def sparse_precision_potential(kappa, gamma, beta, P0=None, r0=None, ctx=None):
    """docs/src/index.md:156-160: returns (P0 + κ·Diagonal(γ)·κᵀ, r0 + κ·β).  In a sharded run pass the same P0 / r0
    on every rank; the sums over ranks are taken here (NCCL) or inside the verb (fused mailbox mode)."""
    ctx = ctx or default_context()
    kappa = _kappa(kappa)
    n, m = kappa.shape
    ctx.enter()
    Pr = ctx.empty((m * m + m,))
    P0, r0 = _prior_terms(ctx, P0, r0)
    check(ctx.lib.aug_sparse_precision_potential(ctx.h, n, m, _ptr(kappa), _ptr(_f64(gamma, "gamma")),
                                                 _ptr(_f64(beta, "beta")), _ptr(P0), _ptr(r0), _ptr(Pr)))
    _reduce_pr(ctx, Pr)
    ctx.leave()
    return _split_pr(Pr, m)


def sparse_cavi_sweep_(qΩ: Optional[AuxPosterior], lik, y, kappa, mvec, B, kdiag, P0=None, r0=None,
                       want_elbo: bool = True, want_marginals: bool = False, want_potentials: bool = False,
                       ctx: Optional[Context] = None):
    """One pass over κ for a whole CAVI iteration of a sparse variational GP (examples/bernoulli/script.jl:29-39 with
    the sparse update of docs/src/index.md:154-163): marginals → aux_posterior!(qΩ, …) → E[β], E[γ] (+ ELBO sums)
    → P = P0 + κ·Diagonal(γ)·κᵀ, rhs = r0 + κ·β.  qΩ may be None (state not materialised).
    Returns (P [m][m], rhs [m], scalars | None, qf | None, (β, γ) | None)."""
    ctx = ctx or default_context()
    if lik.kind in (HETERO, CAT, CAT_BIJ):
        raise ValueError("multi-latent likelihoods: call sparse_marginals / sparse_precision_potential per latent GP")
    y = _check_y(lik, y)
    kappa = _kappa(kappa)
    n, m = kappa.shape
    d = lik._desc()
    ctx.enter()
    Pr = ctx.empty((m * m + m,))
    scal = ctx.empty((NSCALARS,)) if want_elbo else None
    mu = ctx.empty((n,)) if want_marginals else None
    var = ctx.empty((n,)) if want_marginals else None
    beta = ctx.empty((n,)) if want_potentials else None
    gamma = ctx.empty((n,)) if want_potentials else None
    s = [qΩ._s(i) if qΩ is not None else None for i in range(3)]
    P0, r0 = _prior_terms(ctx, P0, r0)
    check(ctx.lib.aug_sparse_cavi_sweep(ctx.h, C.byref(d), n, m, _ptr(y), _ptr(kappa), _ptr(_f64(mvec, "mvec")),
                                        _ptr(_f64(B, "B")), _ptr(_f64(kdiag, "kdiag")), _ptr(mu), _ptr(var),
                                        _ptr(s[0]), _ptr(s[1]), _ptr(s[2]), _ptr(beta), _ptr(gamma), _ptr(P0),
                                        _ptr(r0), _ptr(Pr), _ptr(scal)))
    _reduce_pr(ctx, Pr)
    if scal is not None and ctx.comm_ready and not ctx.fused:
        check(ctx.lib.aug_allreduce_scalars(ctx.h, _ptr(scal), NSCALARS))
    ctx.leave()
    P, rhs = _split_pr(Pr, m)
    return (P, rhs, scal, Normals(mu, var) if want_marginals else None,
            (beta, gamma) if want_potentials else None)


def dense_precision_potential(Kinv, gamma, beta, r0=None, out=None, ctx=None):
    """examples/bernoulli/script.jl:35-36: (inv(K) + Diagonal(γ), β + K \\ mean(fz)); out=Kinv updates in place."""
    ctx = ctx or default_context()
    n = gamma.numel()
    ctx.enter()
    P = out if out is not None else ctx.empty((n, n))
    rhs = ctx.empty((n,))
    check(ctx.lib.aug_dense_precision_potential(ctx.h, n, _ptr(_f64(Kinv, "Kinv")), _ptr(_f64(gamma, "gamma")),
                                                _ptr(_f64(beta, "beta")), _ptr(r0), _ptr(P), _ptr(rhs)))
    ctx.leave()
    return P, rhs
