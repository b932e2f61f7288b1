"""ctypes binding of libaugcuda.so (the C ABI declared in include/augcuda.h).

There is NO fallback: if the shared library is missing the import raises, and every compute call
needs a CUDA device.  Nothing here imports or calls the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AUGCUDA_LIB lets a tuning run load an alternative BUILD of the same library (never a different backend)
LIB_PATH = os.environ.get("AUGCUDA_LIB") or os.path.join(_HERE, "libaugcuda.so")

BERNOULLI, NEGBIN, POISSON, LAPLACE, STUDENTT, HETERO, CAT_BIJ, CAT = range(8)
NSCALARS = 8
S_EXPECTED_LOGTILT, S_KL, S_EXPECTED_AUGLL, S_LOGTILT, S_LOGPRIOR, S_AUGLL, S_FLAGS = range(7)
ERR_PRECONDITION = -3


class AugLik(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nlatent", C.c_int32), ("r_is_int", C.c_int32),
                ("flags", C.c_int32), ("p", C.c_double * 4), ("logtheta", C.c_void_p)]


class AugError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"libaugcuda: {msg} (rc={rc})")
        self.rc = rc


_vp, _i32, _i64, _u64, _d = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double

# name -> argtypes; every function returns int32 except aug_strerror
SIGNATURES = {
    "aug_version": [],
    "aug_ctx_create": [C.POINTER(_vp), _i32, _vp],
    "aug_ctx_destroy": [_vp],
    "aug_ctx_seed": [_vp, _u64, _u64],
    "aug_ctx_get_offset": [_vp, C.POINTER(_u64)],
    "aug_ctx_sync": [_vp],
    "aug_ctx_stream": [_vp, C.POINTER(_vp)],
    "aug_ctx_sm_count": [_vp, C.POINTER(_i32)],
    "aug_ctx_launch_count": [_vp, C.POINTER(_u64)],
    "aug_ctx_error_flag": [_vp, C.POINTER(C.c_uint32)],
    "aug_malloc": [_vp, C.POINTER(_vp), C.c_size_t],
    "aug_free": [_vp, _vp],
    "aug_host_alloc": [C.POINTER(_vp), C.c_size_t],
    "aug_host_free": [_vp],
    "aug_memcpy_h2d": [_vp, _vp, _vp, C.c_size_t],
    "aug_memcpy_d2h": [_vp, _vp, _vp, C.c_size_t],
    "aug_init_aux_posterior": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _vp],
    "aug_aux_posterior": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "aug_expected_potential_precision": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp,
                                         _i64],
    "aug_cavi_step": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "aug_expected_elbo_terms": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "aug_init_aux_variables": [_vp, C.POINTER(AugLik), _i64, _i64, _vp, _vp],
    "aug_aux_sample": [_vp, C.POINTER(AugLik), _i64, _i64, _vp, _vp, _i64, _vp, _vp],
    "aug_potential_precision": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64],
    "aug_sampled_loglik_terms": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _i64, _vp, _vp, _i32, _vp],
    "aug_pg_rand": [_vp, _i64, _i64, _vp, _vp, _i32, _vp],
    "aug_pg_rand_bc": [_vp, _i64, _i64, _d, _d, _i32, _vp],
    "aug_pg_mean": [_vp, _i64, _vp, _vp, _vp],
    "aug_pg_kl": [_vp, _i64, _vp, _vp, _vp],
    "aug_pg_logpdf": [_vp, _i64, _d, _d, _vp, _vp],
    "aug_approx_expected_logistic": [_vp, _i64, _vp, _vp, _vp],
    "aug_second_moment": [_vp, _i64, _vp, _vp, _vp, _vp],
    "aug_fastmath_eval": [_vp, _i32, _i64, _vp, _vp],
    "aug_hetero_lambda_stats": [_vp, _i64, _vp, _vp, _vp, _i64, _vp],
    "aug_hetero_lambda_stats_sampled": [_vp, _i64, _vp, _vp, _i64, _vp],
    "aug_logisticsoftmax": [_vp, C.POINTER(AugLik), _i64, _vp, _vp],
    "aug_approx_expected_logisticsoftmax": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _vp],
    "aug_sparse_marginals": [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "aug_sparse_marginals_strided": [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64],
    "aug_sparse_precision_potential": [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "aug_sparse_cavi_sweep": [_vp, C.POINTER(AugLik), _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                              _vp, _vp, _vp, _vp, _vp, _vp],
    "aug_dense_precision_potential": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "aug_comm_get_unique_id": [C.c_char * 128],
    "aug_comm_init": [_vp, _i32, _i32, C.c_char * 128],
    "aug_comm_destroy": [_vp],
    "aug_allreduce_scalars": [_vp, _vp, _i32],
    "aug_comm_p2p_export": [_vp, _vp, C.POINTER(_vp)],
    "aug_comm_p2p_attach": [_vp, _i32, _i32, C.c_char_p],
    "aug_comm_p2p_attach_ptrs": [_vp, _i32, _i32, C.POINTER(_vp), C.POINTER(_i32)],
    "aug_comm_p2p_detach": [_vp],
    "aug_comm_set_fused": [_vp, _i32],
    "aug_comm_get_fused": [_vp, C.POINTER(_i32)],
    "aug_comm_set_deferred": [_vp, _i32],
    "aug_comm_flush": [_vp],
    "aug_allreduce_scalars_p2p": [_vp, _vp, _i32],
    "aug_cavi_step_host": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "aug_aux_sample_host": [_vp, C.POINTER(AugLik), _i64, _i64, _vp, _vp, _i64, _vp, _vp],
    "aug_init_aux_posterior_host": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _vp],
    "aug_aux_posterior_host": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "aug_expected_potential_precision_host": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp,
                                              _vp, _i64],
    "aug_expected_elbo_terms_host": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "aug_init_aux_variables_host": [_vp, C.POINTER(AugLik), _i64, _i64, _vp, _vp],
    "aug_potential_precision_host": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64],
    "aug_sampled_loglik_terms_host": [_vp, C.POINTER(AugLik), _i64, _vp, _vp, _i64, _vp, _vp, _i32, _vp],
}

_lib = None


def load():
    """Load libaugcuda.so and bind every symbol include/augcuda.h declares.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C augmentedgplikelihoods.jl_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.argtypes = argtypes
        fn.restype = C.c_int32
    lib.aug_strerror.argtypes = [C.c_int32]
    lib.aug_strerror.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise AugError(rc, load().aug_strerror(rc).decode())
