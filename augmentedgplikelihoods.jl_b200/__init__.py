"""augmentedgplikelihoods.jl_b200 — B200-native (sm_100a) augmentation hot path of AugmentedGPLikelihoods.jl.

csrc/      hand-written CUDA kernels + the C ABI of include/augcuda.h  ->  libaugcuda.so
api.py     host-side mirror of the reference verbs (one C-ABI call per verb)
dist.py    observation-axis sharding + NCCL all-reduce of the scalar block
julia/     the Julia glue a maintainer would add (ccall stubs; cannot be run in this image)
"""
from . import _lib
from ._lib import AugError, load  # noqa: F401

load()  # fail loudly at import time if libaugcuda.so is missing

from .api import *  # noqa: F401,F403,E402
from .api import (AugPhilox, AuxPosterior, AuxSamples, BernoulliLikelihood, CategoricalLikelihood,  # noqa: E402,F401
                  Context, HeteroscedasticGaussianLikelihood, LaplaceLikelihood, NegativeBinomialLikelihood,
                  Normals, PoissonLikelihood, StudentTLikelihood, default_context, set_default_context)
from . import dist  # noqa: E402,F401
from . import testutils  # noqa: E402,F401
