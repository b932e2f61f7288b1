"""CPU checks of the boundary: libaugcuda.so loads, exports every symbol include/augcuda.h declares,
and fails loudly (no fallback) when there is no CUDA device.  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

import aug_pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "augcuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aug_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    pkg = aug_pkg.load_package()
    lib = pkg.load()
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/augcuda.h but not exported"
    # and the ctypes table binds exactly the header's functions (plus strerror)
    bound = set(pkg._lib.SIGNATURES) | {"aug_strerror"}
    assert bound == set(syms), bound ^ set(syms)
    assert lib.aug_version() == 100


def test_strerror_and_struct_layout():
    pkg = aug_pkg.load_package()
    lib = pkg.load()
    assert lib.aug_strerror(0) == b"ok"
    assert b"precondition" in lib.aug_strerror(-3)
    assert C.sizeof(pkg._lib.AugLik) == 56          # 4 x int32 + 4 x double + pointer


def test_no_cpu_fallback():
    import torch
    pkg = aug_pkg.load_package()
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        pkg.Context()
    h = C.c_void_p()
    rc = pkg.load().aug_ctx_create(C.byref(h), 0, None)
    assert rc != 0 and not h.value                   # cudaErrorNoDevice / insufficient driver, never a CPU path


def test_product_never_touches_the_oracle():
    """The product path (package + csrc) must not import, link or execute anything under oracle/."""
    pkgdir = os.path.join(ROOT, "augmentedgplikelihoods.jl_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".jl")) or fn == "Makefile":
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "liborc" not in txt and "aug_oracle" not in txt and "oracle/" not in txt \
                    and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dirpath, fn)


def test_likelihood_descriptors():
    pkg = aug_pkg.load_package()
    d = pkg.NegativeBinomialLikelihood(10)._desc()
    assert (d.kind, d.r_is_int, d.p[0]) == (1, 1, 10.0)
    assert pkg.NegativeBinomialLikelihood(5.5)._desc().r_is_int == 0
    lik = pkg.CategoricalLikelihood(100)
    assert pkg.nlatent(lik) == 99 and lik.kind == 6
    assert pkg.nlatent(pkg.CategoricalLikelihood(100, bijective=False)) == 100
    assert pkg.nlatent(pkg.HeteroscedasticGaussianLikelihood(5.0)) == 2
    assert pkg.nlatent(pkg.BernoulliLikelihood()) == 1


def test_julia_glue_binds_only_declared_symbols_with_matching_arity():
    """AugCUDA.jl cannot be executed here (no Julia): check statically that every ccall names a function the header
    declares and passes as many arguments as the C prototype (= the ctypes signature table) has."""
    pkg = aug_pkg.load_package()
    src = open(os.path.join(ROOT, "augmentedgplikelihoods.jl_b200", "julia", "AugCUDA.jl")).read()
    syms = set(header_symbols())
    calls = list(re.finditer(r"ccall\(\(:(aug_[a-z0-9_]+),\s*lib\),\s*(\w+),\s*\(", src))
    assert len(calls) >= 20
    seen = set()
    for m in calls:
        name, ret = m.group(1), m.group(2)
        assert name in syms, f"{name} is not declared in include/augcuda.h"
        seen.add(name)
        # the argument-type tuple: balanced parentheses from the opening "(" of the tuple
        i = m.end()
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        types = src[i:j - 1]
        # split on top-level commas
        parts, d, cur = [], 0, ""
        for ch in types:
            if ch in "({[":
                d += 1
            elif ch in ")}]":
                d -= 1
            if ch == "," and d == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur)
        nargs = len([p for p in parts if p.strip()])
        if name == "aug_strerror":
            assert ret == "Cstring" and nargs == 1
            continue
        assert ret == "Int32", (name, ret)
        assert nargs == len(pkg._lib.SIGNATURES[name]), (name, nargs, len(pkg._lib.SIGNATURES[name]))
    # the verbs of the path and of the rows either side of it are all bound
    for must in ("aug_cavi_step", "aug_aux_sample", "aug_expected_elbo_terms", "aug_sampled_loglik_terms",
                 "aug_sparse_cavi_sweep", "aug_sparse_marginals", "aug_sparse_precision_potential"):
        assert must in seen, must
