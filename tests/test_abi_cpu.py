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


def header_prototypes():
    """name -> list of C argument-type classes parsed from include/augcuda.h (ptr / i32 / i64 / u64 / size / f64)"""
    src = open(os.path.join(ROOT, "include", "augcuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int32_t|const char\s*\*)\s+(aug_[a-z0-9_]+)\s*\(([^;{]*)\)\s*;", src):
        name, args = m.group(1), " ".join(m.group(2).split())
        classes = []
        if args not in ("", "void"):
            for a in args.split(","):
                if "*" in a or "[" in a:
                    classes.append("ptr")
                elif "uint64_t" in a:
                    classes.append("u64")
                elif "int64_t" in a:
                    classes.append("i64")
                elif "int32_t" in a:
                    classes.append("i32")
                elif "size_t" in a:
                    classes.append("u64")      # size_t == uint64_t on LP64 (ctypes aliases them)
                elif "double" in a:
                    classes.append("f64")
                else:
                    raise AssertionError(f"unclassified C argument {a!r} of {name}")
        protos[name] = classes
    return protos


def ctypes_class(t):
    if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or hasattr(t, "_length_"):
        return "ptr"
    return {C.c_int32: "i32", C.c_int64: "i64", C.c_uint64: "u64", C.c_double: "f64", C.c_uint32: "u32"}[t]


def julia_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Int32": "i32", "Int64": "i64", "UInt64": "u64", "Csize_t": "u64", "Float64": "f64", "Cint": "i32"}[t]


def test_header_compiles_as_c_and_a_c_client_links(tmp_path):
    """include/augcuda.h is a C header: it must compile as C99 (not only parse with a regex), and a plain C client
    must link against libaugcuda.so and call through it without a GPU (version + strerror + a failing ctx_create)."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "augcuda.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "client.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "augcuda.h"\n'
        "int main(void) {\n"
        "    aug_lik lik; memset(&lik, 0, sizeof lik); lik.kind = AUG_POISSON; lik.nlatent = 1; lik.p[0] = 10.0;\n"
        "    if (aug_version() != AUGCUDA_VERSION) return 1;\n"
        '    if (strcmp(aug_strerror(AUG_OK), "ok") != 0) return 2;\n'
        "    if (sizeof(aug_lik) != 56) return 3;\n"
        "    /* NULL ctx: every verb must refuse without touching a device */\n"
        "    if (aug_cavi_step(NULL, &lik, 0, NULL, NULL, NULL, 0, NULL, NULL, NULL, NULL, NULL, 0, NULL) != AUG_ERR_NOT_INIT) return 4;\n"
        "    if (aug_cavi_step_host(NULL, &lik, 0, NULL, NULL, NULL, 0, NULL, NULL, NULL, NULL, NULL, 0, NULL) != AUG_ERR_NOT_INIT) return 5;\n"
        "    if (aug_aux_sample_host(NULL, &lik, 0, 0, NULL, NULL, 0, NULL, NULL) != AUG_ERR_NOT_INIT) return 6;\n"
        '    printf("c client ok\\n");\n    return 0;\n}\n')
    exe = tmp_path / "client"
    libdir = os.path.join(ROOT, "augmentedgplikelihoods.jl_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libaugcuda.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "c client ok" in out.stdout, (out.returncode, out.stdout, out.stderr)


def test_ctypes_table_matches_the_c_prototypes_type_by_type():
    pkg = aug_pkg.load_package()
    protos = header_prototypes()
    assert set(protos) == set(pkg._lib.SIGNATURES) | {"aug_strerror"}
    for name, argtypes in pkg._lib.SIGNATURES.items():
        got = [ctypes_class(t) for t in argtypes]
        assert got == protos[name], (name, got, protos[name])


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
    assert len(calls) >= 30
    protos = header_prototypes()
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
        # type class by type class against the C prototype (pointer / Int32 / Int64 / UInt64 / Csize_t / Float64)
        got = [julia_class(p) for p in parts if p.strip()]
        assert got == protos[name], (name, got, protos[name])
    # the verbs of the path and of the rows either side of it are all bound
    for must in ("aug_cavi_step", "aug_aux_sample", "aug_expected_elbo_terms", "aug_sampled_loglik_terms",
                 "aug_sparse_cavi_sweep", "aug_sparse_marginals", "aug_sparse_precision_potential",
                 # the host-Vector methods (generic.jl:1-88 dispatches on plain vectors)
                 "aug_cavi_step_host", "aug_aux_posterior_host", "aug_expected_potential_precision_host",
                 "aug_expected_elbo_terms_host", "aug_aux_sample_host", "aug_init_aux_variables_host",
                 "aug_potential_precision_host", "aug_sampled_loglik_terms_host"):
        assert must in seen, must


def test_julia_glue_methods_cannot_be_ambiguous_with_the_reference():
    """The reference specialises its verbs on the likelihood type with loosely typed arrays (e.g. bernoulli.jl:17-22
    `aux_posterior!(qΩ, ::BernoulliLikelihood{<:LogisticLink}, ::AbstractVector, qf::AbstractVector{<:Normal})`).  A glue
    method with `lik::AbstractLikelihood` and a device / Vector array type is more specific in one argument and less in
    another: Julia raises an ambiguity MethodError at the call.  Static guard (no Julia here): every method AugCUDA.jl adds
    to a verb of the reference takes the likelihood as `lik::$L` inside the loop over REFERENCE_LIKELIHOODS, whose entries
    are the aliases the reference itself dispatches on."""
    src = open(os.path.join(ROOT, "augmentedgplikelihoods.jl_b200", "julia", "AugCUDA.jl")).read()
    code = "\n".join(l.split("#")[0] for l in src.splitlines())
    defs = re.findall(r"^\s*(?:@eval\s+)?(?:function\s+)?AGPL\.([a-z_]+!?)\(([^)]*)\)", code, flags=re.M)
    assert len(defs) >= 40
    verbs = {v for v, _ in defs}
    for must in ("init_aux_posterior", "aux_posterior!", "aux_posterior", "expected_auglik_potential_and_precision",
                 "expected_auglik_potential", "expected_auglik_precision", "expected_logtilt", "aux_kldivergence",
                 "expected_aug_loglik", "init_aux_variables", "aux_sample!", "aux_sample", "auglik_potential_and_precision",
                 "auglik_potential", "auglik_precision", "logtilt", "aug_loglik", "aux_prior", "aux_full_conditional",
                 "logdensity_def"):
        assert must in verbs, must
    for verb, args in defs:
        if "::DeviceAux" in args:          # dispatches on a handle type AugCUDA owns: nothing of the reference's can match
            continue
        assert "lik::$L" in args, (verb, args)
        assert "AbstractLikelihood" not in args, (verb, args)
    # the loop's likelihood types are the reference's own dispatch aliases, read from the reference when it is present
    table = re.search(r"const REFERENCE_LIKELIHOODS = \((.*?)\n\)", src, flags=re.S).group(1)
    for alias in ("BernoulliLikelihood{<:LogisticLink}", "AGPL.NegBinomialLikelihood", "AGPL.AugPoisson", "LaplaceLikelihood",
                  "StudentTLikelihood", "AGPL.AugHeteroGaussian", "AGPL.BijectiveLogisticSoftMaxLikelihood",
                  "AGPL.LogisticSoftMaxLikelihood"):
        assert alias in table, alias
    ref = "/root/reference/src/likelihoods"
    if os.path.isdir(ref):
        txt = "".join(open(os.path.join(ref, f)).read() for f in os.listdir(ref))
        for alias in ("NegBinomialLikelihood", "AugPoisson", "AugHeteroGaussian", "BijectiveLogisticSoftMaxLikelihood",
                      "LogisticSoftMaxLikelihood"):
            assert re.search(rf"^const {alias} = ", txt, flags=re.M), alias
