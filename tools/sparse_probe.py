#!/usr/bin/env python
"""Timing probe of the sparse-GP sweep (aug_sparse.cu) against the library composition (cuBLAS DGEMM via torch +
our streaming CAVI kernel): obs/s, useful FP64 TFLOP/s (3 m² flops per observation) and κ GB/s.
  python tools/sparse_probe.py [--m 128 64 32 16] [--bytes 4e9] [--reps 5]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aug_pkg  # noqa: E402


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, nargs="+", default=[128, 64, 32, 16])
    ap.add_argument("--bytes", type=float, default=4e9)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--peak-tflops", type=float, default=37.1)
    ap.add_argument("--no-lib", action="store_true")
    a = ap.parse_args()
    A = aug_pkg.load_package()
    ctx = A.Context(0)
    A.set_default_context(ctx)
    lik = A.BernoulliLikelihood()
    for m in a.m:
        n = int(a.bytes / (8 * m))
        g = torch.Generator(device="cuda").manual_seed(m)
        kappa = torch.randn(n, m, dtype=torch.float64, device="cuda", generator=g) / m ** 0.5
        Aq = torch.randn(m, m, dtype=torch.float64, device="cuda", generator=g)
        B = Aq @ Aq.T
        B = B * (0.3 * m / torch.trace(B))
        B = 0.5 * (B + B.T)
        kdiag = ((kappa @ B) * kappa).sum(1) + 0.3 + 0.5 * torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
        mvec = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
        y = (torch.rand(n, device="cuda", generator=g) < 0.5).to(torch.uint8)
        if True:
            q = A.init_aux_posterior(lik, n)
            res = {}

            def fused():
                res["f"] = A.sparse_cavi_sweep_(q, lik, y, kappa, mvec, B, kdiag, want_potentials=True)

            def prod():
                res["p"] = A.sparse_marginals(kappa, mvec, B, kdiag)

            def fused_noelbo():
                res["fn"] = A.sparse_cavi_sweep_(q, lik, y, kappa, mvec, B, kdiag, want_elbo=False)

            t_f = timeit(fused, a.reps)
            t_fn = timeit(fused_noelbo, a.reps)
            t_p = timeit(prod, a.reps)
            beta, gamma = res["f"][4]

            def cons():
                res["c"] = A.sparse_precision_potential(kappa, gamma, beta)

            t_c = timeit(cons, a.reps)
            line = {"m": m, "n": n, "ms_fused": t_f, "ms_fused_no_elbo": t_fn, "ms_producer": t_p, "ms_consumer": t_c,
                    "obs_per_s": n / t_f * 1e3, "tflops_fused": 3.0 * m * m * n / t_f * 1e-9,
                    "frac_fp64_peak": 3.0 * m * m * n / t_f * 1e-9 / a.peak_tflops,
                    "tflops_producer": 2.0 * m * m * n / t_p * 1e-9, "tflops_consumer": 1.0 * m * m * n / t_c * 1e-9,
                    "kappa_GBs_fused": 8.0 * m * n / t_f * 1e-6}
            if not a.no_lib:
                def lib():
                    mu = kappa @ mvec
                    T = kappa @ B
                    var = kdiag - (T * kappa).sum(1)
                    _, b, gm, sc = A.cavi_step_(q, lik, y, A.Normals(mu, var))
                    P = (kappa * gm[0][:, None]).T @ kappa
                    rhs = kappa.T @ b[0]
                    res["l"] = (P, rhs, sc)

                t_l = timeit(lib, max(2, a.reps // 2))
                P, rhs, sc = res["l"]
                Pf, rf, sf = res["f"][0], res["f"][1], res["f"][2]
                line.update({"ms_library_composition": t_l, "speedup_vs_library": t_l / t_f,
                             "max_rel_P_vs_library": float(((P - Pf).abs().max() / P.abs().max()).item()),
                             "max_rel_rhs_vs_library": float(((rhs - rf).abs().max() / rhs.abs().max()).item()),
                             "elbo_rel_vs_library": float(((sc[2] - sf[2]).abs() / sc[2].abs()).item())})
                del P, rhs, res["l"]
        print(json.dumps(line), flush=True)
        del kappa, y, kdiag, q, res
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
