#!/usr/bin/env python
"""Per-likelihood roofline table (not the bench contract): times the fused CAVI step and aux_sample! of every
likelihood at BASELINE sizes with CUDA events on the ctx stream and prints achieved GB/s against the measured
HBM peak.  Output is committed under profiles/ by hand.   python tools/roofline_all.py [--n 100000000]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aug_pkg  # noqa: E402

A = aug_pkg.load_package()
BYTES = {"bernoulli": 41, "negbin": 56, "poisson": 64, "laplace": 48, "studentt": 48, "hetero": 96, "cat": 50}


def timeit(fn, st, reps=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--ncat", type=int, default=10_000_000)
    ap.add_argument("--only", type=str, default="", help="comma-separated case names (for ncu captures)")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--clocks", action="store_true", help="sample nvidia-smi SM clocks / throttle reasons during the run")
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    sampler = None
    if args.clocks:
        from bench import ClockSampler
        sampler = ClockSampler(0)
        sampler.start()
    ctx = A.Context(0)
    A.set_default_context(ctx)
    st = ctx.stream
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    rows = []

    def rnd(*shape):
        return torch.randn(*shape, dtype=torch.float64, device=dev, generator=g)

    def uni(*shape):
        return torch.rand(*shape, dtype=torch.float64, device=dev, generator=g)

    cases = [
        ("bernoulli", A.BernoulliLikelihood()), ("negbin", A.NegativeBinomialLikelihood(10)),
        ("negbin_real", A.NegativeBinomialLikelihood(5.5)), ("poisson", A.PoissonLikelihood(10.0)),
        ("laplace", A.LaplaceLikelihood(1.0)), ("studentt", A.StudentTLikelihood(3.0, 1.5)),
        ("hetero", A.HeteroscedasticGaussianLikelihood(5.0)),
        ("cat_bij_K100", A.CategoricalLikelihood(100)), ("cat_K100", A.CategoricalLikelihood(100, bijective=False)),
    ]
    only = [s for s in args.only.split(",") if s]
    for name, lik in cases:
        if only and name not in only:
            continue
        cat = name.startswith("cat")
        n = args.ncat if cat else args.n
        nl = lik.nlatent
        if cat:
            mu, var, f = rnd(n, nl), (0.5 + uni(n, nl)) ** 2, rnd(n, nl)
            cls = torch.randint(0, 100, (n,), device=dev, generator=g)
            y = torch.zeros(n, nl, dtype=torch.uint8, device=dev)
            rows_ok = cls < nl
            y[rows_ok.nonzero().squeeze(1), cls[rows_ok]] = 1
        elif name == "hetero":
            mu, var, f = rnd(2, n), (0.5 + uni(2, n)) ** 2, rnd(2, n)
            y = f[0] + rnd(n)
        else:
            mu, var, f = rnd(n), (0.5 + uni(n)) ** 2, rnd(n)
            sig = torch.sigmoid(f)
            if name == "bernoulli":
                y = (uni(n) < sig).to(torch.uint8)
            elif name.startswith("negbin"):
                y = torch.poisson(10.0 * sig / (1 - sig).clamp_min(0.05)).clamp_max(100).to(torch.int64)
            elif name == "poisson":
                y = torch.poisson(10.0 * sig).to(torch.int64)
            else:
                y = f + rnd(n)
        qf = A.Normals(mu, var)
        q = A.init_aux_posterior(lik, n)
        beta = torch.empty((nl, n), dtype=torch.float64, device=dev)
        gamma = torch.empty((nl, n), dtype=torch.float64, device=dev)
        scal = torch.zeros(8, dtype=torch.float64, device=dev)
        want = name != "cat_K100"
        t_elbo = timeit(lambda: A.cavi_step_(q, lik, y, qf, want_elbo=want, out=(beta, gamma, scal if want else None)), st,
                        reps=args.reps)
        t_plain = timeit(lambda: A.cavi_step_(q, lik, y, qf, want_elbo=False, out=(beta, gamma, None)), st, reps=args.reps)
        del beta, gamma
        Ω = A.init_aux_variables(lik, n)
        t_gibbs = timeit(lambda: A.aux_sample_(Ω, lik, y, f), st, reps=3, warm=1)
        extra = {}
        if name == "hetero":                       # λ statistics, 40 / 24 B per observation
            t_l = timeit(lambda: A.hetero_lambda_stats(y, qf), st, reps=args.reps)
            t_ls = timeit(lambda: A.hetero_lambda_stats_sampled(y, f), st, reps=args.reps)
            extra = dict(ms_lambda_stats=t_l, gbs_lambda_stats=40 * n / t_l / 1e6, frac_lambda_stats=40 * n / t_l / 1e6 / peak,
                         ms_lambda_stats_sampled=t_ls, gbs_lambda_stats_sampled=24 * n / t_ls / 1e6)
        key = "cat" if cat else name.split("_")[0]
        units = n * nl if cat else n
        bpo = BYTES[key]
        rows.append(dict(likelihood=name, n=n, nlatent=nl, bytes_per_unit=bpo, ms_cavi_elbo=t_elbo, ms_cavi=t_plain,
                         gbs_cavi_elbo=bpo * units / t_elbo / 1e6, frac_elbo=bpo * units / t_elbo / 1e6 / peak,
                         gbs_cavi=bpo * units / t_plain / 1e6, frac=bpo * units / t_plain / 1e6 / peak,
                         ms_gibbs=t_gibbs, draws_per_s=units / t_gibbs * 1e3, **extra))
        print(json.dumps(rows[-1]), flush=True)
        del q, Ω, mu, var, f, y, qf
        torch.cuda.empty_cache()
    if sampler:
        print("clocks:", json.dumps(sampler.finish()))
    print("\n| likelihood | N | B/unit | CAVI+ELBO ms | GB/s | of measured | CAVI ms | of measured | Gibbs ms | draws/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['likelihood']} | {r['n']:.0e} x {r['nlatent']} | {r['bytes_per_unit']} | {r['ms_cavi_elbo']:.3f} | "
              f"{r['gbs_cavi_elbo']:.0f} | {r['frac_elbo']:.2f} | {r['ms_cavi']:.3f} | {r['frac']:.2f} | "
              f"{r['ms_gibbs']:.2f} | {r['draws_per_s']:.3g} |")


if __name__ == "__main__":
    main()
