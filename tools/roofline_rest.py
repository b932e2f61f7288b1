#!/usr/bin/env python
"""HBM roofline of the remaining verbs of the path: init_aux_variables, sampled
auglik_potential_and_precision, logtilt, aug_loglik, expected_elbo_terms / expected potential-precision from a state.
Streaming maps / map-reduces timed with CUDA events on the ctx stream; bytes are the algorithmic R + W per observation.
   python tools/roofline_rest.py [--n 100000000]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aug_pkg  # noqa: E402

A = aug_pkg.load_package()


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
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    n = args.n
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    ctx = A.Context(0)
    A.set_default_context(ctx)
    st = ctx.stream
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    rnd = lambda *s: torch.randn(*s, dtype=torch.float64, device=dev, generator=g)
    uni = lambda *s: torch.rand(*s, dtype=torch.float64, device=dev, generator=g)
    # (name, likelihood, y, bytes: potential R/W, logtilt R, elbo-from-state R, expected-from-state R/W)
    cases = {
        "bernoulli": (A.BernoulliLikelihood(), lambda f: (uni(n) < torch.sigmoid(f)).to(torch.uint8), dict(pot=1 + 8 + 16, lt=1 + 8 + 8, elbo=1 + 16 + 8, exp=1 + 8 + 16)),
        "poisson": (A.PoissonLikelihood(10.0), lambda f: torch.poisson(10.0 * torch.sigmoid(f)).to(torch.int64), dict(pot=8 + 16 + 16, lt=8 + 8 + 16, elbo=8 + 16 + 16, exp=8 + 16 + 16)),
        "studentt": (A.StudentTLikelihood(3.0, 1.5), lambda f: f + rnd(n), dict(pot=8 + 8 + 16, lt=8 + 8 + 8, elbo=8 + 16 + 8, exp=8 + 8 + 16)),
    }
    rows = []
    for name, (lik, mk_y, by) in cases.items():
        if args.only and name not in args.only.split(","):
            continue
        f, mu, var = rnd(n), rnd(n), (0.5 + uni(n)) ** 2
        y = mk_y(f)
        qf = A.Normals(mu, var)
        q = A.aux_posterior(lik, y, qf)
        Ω = A.aux_sample(A.AugPhilox(1, 0), lik, y, f)
        nb = 16 if Ω.n is not None else 8
        r = {"likelihood": name, "n": n}
        t = timeit(lambda: A.init_aux_variables(lik, n), st, reps=3, warm=1)
        r["init_aux_variables"] = dict(ms=t, draws_per_s=n / t * 1e3)
        for key, fn, b in [("auglik_potential_and_precision", lambda: A.auglik_potential_and_precision(lik, Ω, y, f), by["pot"]),
                           ("logtilt", lambda: A.logtilt(lik, Ω, y, f), by["lt"]),
                           ("expected_elbo_terms (from state)", lambda: A.expected_logtilt(lik, q, y, qf), by["elbo"]),
                           ("expected_auglik_potential_and_precision (from state)",
                            lambda: A.expected_auglik_potential_and_precision(lik, q, y, qf), by["exp"])]:
            t = timeit(fn, st)
            r[key] = dict(ms=t, bytes_per_obs=b, GBs=b * n / t / 1e6, frac=b * n / t / 1e6 / peak)
        t = timeit(lambda: A.aug_loglik(lik, Ω, y, f), st, reps=2, warm=1)
        r["aug_loglik (PG log-density series)"] = dict(ms=t, obs_per_s=n / t * 1e3, bound="fp64 / issue (alternating series, ~200 exp per obs)")
        rows.append(r)
        print(json.dumps(r), flush=True)
        del f, mu, var, y, qf, q, Ω
        torch.cuda.empty_cache()
    print("\n| likelihood | verb | ms | GB/s | of measured HBM |")
    print("|---|---|---|---|---|")
    for r in rows:
        for k, v in r.items():
            if isinstance(v, dict):
                print(f"| {r['likelihood']} | {k} | {v['ms']:.3f} | {v.get('GBs', float('nan')):.0f} | {v.get('frac', float('nan')):.2f} |")


if __name__ == "__main__":
    main()
