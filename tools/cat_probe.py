#!/usr/bin/env python
"""Where does the Categorical CAVI kernel lose bandwidth?  Times the fused call with subsets of its outputs
(all / state only / beta+gamma only / none) and several row counts, CUDA events on the ctx stream.
   python tools/cat_probe.py [--n 10000000] [--nl 99]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aug_pkg  # noqa: E402

A = aug_pkg.load_package()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--nl", type=str, default="99,100")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    ctx = A.Context(0)
    A.set_default_context(ctx)
    st = ctx.stream
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    for nl in [int(s) for s in args.nl.split(",")]:
        n = args.n
        lik = A.CategoricalLikelihood(nl + 1) if nl % 2 else A.CategoricalLikelihood(nl, bijective=False)
        assert lik.nlatent == nl
        mu = torch.randn(n, nl, dtype=torch.float64, device=dev, generator=g)
        var = (0.5 + torch.rand(n, nl, dtype=torch.float64, device=dev, generator=g)) ** 2
        y = torch.zeros(n, nl, dtype=torch.uint8, device=dev)
        cls = torch.randint(0, nl, (n,), device=dev, generator=g)
        y[torch.arange(n, device=dev), cls] = 1
        s0 = torch.empty(n, nl, dtype=torch.float64, device=dev)
        s1 = torch.empty(n, nl, dtype=torch.float64, device=dev)
        beta = torch.empty(nl, n, dtype=torch.float64, device=dev)
        gamma = torch.empty(nl, n, dtype=torch.float64, device=dev)
        scal = torch.zeros(8, dtype=torch.float64, device=dev)
        d = lik._desc()
        P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None

        def run(state, outs, elbo):
            A.check(ctx.lib.aug_cavi_step(ctx.h, C.byref(d), n, P(y), P(mu), P(var), n, P(s0) if state else None,
                                          P(s1) if state else None, None, P(beta) if outs else None,
                                          P(gamma) if outs else None, n, P(scal) if elbo else None))

        for name, state, outs, elbo, b in (("all", 1, 1, 0, 49), ("all+elbo", 1, 1, 1, 49), ("state only", 1, 0, 0, 33),
                                           ("beta,gamma only", 0, 1, 0, 33), ("no stores (elbo)", 0, 0, 1, 17)):
            if elbo and not lik.bijective:
                continue
            for _ in range(2):
                run(state, outs, elbo)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(st)
            for _ in range(args.reps):
                run(state, outs, elbo)
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            print(json.dumps({"nl": nl, "n": n, "case": name, "ms": round(ms, 3), "bytes_per_elt": b,
                              "GBs": round(b * n * nl / ms / 1e6)}), flush=True)
        del mu, var, y, s0, s1, beta, gamma
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
