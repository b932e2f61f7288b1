#!/usr/bin/env python
"""bench.py — headline benchmark of the augmentation hot path (BASELINE.json metric):
"CAVI obs/sec & Polya-Gamma draws/sec, N=1e8 fp64, 1/2/4/8 B200 vs CPU".

A step = one pass of the hot path over one batch of N synthetic Bernoulli-logistic observations PER GPU:
  (1) the fused CAVI update  — aux_posterior! + expected_auglik_potential_and_precision + the
      expected_logtilt / aux_kldivergence partial sums (one kernel), and
  (2) the Gibbs draw         — aux_sample!: one PG(1, |f_i|) draw per observation (one kernel),
  (3) for N > 1 GPUs, the all-reduce of the 64-byte scalar block: by default FUSED into kernel (1) over the
      peer-memory mailbox (NVLink stores + epoch flags, include/augcuda.h); `--collective nccl` issues the
      library's ncclAllReduce after the step instead.
`value` = observations through BOTH halves per second, whole job, inputs resident in HBM;
`parts` gives each half on its own (CAVI obs/s, PG draws/s) from CUDA events inside the timed region.
`e2e` = the same step through the host-buffer C-ABI calls (aug_cavi_step_host + aug_aux_sample_host)
with pinned HOST inputs/outputs, H2D and D2H inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n OBS_PER_GPU] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CAVI obs/sec & Polya-Gamma draws/sec (Bernoulli-logistic; one fused CAVI update + one PG draw per obs)"
UNIT = "obs/s"
BYTES_CAVI = 41   # R y 1 + mu 8 + var 8, W c 8 + beta 8 + gamma 8
BYTES_GIBBS = 16  # R f 8, W omega 8


_OUT = None


def emit(line):
    out = _OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads(orc):
    """all host cores this process may use (torchrun exports OMP_NUM_THREADS=1: the oracle's num_threads clause overrides it)"""
    try:
        return max(orc.max_threads(), len(os.sched_getaffinity(0)))
    except Exception:
        return orc.max_threads()


def cpu_inputs(n, seed=1):
    import numpy as np
    rng = np.random.default_rng(seed)
    mu = rng.standard_normal(n)
    var = (0.5 + rng.random(n)) ** 2
    f = rng.standard_normal(n)
    y = (rng.random(n) < 1 / (1 + np.exp(-f))).astype(np.uint8)
    return y, mu, var, f


def cpu_reference(n, threads, seed=1, repeats=1, inputs=None):
    """The reference's CPU path (C++ oracle restatement — Julia is not installable here), structured like
    the reference: separate passes + temporaries, sequential ELBO sums, per-element Devroye sampler."""
    from oracle import orc
    orc.lib()
    orc.set_threads(threads)
    y, mu, var, f = inputs if inputs is not None else cpu_inputs(n, seed)
    lik = orc.make_lik(orc.BERNOULLI)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        rc, st, b, g, seq, comp = orc.cavi_step(lik, y, mu, var)
        t1 = time.perf_counter()
        om, nv = orc.aux_sample(lik, seed, y, f)
        t2 = time.perf_counter()
        cur = (t2 - t0, t1 - t0, t2 - t1)
        if best is None or cur[0] < best[0]:
            best = cur
    return best


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (all threads).
    Each step is a bounded sample of the workload, sized from a probe so that K steps take about a minute."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    orc.lib()
    threads = host_threads(orc)
    probe = cpu_inputs(1_000_000)
    for _ in range(max(1, args.warmup)):
        tp = cpu_reference(1_000_000, threads, inputs=probe)[0]
    rate = 1_000_000 / tp
    n = int(min(args.ref_n, max(1_000_000, 60.0 * rate / max(1, args.steps))))
    inputs = cpu_inputs(n)
    tot = cavi = gib = 0.0
    for k in range(args.steps):
        t, tc, tg = cpu_reference(n, threads, seed=k + 1, inputs=inputs)
        tot += t
        cavi += tc
        gib += tg
    value = n * args.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]+configs[0]: Bernoulli-logistic fused CAVI update + PG Gibbs aux_sample!, "
                               f"bounded sample of {n} observations per step on the host CPU"},
        "parts": {"cavi_obs_per_s": n * args.steps / cavi, "pg_draws_per_s": n * args.steps / gib},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} obs/step x {args.steps} steps, OpenMP over {threads} threads; C++ restatement "
                                   "of the Julia reference (Julia is not available in this image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def fp64_peak():
    """Measured FP64 tensor (DMMA m8n8k4) peak of this GPU: tools/fp64_peak (built by build()) run live, else the
    number recorded on this pool (profiles/fp64_peak_r1f.jsonl)."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    if os.path.exists(exe):
        try:
            out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
            best = max(json.loads(l)["tflops"] for l in out.splitlines() if '"dmma' in l)
            return best, "measured live (tools/fp64_peak: mma.sync.m8n8k4.f64, all SMs)"
        except Exception:
            pass
    return 37.1, "recorded (profiles/fp64_peak_r1f.jsonl)"


def sparse_traffic(m, n):
    """dram bytes per launch from the ncu --set full capture recorded in profiles/traffic.json (m = 128 only)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["sparse_sweep_kernel_m128_fused"]
        return t["bytes_per_obs"] * n if m == 128 else None
    except Exception:
        return None


def sparse_leg(A, ctx, torch, dist, world, dev, args, rank):
    """aug_sparse_cavi_sweep on m inducing points: marginals -> aux_posterior! -> E[beta], E[gamma], ELBO sums ->
    P = kappa Diag(gamma) kappa^T, rhs = kappa beta, next to the library composition (cuBLAS DGEMMs through torch +
    our streaming CAVI kernel).  FP64-pipe bound: roofline against the measured DMMA peak."""
    m, n = args.sparse_m, args.sparse_n
    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)
    kappa = torch.randn(n, m, dtype=torch.float64, device=dev, generator=g) / m ** 0.5
    Aq = torch.randn(m, m, dtype=torch.float64, device=dev, generator=g)
    B = Aq @ Aq.T
    B = B * (0.3 * m / torch.trace(B))
    B = 0.5 * (B + B.T)
    kdiag = ((kappa @ B) * kappa).sum(1) + 0.3 + 0.5 * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    mvec = torch.randn(m, dtype=torch.float64, device=dev, generator=g)
    y = (torch.rand(n, device=dev, generator=g) < 0.5).to(torch.uint8)
    lik = A.BernoulliLikelihood()
    q = A.init_aux_posterior(lik, n)
    st = ctx.stream
    res = {}

    def fused():
        res["f"] = A.sparse_cavi_sweep_(q, lik, y, kappa, mvec, B, kdiag, want_elbo=True, want_potentials=True)

    def library():
        mu = kappa @ mvec
        var = kdiag - ((kappa @ B) * kappa).sum(1)
        _, b, gm, sc = A.cavi_step_(q, lik, y, A.Normals(mu, var))
        res["l"] = ((kappa * gm[0][:, None]).T @ kappa, kappa.T @ b[0], sc)

    def timed(fn, k):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), (ctx.launches() - l0) // k

    k = max(3, min(args.steps, 20))
    ms_f, launches = timed(fused, k)
    ms_l, _ = timed(library, max(2, k // 4))
    Pf, rf, sf = res["f"][0], res["f"][1], res["f"][2]
    Pl, rl, sl = res["l"]
    agree = {"P": float(((Pf - Pl).abs().max() / Pl.abs().max()).item()),
             "rhs": float(((rf - rl).abs().max() / rl.abs().max()).item()),
             "elbo": float(((sf[2] - sl[2]).abs() / sl[2].abs()).item())}
    flops = 2.0 * m * m + 6.0 * m            # symmetric quadratic form + symmetric rank-1 update + mu + rhs, per obs
    peak, src = fp64_peak() if rank == 0 else (37.1, "")
    ach = flops * n / (ms_f * 1e-3) / 1e12
    return {"workload": f"sparse-GP CAVI iteration (Bernoulli), m={m} inducing points, {n} observations per GPU: "
                        "aug_sparse_cavi_sweep = SVGP marginals + aux_posterior! + E[beta],E[gamma] + ELBO sums + "
                        "P = kappa Diag(gamma) kappa^T, rhs = kappa beta in one pass over kappa",
            "m": m, "obs_per_gpu": n, "l2": f"kappa is {8.0 * m * n / 1e9:.1f} GB per sweep >> 126 MB L2 (no flush needed)",
            "ms_per_sweep": ms_f, "obs_per_s": n * world / (ms_f * 1e-3),
            "gpu_launches_per_sweep": int(launches),
            "roofline": {"kernel": "sparse_sweep_kernel<128, FUSED, BERNOULLI> (DMMA m8n8k4)", "bound": "tensor",
                         "note": "fp64 tensor pipe (tcgen05 has no f64 kind); HBM traffic 8m B/obs is 10% of the HBM roofline",
                         "achieved": ach, "peak": peak, "peak_source": src, "unit": "TFLOP/s", "frac": ach / peak,
                         "flops_per_obs": flops, "traffic": sparse_traffic(m, n)},
            "library_composition": {"what": "torch (cuBLAS DGEMM) kappa@m, kappa@B, row dots, (kappa*gamma)^T@kappa, "
                                            "kappa^T@beta + aug_cavi_step", "ms": ms_l, "speedup": ms_l / ms_f,
                                    "max_rel_diff": agree}}


def sparse_cpu(m, n=20000):
    """the oracle's separate-pass restatement of the same iteration on the host cores (bounded sample)"""
    import numpy as np
    from oracle import orc
    orc.lib()
    threads = host_threads(orc)
    orc.set_threads(threads)
    rng = np.random.default_rng(0)
    kappa = rng.standard_normal((n, m)) / np.sqrt(m)
    Aq = rng.standard_normal((m, m))
    B = Aq @ Aq.T
    B *= 0.3 * m / np.trace(B)
    B = 0.5 * (B + B.T)
    kdiag = np.einsum("ti,ij,tj->t", kappa, B, kappa) + 0.3 + 0.5 * rng.random(n)
    mvec = rng.standard_normal(m)
    y = (rng.random(n) < 0.5).astype(np.uint8)
    lik = orc.make_lik(orc.BERNOULLI)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        rc, _ = orc.sparse_cavi_sweep(lik, y, kappa, mvec, B, kdiag)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    # what the user's Julia loop would run: BLAS for the matrix products (numpy -> OpenBLAS / MKL, all cores) around
    # the per-observation passes of the oracle (double precision throughout)
    nb = 20 * n
    kb = np.tile(kappa, (20, 1)); yb = np.tile(y, 20); kdb = np.tile(kdiag, 20)
    bestb = None
    for _ in range(2):
        t0 = time.perf_counter()
        mu = kb @ mvec
        var = kdb - np.einsum("ti,ti->t", kb @ B, kb)
        rc, st, be, ga, seq, comp = orc.cavi_step(lik, yb, mu, var)
        P = (kb * ga[0][:, None]).T @ kb
        rhs = kb.T @ be[0]
        dt = time.perf_counter() - t0
        bestb = dt if bestb is None else min(bestb, dt)
    return {"value": nb / bestb, "unit": "obs/s", "cores": threads, "kind": "port",
            "sample": f"{nb} observations, m={m}: BLAS (numpy) matrix products + the oracle's separate per-observation passes, "
                      f"double precision, {threads} threads, best of 2",
            "long_double_oracle": {"value": n / best, "unit": "obs/s",
                                   "sample": f"{n} observations, the parity oracle itself (long double loops, OpenMP {threads})"}}


def main():
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner under
    # NCCL_DEBUG, torchrun notices) is sent to stderr; the result line is written to the saved descriptor.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=100_000_000, help="observations per GPU (weak scaling)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-n", type=int, default=20_000_000, help="CPU sample size per step of the reference arm")
    ap.add_argument("--cpu-n", type=int, default=20_000_000, help="CPU sample of the cpu_baseline leg")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: all-reduce of the scalar block fused into the CAVI kernel over peer memory (p2p) or "
                         "a separate ncclAllReduce (nccl)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sparse", action="store_true", help="skip the secondary sparse-GP sweep measurement")
    ap.add_argument("--sparse-m", type=int, default=128)
    ap.add_argument("--sparse-n", type=int, default=2_000_000, help="observations per GPU of the sparse sweep leg")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import aug_pkg
    A = aug_pkg.load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = A.Context(local_rank)
    A.set_default_context(ctx)
    if world > 1:
        if args.collective == "nccl":
            A.dist.init_comm(ctx)
        else:
            A.dist.init_p2p(ctx, fused=True)
    n = args.n
    i0 = rank * n                                    # global index of this rank's first observation
    lik = A.BernoulliLikelihood()
    dev = torch.device("cuda", local_rank)

    # synthetic inputs of SURVEY §8(d), generated on the device
    g = torch.Generator(device=dev)
    g.manual_seed(1 + rank)
    mu = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    var = (0.5 + torch.rand(n, dtype=torch.float64, device=dev, generator=g)) ** 2
    f = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    y = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) < torch.sigmoid(f)).to(torch.uint8)
    qf = A.Normals(mu, var)
    q = A.init_aux_posterior(lik, n)
    beta = torch.empty((1, n), dtype=torch.float64, device=dev)
    gamma = torch.empty((1, n), dtype=torch.float64, device=dev)
    scal = torch.zeros(8, dtype=torch.float64, device=dev)
    Ω = A.AuxSamples(torch.empty(n, dtype=torch.float64, device=dev), None)
    ctx.seed(2026, 0)
    torch.cuda.synchronize()

    st = ctx.stream                                  # every kernel of the step is launched on this stream

    def step(ev=None):
        if ev:
            ev[0].record(st)
        A.cavi_step_(q, lik, y, qf, want_elbo=True, out=(beta, gamma, scal))
        if ev:
            ev[1].record(st)
        A.aux_sample_(Ω, lik, y, f, i0=i0)
        if ev:
            ev[2].record(st)
        if world > 1 and args.collective == "nccl":
            A.dist.allreduce_scalars_(ctx, scal)          # p2p: already summed over ranks inside the CAVI kernel
        if ev:
            ev[3].record(st)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    time.sleep(0.25)                                 # every rank waits for the sampler's first rows ...
    if world > 1:
        dist.barrier()                               # ... and all ranks enter the timed region together
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    l0 = ctx.launches()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(st)
    for k in range(args.steps):
        step(evs[k])
    t_end.record(st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = ctx.launches() - l0
    clocks = sampler.finish() if rank == 0 else None
    ms_total = t_start.elapsed_time(t_end)
    ms_cavi = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    ms_gibbs = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    ms_coll = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    t = torch.tensor([ms_total, ms_cavi, ms_gibbs, ms_coll], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)      # max over ranks, device-timed
    ms_total, ms_cavi, ms_gibbs, ms_coll = t.tolist()
    elbo = float(scal[2].item())

    # ---------------- e2e: host buffers through the plugin calls, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        hy = torch.empty(n, dtype=torch.uint8).pin_memory()
        hmu = torch.empty(n, dtype=torch.float64).pin_memory()
        hvar = torch.empty(n, dtype=torch.float64).pin_memory()
        hf = torch.empty(n, dtype=torch.float64).pin_memory()
        hy.copy_(y); hmu.copy_(mu); hvar.copy_(var); hf.copy_(f)
        hc = torch.empty(n, dtype=torch.float64).pin_memory()
        hb = torch.empty(n, dtype=torch.float64).pin_memory()
        hg = torch.empty(n, dtype=torch.float64).pin_memory()
        hw = torch.empty(n, dtype=torch.float64).pin_memory()
        hs = (C.c_double * 8)()
        d = lik._desc()
        P = lambda tt: C.c_void_p(tt.data_ptr())

        def e2e_step():
            A.check(ctx.lib.aug_cavi_step_host(ctx.h, C.byref(d), n, P(hy), P(hmu), P(hvar), 0, P(hc), None, None,
                                               P(hb), P(hg), n, hs))
            A.check(ctx.lib.aug_aux_sample_host(ctx.h, C.byref(d), n, i0, P(hy), P(hf), 0, P(hw), None))

        torch.cuda.synchronize()
        e2e_step()                                     # warm-up (allocates the staging slots)
        if world > 1:
            dist.barrier()
        ksteps = max(1, min(args.steps, 5))
        w0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()                                 # returns after results are on the host
        w1 = time.perf_counter()
        te = torch.tensor([w1 - w0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = te.item() / ksteps
        assert abs(hs[2] - elbo) <= 1e-9 * abs(elbo) or world > 1, (hs[2], elbo)
        e2e = {"value": n * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * (1 + 8 + 8 + 1 + 8),
               "d2h_bytes_per_step": n * (24 + 8) + 64, "ms_per_step": 1e3 * e2e_s, "steps": ksteps,
               "path": "aug_cavi_step_host + aug_aux_sample_host (pinned host buffers, 3-slot H2D/kernel/D2H pipeline)",
               "pcie_GBs": (n * (1 + 8 + 8 + 1 + 8) + n * (24 + 8) + 64) / e2e_s * 1e-9,
               "pcie_bound": "measured on this pool (tools/pcie_probe.py, profiles/e2e_pcie_probe_r1k.txt): 55.5 GB/s H2D, "
                             "57.0 D2H, 49.8 each way when both run: the D2H side (32 B/obs) bounds the step at ~64 ms"}
        del hy, hmu, hvar, hf, hc, hb, hg, hw

    # ---------------- secondary leg: one sparse-GP CAVI iteration as one pass over κ
    sparse = None
    if not args.no_sparse:
        sparse = sparse_leg(A, ctx, torch, dist, world, dev, args, rank)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- cpu_baseline: the oracle, timed on this box's host cores (bounded sample)
    cpu = None
    if not args.no_cpu:
        from oracle import orc
        orc.lib()
        threads = host_threads(orc)
        t1, c1, g1 = cpu_reference(min(args.cpu_n // 10, 2_000_000), 1)
        n1 = min(args.cpu_n // 10, 2_000_000)
        tt, tc, tg = cpu_reference(args.cpu_n, threads, repeats=2)
        cpu = {"value": args.cpu_n / tt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_n} Bernoulli obs, same step (separate-pass CAVI + Devroye PG(1) draws), OpenMP "
                         f"{threads} threads, best of 2; C++ restatement of the Julia reference",
               "cavi_obs_per_s": args.cpu_n / tc, "pg_draws_per_s": args.cpu_n / tg,
               "single_thread": {"value": n1 / t1, "cavi_obs_per_s": n1 / c1, "pg_draws_per_s": n1 / g1,
                                 "sample": f"{n1} obs, 1 thread (the reference is single-threaded)"}}

    peak, peak_src = load_peaks()
    ach = BYTES_CAVI * n / (ms_cavi * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("cavi_bernoulli_bytes_per_launch_at_1e8")
            if traffic is not None and n != 100_000_000:
                traffic = traffic * n / 100_000_000
        except Exception:
            traffic = None
    ms_step = ms_total / args.steps
    line = {
        "metric": METRIC, "value": n * world / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] (Bernoulli-logistic Gibbs aux_sample! PG draws, N=1e8 fp64) + "
                               "configs[0]'s fused CAVI update at the same N",
                   "obs_per_gpu": n, "global_obs": n * world, "likelihood": "BernoulliLikelihood(LogisticLink)",
                   "l2": "inputs+outputs 5.7 GB per step >> 126 MB L2 (no flush needed)",
                   "sharding": f"contiguous observation blocks, {world} rank(s); only the 8-double scalar block "
                               "is all-reduced",
                   "collective": ("none (1 rank)" if world == 1 else
                                  "fused into cavi_tma_kernel's finaliser over the peer-memory mailbox (NVLink)"
                                  if args.collective == "p2p" else "ncclAllReduce(sum, double, 8) after the step")},
        "parts": {"cavi_obs_per_s": n * world / (ms_cavi * 1e-3), "pg_draws_per_s": n * world / (ms_gibbs * 1e-3),
                  "ms_cavi": ms_cavi, "ms_gibbs": ms_gibbs, "ms_allreduce": ms_coll},
        "roofline": {"kernel": "cavi_tma_kernel<BERNOULLI, ELBO> (aux_posterior! + expected potential/precision + ELBO sums)", "bound": "hbm", "achieved": ach,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                     "frac_of_8TBs_nominal": ach / 8000.0, "bytes_per_obs": BYTES_CAVI, "traffic": traffic},
        "roofline_gibbs": {"kernel": "pg1_compact_kernel (warp-compacted Devroye PG(1,c))", "bound": "fp64 pipe / issue (not HBM)",
                           "achieved": BYTES_GIBBS * n / (ms_gibbs * 1e-3) / 1e9, "unit": "GB/s",
                           "frac_hbm": BYTES_GIBBS * n / (ms_gibbs * 1e-3) / 1e9 / peak,
                           "pg_draws_per_s_per_gpu": n / (ms_gibbs * 1e-3)},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "elbo_check": elbo,
    }
    if sparse:
        if not args.no_cpu:
            sparse["cpu_baseline"] = sparse_cpu(args.sparse_m)
        line["sparse_sweep"] = sparse
    if cpu:
        line["speedup_vs_cpu"] = {"value_vs_all_threads": line["value"] / cpu["value"],
                                  "pg_draws_vs_single_thread": line["parts"]["pg_draws_per_s"] /
                                  cpu["single_thread"]["pg_draws_per_s"]}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
