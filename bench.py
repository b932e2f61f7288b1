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
# pg1_compact_kernel: warp-instructions per draw = ncu smsp__inst_executed.sum / draws at the bench's input law (f ~ N(0,1)),
# profiles/r2zz_ncu_summary.txt (1 126 289 538 for 1e8 draws; 28.4 of 32 lanes active per instruction)
PG1_WARP_INSTR_PER_DRAW = 11.263


def issue_roofline(draws, ms, sms, clocks):
    """Issue-slot roofline of the sampler, the time-dominant and issue-bound kernel of the step: achieved = executed
    warp-instructions per second from the LIVE kernel time and the instruction count ncu measured for this kernel; peak = SMs x 4
    schedulers x SM clock (one warp-instruction per scheduler per cycle, clock = the median sampled under load)."""
    try:
        mhz = float((clocks or {}).get("sm_mhz") or 0.0)
        if mhz <= 0.0 or not sms or ms <= 0.0:
            return None
        ach = PG1_WARP_INSTR_PER_DRAW * draws / (ms * 1e-3)
        peak = float(sms) * 4.0 * mhz * 1e6
        return {"bound": "issue", "achieved": ach, "peak": peak, "unit": "warp-instructions/s", "frac": ach / peak,
                "warp_instructions_per_draw": PG1_WARP_INSTR_PER_DRAW,
                "source": "instruction count: ncu (profiles/r2zz_ncu_summary.txt, issue-active 71.1 % under ncu); "
                          "time and clock: this run"}
    except Exception:
        return None


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
            os.chmod(exe, 0o755)                 # a snapshot copy may drop the executable bit
            out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
            best = max(json.loads(l)["tflops"] for l in out.splitlines() if '"kernel": "dmma' in l)
            return best, "measured live (tools/fp64_peak: mma.sync.m8n8k4.f64, all SMs)"
        except Exception as e:
            return 37.1, f"recorded (profiles/fp64_peak_r1l.jsonl); the live run of tools/fp64_peak failed: {type(e).__name__}: {e}"
    return 37.1, "recorded (profiles/fp64_peak_r1l.jsonl); tools/fp64_peak is not built on this box (build() makes it)"


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
        ctx.flush()       # split-phase mode: no sampling launch follows here to gather the published sums
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
    Pl, rl = Pl.clone(), rl.clone()
    if world > 1:
        # the fused sweep returns the sums over ALL ranks (exchanged inside its finalise launch, or by the library's
        # ncclAllReduce): the composition must be summed over ranks too before the two can be compared.  In fused
        # mode cavi_step_ already returned the global ELBO sums; in nccl mode they are rank-local.
        dist.all_reduce(Pl)
        dist.all_reduce(rl)
        if not ctx.fused:
            sl = sl.clone()
            dist.all_reduce(sl)
    scaleP = (torch.sqrt(torch.diag(Pl))[:, None] * torch.sqrt(torch.diag(Pl))[None, :])
    agree = {"P": float(((Pf - Pl).abs() / scaleP).max().item()),
             "rhs": float(((rf - rl).abs().max() / rl.abs().max()).item()),
             "elbo": float(((sf[2] - sl[2]).abs() / sl[2].abs()).item())}
    # cuBLAS sums in another order: agreement to rounding of an n-term fp64 sum, far below the 1e-12 parity bar
    assert agree["P"] <= 1e-11 and agree["rhs"] <= 1e-11 and agree["elbo"] <= 1e-12, agree
    agree["ok"] = True
    flops = 2.0 * m * m + 6.0 * m            # symmetric quadratic form + symmetric rank-1 update + mu + rhs, per obs
    peak, src = fp64_peak() if rank == 0 else (37.1, "")
    if world > 1:
        pk = torch.tensor([peak], dtype=torch.float64, device=dev)
        dist.broadcast(pk, 0)
        peak = pk.item()
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


# ---------------------------------------------------------------------------------------------- workloads
# (name, BASELINE config, likelihood factory, CAVI bytes / unit with the reference's y copy materialised, Gibbs bytes / unit)
#   bytes per SURVEY §8(d): CAVI = R y, mu, var + W state, beta, gamma;  Gibbs = R y, f (+g) + W omega (+n)
WORKLOADS = {
    "bernoulli": dict(cfg="configs[0]+[1]", bytes_cavi=41, bytes_gibbs=16, n=100_000_000),
    "negbin": dict(cfg="configs[2]", bytes_cavi=56, bytes_gibbs=24, n=100_000_000),
    "poisson": dict(cfg="configs[2]", bytes_cavi=64, bytes_gibbs=32, n=100_000_000),
    "studentt": dict(cfg="configs[3]", bytes_cavi=48, bytes_gibbs=24, n=100_000_000),
    "laplace": dict(cfg="configs[3]", bytes_cavi=48, bytes_gibbs=24, n=100_000_000),
    "hetero": dict(cfg="configs[3]", bytes_cavi=96, bytes_gibbs=40, n=100_000_000),
    "categorical": dict(cfg="configs[4]", bytes_cavi=50, bytes_gibbs=25, n=10_000_000),   # per (obs, class) element, K = 100
}
CAVI_KERNEL = {"bernoulli": "cavi_tma_kernel<BERNOULLI, ELBO>", "negbin": "cavi_tma_kernel<NEGBIN, ELBO>",
               "poisson": "cavi_tma_kernel<POISSON, ELBO>", "studentt": "cavi_tma_kernel<STUDENTT, ELBO>",
               "laplace": "cavi_tma_kernel<LAPLACE, ELBO>", "hetero": "cavi_tma_kernel<HETERO, ELBO>",
               "categorical": "cat_row_kernel<ELBO> (row-aligned two-warp tiles, bulk-async ring)"}
GIBBS_KERNEL = {"bernoulli": "pg1_compact_kernel (warp-compacted Devroye PG(1,c))",
                "negbin": "pgb_kernel<NEGBIN> (warp-compacted PG(y+r,c): certified Gamma convolution / exact pieces)",
                "poisson": "pgb_kernel<POISSON> (Poisson draw + warp-compacted PG(y+n,c))",
                "studentt": "aux_sample_kernel<STUDENTT> (Marsaglia-Tsang Gamma)",
                "laplace": "aux_sample_kernel<LAPLACE> (inverse Gaussian)",
                "hetero": "pgb_kernel<HETERO> (Poisson draw + exact PG(n+1/2,c))",
                "categorical": "cat_gibbs_kernel (geometric row total + class picks of the NegativeMultinomial, per-warp PG queues)"}


def make_lik(A, name):
    return {"bernoulli": lambda: A.BernoulliLikelihood(), "negbin": lambda: A.NegativeBinomialLikelihood(10),
            "poisson": lambda: A.PoissonLikelihood(10.0), "studentt": lambda: A.StudentTLikelihood(3.0, 1.5),
            "laplace": lambda: A.LaplaceLikelihood(1.0), "hetero": lambda: A.HeteroscedasticGaussianLikelihood(5.0),
            "categorical": lambda: A.CategoricalLikelihood(100)}[name]()


def make_orc_lik(orc, name):
    return {"bernoulli": lambda: orc.make_lik(orc.BERNOULLI), "negbin": lambda: orc.make_lik(orc.NEGBIN, 10, r_is_int=True),
            "poisson": lambda: orc.make_lik(orc.POISSON, 10.0), "studentt": lambda: orc.make_lik(orc.STUDENTT, 3.0, 1.5),
            "laplace": lambda: orc.make_lik(orc.LAPLACE, 1.0), "hetero": lambda: orc.make_lik(orc.HETERO, 5.0),
            "categorical": lambda: orc.make_lik(orc.CAT_BIJ, nlatent=99)}[name]()


def make_inputs(torch, name, lik, n, dev, seed):
    """synthetic inputs of SURVEY §8(d), generated on the device: mu ~ N(0,1), var = (0.5+U)^2, f ~ N(0,1), y from the likelihood"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rnd = lambda *sh: torch.randn(*sh, dtype=torch.float64, device=dev, generator=g)
    uni = lambda *sh: torch.rand(*sh, dtype=torch.float64, device=dev, generator=g)
    nl = lik.nlatent
    if name == "categorical":
        mu, var, f = rnd(n, nl), (0.5 + uni(n, nl)) ** 2, rnd(n, nl)
        cls = torch.randint(0, nl + 1, (n,), device=dev, generator=g)
        y = torch.zeros(n, nl, dtype=torch.uint8, device=dev)
        ok = (cls < nl).nonzero().squeeze(1)
        y[ok, cls[ok]] = 1
    elif name == "hetero":
        mu, var, f = rnd(2, n), (0.5 + uni(2, n)) ** 2, rnd(2, n)
        y = f[0] + rnd(n) / torch.sqrt(5.0 * torch.sigmoid(f[1]))
    else:
        mu, var, f = rnd(n), (0.5 + uni(n)) ** 2, rnd(n)
        sig = torch.sigmoid(f)
        if name == "bernoulli":
            y = (uni(n) < sig).to(torch.uint8)
        elif name == "negbin":      # y ~ NB(r = 10, p = 1 - sigma(f)) as a Gamma-Poisson mixture
            lam = torch.distributions.Gamma(torch.full((1,), 10.0, dtype=torch.float64, device=dev),
                                            torch.ones(1, dtype=torch.float64, device=dev)).sample((n,)).view(n)
            y = torch.poisson(lam * sig / (1 - sig).clamp_min(1e-3)).clamp_max(400).to(torch.int64)
        elif name == "poisson":
            y = torch.poisson(10.0 * sig).to(torch.int64)
        elif name == "laplace":
            u = uni(n) - 0.5
            y = f - torch.sign(u) * torch.log1p(-2 * u.abs())
        else:
            y = f + 1.5 * rnd(n) / torch.sqrt(torch.distributions.Chi2(torch.tensor([3.0], dtype=torch.float64, device=dev)).sample((n,)).view(n) / 3.0)
    return y, mu, var, f


class Leg:
    """one likelihood at n observations on this rank: device-resident state, a step = fused CAVI update (+ELBO) + aux_sample!"""

    def __init__(self, A, ctx, torch, name, n, dev, rank, i0, want_elbo=True):
        self.A, self.ctx, self.torch, self.name, self.n, self.i0 = A, ctx, torch, name, n, i0
        self.lik = make_lik(A, name)
        nl = self.lik.nlatent
        self.y, self.mu, self.var, self.f = make_inputs(torch, name, self.lik, n, dev, 1000 + 17 * rank + sum(map(ord, name)))
        self.qf = A.Normals(self.mu, self.var)
        self.q = A.init_aux_posterior(self.lik, n)
        self.beta = torch.empty((nl, n), dtype=torch.float64, device=dev)
        self.gamma = torch.empty((nl, n), dtype=torch.float64, device=dev)
        self.scal = torch.zeros(8, dtype=torch.float64, device=dev)
        self.Ω = A.init_aux_variables(A.AugPhilox(7, 0), self.lik, n, i0=i0)
        self.units = n * nl if name == "categorical" else n

    def cavi(self):
        self.A.cavi_step_(self.q, self.lik, self.y, self.qf, want_elbo=True, out=(self.beta, self.gamma, self.scal))

    def gibbs(self):
        self.A.aux_sample_(self.Ω, self.lik, self.y, self.f, i0=self.i0)

    def free(self):
        for k in ("y", "mu", "var", "f", "qf", "q", "beta", "gamma", "Ω"):
            setattr(self, k, None)
        self.torch.cuda.empty_cache()


def time_leg(leg, torch, dist, world, steps, warmup, collective=None):
    """W warm-up steps, then EXACTLY `steps` timed steps between barriers; CUDA events on the ctx stream around each half;
    max over ranks.  Returns ms_total / ms_cavi / ms_gibbs / ms_coll per step and the launch count."""
    ctx, st = leg.ctx, leg.ctx.stream

    def step(ev=None):
        if ev:
            ev[0].record(st)
        leg.cavi()
        if ev:
            ev[1].record(st)
        leg.gibbs()
        if ev:
            ev[2].record(st)
        if collective is not None:
            collective(leg.scal)
        if ev:
            ev[3].record(st)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    l0 = ctx.launches()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(st)
    for k in range(steps):
        step(evs[k])
    t1.record(st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = ctx.launches() - l0
    vals = [t0.elapsed_time(t1) / steps, sum(e[0].elapsed_time(e[1]) for e in evs) / steps,
            sum(e[1].elapsed_time(e[2]) for e in evs) / steps, sum(e[2].elapsed_time(e[3]) for e in evs) / steps]
    t = torch.tensor(vals, dtype=torch.float64, device=leg.mu.device)
    per_rank = None
    if world > 1:
        # every rank's own times next to the max: the spread between the GPUs of the box (HBM-bound kernels differ by a
        # percent or two from GPU to GPU) is what the max-over-ranks weak-scaling efficiency mostly measures
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = {"ms_per_step": [round(float(a[0]), 5) for a in allr], "ms_cavi": [round(float(a[1]), 5) for a in allr],
                    "ms_gibbs": [round(float(a[2]), 5) for a in allr]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_cavi, ms_gibbs, ms_coll = t.tolist()
    return dict(ms_per_step=ms_total, ms_cavi=ms_cavi, ms_gibbs=ms_gibbs, ms_allreduce=ms_coll, launches=int(launches),
                per_rank=per_rank)


def leg_report(leg, tm, world, peak, peak_src, steps, traffic=None):
    w = WORKLOADS[leg.name]
    units = leg.units
    ach = w["bytes_cavi"] * units / (tm["ms_cavi"] * 1e-3) / 1e9
    gach = w["bytes_gibbs"] * units / (tm["ms_gibbs"] * 1e-3) / 1e9
    unit = "(obs, class) elements" if leg.name == "categorical" else "obs"
    return {
        "baseline_config": w["cfg"], "likelihood": type(leg.lik).__name__, "obs_per_gpu": leg.n, "units_per_gpu": units,
        "unit_of_rates": unit, "ms_per_step": tm["ms_per_step"], "obs_per_s": leg.n * world / (tm["ms_per_step"] * 1e-3),
        "cavi": {"ms": tm["ms_cavi"], "units_per_s": units * world / (tm["ms_cavi"] * 1e-3),
                 "roofline": {"kernel": CAVI_KERNEL[leg.name], "bound": "hbm", "achieved": ach, "peak": peak,
                              "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                              "bytes_per_unit": w["bytes_cavi"], "traffic": traffic}},
        "gibbs": {"ms": tm["ms_gibbs"], "draws_per_s": units * world / (tm["ms_gibbs"] * 1e-3),
                  "kernel": GIBBS_KERNEL[leg.name], "bound": "fp64 pipe / issue (not HBM)",
                  "hbm_GBs": gach, "frac_hbm": gach / peak, "bytes_per_unit": w["bytes_gibbs"]},
        "gpu_launches_per_step": tm["launches"] // steps,
    }


def oracle_check_and_cpu(leg, torch, m, threads, seed=5):
    """(1) parity at bench scale: the first m observations of the DEVICE inputs are copied to the host and the CPU oracle is
    run on exactly those bits; state / beta / gamma of the timed run and the scalars of a GPU call on the same slice must
    agree to 1e-12.  (2) the same oracle calls, timed, are the cpu_baseline of this workload (`kind: port`)."""
    import numpy as np
    from oracle import orc
    A, name = leg.A, leg.name
    cat, het = name == "categorical", name == "hetero"
    m = min(m, leg.n)
    sl = (slice(None), slice(0, m)) if het else slice(0, m)
    y_d = leg.y[:m].contiguous()
    mu_d, var_d, f_d = leg.mu[sl].contiguous(), leg.var[sl].contiguous(), leg.f[sl].contiguous()
    y, mu, var, f = (t.cpu().numpy() for t in (y_d, mu_d, var_d, f_d))
    olik = make_orc_lik(orc, name)
    orc.set_threads(threads)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        rc, st, ob, og, seq, comp = orc.cavi_step(olik, y, mu, var)
        t1 = time.perf_counter()
        assert rc == 0
        best = (t1 - t0) if best is None else min(best, t1 - t0)
    t_cavi = best
    mg = m if name in ("bernoulli", "laplace", "studentt") else max(1, m // 8)      # the reference sums b Devroye draws: slower
    fg = f[:, :mg] if het else f[:mg]
    t0 = time.perf_counter()
    orc.aux_sample(olik, seed, np.ascontiguousarray(y[:mg]), np.ascontiguousarray(fg))
    t_gibbs = (time.perf_counter() - t0) * (m / mg)
    # --- parity of the timed run's outputs on the slice
    rel = lambda a, b, floor=1e-300: float(np.max(np.where(a == b, 0.0, np.abs(a - b) / np.maximum(np.abs(b), floor))))
    state0 = leg.q._s(0)[:m].cpu().numpy()
    errs = {"state": rel(state0, st[0]),
            "beta": rel(leg.beta[:, :m].cpu().numpy(), ob, 1.0), "gamma": rel(leg.gamma[:, :m].cpu().numpy(), og)}
    # a rank-local call (only rank 0 is here): in fused multi-GPU mode the scalar-producing verbs are COLLECTIVE, so the
    # in-kernel exchange is switched off around it
    fused = leg.ctx.fused
    if fused:
        A.dist.set_fused(leg.ctx, False)
    qs = A.init_aux_posterior(leg.lik, m)
    _, _, _, sc = A.cavi_step_(qs, leg.lik, y_d, A.Normals(mu_d, var_d), want_elbo=True)
    sc = sc.cpu().numpy()
    if fused:
        A.dist.set_fused(leg.ctx, True)
    errs["scalars"] = max(abs(sc[k] - comp[k]) / abs(comp[k]) for k in range(3))
    ok = all(v is None or v <= 1e-12 for v in errs.values())
    assert ok, (name, errs)
    units = m * leg.lik.nlatent if cat else m
    return {"value": m / (t_cavi + t_gibbs), "unit": "obs/s", "cores": threads, "kind": "port",
            "sample": f"first {m} observations of the GPU leg's own inputs (copied D2H): separate-pass CAVI + ELBO sums on all of "
                      f"them, aux_sample! on {mg} (extrapolated x{m // mg}); OpenMP {threads} threads; C++ restatement of the Julia "
                      "reference",
            "cavi_units_per_s": units / t_cavi, "draws_per_s": units / t_gibbs,
            "parity_on_this_sample": {"max_rel_err": errs, "tolerance": 1e-12, "ok": ok}}


def main():
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner under
    # NCCL_DEBUG, torchrun notices) is sent to stderr; the result line is written to the saved descriptor.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--obs-per-gpu", "--n", dest="n", type=int, default=100_000_000,
                    help="observations per GPU (weak scaling); under torchrun use --obs-per-gpu (its parser trips over --n)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-n", type=int, default=20_000_000, help="CPU sample size per step of the reference arm")
    ap.add_argument("--cpu-n", type=int, default=20_000_000, help="CPU sample of the cpu_baseline leg")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: all-reduce of the scalar block fused into the CAVI kernel over peer memory (p2p) or "
                         "a separate ncclAllReduce (nccl)")
    ap.add_argument("--no-defer", action="store_true",
                    help="p2p: exchange inside the CAVI kernel's finaliser (it then waits for the slowest rank every step) "
                         "instead of the split-phase publish / deferred gather")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sparse", action="store_true", help="skip the secondary sparse-GP sweep measurement")
    ap.add_argument("--sparse-m", type=int, default=128)
    ap.add_argument("--sparse-n", type=int, default=2_000_000, help="observations per GPU of the sparse sweep leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--configs", default="negbin,poisson,studentt,laplace,hetero,categorical",
                    help="BASELINE configs[2]-[4] legs to run besides the headline (comma-separated, or 'none')")
    ap.add_argument("--config-steps", type=int, default=10, help="timed steps of each configs[2]-[4] leg")
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
    # before anything is pinned: keep this rank's host buffers and staging copies on the GPU's own NUMA node
    orig_affinity = os.sched_getaffinity(0)
    numa = A.dist.bind_host_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    ctx = A.Context(local_rank)
    A.set_default_context(ctx)
    if world > 1:
        A.dist.init_comm(ctx)                                    # NCCL communicator: the `nccl` collective and the cross-check
        if args.collective == "p2p":
            A.dist.init_p2p(ctx, fused=True)
            if not args.no_defer:                                # split-phase: publish in the CAVI kernel's finaliser, gather
                A.dist.set_deferred(ctx, True)                   # in an extra CTA of the next sampling launch
    dev = torch.device("cuda", local_rank)
    peak, peak_src = load_peaks()
    nccl_coll = (lambda scal: A.dist.allreduce_scalars_(ctx, scal)) if (world > 1 and args.collective == "nccl") else None
    checks = {}

    def fused_vs_nccl(leg):
        """N > 1: the scalars the fused mailbox exchange returned == the NCCL sum of the ranks' local scalars"""
        if world == 1 or args.collective != "p2p":
            return None
        leg.cavi()
        ctx.flush()
        got = leg.scal.clone()
        A.dist.set_fused(ctx, False)
        leg.cavi()
        loc = leg.scal.clone()
        A.dist.allreduce_scalars_(ctx, loc)
        A.dist.set_fused(ctx, True)
        torch.cuda.synchronize()
        err = float(((got[:3] - loc[:3]).abs() / loc[:3].abs()).max().item())
        assert err <= 1e-13, (leg.name, got.tolist(), loc.tolist())
        return err

    # ------------------------------------------------------------------ headline: Bernoulli, weak scaling (n per GPU)
    n = args.n
    head = Leg(A, ctx, torch, "bernoulli", n, dev, rank, rank * n)
    ctx.seed(2026, 0)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    time.sleep(0.25)
    tm = time_leg(head, torch, dist, world, args.steps, args.warmup, nccl_coll)
    clocks = sampler.finish() if rank == 0 else None
    ctx.flush()
    elbo = float(head.scal[2].item())
    checks["fused_mailbox_equals_nccl_sum_of_locals"] = {"bernoulli": fused_vs_nccl(head)}

    # ---------------- e2e: host buffers through the plugin calls, copies inside the timed region
    e2e = e2e_full = None
    if not args.no_e2e:
        lik = head.lik
        pin = lambda t: t.to("cpu").pin_memory()
        hy, hmu, hvar, hf = pin(head.y), pin(head.mu), pin(head.var), pin(head.f)
        hq = A.init_aux_posterior(lik, n, host=True)
        hb = torch.empty((1, n), dtype=torch.float64).pin_memory()
        hg = torch.empty((1, n), dtype=torch.float64).pin_memory()
        hΩ = A.AuxSamples(torch.empty(n, dtype=torch.float64).pin_memory(), None)
        res = {}

        def e2e_step(full):
            # the public API with HOST arrays: api.cavi_step_ / api.aux_sample_ -> aug_cavi_step_host / aug_aux_sample_host
            if full:
                res["c"] = A.cavi_step_(hq, lik, hy, A.Normals(hmu, hvar), want_elbo=True, out=(hb, hg))
            else:   # optional outputs: beta = sign(y - 1/2)/2 is a function of y the caller holds, c only feeds the ELBO
                res["c"] = A.cavi_step_(None, lik, hy, A.Normals(hmu, hvar), want_elbo=True, out=(None, hg), want_beta=False)
            ctx.seed(2026, 99)
            A.aux_sample_(hΩ, lik, hy, hf, i0=rank * n)

        def run_e2e(full):
            torch.cuda.synchronize()
            e2e_step(full)                                 # warm-up (allocates the staging slots)
            if world > 1:
                dist.barrier()
            ksteps = max(1, min(args.steps, 5))
            w0 = time.perf_counter()
            for _ in range(ksteps):
                e2e_step(full)                             # returns after results are on the host
            w1 = time.perf_counter()
            te = torch.tensor([w1 - w0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return te.item() / ksteps, ksteps

        s_full, k_full = run_e2e(True)
        # parity of the e2e path at bench scale: every array it returned == the device-resident run on the same inputs
        ctx.seed(2026, 99)
        head.gibbs()
        torch.cuda.synchronize()
        eq = {"c": bool(torch.equal(hq.c, head.q.c.cpu())), "beta": bool(torch.equal(hb, head.beta.cpu())),
              "gamma": bool(torch.equal(hg, head.gamma.cpu())), "omega": bool(torch.equal(hΩ.omega, head.Ω.omega.cpu()))}
        hs = res["c"][3]
        head.cavi()
        ctx.flush()
        loc_elbo = head.scal.clone()
        if world > 1 and ctx.fused:                        # host verbs are rank-local: compare with this rank's local sums
            A.dist.set_fused(ctx, False)
            head.cavi()
            loc_elbo = head.scal.clone()
            A.dist.set_fused(ctx, True)
        torch.cuda.synchronize()
        eq["scalars_rel_err"] = float(((hs[:3] - loc_elbo[:3].cpu()).abs() / loc_elbo[:3].cpu().abs()).max().item())
        assert all(eq[k] for k in ("c", "beta", "gamma", "omega")) and eq["scalars_rel_err"] <= 1e-12, eq
        checks["e2e_outputs_equal_device_run"] = eq
        s_min, k_min = run_e2e(False)
        assert bool(torch.equal(hg, head.gamma.cpu()))
        h2d = n * (1 + 8 + 8) + n * 8                       # y, mu, var | f  (the Bernoulli aux_sample! reads no y)
        d2h_full, d2h_min = n * (24 + 8) + 64, n * (8 + 8) + 64
        pcie = ("measured on this pool (tools/pcie_probe.py, profiles/e2e_pcie_probe_r1k.txt): 55.5 GB/s H2D, 57.0 D2H, "
                "49.8 each way when both run")
        e2e = {"value": n * world / s_min, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_min,
               "ms_per_step": 1e3 * s_min, "steps": k_min,
               "path": "api.cavi_step_ + api.aux_sample_ on pinned HOST arrays -> aug_cavi_step_host + aug_aux_sample_host "
                       "(3-slot H2D / kernel / D2H pipeline inside the library)",
               "outputs": "gamma, omega, ELBO scalars (optional outputs skipped: beta = sign(y-1/2)/2 is a function of y the "
                          "caller holds — bernoulli.jl:28 — and the state c only feeds the ELBO terms the call already returns)",
               "pcie_GBs": (h2d + d2h_min) / s_min * 1e-9, "pcie_bound": pcie}
        e2e_full = {"value": n * world / s_full, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_full,
                    "ms_per_step": 1e3 * s_full, "steps": k_full, "outputs": "c, beta, gamma, omega, ELBO scalars (everything)",
                    "pcie_GBs": (h2d + d2h_full) / s_full * 1e-9}
        # the platform's bound for this step: the same bytes as plain concurrent H2D + D2H copies (no kernels), all ranks at once
        def copy_only():
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            d_y, d_mu, d_var, d_f = (torch.empty_like(t, device=dev) for t in (hy, hmu, hvar, hf))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            reps = 3
            w0 = time.perf_counter()
            for _ in range(reps):
                with torch.cuda.stream(s_in):
                    for d_, h_ in ((d_y, hy), (d_mu, hmu), (d_var, hvar), (d_f, hf)):
                        d_.copy_(h_, non_blocking=True)
                with torch.cuda.stream(s_out):
                    hg.copy_(head.gamma, non_blocking=True)
                    hΩ.omega.copy_(head.Ω.omega, non_blocking=True)
                torch.cuda.synchronize()
            w1 = time.perf_counter()
            tc = torch.tensor([(w1 - w0) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            return tc.item()
        s_copy = copy_only()
        e2e["copy_only"] = {"ms_per_step": 1e3 * s_copy, "GBs_per_gpu": (h2d + d2h_min) / s_copy * 1e-9,
                            "what": "the step's H2D and D2H bytes as plain concurrent cudaMemcpyAsync on two streams, no kernels, "
                                    "all ranks at once (max over ranks): the host / PCIe bound of this box for the e2e step",
                            "e2e_over_copy_only": s_min / s_copy}
        if world > 1:
            bound = [None] * world
            dist.all_gather_object(bound, numa)
            e2e["numa_binding"] = {"what": "each rank pinned to the CPUs of its GPU's NUMA node before allocating pinned host "
                                           "buffers (dist.bind_host_to_gpu_numa_node); null = the platform exposes a single NUMA "
                                           "node (e.g. a one-node VM), nothing to bind", "per_rank": bound}
        del hy, hmu, hvar, hf, hq, hb, hg, hΩ

    # ---------------- cpu_baseline + parity on a slice of the device inputs (rank 0)
    cpu = None
    threads = 1
    if rank == 0 and not args.no_cpu:
        os.sched_setaffinity(0, orig_affinity)     # the CPU arm uses every host core, not just this rank's NUMA node
        from oracle import orc
        orc.lib()
        threads = host_threads(orc)
        cpu = oracle_check_and_cpu(head, torch, args.cpu_n, threads)
        n1 = min(args.cpu_n // 10, 2_000_000)
        t1, c1, g1 = cpu_reference(n1, 1)
        cpu["single_thread"] = {"value": n1 / t1, "cavi_obs_per_s": n1 / c1, "pg_draws_per_s": n1 / g1,
                                "sample": f"{n1} obs, 1 thread (the reference is single-threaded)"}
    if world > 1:
        dist.barrier()
    head_report = leg_report(head, tm, world, peak, peak_src, args.steps)
    head.free()

    # ---------------- strong scaling: the SAME global problem (N = 1e8 in total) sharded over the ranks
    strong = None
    if world > 1:
        lo, hi = A.dist.shard_bounds(args.n, world, rank)
        sleg = Leg(A, ctx, torch, "bernoulli", hi - lo, dev, rank, lo)
        stm = time_leg(sleg, torch, dist, world, args.steps, args.warmup, nccl_coll)
        strong = {"global_obs": args.n, "obs_per_gpu": hi - lo, "ms_per_step": stm["ms_per_step"],
                  "value": args.n / (stm["ms_per_step"] * 1e-3), "unit": UNIT, "ms_cavi": stm["ms_cavi"],
                  "ms_gibbs": stm["ms_gibbs"], "ms_allreduce": stm["ms_allreduce"],
                  "speedup_vs_1gpu_weak_step": tm["ms_per_step"] / stm["ms_per_step"],
                  "efficiency_vs_ideal": (tm["ms_per_step"] / n) * (hi - lo) / stm["ms_per_step"],
                  "note": "ideal = this run's per-observation time of the weak leg times the shard size; SURVEY §8(e) "
                          "expects ~85-90% at 8 GPUs for the scalar-returning variant (launch + exchange latency against a "
                          "~0.35 ms step)"}
        sleg.free()

    # ---------------- BASELINE configs[2]-[4]
    cfgs = {}
    names = [] if args.configs in ("", "none") else [c for c in args.configs.split(",") if c]
    for name in names:
        w = WORKLOADS[name]
        nn = w["n"] if args.n >= 100_000_000 else max(1, int(w["n"] * args.n / 100_000_000))
        leg = Leg(A, ctx, torch, name, nn, dev, rank, rank * nn)
        ltm = time_leg(leg, torch, dist, world, args.config_steps, 3, nccl_coll)
        rep = leg_report(leg, ltm, world, peak, peak_src, args.config_steps)
        rep["scaling"] = "weak"
        fv = fused_vs_nccl(leg)
        if fv is not None:
            rep["fused_mailbox_vs_nccl_sum_rel_err"] = fv
        if rank == 0 and not args.no_cpu:
            m = {"categorical": 20_000, "hetero": 4_000_000}.get(name, 8_000_000)
            rep["cpu_baseline"] = oracle_check_and_cpu(leg, torch, m, threads)
            rep["speedup_vs_cpu"] = {"cavi": rep["cavi"]["units_per_s"] / rep["cpu_baseline"]["cavi_units_per_s"],
                                     "gibbs": rep["gibbs"]["draws_per_s"] / rep["cpu_baseline"]["draws_per_s"]}
        if world > 1:
            dist.barrier()
        leg.free()
        if world > 1:                                      # strong: the config's global size sharded over the ranks
            lo, hi = A.dist.shard_bounds(w["n"], world, rank)
            sl = Leg(A, ctx, torch, name, hi - lo, dev, rank, lo)
            stm = time_leg(sl, torch, dist, world, args.config_steps, 3, nccl_coll)
            rep["strong"] = {"global_obs": w["n"], "obs_per_gpu": hi - lo, "ms_per_step": stm["ms_per_step"],
                             "ms_cavi": stm["ms_cavi"], "ms_gibbs": stm["ms_gibbs"],
                             "obs_per_s": w["n"] / (stm["ms_per_step"] * 1e-3),
                             "efficiency_vs_ideal": (ltm["ms_per_step"] / nn) * (hi - lo) / stm["ms_per_step"]}
            sl.free()
        cfgs[name] = rep

    # ---------------- secondary leg: one sparse-GP CAVI iteration as one pass over κ
    sparse = None
    if not args.no_sparse:
        sparse = sparse_leg(A, ctx, torch, dist, world, dev, args, rank)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("cavi_bernoulli_bytes_per_launch_at_1e8")
            if traffic is not None and n != 100_000_000:
                traffic = traffic * n / 100_000_000
        except Exception:
            traffic = None
    ms_step, ms_cavi, ms_gibbs = tm["ms_per_step"], tm["ms_cavi"], tm["ms_gibbs"]
    try:
        sm_count = int(ctx.sm_count())
    except Exception:
        sm_count = 0
    ach = BYTES_CAVI * n / (ms_cavi * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": n * world / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] (Bernoulli-logistic Gibbs aux_sample! PG draws, N=1e8 fp64) + "
                               "configs[0]'s fused CAVI update at the same N; configs[2]-[4] in `configs`",
                   "obs_per_gpu": n, "global_obs": n * world, "likelihood": "BernoulliLikelihood(LogisticLink)",
                   "l2": "inputs+outputs 5.7 GB per step >> 126 MB L2 (no flush needed)",
                   "sharding": f"contiguous observation blocks, {world} rank(s); only the 8-double scalar block "
                               "is all-reduced",
                   "collective": ("none (1 rank)" if world == 1 else
                                  ("split-phase over the peer-memory mailbox (NVLink): published by cavi_tma_kernel's finaliser, "
                                   "gathered by an extra CTA of the sampling kernel" if not args.no_defer else
                                   "fused into cavi_tma_kernel's finaliser over the peer-memory mailbox (NVLink)")
                                  if args.collective == "p2p" else "ncclAllReduce(sum, double, 8) after the step")},
        "parts": {"cavi_obs_per_s": n * world / (ms_cavi * 1e-3), "pg_draws_per_s": n * world / (ms_gibbs * 1e-3),
                  "ms_cavi": ms_cavi, "ms_gibbs": ms_gibbs, "ms_allreduce": tm["ms_allreduce"],
                  "per_rank": tm.get("per_rank")},
        "roofline": {"kernel": "cavi_tma_kernel<BERNOULLI, ELBO> (aux_posterior! + expected potential/precision + ELBO sums): "
                               "the HBM-bound kernel of the step; the time-dominant one is the issue-bound sampler in roofline_gibbs",
                     "bound": "hbm", "achieved": ach,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                     "frac_of_8TBs_nominal": ach / 8000.0, "bytes_per_obs": BYTES_CAVI, "traffic": traffic,
                     "share_of_step": ms_cavi / ms_step},
        "roofline_gibbs": {"kernel": GIBBS_KERNEL["bernoulli"], "bound": "fp64 pipe / issue (not HBM)",
                           "achieved": BYTES_GIBBS * n / (ms_gibbs * 1e-3) / 1e9, "unit": "GB/s",
                           "frac_hbm": BYTES_GIBBS * n / (ms_gibbs * 1e-3) / 1e9 / peak,
                           "pg_draws_per_s_per_gpu": n / (ms_gibbs * 1e-3), "share_of_step": ms_gibbs / ms_step,
                           "issue": issue_roofline(n, ms_gibbs, sm_count, clocks),
                           "ncu": "profiles/r2zz_ncu_summary.txt (issue-slot utilisation, active lanes, pipe shares of this kernel)"},
        "cpu_baseline": cpu, "e2e": e2e, "e2e_all_outputs": e2e_full, "gpu_launches": int(tm["launches"]), "clocks": clocks,
        "elbo_check": elbo, "checks": checks, "strong_scaling": strong, "configs": cfgs,
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
    try:
        main()
    except BaseException as exc:                # a failed self-check must end the WHOLE job at once: a rank that lingers in
        if isinstance(exc, SystemExit) and exc.code in (0, None):       # interpreter shutdown keeps its peers in a barrier
            raise
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
