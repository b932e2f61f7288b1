"""Multi-GPU parity (needs >= 2 GPUs of one box; skipped otherwise — run with `gpurun --gpus 2`).

The observation axis is cut into contiguous shards (dist.shard_bounds); the only exchange of the path is the
sum of the 64-byte scalar block.  Two implementations are checked against the single-GPU result and
the oracle: the NCCL all-reduce issued by the library, and the peer-memory mailbox that performs the all-reduce
INSIDE the reducing kernel (include/augcuda.h, "Peer-memory mailbox")."""
import os
import socket

import numpy as np
import pytest
import torch

import aug_pkg
from common import BERNOULLI, CAT_BIJ, HETERO, NEGBIN, POISSON, STUDENTT, synth_inputs
from gpu_common import make_lik

pytestmark = pytest.mark.gpu


def _need2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


CASES = [("bern", BERNOULLI, (), {}, 40_001), ("negbin", NEGBIN, (10,), dict(r_is_int=True), 20_000),
         ("pois", POISSON, (10.0,), {}, 33_333), ("studentt", STUDENTT, (3.0, 1.5), {}, 25_000),
         ("hetero", HETERO, (5.0,), dict(nlatent=2), 30_000), ("cat", CAT_BIJ, (), dict(nlatent=9), 7_001)]


def _shard(kind, arr, lo, hi):
    if kind == HETERO and arr.ndim == 2:
        return np.ascontiguousarray(arr[:, lo:hi])
    return np.ascontiguousarray(arr[lo:hi])


def _run_rank(A, ctx, rank, world, fused):
    """every scalar-producing verb on this rank's shard; returns {case: scalars (8,)}"""
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.tdev)
    out = {}
    for name, kind, params, kw, n in CASES:
        y, mu, var, f = synth_inputs(kind, n, 7, params, kw.get("nlatent", 1))
        lo, hi = A.dist.shard_bounds(n, world, rank)
        lik = make_lik(kind, params, kw)
        ys, ms, vs, fs = (_shard(kind, a, lo, hi) for a in (y, mu, var, f))
        q = A.init_aux_posterior(lik, hi - lo, ctx=ctx)
        q, beta, gamma, scal = A.cavi_step_(q, lik, dev(ys), A.Normals(dev(ms), dev(vs)), want_elbo=True, ctx=ctx)
        if not fused:
            A.dist.allreduce_scalars_(ctx, scal)
        Ω = A.aux_sample(A.AugPhilox(11, 0), lik, dev(ys), dev(fs), ctx=ctx, i0=lo)
        lt = A.logtilt(lik, Ω, dev(ys), dev(fs), ctx=ctx)          # reduced over ranks by the verb itself
        el = A.expected_logtilt(lik, q, dev(ys), A.Normals(dev(ms), dev(vs)), ctx=ctx)
        ctx.sync()
        out[name] = (scal.cpu().numpy().copy(), lt, el, Ω.omega.cpu().numpy().copy(), lo, hi)
    # an empty shard still takes part in the exchange
    lik = make_lik(BERNOULLI, (), {})
    n = 5000
    y, mu, var, f = synth_inputs(BERNOULLI, n, 3, (), 1)
    lo, hi = (0, n) if rank == 0 else (n, n)
    q = A.init_aux_posterior(lik, hi - lo, ctx=ctx)
    q, beta, gamma, scal = A.cavi_step_(q, lik, dev(y[lo:hi]), A.Normals(dev(mu[lo:hi]), dev(var[lo:hi])), ctx=ctx)
    if not fused:
        A.dist.allreduce_scalars_(ctx, scal)
    ctx.sync()
    out["empty_shard"] = (scal.cpu().numpy().copy(), 0.0, 0.0, np.zeros(0), lo, hi)
    return out


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    A = aug_pkg.load_package()
    ctx = A.Context(rank)
    A.set_default_context(ctx)
    if mode == "nccl":
        A.dist.init_comm(ctx)
    else:
        A.dist.init_p2p(ctx, fused=True)
        if mode == "p2p_deferred":      # split-phase: publish in the reducing kernel, gather in the next sampling launch
            A.dist.set_deferred(ctx, True)
    res = _run_rank(A, ctx, rank, world, fused=(mode != "nccl"))
    flag = ctx.error_flag()
    q.put((rank, res, flag))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _single_gpu_reference(A, orc):
    ctx = A.Context(0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.tdev)
    ref = {}
    for name, kind, params, kw, n in CASES:
        y, mu, var, f = synth_inputs(kind, n, 7, params, kw.get("nlatent", 1))
        lik = make_lik(kind, params, kw)
        q = A.init_aux_posterior(lik, n, ctx=ctx)
        q, beta, gamma, scal = A.cavi_step_(q, lik, dev(y), A.Normals(dev(mu), dev(var)), ctx=ctx)
        Ω = A.aux_sample(A.AugPhilox(11, 0), lik, dev(y), dev(f), ctx=ctx)
        lt = A.logtilt(lik, Ω, dev(y), dev(f), ctx=ctx)
        ctx.sync()
        olik = orc.make_lik(kind, *params, **kw)
        rc, st, b, g, seq, comp = orc.cavi_step(olik, y, mu, var)
        assert rc == 0
        ref[name] = (scal.cpu().numpy().copy(), lt, Ω.omega.cpu().numpy().copy(), comp)
    lik = make_lik(BERNOULLI, (), {})
    y, mu, var, f = synth_inputs(BERNOULLI, 5000, 3, (), 1)
    q = A.init_aux_posterior(lik, 5000, ctx=ctx)
    q, beta, gamma, scal = A.cavi_step_(q, lik, dev(y), A.Normals(dev(mu), dev(var)), ctx=ctx)
    ctx.sync()
    ref["empty_shard"] = (scal.cpu().numpy().copy(), 0.0, np.zeros(0), None)
    ctx.close()
    return ref


@pytest.mark.parametrize("mode", ["nccl", "p2p_fused", "p2p_deferred"])
def test_two_rank_scalars_and_draws_match_single_gpu(orc, mode):
    _need2()
    import torch.multiprocessing as mp
    A = aug_pkg.load_package()
    ref = _single_gpu_reference(A, orc)
    mpc = mp.get_context("spawn")
    qu = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, mode, qu)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, res, flag = qu.get(timeout=300)
        assert flag == 0, f"rank {rank}: device error flag {flag}"
        got[rank] = res
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for name in ref:
        s_ref, lt_ref, om_ref, comp = ref[name]
        s0, lt0, el0, om0, lo0, hi0 = got[0][name]
        s1, lt1, el1, om1, lo1, hi1 = got[1][name]
        # every rank holds the SAME sums (bit-identical for the mailbox: fixed rank-order addition)
        assert np.array_equal(s0[:3], s1[:3]), name
        assert lt0 == lt1 and el0 == el1
        for k in range(3):
            assert s0[k] == pytest.approx(s_ref[k], rel=1e-12), (name, k)
            if comp is not None:
                assert s0[k] == pytest.approx(comp[k], rel=1e-12), (name, k)
        if name != "empty_shard":
            assert lt0 == pytest.approx(lt_ref, rel=1e-11, abs=1e-9)
            assert el0 == pytest.approx(s_ref[0], rel=1e-12)
            # draws depend on (seed, offset, GLOBAL index) only: the shards concatenate to the single-GPU sample
            assert np.array_equal(np.concatenate([om0, om1]), om_ref), name


def test_in_process_mailbox_two_contexts():
    """two ctxs of ONE process attached through raw peer pointers (aug_comm_p2p_attach_ptrs)"""
    _need2()
    import ctypes as C
    A = aug_pkg.load_package()
    ctxs = [A.Context(0), A.Context(1)]
    ptrs = (C.c_void_p * 2)()
    for r, c in enumerate(ctxs):
        p = C.c_void_p()
        A.check(c.lib.aug_comm_p2p_export(c.h, None, C.byref(p)))
        ptrs[r] = p.value
    devs = (C.c_int32 * 2)(0, 1)
    for r, c in enumerate(ctxs):
        A.check(c.lib.aug_comm_p2p_attach_ptrs(c.h, 2, r, ptrs, devs))
    vals = []
    for r, c in enumerate(ctxs):
        with torch.cuda.device(r):
            vals.append(torch.arange(8, dtype=torch.float64, device=c.tdev) * (r + 1) + 0.25)
    for rep in range(3):                                   # epochs advance on the device
        for r, c in enumerate(ctxs):
            c.p2p_ready = True
            A.dist.allreduce_scalars_p2p_(c, vals[r], 7)
        for c in ctxs:
            c.sync()
        assert c.error_flag() == 0
        assert torch.equal(vals[0].cpu()[:7], vals[1].cpu()[:7])
    base = np.arange(8) * 1.0 + 0.25, np.arange(8) * 2.0 + 0.25
    exp = base[0][:7] + base[1][:7]
    for _ in range(2):
        exp = exp + exp
    assert np.array_equal(vals[0].cpu().numpy()[:7], exp)
    assert vals[0].cpu().numpy()[7] == base[0][7]           # slot 7 is not exchanged
    for c in ctxs:
        c.close()


def test_mailbox_timeout_raises_flag_instead_of_hanging():
    """a rank whose peer never launches the matching verb gets NaN + error-flag bit 1 after the timeout"""
    _need2()
    import ctypes as C
    os.environ["AUGCUDA_XCH_TIMEOUT_MS"] = "200"
    try:
        A = aug_pkg.load_package()
        ctxs = [A.Context(0), A.Context(1)]
        ptrs = (C.c_void_p * 2)()
        for r, c in enumerate(ctxs):
            p = C.c_void_p()
            A.check(c.lib.aug_comm_p2p_export(c.h, None, C.byref(p)))
            ptrs[r] = p.value
        devs = (C.c_int32 * 2)(0, 1)
        for r, c in enumerate(ctxs):
            A.check(c.lib.aug_comm_p2p_attach_ptrs(c.h, 2, r, ptrs, devs))
        c = ctxs[0]
        c.p2p_ready = True
        v = torch.ones(8, dtype=torch.float64, device=c.tdev)
        A.dist.allreduce_scalars_p2p_(c, v, 3)             # rank 1 never joins
        c.sync()
        assert c.error_flag() & 2
        assert torch.isnan(v[:3]).all()
        for c in ctxs:
            c.close()
    finally:
        del os.environ["AUGCUDA_XCH_TIMEOUT_MS"]


def _sparse_worker(rank, world, port, mode, qu):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from common import synth_sparse
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    A = aug_pkg.load_package()
    ctx = A.Context(rank)
    A.set_default_context(ctx)
    if mode == "nccl":
        A.dist.init_comm(ctx)
    else:
        A.dist.init_p2p(ctx, fused=True)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.tdev)
    n, m = 30_001, 64
    kappa, mvec, B, kdiag = synth_sparse(n, m, 5)
    y, _, _, _ = synth_inputs(BERNOULLI, n, 6)
    P0 = np.eye(m) * 0.5
    lo, hi = A.dist.shard_bounds(n, world, rank)
    lik = A.BernoulliLikelihood()
    res = []
    for rep in range(3):                                   # epochs advance on the device: repeat the collective verb
        P, rhs, scal, _, bg = A.sparse_cavi_sweep_(None, lik, dev(y[lo:hi]), dev(kappa[lo:hi]), dev(mvec), dev(B),
                                                   dev(kdiag[lo:hi]), P0=dev(P0), want_potentials=True)
        P2, rhs2 = A.sparse_precision_potential(dev(kappa[lo:hi]), bg[1], bg[0], P0=dev(P0))   # the stand-alone verb too
        ctx.sync()
        res.append((P.cpu().numpy().copy(), rhs.cpu().numpy().copy(), scal.cpu().numpy().copy(),
                    P2.cpu().numpy().copy(), rhs2.cpu().numpy().copy()))
    for r in res[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(r, res[0]))
    assert ctx.error_flag() == 0
    qu.put((rank,) + res[0])
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "p2p_fused"])
def test_two_rank_sparse_sweep_matches_single_gpu(orc, mode):
    """SURVEY §8(f) rows 1-2 sharded over 2 GPUs: the exchange is one ncclAllReduce of m*m + m doubles (and the
    scalar block); every rank ends with the P, rhs and ELBO sums of the whole data set."""
    _need2()
    import torch.multiprocessing as mp
    from common import synth_sparse
    A = aug_pkg.load_package()
    n, m = 30_001, 64
    kappa, mvec, B, kdiag = synth_sparse(n, m, 5)
    y, _, _, _ = synth_inputs(BERNOULLI, n, 6)
    P0 = np.eye(m) * 0.5
    orc.set_threads(8)
    rc, o = orc.sparse_cavi_sweep(orc.make_lik(orc.BERNOULLI), y, kappa, mvec, B, kdiag, P0, None)
    orc.set_threads(1)
    mpc = mp.get_context("spawn")
    qu = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_sparse_worker, args=(r, 2, port, mode, qu)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, P, rhs, scal, P2, rhs2 = qu.get(timeout=300)
        got[rank] = (P, rhs, scal)
        assert np.all(np.abs(P2 - P) <= 1e-12 * np.abs(P).max()) and np.all(np.abs(rhs2 - rhs) <= 1e-12 * np.abs(rhs).max())
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    d = np.sqrt(np.diag(o["P"]))
    assert np.all(np.abs(got[0][0] - o["P"]) <= 1e-12 * np.outer(d, d))
    assert np.all(np.abs(got[0][1] - o["rhs"]) <= 1e-12 * (np.abs(kappa) * np.abs(o["beta"])[:, None]).sum(0))
    for k in range(3):
        assert got[0][2][k] == pytest.approx(o["comp"][k], rel=1e-12)
