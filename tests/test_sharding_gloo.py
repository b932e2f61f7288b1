"""Multi-rank host logic on CPU (gloo, world_size 2): the observation axis is cut into contiguous blocks,
each rank computes its partial scalar block, and the all-reduced sum equals the single-rank result.
The per-shard arithmetic is done by the ORACLE here (no GPU in this container); the sharding /
combination logic under test (dist.shard_bounds, the additive scalar block) is the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import aug_pkg
from common import BERNOULLI, CAT_BIJ, POISSON, synth_inputs


def test_shard_bounds_partition():
    D = aug_pkg.load_package().dist
    for n in (0, 1, 7, 100, 10**8 + 3):
        for world in (1, 2, 3, 8):
            prev = 0
            sizes = []
            for r in range(world):
                lo, hi = D.shard_bounds(n, world, r)
                assert lo == prev and hi >= lo
                prev = hi
                sizes.append(hi - lo)
            assert prev == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    orc.set_threads(1)
    D = aug_pkg.load_package().dist
    res = {}
    for name, kind, params, kw in [("bern", BERNOULLI, (), {}), ("pois", POISSON, (10.0,), {}),
                                   ("cat", CAT_BIJ, (), dict(nlatent=4))]:
        n = 1001
        y, mu, var, f = synth_inputs(kind, n, 42, params, kw.get("nlatent", 1))
        lo, hi = D.shard_bounds(n, world, rank)
        lik = orc.make_lik(kind, *params, **kw)
        rc, st, b, g, seq, comp = orc.cavi_step(lik, np.ascontiguousarray(y[lo:hi]),
                                                np.ascontiguousarray(mu[lo:hi]), np.ascontiguousarray(var[lo:hi]))
        assert rc == 0
        block = torch.from_numpy(comp.copy())
        dist.all_reduce(block, op=dist.ReduceOp.SUM)          # the one collective of the path
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, g[:, :].tolist()))
        res[name] = (block.numpy().tolist(), gathered)
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scalar_allreduce_matches_single_rank():
    from oracle import orc
    orc.set_threads(1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    D = aug_pkg.load_package().dist
    for name, kind, params, kw in [("bern", BERNOULLI, (), {}), ("pois", POISSON, (10.0,), {}),
                                   ("cat", CAT_BIJ, (), dict(nlatent=4))]:
        y, mu, var, f = synth_inputs(kind, 1001, 42, params, kw.get("nlatent", 1))
        lik = orc.make_lik(kind, *params, **kw)
        rc, st, b, g, seq, comp = orc.cavi_step(lik, y, mu, var)
        block, gathered = res[name]
        for k in range(3):
            assert block[k] == pytest.approx(comp[k], rel=1e-12)
        # sharded β/γ outputs concatenate to the unsharded ones (state stays sharded)
        gg = np.concatenate([np.array(part[2]) for part in sorted(gathered)], axis=1)
        assert np.array_equal(gg, g)
        assert D.combine_scalars_host([block])[0] == block[0]


def _sparse_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import synth_sparse
    from oracle import orc
    orc.set_threads(1)
    D = aug_pkg.load_package().dist
    n, m = 1501, 12
    kappa, mvec, B, kdiag = synth_sparse(n, m, 5)
    y, _, _, _ = synth_inputs(BERNOULLI, n, 6)
    rng = np.random.default_rng(1)
    P0 = rng.standard_normal((m, m)); P0 = P0 @ P0.T
    r0 = rng.standard_normal(m)
    lo, hi = D.shard_bounds(n, world, rank)
    # P0 / r0 enter on rank 0 only (include/augcuda.h, "Multi-GPU"); every other rank contributes its shard's sums
    rc, o = orc.sparse_cavi_sweep(orc.make_lik(orc.BERNOULLI), np.ascontiguousarray(y[lo:hi]),
                                  np.ascontiguousarray(kappa[lo:hi]), mvec, B, np.ascontiguousarray(kdiag[lo:hi]),
                                  P0 if rank == 0 else None, r0 if rank == 0 else None)
    assert rc == 0
    Pr = torch.from_numpy(np.concatenate([o["P"].ravel(), o["rhs"]]))
    dist.all_reduce(Pr, op=dist.ReduceOp.SUM)              # the exchange of the sparse rows: m*m + m doubles
    sc = torch.from_numpy(o["comp"].copy())
    dist.all_reduce(sc, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((Pr.numpy().tolist(), sc.numpy().tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sparse_sweep_matches_single_rank():
    """SURVEY §8(f) rows 1-2 under sharding: P, rhs and the ELBO sums are plain sums over observation shards."""
    from common import synth_sparse
    from oracle import orc
    orc.set_threads(1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sparse_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    Pr, sc = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, m = 1501, 12
    kappa, mvec, B, kdiag = synth_sparse(n, m, 5)
    y, _, _, _ = synth_inputs(BERNOULLI, n, 6)
    rng = np.random.default_rng(1)
    P0 = rng.standard_normal((m, m)); P0 = P0 @ P0.T
    r0 = rng.standard_normal(m)
    rc, o = orc.sparse_cavi_sweep(orc.make_lik(orc.BERNOULLI), y, kappa, mvec, B, kdiag, P0, r0)
    np.testing.assert_allclose(np.array(Pr[: m * m]).reshape(m, m), o["P"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(np.array(Pr[m * m:]), o["rhs"], rtol=0, atol=1e-12)
    for k in range(3):
        assert sc[k] == pytest.approx(o["comp"][k], rel=1e-12)
