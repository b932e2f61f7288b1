"""GPU parity tests of the HOST-buffer verbs (the `e2e` path of bench.py): every verb of the path called with host
arrays, through the *_host entry points of the C ABI (chunked H2D / kernel / D2H pipeline inside the library).

Bar: arrays bit-equal to the device-pointer verbs on the same inputs (same kernels, other chunking) and within 1e-12
of the CPU oracle; scalars within 1e-12 of the oracle's compensated sums; samples bit-equal to aug_aux_sample /
aug_init_aux_variables for the same (seed, offset, i0).  Sizes straddle the chunk boundaries: 1, chunk-1, chunk,
chunk+1, 3 chunks + an odd tail — with a small chunk (AUGCUDA_HOST_CHUNK_LOG2) for all eight kinds and with the
default 4 Mi-element chunk for three of them.
"""
import os

import numpy as np
import pytest
import torch

from common import (BERNOULLI, CAT, CAT_BIJ, HETERO, LAPLACE, NEGBIN, POISSON, STUDENTT, relerr, synth_inputs)

pytestmark = pytest.mark.gpu
RTOL = 1e-12

CASES = [
    ("bernoulli", BERNOULLI, (), {}),
    ("negbin10", NEGBIN, (10,), dict(r_is_int=True)),
    ("negbin5.5", NEGBIN, (5.5,), {}),
    ("poisson10", POISSON, (10.0,), {}),
    ("laplace1", LAPLACE, (1.0,), {}),
    ("studentt", STUDENTT, (3.0, 1.5), {}),
    ("hetero5", HETERO, (5.0,), {}),
    ("cat_bij_K12", CAT_BIJ, (), dict(nlatent=11, logtheta=[0.1 * (j % 5) - 0.2 for j in range(12)])),
    ("cat_K7", CAT, (), dict(nlatent=7)),
]


@pytest.fixture(scope="module")
def A():
    from gpu_common import pkg
    return pkg()


@pytest.fixture()
def small_chunks():
    old = os.environ.get("AUGCUDA_HOST_CHUNK_LOG2")
    os.environ["AUGCUDA_HOST_CHUNK_LOG2"] = "12"
    yield 4096
    if old is None:
        del os.environ["AUGCUDA_HOST_CHUNK_LOG2"]
    else:
        os.environ["AUGCUDA_HOST_CHUNK_LOG2"] = old


def _rows_per_chunk(elems, nl):
    c = max(2, elems // nl)
    return c & ~1


def _sizes(chunk_rows):
    return [1, chunk_rows - 1, chunk_rows, chunk_rows + 1, 3 * chunk_rows + 1235]


def _eq(a, b):
    if a is None and b is None:
        return True
    return torch.equal(a.cpu(), b.cpu())


def _tup_eq(a, b):
    return all(_eq(x, y) for x, y in zip(a, b))


def _check_all_verbs(A, orc, name, kind, params, kw, n, seed, oracle=True):
    from gpu_common import dev, host, make_lik, stack
    lik = make_lik(kind, params, kw)
    nl = kw.get("nlatent", 1)
    y, mu, var, f = synth_inputs(kind, n, seed, params, nl)
    want_elbo = kind != CAT
    tag = (name, n)
    # ---------------- variational side
    qd = A.init_aux_posterior(lik, n)
    qd, bd, gd, sd = A.cavi_step_(qd, lik, dev(y), A.Normals(dev(mu), dev(var)), want_elbo=want_elbo)
    qh = A.init_aux_posterior(lik, n, host=True)
    for i in range(3):
        if qh._s(i) is not None:
            assert not qh._s(i).is_cuda and float(qh._s(i).double().abs().sum()) == 0.0
    qh, bh, gh, sh = A.cavi_step_(qh, lik, y, A.Normals(mu, var), want_elbo=want_elbo)      # numpy arrays in
    assert all(not t.is_cuda for t in bh) and len(bh) == lik.nlatent and bh[0].shape == (n,)
    for i in range(3):
        assert _eq(qh._s(i), qd._s(i)), tag + ("state", i)
    assert _tup_eq(bh, bd) and _tup_eq(gh, gd), tag
    if want_elbo:
        s_d, s_h = host(sd), sh.numpy()
        for k in range(3):
            assert s_h[k] == pytest.approx(s_d[k], rel=RTOL, abs=1e-12), tag + (k,)
        assert np.all(s_h[3:6] == 0.0)
    if oracle:
        olik = orc.make_lik(kind, *params, **kw)
        rc, ostate, obeta, ogamma, oseq, ocomp = orc.cavi_step(olik, y, mu, var, want_scalars=want_elbo)
        assert rc == 0
        for i in range(3):
            if ostate[i] is not None and qh._s(i) is not None:
                assert relerr(qh._s(i).numpy(), ostate[i]) < RTOL, tag + ("oracle state", i)
        assert relerr(np.stack([t.numpy() for t in bh]), obeta, floor=1.0) < RTOL, tag
        assert relerr(np.stack([t.numpy() for t in gh]), ogamma) < RTOL, tag
        if want_elbo:
            for k in range(3):
                assert sh.numpy()[k] == pytest.approx(ocomp[k], rel=RTOL, abs=1e-12), tag + ("oracle scalar", k)
    # separate verbs on host arrays
    q2 = A.aux_posterior(lik, y, A.Normals(mu, var))
    for i in range(3):
        assert _eq(q2._s(i), qd._s(i)), tag + ("aux_posterior", i)
    b2, g2 = A.expected_auglik_potential_and_precision(lik, q2, y, A.Normals(mu, var))
    b2d, g2d = A.expected_auglik_potential_and_precision(lik, qd, dev(y), A.Normals(dev(mu), dev(var)))
    assert _tup_eq(b2, b2d) and _tup_eq(g2, g2d), tag
    assert A.expected_auglik_precision(lik, q2, y, A.Normals(mu, var))[0].is_cuda is False
    if want_elbo:
        qfd = A.Normals(dev(mu), dev(var))
        for fn in (A.expected_logtilt, A.aux_kldivergence, A.expected_aug_loglik):
            assert fn(lik, q2, y, A.Normals(mu, var)) == pytest.approx(fn(lik, qd, dev(y), qfd), rel=RTOL, abs=1e-12)
    # optional outputs: no state, no beta -> same gamma and scalars
    _, b3, g3, s3 = A.cavi_step_(None, lik, y, A.Normals(mu, var), want_elbo=want_elbo, want_beta=False)
    assert b3 is None and _tup_eq(g3, gd)
    if want_elbo:
        assert np.array_equal(s3.numpy(), sh.numpy())
    # ---------------- sampling side
    i0 = 12345
    Wd = A.aux_sample(A.AugPhilox(7, 3), lik, dev(y), dev(f), i0=i0)
    rng = A.AugPhilox(7, 3)
    Wh = A.aux_sample(rng, lik, y, f, i0=i0)
    assert rng.offset == 4                                        # one tick per call, whatever the chunking
    assert _eq(Wh.omega, Wd.omega) and _eq(Wh.n, Wd.n), tag
    Id = A.init_aux_variables(A.AugPhilox(9, 0), lik, n, i0=i0)
    Ih = A.init_aux_variables(A.AugPhilox(9, 0), lik, n, i0=i0, host=True)
    assert _eq(Ih.omega, Id.omega) and _eq(Ih.n, Id.n), tag
    fd = dev(f)
    bs, gs = A.auglik_potential_and_precision(lik, Wh, y, f)
    bsd, gsd = A.auglik_potential_and_precision(lik, Wd, dev(y), fd)
    assert _tup_eq(bs, bsd) and _tup_eq(gs, gsd), tag
    lt_h, lt_d = A.logtilt(lik, Wh, y, f), A.logtilt(lik, Wd, dev(y), fd)
    assert lt_h == pytest.approx(lt_d, rel=RTOL, abs=1e-12), tag
    if kind != CAT:
        al_h, al_d = A.aug_loglik(lik, Wh, y, f), A.aug_loglik(lik, Wd, dev(y), fd)
        assert al_h == pytest.approx(al_d, rel=1e-11, abs=1e-9), tag
    if oracle:
        olik = orc.make_lik(kind, *params, **kw)
        w_np = Wh.omega.numpy()
        n_np = Wh.n.numpy() if Wh.n is not None else np.zeros(w_np.shape, dtype=np.int64)
        ob, og = orc.potential_precision(olik, y, f, w_np, n_np)
        assert relerr(np.stack([t.numpy() for t in bs]), ob, floor=1.0) < RTOL, tag
        assert relerr(np.stack([t.numpy() for t in gs]), og) < RTOL, tag
        seq, comp = orc.sampled_loglik_terms(olik, y, f, w_np, n_np, with_prior=False)
        assert lt_h == pytest.approx(comp[3], rel=RTOL, abs=1e-12), tag


@pytest.mark.parametrize("name,kind,params,kw", CASES)
def test_host_verbs_match_device_verbs_and_oracle(A, orc, small_chunks, name, kind, params, kw):
    nl = kw.get("nlatent", 1)
    rows = _rows_per_chunk(small_chunks, nl) if kind in (CAT, CAT_BIJ) else small_chunks
    for j, n in enumerate(_sizes(rows)):
        _check_all_verbs(A, orc, name, kind, params, kw, n, 500 + j)


@pytest.mark.parametrize("name,kind,params,kw", [CASES[0], CASES[3], CASES[6]])
def test_host_verbs_at_the_default_chunk_size(A, orc, name, kind, params, kw):
    """the chunk the bench's e2e leg runs with: 4 Mi observations"""
    assert "AUGCUDA_HOST_CHUNK_LOG2" not in os.environ
    orc.set_threads(orc.max_threads())
    try:
        chunk = 1 << 22
        for j, n in enumerate([chunk - 1, chunk, chunk + 1, 3 * chunk + 12345]):
            _check_all_verbs(A, orc, name, kind, params, kw, n, 900 + j, oracle=(j == 3))
    finally:
        orc.set_threads(1)


def test_host_verbs_with_pageable_and_pinned_buffers(A, small_chunks):
    """pinned torch tensors, pageable torch tensors and numpy arrays give the same bits"""
    from gpu_common import make_lik
    lik = make_lik(POISSON, (10.0,), {})
    n = 3 * small_chunks + 77
    y, mu, var, f = synth_inputs(POISSON, n, 42, (10.0,))
    outs = []
    for mode in ("numpy", "pageable", "pinned"):
        conv = {"numpy": lambda a: a, "pageable": torch.from_numpy,
                "pinned": lambda a: torch.from_numpy(a).pin_memory()}[mode]
        q = A.init_aux_posterior(lik, n, host=True)
        q, b, g, s = A.cavi_step_(q, lik, conv(y), A.Normals(conv(mu), conv(var)))
        W = A.aux_sample(A.AugPhilox(3, 0), lik, conv(y), conv(f))
        outs.append((q.c.clone(), q.λ.clone(), b[0].clone(), g[0].clone(), s.clone(), W.omega.clone(), W.n.clone()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)


def test_host_and_device_arrays_cannot_be_mixed(A):
    from gpu_common import dev, make_lik
    lik = make_lik(BERNOULLI, (), {})
    y, mu, var, f = synth_inputs(BERNOULLI, 64, 1)
    q = A.init_aux_posterior(lik, 64)
    with pytest.raises(ValueError):
        A.cavi_step_(q, lik, y, A.Normals(dev(mu), dev(var)))


def test_host_verb_error_codes(A):
    """same error behaviour as the device verbs: non-bijective KL -> precondition (categorical.jl:165-170)"""
    from gpu_common import make_lik
    lik = make_lik(CAT, (), dict(nlatent=5))
    y, mu, var, f = synth_inputs(CAT, 8, 3, (), 5)
    q = A.init_aux_posterior(lik, 8, host=True)
    with pytest.raises(A.AugError) as ei:
        A.cavi_step_(q, lik, y, A.Normals(mu, var), want_elbo=True)
    assert ei.value.rc == -3
    W = A.aux_sample(A.AugPhilox(1, 0), lik, y, f)
    with pytest.raises(A.AugError) as ei:
        A.aug_loglik(lik, W, y, f)
    assert ei.value.rc == -3
