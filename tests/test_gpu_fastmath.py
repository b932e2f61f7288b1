"""Accuracy of the straight-line fp64 functions the map/reduce kernels use (csrc/aug_fastmath.cuh),
measured on the GPU against numpy's correctly-rounded-ish libm.  Bar: a few ulp — three orders of magnitude
inside the 1e-12 parity tolerance of the path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = 2.220446049250313e-16


@pytest.fixture(scope="module")
def A():
    from gpu_common import pkg
    return pkg()


def _run(A, fn, x):
    from gpu_common import dev, host
    return host(A.fastmath_eval(fn, dev(x)))


def test_rcp_rsqrt_sqrt(A):
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-280, 280, 200000), rng.uniform(0.5, 4.0, 200000), [1.0, 2.0, 1e-290, 1e290]])
    assert np.max(np.abs(_run(A, 0, x) * x - 1)) < 3 * EPS
    assert np.max(np.abs(_run(A, 1, x) * np.sqrt(x) - 1)) < 3 * EPS
    assert np.max(np.abs(_run(A, 5, x) / np.sqrt(x) - 1)) < 3 * EPS


def test_exp(A):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-708, 708, 400000), rng.uniform(-2, 2, 200000), [0.0, -708.0, 708.0, -1e-300]])
    got, ref = _run(A, 2, x), np.exp(x)
    assert np.max(np.abs(got / ref - 1)) < 4 * EPS


def test_log(A):
    rng = np.random.default_rng(2)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 300000), rng.uniform(0.5, 2.0, 300000),
                        1 + rng.uniform(-1e-3, 1e-3, 100000), [1.0, 2.2250738585072014e-308, 1.7e308]])
    got, ref = _run(A, 3, x), np.log(x)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-15
    assert np.max(np.abs(got - ref)[np.abs(ref) < 1e-3]) < 4e-19 + 4 * EPS * 1e-3
    d = 1 + rng.uniform(0, 1, 300000)
    got, ref = _run(A, 4, d), np.log(d)
    assert np.max(np.abs(got - ref)) < 3 * EPS
