"""Pins the oracle's port of Eigen's LevenbergMarquardt / NumericalDiff / lmpar2 / qrsolv
(third-party, not vendored by the reference) against MINPACK lmdif via scipy, and the private
glibc rand() restatement against libc."""
import ctypes as C

import numpy as np
import pytest

scipy_opt = pytest.importorskip("scipy.optimize")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_lm_iterates_match_minpack(oracle_lib, seed):
    rng = np.random.default_rng(seed)
    m = 40
    t = np.linspace(0, 4, m)
    y = 2.5 * np.exp(-1.3 * t) + rng.normal(0, 0.05, m)
    x0 = np.array([1.0 + seed, 0.3 + 0.4 * seed])
    calls = []

    def fun(x):
        calls.append(x.copy())
        return y - x[0] * np.exp(-x[1] * t)

    sol, cov, info, msg, ier = scipy_opt.leastsq(fun, x0.copy(), full_output=True, ftol=1e-10, xtol=1e-10, maxfev=400)
    calls = np.array(calls)
    x = x0.copy()
    trace = np.zeros((60, 2))
    nfev = C.c_int(0)
    f = oracle_lib.lib.esvo_oracle_op_lm_expfit
    f.restype = C.c_int
    rc = f(_p(t, C.c_double), _p(y, C.c_double), m, _p(x, C.c_double), C.c_double(1e-10), C.c_double(1e-10), 400, 60,
           _p(trace, C.c_double), C.byref(nfev))
    status, k = divmod(rc, 1000)
    assert status in (1, 2, 3), (status, k)
    assert np.allclose(x, sol, rtol=1e-8, atol=1e-10)
    # every accepted iterate of the port was a trial point of MINPACK
    for i in range(k):
        d = np.abs(calls - trace[i]).max(axis=1)
        assert d.min() < 1e-7 * max(1.0, np.abs(trace[i]).max()), (i, trace[i])


def test_private_rand_matches_glibc(oracle_lib):
    libc = C.CDLL("libc.so.6")
    for seed in (1, 42, 123456):
        libc.srand(seed)
        ref = np.array([libc.rand() for _ in range(500)], np.int32)
        out = np.zeros(500, np.int32)
        oracle_lib.lib.esvo_oracle_op_rand(C.c_uint(seed), _p(out, C.c_int), 500)
        assert np.array_equal(out, ref)


def test_polar_factor(oracle_lib):
    rng = np.random.default_rng(5)
    for _ in range(20):
        A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        if np.linalg.det(A) < 0:
            A[:, 0] *= -1
        M = A + 1e-3 * rng.normal(size=(3, 3))
        Q = np.zeros(9)
        oracle_lib.lib.esvo_oracle_op_polar(_p(np.ascontiguousarray(M), C.c_double), _p(Q, C.c_double))
        U, _, Vt = np.linalg.svd(M)
        assert np.allclose(Q.reshape(3, 3), U @ Vt, atol=1e-12)
