"""CPU test: the plain-C scalar restatement (oracle/oracle_c.c -> oracle/_ref/liboracle_c.so, built by
__graft_entry__.build()) against the reference's golden vectors and the numpy oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle import oracle_np as onp

SO = os.path.join(ROOT, "oracle", "_ref", "liboracle_c.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    return C.CDLL(SO)


def _fp(x):
    x = np.ascontiguousarray(x, np.float32)
    return x, x.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", ["tin_cfg1.npz", "tin_400_300.npz"])
def test_c_tin_eval_matches_reference_golden(lib, name):
    g = golden(name)
    s, a = g["s"], g["a"]
    B, S = s.shape
    N, A = a.shape
    H1, H2 = g["W1"].shape[0], g["W2"].shape[0]
    keep = [_fp(x) for x in (s, a, g["W1"], g["b1"], g["W2"], g["b2"], g["W3"], g["b3"])]
    q = np.zeros(B * N, np.float32)
    rc = lib.oracle_tin_eval(keep[0][1], keep[1][1], 0, B, N, S, A, H1, H2, *[k[1] for k in keep[2:]],
                             q.ctypes.data_as(C.c_void_p), C.c_long(0), C.c_long(B * N))
    assert rc == 0
    np.testing.assert_allclose(q.reshape(B, N), g["q"], rtol=2e-5, atol=2e-6)


def test_c_tmid_eval_matches_numpy_oracle(lib):
    rng = np.random.RandomState(0)
    S, A, H1, H2, B, N = 4, 2, 30, 20, 5, 17
    p = [rng.randn(S, H1) * .3, rng.randn(H1) * .1, rng.randn(H1 + A, H2) * .3, rng.randn(H2) * .1, rng.randn(H2, 1), rng.randn(1)]
    s, a = rng.randn(B, S) * 2, rng.randn(B, N, A)
    smin, smax = -np.ones(S), np.ones(S) * 1.5
    keep = [_fp(x) for x in (s, a, *p, smin, smax)]
    q = np.zeros(B * N, np.float32)
    rc = lib.oracle_tmid_eval(keep[0][1], keep[1][1], 1, B, N, S, A, H1, H2, *[k[1] for k in keep[2:8]],
                              keep[8][1], keep[9][1], q.ctypes.data_as(C.c_void_p), C.c_long(0), C.c_long(B * N))
    assert rc == 0
    ref = onp.tmid_eval(s, a, p, smin, smax, dtype=np.float64)
    np.testing.assert_allclose(q.reshape(B, N), ref, rtol=2e-5, atol=2e-5)
