"""CPU tests: the C-ABI library loads and exports every symbol include/rlc.h declares (no
compute calls without a GPU), and the pure-host helpers behave."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rlc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rlc_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    from rlcontrol_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in rlc.h but not exported by librlc.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.rlc_version() == 100


def test_host_only_entry_points():
    from rlcontrol_b200 import _lib
    lib = _lib.load()
    assert lib.rlc_status_string(0) == b"ok"
    assert b"invalid" in lib.rlc_status_string(-1)
    # theta layout arithmetic (no device needed)
    n_tin = lib.rlc_theta_numel(_lib.TIN, 17, 6, 400, 300)
    assert n_tin == 23 * 400 + 400 + 400 * 300 + 300 + 300 + 1 == 130201
    n_tmid = lib.rlc_theta_numel(_lib.TMID, 1, 1, 200, 200)
    assert n_tmid == 200 + 200 + 201 * 200 + 200 + 200 + 1
    off = (ctypes.c_int64 * 6)()
    assert lib.rlc_theta_offsets(_lib.TMID, 1, 1, 200, 200, off) == 0
    assert list(off) == [0, 200, 400, 400 + 201 * 200, 400 + 201 * 200 + 200, 400 + 201 * 200 + 400]
    assert lib.rlc_theta_numel(7, 1, 1, 1, 1) == -1
    assert lib.rlc_theta_offsets(_lib.TIN, 0, 1, 1, 1, off) == -1
    # dispatcher pins of the round-2 tensor-core paths: per-thread state, previous setting returned, junk = default
    for fn in (lib.rlc_rows_gemm_force, lib.rlc_tmid_tc_force):
        assert fn(0) == -1 and fn(2) == 0 and fn(7) == 2 and fn(-1) == -1
    # null / malformed arguments are rejected before any CUDA call
    assert lib.rlc_rows_gemm(None, 0, 0, 1, 1, 1, None, 1, None, 1, None, 1, None, None, 0, 0, 1.0, 0, 0, None) == -1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rlcontrol_b200 as rb
    with pytest.raises(rb.RlcError):
        rb.Engine(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under rlcontrol_b200/ may reference it."""
    pkg = os.path.join(ROOT, "rlcontrol_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_invalid_arguments_return_status_not_crash():
    """C-ABI error behaviour: bad shapes / null pointers come back as RLC_ERR_INVALID through
    the int status, nothing throws across the ABI; the Python layer maps it to RlcError(RuntimeError)."""
    import ctypes as C
    from rlcontrol_b200 import _lib
    lib = _lib.load()
    # no handle, no device needed: argument validation happens first
    cr = _lib.RlcCritic(0, 3, 1, 16, 16, None, None, None)
    assert lib.rlc_critic_eval(None, C.byref(cr), None, 1, None, 1, 0, 0, None, None) == -1
    assert lib.rlc_reduce_topk(None, None, 1, 4, 2, None, None, None, 0, 0, None, None) == -1
    assert lib.rlc_reduce_fkl(None, None, None, None, 1, 1, 0.1, 1, None, None, None, None) == -1
    assert lib.rlc_cem(None, C.byref(cr), None, 1, 1, 1, 1, 1, None, None, None, None, None, None, None, None, None, None, None) == -1
    assert lib.rlc_adam_step(None, None, None, None, None, 4, 1, 1e-3, .9, .999, 1e-8, 0, None, 0.0, None) == -1
    assert lib.rlc_umma_mode(C.byref(cr), 0) == -1          # theta == NULL -> not a valid critic
    assert lib.rlc_destroy(None) == 0 and lib.rlc_launch_count(None) == 0
    with pytest.raises(_lib.RlcError):
        _lib.check(-1)
    with pytest.raises(RuntimeError):
        _lib.check(-5)


def test_gather_reciprocal_is_exact():
    """csrc/replay.cu k_replay_gather_flat maps output float i -> sampled row i // E with __umulhi(i, 2^32 // E + 1);
    the launcher takes that kernel only for E = 2S+A+2 <= 8192 and at most 32 rows per CTA (the record-layout kernels
    use the same split per field, widths S, A <= 4000).  Every (E, i) it can see:"""
    for E in range(2, 8193):
        magic = np.uint64((1 << 32) // E + 1)
        i = np.arange(0, 32 * E, dtype=np.uint64)
        assert np.array_equal((i * magic) >> np.uint64(32), i // np.uint64(E)), E
