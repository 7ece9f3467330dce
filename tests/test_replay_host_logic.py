"""Host logic of rlcontrol_b200.replaybuffer.ReplayBuffer on the CPU: ring bookkeeping (FIFO eviction, deferred flushes,
head/count), the reference's index stream and both storage layouts, against a list-based model of the reference buffer
(utils/replaybuffer.py:14-42 on RandomAccessQueue(maxlen), utils/custom_collections.py:85-131).

The kernels are NOT under test here: a stand-in `lib` moves the rows with numpy so that the host side can run without a GPU
(the product path has no such fallback -- Engine() raises without CUDA; the real kernels are compared with the oracle in
tests/test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle_np as onp
from rlcontrol_b200.replaybuffer import ReplayBuffer


def _f32(p, *shape):
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (n,)).reshape(shape)


def _i64(p, n):
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int64)), (n,))


class _NumpyMover:
    """Same signatures as the C-ABI entry points the buffer calls (include/rlc.h), rows moved with numpy."""
    calls = 0

    def rlc_replay_rec_stride(self, S, A):
        return (2 * S + A + 2 + 15) // 16 * 16

    def rlc_replay_scatter(self, h, st, ac, rw, s2, gm, cap, S, A, slot, n, s_in, a_in, r_in, s2_in, g_in, stream):
        j = _i64(slot, n)
        _f32(st, cap, S)[j] = _f32(s_in, n, S); _f32(ac, cap, A)[j] = _f32(a_in, n, A); _f32(rw, cap)[j] = _f32(r_in, n)
        _f32(s2, cap, S)[j] = _f32(s2_in, n, S); _f32(gm, cap)[j] = _f32(g_in, n)
        self.calls += 1
        return 0

    def rlc_replay_gather(self, h, st, ac, rw, s2, gm, cap, S, A, idx, B, s_o, a_o, r_o, s2_o, g_o, stream):
        j = _i64(idx, B)
        _f32(s_o, B, S)[:] = _f32(st, cap, S)[j]; _f32(a_o, B, A)[:] = _f32(ac, cap, A)[j]; _f32(r_o, B)[:] = _f32(rw, cap)[j]
        _f32(s2_o, B, S)[:] = _f32(s2, cap, S)[j]; _f32(g_o, B)[:] = _f32(gm, cap)[j]
        return 0

    def rlc_replay_scatter_rec(self, h, rec, cap, stride, S, A, slot, n, s_in, a_in, r_in, s2_in, g_in, stream):
        r, j = _f32(rec, cap, stride), _i64(slot, n)
        r[j, :S] = _f32(s_in, n, S); r[j, S:S + A] = _f32(a_in, n, A); r[j, S + A] = _f32(r_in, n)
        r[j, S + A + 1:2 * S + A + 1] = _f32(s2_in, n, S); r[j, 2 * S + A + 1] = _f32(g_in, n)
        self.calls += 1
        return 0

    def rlc_replay_gather_rec(self, h, rec, cap, stride, S, A, idx, B, s_o, a_o, r_o, s2_o, g_o, stream):
        r, j = _f32(rec, cap, stride), _i64(idx, B)
        _f32(s_o, B, S)[:] = r[j, :S]; _f32(a_o, B, A)[:] = r[j, S:S + A]; _f32(r_o, B)[:] = r[j, S + A]
        _f32(s2_o, B, S)[:] = r[j, S + A + 1:2 * S + A + 1]; _f32(g_o, B)[:] = r[j, 2 * S + A + 1]
        return 0


class _HostEngine:
    device = torch.device("cpu")
    h = None

    def __init__(self):
        self.lib = _NumpyMover()


@pytest.fixture(autouse=True)
def _no_cuda_stream(monkeypatch):
    import rlcontrol_b200.replaybuffer as rbm
    monkeypatch.setattr(rbm, "_stream", lambda: None)


@pytest.mark.parametrize("layout", ["soa", "record"])
@pytest.mark.parametrize("flush_every", [1, 7, 256, 10 ** 9])
def test_ring_bookkeeping_and_index_stream_match_the_reference_model(layout, flush_every):
    S, A, cap, seed = 3, 2, 50, 11
    buf = ReplayBuffer(cap, seed, S, A, engine=_HostEngine(), flush_every=flush_every, layout=layout)
    model, ref_rng = [], np.random.RandomState(seed)           # the reference: a FIFO list with maxlen + the same RandomState
    rng = np.random.RandomState(5)
    t = 0
    for burst, B in [(3, 1), (20, 8), (40, 16), (1, 16), (130, 50), (60, 5)]:   # under-full, exactly full, wrapped > 2x
        for _ in range(burst):
            tr = (rng.randn(S).astype(np.float32), rng.randn(A).astype(np.float32), np.float32(rng.randn()),
                  rng.randn(S).astype(np.float32), np.float32(0.0 if t % 9 == 0 else 0.99))
            buf.add(*tr)
            model.append(tr)
            model[:] = model[-cap:]
            t += 1
        assert buf.get_size() == len(model) == len(buf)
        got = buf.sample_batch(B, as_numpy=True)
        idx = onp.sample_n_k(ref_rng, len(model), B)
        want = [np.array(x) for x in zip(*[model[i] for i in idx])]        # map(np.array, zip(*batch)), replaybuffer.py:32-37
        for g, w in zip(got, want):
            assert g.dtype == np.float32 and np.array_equal(g, w)
    # same draws on both sides => the two generators are in the same state
    assert buf.rng.randint(1 << 30) == ref_rng.randint(1 << 30)
    # the ring in FIFO order is the model
    order = (buf._head + np.arange(len(model))) % cap
    assert np.array_equal(buf.state.numpy()[order], np.stack([m[0] for m in model]))
    assert np.array_equal(buf.reward.numpy()[order], np.array([m[2] for m in model], np.float32))
    buf.clear()
    assert buf.get_size() == 0


def test_pending_longer_than_capacity_keeps_the_newest(monkeypatch):
    S, A, cap = 2, 1, 8
    buf = ReplayBuffer(cap, 0, S, A, engine=_HostEngine(), flush_every=10 ** 9)
    rows = [(np.full(S, i, np.float32), np.full(A, -i, np.float32), float(i), np.full(S, i + .5, np.float32), 0.9) for i in range(29)]
    for r in rows:
        buf.add(*r)
    assert buf.get_size() == cap and buf.eng.lib.calls == 0      # nothing flushed yet
    s, a, r, s2, g = buf.sample_batch(cap, as_numpy=True)         # 3k >= n: the permutation branch of sample_n_k
    assert buf.eng.lib.calls == 1
    assert sorted(r.tolist()) == [float(i) for i in range(21, 29)]
    assert np.array_equal(s[:, 0], r) and np.array_equal(a[:, 0], -r) and np.array_equal(s2[:, 0], r + .5)


def test_argument_errors():
    eng = _HostEngine()
    with pytest.raises(ValueError):
        ReplayBuffer(4, 0, 2, 1, engine=eng, layout="aos")
    buf = ReplayBuffer(4, 0, 2, 1, engine=eng)
    with pytest.raises(ValueError):
        buf.add(np.zeros(3), np.zeros(1), 0.0, np.zeros(2), 0.9)
    with pytest.raises(AssertionError):
        buf.sample_batch(1)
    lazy = ReplayBuffer(4, 0, engine=eng, layout="record")       # dims taken from the first transition
    lazy.add(np.zeros(5), np.zeros(2), 1.0, np.ones(5), 0.9)
    assert (lazy.S, lazy.A, lazy.stride) == (5, 2, 16)
    s, a, r, s2, g = lazy.sample_batch(1, as_numpy=True)
    assert r[0] == 1.0 and np.all(s2 == 1.0)
