"""GPU tests of the device-side minibatch index sampler (rlc_replay_sample): distinct, in range,
a pure function of (seed, counter), uniform; and the ReplayBuffer option that uses it."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sample(eng, n, k, seed, counter, head=0, cap=None):
    import torch
    from rlcontrol_b200._lib import check
    from rlcontrol_b200.engine import _ptr, _stream
    cap = n if cap is None else cap
    idx = torch.full((k,), -1, dtype=torch.int64, device=eng.device)
    slot = torch.full((k,), -1, dtype=torch.int64, device=eng.device)
    check(eng.lib.rlc_replay_sample(eng.h, n, k, seed, counter, head, cap, _ptr(idx), _ptr(slot), _stream()))
    return idx.cpu().numpy(), slot.cpu().numpy()


@pytest.mark.parametrize("n,k", [(100, 32), (1000, 32), (1_000_000, 4096), (13, 4), (4097 * 3, 4096), (50, 1)])
def test_device_sampler_distinct_in_range_deterministic(eng, n, k):
    a, sa = _sample(eng, n, k, seed=7, counter=3, head=5 % n, cap=n + 11)
    assert a.min() >= 0 and a.max() < n and len(set(a.tolist())) == k
    np.testing.assert_array_equal(sa, (5 % n + a) % (n + 11))
    b, _ = _sample(eng, n, k, seed=7, counter=3, head=5 % n, cap=n + 11)
    np.testing.assert_array_equal(a, b)                          # pure function of (seed, counter)
    c, _ = _sample(eng, n, k, seed=7, counter=4)
    d, _ = _sample(eng, n, k, seed=8, counter=3)
    if n > 100:
        assert not np.array_equal(a, c) and not np.array_equal(a, d)


def test_device_sampler_is_uniform(eng):
    n, k, draws = 64, 16, 4000
    counts = np.zeros(n)
    for c in range(draws):
        idx, _ = _sample(eng, n, k, seed=1, counter=c)
        counts[idx] += 1
    expect = draws * k / n
    # binomial std ~ sqrt(expect * (1 - k/n)) = 27; 5 sigma
    assert np.abs(counts - expect).max() < 5 * np.sqrt(expect)
    # pairs are not correlated with the slot order either: first slot uniform
    first = np.array([_sample(eng, n, k, seed=2, counter=c)[0][0] for c in range(2000)])
    assert np.abs(np.bincount(first, minlength=n) - 2000 / n).max() < 6 * np.sqrt(2000 / n)


def test_device_sampler_rejects_what_the_host_covers(eng):
    from rlcontrol_b200._lib import RlcError
    with pytest.raises(RlcError):
        _sample(eng, 30, 10, 0, 0)            # 3k >= n -> host sampler (choice without replacement)
    with pytest.raises(RlcError):
        _sample(eng, 100000, 5000, 0, 0)      # k > 4096


def test_replay_buffer_device_sampling_option(eng):
    from rlcontrol_b200.replaybuffer import ReplayBuffer
    rb_ = ReplayBuffer(500, 3, state_dim=3, action_dim=1, engine=eng, sample_on_device=True)
    for i in range(700):                       # wraps the ring: logical index 0 is transition 200
        rb_.add([i, i + 0.5, -i], [i * 0.25], float(i), [i + 1, i + 1.5, -i - 1], 0.99)
    s, a, r, s2, g = rb_.sample_batch(32, as_numpy=True)
    assert s.shape == (32, 3) and len(set(r.tolist())) == 32 and r.min() >= 200 and r.max() < 700
    np.testing.assert_array_equal(s[:, 0], r)
    np.testing.assert_array_equal(a[:, 0], r * 0.25)
    np.testing.assert_array_equal(s2[:, 0], r + 1)
    r2 = rb_.sample_batch(32, as_numpy=True)[2]
    assert not np.array_equal(r, r2)           # the call counter advances the stream
    # small buffer: falls back to the reference's host stream
    small = ReplayBuffer(64, 3, state_dim=1, action_dim=1, engine=eng, sample_on_device=True)
    ref = ReplayBuffer(64, 3, state_dim=1, action_dim=1, engine=eng)
    for i in range(40):
        small.add([i], [0.0], float(i), [i], 1.0)
        ref.add([i], [0.0], float(i), [i], 1.0)
    np.testing.assert_array_equal(small.sample_batch(32, as_numpy=True)[2], ref.sample_batch(32, as_numpy=True)[2])
