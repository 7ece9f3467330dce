"""Device-resident replay buffer with the reference's interface and sampling stream.

Mirrors ``utils/replaybuffer.py:14-42`` (``ReplayBuffer.add / get_size / sample_batch``) and the
index sampler of ``utils/custom_collections.py:103-131`` (``RandomAccessQueue.sample_n_k`` on a
``np.random.RandomState(seed)``), so a run with the same seed draws the same minibatches.  The
list-of-namedtuples storage is replaced by a ring in HBM; ``sample_batch`` is one gather kernel.
Logical FIFO index i <-> ring slot (head + i) % capacity.  Two layouts: ``"record"`` (default: one array of
64-byte-aligned fixed-stride records, ``rlc_replay_gather_rec`` -- a random transition is one contiguous DRAM read
instead of five, 2.5x less traffic than the struct-of-arrays gather at 1M transitions) and ``"soa"`` (five arrays
``state[cap,S] action[cap,A] reward[cap] next_state[cap,S] gamma[cap]``, ``rlc_replay_gather``)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check
from .engine import Engine, _ptr, _stream


class ReplayBuffer(object):
    def __init__(self, buffer_size, random_seed, state_dim=None, action_dim=None, engine: Engine = None,
                 flush_every: int = 256, sample_on_device: bool = False, layout: str = "record"):
        if layout not in ("soa", "record"):
            raise ValueError("layout must be 'soa' or 'record'")
        self.layout = layout
        # sample_on_device: draw the minibatch indices with rlc_replay_sample (Philox, no host round trip) instead of
        # the reference's numpy stream; falls back to the host sampler when 3k >= n or k > 4096
        self.sample_on_device = bool(sample_on_device)
        self._seed = int(random_seed) & 0xFFFFFFFFFFFFFFFF
        self._draws = 0
        self.buffer_size = int(buffer_size)
        self.rng = np.random.RandomState(random_seed)       # custom_collections.py:14-15
        self.eng = engine if engine is not None else Engine()
        self.S, self.A = state_dim, action_dim
        self._count = 0          # number of stored transitions (<= capacity)
        self._head = 0           # ring slot of logical index 0
        self._pending = []       # host staging of not-yet-flushed transitions
        self._flush_every = int(flush_every)
        self._alloc_done = False
        if state_dim is not None and action_dim is not None:
            self._alloc()

    def _alloc(self):
        dev, cap = self.eng.device, self.buffer_size
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        if self.layout == "record":
            S, A = self.S, self.A
            self.stride = int(self.eng.lib.rlc_replay_rec_stride(S, A))
            if self.stride < 0:
                check(self.stride)
            self.rec = z(cap, self.stride)
            # field views of the records (inspection / checkpointing; the kernels take `rec`)
            self.state, self.action, self.reward = self.rec[:, :S], self.rec[:, S:S + A], self.rec[:, S + A]
            self.next_state, self.gamma = self.rec[:, S + A + 1:2 * S + A + 1], self.rec[:, 2 * S + A + 1]
        else:
            self.state, self.next_state = z(cap, self.S), z(cap, self.S)
            self.action = z(cap, self.A)
            self.reward, self.gamma = z(cap), z(cap)
        self._alloc_done = True

    # utils/replaybuffer.py:25-27
    def add(self, state, action, reward, next_state, transition_gamma):
        state = np.asarray(state, np.float32).reshape(-1)
        action = np.asarray(action, np.float32).reshape(-1)
        next_state = np.asarray(next_state, np.float32).reshape(-1)
        if not self._alloc_done:
            self.S, self.A = state.size, action.size
            self._alloc()
        if state.size != self.S or next_state.size != self.S or action.size != self.A:
            raise ValueError("transition does not match the buffer's state/action dims")
        self._pending.append((state, action, np.float32(reward), next_state, np.float32(transition_gamma)))
        if len(self._pending) >= self._flush_every:
            self._flush()

    def _flush(self):
        if not self._pending:
            return
        pend = self._pending[-self.buffer_size:]            # older ones would be evicted anyway
        dropped = len(self._pending) - len(pend)
        self._pending = []
        cap = self.buffer_size
        # FIFO semantics of RandomAccessQueue.append with maxlen (custom_collections.py:85-88)
        for _ in range(dropped):
            self._advance_one()
        slots = np.empty(len(pend), np.int64)
        for i in range(len(pend)):
            slots[i] = self._advance_one()
        dev = self.eng.device
        h2d = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).to(dev, non_blocking=False)
        s_in = h2d(np.stack([p[0] for p in pend]))
        a_in = h2d(np.stack([p[1] for p in pend]))
        r_in = h2d(np.array([p[2] for p in pend], np.float32))
        s2_in = h2d(np.stack([p[3] for p in pend]))
        g_in = h2d(np.array([p[4] for p in pend], np.float32))
        slot_t = h2d(slots)
        if self.layout == "record":
            check(self.eng.lib.rlc_replay_scatter_rec(self.eng.h, _ptr(self.rec), cap, self.stride, self.S, self.A,
                                                      _ptr(slot_t), len(pend), _ptr(s_in), _ptr(a_in), _ptr(r_in),
                                                      _ptr(s2_in), _ptr(g_in), _stream()))
            return
        check(self.eng.lib.rlc_replay_scatter(self.eng.h, _ptr(self.state), _ptr(self.action), _ptr(self.reward),
                                              _ptr(self.next_state), _ptr(self.gamma), cap, self.S, self.A,
                                              _ptr(slot_t), len(pend), _ptr(s_in), _ptr(a_in), _ptr(r_in),
                                              _ptr(s2_in), _ptr(g_in), _stream()))

    def _advance_one(self):
        """Slot for the next appended transition; evicts the oldest when full."""
        cap = self.buffer_size
        if self._count < cap:
            slot = (self._head + self._count) % cap
            self._count += 1
        else:
            slot = self._head
            self._head = (self._head + 1) % cap
        return slot

    def get_size(self):
        return min(self._count + len(self._pending), self.buffer_size)

    def __len__(self):
        return self.get_size()

    def sample_indices(self, batch_size):
        """``RandomAccessQueue.sample_n_k(len(self), k)`` (custom_collections.py:107-131)."""
        n, k = self.get_size(), int(batch_size)
        rng = self.rng
        if not 0 <= k <= n:
            raise ValueError("Sample larger than population or is negative")
        if k == 0:
            return np.empty((0,), dtype=np.int64)
        if 3 * k >= n:
            return rng.choice(n, k, replace=False)
        result = rng.choice(n, 2 * k)
        selected = set()
        j = k
        for i in range(k):
            x = result[i]
            while x in selected:
                x = result[i] = result[j]
                j += 1
                if j == 2 * k:
                    result[k:] = rng.choice(n, k)
                    j = k
            selected.add(x)
        return result[:k]

    # utils/replaybuffer.py:32-37
    def sample_batch(self, batch_size, as_numpy=True):
        """Five numpy arrays like the reference (``map(np.array, zip(*batch))``, utils/replaybuffer.py:36-37), which is
        what ``BaseAgent.learn`` hands to ``update_network``; ``as_numpy=False`` keeps the gathered minibatch on the
        device (the drop-in networks accept both)."""
        assert self.get_size() >= batch_size
        self._flush()
        if self.sample_on_device and 0 < batch_size <= 4096 and 3 * batch_size < self.get_size():
            slots = torch.empty((int(batch_size),), dtype=torch.int64, device=self.eng.device)
            check(self.eng.lib.rlc_replay_sample(self.eng.h, self.get_size(), int(batch_size), self._seed, self._draws,
                                                 self._head, self.buffer_size, None, _ptr(slots), _stream()))
            self._draws += 1
            return self.gather_slots(slots, as_numpy=as_numpy)
        idx = np.asarray(self.sample_indices(batch_size), np.int64)
        slots = (self._head + idx) % self.buffer_size
        return self.gather_slots(slots, as_numpy=as_numpy)

    def gather_slots(self, slots, as_numpy=True):
        dev, B = self.eng.device, len(slots)
        slot_t = slots if isinstance(slots, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(slots, dtype=np.int64)).to(dev)
        e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        s, a, r, s2, g = e(B, self.S), e(B, self.A), e(B), e(B, self.S), e(B)
        if self.layout == "record":
            check(self.eng.lib.rlc_replay_gather_rec(self.eng.h, _ptr(self.rec), self.buffer_size, self.stride, self.S,
                                                     self.A, _ptr(slot_t), B, _ptr(s), _ptr(a), _ptr(r), _ptr(s2),
                                                     _ptr(g), _stream()))
        else:
            check(self.eng.lib.rlc_replay_gather(self.eng.h, _ptr(self.state), _ptr(self.action), _ptr(self.reward),
                                                 _ptr(self.next_state), _ptr(self.gamma), self.buffer_size, self.S,
                                                 self.A, _ptr(slot_t), B, _ptr(s), _ptr(a), _ptr(r), _ptr(s2), _ptr(g),
                                                 _stream()))
        out = (s, a, r, s2, g)
        if as_numpy:
            return tuple(t.cpu().numpy() for t in out)
        return out

    def clear(self):
        """utils/replaybuffer.py:40-42: the reference re-creates the queue WITHOUT a seed
        (``RandomAccessQueue(maxlen=self.buffer_size)`` -> ``RandomState(None)``), so the index stream after a clear
        is freshly (OS-)seeded there too; the ring is emptied in place."""
        self._count = self._head = 0
        self._pending = []
        self.rng = np.random.RandomState(None)
        self._seed = int(self.rng.randint(0, 2 ** 31 - 1))
        self._draws = 0
