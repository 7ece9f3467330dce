"""Device-resident experiment loop: ``Experiment.run()`` of the reference
(experiment.py:48-214) with the environment, the replay ring and the agent all in HBM.

The reference crosses Python -> gym -> list-of-namedtuples replay -> torch on every environment step
(``run_episode_train``, experiment.py:101-161; ``BaseAgent.update/learn``, agents/base_agent.py:52-70).  Here one
training step -- ``env.step`` -> ``ReplayBuffer.add`` -> ``sample_batch`` -> ``update_network`` ->
``update_target_network`` -> ``sample_action`` -- is ONE captured CUDA graph (csrc/envloop.cu glue kernels around the
four-branch update of ``kl_networks``), replayed back to back with no host synchronisation inside a chunk of steps;
an evaluation session (``eval``, experiment.py:163-194: EvalEpisodes greedy episodes) is one more graph that runs its
episodes side by side.  The host only draws randomness a chunk ahead and reads the reward log a chunk behind:

* minibatch indices: the reference's own stream -- ``RandomAccessQueue.sample_n_k`` on ``RandomState(random_seed)``
  (utils/custom_collections.py:107-131, utils/replaybuffer.py:18), bit-exact;
* environment resets: gym 0.18.0's ``seeding.np_random(seed)`` + ``uniform(-[pi,1],[pi,1])`` (restated from the
  published source; gym is not vendored by the reference, parity unpinned);
* the policy's N(0,1) draws: the reference never seeds torch for these agents, so there is no stream to reproduce;
  they come from a per-run ``torch.Generator(seed)`` (runs are reproducible and independent of how a sweep
  interleaves them).

Results use the reference's names (``run_data`` of main.py:188-203, including its swapped ``train_time`` /
``eval_time`` keys)."""
from __future__ import annotations

import ctypes as C
import hashlib
import struct
import time
from typing import Sequence

import numpy as np
import torch

from ._lib import ENV_BIMODAL1D, ENV_PENDULUM, SB_MAX_B, RlcEnv, RlcSbNet, check
from .engine import _ptr, _stream
from .kl_networks import LOG_STD_MAX, LOG_STD_MIN

# (maxima1, maxima2, stddev1, stddev2, height1, height2) of reward_func, environments/environments.py
BIMODAL = {
    "Bimodal1DEnv": (-1.0, 1.0, 0.2, 0.2, 1.0, 1.5),               # :226-238
    "Bimodal1DEnv_uneq_var1": (-1.0, 1.0, 0.4, 0.2, 1.0, 1.5),     # :312-326
    "Bimodal1DEnv_uneq_var2": (-1.0, 1.0, 0.3, 0.1, 1.0, 1.5),     # :399-413
    "Bimodal1DEnv_uneq_var3": (-1.0, 1.0, 0.3, 0.1, 1.0, 1.0),     # :486-500
    "Bimodal1DEnv_eq_var1": (-0.6, 0.6, 0.2, 0.2, 1.0, 1.0),       # :573-587
    "Bimodal1DEnv_eq_var2": (-0.8, 0.8, 0.2, 0.2, 1.0, 1.0),       # :660-674
    "Bimodal1DEnv_eq_var3": (-1.0, 1.0, 0.2, 0.2, 1.0, 1.0),       # :747-761
}


def gym_np_random(seed: int) -> np.random.RandomState:
    """gym 0.18.0 ``utils/seeding.np_random``: RandomState seeded with the 32-bit words of the first 8 bytes of
    sha512(str(seed))."""
    seed = int(seed) % 2 ** 64
    digest = hashlib.sha512(str(seed).encode("utf8")).digest()[:8] + b"\0" * 4
    words = struct.unpack("3I", digest)
    big = sum(2 ** (32 * i) * w for i, w in enumerate(words))
    ints = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        ints.append(mod)
    rng = np.random.RandomState()
    rng.seed(ints or [0])
    return rng


class EnvSpec:
    """What ``create_environment(env_json)`` contributes (environments.py:16-75,158-200): the dims and bounds
    main.py:67-77 hands to the agent, the step limits, and the device descriptor of the dynamics."""

    def __init__(self, env_json: dict):
        name = env_json["environment"]
        self.name = name
        self.total_steps = int(env_json["TotalMilSteps"] * 1000000)
        self.eval_interval = int(env_json["EvalIntervalMilSteps"] * 1000000)
        self.eval_episodes = int(env_json["EvalEpisodes"])
        ep = int(env_json.get("EpisodeSteps", -1))
        d = RlcEnv()
        if name == "Pendulum-v0":
            d.kind, d.S, d.A = ENV_PENDULUM, 3, 1
            d.episode_limit = ep if ep != -1 else 200                  # gym registration: max_episode_steps=200
            self.state_min, self.state_max = -np.array([1., 1., 8.], np.float32), np.array([1., 1., 8.], np.float32)
            self.action_min, self.action_max = np.array([-2.], np.float32), np.array([2.], np.float32)
        elif name in BIMODAL:
            d.kind, d.S, d.A = ENV_BIMODAL1D, 1, 1
            d.episode_limit = ep if ep != -1 else 1
            for i, v in enumerate(BIMODAL[name]):
                d.p[i] = v
            self.state_min, self.state_max = np.array([-2.]), np.array([2.])
            self.action_min, self.action_max = np.array([-2.]), np.array([2.])
        else:
            raise NotImplementedError("device-resident environments: Pendulum-v0 and Bimodal1DEnv*; got %r" % name)
        self.desc = d
        self.state_dim, self.action_dim, self.episode_limit = int(d.S), int(d.A), int(d.episode_limit)

    def env_params(self) -> dict:
        """The ``env_params`` dict of main.py:67-77."""
        return dict(env_name=self.name, state_dim=self.state_dim, state_min=self.state_min, state_max=self.state_max,
                    action_dim=self.action_dim, action_min=self.action_min, action_max=self.action_max)

    def reset_states(self, rng: np.random.RandomState, n: int) -> np.ndarray:
        """n consecutive ``env.reset()`` internal states [n,2] (float64)."""
        out = np.zeros((max(n, 1), 2), np.float64)
        if self.desc.kind == ENV_PENDULUM:
            high = np.array([np.pi, 1.0])
            for i in range(n):
                out[i] = rng.uniform(low=-high, high=high)
        return out


def sample_n_k(rng: np.random.RandomState, n: int, k: int) -> np.ndarray:
    """``RandomAccessQueue.sample_n_k`` (utils/custom_collections.py:107-131) on the caller's RandomState: same
    draws, same result; the common no-collision case skips the Python loop (it would not change anything)."""
    if not 0 <= k <= n:
        raise ValueError("Sample larger than population or is negative")
    if k == 0:
        return np.empty((0,), dtype=np.int64)
    if 3 * k >= n:
        return rng.choice(n, k, replace=False)
    # RandomState.choice(int n, size) without p is randint(0, n, size) on the same stream (same draws, same state
    # afterwards: tests/test_oracle_env.py) at half the call overhead
    result = rng.randint(0, n, size=2 * k)
    if len(set(result[:k].tolist())) == k:
        return result[:k]
    selected = set()
    j = k
    for i in range(k):
        x = result[i]
        while x in selected:
            x = result[i] = result[j]
            j += 1
            if j == 2 * k:
                result[k:] = rng.randint(0, n, size=k)
                j = k
        selected.add(x)
    return result[:k]


_HOST_LIB = None


def _host_lib():
    """librlc_host.so (csrc/hostrng.c): the same index stream for a whole chunk of steps in C; None when not built."""
    global _HOST_LIB
    if _HOST_LIB is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "librlc_host.so")
        try:
            lib = C.CDLL(path)
            lib.rlc_host_sample_chunk.restype = C.c_int
            lib.rlc_host_sample_chunk.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            _HOST_LIB = lib
        except OSError:
            _HOST_LIB = False
    return _HOST_LIB or None


def sample_chunk(rng: np.random.RandomState, sizes: np.ndarray, k: int, out: np.ndarray, use_c: bool = True):
    """``out[i] = sample_n_k(rng, sizes[i], k)`` for every step i with ``sizes[i] > 0``, in step order, on the caller's
    RandomState (same draws and same final state as calling :func:`sample_n_k` step by step).  Steps in the rejection
    branch (3k < n) go through the C helper in one call; the few early permutation-branch steps stay in Python."""
    sizes = np.ascontiguousarray(sizes, dtype=np.int64)
    n_steps = len(sizes)
    lib = _host_lib() if use_c else None
    i = 0
    while i < n_steps:
        if sizes[i] == 0:
            i += 1
            continue
        if lib is None or 3 * k >= sizes[i] or sizes[i] > 2 ** 32:
            out[i] = sample_n_k(rng, int(sizes[i]), k)
            i += 1
            continue
        # the longest run of steps the C helper covers (sizes never shrink below 3k again in a run, but be general)
        j = i
        while j < n_steps and (sizes[j] == 0 or (3 * k < sizes[j] <= 2 ** 32)):
            j += 1
        st = rng.get_state()
        key, pos = np.ascontiguousarray(st[1], dtype=np.uint32).copy(), C.c_int(int(st[2]))
        block = np.empty((j - i, k), np.int32)
        rc = lib.rlc_host_sample_chunk(key.ctypes.data, C.byref(pos), sizes[i:j].ctypes.data, j - i, int(k), block.ctypes.data)
        if rc != 0:
            raise RuntimeError("rlc_host_sample_chunk failed")
        rng.set_state((st[0], key, pos.value, st[3], st[4]))
        sel = sizes[i:j] > 0
        out[i:j][sel] = block[sel]
        i = j


class DeviceExperiment:
    """One run (one sweep INDEX) of ``Experiment`` on the device.

    ``network``: a drop-in ``ReverseKLNetwork`` / ``ForwardKLNetwork`` (kl_networks.py) built from the merged config.
    ``config`` supplies ``batch_size, gamma, warmup_steps, buffer_size, random_seed`` (utils/config.py:8-22).
    ``run()`` returns the tuple ``Experiment.run()`` returns; ``run_data()`` the dict main.py stores per run.
    ``begin() / launch_chunk() / finish()`` let a sweep interleave several runs on one GPU."""

    def __init__(self, network, env_json: dict, config, chunk_steps: int = 250, buffer_capacity: int | None = None,
                 steps_per_graph: int = 25):
        self.net, self.spec, self.cfg = network, EnvSpec(env_json), config
        sp = self.spec
        if (sp.state_dim, sp.action_dim) != (network.state_dim, network.action_dim):
            raise ValueError("network and environment dimensions differ")
        if getattr(config, "sample_for_eval", "False") in (True, "True"):
            raise NotImplementedError("sample_for_eval")          # greedy evaluation only (ReverseKL.py:64-70)
        if getattr(config, "norm_type", "none") == "layer":
            raise NotImplementedError("layer norm")
        self.B, self.gamma = int(config.batch_size), float(config.gamma)
        self.warmup = int(getattr(config, "warmup_steps", 0))
        self.seed = int(config.random_seed)
        self.K = int(chunk_steps)
        self.U = max(1, int(steps_per_graph))                          # learning steps captured per graph launch
        self.ring = 2 * self.K                                         # two halves: one in flight, one being filled
        # every stored transition needs a slot only until evicted: the run adds at most total_steps of them
        cap = int(getattr(config, "buffer_size", 1e6))
        self.cap = min(cap, sp.total_steps + 1) if buffer_capacity is None else int(buffer_capacity)
        self.eng, self.lib, self.dev = network.eng, network.eng.lib, network.device
        dev, S, A, B, E = self.dev, sp.state_dim, sp.action_dim, self.B, sp.eval_episodes
        z = lambda *sh, dt=torch.float32: torch.zeros(sh, dtype=dt, device=dev)
        # environment + loop state
        self.cur = z(8, dt=torch.int64)
        self.env_state, self.ep_step, self.obs = z(1, 2, dt=torch.float64), z(1, dt=torch.int32), z(1, S)
        self.ev_state, self.ev_step, self.ev_ret = z(E, 2, dt=torch.float64), z(E, dt=torch.int32), z(E, dt=torch.float64)
        self.ev_done, self.ev_obs, self.ev_cur = z(E, dt=torch.int32), z(E, S), z(2, dt=torch.int64)
        n_sessions = sp.total_steps // max(sp.eval_interval, 1) + 2
        self.ev_cursor_reset, self.ev_cursor_log = self.ev_cur[0:1], self.ev_cur[1:2]
        self.ev_ret_log, self.ev_steps_log = z(n_sessions, E, dt=torch.float64), z(n_sessions, E, dt=torch.int32)
        self.n_sessions = n_sessions
        # replay ring (struct of arrays, as replaybuffer.py)
        # replay ring in the RECORD layout (one 64-byte-aligned record [s | a | r | s' | gamma | pad] per transition, the
        # default of replaybuffer.ReplayBuffer): the five arrays the loop kernels take are field views sharing one row pitch
        self.rb_stride = int(self.lib.rlc_replay_rec_stride(S, A))
        self.rb_rec = z(self.cap, self.rb_stride)
        rec = self.rb_rec
        self.rb = dict(s=rec[:, :S], a=rec[:, S:S + A], r=rec[:, S + A], s2=rec[:, S + A + 1:2 * S + A + 1],
                       g=rec[:, 2 * S + A + 1])
        # per-step feeds and logs: ring of 2K rows, indexed by cur[0] on the device
        R = self.ring
        self.f_eps_act, self.f_eps_upd, self.f_idx = z(R, A), z(R, B, A), z(R, B, dt=torch.int32)
        self.reward_log, self.flag_log = z(R, dt=torch.float64), z(R, dt=torch.int32)
        pin = lambda *sh, dt=torch.float32: torch.zeros(sh, dtype=dt).pin_memory()
        self.h_eps_act, self.h_eps_upd, self.h_idx = pin(R, A), pin(R, B, A), pin(R, B, dt=torch.int32)
        self.h_reward, self.h_flag = pin(R, dt=torch.float64), pin(R, dt=torch.int32)
        # fixed buffers the graphs read
        self.eps_act, self.slots = z(1, A), z(B, dt=torch.int64)
        self.head_act, self.head_ev = z(1, 2 * A), z(E, 2 * A)
        f = lambda n: dict(action=z(n, A), logp=z(n), mean=z(n, A), mu_raw=z(n, A), log_std=z(n, A), z=z(n, A))
        self.act, self.ev_act = f(1), f(E)
        # host-side streams of randomness
        self.rng_replay = np.random.RandomState(self.seed)              # utils/replaybuffer.py:18
        self.gen = torch.Generator().manual_seed(self.seed)
        n_train_resets = sp.total_steps // sp.episode_limit + 2 if sp.desc.kind == ENV_PENDULUM else 1
        self.train_resets = torch.from_numpy(sp.reset_states(gym_np_random(self.seed), n_train_resets)).to(dev)
        n_ev = n_sessions * E if sp.desc.kind == ENV_PENDULUM else 1
        self.eval_resets = torch.from_numpy(sp.reset_states(gym_np_random(self.seed), n_ev)).to(dev)
        self.st = network._build_step(B)
        self.stream = self.st.stream
        self._built = False
        # small minibatches: the fused launches of csrc/small_batch.cu, with the Polyak step folded into V's Adam and
        # the acting / evaluation policy passes as one launch each
        self.small = network._small_ok(B) and E <= SB_MAX_B
        if self.small:
            network._build_small(self.st, B)
            self.st.sb_upd[0].target, self.st.sb_upd[0].tau = network.target_v.theta.data_ptr(), network.tau
            self.sb_act, self.sb_ev = self._policy_pass(self.obs, self.eps_act, self.head_act, self.act), \
                self._policy_pass(self.ev_obs, None, self.head_ev, self.ev_act)

    def _policy_pass(self, obs, eps, head, ev):
        net, n = self.net, (RlcSbNet * 1)()
        d = n[0]
        d.theta, d.inp, d.H1, d.H2, d.O = net.pi.theta.data_ptr(), net.pi.inp, net.pi.H1, net.pi.H2, net.pi.O
        d.x0, d.n0, d.n1, d.out, d.policy = obs.data_ptr(), net.pi.inp, 0, head.data_ptr(), 1
        d.eps = None if eps is None else eps.data_ptr()
        d.action_scale, d.log_std_min, d.log_std_max = net.action_scale, LOG_STD_MIN, LOG_STD_MAX
        d.action, d.logp, d.mean, d.mu_raw, d.log_std, d.z = (ev[k].data_ptr() for k in
                                                              ("action", "logp", "mean", "mu_raw", "log_std", "z"))
        return n

    # ------------------------------------------------------------------ device pieces (enqueue on the current stream)
    def _act(self):
        net = self.net
        if self.small:
            check(self.lib.rlc_sb_forward(self.eng.h, self.sb_act, 1, 1, _stream()))
            return
        net.pi.forward(self.obs, out=self.head_act)
        net.eng_pi.policy_evaluate(self.head_act, self.eps_act, net.action_scale, LOG_STD_MIN, LOG_STD_MAX, out=self.act)

    def _env_step(self):
        sp, rb = self.spec, self.rb
        check(self.lib.rlc_env_step_train(self.eng.h, C.byref(sp.desc), _ptr(self.cur), _ptr(self.env_state),
                                          _ptr(self.ep_step), _ptr(self.obs), _ptr(self.act["action"]),
                                          _ptr(self.train_resets), self.train_resets.shape[0], _ptr(rb["s"]),
                                          _ptr(rb["a"]), _ptr(rb["r"]), _ptr(rb["s2"]), _ptr(rb["g"]), self.cap,
                                          self.gamma, self.ring, _ptr(self.reward_log), _ptr(self.flag_log), self.rb_stride,
                                          _stream()))

    def _stage(self, learn: bool):
        sp = self.spec
        check(self.lib.rlc_loop_stage(self.eng.h, _ptr(self.cur), self.B, sp.action_dim, self.ring, _ptr(self.f_eps_act),
                                      _ptr(self.f_eps_upd) if learn else None, _ptr(self.f_idx) if learn else None,
                                      self.cap, _ptr(self.eps_act), _ptr(self.st.d["eps"]) if learn else None,
                                      _ptr(self.slots) if learn else None, _stream()))

    def _train_step(self, learn: bool):
        """experiment.py:118-142 for one step; ``learn`` = BaseAgent.learn's size test (base_agent.py:64-66)."""
        sp, rb, d = self.spec, self.rb, self.st.d
        # env.step + replay append + this step's feeds + minibatch gather: one launch (rlc_loop_step)
        check(self.lib.rlc_loop_step(
            self.eng.h, C.byref(sp.desc), _ptr(self.cur), _ptr(self.env_state), _ptr(self.ep_step), _ptr(self.obs),
            _ptr(self.act["action"]), _ptr(self.train_resets), self.train_resets.shape[0], _ptr(rb["s"]), _ptr(rb["a"]),
            _ptr(rb["r"]), _ptr(rb["s2"]), _ptr(rb["g"]), self.cap, self.gamma, self.ring, _ptr(self.reward_log),
            _ptr(self.flag_log), self.B, _ptr(self.f_eps_act), _ptr(self.f_eps_upd) if learn else None,
            _ptr(self.f_idx) if learn else None, _ptr(self.eps_act), _ptr(d["eps"]) if learn else None,
            _ptr(d["s"]) if learn else None, _ptr(d["a"]) if learn else None, _ptr(d["r"]) if learn else None,
            _ptr(d["s2"]) if learn else None, _ptr(d["g"]) if learn else None, self.rb_stride, _stream()))
        if learn:
            self.net._enqueue(self.st, self.B, device_inputs=True)
            if not self.small:
                self.net.eng_v.soft_update(self.net.target_v.theta, self.net.v.theta, self.net.tau)
        self._act()

    def _eval_session(self):
        """eval() (experiment.py:163-194): EvalEpisodes greedy episodes, side by side."""
        sp, net, E = self.spec, self.net, self.spec.eval_episodes
        check(self.lib.rlc_env_reset(self.eng.h, C.byref(sp.desc), E, _ptr(self.eval_resets), self.eval_resets.shape[0],
                                     _ptr(self.ev_cursor_reset), _ptr(self.ev_state), _ptr(self.ev_step),
                                     _ptr(self.ev_ret), _ptr(self.ev_done), _ptr(self.ev_obs), _stream()))
        for _ in range(sp.episode_limit):
            if self.small:
                check(self.lib.rlc_sb_forward(self.eng.h, self.sb_ev, 1, E, _stream()))
            else:
                net.pi.forward(self.ev_obs, out=self.head_ev)
                net.eng_pi.policy_evaluate(self.head_ev, None, net.action_scale, LOG_STD_MIN, LOG_STD_MAX, out=self.ev_act)
            check(self.lib.rlc_env_step_eval(self.eng.h, C.byref(sp.desc), E, _ptr(self.ev_state), _ptr(self.ev_step),
                                             _ptr(self.ev_ret), _ptr(self.ev_done), _ptr(self.ev_obs),
                                             _ptr(self.ev_act["mean"]), _stream()))
        check(self.lib.rlc_eval_store(self.eng.h, E, _ptr(self.ev_ret), _ptr(self.ev_step), _ptr(self.ev_cursor_log),
                                      self.n_sessions, _ptr(self.ev_ret_log), _ptr(self.ev_steps_log), _stream()))

    def _reset_train_env(self):
        sp = self.spec
        check(self.lib.rlc_env_reset(self.eng.h, C.byref(sp.desc), 1, _ptr(self.train_resets), self.train_resets.shape[0],
                                     _ptr(self.cur[3:4]), _ptr(self.env_state), _ptr(self.ep_step), None, None,
                                     _ptr(self.obs), _stream()))

    def _capture(self, fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream):
            fn()
        return g

    def _build(self):
        """Warm every kernel once (workspace growth must not happen inside a capture), capture the three graphs, then put
        the agent and the loop back to their initial state."""
        net = self.net
        ts, saved = net._snapshot()
        tv = net.target_v.theta.clone()
        with torch.cuda.stream(self.stream):
            self._reset_train_env()
            self._act()
            self.cur[1] = self.B + 1                               # pretend the ring holds enough rows for the warm-up
            self._train_step(True)
            self._train_step(False)
            self._eval_session()
        self.stream.synchronize()
        self.g_learn = self._capture(lambda: self._train_step(True))
        # U consecutive learning steps as ONE graph: the host's launch cost per step drops by U (with 8 runs sharing a
        # GPU the single host thread, not the device, was the limit)
        self.g_learn_u = self._capture(lambda: [self._train_step(True) for _ in range(self.U)]) if self.U > 1 else None
        self.g_nolearn = self._capture(lambda: self._train_step(False))
        self.g_eval = self._capture(self._eval_session)
        for t, s_ in zip(ts, saved):
            t.copy_(s_)
        net.target_v.theta.copy_(tv)
        net.critic.invalidate()
        net.critic_grid.invalidate()
        self.cur.zero_()
        self.ev_cur.zero_()
        torch.cuda.synchronize(self.dev)
        self._built = True

    # ------------------------------------------------------------------ host side
    def begin(self):
        if not self._built:
            self._build()
        sp = self.spec
        self.t = 0                                   # total_step_count
        self.replay_n = 0                            # host mirror of the replay size (deterministic for these envs)
        self.ep_t = 0                                # step count inside the current episode
        self.chunk = 0
        self.events = {}
        self.pending = []                            # (chunk, n_steps) whose logs are not read yet
        self.sessions = 0
        self.train_rewards_per_episode, self.train_steps_per_episode, self.train_cum_steps = [], [], []
        self.timesteps_at_eval, self.train_episodes = [], 0
        self._ep_reward, self._ep_len, self._seen = 0.0, 0, 0
        self.t_begin = time.time()
        with torch.cuda.stream(self.stream):
            # evaluate once at the beginning (experiment.py:57-59), then reset + agent.start (:105-111)
            self.g_eval.replay()
            self.sessions += 1
            self.timesteps_at_eval.append(0)
            self._reset_train_env()
            self.eps_act.copy_(torch.randn(1, sp.action_dim, generator=self.gen).to(self.dev, non_blocking=False))
            self._act()
        self.train_episodes = 1

    @property
    def finished(self) -> bool:
        return self.t >= self.spec.total_steps

    def _fill_feeds(self, half: int, n: int):
        """Draw chunk ``self.chunk``'s randomness into the pinned half; returns the per-step learn flags."""
        sp, B, A, K = self.spec, self.B, self.spec.action_dim, self.K
        lo = half * K
        self.h_eps_act[lo:lo + n] = torch.randn(n, A, generator=self.gen)
        self.h_eps_upd[lo:lo + n] = torch.randn(n, B, A, generator=self.gen)
        idx = self.h_idx.numpy()
        bandit = sp.desc.kind == ENV_BIMODAL1D
        thresh = max(self.warmup, B)
        # per-step bookkeeping of the reference loop, vectorised over the chunk: position inside the episode, whether the
        # step is cut by the episode limit (not stored, experiment.py:127-134), replay size after the step
        L = 1 if bandit else sp.episode_limit
        pos = (self.ep_t + np.arange(n)) % L + 1
        truncated = np.zeros(n, bool) if bandit else pos == L
        sizes = np.minimum(self.replay_n + np.cumsum(~truncated), self.cap)
        learn = sizes > thresh                                  # BaseAgent.learn (base_agent.py:64-66)
        self.ep_t, self.replay_n = int(pos[-1] % L), int(sizes[-1])
        sample_chunk(self.rng_replay, np.where(learn, sizes, 0), B, idx[lo:lo + n])
        return learn

    def launch_chunk(self):
        """Enqueue the next chunk of steps (asynchronous), then read the logs of the one before it."""
        sp, K = self.spec, self.K
        n = min(K, sp.total_steps - self.t)
        if n <= 0:
            return
        half = self.chunk & 1
        ev_old = self.events.pop(self.chunk - 2, None)
        if ev_old is not None:
            ev_old.synchronize()                     # this half's pinned buffers are free again
            self._drain(self.chunk - 2)
        learn = self._fill_feeds(half, n)
        lo = half * K
        with torch.cuda.stream(self.stream):
            for dst, src in ((self.f_eps_act, self.h_eps_act), (self.f_eps_upd, self.h_eps_upd), (self.f_idx, self.h_idx)):
                dst[lo:lo + n].copy_(src[lo:lo + n], non_blocking=True)
            i, U, ei = 0, self.U, sp.eval_interval
            while i < n:
                # a U-step graph when the next U steps all learn and no evaluation session falls inside them
                if self.g_learn_u is not None and i + U <= n and learn[i:i + U].all() and \
                        (ei <= 0 or (self.t % ei) + U <= ei):
                    self.g_learn_u.replay()
                    self.t += U
                    i += U
                else:
                    (self.g_learn if learn[i] else self.g_nolearn).replay()
                    self.t += 1
                    i += 1
                if ei > 0 and self.t % ei == 0:
                    self.g_eval.replay()
                    self.sessions += 1
                    self.timesteps_at_eval.append(self.t)
            self.h_reward[lo:lo + n].copy_(self.reward_log[lo:lo + n], non_blocking=True)
            self.h_flag[lo:lo + n].copy_(self.flag_log[lo:lo + n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        # the replays moved theta_Q on the device: drop the host-side validity of the cached tensor-core operand packs, so
        # an eager evaluation between chunks (q_net.eval_grid, getQFunction) repacks from the current parameters
        self.net.critic.invalidate()
        self.net.critic_grid.invalidate()
        self.events[self.chunk] = ev
        self.pending.append((self.chunk, n))
        self.chunk += 1

    def _drain(self, chunk: int):
        """Fold a finished chunk's reward/flag log into the per-episode lists (experiment.py:62-75)."""
        for j, (c, n) in enumerate(self.pending):
            if c == chunk:
                self.pending.pop(j)
                break
        else:
            return
        lo = (chunk & 1) * self.K
        r, f = self.h_reward.numpy()[lo:lo + n], self.h_flag.numpy()[lo:lo + n]
        for i in range(n):
            self._ep_reward += float(r[i])
            self._ep_len += 1
            self._seen += 1
            if f[i] & 1:
                self.train_rewards_per_episode.append(self._ep_reward)
                self.train_steps_per_episode.append(self._ep_len)
                self.train_cum_steps.append(self._seen)
                self._ep_reward, self._ep_len = 0.0, 0
                if self._seen < self.spec.total_steps:
                    self.train_episodes += 1

    def finish(self):
        for c in sorted(self.events):
            self.events[c].synchronize()
        self.stream.synchronize()
        for c, _ in sorted(self.pending):
            self._drain(c)
        self.events = {}
        ns = self.sessions
        self.eval_rewards_per_episode = self.ev_ret_log[:ns].cpu().numpy().tolist()
        self.eval_steps_per_episode = self.ev_steps_log[:ns].cpu().numpy().tolist()
        self.wall = time.time() - self.t_begin
        return self.results()

    def run(self):
        self.begin()
        while not self.finished:
            self.launch_chunk()
        return self.finish()

    def results(self):
        """The tuple of Experiment.run() (experiment.py:96-98).  Train and evaluation time are not separable when the
        device runs ahead of the host: the whole wall time is reported as training time."""
        return (self.train_rewards_per_episode, self.eval_rewards_per_episode, self.train_steps_per_episode,
                self.eval_steps_per_episode, self.timesteps_at_eval, self.wall, 0.0, self.train_episodes,
                self.train_cum_steps)

    def run_data(self, env_json: dict) -> dict:
        """``run_data`` of main.py:188-203 (its ``eval_time`` / ``train_time`` keys are swapped there; kept)."""
        (ep_r, ev_r, ep_s, ev_s, t_ev, train_time, eval_time, train_ep, _) = self.results()
        return {
            "random_seed": self.seed,
            "total_timesteps": env_json["TotalMilSteps"] * 1000000,
            "eval_interval_timesteps": env_json["EvalIntervalMilSteps"] * 1000000,
            "episodes_per_eval": env_json["EvalEpisodes"],
            "eval_episode_rewards": np.array(ev_r), "eval_episode_steps": np.array(ev_s),
            "timesteps_at_eval": np.array(t_ev), "train_episode_steps": np.array(ep_s),
            "train_episode_rewards": np.array(ep_r), "total_train_episodes": train_ep,
            "eval_time": train_time, "train_time": eval_time,
        }


def run_interleaved(experiments: Sequence[DeviceExperiment]):
    """Several independent runs (sweep INDEX settings, cfg5) on one GPU: every run owns its stream and graphs; the
    host enqueues one chunk per run in turn, so the device always has work from all of them.  Replicas only."""
    for e in experiments:
        e.begin()
    while not all(e.finished for e in experiments):
        for e in experiments:
            if not e.finished:
                e.launch_chunk()
    return [e.finish() for e in experiments]
