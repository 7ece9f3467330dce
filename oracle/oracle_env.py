"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the reference's experiment loop
and of the two environments the device-resident loop (rlcontrol_b200/device_loop.py, csrc/envloop.cu) covers.

* ``run_experiment``     /root/reference/experiment.py:48-214 (run, run_episode_train, eval, run_episode_eval) +
                         agents/base_agent.py:35-70 (start/step/update/learn) + agents/ReverseKL.py:31-90 (take_action,
                         update_network) + utils/replaybuffer.py:14-42, with the policy's N(0,1) draws fed in.
* ``Bimodal1D``          /root/reference/environments/environments.py:158-764.  Parity: PINNED on the reference classes
                         (tests/golden/bimodal_env.npz, recorded by oracle/make_golden.py ``gen_bimodal_env``).
* ``PendulumV0``         gym==0.18.0 (requirements.txt:9) ``envs/classic_control/pendulum.py`` behind ``TimeLimit``
                         (max_episode_steps=200) and ``utils/seeding.np_random``.  gym is a third-party dependency that
                         is not vendored under /root/reference and not installed here: restated from its published
                         source, **parity unpinned**; anchored on the reference's call sites
                         (environments.py:47,95-111) and on known answers (tests/test_oracle_env.py).
"""
import hashlib
import math
import struct

import numpy as np

from . import oracle_np as onp

BIMODAL = {
    "Bimodal1DEnv": (-1.0, 1.0, 0.2, 0.2, 1.0, 1.5),
    "Bimodal1DEnv_uneq_var1": (-1.0, 1.0, 0.4, 0.2, 1.0, 1.5),
    "Bimodal1DEnv_uneq_var2": (-1.0, 1.0, 0.3, 0.1, 1.0, 1.5),
    "Bimodal1DEnv_uneq_var3": (-1.0, 1.0, 0.3, 0.1, 1.0, 1.0),
    "Bimodal1DEnv_eq_var1": (-0.6, 0.6, 0.2, 0.2, 1.0, 1.0),
    "Bimodal1DEnv_eq_var2": (-0.8, 0.8, 0.2, 0.2, 1.0, 1.0),
    "Bimodal1DEnv_eq_var3": (-1.0, 1.0, 0.2, 0.2, 1.0, 1.0),
}


def gym_np_random(seed):
    """gym 0.18.0 utils/seeding.py: np_random -> create_seed -> hash_seed -> _bigint_from_bytes ->
    _int_list_from_bigint -> RandomState.seed(list of uint32)."""
    seed = int(seed) % 2 ** (8 * 8)
    h = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    h += b"\0" * (4 - len(h) % 4)
    vals = struct.unpack("{}I".format(len(h) // 4), h)
    big = 0
    for i, v in enumerate(vals):
        big += 2 ** (32 * i) * v
    ints = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        ints.append(mod)
    rng = np.random.RandomState()
    rng.seed(ints if ints else [0])
    return rng


class PendulumV0:
    """gym 0.18.0 PendulumEnv + TimeLimit.  step() takes the agent's float32 action array."""
    name = "Pendulum-v0"
    state_dim, action_dim = 3, 1

    def __init__(self, episode_steps=-1):
        self.max_speed, self.max_torque, self.dt, self.g, self.m, self.l = 8, 2., .05, 10.0, 1., 1.
        self.EPISODE_STEPS_LIMIT = 200 if episode_steps == -1 else int(episode_steps)
        self.np_random = gym_np_random(0)             # gym draws a random seed until env.seed() is called
        self._elapsed = 0

    def set_random_seed(self, seed):
        self.np_random = gym_np_random(seed)

    def _obs(self):
        th, thdot = self.state
        return np.array([np.cos(th), np.sin(th), thdot])

    def reset(self):
        high = np.array([np.pi, 1])
        self.state = self.np_random.uniform(low=-high, high=high)
        self._elapsed = 0
        return self._obs()

    def step(self, u):
        th, thdot = self.state
        g, m, l, dt = self.g, self.m, self.l, self.dt
        u = np.float64(np.clip(u, -self.max_torque, self.max_torque)[0])     # numpy<2: float32 scalar (op) python float -> float64
        costs = (((th + np.pi) % (2 * np.pi)) - np.pi) ** 2 + .1 * thdot ** 2 + .001 * (u ** 2)
        newthdot = thdot + (-3 * g / (2 * l) * np.sin(th + np.pi) + 3. / (m * l ** 2) * u) * dt
        newth = th + newthdot * dt
        newthdot = np.clip(newthdot, -self.max_speed, self.max_speed)
        self.state = np.array([newth, newthdot])
        self._elapsed += 1
        done = self._elapsed >= self.EPISODE_STEPS_LIMIT                       # TimeLimit
        return self._obs(), -costs, done, {}


class Bimodal1D:
    """The one-step bandits of environments.py (all seven share this body; only reward_func's constants differ)."""
    state_dim, action_dim = 1, 1

    def __init__(self, name, episode_steps=-1):
        self.name = name
        self.m1, self.m2, self.sd1, self.sd2, self.h1, self.h2 = BIMODAL[name]
        self.EPISODE_STEPS_LIMIT = 1 if episode_steps == -1 else int(episode_steps)

    def set_random_seed(self, seed):
        pass

    def reset(self):
        self.state = np.array([0.])
        return self.state

    def reward_func(self, action):
        # ``action`` is the agent's float32 ARRAY: (action - m) / sd, ** 2 and -0.5 * stay float32; math.exp is double
        action = np.asarray(action, np.float32).reshape(1)
        modal1 = self.h1 * math.exp(float((-0.5 * ((action - self.m1) / self.sd1) ** 2)[0]))
        modal2 = self.h2 * math.exp(float((-0.5 * ((action - self.m2) / self.sd2) ** 2)[0]))
        return modal1 + modal2

    def step(self, action):
        self.state = self.state + action
        return self.state, self.reward_func(action), True, {}


def make_env(env_json):
    name = env_json["environment"]
    ep = env_json.get("EpisodeSteps", -1)
    return PendulumV0(ep) if name == "Pendulum-v0" else Bimodal1D(name, ep)


def run_experiment(agent, env_json, seed, batch_size, gamma, draws, warmup_steps=0):
    """Experiment.run() with an ``oracle_kl.KLAgent``.  ``draws(kind, t)`` returns the N(0,1) array the policy
    consumes: kind 'act' -> [1,A] for the sample_action that picks the action of step t (0-based), kind 'upd' ->
    [B,A] for the update after step t.  Returns a dict with the reference's result lists plus the per-step rewards."""
    train_env, test_env = make_env(env_json), make_env(env_json)
    train_env.set_random_seed(seed)
    test_env.set_random_seed(seed)
    total = env_json["TotalMilSteps"] * 1000000
    eval_interval = env_json["EvalIntervalMilSteps"] * 1000000
    bandit = train_env.name.startswith("Bimodal1DEnv")
    rng = np.random.RandomState(seed)                     # ReplayBuffer(buffer_size, random_seed)
    store = []
    out = dict(train_rewards_per_episode=[], train_steps_per_episode=[], train_cum_steps=[], eval_rewards_per_episode=[],
               eval_steps_per_episode=[], timesteps_at_eval=[], step_rewards=[], train_episodes=0, actions=[])

    def evaluate():
        rs, ns = [], []
        for _ in range(env_json["EvalEpisodes"]):
            obs = test_env.reset()
            ep_r, done, n = 0., False, 0
            a = agent.predict_action(np.asarray(obs, np.float32)[None])[0].astype(np.float32)
            while not (done or n == test_env.EPISODE_STEPS_LIMIT):
                obs, r, done, _ = test_env.step(a)
                ep_r += r
                if not done:
                    a = agent.predict_action(np.asarray(obs, np.float32)[None])[0].astype(np.float32)
                n += 1
            rs.append(float(ep_r))
            ns.append(n)
        out["eval_rewards_per_episode"].append(rs)
        out["eval_steps_per_episode"].append(ns)

    evaluate()
    out["timesteps_at_eval"].append(0)
    t = 0
    while t < total:
        out["train_episodes"] += 1
        obs = train_env.reset()
        ep_r, done, n = 0., False, 0
        a = agent.sample_action(np.asarray(obs, np.float32)[None], draws("act", t))[0].astype(np.float32)
        while not (done or n == train_env.EPISODE_STEPS_LIMIT or t == total):
            n += 1
            t += 1
            out["actions"].append(float(a[0]))
            obs_n, r, done, _ = train_env.step(a)
            ep_r += r
            out["step_rewards"].append(float(r))
            truncated = (not bandit) and done and n == train_env.EPISODE_STEPS_LIMIT
            if not truncated:
                store.append((np.asarray(obs, np.float32), a, np.float32(r), np.asarray(obs_n, np.float32),
                              np.float32(0.0 if done else gamma)))
            if len(store) > max(warmup_steps, batch_size):
                idx = onp.sample_n_k(rng, len(store), batch_size)
                s, a_b, r_b, s2, g_b = (np.array(x) for x in zip(*[store[i] for i in idx]))
                agent.update(s, a_b, s2, r_b, g_b, draws("upd", t - 1))
            if not done:
                a = agent.sample_action(np.asarray(obs_n, np.float32)[None], draws("act", t))[0].astype(np.float32)
            obs = obs_n
            if t % eval_interval == 0:
                out["timesteps_at_eval"].append(t)
                evaluate()
        if done or n == train_env.EPISODE_STEPS_LIMIT:
            out["train_rewards_per_episode"].append(float(ep_r))
            out["train_cum_steps"].append(t)
            out["train_steps_per_episode"].append(n)
    return out
