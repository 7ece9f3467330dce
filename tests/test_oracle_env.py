"""CPU tests of the environment / experiment-loop oracle (oracle/oracle_env.py) and of the host side of the
device-resident loop (rlcontrol_b200/device_loop.py): the Bimodal bandits against the fixture recorded from the
reference classes, gym's Pendulum-v0 restatement against known answers (gym itself is absent: parity unpinned), the
reference's index-sampling stream, and the loop's bookkeeping on a short bandit run."""
import os

import numpy as np
import pytest

from oracle import oracle_env as oenv
from oracle import oracle_kl as okl
from oracle import oracle_np as onp
from rlcontrol_b200 import device_loop as dl

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bimodal_env.npz")


def test_bimodal_envs_match_reference_fixture():
    g = np.load(GOLD)
    for name in oenv.BIMODAL:
        env = oenv.Bimodal1D(name)
        for a, r, s2 in zip(g["actions"], g[name + "_reward"], g[name + "_next"]):
            env.reset()
            obs, rew, done, _ = env.step(np.array([a], np.float32))
            assert done is True
            assert rew == r                                  # same float32 argument, same double exp
            assert obs[0] == s2
        spec = dl.EnvSpec({"environment": name, "TotalMilSteps": 0.001, "EvalIntervalMilSteps": 0.0001, "EvalEpisodes": 1})
        np.testing.assert_array_equal([spec.state_min[0], spec.state_max[0], spec.action_min[0], spec.action_max[0]],
                                      g[name + "_bounds"])
        assert tuple(spec.desc.p[:6]) == oenv.BIMODAL[name] == dl.BIMODAL[name] and spec.episode_limit == 1


def test_pendulum_known_answers():
    env = oenv.PendulumV0()
    env.set_random_seed(3)
    obs = env.reset()
    th, thdot = env.state
    assert -np.pi <= th <= np.pi and -1 <= thdot <= 1
    np.testing.assert_allclose(obs, [np.cos(th), np.sin(th), thdot])
    # upright and at rest with no torque: stays (sin(pi) is 1.2e-16, not 0), cost 0
    env.state = np.array([0.0, 0.0])
    obs, r, done, _ = env.step(np.array([0.0], np.float32))
    np.testing.assert_allclose(obs, [1, 0, 0], atol=1e-14)
    assert abs(r) < 1e-30 and not done
    # hanging down, full torque beyond the limit is clipped to 2: cost = pi^2 + .001*4
    env.state = np.array([np.pi, 0.0])
    obs, r, _, _ = env.step(np.array([5.0], np.float32))
    np.testing.assert_allclose(r, -(np.pi ** 2 + 0.004), rtol=1e-14)
    np.testing.assert_allclose(env.state[1], (-15.0 * np.sin(2 * np.pi) + 6.0) * 0.05, rtol=1e-12)
    # angle normalisation: 3*pi/2 is -pi/2
    env.state = np.array([1.5 * np.pi, 1.0])
    _, r, _, _ = env.step(np.array([0.0], np.float32))
    np.testing.assert_allclose(r, -((np.pi / 2) ** 2 + 0.1), rtol=1e-13)
    # speed clip and the TimeLimit
    env.state = np.array([np.pi / 2, 7.9])
    env.step(np.array([2.0], np.float32))
    assert env.state[1] == 8.0
    env.reset()
    flags = [env.step(np.array([0.3], np.float32))[2] for _ in range(200)]
    assert flags[:199] == [False] * 199 and flags[199] is True or flags[199] == True  # noqa: E712


def test_gym_seeding_restatements_agree_and_are_deterministic():
    for seed in (0, 1, 7, 12345):
        a, b = oenv.gym_np_random(seed), dl.gym_np_random(seed)
        np.testing.assert_array_equal(a.uniform(size=5), b.uniform(size=5))
    assert oenv.gym_np_random(1).uniform() != oenv.gym_np_random(2).uniform()
    spec = dl.EnvSpec({"environment": "Pendulum-v0", "TotalMilSteps": 0.001, "EvalIntervalMilSteps": 0.0005,
                       "EvalEpisodes": 2, "EpisodeSteps": -1})
    env = oenv.PendulumV0()
    env.set_random_seed(5)
    want = []
    for _ in range(4):
        env.reset()
        want.append(env.state.copy())
    np.testing.assert_array_equal(spec.reset_states(dl.gym_np_random(5), 4), np.array(want))
    assert spec.episode_limit == 200 and spec.total_steps == 1000 and spec.eval_interval == 500


def test_sample_n_k_fast_path_is_the_reference_stream():
    for n, k, seed in [(40, 8, 0), (100, 32, 1), (97, 32, 2), (33, 32, 3), (5000, 32, 4), (20, 8, 5), (25, 8, 6)]:
        r1, r2 = np.random.RandomState(seed), np.random.RandomState(seed)
        for _ in range(50):
            a, b = onp.sample_n_k(r1, n, k), dl.sample_n_k(r2, n, k)
            np.testing.assert_array_equal(a, b)
            assert len(set(b.tolist())) == k
        assert r1.randint(1 << 30) == r2.randint(1 << 30)   # both consumed the same number of draws


def _tiny_agent(S, A, scale, rng):
    u = lambda k, *sh: rng.uniform(-k, k, sh)
    mlp = lambda i, o: [u(.5, 8, i), u(.1, 8), u(.3, 8, 8), u(.1, 8), u(.1, o, 8), u(.1, o)]
    pi = mlp(S, A)[:4] + [u(.1, A, 8), u(.1, A), u(.1, A, 8), u(.1, A)]
    v = mlp(S, 1)
    grid_a, grid_w = onp.intg_grid_1d(10, scale)
    return okl.KLAgent("rkl", mlp(S + A, 1), v, [p.copy() for p in v], pi, grid_a.reshape(-1, 1), grid_w, scale, 0.1,
                       1e-3, 1e-2, 0.01)


@pytest.mark.parametrize("env_name,ep", [("Bimodal1DEnv_eq_var1", -1), ("Pendulum-v0", 7)])
def test_oracle_experiment_loop_bookkeeping(env_name, ep):
    env_json = {"environment": env_name, "TotalMilSteps": 30e-6, "EpisodeSteps": ep, "EvalIntervalMilSteps": 10e-6,
                "EvalEpisodes": 2}
    rng = np.random.RandomState(0)
    S = 1 if env_name.startswith("Bimodal") else 3
    agent = _tiny_agent(S, 1, 2.0, rng)
    eps = np.random.RandomState(1).randn(64, 8, 1)
    calls = []

    def draws(kind, t):
        calls.append((kind, t))
        return eps[t, :1] if kind == "act" else eps[t + 31, :4]
    out = oenv.run_experiment(agent, env_json, seed=2, batch_size=4, gamma=0.99, draws=draws)
    assert len(out["step_rewards"]) == 30 and out["timesteps_at_eval"] == [0, 10, 20, 30]
    assert len(out["eval_rewards_per_episode"]) == 4 and all(len(r) == 2 for r in out["eval_rewards_per_episode"])
    limit = 1 if ep == -1 else ep
    n_ep = 30 // limit
    assert out["train_steps_per_episode"] == [limit] * n_ep
    assert out["train_episodes"] == n_ep + (1 if 30 % limit else 0)
    np.testing.assert_allclose(sum(out["train_rewards_per_episode"]), sum(out["step_rewards"][:n_ep * limit]), rtol=1e-12)
    # learning starts once the buffer holds MORE than batch_size rows (base_agent.py:64-66); steps cut by the episode
    # limit are not stored (experiment.py:127-134), the bandits' single step is
    stored_per_ep = limit if env_name.startswith("Bimodal") else limit - 1
    upd = [t for k, t in calls if k == "upd"]
    first = next(t for t in range(30) if (t + 1) // limit * stored_per_ep + min((t + 1) % limit, stored_per_ep) > 4)
    assert upd == list(range(first, 30))
    assert [t for k, t in calls if k == "act"] == list(range(30)) + ([30] if 30 % limit else [])


@pytest.mark.parametrize("k,n0,steps,seed", [(32, 20, 400, 0), (8, 25, 300, 1), (32, 97, 50, 2), (4, 13, 200, 3), (32, 100000, 64, 4)])
def test_chunked_c_sampler_is_numpy_stream_bit_for_bit(k, n0, steps, seed):
    """csrc/hostrng.c (MT19937 + masked rejection + the duplicate-replacement loop) against numpy's RandomState driven
    step by step by the oracle's sample_n_k: same indices, same generator state afterwards -- through the small-n
    permutation branch (left to numpy), populations just above 3k (many collisions, second-half refills) and steps
    without a minibatch."""
    assert dl._host_lib() is not None, "librlc_host.so is not built (make -C rlcontrol_b200/csrc)"
    sizes = np.array([0 if (i % 7 == 3 or n0 + i <= k) else n0 + i for i in range(steps)], np.int64)
    sizes[steps // 2:] = np.minimum(sizes[steps // 2:], n0 + steps // 2 + 5)        # the ring is full: the size stops growing
    r_ref, r_c, r_py = (np.random.RandomState(seed) for _ in range(3))
    want = np.zeros((steps, k), np.int64)
    for i, n in enumerate(sizes):
        if n:
            want[i] = onp.sample_n_k(r_ref, int(n), k)
    got_c, got_py = np.zeros((steps, k), np.int32), np.zeros((steps, k), np.int32)
    dl.sample_chunk(r_c, sizes, k, got_c)
    dl.sample_chunk(r_py, sizes, k, got_py, use_c=False)
    np.testing.assert_array_equal(got_c, want)
    np.testing.assert_array_equal(got_py, want)
    a, b, c = r_ref.get_state(), r_c.get_state(), r_py.get_state()
    np.testing.assert_array_equal(a[1], b[1])
    assert a[2] == b[2] and a[2] == c[2]
    assert r_ref.randint(1 << 30) == r_c.randint(1 << 30) == r_py.randint(1 << 30)
