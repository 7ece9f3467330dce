"""GPU tests of the device-resident experiment loop (csrc/envloop.cu, rlcontrol_b200/device_loop.py): the environment
kernels against the oracle environments (Bimodal: the fixture recorded from the reference classes), and whole runs --
environment steps, replay, the reference's minibatch index stream, updates, evaluation sessions -- against
oracle_env.run_experiment driving oracle_kl.KLAgent on the same draws."""
import ctypes as C
import os
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import oracle_env as oenv
from oracle import oracle_kl as okl

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "bimodal_env.npz")


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _env_buffers(torch, dev, E, S):
    z = lambda *sh, dt: torch.zeros(sh, dtype=dt, device=dev)
    return dict(state=z(E, 2, dt=torch.float64), step=z(E, dt=torch.int32), ret=z(E, dt=torch.float64),
                done=z(E, dt=torch.int32), obs=z(E, S, dt=torch.float32))


@pytest.mark.parametrize("name", list(oenv.BIMODAL))
def test_bimodal_step_kernel_matches_reference_fixture(eng, name):
    import torch
    from rlcontrol_b200 import device_loop as dl
    from rlcontrol_b200._lib import check
    g = np.load(GOLD)
    a = g["actions"]
    E = len(a)
    spec = dl.EnvSpec({"environment": name, "TotalMilSteps": 0.001, "EvalIntervalMilSteps": 0.0001, "EvalEpisodes": 1})
    b = _env_buffers(torch, eng.device, E, 1)
    feed = torch.zeros((1, 2), dtype=torch.float64, device=eng.device)
    act = torch.as_tensor(a.reshape(E, 1)).to(eng.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(eng.lib.rlc_env_reset(eng.h, C.byref(spec.desc), E, _ptr(feed), 1, None, _ptr(b["state"]), _ptr(b["step"]),
                                _ptr(b["ret"]), _ptr(b["done"]), _ptr(b["obs"]), st))
    assert float(b["obs"].abs().max()) == 0.0
    for _ in range(2):                                            # the second call must be a no-op: episodes are over
        check(eng.lib.rlc_env_step_eval(eng.h, C.byref(spec.desc), E, _ptr(b["state"]), _ptr(b["step"]), _ptr(b["ret"]),
                                        _ptr(b["done"]), _ptr(b["obs"]), _ptr(act), st))
    # float32 argument like the reference, double exp: libm vs CUDA exp differ by an ulp at most
    np.testing.assert_allclose(b["ret"].cpu().numpy(), g[name + "_reward"], rtol=4e-16, atol=1e-300)
    np.testing.assert_array_equal(b["obs"].cpu().numpy().reshape(-1), g[name + "_next"].astype(np.float32))
    assert b["done"].cpu().numpy().tolist() == [1] * E and b["step"].cpu().numpy().tolist() == [1] * E


def test_pendulum_step_kernel_matches_oracle_rollouts(eng):
    import torch
    from rlcontrol_b200 import device_loop as dl
    from rlcontrol_b200._lib import check
    E, T = 16, 60
    spec = dl.EnvSpec({"environment": "Pendulum-v0", "TotalMilSteps": 0.001, "EvalIntervalMilSteps": 0.0005,
                       "EvalEpisodes": E, "EpisodeSteps": 50})
    rng = np.random.RandomState(0)
    resets = spec.reset_states(dl.gym_np_random(9), E + 3)
    actions = rng.uniform(-2.5, 2.5, (T, E, 1)).astype(np.float32)
    b = _env_buffers(torch, eng.device, E, 3)
    cursor = torch.tensor([3], dtype=torch.int64, device=eng.device)
    feed = torch.as_tensor(resets).to(eng.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(eng.lib.rlc_env_reset(eng.h, C.byref(spec.desc), E, _ptr(feed), feed.shape[0], _ptr(cursor), _ptr(b["state"]),
                                _ptr(b["step"]), _ptr(b["ret"]), _ptr(b["done"]), _ptr(b["obs"]), st))
    assert int(cursor[0]) == 3 + E
    envs = []
    for e in range(E):
        env = oenv.PendulumV0(50)
        env.state = resets[3 + e].copy()
        env._elapsed = 0
        envs.append(env)
    np.testing.assert_allclose(b["obs"].cpu().numpy(), np.array([env._obs() for env in envs], np.float32), rtol=0, atol=1e-7)
    ret = np.zeros(E)
    for t in range(T):
        act = torch.as_tensor(actions[t]).to(eng.device)
        check(eng.lib.rlc_env_step_eval(eng.h, C.byref(spec.desc), E, _ptr(b["state"]), _ptr(b["step"]), _ptr(b["ret"]),
                                        _ptr(b["done"]), _ptr(b["obs"]), _ptr(act), st))
        if t < 50:
            obs = []
            for e, env in enumerate(envs):
                o, r, done, _ = env.step(actions[t, e])
                ret[e] += r
                obs.append(o)
                assert done == (t == 49)
            np.testing.assert_allclose(b["obs"].cpu().numpy(), np.array(obs, np.float32), rtol=0, atol=2e-6)
            np.testing.assert_allclose(b["state"].cpu().numpy(), np.array([env.state for env in envs]), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(b["ret"].cpu().numpy(), ret, rtol=1e-12)       # frozen after the 50-step limit
    assert b["step"].cpu().numpy().tolist() == [50] * E and b["done"].cpu().numpy().tolist() == [1] * E


def _config(spec, seed, kind="rkl", batch=8, engine=None):
    p = spec.env_params()
    return SimpleNamespace(pi_lr=1e-3, qf_vf_lr=1e-2, tau=0.01, norm_type="none", optim_type="intg", q_update_type="non_sac",
                           use_true_q="False", sample_for_eval="False", random_seed=seed, entropy_scale=0.1,
                           actor_l1_dim=32, actor_l2_dim=24, critic_l1_dim=40, critic_l2_dim=32, N_param=18, l_param=6,
                           batch_size=batch, gamma=0.99, warmup_steps=0, buffer_size=1e6, precision="fp32", engine=engine, **p)


def _oracle_twin(net, kind):
    p = net.export_parameters()
    return okl.KLAgent(kind, p["q"], p["v"], p["tv"], p["pi"], net.intgrl_actions.cpu().numpy(), net.intgrl_weights.cpu().numpy(),
                       net.action_scale, net.entropy_scale, net.learning_rate[0], net.learning_rate[1], net.tau,
                       optim_type=net.optim_type, q_update_type=net.q_update_type)


def _draws_like_device(torch, seed, A, B, K, total):
    """The N(0,1) feeds exactly as DeviceExperiment draws them (per-run generator; start draw, then per chunk
    [n,A] and [n,B,A])."""
    gen = torch.Generator().manual_seed(seed)
    start = torch.randn(1, A, generator=gen).numpy()
    act, upd, t = [], [], 0
    while t < total:
        n = min(K, total - t)
        act.append(torch.randn(n, A, generator=gen).numpy())
        upd.append(torch.randn(n, B, A, generator=gen).numpy())
        t += n
    act, upd = np.concatenate(act), np.concatenate(upd)
    return lambda kind, t: (start if t == 0 else act[t - 1:t]) if kind == "act" else upd[t]


@pytest.mark.parametrize("env_name,ep,kind,fused", [("Pendulum-v0", 12, "rkl", True), ("Bimodal1DEnv_uneq_var1", -1, "rkl", True),
                                                   ("Pendulum-v0", 9, "fkl", True), ("Pendulum-v0", 12, "rkl", False)])
def test_device_experiment_matches_oracle_loop(eng, env_name, ep, kind, fused):
    import torch
    from rlcontrol_b200 import device_loop as dl
    from rlcontrol_b200 import kl_networks
    env_json = {"environment": env_name, "TotalMilSteps": 70e-6, "EpisodeSteps": ep, "EvalIntervalMilSteps": 20e-6,
                "EvalEpisodes": 3}
    spec = dl.EnvSpec(env_json)
    seed, B, K = 3, 8, 16                                     # 70 steps = 4 full chunks of 16 + 6: the feed ring wraps
    cfg = _config(spec, seed, kind, B, engine=eng)
    cfg.fused_small_batch = fused                             # False: the generic multi-kernel update and acting path
    torch.manual_seed(seed)
    net = (kl_networks.ReverseKLNetwork if kind == "rkl" else kl_networks.ForwardKLNetwork)(None, None, cfg)
    assert net._small_ok(B) == fused
    twin = _oracle_twin(net, kind)
    want = oenv.run_experiment(twin, env_json, seed, B, 0.99, _draws_like_device(torch, seed, 1, B, K, 70))
    exp = dl.DeviceExperiment(net, env_json, cfg, chunk_steps=K, steps_per_graph=1 if kind == "fkl" else 4)
    before = net.export_parameters()
    got = exp.run()
    (ep_r, ev_r, ep_s, ev_s, t_ev, _, _, n_ep, cum) = got
    assert t_ev == want["timesteps_at_eval"] == [0, 20, 40, 60]
    assert ep_s == want["train_steps_per_episode"] and cum == want["train_cum_steps"] and n_ep == want["train_episodes"]
    assert ev_s == want["eval_steps_per_episode"]
    # fp32 networks on the device vs the float64 oracle, ~60 chained updates: the trajectories stay together to ~1e-4
    np.testing.assert_allclose(ep_r, want["train_rewards_per_episode"], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(ev_r, want["eval_rewards_per_episode"], rtol=2e-3, atol=2e-3)
    after = net.export_parameters()
    for k in ("q", "v", "tv", "pi"):
        ref = getattr(twin, k)
        for a, b0, r in zip(after[k], before[k], ref):
            move = np.abs(np.asarray(r) - b0).max()
            np.testing.assert_allclose(a, r, rtol=0, atol=2e-2 * move + 1e-6)
    d = exp.run_data(env_json)
    assert d["random_seed"] == seed and d["total_timesteps"] == 70 and d["episodes_per_eval"] == 3
    assert d["eval_episode_rewards"].shape == (4, 3) and d["total_train_episodes"] == n_ep


def test_interleaved_runs_equal_solo_runs(eng):
    import torch
    from rlcontrol_b200 import device_loop as dl
    from rlcontrol_b200 import kl_networks
    env_json = {"environment": "Pendulum-v0", "TotalMilSteps": 90e-6, "EpisodeSteps": 15, "EvalIntervalMilSteps": 30e-6,
                "EvalEpisodes": 2}
    spec = dl.EnvSpec(env_json)

    def make(seed):
        cfg = _config(spec, seed, engine=None)
        torch.manual_seed(seed)
        return dl.DeviceExperiment(kl_networks.ReverseKLNetwork(None, None, cfg), env_json, cfg, chunk_steps=20, steps_per_graph=5)
    solo = [make(s).run() for s in (0, 1, 2)]
    exps = [make(s) for s in (0, 1, 2)]
    together = dl.run_interleaved(exps)
    for a, b in zip(solo, together):
        for x, y in zip(a[:5], b[:5]):
            assert x == y                                      # bit-identical: runs share nothing
    assert solo[0][0] != solo[1][0]


def test_main_device_writes_the_reference_pickle(eng, tmp_path):
    """`python -m rlcontrol_b200.main_device` with main.py's arguments: INDEX decoding, grouping of runs on the GPU and
    the pickled dictionary of main.py:80-95,188-203."""
    import json
    import pickle
    from rlcontrol_b200 import main_device
    env_json = {"environment": "Pendulum-v0", "TotalMilSteps": 60e-6, "EpisodeSteps": 12, "EvalIntervalMilSteps": 20e-6,
                "EvalEpisodes": 2}
    agent_json = {"agent": "ReverseKL", "sweeps": {
        "norm_type": ["input_norm"], "exploration_policy": ["none"], "actor_l1_dim": [24], "actor_l2_dim": [24],
        "critic_l1_dim": [32], "critic_l2_dim": [24], "pi_lr": [1e-3], "qf_vf_lr": [1e-2, 1e-3], "sample_for_eval": ["False"],
        "use_true_q": ["False"], "entropy_scale": [0.1], "l_param": [6], "N_param": [16], "optim_type": ["intg"],
        "q_update_type": ["non_sac"]}}
    (tmp_path / "Pendulum-v0.json").write_text(json.dumps(env_json))
    (tmp_path / "reverse_kl.json").write_text(json.dumps(agent_json))
    main_device.main(["--env_json", str(tmp_path / "Pendulum-v0.json"), "--agent_json", str(tmp_path / "reverse_kl.json"),
                      "--indices", "0", "1", "4", "--save_dir", str(tmp_path / "results"), "--runs_per_gpu", "3"])
    path = tmp_path / "results" / "Pendulum-v0_reverse_klresults" / "data_0_1_4.pkl"
    data = pickle.loads(path.read_bytes())
    assert data["experiment"]["agent"]["agent_name"] == "ReverseKL"
    assert data["experiment"]["environment"] == {"env_name": "Pendulum-v0", "total_timesteps": 60.0, "steps_per_episode": 12,
                                                 "eval_interval_timesteps": 20.0, "eval_episodes": 2}
    assert sorted(data["experiment_data"]) == [0, 1]                       # two settings (qf_vf_lr), two runs each
    for setting, lr in ((0, 1e-2), (1, 1e-3)):
        d = data["experiment_data"][setting]
        assert d["agent_params"]["qf_vf_lr"] == lr and len(d["runs"]) == 2
        assert [r["random_seed"] for r in d["runs"]] == [0, 1]             # RANDOM_SEED = RUN_NUM (main.py:131-141)
        for r in d["runs"]:
            assert r["eval_episode_rewards"].shape == (4, 2) and r["timesteps_at_eval"].tolist() == [0, 20, 40, 60]
            assert r["train_episode_steps"].tolist() == [12] * 5 and r["total_train_episodes"] == 5
            assert np.all(r["train_episode_rewards"] < 0) and r["episodes_per_eval"] == 2
    # same seed, different learning rate: the first episodes coincide until learning starts to matter
    a, b = data["experiment_data"][0]["runs"][0], data["experiment_data"][1]["runs"][0]
    np.testing.assert_allclose(a["eval_episode_rewards"][0], b["eval_episode_rewards"][0], rtol=1e-12)
    assert not np.allclose(a["eval_episode_rewards"][-1], b["eval_episode_rewards"][-1], rtol=1e-9)
