"""Sweep INDEX decoding / config assembly (rlcontrol_b200/sweep.py) against the reference's own function and JSON files
(CPU; the reference-dependent cases skip where /root/reference is absent), and the concurrent runner on the GPU."""
import importlib.util
import os

import numpy as np
import pytest

from rlcontrol_b200 import sweep

REF = "/root/reference"
need_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference not present")
PENDULUM = dict(state_dim=3, state_min=[-1.0, -1.0, -8.0], state_max=[1.0, 1.0, 8.0], action_dim=1, action_min=[-2.0], action_max=[2.0])


def test_sweep_setting_mixed_radix_first_key_fastest():
    sw = {"a": [1, 2, 3], "b": ["x", "y"], "c": [0.1]}
    seen = [tuple(sweep.sweep_setting(sw, i)[0].values()) for i in range(6)]
    assert seen == [(1, "x", 0.1), (2, "x", 0.1), (3, "x", 0.1), (1, "y", 0.1), (2, "y", 0.1), (3, "y", 0.1)]
    assert sweep.sweep_setting(sw, 7)[0] == sweep.sweep_setting(sw, 1)[0] and sweep.sweep_setting(sw, 0)[1] == 6
    assert sweep.index_to_run(13, 6) == (1, 2, 2)              # setting 1, third run, seed 2 (main.py:131-141)
    cfg = sweep.make_config(PENDULUM, {"tau": 0.5, "pi_lr": 1e-3}, {"random_seed": 4}, precision="fp32")
    assert cfg.batch_size == 32 and cfg.gamma == 0.99 and cfg.tau == 0.5 and cfg.random_seed == 4 and cfg.precision == "fp32"


@need_ref
def test_matches_reference_get_sweep_parameters_on_shipped_jsons():
    spec = importlib.util.spec_from_file_location("ref_main_utils", os.path.join(REF, "utils", "main_utils.py"))
    src = open(os.path.join(REF, "utils", "main_utils.py")).read()
    ns = {}
    start = src.index("def get_sweep_parameters")
    exec("from collections import OrderedDict\n" + src[start:], ns)          # the function only (the module imports every agent)
    ref = ns["get_sweep_parameters"]
    for name in ("reverse_kl.json", "forward_kl.json", "qt_opt.json", "ae.json", "sql.json"):
        agent, sweeps = sweep.load_agent_json(os.path.join(REF, "jsonfiles", "agent", name))
        total = ref(sweeps, 0)[1]
        for idx in list(range(0, min(total, 40))) + [total - 1, total, total + 3, 5 * total + 7]:
            mine, t1 = sweep.sweep_setting(sweeps, idx)
            theirs, t2 = ref(sweeps, idx)
            assert t1 == t2 and list(mine.items()) == list(theirs.items())
    agent, sweeps = sweep.load_agent_json(os.path.join(REF, "jsonfiles", "agent", "reverse_kl.json"))
    assert agent == "ReverseKL" and sweep.sweep_setting(sweeps, 0)[1] == 36   # 3 pi_lr x 3 qf_vf_lr x 4 entropy scales


@pytest.mark.gpu
def test_sweep_runner_agents_are_independent_and_overlapped(eng):
    """8 INDEX runs of the shipped reverse_kl sweep shape on one GPU: each agent gets its INDEX's setting and seed, and
    updating them together gives exactly what updating each one alone gives."""
    import torch
    from rlcontrol_b200 import kl_networks
    from rlcontrol_b200.engine import Engine
    sweeps = {"norm_type": ["input_norm"], "actor_l1_dim": [64], "actor_l2_dim": [48], "critic_l1_dim": [64], "critic_l2_dim": [48],
              "pi_lr": [1e-3, 1e-4], "qf_vf_lr": [1e-2, 1e-3], "use_true_q": ["False"], "entropy_scale": [1, 0.1], "l_param": [6],
              "N_param": [64], "optim_type": ["intg"], "q_update_type": ["non_sac"]}
    idx = list(range(8)) + [9]
    run = sweep.SweepRunner("ReverseKL", sweeps, PENDULUM, idx)
    assert [s["setting"] for s in run.settings] == [0, 1, 2, 3, 4, 5, 6, 7, 1] and run.settings[-1]["run"] == 1
    assert run.agents[1].learning_rate == [1e-4, 1e-2] and run.agents[4].entropy_scale == 0.1
    rng = np.random.RandomState(0)
    B = 32
    batches = [(rng.randn(B, 3), rng.uniform(-2, 2, (B, 1)), rng.randn(B, 3), rng.randn(B), np.full(B, 0.99)) for _ in idx]
    eps = [rng.randn(B, 1).astype(np.float32) for _ in idx]
    losses = run.update_all([b + (e,) for b, e in zip(batches, eps)])
    for i in (0, 4, 8):                                          # the same agent, alone
        params, total = sweep.sweep_setting(sweeps, idx[i])
        seed = sweep.index_to_run(idx[i], total)[2]
        torch.manual_seed(seed)
        solo = kl_networks.ReverseKLNetwork(None, None, sweep.make_config(PENDULUM, params, dict(random_seed=seed), engine=Engine(0)))
        solo.update_network(*batches[i], eps=eps[i])
        solo.update_target_network()
        np.testing.assert_allclose(solo.last_losses, losses[i], rtol=1e-6)       # the loss scalars are atomic sums
        a, b = solo.export_parameters(), run.agents[i].export_parameters()
        for k in a:
            for x, y in zip(a[k], b[k]):
                np.testing.assert_array_equal(x, y)
    with pytest.raises(NotImplementedError):
        sweep.SweepRunner("DDPG", sweeps, PENDULUM, [0])


def test_main_device_rank_partition_covers_every_index_once():
    """torchrun ranks of rlcontrol_b200.main_device split range(START, STOP, STEP) without overlap (replicas only)."""
    from rlcontrol_b200.main_device import rank_indices
    for indices in [(0, 1, 64), (3, 2, 50), (0, 1, 5), (10, 7, 11)]:
        rng_of = lambda t: list(range(t[0], t[2], t[1]))            # main.py:113: range(indices[0], indices[2], indices[1])
        want = rng_of(indices)
        for world in (1, 2, 4, 8):
            got = [rng_of(rank_indices(indices, r, world)) for r in range(world)]
            flat = sorted(i for g in got for i in g)
            assert flat == want
            assert all(g == want[r::world] for r, g in enumerate(got))
