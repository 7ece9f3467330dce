"""GPU tests of the drop-in mirrors of the reference's agents/network critic entry points
(rlcontrol_b200/networks.py): same method names / positional signatures / shapes as the reference
classes, results equal to the oracle restatement of the reference arithmetic."""
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def _cfg(eng, **kw):
    base = dict(state_dim=4, state_min=[-2.0] * 4, state_max=[2.0] * 4, action_dim=2, action_min=[-1.0, -1.0],
                action_max=[1.0, 1.0], tau=0.01, norm_type="input_norm", random_seed=3, engine=eng)
    base.update(kw)
    return SimpleNamespace(**base)


def test_qtopt_network_entry_points(eng):
    from rlcontrol_b200.networks import QTOPTNetwork
    cfg = _cfg(eng, qnet_lr=1e-3, qnet_l1_dim=48, qnet_l2_dim=40, num_iter=3, num_samples=64, top_m=6, num_modal=2)
    net = QTOPTNetwork(None, None, cfg)
    rng = np.random.RandomState(0)
    R = 50
    s = rng.randn(R, 4) * 1.5                       # float64 in, some rows get clipped
    a = rng.uniform(-1, 1, (R, 2))
    p = net.get_weights()
    q = net.predict_q(s, a, True)
    assert q.shape == (R, 1) and q.dtype == np.float32
    ref = onp.tmid_forward(s, a, *p, smin=np.array(cfg.state_min), smax=np.array(cfg.state_max), dtype=np.float64)
    np.testing.assert_allclose(q[:, 0], ref, rtol=1e-4, atol=1e-5)
    # target net is an independent init until init_target_network()
    assert np.abs(net.predict_q_target(s, a, True) - q).max() > 1e-6
    net.init_target_network()
    np.testing.assert_array_equal(net.predict_q_target(s, a, True), q)
    # train == one TF-Adam step on the MSE (critic_network.py:54-55)
    y = rng.randn(R, 1)
    out = net.train(s, a, y)
    assert len(out) == 2 and out[1] is None
    np.testing.assert_allclose(out[0], q, rtol=1e-5, atol=1e-6)          # q before the step, like sess.run([outputs, optimize])
    loss, grads = onp.tmid_mse_grads(s, a, y[:, 0], p, np.array(cfg.state_min), np.array(cfg.state_max))
    th = np.concatenate([x.ravel() for x in p]).astype(np.float64)
    th2, _, _ = onp.adam_step_tf(th, np.concatenate([g.ravel() for g in grads]), np.zeros_like(th), np.zeros_like(th), 1, 1e-3)
    np.testing.assert_allclose(np.concatenate([x.ravel() for x in net.get_weights()]), th2, rtol=0, atol=2e-6)
    # soft target update
    tgt0 = np.concatenate([x.ravel() for x in net.get_weights(target=True)])
    net.update_target_network()
    tgt1 = np.concatenate([x.ravel() for x in net.get_weights(target=True)])
    np.testing.assert_allclose(tgt1, tgt0 + 0.01 * (th2 - tgt0), rtol=0, atol=2e-6)
    # CEM: same draws through the oracle reproduce the device result
    B = 7
    sb = rng.randn(B, 4)
    net.rng = np.random.RandomState(11)
    gmms = net.iterate_cem_multidim(sb)
    r2 = np.random.RandomState(11)
    u0 = r2.uniform(size=(B, 64, 2)).astype(np.float32)
    noise = r2.randn(2, B, 64, 2).astype(np.float32)
    cu = r2.uniform(size=(2, B, 64)).astype(np.float32)
    pw = net.get_weights()
    qf = lambda st, ac: onp.tmid_eval(st, ac, pw, np.array(cfg.state_min), np.array(cfg.state_max))
    W, Mu, Cv, _ = onp.cem_iterate(qf, sb.astype(np.float32), u0, noise, cu, 6, 2, -np.ones(2), np.ones(2))
    got_mu = np.array([g.means_ for g in gmms])
    ok = np.isclose(got_mu, Mu, rtol=1e-3, atol=1e-4).all(axis=(1, 2))
    assert ok.mean() >= 0.7                       # near-ties in top-m may diverge a state (see test_cem_matches_oracle)
    act = net.predict_action(sb)
    assert act.shape == (B, 2) and np.all(np.abs(act) <= 2.0)
    samp, mean, wmv = net.sample_action(sb)
    assert samp.shape == (B, 1, 2) and mean.shape == (B, 2) and len(wmv) == B and wmv[0][1].shape == (2, 2)
    assert net.getQFunction(sb[0])(np.array([0.1, -0.2])).shape == (1, 1)


def test_critic_network_and_actor_expert_entry_points(eng):
    from rlcontrol_b200.networks import ActorExpertCritic, CriticNetwork
    cfg = _cfg(eng, critic_lr=1e-3, critic_l1_dim=32, critic_l2_dim=24, shared_l1_dim=32, expert_l2_dim=24,
               expert_lr=1e-3, better_q_gd_alpha=0.05, better_q_gd_max_steps=10, better_q_gd_stop=1e-3, norm_type="none")
    rng = np.random.RandomState(1)
    s, a = rng.randn(20, 4) * 3, rng.uniform(-1, 1, (20, 2))
    cn = CriticNetwork(None, None, cfg)
    p = cn.get_weights()
    np.testing.assert_allclose(cn.predict(s, a, False)[:, 0], onp.tmid_forward(s, a, *p, dtype=np.float64), rtol=1e-4, atol=1e-5)
    g = cn.action_gradients(s, a, False)
    assert isinstance(g, list) and g[0].shape == (20, 2)
    np.testing.assert_allclose(g[0], onp.tmid_dq_da(s, a, p), rtol=1e-3, atol=1e-6)     # norm_type none: no clip
    ae = ActorExpertCritic(None, None, cfg)
    ae.set_weights(*[x * (30.0 if i >= 4 else 1.0) for i, x in enumerate(ae.get_weights())])   # steeper Q
    p = ae.get_weights()
    a2 = ae.q_gradient_ascent(s, a.copy(), True, is_better_q_gd=True)
    ref = onp.q_gradient_ascent(lambda st, ac: onp.tmid_dq_da(st, ac, p), s, a, 0.05, -1.0, 1.0, 10, 1e-3)
    np.testing.assert_allclose(a2, ref, rtol=1e-3, atol=1e-4)
    with pytest.raises(AssertionError):
        ae.q_gradient_ascent(s, a, True)
    # ActorExpert.py:162-181: sample -> stack -> predict_q -> per-state argsort()[::-1][:k] -> gather
    B, N, k = 6, 40, 5
    sb, ab = rng.randn(B, 4), rng.uniform(-1, 1, (B, N, 2))
    q, idx, el = ae.select_elites(sb, ab, k)
    qs = ae.predict_q(np.repeat(sb, N, axis=0), ab.reshape(B * N, 2), True).reshape(B, N)
    np.testing.assert_allclose(q, qs, rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(idx, onp.topk_desc(q, k))
    np.testing.assert_array_equal(el, onp.gather_elites(ab.astype(np.float32), idx))
    with pytest.raises(NotImplementedError):
        CriticNetwork(None, None, _cfg(eng, critic_lr=1e-3, critic_l1_dim=8, critic_l2_dim=8, norm_type="batch"))
    with pytest.raises(ValueError):
        cn.predict(s, a[:5], False)


def test_soft_q_network_mirror(eng):
    import torch
    from rlcontrol_b200.networks import SoftQNetwork

    class RefQ(torch.nn.Module):                       # forwardkl_network.py:250-268
        def __init__(self, S, A, l1, l2):
            super().__init__()
            self.linear1, self.linear2, self.linear3 = torch.nn.Linear(S + A, l1), torch.nn.Linear(l1, l2), torch.nn.Linear(l2, 1)

        def forward(self, state, action):
            x = torch.cat([state, action], 1)
            return self.linear3(torch.relu(self.linear2(torch.relu(self.linear1(x)))))

    torch.manual_seed(0)
    ref = RefQ(3, 1, 64, 48)
    net = SoftQNetwork(3, 1, 64, 48, engine=eng).load_from_torch(ref)
    s, a = torch.randn(30, 3), torch.rand(30, 1) * 2 - 1
    q = net(s, a)
    assert tuple(q.shape) == (30, 1)
    np.testing.assert_allclose(q.cpu().numpy(), ref(s, a).detach().numpy(), rtol=1e-4, atol=1e-5)
    grid = torch.linspace(-1, 1, 62).reshape(-1, 1)
    qg = net.eval_grid(s, grid)
    stacked = ref(s.unsqueeze(1).repeat(1, 62, 1).reshape(-1, 3), grid.repeat(30, 1, 1).reshape(-1, 1)).reshape(30, 62)
    np.testing.assert_allclose(qg.cpu().numpy(), stacked.detach().numpy(), rtol=1e-4, atol=1e-5)


def test_critic_tf_checkpoint_round_trip(eng, tmp_path):
    """save_tf_checkpoint / load_tf_checkpoint (tf_bundle.py): the restored critic evaluates identically, and the
    container holds the reference's variable names (agents/SoftActorCritic.py:37-49 restores ``main/qf/*``)."""
    from rlcontrol_b200.networks import CriticNetwork
    from rlcontrol_b200 import tf_bundle
    cfg = _cfg(eng, critic_lr=1e-3, critic_l1_dim=40, critic_l2_dim=24, state_dim=1, state_min=[-1.0], state_max=[1.0], action_dim=1,
               action_min=[-1.0], action_max=[1.0])
    a = CriticNetwork(None, None, cfg)
    cfg.random_seed = 99
    b = CriticNetwork(None, None, cfg)
    rng = np.random.RandomState(1)
    b.set_weights(*[w + rng.randn(*w.shape).astype(np.float32) * 0.1 for w in b.get_weights()])
    s, act = rng.uniform(-1, 1, (20, 1)), rng.uniform(-1, 1, (20, 1))
    assert np.abs(a.predict(s, act, True) - b.predict(s, act, True)).max() > 1e-6
    pre = str(tmp_path / "ckpt")
    a.save_tf_checkpoint(pre)
    names = set(tf_bundle.read_index(pre))
    assert names == {"main/qf/fully_connected%s/%s" % (sfx, k) for sfx in ("", "_1", "_2") for k in ("weights", "biases")}
    b.load_tf_checkpoint(pre)
    np.testing.assert_array_equal(a.predict(s, act, True), b.predict(s, act, True))
