"""GPU tests of the ForwardKL / ReverseKL update path: the B-row C-ABI pieces (rlc_mlp_*,
rlc_policy_evaluate, rlc_kl_targets, rlc_policy_head_grad) against the numpy oracle, and the drop-in
``ForwardKLNetwork`` / ``ReverseKLNetwork`` (rlcontrol_b200/kl_networks.py) against fixtures recorded
from the unmodified reference classes (tests/golden/full_*.npz, oracle/make_golden.py)."""
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import oracle_kl as okl
from test_oracle_kl import FULL, FULL_DENSE, load_full, make_agent

pytestmark = pytest.mark.gpu


def _t(eng, x):
    import torch
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(eng.device)


def _rand_mlp(rng, inp, H1, H2, O):
    u = lambda k, *sh: rng.uniform(-k, k, sh).astype(np.float32)
    return [u(1 / np.sqrt(inp), H1, inp), u(0.3, H1), u(1 / np.sqrt(H1), H2, H1), u(0.3, H2), u(0.5, O, H2), u(0.5, O)]


@pytest.mark.parametrize("B,inp,H1,H2,O", [(32, 3, 200, 200, 2), (7, 5, 33, 17, 1), (300, 17, 400, 300, 12), (1, 3, 64, 48, 4)])
def test_mlp_forward_and_grads_match_oracle(eng, B, inp, H1, H2, O):
    import rlcontrol_b200 as rb
    rng = np.random.RandomState(B + O)
    p = _rand_mlp(rng, inp, H1, H2, O)
    x = rng.randn(B, inp).astype(np.float32)
    dout = rng.randn(B, O).astype(np.float32)
    m = rb.Mlp(eng, inp, H1, H2, O).load_torch(*p)
    for a, b in zip(m.export_torch(), p):                       # pack/unpack round trip is exact
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    ref_out, cache = okl.mlp_forward(x, *p)
    ref_g = okl.mlp_grads(cache, dout, p[2], p[4])
    act = m.act_buffer(B)
    out = m.forward(_t(eng, x), act=act)
    np.testing.assert_allclose(out.cpu().numpy(), ref_out, rtol=2e-5, atol=2e-5)
    for use_act in (True, False):                               # kept activations and recomputed forward agree
        g, dx = m.grads(_t(eng, x), _t(eng, dout), act=act if use_act else None, want_dx=True)
        g = g.cpu().numpy()
        o = m.offsets + [g.size]
        shapes = [(inp, H1), (H1,), (H1, H2), (H2,), (H2, O), (O,)]
        for i, (r, sh) in enumerate(zip(ref_g, shapes)):
            mine = g[o[i]:o[i + 1]].reshape(sh)
            r = r.T if r.ndim == 2 else r                       # oracle is [out,in], theta is [in,out]
            np.testing.assert_allclose(mine, r, rtol=1e-4, atol=1e-4 * max(1.0, np.abs(r).max()))
        g1 = (np.asarray(dout, np.float64) @ p[4]) * (cache[3] > 0)
        g1 = (g1 @ p[2]) * (cache[1] > 0)
        np.testing.assert_allclose(dx.cpu().numpy(), g1 @ p[0], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("A", [1, 2, 3])
def test_policy_evaluate_matches_oracle(eng, A):
    rng = np.random.RandomState(A)
    B = 37
    head = rng.randn(B, 2 * A).astype(np.float32)
    head[:, A:] = head[:, A:] * 2 - 1
    head[0, A:] = 5.0                                           # clamped above (LOG_STD_MAX = 2)
    head[1, A:] = -30.0                                         # clamped below
    eps = rng.randn(B, A).astype(np.float32)
    ev = eng.policy_evaluate(_t(eng, head), _t(eng, eps), 2.0)
    mean, ls = head[:, :A], np.clip(head[:, A:], -20, 2)
    act, lp, z, mt = okl.policy_evaluate(mean, ls, eps, 2.0)
    np.testing.assert_allclose(ev["action"].cpu().numpy(), act, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ev["z"].cpu().numpy(), z, rtol=2e-6, atol=2e-6)
    # log(1 - tanh(z)^2 + 1e-6) is evaluated in fp32 by the reference as well (forwardkl_network.py:313): where
    # tanh saturates, 1 - a^2 cancels and one ulp of tanh moves the term by O(1) -- those rows get a loose bound
    # (same for std = e^-20: z = mean + std * eps rounds back to mean in fp32, so the Gaussian term loses eps^2/2)
    sat = (np.abs(np.tanh(z)) > 0.999).any(axis=1) | (ls < -10).any(axis=1)
    mine = ev["logp"].cpu().numpy()
    np.testing.assert_allclose(mine[~sat], lp[~sat], rtol=3e-5, atol=3e-4)
    np.testing.assert_allclose(mine[sat], lp[sat], rtol=0, atol=3.0 * A)
    assert sat.sum() < len(sat) // 2
    np.testing.assert_allclose(ev["mean"].cpu().numpy(), mt, rtol=2e-6, atol=2e-6)
    np.testing.assert_array_equal(ev["mu_raw"].cpu().numpy(), mean)
    np.testing.assert_array_equal(ev["log_std"].cpu().numpy(), ls)
    ev0 = eng.policy_evaluate(_t(eng, head), None, 2.0)         # eps = NULL -> the mean action
    np.testing.assert_allclose(ev0["action"].cpu().numpy(), ev0["mean"].cpu().numpy(), rtol=0, atol=1e-7)


@pytest.mark.parametrize("sac", [False, True])
def test_kl_targets_match_oracle(eng, sac):
    rng = np.random.RandomState(5)
    B, alpha = 45, 0.3
    r, g, vn, qn, lp, v = [rng.randn(B).astype(np.float32) for _ in range(6)]
    y, dv, vl = eng.kl_targets(*[_t(eng, x) for x in (r, g, vn, qn, lp, v)], alpha, sac, b_total=2 * B)
    ry = r.astype(np.float64) + g * vn
    tv = (qn - alpha * lp) if sac else ((r - alpha * lp) + g * vn)
    np.testing.assert_allclose(y.cpu().numpy(), ry, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(dv.cpu().numpy(), 2 * (v - tv) / (2 * B), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(float(vl.cpu()), np.sum((v - tv) ** 2) / (2 * B), rtol=1e-5)


@pytest.mark.parametrize("A,mode", [(1, 0), (1, 1), (1, 2), (3, 0), (3, 1), (3, 2)])
def test_policy_head_grad_matches_oracle(eng, A, mode):
    rng = np.random.RandomState(10 * A + mode)
    B, alpha = 29, 0.2
    head = rng.randn(B, 2 * A).astype(np.float32)
    head[0, A:] = 3.0
    head[1, A:] = -25.0                                         # clamp -> no gradient into log_std
    dm, ds = rng.randn(B, A).astype(np.float32), rng.randn(B, A).astype(np.float32)
    z, lp, qn, v = (rng.randn(B, A).astype(np.float32), rng.randn(B).astype(np.float32),
                    rng.randn(B).astype(np.float32), rng.randn(B).astype(np.float32))
    import torch
    loss = torch.zeros(1, device=eng.device)
    dh = eng.policy_head_grad(_t(eng, head), mode, dmean=_t(eng, dm), dlog_std=_t(eng, ds), z=_t(eng, z), logp=_t(eng, lp),
                              q_new=_t(eng, qn), v=_t(eng, v), entropy_scale=alpha, loss_out=loss).cpu().numpy()
    raw = head[:, A:]
    passm = ((raw >= -20) & (raw <= 2)).astype(np.float64)
    if mode == 0:
        rm, rs = dm, ds * passm
    else:
        c = (qn - v - alpha * lp) if mode == 1 else (qn - v)
        coef = (-c.astype(np.float64) / B)[:, None]
        mu, ls = head[:, :A].astype(np.float64), np.clip(raw, -20, 2).astype(np.float64)
        t = z - mu
        if A == 1:
            rm, rs = coef * t * np.exp(-2 * ls), coef * (t * t * np.exp(-2 * ls) - 1) * passm
        else:
            rm, rs = coef * t * np.exp(-ls), coef * (0.5 * t * t * np.exp(-ls) - 0.5) * passm
        np.testing.assert_allclose(float(loss.cpu()), np.mean(-lp.astype(np.float64) * c), rtol=1e-4, atol=1e-6)
    scale = max(1.0, np.abs(rm).max(), np.abs(rs).max())
    np.testing.assert_allclose(dh[:, :A], rm, rtol=1e-4, atol=1e-6 * scale)
    np.testing.assert_allclose(dh[:, A:], rs, rtol=1e-4, atol=1e-6 * scale)


def _config(eng, g, **kw):
    am = float(g["action_max"])
    base = dict(state_dim=3, state_min=[-1.0, -1.0, -8.0], state_max=[1.0, 1.0, 8.0], action_dim=1, action_min=[-am],
                action_max=[am], tau=float(g["tau"]), norm_type="input_norm", random_seed=0,
                pi_lr=float(g["pi_lr"]), qf_vf_lr=float(g["qf_vf_lr"]), optim_type=str(g["optim_type"]),
                q_update_type=str(g["q_update_type"]), use_true_q="False", actor_l1_dim=int(g["l1"]),
                actor_l2_dim=int(g["l2"]), critic_l1_dim=int(g["l1"]), critic_l2_dim=int(g["l2"]),
                entropy_scale=float(g["alpha"]), N_param=int(g["n_param"]), l_param=6, batch_size=32, engine=eng)
    base.update(kw)
    return SimpleNamespace(**base)


@pytest.mark.parametrize("use_graph", [True, False])
@pytest.mark.parametrize("name", FULL + FULL_DENSE)
def test_dropin_update_network_matches_reference(eng, name, use_graph):
    """The recorded reference run, replayed through the drop-in class: same batches, same normal draws ->
    every parameter of q_net / v_net / target_v_net / pi_net after update 1 and update 2, the three
    losses, and sample_action / predict_action on the final networks."""
    from rlcontrol_b200 import kl_networks
    g, pre, post = load_full(name)
    cls = kl_networks.ForwardKLNetwork if "fkl" in name else kl_networks.ReverseKLNetwork
    net = cls(None, None, _config(eng, g, use_cuda_graph=use_graph))
    assert net.intgrl_actions_len == g["grid_a"].shape[0]
    net.load_reference_parameters(pre["q"], pre["v"], pre["tv"], pre["pi"])
    for u in range(g["s"].shape[0]):
        net.update_network(g["s"][u], g["a"][u], g["s2"][u], g["r"][u], g["g"][u], eps=g["eps"][u])
        net.update_target_network()
        np.testing.assert_allclose(net.last_losses, g["losses"][u], rtol=3e-4, atol=3e-5)
        mine = net.export_parameters()
        for k in ("q", "v", "tv", "pi"):
            for i, (m, ref) in enumerate(zip(mine[k], post[u][k])):
                move = np.abs(ref - pre[k][i]).max() + 1e-12
                err = np.abs(m.reshape(ref.shape) - ref).max()
                if name in FULL_DENSE:
                    # Adam's first step is lr * g / (|g| + 1e-8): where the 2048-row mean gradient of an element is within
                    # ~100x of eps, an fp32-class difference in g (3e-7 of the largest gradient) moves the STEP by percents
                    # of lr.  So: all but 0.1 % of the elements within 1e-2 of the move, none beyond 1e-1 of it.
                    e = np.abs(m.reshape(ref.shape) - ref).ravel()
                    assert np.quantile(e, 0.999) <= 1e-2 * move + 5e-7, (name, u, k, i, np.quantile(e, 0.999), move)
                    assert err <= 1e-1 * move + 5e-7, (name, u, k, i, err, move)
                else:
                    assert err <= 5e-3 * move + 5e-7, (name, u, k, i, err, move)
    st = g["act_states"]
    np.testing.assert_allclose(net.sample_action(st, eps=g["act_eps"]), g["act_sample"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(net.predict_action(st), g["act_predict"], rtol=1e-4, atol=2e-5)


def test_dropin_uses_the_references_random_streams(eng):
    """Under torch.manual_seed the drop-in draws its initial weights exactly as the reference's constructors do
    (nn.Linear order pi, q, v, target_v) and its policy noise from the stream normal.sample() consumes."""
    import torch
    from rlcontrol_b200 import kl_networks
    g, _, _ = load_full("full_fkl_intg_nonsac")
    torch.manual_seed(123)
    net = kl_networks.ForwardKLNetwork(None, None, _config(eng, g))
    torch.manual_seed(123)
    S, A, l1, l2 = 3, 1, int(g["l1"]), int(g["l2"])
    lin = torch.nn.Linear
    pi = [lin(S, l1), lin(l1, l2)]
    heads = []
    for _ in range(2):
        h = lin(l2, A)
        h.weight.data.uniform_(-3e-3, 3e-3)
        h.bias.data.uniform_(-3e-3, 3e-3)
        heads.append(h)
    q1 = lin(S + A, l1)
    p = net.export_parameters()
    np.testing.assert_array_equal(p["pi"][0], pi[0].weight.detach().numpy())
    np.testing.assert_array_equal(p["pi"][6], heads[1].weight.detach().numpy())
    np.testing.assert_array_equal(p["q"][0], q1.weight.detach().numpy())
    for a_, b_ in zip(p["v"], p["tv"]):
        np.testing.assert_array_equal(a_, b_)                   # target_v starts as a copy of v
    torch.manual_seed(7)
    a1 = net.sample_action(g["act_states"])
    torch.manual_seed(7)
    a2 = net.sample_action(g["act_states"], eps=torch.randn(5, 1))
    np.testing.assert_array_equal(a1, a2)


def test_dropin_rejects_what_the_reference_rejects(eng):
    from rlcontrol_b200 import kl_networks
    g, _, _ = load_full("full_fkl_intg_nonsac")
    with pytest.raises(NotImplementedError):
        kl_networks.ForwardKLNetwork(None, None, _config(eng, g, optim_type="ll"))     # forwardkl_network.py:152-153
    with pytest.raises(ValueError):
        kl_networks.ReverseKLNetwork(None, None, _config(eng, g, q_update_type="td"))  # :157 invalid q_update_type


def test_dropin_two_action_dims_matches_oracle(eng):
    """action_dim = 2 (Smolyak grid, MVN-with-std-as-covariance log-density): the reference itself cannot run this
    branch under torch 2.x (float64 grid), so the check is against the oracle restatement only."""
    from rlcontrol_b200 import kl_networks
    g, _, _ = load_full("full_fkl_intg_nonsac")
    rng = np.random.RandomState(3)
    S, A, B, l1, l2 = 4, 2, 16, 40, 24
    for cls, kind in ((kl_networks.ForwardKLNetwork, "fkl"), (kl_networks.ReverseKLNetwork, "rkl")):
        cfg = _config(eng, g, state_dim=S, state_min=[-3.0] * S, state_max=[3.0] * S, action_dim=A, action_min=[-1.5] * A,
                      action_max=[1.5] * A, actor_l1_dim=l1, actor_l2_dim=l2, critic_l1_dim=l1, critic_l2_dim=l2,
                      l_param=4, optim_type="intg", q_update_type="sac", entropy_scale=0.3)
        net = cls(None, None, cfg)
        p = net.export_parameters()
        p["q"][4] = p["q"][4] * 100
        p["pi"][4] = p["pi"][4] * 50
        p["pi"][6] = p["pi"][6] * 30
        net.load_reference_parameters(p["q"], p["v"], p["tv"], p["pi"])
        ag = okl.KLAgent(kind, p["q"], p["v"], p["tv"], p["pi"], net.intgrl_actions.cpu().numpy(),
                         net.intgrl_weights.cpu().numpy(), 1.5, 0.3, cfg.pi_lr, cfg.qf_vf_lr, cfg.tau, "intg", "sac")
        for u in range(2):
            s, a, s2 = rng.randn(B, S), rng.uniform(-1.5, 1.5, (B, A)), rng.randn(B, S)
            r, gm, eps = rng.randn(B), np.full(B, 0.99), rng.randn(B, A).astype(np.float32)
            net.update_network(s, a, s2, r, gm, eps=eps)
            net.update_target_network()
            ref_losses = ag.update(s.astype(np.float32), a.astype(np.float32), s2.astype(np.float32),
                                   r.astype(np.float32), gm.astype(np.float32), eps)
            np.testing.assert_allclose(net.last_losses, ref_losses, rtol=5e-4, atol=5e-5)
        mine = net.export_parameters()
        for k, ref in (("q", ag.q), ("v", ag.v), ("tv", ag.tv), ("pi", ag.pi)):
            for i, (m, r_) in enumerate(zip(mine[k], ref)):
                move = np.abs(r_ - p[k][i]).max() + 1e-12
                assert np.abs(m.reshape(r_.shape) - r_).max() <= 1e-2 * move + 1e-6, (kind, k, i)


# ---------------------------------------------------------------------------------------------
# small-minibatch fast path (csrc/small_batch.cu): the two fused launches against the oracle, through the C-ABI
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,inp,H1,H2,O", [(32, 3, 200, 200, 2), (7, 5, 33, 17, 3), (64, 23, 400, 300, 1), (1, 3, 64, 48, 4),
                                           (9, 4, 16, 200, 1)])
def test_small_batch_forward_and_update_match_oracle(eng, B, inp, H1, H2, O):
    import ctypes as C
    import torch
    from rlcontrol_b200 import _lib
    from rlcontrol_b200._lib import check
    rng = np.random.RandomState(B * 7 + O)
    p = _rand_mlp(rng, inp, H1, H2, O)
    n0 = inp // 2                                                     # two-part input rows, like Q(s, a)
    x = rng.randn(B, inp).astype(np.float32)
    dout = rng.randn(B, O).astype(np.float32)
    dev = eng.device
    m = rb_mlp(eng, inp, H1, H2, O, p)
    theta0 = m.theta.clone()
    x0, x1 = _t(eng, x[:, :n0].copy()), _t(eng, x[:, n0:].copy())
    z = lambda *sh: torch.zeros(sh, dtype=torch.float32, device=dev)
    h1, h2, out, w3 = z(B, H1), z(B, H2), z(B, O), z(H2 * O)
    state = torch.zeros((4,), dtype=torch.int32, device=dev)
    state[0] = 4                                                      # the fifth Adam step
    mom, var, target = _t(eng, rng.randn(m.theta.numel()) * 1e-2), _t(eng, rng.rand(m.theta.numel()) * 1e-3), m.theta.clone() + 0.5
    mom0, var0, target0 = mom.clone(), var.clone(), target.clone()
    lr, tau = 1e-2, 0.05
    net = (_lib.RlcSbNet * 1)()
    n = net[0]
    n.theta, n.inp, n.H1, n.H2, n.O = m.theta.data_ptr(), inp, H1, H2, O
    n.x0, n.n0, n.x1, n.n1 = (x0.data_ptr() if n0 else None), n0, x1.data_ptr(), inp - n0
    n.h1, n.h2, n.out, n.w3_snapshot = h1.data_ptr(), h2.data_ptr(), out.data_ptr(), w3.data_ptr()
    n.adam_state, n.lr, n.beta1, n.beta2, n.adam_variant = state.data_ptr(), lr, 0.9, 0.999, 0
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(eng.lib.rlc_sb_forward(eng.h, net, 1, B, st))
    ref_out, cache = okl.mlp_forward(x, *p)
    np.testing.assert_allclose(out.cpu().numpy(), ref_out, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(h1.cpu().numpy(), np.maximum(cache[1], 0), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(h2.cpu().numpy(), np.maximum(cache[3], 0), rtol=2e-5, atol=2e-5)
    np.testing.assert_array_equal(w3.cpu().numpy().reshape(H2, O), p[4].T)
    assert int(state[0]) == 5
    tr = (_lib.RlcSbTrain * 1)()
    t = tr[0]
    d_dev = _t(eng, dout)
    t.theta, t.m, t.v, t.adam_state = m.theta.data_ptr(), mom.data_ptr(), var.data_ptr(), state.data_ptr()
    t.beta1, t.beta2, t.eps, t.target, t.tau = 0.9, 0.999, 1e-8, target.data_ptr(), tau
    t.inp, t.H1, t.H2, t.O, t.x0, t.n0, t.x1, t.n1 = inp, H1, H2, O, n.x0, n0, n.x1, inp - n0
    t.h1, t.h2, t.out, t.w3_snapshot, t.role, t.dout = n.h1, n.h2, n.out, n.w3_snapshot, _lib.SB_ROLE_DOUT, d_dev.data_ptr()
    check(eng.lib.rlc_sb_update(eng.h, tr, 1, B, B, st))
    # reference: oracle gradients, then torch-flavoured Adam (step 5) and the Polyak step, in float64
    g = okl.mlp_grads(cache, dout, p[2], p[4])
    flat = np.concatenate([(a.T if a.ndim == 2 else a).reshape(-1) for a in g])           # theta is [in,out]-major
    m1 = 0.9 * mom0.cpu().numpy().astype(np.float64) + 0.1 * flat
    v1 = 0.999 * var0.cpu().numpy().astype(np.float64) + 0.001 * flat * flat
    step = lr / (1 - 0.9 ** 5) * m1 / (np.sqrt(v1) / np.sqrt(1 - 0.999 ** 5) + 1e-8)
    want = theta0.cpu().numpy().astype(np.float64) - step
    np.testing.assert_allclose(mom.cpu().numpy(), m1, rtol=1e-4, atol=1e-6 * max(1.0, np.abs(flat).max()))
    np.testing.assert_allclose(var.cpu().numpy(), v1, rtol=1e-4, atol=1e-9)
    got = m.theta.cpu().numpy()
    np.testing.assert_allclose(got - theta0.cpu().numpy(), -step, rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(target.cpu().numpy(), target0.cpu().numpy() + tau * (want - target0.cpu().numpy()), rtol=1e-5, atol=1e-6)


def rb_mlp(eng, inp, H1, H2, O, p):
    import rlcontrol_b200 as rb
    return rb.Mlp(eng, inp, H1, H2, O).load_torch(*p)


@pytest.mark.parametrize("name", FULL)
def test_small_batch_path_equals_generic_path(eng, name):
    """The fused launches and the generic multi-kernel update are the same arithmetic in a different summation order
    (the recorded reference run above goes through the fused path wherever it applies)."""
    from rlcontrol_b200 import kl_networks
    g, pre, _ = load_full(name)
    cls = kl_networks.ForwardKLNetwork if "fkl" in name else kl_networks.ReverseKLNetwork
    outs = []
    for fused in (True, False):
        net = cls(None, None, _config(eng, g, fused_small_batch=fused))
        if fused and not net._small_ok(int(g["s"].shape[1])):
            pytest.skip("variant not covered by the fused path (likelihood-ratio optim types)")
        assert fused or not net._small_ok(int(g["s"].shape[1]))
        net.load_reference_parameters(pre["q"], pre["v"], pre["tv"], pre["pi"])
        losses = []
        for u in range(g["s"].shape[0]):
            net.update_network(g["s"][u], g["a"][u], g["s2"][u], g["r"][u], g["g"][u], eps=g["eps"][u])
            net.update_target_network()
            losses.append(net.last_losses.copy())
        outs.append((net.export_parameters(), np.array(losses)))
    (pa, la), (pb, lb) = outs
    np.testing.assert_allclose(la, lb, rtol=2e-4, atol=1e-6)
    for k in ("q", "v", "tv", "pi"):
        for i, (a, b) in enumerate(zip(pa[k], pb[k])):
            move = np.abs(np.asarray(b) - pre[k][i].reshape(np.asarray(b).shape)).max() + 1e-12
            assert np.abs(a - b).max() <= 2e-3 * move + 2e-7, (k, i)


# ---------------------------------------------------------------------------------------------
# round-2 regressions (advisor findings)
# ---------------------------------------------------------------------------------------------
def _big_grid_agent(eng_factory, use_graph, alpha=0.2, B=64, seed=11):
    """An FKL agent whose grid evaluation takes the tensor path under precision='auto' (B * N >= 16384 rows, shared grid)."""
    import torch
    from rlcontrol_b200 import kl_networks
    g, _, _ = load_full("full_fkl_intg_nonsac")
    S, A, N = 5, 2, 512
    rng = np.random.RandomState(seed)
    grid = (rng.uniform(-1.5, 1.5, (N, A)).astype(np.float32), np.full(N, 2.0 / N, np.float32))
    cfg = _config(eng_factory(), g, state_dim=S, state_min=[-3.0] * S, state_max=[3.0] * S, action_dim=A,
                  action_min=[-1.5] * A, action_max=[1.5] * A, actor_l1_dim=64, actor_l2_dim=48, critic_l1_dim=128,
                  critic_l2_dim=96, optim_type="intg", q_update_type="non_sac", entropy_scale=alpha, batch_size=B,
                  integration_grid=grid, use_cuda_graph=use_graph, fused_small_batch=False)
    torch.manual_seed(seed)
    net = kl_networks.ForwardKLNetwork(None, None, cfg)
    p = net.export_parameters()
    p["q"][4] = p["q"][4] * 100          # a head that makes Q O(1) instead of O(1e-3)
    net.load_reference_parameters(p["q"], p["v"], p["tv"], p["pi"])
    batch = lambda: (rng.randn(B, S), rng.uniform(-1.5, 1.5, (B, A)), rng.randn(B, S), rng.randn(B), np.full(B, 0.99))
    return net, grid, batch


@pytest.mark.parametrize("use_graph", [True, False])
def test_eval_grid_after_graph_replay_uses_current_theta(eng, use_graph):
    """update -> eval_grid('auto') -> update -> eval_grid('auto'): the eager tensor-path evaluation between graph replays
    must see the parameters the replay just wrote (the cached operand pack is invalidated after every replay)."""
    import rlcontrol_b200 as rb
    net, grid, batch = _big_grid_agent(lambda: rb.Engine(0), use_graph)
    s_eval = np.random.RandomState(1).randn(64, 5).astype(np.float32)
    seen = []
    for _ in range(3):
        net.update_network(*batch())
        net.update_target_network()
        q_auto = net.q_net.eval_grid(s_eval, grid[0], "auto").cpu().numpy()       # 64 x 512 rows: split tensor mode
        q_32 = net.q_net.eval_grid(s_eval, grid[0], "fp32").cpu().numpy()
        assert net.eng.umma_error() == 0
        den = np.maximum(np.abs(q_32), np.sqrt((q_32 ** 2).mean(1, keepdims=True)))
        assert (np.abs(q_auto - q_32) / den).max() < 4e-5
        seen.append(q_32)
    assert np.abs(seen[2] - seen[0]).max() > 1e-4       # theta_Q really moved between the evaluations


def test_policy_gradient_at_small_entropy_scale_on_the_default_path(eng):
    """entropy_scale = 0.01 (the reference sweeps down to 0.001): exp(q / alpha) turns a 5e-3 error of Q into an O(1)
    error of the Boltzmann weights.  The default precision ('auto' = split tensor mode here) must reproduce the fp64
    oracle's per-state loss and policy-head gradients; the opt-in single-rounding fp16 mode is shown not to."""
    import bench
    import torch
    from oracle import oracle_np as onp
    from test_gpu_parity import _tin
    W = bench.WORKLOAD
    params = bench.make_params(np.random.RandomState(0), W["S"], W["A"], W["H1"], W["H2"])
    s, a, w, (mean, lstd) = bench.make_inputs(np.random.RandomState(1000), 32, W["N"], W["S"], W["A"])
    alpha = 0.01
    cr = _tin(eng, params, W["S"], W["A"], W["H1"], W["H2"])
    t = lambda x: torch.as_tensor(x, device=eng.device)
    q_ref = onp.tin_eval(s, a, params, dtype=np.float64)
    ref = onp.fkl_policy_reduce(q_ref, w, a, mean, lstd, 1.0, alpha)

    def run(prec):
        q = cr.eval(s, a, prec)
        loss_b, dm, ds, _ = eng.fkl_policy(q, t(w), t(a), 1.0, t(mean), t(lstd), alpha)
        return [x.cpu().numpy() for x in (loss_b, dm, ds)]

    def worst(out):
        return max(np.abs(o - r).max() / max(np.abs(r).max(), 1e-30) for o, r in zip(out, ref[:3]))
    assert cr.tensor_arithmetic(True, "fp16x3") == "grid3"
    assert torch.equal(cr.eval(s, a, "auto"), cr.eval(s, a, "fp16x3"))
    e_auto, e_fast = worst(run("auto")), worst(run("fp16"))
    assert e_auto < 2e-3, e_auto                    # q error 1e-5 * |q| / alpha ~ 1e-3 in the exponent
    assert e_fast > 10 * e_auto, (e_fast, e_auto)   # why 'fp16' is opt-in


def test_replay_sample_feeds_update_network_directly(eng):
    """The two drop-ins compose like the reference's BaseAgent.learn (base_agent.py:64-70): sample_batch's output goes
    straight into update_network -- numpy by default, device tensors with as_numpy=False."""
    import rlcontrol_b200 as rb
    from rlcontrol_b200 import kl_networks
    from rlcontrol_b200.replaybuffer import ReplayBuffer
    g, pre, _ = load_full("full_rkl_intg_nonsac")
    net = kl_networks.ReverseKLNetwork(None, None, _config(rb.Engine(0), g))
    net.load_reference_parameters(pre["q"], pre["v"], pre["tv"], pre["pi"])
    buf = ReplayBuffer(1000, 0, 3, 1, engine=eng)
    rng = np.random.RandomState(0)
    for _ in range(200):
        buf.add(rng.randn(3), rng.uniform(-2, 2, 1), rng.randn(), rng.randn(3), 0.99)
    out = buf.sample_batch(32)
    assert all(isinstance(x, np.ndarray) for x in out)
    state, action, reward, next_state, gamma = out               # base_agent.py:66-67, the same unpacking and call
    net.update_network(state, action, next_state, reward, gamma)
    assert np.isfinite(net.last_losses).all()
    state, action, reward, next_state, gamma = buf.sample_batch(32, as_numpy=False)
    assert all(x.is_cuda for x in (state, action, reward, next_state, gamma))
    net.update_network(state, action, next_state, reward, gamma)
    assert np.isfinite(net.last_losses).all()
