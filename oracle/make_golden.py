#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ by running the UNMODIFIED
reference classes from /root/reference in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/oracle_np.py).  /root/reference does not exist on the GPU
box, so nothing at test/bench time reads it: this script is run once here and its outputs are
committed.  Re-run:  python oracle/make_golden.py

What is generated
-----------------
tin_cfg1.npz / tin_400_300.npz / tin_cfg4_exact.npz
    ``SoftQNetwork`` weights (torch init, seeded), states, CC grid, and the q values returned by
    the reference ``q_net(stacked_s, stacked_a)`` (forwardkl_network.py:160-164).
fkl_update.npz / rkl_update.npz
    one full reference ``update_network`` (forwardkl_network.py:123-209 /
    reversekl_network.py:130-217) with the policy sample fixed by seeding torch: inputs,
    pre/post q_net parameters (pins the critic regression step, a15), grid q, grid logp,
    and the policy loss (pins a3 / a4).
trueq.npz
    the five Bimodal1DEnv_trueQ_ckpt critics decoded from the TF bundle + Q on a 401 grid.
gmm.npz
    scikit-learn GaussianMixture run through a BoundedVarGaussianMixture-equivalent subclass
    (utils/boundedvar_gaussian_mixture.py) with the k-means responsibilities captured.
cc.npz
    Clenshaw-Curtis nodes/weights checks.
"""
import importlib.machinery
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import oracle_np as onp  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    """tensorflow / gym / matplotlib / quadpy are absent here; the torch agents import them at
    module load but only call quadpy.c1.clenshaw_curtis."""
    _stub("tensorflow")
    _stub("gym")
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = _stub("matplotlib.pyplot")

    class _Scheme:
        def __init__(self, n):
            self.points, self.weights = onp.clenshaw_curtis(n)

    c1 = types.SimpleNamespace(clenshaw_curtis=lambda n: _Scheme(n))
    _stub("quadpy", c1=c1, line_segment=c1)


def make_config(state_dim, action_dim, action_max, batch_size, n_param, l1, l2, alpha, state_max=None):
    cfg = types.SimpleNamespace()
    cfg.state_dim = state_dim
    cfg.state_min = -np.ones(state_dim) * 10 if state_max is None else -np.asarray(state_max)
    cfg.state_max = np.ones(state_dim) * 10 if state_max is None else np.asarray(state_max)
    cfg.action_dim = action_dim
    cfg.action_min = np.array([-action_max] * action_dim)
    cfg.action_max = np.array([action_max] * action_dim)
    cfg.tau = 0.01
    cfg.norm_type = "input_norm"
    cfg.pi_lr = 1e-3
    cfg.qf_vf_lr = 1e-3
    cfg.optim_type = "intg"
    cfg.use_true_q = "False"
    cfg.random_seed = 0
    cfg.actor_l1_dim = l1
    cfg.actor_l2_dim = l2
    cfg.critic_l1_dim = l1
    cfg.critic_l2_dim = l2
    cfg.entropy_scale = alpha
    cfg.N_param = n_param
    cfg.l_param = 6
    cfg.batch_size = batch_size
    cfg.q_update_type = "non_sac"
    return cfg


def q_params(qnet):
    return [p.detach().numpy().copy() for p in
            (qnet.linear1.weight, qnet.linear1.bias, qnet.linear2.weight, qnet.linear2.bias,
             qnet.linear3.weight, qnet.linear3.bias)]


def gen_tin(torch, fkl_mod, name, S, A, B, n_param, l1, l2, action_max, scale_last=1.0, seed=0):
    torch.manual_seed(seed)
    net = fkl_mod.SoftQNetwork(S, A, l1, l2)
    with torch.no_grad():
        net.linear3.weight.mul_(scale_last)
        net.linear3.bias.mul_(scale_last)
    rng = np.random.RandomState(seed)
    s = np.clip(rng.randn(B, S), -10, 10).astype(np.float32)
    if A == 1:
        acts, w = onp.intg_grid_1d(n_param, action_max)
    else:
        N = n_param - 2
        acts = (rng.uniform(-1, 1, size=(N, A)) * action_max).astype(np.float32)
        _, w = onp.intg_grid_1d(n_param, 1.0)
    N = acts.shape[0]
    ts = torch.from_numpy(s)
    ta = torch.from_numpy(acts)
    stacked_s = ts.unsqueeze(1).repeat(1, N, 1).reshape(-1, S)          # forwardkl_network.py:160-161
    stacked_a = ta.unsqueeze(0).repeat(B, 1, 1).reshape(-1, A)          # forwardkl_network.py:104-105
    with torch.no_grad():
        q = net(stacked_s, stacked_a).reshape(B, N).numpy()
    W1, b1, W2, b2, W3, b3 = q_params(net)
    np.savez_compressed(os.path.join(OUT, name), s=s, a=acts, w=w, q=q,
                        W1=W1, b1=b1, W2=W2, b2=b2, W3=W3, b3=b3,
                        action_max=np.float32(action_max))
    print(name, "q", q.shape, float(np.abs(q).mean()))


def gen_update(torch, mod, cls_name, name, alpha, seed=0):
    """One reference update_network on a Pendulum-shaped batch (cfg1: S=3, A=1, B=32, N_param=64,
    200-200).  Hooks record the grid q / logp the reference computes internally."""
    S, A, B = 3, 1, 32
    cfg = make_config(S, A, 2.0, B, 64, 200, 200, alpha, state_max=[1.0, 1.0, 8.0])
    torch.manual_seed(seed)
    net = getattr(mod, cls_name)(None, None, cfg)
    rng = np.random.RandomState(seed + 1)
    s = rng.uniform(-1, 1, size=(B, S)) * np.array([1, 1, 8.0])
    a = rng.uniform(-2, 2, size=(B, A))
    s2 = rng.uniform(-1, 1, size=(B, S)) * np.array([1, 1, 8.0])
    r = -rng.uniform(0, 16, size=(B,))
    g = np.full((B,), 0.99)
    # make Q non-trivial: scale the last critic layer so q is O(1)
    with torch.no_grad():
        net.q_net.linear3.weight.mul_(100.0)
        net.q_net.linear3.bias.mul_(100.0)
    pre_q = q_params(net.q_net)
    pre_v = [p.detach().numpy().copy() for p in net.v_net.parameters()]
    pre_tv = [p.detach().numpy().copy() for p in net.target_v_net.parameters()]
    pre_pi = [p.detach().numpy().copy() for p in net.pi_net.parameters()]

    rec = {}
    orig_q_forward = net.q_net.forward
    calls = []

    def q_hook(state, action):
        out = orig_q_forward(state, action)
        calls.append((state.detach().numpy().copy(), action.detach().numpy().copy(),
                      out.detach().numpy().copy()))
        return out

    net.q_net.forward = q_hook
    orig_logprob = net.pi_net.get_logprob

    def lp_hook(states, tiled_actions, epsilon=1e-6):
        out = orig_logprob(states, tiled_actions, epsilon)
        rec["logp"] = out.detach().numpy().reshape(B, -1).copy()
        return out

    net.pi_net.get_logprob = lp_hook
    orig_v = net.v_net.forward

    def v_hook(state):
        out = orig_v(state)
        rec.setdefault("v", out.detach().numpy().copy())
        return out

    net.v_net.forward = v_hook

    # capture the policy loss through the optimizer's backward: wrap pi_optimizer.step
    losses = {}
    orig_backward = torch.Tensor.backward
    order = []

    def backward_hook(self, *a_, **k_):
        order.append(float(self.detach()))
        return orig_backward(self, *a_, **k_)

    torch.Tensor.backward = backward_hook
    try:
        torch.manual_seed(seed + 2)
        net.update_network(s, a, s2, r, g)
    finally:
        torch.Tensor.backward = orig_backward
    losses["q_loss"], losses["v_loss"], losses["pi_loss"] = order
    post_q = q_params(net.q_net)
    # the grid evaluation is the q_net call with B*N rows
    N = net.intgrl_actions_len
    grid_call = [c for c in calls if c[0].shape[0] == B * N][0]
    reg_call = calls[0]                                   # q_net(state_batch, action_batch)
    save = dict(
        s=s, a=a, s2=s2, r=r, g=g, alpha=np.float64(alpha),
        grid_a=net.intgrl_actions.numpy(), grid_w=net.intgrl_weights.numpy(),
        grid_q=grid_call[2].reshape(B, N), logp=rec["logp"], v=rec["v"].reshape(B),
        q_reg=reg_call[2].reshape(B),
        q_loss=np.float64(losses["q_loss"]), pi_loss=np.float64(losses["pi_loss"]),
        lr=np.float64(cfg.qf_vf_lr),
    )
    # regression target used by the reference: r + gamma * target_v(s2) (forwardkl_network.py:137-138)
    with torch.no_grad():
        tv = net.target_v_net
        # target net is untouched by update_network (Polyak happens in update_target_network)
        x = torch.from_numpy(s2.astype(np.float32))
        y = torch.from_numpy(r.astype(np.float32)).unsqueeze(-1) + \
            torch.from_numpy(g.astype(np.float32)).unsqueeze(-1) * tv(x)
    save["y"] = y.numpy().reshape(B)
    for i, nm in enumerate(["W1", "b1", "W2", "b2", "W3", "b3"]):
        save["pre_" + nm] = pre_q[i]
        save["post_" + nm] = post_q[i]
    np.savez_compressed(os.path.join(OUT, name), **save)
    print(name, "pi_loss", losses["pi_loss"], "q_loss", losses["q_loss"], "N", N)


def _net_params(net):
    """(q, v, target_v, pi) parameter lists in torch layout, copied."""
    cp = lambda ps: [p.detach().numpy().copy() for p in ps]
    pi = net.pi_net
    return dict(
        q=q_params(net.q_net),
        v=cp([net.v_net.linear1.weight, net.v_net.linear1.bias, net.v_net.linear2.weight,
              net.v_net.linear2.bias, net.v_net.linear3.weight, net.v_net.linear3.bias]),
        tv=cp([net.target_v_net.linear1.weight, net.target_v_net.linear1.bias, net.target_v_net.linear2.weight,
               net.target_v_net.linear2.bias, net.target_v_net.linear3.weight, net.target_v_net.linear3.bias]),
        pi=cp([pi.linear1.weight, pi.linear1.bias, pi.linear2.weight, pi.linear2.bias, pi.mean_linear.weight,
               pi.mean_linear.bias, pi.log_std_linear.weight, pi.log_std_linear.bias]))


def gen_full_update(torch, mod, cls_name, name, alpha, optim_type="intg", q_update_type="non_sac", seed=0,
                    n_updates=2, l1=64, l2=48, B=32):
    """n_updates consecutive reference ``update_network`` + ``update_target_network`` calls
    (forwardkl_network.py:123-215 / reversekl_network.py:130-224) on Pendulum-shaped batches, ALL three
    networks recorded before and after every update, and the N(0,1) draws behind ``normal.sample()``
    recorded as ``eps`` (torch.randn under the same seed; asserted to reproduce the reference's z).
    B = 32 is the reference's minibatch; the ``*_dense`` fixtures (B = 2048, 128-128) put the same update on the
    tensor-core training GEMMs of the B200 path (csrc/rows_gemm_tc.cu)."""
    S, A = 3, 1
    cfg = make_config(S, A, 2.0, B, 64, l1, l2, alpha, state_max=[1.0, 1.0, 8.0])
    cfg.optim_type, cfg.q_update_type = optim_type, q_update_type
    cfg.pi_lr, cfg.qf_vf_lr = 1e-3, 2e-3
    torch.manual_seed(seed)
    net = getattr(mod, cls_name)(None, None, cfg)
    with torch.no_grad():                      # O(1) Q, V and policy heads so every term of the update matters
        net.q_net.linear3.weight.mul_(100.0); net.q_net.linear3.bias.mul_(100.0)
        net.v_net.linear3.weight.mul_(100.0); net.v_net.linear3.bias.mul_(100.0)
        net.target_v_net.linear3.weight.mul_(80.0); net.target_v_net.linear3.bias.mul_(80.0)
        net.pi_net.mean_linear.weight.mul_(60.0); net.pi_net.log_std_linear.weight.mul_(40.0)
        net.pi_net.log_std_linear.bias.sub_(0.7)
    rng = np.random.RandomState(seed + 11)
    save = dict(alpha=np.float64(alpha), pi_lr=np.float64(cfg.pi_lr), qf_vf_lr=np.float64(cfg.qf_vf_lr),
                tau=np.float64(cfg.tau), optim_type=np.array(optim_type), q_update_type=np.array(q_update_type),
                action_max=np.float64(2.0), n_param=np.int64(64), l1=np.int64(l1), l2=np.int64(l2),
                grid_a=net.intgrl_actions.numpy(), grid_w=net.intgrl_weights.numpy())
    for k, ps in _net_params(net).items():
        for i, p_ in enumerate(ps):
            save["pre_%s_%d" % (k, i)] = p_
    orig_eval = net.pi_net.evaluate
    zs = []

    def eval_hook(state, epsilon=1e-6):
        out = orig_eval(state, epsilon)
        zs.append((out[2].detach().numpy().copy(), out[1].detach().numpy().copy()))
        return out

    net.pi_net.evaluate = eval_hook
    orig_backward = torch.Tensor.backward
    order = []

    def backward_hook(self, *a_, **k_):
        order.append(float(self.detach()))
        return orig_backward(self, *a_, **k_)

    batches = {k: [] for k in ("s", "a", "s2", "r", "g", "eps", "z", "logp_sample")}
    losses = []
    for u in range(n_updates):
        s = rng.uniform(-1, 1, size=(B, S)) * np.array([1, 1, 8.0])
        a = rng.uniform(-2, 2, size=(B, A))
        s2 = rng.uniform(-1, 1, size=(B, S)) * np.array([1, 1, 8.0])
        r = -rng.uniform(0, 16, size=(B,))
        g = np.where(rng.uniform(size=B) < 0.1, 0.0, 0.99)
        torch.manual_seed(seed + 100 + u)
        eps = torch.randn(B, A).numpy().copy()
        with torch.no_grad():
            mean0, log_std0 = net.pi_net.forward(torch.from_numpy(s.astype(np.float32)))
        torch.manual_seed(seed + 100 + u)
        zs.clear(); order.clear()
        torch.Tensor.backward = backward_hook
        try:
            net.update_network(s, a, s2, r, g)
        finally:
            torch.Tensor.backward = orig_backward
        net.update_target_network()
        z_ref = zs[0][0]
        z_mine = mean0.numpy() + np.exp(log_std0.numpy()) * eps
        assert np.allclose(z_ref, z_mine, rtol=0, atol=1e-6), "normal.sample() is not mean + std * randn"
        losses.append(order[:3])
        for k, v_ in zip(("s", "a", "s2", "r", "g", "eps", "z", "logp_sample"),
                         (s, a, s2, r, g, eps, z_ref, zs[0][1].reshape(B))):
            batches[k].append(v_)
        for k, ps in _net_params(net).items():
            for i, p_ in enumerate(ps):
                save["post%d_%s_%d" % (u, k, i)] = p_
    net.pi_net.evaluate = orig_eval
    for k, v_ in batches.items():
        save[k] = np.asarray(v_)
    save["losses"] = np.asarray(losses, np.float64)          # [n_updates, (q_loss, v_loss, pi_loss)]
    # sample_action / predict_action on the final networks (forwardkl_network.py:109-121)
    st = rng.uniform(-1, 1, size=(5, S)) * np.array([1, 1, 8.0])
    torch.manual_seed(seed + 500)
    eps_act = torch.randn(5, A).numpy().copy()
    torch.manual_seed(seed + 500)
    save["act_states"], save["act_eps"] = st, eps_act
    save["act_sample"] = net.sample_action(st)
    save["act_predict"] = net.predict_action(st)
    np.savez_compressed(os.path.join(OUT, name), **save)
    print(name, "losses", np.asarray(losses))


def gen_trueq():
    variants = {
        "eq_var1": ((-0.6, 0.6), (0.2, 0.2), (1.0, 1.0)),      # environments.py:573-587
        "eq_var2": ((-0.8, 0.8), (0.2, 0.2), (1.0, 1.0)),      # :660-674
        "eq_var3": ((-1.0, 1.0), (0.2, 0.2), (1.0, 1.0)),      # :747-761
        "uneq_var1": ((-1.0, 1.0), (0.4, 0.2), (1.0, 1.5)),    # :312-326
        "uneq_var2": ((-1.0, 1.0), (0.3, 0.1), (1.0, 1.5)),    # :399-413
    }
    grid = np.linspace(-2, 2, 401).astype(np.float32)
    save = {"grid": grid}
    for v, (mx, sd, h) in variants.items():
        p = onp.decode_trueq_checkpoint(
            os.path.join(REF, "Bimodal1DEnv_trueQ_ckpt", f"Bimodal1DEnv_{v}_trueQ_learned.data-00000-of-00001"))
        for k, arr in p.items():
            save[f"{v}_{k}"] = arr
        save[f"{v}_reward"] = onp.bimodal_reward(grid, mx, sd, h)
        q = onp.tmid_forward(np.zeros((401, 1), np.float32), grid[:, None],
                             p["W1"], p["b1"], p["W2"], p["b2"], p["W3"], p["b3"])
        print("trueq", v, "max|Q-r|", float(np.max(np.abs(q - save[f"{v}_reward"]))))
    np.savez_compressed(os.path.join(OUT, "trueq.npz"), **save)


def gen_gmm():
    from sklearn.mixture import GaussianMixture
    from sklearn.mixture._gaussian_mixture import (_compute_precision_cholesky,
                                                   _estimate_gaussian_parameters)

    captured = {}

    class Bounded(GaussianMixture):
        """Same two overrides as utils/boundedvar_gaussian_mixture.py:13-75 on sklearn>=0.24's
        module path (the reference imports the removed ``sklearn.mixture.gaussian_mixture``)."""

        def _initialize(self, X, resp, xp=None):
            captured["resp0"] = resp.copy()
            n_samples, _ = X.shape
            weights, means, covariances = _estimate_gaussian_parameters(
                X, resp, self.reg_covar, self.covariance_type)
            means = np.clip(means, -2, 2)
            covariances = np.clip(covariances, np.exp(-2), np.exp(2))
            weights /= n_samples
            self.weights_ = weights
            self.means_ = means
            self.covariances_ = covariances
            self.precisions_cholesky_ = _compute_precision_cholesky(covariances, self.covariance_type)

        def _m_step(self, X, log_resp, xp=None):
            n_samples, _ = X.shape
            self.weights_, self.means_, self.covariances_ = _estimate_gaussian_parameters(
                X, np.exp(log_resp), self.reg_covar, self.covariance_type)
            self.means_ = np.clip(self.means_, -2, 2)
            self.covariances_ = np.clip(self.covariances_, np.exp(-2), np.exp(2))
            self.weights_ /= n_samples
            self.precisions_cholesky_ = _compute_precision_cholesky(self.covariances_, self.covariance_type)

    rng = np.random.RandomState(0)
    Xs, R0, Ws, Ms, Cs, Its = [], [], [], [], [], []
    import warnings
    for case in range(24):
        A = [1, 2, 6][case % 3]
        k = [6, 6, 12, 20][case % 4]
        centers = rng.uniform(-2.5, 2.5, size=(2, A))
        lab = rng.randint(0, 2, size=k)
        X = centers[lab] + rng.randn(k, A) * rng.uniform(0.05, 1.5)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gm = Bounded(n_components=2, random_state=rng, covariance_type="diag", tol=1e-2).fit(X)
        pad = np.zeros((20, 6))
        pad[:k, :A] = X
        Xs.append(pad)
        r0 = np.zeros((20, 2))
        r0[:k] = captured["resp0"]
        R0.append(r0)
        Ws.append(gm.weights_.copy())
        m = np.zeros((2, 6)); m[:, :A] = gm.means_
        c = np.zeros((2, 6)); c[:, :A] = gm.covariances_
        Ms.append(m); Cs.append(c); Its.append(gm.n_iter_)
    np.savez_compressed(os.path.join(OUT, "gmm.npz"), X=np.array(Xs), resp0=np.array(R0),
                        weights=np.array(Ws), means=np.array(Ms), covs=np.array(Cs),
                        n_iter=np.array(Its),
                        k=np.array([[6, 6, 12, 20][c % 4] for c in range(24)]),
                        A=np.array([[1, 2, 6][c % 3] for c in range(24)]))
    print("gmm cases", len(Xs), "n_iter", Its)


def gen_policy_logp(torch, fkl_mod, rkl_mod):
    """The reference's own PolicyNetwork.forward + get_logprob (forwardkl_network.py:288-344,
    reversekl_network.py twin) for action_dim 1, 2 and 3 -- pins the tanh-Gaussian log-density
    including the A>1 quirk (MultivariateNormal(mean, diag_embed(std)): std used as the covariance),
    and, through autograd, d(sum c*logp)/d(mean, log_std)."""
    save = {}
    for A, mod in ((1, fkl_mod), (2, fkl_mod), (3, rkl_mod)):
        torch.manual_seed(10 + A)
        rng = np.random.RandomState(20 + A)
        S, B, N, scale = 4, 6, 40, 2.0
        pi = mod.PolicyNetwork(S, A, 16, 16, scale)
        with torch.no_grad():                       # make the heads non-trivial
            pi.mean_linear.weight.mul_(100.0)
            pi.log_std_linear.weight.mul_(150.0)
        states = torch.tensor(rng.randn(B, S), dtype=torch.float32)
        acts = torch.tensor(rng.uniform(-0.98, 0.98, (N, A)) * scale, dtype=torch.float32)
        tiled = acts.unsqueeze(0).repeat(B, 1, 1)                     # [B,N,A] as the reference tiles it
        mean, log_std = pi.forward(states)
        mean.retain_grad()
        log_std.retain_grad()
        std = log_std.exp()
        normal = pi.get_distribution(mean, std)
        na = tiled.permute(1, 0, 2) / pi.action_scale
        lp = normal.log_prob(pi.atanh(na))
        if len(lp.shape) == 2:
            lp = lp.unsqueeze(-1)
        lp = lp - torch.log(1 - na.pow(2) + 1e-6).sum(dim=-1, keepdim=True)
        lp = lp.permute(1, 0, 2).reshape(B, N)
        ref = pi.get_logprob(states, tiled).reshape(B, N)
        assert torch.allclose(lp, ref)
        c = torch.tensor(rng.randn(B, N), dtype=torch.float32)
        (c * lp).sum().backward()
        save.update({f"A{A}_mean": mean.detach().numpy(), f"A{A}_log_std": log_std.detach().numpy(),
                     f"A{A}_actions": acts.numpy(), f"A{A}_scale": np.float32(scale),
                     f"A{A}_logp": ref.detach().numpy(), f"A{A}_c": c.numpy(),
                     f"A{A}_dmean": mean.grad.numpy(), f"A{A}_dlog_std": log_std.grad.numpy()})
    np.savez_compressed(os.path.join(OUT, "policy_logp.npz"), **save)
    print("policy_logp", {k: v.shape for k, v in save.items() if k.endswith("logp")})


def gen_smolyak(forwardkl_network):
    """The action_dim > 1 integration grid as the reference's constructor builds it
    (forwardkl_network.py:73-102).  (Its update_network cannot run for A > 1 under torch 2.x: the grid is
    float64 and torch.cat/Linear refuse the mixed dtypes -- so only the grid is pinned.)"""
    save = {}
    for A, l in ((2, 6), (3, 4)):
        cfg = make_config(3, A, 1.5, 8, 64, 32, 32, 0.1)
        cfg.l_param = l
        net = forwardkl_network.ForwardKLNetwork(None, None, cfg)
        save["a_%d_%d" % (A, l)] = net.intgrl_actions.numpy()
        save["w_%d_%d" % (A, l)] = net.intgrl_weights.numpy()
    np.savez_compressed(os.path.join(OUT, "smolyak.npz"), **save)


def gen_full_updates(torch, forwardkl_network, reversekl_network):
    gen_smolyak(forwardkl_network)
    F, R = (forwardkl_network, "ForwardKLNetwork"), (reversekl_network, "ReverseKLNetwork")
    gen_full_update(torch, *F, "full_fkl_intg_nonsac.npz", alpha=0.1)
    gen_full_update(torch, *F, "full_fkl_intg_sac.npz", alpha=0.5, q_update_type="sac", seed=3)
    gen_full_update(torch, *R, "full_rkl_intg_nonsac.npz", alpha=0.1, seed=4)
    gen_full_update(torch, *R, "full_rkl_hardintg_sac.npz", alpha=0.2, optim_type="hard_intg", q_update_type="sac", seed=5)
    gen_full_update(torch, *R, "full_rkl_ll_nonsac.npz", alpha=0.1, optim_type="ll", seed=6)
    gen_full_update(torch, *R, "full_rkl_hardll_sac.npz", alpha=0.1, optim_type="hard_ll", q_update_type="sac", seed=7)
    gen_full_update(torch, *F, "full_fkl_intg_nonsac_dense.npz", alpha=0.1, seed=8, n_updates=1, l1=128, l2=128, B=2048)
    gen_full_update(torch, *R, "full_rkl_intg_sac_dense.npz", alpha=0.2, q_update_type="sac", seed=9, n_updates=1, l1=128,
                    l2=128, B=2048)


def gen_bimodal_env():
    """step() of the reference's seven Bimodal1D bandit classes on float32 action arrays (what the agents hand to
    env.step): rewards and next states, for csrc/envloop.cu and oracle/oracle_env.py."""
    import warnings
    warnings.simplefilter("ignore", DeprecationWarning)
    from environments import environments as envs
    rng = np.random.RandomState(11)
    actions = np.concatenate([np.linspace(-2, 2, 81), rng.uniform(-2, 2, 47)]).astype(np.float32)
    save = {"actions": actions}
    for name in ("Bimodal1DEnv", "Bimodal1DEnv_uneq_var1", "Bimodal1DEnv_uneq_var2", "Bimodal1DEnv_uneq_var3",
                 "Bimodal1DEnv_eq_var1", "Bimodal1DEnv_eq_var2", "Bimodal1DEnv_eq_var3"):
        env = envs.create_environment({"environment": name, "TotalMilSteps": 0.001, "EpisodeSteps": -1,
                                       "EvalIntervalMilSteps": 0.0001, "EvalEpisodes": 1})
        rew, nxt = [], []
        for a in actions:
            env.reset()
            s2, r, done, _ = env.step(np.array([a], np.float32))
            assert done is True and env.EPISODE_STEPS_LIMIT == 1
            rew.append(float(r))
            nxt.append(float(s2[0]))
        save[name + "_reward"], save[name + "_next"] = np.array(rew, np.float64), np.array(nxt, np.float64)
        save[name + "_bounds"] = np.array([env.state_min[0], env.state_max[0], env.action_min[0], env.action_max[0]])
    np.savez_compressed(os.path.join(OUT, "bimodal_env.npz"), **save)
    print("bimodal_env.npz written")


def main():
    os.makedirs(OUT, exist_ok=True)
    install_stubs()
    sys.path.insert(0, REF)
    import torch
    torch.set_num_threads(1)
    from agents.network import forwardkl_network, reversekl_network

    gen_tin(torch, forwardkl_network, "tin_cfg1.npz", S=3, A=1, B=32, n_param=64, l1=200, l2=200,
            action_max=2.0, scale_last=100.0)
    gen_tin(torch, forwardkl_network, "tin_400_300.npz", S=17, A=6, B=8, n_param=258, l1=400, l2=300,
            action_max=1.0, scale_last=100.0, seed=1)
    gen_tin(torch, reversekl_network, "tin_cfg4_exact.npz", S=3, A=1, B=4, n_param=1026, l1=400, l2=300,
            action_max=2.0, scale_last=1.0, seed=2)
    gen_update(torch, forwardkl_network, "ForwardKLNetwork", "fkl_update.npz", alpha=0.1)
    gen_update(torch, reversekl_network, "ReverseKLNetwork", "rkl_update.npz", alpha=0.1)
    gen_policy_logp(torch, forwardkl_network, reversekl_network)
    gen_full_updates(torch, forwardkl_network, reversekl_network)
    gen_trueq()
    gen_gmm()
    gen_bimodal_env()
    x, w = onp.clenshaw_curtis(64)
    np.savez_compressed(os.path.join(OUT, "cc.npz"), x64=x, w64=w)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "full_updates":      # only the full-update fixtures
        os.makedirs(OUT, exist_ok=True)
        install_stubs()
        sys.path.insert(0, REF)
        import torch
        torch.set_num_threads(1)
        from agents.network import forwardkl_network, reversekl_network
        gen_full_updates(torch, forwardkl_network, reversekl_network)
    elif len(sys.argv) > 1 and sys.argv[1] == "dense_updates":   # only the dense-minibatch full-update fixtures
        os.makedirs(OUT, exist_ok=True)
        install_stubs()
        sys.path.insert(0, REF)
        import torch
        torch.set_num_threads(1)
        from agents.network import forwardkl_network, reversekl_network
        gen_full_update(torch, forwardkl_network, "ForwardKLNetwork", "full_fkl_intg_nonsac_dense.npz", alpha=0.1, seed=8,
                        n_updates=1, l1=128, l2=128, B=2048)
        gen_full_update(torch, reversekl_network, "ReverseKLNetwork", "full_rkl_intg_sac_dense.npz", alpha=0.2,
                        q_update_type="sac", seed=9, n_updates=1, l1=128, l2=128, B=2048)
    elif len(sys.argv) > 1 and sys.argv[1] == "bimodal_env":     # only the bandit-environment fixture
        install_stubs()
        sys.path.insert(0, REF)
        gen_bimodal_env()
    else:
        main()
