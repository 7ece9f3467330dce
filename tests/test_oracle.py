"""CPU tests: the oracle restatement against the committed golden vectors (generated from the
reference's own classes by oracle/make_golden.py) and internal consistency."""
import numpy as np
import pytest

from conftest import golden
from oracle import oracle_np as onp


def _params(g, prefix=""):
    return [g[prefix + k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")]


@pytest.mark.parametrize("name", ["tin_cfg1.npz", "tin_400_300.npz", "tin_cfg4_exact.npz"])
def test_tin_forward_matches_reference_softqnetwork(name):
    g = golden(name)
    q = onp.tin_eval(g["s"], g["a"], _params(g))
    assert q.shape == g["q"].shape
    np.testing.assert_allclose(q, g["q"], rtol=2e-5, atol=2e-6)


def test_trueq_checkpoints_reproduce_reward():
    g = golden("trueq.npz")
    grid = g["grid"]
    bounds = {"eq_var1": 0.05, "eq_var2": 0.08, "eq_var3": 0.11, "uneq_var1": 0.05, "uneq_var2": 0.11}
    for v, bound in bounds.items():
        p = [g[f"{v}_{k}"] for k in ("W1", "b1", "W2", "b2", "W3", "b3")]
        assert p[2].shape == (201, 200)            # rows 0..199 <-> h1, row 200 <-> action
        q = onp.tmid_forward(np.zeros((grid.size, 1), np.float32), grid[:, None], *p)
        assert np.max(np.abs(q - g[f"{v}_reward"])) < bound
        # concat order matters: swapping the action row to the front must break the fit
        W2s = np.concatenate([p[2][200:], p[2][:200]], axis=0)
        qs = onp.tmid_forward(np.zeros((grid.size, 1), np.float32), grid[:, None], p[0], p[1], W2s, *p[3:])
        assert np.max(np.abs(qs - g[f"{v}_reward"])) > 2 * bound


def test_tmid_hoisted_equals_naive():
    rng = np.random.RandomState(0)
    S, A, H1, H2, B, N = 5, 3, 40, 30, 7, 33
    p = [rng.randn(S, H1) * .3, rng.randn(H1) * .1, rng.randn(H1 + A, H2) * .3, rng.randn(H2) * .1,
         rng.randn(H2, 1), rng.randn(1)]
    s = rng.randn(B, S) * 3
    smin, smax = -np.ones(S), np.ones(S) * 2
    for a in (rng.randn(N, A), rng.randn(B, N, A)):
        q0 = onp.tmid_eval(s, a, p, smin, smax, dtype=np.float64)
        q1 = onp.tmid_eval_hoisted(s, a, p, smin, smax, dtype=np.float64)
        np.testing.assert_allclose(q0, q1, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name,kind", [("fkl_update.npz", "fkl"), ("rkl_update.npz", "rkl")])
def test_policy_loss_matches_reference_update(name, kind):
    g = golden(name)
    p = _params(g, "pre_")
    q = onp.tin_eval(g["s"].astype(np.float32), g["grid_a"], p)
    np.testing.assert_allclose(q, g["grid_q"], rtol=2e-5, atol=2e-5)
    if kind == "fkl":
        loss, per_state, boltz, dlogp = onp.fkl_reduce(g["grid_q"], g["grid_w"], g["logp"], float(g["alpha"]))
        assert abs(boltz @ g["grid_w"] - 1).max() < 1e-5
    else:
        loss, per_state, dlogp = onp.rkl_reduce(g["grid_q"], g["v"], g["grid_w"], g["logp"], float(g["alpha"]))
    assert abs(loss - float(g["pi_loss"])) <= 2e-5 * max(1.0, abs(float(g["pi_loss"])))
    # gradient wrt logp by central differences (fp64)
    lp = g["logp"].astype(np.float64)
    eps = 1e-6
    for (b, n) in [(0, 0), (3, 17), (31, 61)]:
        d = np.zeros_like(lp)
        d[b, n] = eps
        if kind == "fkl":
            f = lambda l: onp.fkl_reduce(g["grid_q"], g["grid_w"], l, float(g["alpha"]), dtype=np.float64)[0]
        else:
            f = lambda l: onp.rkl_reduce(g["grid_q"], g["v"], g["grid_w"], l, float(g["alpha"]), dtype=np.float64)[0]
        fd = (f(lp + d) - f(lp - d)) / (2 * eps)
        assert abs(fd - dlogp[b, n]) <= 1e-4 * max(1e-3, abs(fd))


def test_critic_step_matches_reference_adam():
    """a15: one MSE/Adam step reproduces the q_net parameters after the reference update."""
    g = golden("fkl_update.npz")
    pre = [g["pre_" + k].astype(np.float64) for k in ("W1", "b1", "W2", "b2", "W3", "b3")]
    post = [g["post_" + k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")]
    loss, grads = onp.tin_mse_grads(g["s"].astype(np.float32), g["a"].astype(np.float32), g["y"], pre)
    assert abs(loss - float(g["q_loss"])) < 1e-4 * float(g["q_loss"])
    for p0, gr, p1 in zip(pre, grads, post):
        new, _, _ = onp.adam_step_torch(p0, gr, np.zeros_like(p0), np.zeros_like(p0), 1, float(g["lr"]))
        np.testing.assert_allclose(new, p1, rtol=0, atol=2e-6)


def test_mse_grads_finite_difference():
    rng = np.random.RandomState(1)
    S, A, H1, H2, B = 4, 2, 9, 7, 6
    s, a, y = rng.randn(B, S), rng.randn(B, A), rng.randn(B)
    tin = [rng.randn(H1, S + A), rng.randn(H1), rng.randn(H2, H1), rng.randn(H2), rng.randn(1, H2), rng.randn(1)]
    tmid = [rng.randn(S, H1), rng.randn(H1), rng.randn(H1 + A, H2), rng.randn(H2), rng.randn(H2, 1), rng.randn(1)]
    for fn, p in ((onp.tin_mse_grads, tin), (onp.tmid_mse_grads, tmid)):
        loss, grads = fn(s, a, y, p)
        for i in range(6):
            idx = tuple(rng.randint(0, d) for d in p[i].shape)
            pp = [x.copy() for x in p]; pm = [x.copy() for x in p]
            pp[i][idx] += 1e-6; pm[i][idx] -= 1e-6
            fd = (fn(s, a, y, pp)[0] - fn(s, a, y, pm)[0]) / 2e-6
            assert abs(fd - grads[i][idx]) < 1e-5 * max(1.0, abs(fd))


def test_dq_da_finite_difference():
    rng = np.random.RandomState(2)
    S, A, H1, H2, R = 4, 3, 11, 8, 5
    s, a = rng.randn(R, S), rng.randn(R, A)
    tin = [rng.randn(H1, S + A), rng.randn(H1), rng.randn(H2, H1), rng.randn(H2), rng.randn(1, H2), rng.randn(1)]
    tmid = [rng.randn(S, H1), rng.randn(H1), rng.randn(H1 + A, H2), rng.randn(H2), rng.randn(H2, 1), rng.randn(1)]
    g = onp.tin_dq_da(s, a, tin)
    g2 = onp.tmid_dq_da(s, a, tmid)
    for i in range(A):
        d = np.zeros_like(a); d[:, i] = 1e-6
        fd = (onp.tin_forward(s, a + d, *tin, dtype=np.float64) - onp.tin_forward(s, a - d, *tin, dtype=np.float64)) / 2e-6
        np.testing.assert_allclose(g[:, i], fd, rtol=1e-5, atol=1e-6)
        fd2 = (onp.tmid_forward(s, a + d, *tmid, dtype=np.float64) - onp.tmid_forward(s, a - d, *tmid, dtype=np.float64)) / 2e-6
        np.testing.assert_allclose(g2[:, i], fd2, rtol=1e-5, atol=1e-6)


def test_clenshaw_curtis():
    g = golden("cc.npz")
    x, w = onp.clenshaw_curtis(64)
    np.testing.assert_allclose(x, g["x64"], atol=0)
    np.testing.assert_allclose(w, g["w64"], atol=0)
    xf, wf = onp.clenshaw_curtis_fast(64)
    np.testing.assert_allclose(xf, x, atol=1e-15)
    np.testing.assert_allclose(wf, w, atol=1e-14)
    assert abs(w.sum() - 2.0) < 1e-12
    # n-point CC integrates polynomials of degree <= n-1 exactly
    for deg in (0, 2, 10, 40, 62):
        exact = 2.0 / (deg + 1)
        assert abs(np.sum(w * x ** deg) - exact) < 1e-12
    # the reference drops both endpoints: the remaining weights sum to just under 2
    assert abs(w[1:-1].sum() - 1.9995) < 1e-3
    x2, w2 = onp.clenshaw_curtis_fast(1026)
    assert abs(w2[1:-1].sum() - 1.999998) < 1e-5
    acts, ww = onp.intg_grid_1d(64, 2.0)
    assert acts.shape == (62, 1) and ww.shape == (62,) and acts.dtype == np.float32
    assert np.all(np.abs(acts) < 2.0)


def test_gmm_refit_matches_sklearn_bounded():
    g = golden("gmm.npz")
    for c in range(g["X"].shape[0]):
        k, A = int(g["k"][c]), int(g["A"][c])
        X = g["X"][c][:k, :A]
        w, mu, cv, nit = onp.gmm_fit_bounded(X, g["resp0"][c][:k])
        assert nit == int(g["n_iter"][c])
        np.testing.assert_allclose(w, g["weights"][c], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(mu, g["means"][c][:, :A], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cv, g["covs"][c][:, :A], rtol=1e-9, atol=1e-12)
        assert np.all(np.abs(mu) <= 2) and np.all(cv >= np.exp(-2) - 1e-15) and np.all(cv <= np.exp(2) + 1e-15)


def test_gmm_one_component_closed_form():
    rng = np.random.RandomState(3)
    X = rng.randn(6, 4) * 3
    w, mu, cv = onp.gmm_fit_1comp(X)
    w2, mu2, cv2, nit = onp.gmm_fit_bounded(X, np.ones((6, 1)))
    assert nit == 2
    np.testing.assert_allclose(mu, mu2, atol=1e-12)
    np.testing.assert_allclose(cv, cv2, atol=1e-12)
    m0 = X.mean(0)
    np.testing.assert_allclose(mu[0], np.clip(m0, -2, 2), atol=1e-12)
    np.testing.assert_allclose(cv[0], np.clip((X * X).mean(0) - m0 ** 2 + 1e-6, np.exp(-2), np.exp(2)), atol=1e-9)


def test_topk_and_stats_semantics():
    q = np.array([[1., 3., 3., 2., -1.], [0., 0., 0., 0., 0.]], np.float32)
    idx = onp.topk_desc(q, 3)
    assert idx.tolist() == [[2, 1, 3], [4, 3, 2]]       # ties: larger index first
    am, mx, mean = onp.argmax_max_mean(q)
    assert am.tolist() == [1, 0] and mx.tolist() == [3., 0.]
    rng = np.random.RandomState(4)
    q = rng.randn(16, 120).astype(np.float32)
    idx = onp.topk_desc(q, 6)
    for b in range(16):
        assert idx[b].tolist() == list(q[b].argsort()[::-1][:6])     # no ties -> same as reference
    el = onp.gather_elites(rng.randn(16, 120, 2), idx)
    assert el.shape == (16, 6, 2)


def test_sql_soft_value():
    rng = np.random.RandomState(5)
    q = rng.randn(8, 30).astype(np.float32) * 5
    v = onp.sql_soft_value(q, 2)
    ref = np.log(np.exp(q.astype(np.float64)).sum(1)) - np.log(30) + 2 * np.log(2)
    np.testing.assert_allclose(v, ref, rtol=1e-5)


def test_sample_n_k_matches_reference_scheme():
    for n, k in ((10, 5), (1000, 32), (100000, 32), (50, 0)):
        r1 = np.random.RandomState(7)
        idx = onp.sample_n_k(r1, n, k)
        assert len(idx) == k and len(set(idx.tolist())) == k and all(0 <= i < n for i in idx)
        # same stream of draws as the reference algorithm (custom_collections.py:107-131)
        r2 = np.random.RandomState(7)
        if k and 3 * k >= n:
            assert idx.tolist() == r2.choice(n, k, replace=False).tolist()
    with pytest.raises(ValueError):
        onp.sample_n_k(np.random.RandomState(0), 3, 5)


def test_adam_variants_and_soft_update():
    p, g = np.array([1.0, -2.0]), np.array([0.5, -0.25])
    m = v = np.zeros(2)
    pt, _, _ = onp.adam_step_torch(p, g, m, v, 1, 1e-3)
    ptf, _, _ = onp.adam_step_tf(p, g, m, v, 1, 1e-3)
    np.testing.assert_allclose(pt, p - 1e-3 * np.sign(g), atol=1e-9)
    np.testing.assert_allclose(ptf, pt, atol=1e-9)          # differ only through eps placement
    assert not np.array_equal(pt, ptf)
    t = onp.soft_update(np.zeros(2), np.ones(2), 0.01)
    np.testing.assert_allclose(t, 0.01)


def test_cem_runs_and_respects_bounds():
    rng = np.random.RandomState(6)
    S, A, H1, H2, B, N = 3, 2, 16, 12, 4, 64
    p = [rng.randn(S, H1) * .5, rng.randn(H1) * .1, rng.randn(H1 + A, H2) * .5, rng.randn(H2) * .1,
         rng.randn(H2, 1), rng.randn(1)]
    s = rng.randn(B, S)
    u0 = rng.uniform(size=(B, N, A))
    noise = rng.randn(2, B, N, A)
    cu = rng.uniform(size=(2, B, N))
    qf = lambda st, ac: onp.tmid_eval(st, ac, p)
    for M in (1, 2):
        W, Mu, Cv, idx = onp.cem_iterate(qf, s, u0, noise, cu, 6, M, -np.ones(A), np.ones(A))
        assert W.shape == (B, M) and np.allclose(W.sum(1), 1, atol=1e-9)
        assert np.all(np.abs(Mu) <= 2) and np.all(Cv >= np.exp(-2) - 1e-12) and np.all(Cv <= np.exp(2) + 1e-12)
        assert idx.shape == (3, B, 6)
        act = onp.cem_final_action(W, Mu)
        assert act.shape == (B, A)


# ----------------------------------------------------------------------------- ports timed by bench.py
@pytest.mark.parametrize("name", ["tin_cfg1.npz", "tin_400_300.npz"])
def test_torch_port_matches_reference_golden(name):
    """oracle_torch.SoftQNetworkPort (the CPU arm bench.py times) == the reference's SoftQNetwork."""
    import torch
    from oracle import oracle_torch as ot
    g = golden(name)
    net = ot.SoftQNetworkPort(*_params(g))
    s, a = torch.as_tensor(g["s"]), torch.as_tensor(g["a"])
    B, N = s.shape[0], a.shape[0]
    q = net(s.unsqueeze(1).repeat(1, N, 1).reshape(-1, s.shape[1]), a.repeat(B, 1, 1).reshape(-1, a.shape[1]))
    np.testing.assert_allclose(q.reshape(B, N).numpy(), g["q"], rtol=2e-5, atol=2e-6)


def test_torch_port_fkl_step_matches_reference_update():
    import torch
    from oracle import oracle_torch as ot
    g = golden("fkl_update.npz")
    net = ot.SoftQNetworkPort(*_params(g, "pre_"))
    t = lambda x: torch.as_tensor(np.asarray(x, np.float32))
    loss_b, q = ot.fkl_sampled_step(net, t(g["s"]), t(g["grid_a"]), t(g["grid_w"]), t(g["logp"]), float(g["alpha"]))
    np.testing.assert_allclose(q.numpy(), g["grid_q"], rtol=2e-5, atol=2e-5)
    assert abs(float(loss_b.mean()) - float(g["pi_loss"])) <= 2e-5 * max(1.0, abs(float(g["pi_loss"])))


def test_rounded_operand_oracle():
    """tin_eval_rounded: exact when the operands are representable; otherwise within the stated
    precision budget of the operand type (same figures test_gpu_parity gates the kernel on)."""
    from conftest import rel_err
    g = golden("tin_400_300.npz")
    p = _params(g)
    exact = onp.tin_eval(g["s"], g["a"], p, dtype=np.float64)
    for kind, (rms_b, max_b) in {"fp16": (2.5e-3, 2e-2), "bf16": (2e-2, 1.5e-1)}.items():
        e = rel_err(onp.tin_eval_rounded(g["s"], g["a"], p, kind), exact)
        assert 1e-6 < np.sqrt((e ** 2).mean()) < rms_b and e.max() < max_b
    # representable operands -> no rounding anywhere -> equals the exact path
    rng = np.random.RandomState(0)
    S, A, H1, H2 = 3, 2, 8, 8
    ints = lambda *sh: rng.randint(-2, 3, sh).astype(np.float32)
    p = [ints(H1, S + A), ints(H1), ints(H2, H1), ints(H2), ints(1, H2), ints(1)]
    s, a = ints(4, S), ints(5, A)
    np.testing.assert_array_equal(onp.tin_eval_rounded(s, a, p, "bf16"), onp.tin_eval(s, a, p, dtype=np.float64))
    x = np.array([1.0, 1.0 + 2 ** -11, 1.0 + 3 * 2 ** -11, 70000.0, -1e-8], np.float32)
    np.testing.assert_array_equal(onp.round_operand(x, "fp16"), [1.0, 1.0, 1.0 + 2 ** -9, 65504.0, -0.0])


@pytest.mark.parametrize("A", [1, 2, 3])
def test_policy_logprob_matches_reference(A):
    """oracle.tanh_gauss_logprob == the reference's own PolicyNetwork.get_logprob and its autograd
    (golden policy_logp.npz; A>1 pins the diag_embed(std)-as-covariance quirk)."""
    g = golden("policy_logp.npz")
    k = lambda n: g[f"A{A}_{n}"]
    lp, dm, ds = onp.tanh_gauss_logprob(k("mean"), k("log_std"), k("actions"), float(k("scale")))
    np.testing.assert_allclose(lp, k("logp"), rtol=2e-5, atol=2e-5)
    c = k("c").astype(np.float64)
    np.testing.assert_allclose((c[:, :, None] * dm).sum(1), k("dmean"), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose((c[:, :, None] * ds).sum(1), k("dlog_std"), rtol=1e-4, atol=1e-4)
    if A == 1:
        lp1 = onp.tanh_gauss_logprob_1d(k("mean"), k("log_std"), k("actions"), float(k("scale")))
        np.testing.assert_allclose(lp1, k("logp"), rtol=2e-5, atol=2e-5)


def test_fkl_policy_reduce_consistent_with_golden_update():
    """With logp taken from the reference's update (fkl_update.npz) the fused reduction equals
    fkl_reduce; with logp recomputed from (mean, log_std) it equals the composition."""
    g = golden("policy_logp.npz")
    rng = np.random.RandomState(3)
    B, N = g["A2_logp"].shape
    q, w = rng.randn(B, N), rng.uniform(0.01, 0.1, N)
    per, dmean, dls, lp = onp.fkl_policy_reduce(q, w, g["A2_actions"], g["A2_mean"], g["A2_log_std"], 2.0, 0.3)
    _, per2, _, dlogp = onp.fkl_reduce(q, w, g["A2_logp"].astype(np.float64), 0.3, dtype=np.float64)
    np.testing.assert_allclose(per, per2, rtol=1e-4, atol=1e-5)
    # finite-difference check of dL/dmean
    eps = 1e-6
    m = g["A2_mean"].astype(np.float64)
    m2 = m.copy(); m2[1, 0] += eps
    f = lambda mm: onp.fkl_policy_reduce(q, w, g["A2_actions"], mm, g["A2_log_std"], 2.0, 0.3)[0].mean()
    assert abs((f(m2) - f(m)) / eps - dmean[1, 0]) < 1e-4 * max(1.0, abs(dmean[1, 0]))


@pytest.mark.parametrize("A", [1, 2, 3])
def test_torch_port_get_logprob_matches_reference(A):
    import torch
    from oracle import oracle_torch as ot
    g = golden("policy_logp.npz")
    k = lambda n: torch.as_tensor(g[f"A{A}_{n}"])
    lp = ot.get_logprob_port(k("mean"), k("log_std"), k("actions"), float(g[f"A{A}_scale"]))
    np.testing.assert_allclose(lp.numpy(), g[f"A{A}_logp"], rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# Actor-Expert actor side (N1): the sampling restatement is pinned on numpy's own RandomState
# ---------------------------------------------------------------------------------------------
def test_mixture_sample_reproduces_numpy_choice_and_normal():
    """ae_network.py:483-488 calls rng.choice(M, N, p=alpha_b) then rng.normal(m[idx], s[idx]) per state; feeding
    the uniforms / normals numpy itself would draw (same RandomState) reproduces its output bit for bit."""
    B, M, A, N = 4, 3, 2, 60
    r = np.random.RandomState(7)
    alpha = r.dirichlet(np.ones(M), B)
    mean = r.randn(B, M, A).astype(np.float32)
    sig = (0.2 + r.rand(B, M, A)).astype(np.float32)
    rng1, rng2 = np.random.RandomState(5), np.random.RandomState(5)
    ref, ref_idx, us, zs = [], [], [], []
    for b in range(B):
        idx = rng1.choice(M, N, p=alpha[b])
        ref_idx.append(idx)
        ref.append(np.clip(rng1.normal(mean[b][idx], sig[b][idx]), -1.0, 1.0))
        us.append(rng2.random_sample(N))
        zs.append(rng2.standard_normal((N, A)))
    act, idx = onp.mixture_sample(alpha, mean, sig, np.array(us), np.array(zs), [-1, -1], [1, 1])
    np.testing.assert_array_equal(idx, np.array(ref_idx))
    np.testing.assert_array_equal(act, np.array(ref))
    # uniform replacement of the first samples (ae_network.py:489-493): low + (high-low)*u
    uu = r.rand(B, 5, A)
    act2, _ = onp.mixture_sample(alpha, mean, sig, np.array(us), np.array(zs), [-1, -2], [1, 2], uni_u=uu)
    np.testing.assert_allclose(act2[:, :5], np.array([-1, -2]) + np.array([2, 4]) * uu)
    eq = onp.mixture_pick(alpha, np.array(us), equal_modal=True)
    assert eq.min() >= 0 and eq.max() <= M - 1


def test_mixture_nll_gradients_by_finite_differences():
    r = np.random.RandomState(3)
    B, M, A, k = 3, 2, 2, 5
    alpha = r.dirichlet(np.ones(M), B)
    mean, sig = r.randn(B, M, A) * 0.5, 0.3 + r.rand(B, M, A)
    y = r.uniform(-1, 1, (B, k, A))
    loss, nll, da, dm, ds = onp.mixture_nll(alpha, mean, sig, y)
    assert nll.shape == (B, k) and abs(loss - nll.mean()) < 1e-12
    eps = 1e-6
    for arr, grad, which in ((alpha, da, 0), (mean, dm, 1), (sig, ds, 2)):
        it = np.nditer(arr, flags=["multi_index"])
        for _ in it:
            pert = [alpha.copy(), mean.copy(), sig.copy()]
            pert[which][it.multi_index] += eps
            num = (onp.mixture_nll(*pert, y)[0] - loss) / eps
            assert abs(num - grad[it.multi_index]) < 1e-4 * max(1.0, abs(num))
    # the single-Gaussian closed form: -log N(y; m, s)
    l1 = onp.mixture_nll(np.ones((1, 1)), [[[0.3]]], [[[0.5]]], [[[0.1]]])[0]
    assert abs(l1 - (0.5 * np.log(2 * np.pi * 0.25) + 0.04 / 0.5)) < 1e-12
    # underflow is clipped at 1e-30 and carries no gradient (tf.clip_by_value, ae_network.py:276)
    lo = onp.mixture_nll(np.ones((1, 1)), [[[0.0]]], [[[0.01]]], [[[5.0]]])
    assert abs(lo[0] - (-np.log(1e-30))) < 1e-9 and np.all(lo[3] == 0)


def test_svgd_kernel_restatement_properties():
    """utils/sql_kernel.py: kappa in (0,1], kappa = 1 on coincident particles, the gradient is the analytic
    derivative of kappa wrt xs at fixed bandwidth, and the bandwidth is median/log(Kx) floored at h_min."""
    r = np.random.RandomState(0)
    xs, ys = r.uniform(-1, 1, (3, 5, 2)), r.uniform(-1, 1, (3, 4, 2))
    kap, grad, h = onp.adaptive_isotropic_gaussian_kernel(xs, ys)
    assert kap.shape == (3, 5, 4) and grad.shape == (3, 5, 4, 2) and np.all((kap > 0) & (kap <= 1))
    d = ((xs[:, :, None] - ys[:, None]) ** 2).sum(-1).reshape(3, -1)
    np.testing.assert_allclose(h, np.sort(d, 1)[:, ::-1][:, 10] / np.log(5))      # 20 pairs: the 11th largest
    e = 1e-6
    x2 = xs.copy()
    x2[1, 2, 0] += e
    d2 = ((x2[:, :, None] - ys[:, None]) ** 2).sum(-1)
    num = (np.exp(-d2[1, 2] / h[1]) - kap[1, 2]) / e
    np.testing.assert_allclose(num, grad[1, 2, :, 0], rtol=1e-4, atol=1e-6)
    twin = np.repeat(xs[:, :1], 2, axis=1)
    same = onp.adaptive_isotropic_gaussian_kernel(twin, twin, h_min=0.5)
    assert np.all(same[0] == 1.0) and np.all(same[2] == 0.5)                      # zero distances -> h_min floor
    g, _, _ = onp.svgd_action_gradients(np.zeros((3, 5, 2)), xs * 0.5, ys * 0.5)
    assert g.shape == (3, 4, 2) and np.all(np.isfinite(g))


def test_cem_forced_replay_equals_free_run_on_its_own_elites():
    """cem_iterate_forced fed with cem_iterate's own elite indices reproduces it (mixture and per-iteration top-m)."""
    rng = np.random.RandomState(3)
    S, A, H1, H2, B, N, iters, M = 3, 2, 24, 16, 6, 64, 3, 2
    k1, k2 = np.sqrt(3 / S), np.sqrt(3 / (H1 + A))
    p = [rng.uniform(-k1, k1, (S, H1)).astype(np.float32), rng.uniform(-k1, k1, H1).astype(np.float32),
         rng.uniform(-k2, k2, (H1 + A, H2)).astype(np.float32), rng.uniform(-k2, k2, H2).astype(np.float32),
         rng.uniform(-1, 1, (H2, 1)).astype(np.float32), rng.uniform(-1, 1, 1).astype(np.float32)]
    s = rng.randn(B, S).astype(np.float32)
    u0 = rng.uniform(size=(B, N, A)).astype(np.float32)
    noise = rng.randn(iters - 1, B, N, A).astype(np.float32)
    cu = rng.uniform(size=(iters - 1, B, N)).astype(np.float32)
    qf = lambda st, ac: onp.tmid_eval_hoisted(st, ac, p, dtype=np.float64)
    W, Mu, Cv, idx = onp.cem_iterate(qf, s, u0, noise, cu, 6, M, -np.ones(A), np.ones(A))
    Q, own, W2, Mu2, Cv2 = onp.cem_iterate_forced(qf, s, u0, noise, cu, 6, M, -np.ones(A), np.ones(A), idx)
    assert Q.shape == (iters, B, N)
    np.testing.assert_array_equal(own, idx)
    np.testing.assert_array_equal(W2, W)
    np.testing.assert_array_equal(Mu2, Mu)
    np.testing.assert_array_equal(Cv2, Cv)


def test_reference_class_cpu_arm_equals_port():
    """bench.py's CPU arm: the reference's own SoftQNetwork (unmodified copy under oracle/_ref/, made by build()) and the
    line-by-line port produce the same Q; skipped where the copy does not exist."""
    import torch
    from oracle import oracle_torch as ot, ref_loader
    rng = np.random.RandomState(0)
    S, A, H1, H2 = 5, 2, 32, 24
    p = [rng.randn(H1, S + A).astype(np.float32), rng.randn(H1).astype(np.float32), rng.randn(H2, H1).astype(np.float32),
         rng.randn(H2).astype(np.float32), rng.randn(1, H2).astype(np.float32), rng.randn(1).astype(np.float32)]
    net = ref_loader.reference_softq(p)
    if net is None:
        pytest.skip("oracle/_ref/ not built (no /root/reference at build time)")
    s, a = torch.randn(40, S), torch.randn(40, A)
    with torch.no_grad():
        np.testing.assert_array_equal(net(s, a).numpy(), ot.SoftQNetworkPort(*p)(s, a).numpy())
