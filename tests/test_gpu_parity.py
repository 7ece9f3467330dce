"""GPU parity tests: every call goes through the C-ABI (librlc.so) and is compared with the CPU
oracle on the same seeded inputs, plus the committed golden vectors from the reference."""
import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu

# Tolerances (stated per north_star): fp32 CUDA-core path ~1e-5; fp16-operand tensor path 1e-3 on
# the metric |dq| / max(|q|, rms_state(q)); bf16 operands are reported, not gated at 1e-3.
TOL_FP32 = 2e-5
TOL_FP16 = 1e-3
TOL_BF16 = 2e-2


def _p(g, prefix=""):
    return [g[prefix + k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")]


def _tin(eng, g_or_params, S, A, H1, H2):
    import rlcontrol_b200 as rb
    cr = rb.Critic(eng, rb.TIN, S, A, H1, H2)
    cr.load(*g_or_params, rb.LAYOUT_OUT_IN)
    return cr


def _rand_tin(rng, S, A, H1, H2, last=1.0):
    k1, k2, k3 = 1 / np.sqrt(S + A), 1 / np.sqrt(H1), last / np.sqrt(H2)
    return [rng.uniform(-k1, k1, (H1, S + A)).astype(np.float32), rng.uniform(-k1, k1, H1).astype(np.float32),
            rng.uniform(-k2, k2, (H2, H1)).astype(np.float32), rng.uniform(-k2, k2, H2).astype(np.float32),
            rng.uniform(-k3, k3, (1, H2)).astype(np.float32), rng.uniform(-k3, k3, 1).astype(np.float32)]


def _rand_tmid(rng, S, A, H1, H2, last=1.0):
    k1, k2, k3 = np.sqrt(3 / S), np.sqrt(3 / (H1 + A)), last
    return [rng.uniform(-k1, k1, (S, H1)).astype(np.float32), rng.uniform(-k1, k1, H1).astype(np.float32),
            rng.uniform(-k2, k2, (H1 + A, H2)).astype(np.float32), rng.uniform(-k2, k2, H2).astype(np.float32),
            rng.uniform(-k3, k3, (H2, 1)).astype(np.float32), rng.uniform(-k3, k3, 1).astype(np.float32)]


# ----------------------------------------------------------------------------- a1/a2: T-in eval
@pytest.mark.parametrize("name,dims", [("tin_cfg1.npz", (3, 1, 200, 200)),
                                       ("tin_400_300.npz", (17, 6, 400, 300)),
                                       ("tin_cfg4_exact.npz", (3, 1, 400, 300))])
@pytest.mark.parametrize("prec,tol", [("fp32", TOL_FP32), ("fp16", TOL_FP16), ("bf16", TOL_BF16)])
def test_tin_eval_golden(eng, name, dims, prec, tol):
    g = golden(name)
    cr = _tin(eng, _p(g), *dims)
    q = cr.eval(g["s"], g["a"], prec).cpu().numpy()
    assert eng.umma_error() == 0
    err = rel_err(q, g["q"]).max()
    assert err < tol, f"{name} {prec}: rel err {err:.3e}"


@pytest.mark.parametrize("prec,tol", [("fp32", TOL_FP32), ("fp16", TOL_FP16)])
@pytest.mark.parametrize("B,N,per_state", [(1, 1, False), (3, 130, True), (5, 257, False), (2, 1000, True),
                                           (64, 62, False), (7, 128, True)])
def test_tin_eval_ragged_shapes(eng, prec, tol, B, N, per_state):
    rng = np.random.RandomState(B * 1000 + N)
    S, A, H1, H2 = 5, 2, 72, 40
    p = _rand_tin(rng, S, A, H1, H2, last=3.0)
    cr = _tin(eng, p, S, A, H1, H2)
    s = rng.randn(B, S).astype(np.float32)
    a = rng.uniform(-1, 1, (B, N, A) if per_state else (N, A)).astype(np.float32)
    ref = onp.tin_eval(s, a, p, dtype=np.float64)
    q = cr.eval(s, a, prec).cpu().numpy()
    assert eng.umma_error() == 0
    assert rel_err(q, ref).max() < tol


def test_tin_eval_empty(eng):
    rng = np.random.RandomState(0)
    p = _rand_tin(rng, 3, 1, 16, 16)
    cr = _tin(eng, p, 3, 1, 16, 16)
    q = cr.eval(np.zeros((0, 3), np.float32), np.zeros((5, 1), np.float32), "fp32")
    assert tuple(q.shape) == (0, 5)
    q = cr.eval(np.zeros((4, 3), np.float32), np.zeros((0, 1), np.float32), "fp32")
    assert tuple(q.shape) == (4, 0)
    with pytest.raises(ValueError):
        cr.eval(np.zeros((4, 2), np.float32), np.zeros((5, 1), np.float32))


def test_tin_eval_tensor_path_size_independent_properties(eng):
    """cfg4-sized (B=512 slice of it) checks that do not need the CPU oracle at full size:
    (i) shared-grid result == per-state result on the tiled grid; (ii) permuting states permutes
    rows; (iii) a random sample of rows matches the fp64 oracle within 1e-3."""
    import torch
    rng = np.random.RandomState(11)
    S, A, H1, H2, B, N = 17, 6, 400, 300, 512, 1024
    p = _rand_tin(rng, S, A, H1, H2, last=30.0)
    cr = _tin(eng, p, S, A, H1, H2)
    s = np.clip(rng.randn(B, S), -10, 10).astype(np.float32)
    a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
    q = cr.eval(s, a, "fp16")
    q_ps = cr.eval(s, np.broadcast_to(a, (B, N, A)).copy(), "fp16")
    assert eng.umma_error() == 0
    assert torch.equal(q, q_ps)
    perm = rng.permutation(B)
    q_perm = cr.eval(s[perm], a, "fp16")
    assert torch.equal(q_perm, q[torch.as_tensor(perm, device=q.device)])
    rows = rng.choice(B, 24, replace=False)
    ref = onp.tin_eval(s[rows], a, p, dtype=np.float64)
    assert rel_err(q.cpu().numpy()[rows], ref).max() < TOL_FP16
    # argmax / elite agreement wherever the Q gap exceeds the tolerance
    qn = q.cpu().numpy()[rows]
    for i in range(len(rows)):
        order = np.argsort(-ref[i])
        gap = ref[i][order[0]] - ref[i][order[1]]
        if gap > 2 * TOL_FP16 * max(np.abs(ref[i]).max(), np.sqrt(np.mean(ref[i] ** 2))):
            assert int(np.argmax(qn[i])) == int(order[0])


# ----------------------------------------------------------------------------- a6/a7: T-mid eval
def test_tmid_eval_trueq_checkpoints(eng):
    import rlcontrol_b200 as rb
    g = golden("trueq.npz")
    grid = g["grid"]
    for v in ("eq_var1", "eq_var2", "eq_var3", "uneq_var1", "uneq_var2"):
        p = [g[f"{v}_{k}"] for k in ("W1", "b1", "W2", "b2", "W3", "b3")]
        cr = rb.Critic(eng, rb.TMID, 1, 1, 200, 200)
        cr.load(*p, rb.LAYOUT_IN_OUT)
        q = cr.eval(np.zeros((3, 1), np.float32), grid[:, None], "fp32").cpu().numpy()
        ref = onp.tmid_eval(np.zeros((3, 1), np.float32), grid[:, None], p)
        np.testing.assert_allclose(q, ref, rtol=1e-5, atol=1e-5)
        assert np.abs(q[0] - g[f"{v}_reward"]).max() < 0.11


@pytest.mark.parametrize("S,A,H1,H2,B,N", [(1, 1, 200, 200, 32, 120), (17, 6, 400, 300, 16, 200),
                                           (3, 2, 50, 33, 5, 77), (4, 12, 64, 64, 3, 40)])
@pytest.mark.parametrize("per_state", [True, False])
def test_tmid_eval_matches_oracle(eng, S, A, H1, H2, B, N, per_state):
    import rlcontrol_b200 as rb
    rng = np.random.RandomState(S * 100 + A)
    p = _rand_tmid(rng, S, A, H1, H2)
    smin, smax = -np.ones(S) * 1.5, np.ones(S) * 1.5
    cr = rb.Critic(eng, rb.TMID, S, A, H1, H2, smin, smax)
    cr.load(*p, rb.LAYOUT_IN_OUT)
    s = (rng.randn(B, S) * 2).astype(np.float32)           # some states get clipped
    a = rng.uniform(-1, 1, (B, N, A) if per_state else (N, A)).astype(np.float32)
    ref = onp.tmid_eval(s, a, p, smin, smax, dtype=np.float64)
    q = cr.eval(s, a, "fp32").cpu().numpy()
    assert rel_err(q, ref).max() < TOL_FP32
    qg, g = cr.eval_grad(s, a)
    np.testing.assert_allclose(qg.cpu().numpy(), q, rtol=1e-6, atol=1e-6)
    gref = onp.tmid_dq_da(onp.stack_state_major(s, N), onp.stack_actions(a, B), p, smin, smax)
    np.testing.assert_allclose(g.cpu().numpy().reshape(-1, A), gref, rtol=1e-4, atol=1e-5)


def test_pack_unpack_roundtrip(eng):
    import rlcontrol_b200 as rb
    rng = np.random.RandomState(3)
    for topo, layout, mk in ((rb.TIN, rb.LAYOUT_OUT_IN, _rand_tin), (rb.TMID, rb.LAYOUT_IN_OUT, _rand_tmid)):
        p = mk(rng, 4, 3, 20, 12)
        cr = rb.Critic(eng, topo, 4, 3, 20, 12)
        cr.load(*p, layout)
        out = cr.export(layout)
        for x, y in zip(p, out):
            assert np.array_equal(np.asarray(x).reshape(-1), y.cpu().numpy().reshape(-1))


# ----------------------------------------------------------------------------- a8/a9/a10 reductions
@pytest.mark.parametrize("B,N,k", [(32, 120, 6), (256, 1024, 6), (3, 7, 7), (5, 1000, 64), (1, 1, 1)])
def test_topk_bit_exact(eng, B, N, k):
    import torch
    rng = np.random.RandomState(N)
    q = rng.randn(B, N).astype(np.float32)
    q[:, : N // 3] = np.round(q[:, : N // 3], 1)            # force ties
    acts = rng.randn(B, N, 3).astype(np.float32)
    dev = eng.device
    idx, qs, el = eng.topk(torch.as_tensor(q, device=dev), k, torch.as_tensor(acts, device=dev))
    ref = onp.topk_desc(q, k)
    assert np.array_equal(idx.cpu().numpy(), ref)
    assert np.array_equal(qs.cpu().numpy(), np.take_along_axis(q, ref, 1))
    assert np.array_equal(el.cpu().numpy(), onp.gather_elites(acts, ref))


def test_stats_and_lse(eng):
    import torch
    rng = np.random.RandomState(9)
    q = (rng.randn(37, 301) * 4).astype(np.float32)
    q[3, 5] = q[3, 200] = q[3].max() + 1                    # tie -> first index
    am, mx, mean = eng.stats(torch.as_tensor(q, device=eng.device))
    ram, rmx, rmean = onp.argmax_max_mean(q)
    assert np.array_equal(am.cpu().numpy(), ram) and np.array_equal(mx.cpu().numpy(), rmx)
    np.testing.assert_allclose(mean.cpu().numpy(), rmean, rtol=1e-5, atol=1e-6)
    v = eng.soft_value(torch.as_tensor(q, device=eng.device), 2).cpu().numpy()
    np.testing.assert_allclose(v, onp.sql_soft_value(q, 2), rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------------------- a3/a4 FKL / RKL
@pytest.mark.parametrize("name,kind", [("fkl_update.npz", "fkl"), ("rkl_update.npz", "rkl")])
def test_policy_reductions_match_reference(eng, name, kind):
    import torch
    g = golden(name)
    dev = eng.device
    t = lambda x: torch.as_tensor(np.asarray(x, np.float32), device=dev)
    # grid q through the CUDA critic on the reference's pre-update weights
    cr = _tin(eng, _p(g, "pre_"), 3, 1, 200, 200)
    q = cr.eval(g["s"], g["grid_a"], "fp32")
    assert rel_err(q.cpu().numpy(), g["grid_q"]).max() < TOL_FP32
    alpha = float(g["alpha"])
    if kind == "fkl":
        loss_b, boltz, dlogp = eng.fkl(q, t(g["grid_w"]), t(g["logp"]), alpha)
        rl, rper, rboltz, rd = onp.fkl_reduce(g["grid_q"], g["grid_w"], g["logp"], alpha, dtype=np.float64)
        np.testing.assert_allclose(boltz.cpu().numpy(), rboltz, rtol=2e-3, atol=1e-7)
    else:
        loss_b, dlogp = eng.rkl(q, t(g["v"]), t(g["grid_w"]), t(g["logp"]), alpha)
        rl, rper, rd = onp.rkl_reduce(g["grid_q"], g["v"], g["grid_w"], g["logp"], alpha, dtype=np.float64)
    loss = float(loss_b.mean())
    assert abs(loss - float(g["pi_loss"])) < 1e-4 * max(1.0, abs(float(g["pi_loss"])))
    np.testing.assert_allclose(loss_b.cpu().numpy(), rper, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(dlogp.cpu().numpy(), rd, rtol=2e-3, atol=1e-7)


def test_rkl_hard_and_sharded_mean(eng):
    import torch
    rng = np.random.RandomState(4)
    B, N = 16, 62
    q, v = rng.randn(B, N).astype(np.float32), rng.randn(B).astype(np.float32)
    _, w = onp.intg_grid_1d(64, 1.0)
    lp = (rng.randn(B, N) - 1).astype(np.float32)
    t = lambda x: torch.as_tensor(x, device=eng.device)
    lb, d = eng.rkl(t(q), t(v), t(w), t(lp), 0.3, hard=True, b_total=4 * B)
    rl, rper, rd = onp.rkl_reduce(q, v, w, lp, 0.3, hard=True, dtype=np.float64)
    np.testing.assert_allclose(lb.cpu().numpy(), rper, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(d.cpu().numpy(), rd / 4, rtol=1e-4, atol=1e-7)


# ----------------------------------------------------------------------------- a11/a12 CEM
def test_gmm_refit_golden(eng):
    import torch
    g = golden("gmm.npz")
    for c in range(g["X"].shape[0]):
        k, A = int(g["k"][c]), int(g["A"][c])
        X = torch.as_tensor(g["X"][c][None, :k, :A].astype(np.float32), device=eng.device).contiguous()
        r0 = torch.as_tensor(g["resp0"][c][None, :k].astype(np.float32), device=eng.device).contiguous()
        w, mu, var, nit = eng.gmm_refit(X, 2, r0)
        # fp32 I/O of fp64 sklearn data: compare against the oracle on the same fp32 inputs
        ow, om, oc, on = onp.gmm_fit_bounded(X.cpu().numpy()[0].astype(np.float64), g["resp0"][c][:k])
        assert int(nit[0]) == on
        np.testing.assert_allclose(w.cpu().numpy()[0], ow, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(mu.cpu().numpy()[0], om, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(var.cpu().numpy()[0], oc, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(w.cpu().numpy()[0], g["weights"][c], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("M", [1, 2])
@pytest.mark.parametrize("S,A,H1,H2,B,N,iters", [(1, 1, 200, 200, 32, 120, 2), (17, 6, 400, 300, 24, 1024, 3)])
def test_cem_matches_oracle(eng, M, S, A, H1, H2, B, N, iters):
    import rlcontrol_b200 as rb
    rng = np.random.RandomState(5 + M)
    p = _rand_tmid(rng, S, A, H1, H2, last=1.0)
    smin, smax = -np.ones(S) * 10, np.ones(S) * 10
    cr = rb.Critic(eng, rb.TMID, S, A, H1, H2, smin, smax)
    cr.load(*p, rb.LAYOUT_IN_OUT)
    s = rng.randn(B, S).astype(np.float32)
    u0 = rng.uniform(size=(B, N, A)).astype(np.float32)
    noise = rng.randn(iters - 1, B, N, A).astype(np.float32)
    cu = rng.uniform(size=(iters - 1, B, N)).astype(np.float32)
    qf = lambda st, ac: onp.tmid_eval(st, ac, p, smin, smax)
    W, Mu, Cv, idx = onp.cem_iterate(qf, s, u0, noise, cu, 6, M, -np.ones(A), np.ones(A))
    w, mu, var, best, gidx = cr.cem(s, u0, noise, cu, 6, M, -np.ones(A), np.ones(A), want_idx=True)
    gidx = gidx.cpu().numpy()
    # first-iteration elites are a pure function of identical inputs: bit-exact unless two
    # candidates tie within fp32 summation noise
    same0 = (gidx[0] == idx[0]).all(axis=1)
    assert same0.mean() > 0.9
    ok = np.array([(gidx[:, b] == idx[:, b]).all() for b in range(B)])
    assert ok.mean() > 0.7                      # later iterations diverge only after a near-tie
    np.testing.assert_allclose(w.cpu().numpy()[ok], W[ok], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(mu.cpu().numpy()[ok], Mu[ok], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(var.cpu().numpy()[ok], Cv[ok], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(best.cpu().numpy()[ok], onp.cem_final_action(W, Mu)[ok], rtol=1e-4, atol=1e-5)
    assert np.all(np.abs(mu.cpu().numpy()) <= 2) and np.all(var.cpu().numpy() >= np.exp(-2) - 1e-6)


# ----------------------------------------------------------------------------- a14/a15/a16 backward
@pytest.mark.parametrize("topo", ["tin", "tmid"])
def test_grad_action_matches_oracle(eng, topo):
    import rlcontrol_b200 as rb
    rng = np.random.RandomState(8)
    S, A, H1, H2, R = 6, 3, 90, 70, 333
    s, a = rng.randn(R, S).astype(np.float32), rng.uniform(-1, 1, (R, A)).astype(np.float32)
    if topo == "tin":
        p = _rand_tin(rng, S, A, H1, H2)
        cr = _tin(eng, p, S, A, H1, H2)
        gref, qref = onp.tin_dq_da(s, a, p), onp.tin_forward(s, a, *p, dtype=np.float64)
    else:
        p = _rand_tmid(rng, S, A, H1, H2)
        cr = rb.Critic(eng, rb.TMID, S, A, H1, H2)
        cr.load(*p, rb.LAYOUT_IN_OUT)
        gref, qref = onp.tmid_dq_da(s, a, p), onp.tmid_forward(s, a, *p, dtype=np.float64)
    g, q = cr.grad_action(s, a)
    np.testing.assert_allclose(q.cpu().numpy(), qref, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(g.cpu().numpy(), gref, rtol=1e-3, atol=1e-5)


def test_critic_step_matches_reference_update(eng):
    """a15/a16: rlc_critic_grads + rlc_adam_step reproduce the q_net parameters after the
    reference's own update_network (golden, torch Adam)."""
    import rlcontrol_b200 as rb
    g = golden("fkl_update.npz")
    cr = _tin(eng, _p(g, "pre_"), 3, 1, 200, 200)
    opt = rb.CriticOptimizer(cr, lr=float(g["lr"]), variant=rb.ADAM_TORCH)
    loss, q = opt.step(g["s"].astype(np.float32), g["a"].astype(np.float32), g["y"])
    assert abs(float(loss) - float(g["q_loss"])) < 1e-4 * float(g["q_loss"])
    np.testing.assert_allclose(q.cpu().numpy(), g["q_reg"], rtol=1e-4, atol=1e-4)
    out = cr.export(rb.LAYOUT_OUT_IN)
    for x, k in zip(out, ("W1", "b1", "W2", "b2", "W3", "b3")):
        np.testing.assert_allclose(x.cpu().numpy().reshape(g["post_" + k].shape), g["post_" + k],
                                   rtol=0, atol=5e-6)


@pytest.mark.parametrize("topo", ["tin", "tmid"])
def test_critic_grads_and_adam_variants(eng, topo):
    import rlcontrol_b200 as rb
    rng = np.random.RandomState(12)
    S, A, H1, H2, B = 5, 2, 48, 36, 96
    s, a, y = rng.randn(B, S).astype(np.float32), rng.randn(B, A).astype(np.float32), rng.randn(B).astype(np.float32)
    if topo == "tin":
        p = _rand_tin(rng, S, A, H1, H2)
        cr = _tin(eng, p, S, A, H1, H2)
        loss, grads = onp.tin_mse_grads(s, a, y, p)
        # oracle grads are in torch layout [out,in]; canonical theta is [in,out]
        flat = np.concatenate([grads[0].T.ravel(), grads[1], grads[2].T.ravel(), grads[3], grads[4].ravel(), grads[5]])
        variant = rb.ADAM_TORCH
    else:
        p = _rand_tmid(rng, S, A, H1, H2)
        cr = rb.Critic(eng, rb.TMID, S, A, H1, H2)
        cr.load(*p, rb.LAYOUT_IN_OUT)
        loss, grads = onp.tmid_mse_grads(s, a, y, p)
        flat = np.concatenate([x.ravel() for x in grads])
        variant = rb.ADAM_TF
    grad, gl, q = cr.grads(s, a, y)
    assert abs(float(gl) - loss) < 1e-4 * max(1, loss)
    np.testing.assert_allclose(grad.cpu().numpy(), flat, rtol=2e-3, atol=2e-5)
    # sharded mean: two half-batches with b_total=B sum to the full gradient
    g1, _, _ = cr.grads(s[: B // 2], a[: B // 2], y[: B // 2], b_total=B)
    g2, _, _ = cr.grads(s[B // 2:], a[B // 2:], y[B // 2:], b_total=B)
    np.testing.assert_allclose((g1 + g2).cpu().numpy(), flat, rtol=2e-3, atol=2e-5)
    # three optimiser steps against the oracle's Adam, with a Polyak target
    import torch
    theta0 = cr.theta.clone().cpu().numpy().astype(np.float64)
    tgt = rb.Critic(eng, cr.topology, S, A, H1, H2)
    tgt.copy_from(cr)
    opt = rb.CriticOptimizer(cr, lr=1e-2, variant=variant, target=tgt, tau=0.01)
    m = np.zeros_like(theta0); v = np.zeros_like(theta0); th = theta0.copy(); tg = theta0.copy()
    step_fn = onp.adam_step_torch if variant == rb.ADAM_TORCH else onp.adam_step_tf
    for t in range(1, 4):
        gr = cr.grads(s, a, y)[0].cpu().numpy().astype(np.float64)
        th, m, v = step_fn(th, gr, m, v, t, 1e-2)
        tg = onp.soft_update(tg, th, 0.01)
        opt.step(s, a, y)
        np.testing.assert_allclose(cr.theta.cpu().numpy(), th, rtol=0, atol=3e-5)
    np.testing.assert_allclose(tgt.theta.cpu().numpy(), tg, rtol=0, atol=3e-5)


# ----------------------------------------------------------------------------- a17 replay
def test_replay_gather_bit_exact(eng):
    from rlcontrol_b200.replaybuffer import ReplayBuffer
    rng = np.random.RandomState(13)
    S, A = 17, 6
    buf = ReplayBuffer(buffer_size=500, random_seed=3, state_dim=S, action_dim=A, engine=eng)
    store = {k: [] for k in ("state", "action", "reward", "next_state", "gamma")}
    for i in range(1234):                                     # wraps the ring twice
        tr = (rng.randn(S), rng.randn(A), float(rng.randn()), rng.randn(S), 0.99 if i % 7 else 0.0)
        buf.add(*tr)
        for k, x in zip(store, tr):
            store[k].append(np.asarray(x, np.float32))
    fifo = {k: np.array(v[-500:]) for k, v in store.items()}
    assert buf.get_size() == 500
    ref_rng = np.random.RandomState(3)
    for B in (32, 300, 1):
        s, a, r, s2, g = buf.sample_batch(B, as_numpy=True)
        idx = onp.sample_n_k(ref_rng, 500, B)
        rs, ra, rr, rs2, rg = onp.replay_gather(fifo, idx)
        assert np.array_equal(s, rs) and np.array_equal(a, ra) and np.array_equal(r, rr)
        assert np.array_equal(s2, rs2) and np.array_equal(g, rg)
    with pytest.raises(AssertionError):
        ReplayBuffer(10, 0, S, A, engine=eng).sample_batch(1)
