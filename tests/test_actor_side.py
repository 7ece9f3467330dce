"""GPU tests of the Actor-Expert actor side fused with the sampled-action path:
rlc_mixture_sample, rlc_ae_expert_step, rlc_mixture_nll against the numpy oracle (the sampling restatement is
itself pinned on numpy's RandomState in tests/test_oracle.py)."""
import numpy as np
import pytest

from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def _mixture(rng, B, M, A, amax=1.0):
    alpha = rng.dirichlet(np.ones(M), B).astype(np.float32)
    mean = (np.tanh(rng.randn(B, M, A)) * amax).astype(np.float32)
    sigma = np.exp(rng.uniform(-3, 0, (B, M, A))).astype(np.float32)
    return alpha, mean, sigma


@pytest.mark.parametrize("B,M,A,N,n_uni,equal", [(32, 1, 1, 120, 0, False), (5, 3, 2, 77, 0, False), (4, 2, 6, 1024, 100, False),
                                                  (3, 4, 3, 33, 0, True), (1, 8, 1, 1, 1, False)])
def test_mixture_sample_bit_exact(eng, B, M, A, N, n_uni, equal):
    rng = np.random.RandomState(B * 7 + M)
    alpha, mean, sigma = _mixture(rng, B, M, A, 2.0)
    comp_u = rng.random_sample((B, N)).astype(np.float32)
    normal = rng.standard_normal((B, N, A)).astype(np.float32)
    uni = rng.random_sample((B, n_uni, A)).astype(np.float32) if n_uni else None
    amin, amax = -2.0 * np.ones(A), 2.0 * np.ones(A)
    acts, comp = eng.mixture_sample(None if equal else alpha, mean, sigma, comp_u, normal, amin, amax,
                                    equal_modal=equal, uni_u=uni, want_comp=True)
    ref, idx = onp.mixture_sample(alpha, mean, sigma, comp_u, normal, amin, amax, equal_modal=equal, uni_u=uni)
    np.testing.assert_array_equal(comp.cpu().numpy(), idx)
    np.testing.assert_array_equal(acts.cpu().numpy(), ref.astype(np.float32))      # fp64 arithmetic, one fp32 cast


def _tmid(eng, rng, S, A, H1, H2):
    import rlcontrol_b200 as rb
    k1, k2 = np.sqrt(3 / S), np.sqrt(3 / (H1 + A))
    u = lambda k, *sh: rng.uniform(-k, k, sh).astype(np.float32)
    p = [u(k1, S, H1), u(k1, H1), u(k2, H1 + A, H2), u(k2, H2), u(0.3, H2, 1), u(0.3, 1)]
    smin, smax = -2 * np.ones(S), 2 * np.ones(S)
    return rb.Critic(eng, rb.TMID, S, A, H1, H2, smin, smax).load(*p, rb.LAYOUT_IN_OUT), p, smin, smax


@pytest.mark.parametrize("S,A,H1,H2,B,N,k,M", [(1, 1, 200, 200, 32, 120, 6, 1), (17, 6, 400, 300, 64, 1024, 6, 2),
                                                (4, 3, 33, 47, 7, 50, 50, 3), (2, 2, 16, 16, 1, 9, 1, 2)])
def test_ae_expert_step_matches_unfused_path_and_oracle(eng, S, A, H1, H2, B, N, k, M):
    """cfg2 (Bimodal1DEnv Actor-Expert: S=1, A=1, 200-200, B=32, N=120, k=6) and friends: the fused launch returns
    the sampled actions bit-exactly, the same elite indices as predict_q + argsort()[::-1][:k] wherever the Q gap
    exceeds the fp32 noise, and exactly the actions at those indices."""
    rng = np.random.RandomState(S * 100 + N)
    cr, p, smin, smax = _tmid(eng, rng, S, A, H1, H2)
    alpha, mean, sigma = _mixture(rng, B, M, A)
    s = (rng.randn(B, S) * 1.5).astype(np.float32)
    comp_u = rng.random_sample((B, N)).astype(np.float32)
    normal = rng.standard_normal((B, N, A)).astype(np.float32)
    amin, amax = -np.ones(A), np.ones(A)
    out = cr.ae_expert_step(s, k, alpha, mean, sigma, comp_u, normal, amin, amax, want_actions=True, want_q=True)
    acts_ref, _ = onp.mixture_sample(alpha, mean, sigma, comp_u, normal, amin, amax)
    acts_ref = acts_ref.astype(np.float32)
    np.testing.assert_array_equal(out["actions"].cpu().numpy(), acts_ref)
    q_ref = onp.tmid_eval(s, acts_ref, p, smin, smax, dtype=np.float64)
    q = out["q"].cpu().numpy()
    np.testing.assert_allclose(q, q_ref, rtol=2e-5, atol=2e-5)
    idx = out["idx"].cpu().numpy()
    np.testing.assert_array_equal(idx, onp.topk_desc(q, k))                  # bit-exact on the kernel's own q
    ref_idx = onp.topk_desc(q_ref.astype(np.float32), k)
    for b in range(B):                                                       # vs the fp64 oracle: exact where gaps > noise
        if not np.array_equal(idx[b], ref_idx[b]):
            srt = np.sort(q_ref[b])[::-1]
            assert np.min(np.abs(np.diff(srt[:k + 1]))) < 1e-4 * max(1.0, np.abs(srt).max())
    elites = out["elites"].cpu().numpy()
    np.testing.assert_array_equal(elites, onp.gather_elites(acts_ref, idx))
    np.testing.assert_array_equal(out["q_sel"].cpu().numpy(), np.take_along_axis(q, idx, 1))
    # the unfused C-ABI route gives the same indices (same q arithmetic is not required, same ordering is)
    import torch
    q2 = cr.eval(s, torch.as_tensor(acts_ref, device=eng.device), "fp32")
    idx2, _, el2 = eng.topk(q2, k, torch.as_tensor(acts_ref, device=eng.device))
    same = (idx2.cpu().numpy() == idx).all(axis=1).mean()
    assert same >= 0.9


@pytest.mark.parametrize("B,M,A,k,equal", [(32, 1, 1, 6, False), (9, 2, 3, 40, False), (4, 3, 2, 5, True), (2, 8, 6, 64, False)])
def test_mixture_nll_matches_oracle(eng, B, M, A, k, equal):
    rng = np.random.RandomState(B + k)
    alpha, mean, sigma = _mixture(rng, B, M, A)
    sigma = np.maximum(sigma, 0.2).astype(np.float32)
    y = rng.uniform(-1, 1, (B, k, A)).astype(np.float32)
    loss, nll, da, dm, ds = eng.mixture_nll(None if equal else alpha, mean, sigma, y, equal_modal=equal, b_total=2 * B)
    rl, rn, rda, rdm, rds = onp.mixture_nll(alpha, mean, sigma, y, equal_modal=equal, b_total=2 * B)
    np.testing.assert_allclose(float(loss.cpu()), rl, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(nll.cpu().numpy(), rn, rtol=2e-5, atol=2e-5)
    for mine, ref in ((da, rda), (dm, rdm), (ds, rds)):
        np.testing.assert_allclose(mine.cpu().numpy(), ref, rtol=2e-4, atol=2e-6 * max(1.0, np.abs(ref).max()))


def test_mixture_nll_clip_has_no_gradient(eng):
    """density underflow: clip_by_value(mix, 1e-30, 1e30) -> loss = -log(1e-30), zero gradient (ae_network.py:276)."""
    loss, nll, da, dm, ds = eng.mixture_nll(np.ones((1, 1)), np.zeros((1, 1, 1)), np.full((1, 1, 1), 0.01), np.full((1, 1, 1), 5.0))
    assert abs(float(loss.cpu()) + np.log(1e-30)) < 1e-3
    assert float(dm.abs().max().cpu()) == 0.0 and float(ds.abs().max().cpu()) == 0.0


def test_actor_expert_critic_sample_and_select(eng):
    """networks.ActorExpertCritic.sample_and_select_elites draws from its RandomState the way the reference's
    sample_action does and returns elites = actions[idx]."""
    from types import SimpleNamespace
    from rlcontrol_b200.networks import ActorExpertCritic
    cfg = SimpleNamespace(state_dim=1, state_min=[-1.0], state_max=[1.0], action_dim=1, action_min=[-1.0], action_max=[1.0],
                          tau=0.01, norm_type="input_norm", random_seed=3, engine=eng, expert_lr=1e-3, shared_l1_dim=200,
                          expert_l2_dim=200)
    net = ActorExpertCritic(None, None, cfg)
    rng = np.random.RandomState(0)
    B, N, k, M = 32, 120, 6, 1
    alpha, mean, sigma = _mixture(rng, B, M, 1)
    s = rng.uniform(-1, 1, (B, 1))
    elites, idx, acts = net.sample_and_select_elites(s, alpha[:, :, None], mean, sigma, k, N, rng=np.random.RandomState(9))
    assert elites.shape == (B, k, 1) and idx.shape == (B, k) and acts.shape == (B, N, 1)
    np.testing.assert_array_equal(elites, onp.gather_elites(acts, idx))
    q = net.predict_q(np.repeat(s, N, axis=0), acts.reshape(B * N, 1), True).reshape(B, N)
    agree = (onp.topk_desc(q, k) == idx).all(axis=1).mean()
    assert agree >= 0.9


@pytest.mark.parametrize("S,A,H1,H2,B,Kf,Ku", [(3, 1, 200, 200, 32, 15, 15), (17, 6, 400, 300, 8, 16, 16), (2, 2, 24, 20, 5, 7, 3),
                                                (4, 3, 16, 16, 1, 64, 64)])
def test_svgd_action_gradients_match_oracle(eng, S, A, H1, H2, B, Kf, Ku):
    """sql.json shape (kernel_n_particles 30, ratio 0.5 -> 15 + 15) and others: dQ/da on the B*Kf rows without
    materialised states, median bandwidth, Stein direction."""
    import rlcontrol_b200 as rb
    rng = np.random.RandomState(Kf * 10 + A)
    k1, k2 = 1 / np.sqrt(S + A), 1 / np.sqrt(H1)
    u = lambda k, *sh: rng.uniform(-k, k, sh).astype(np.float32)
    p = [u(k1, H1, S + A), u(k1, H1), u(k2, H2, H1), u(k2, H2), u(0.5, 1, H2), u(0.5, 1)]
    cr = rb.Critic(eng, rb.TIN, S, A, H1, H2).load(*p, rb.LAYOUT_OUT_IN)
    s = rng.randn(B, S).astype(np.float32)
    parts = np.tanh(rng.randn(B, Kf + Ku, A)).astype(np.float32) * 0.98
    fixed, updated = np.ascontiguousarray(parts[:, :Kf]), np.ascontiguousarray(parts[:, Kf:])
    g, aux = cr.svgd_action_grads(s, fixed, updated, want_aux=True)
    s_rows = np.repeat(s, Kf, axis=0)
    dq_ref = onp.tin_dq_da(s_rows, fixed.reshape(B * Kf, A), p).reshape(B, Kf, A)
    np.testing.assert_allclose(aux["dqda"].cpu().numpy(), dq_ref, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(aux["q_fixed"].cpu().numpy(), onp.tin_forward(s_rows, fixed.reshape(B * Kf, A), *p, dtype=np.float64).reshape(B, Kf),
                               rtol=2e-5, atol=2e-5)
    ref, kap, h = onp.svgd_action_gradients(dq_ref, fixed, updated)
    np.testing.assert_allclose(aux["h"].cpu().numpy(), h, rtol=1e-5)
    np.testing.assert_allclose(aux["kappa"].cpu().numpy(), kap, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(g.cpu().numpy(), ref, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(ref).max()))
