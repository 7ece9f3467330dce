"""CPU tests: the full FKL / RKL update restatement (oracle/oracle_kl.py) and the product-side quadrature
(rlcontrol_b200/quadrature.py) against fixtures recorded from the UNMODIFIED reference classes
(tests/golden/full_*.npz, smolyak.npz, cc.npz; generator: oracle/make_golden.py)."""
import numpy as np
import pytest

from conftest import golden
from oracle import oracle_kl as okl
from rlcontrol_b200 import quadrature

FULL = ["full_fkl_intg_nonsac", "full_fkl_intg_sac", "full_rkl_intg_nonsac", "full_rkl_hardintg_sac",
        "full_rkl_ll_nonsac", "full_rkl_hardll_sac"]
# the same reference update at a DENSE minibatch (B = 2048, 128-128 networks): one update each; on the GPU these run the B-row
# training path on the tcgen05 3xTF32 GEMMs (csrc/rows_gemm_tc.cu) instead of the small-minibatch kernels
FULL_DENSE = ["full_fkl_intg_nonsac_dense", "full_rkl_intg_sac_dense"]
NP = dict(q=6, v=6, tv=6, pi=8)


def load_full(name):
    g = golden(name + ".npz")
    pre = {k: [g["pre_%s_%d" % (k, i)] for i in range(n)] for k, n in NP.items()}
    post = [{k: [g["post%d_%s_%d" % (u, k, i)] for i in range(n)] for k, n in NP.items()}
            for u in range(g["s"].shape[0])]
    return g, pre, post


def make_agent(name, g, pre):
    kind = "fkl" if "fkl" in name else "rkl"
    return okl.KLAgent(kind, pre["q"], pre["v"], pre["tv"], pre["pi"], g["grid_a"], g["grid_w"],
                       float(g["action_max"]), float(g["alpha"]), float(g["pi_lr"]), float(g["qf_vf_lr"]),
                       float(g["tau"]), optim_type=str(g["optim_type"]), q_update_type=str(g["q_update_type"]))


@pytest.mark.parametrize("name", FULL + FULL_DENSE)
def test_oracle_full_update_matches_reference(name):
    """Two consecutive update_network + update_target_network calls: every parameter of q_net, v_net,
    target_v_net and pi_net lands where the reference's torch autograd + Adam put it (fp32 there, fp64 here)."""
    g, pre, post = load_full(name)
    ag = make_agent(name, g, pre)
    for u in range(g["s"].shape[0]):
        losses = ag.update(g["s"][u], g["a"][u], g["s2"][u], g["r"][u], g["g"][u], g["eps"][u])
        np.testing.assert_allclose(losses, g["losses"][u], rtol=2e-4, atol=2e-5)
        for k, mine in (("q", ag.q), ("v", ag.v), ("tv", ag.tv), ("pi", ag.pi)):
            for i, (m, ref) in enumerate(zip(mine, post[u][k])):
                # one Adam step moves a weight by ~lr; the comparison is on the MOVE (parameters barely change)
                move = np.abs(ref - pre[k][i]).max() + 1e-12
                err = np.abs(m - ref).max()
                # dense fixtures: the first Adam step is lr * g / (|g| + 1e-8), and with a 2048-row mean more elements of g sit
                # near 1e-8, where the reference's fp32 gradient noise moves the step itself
                tol = 6e-3 if name in FULL_DENSE else 2e-3
                assert err <= tol * move + 2e-7, (name, u, k, i, err, move)
    st = g["act_states"]
    np.testing.assert_allclose(ag.sample_action(st, g["act_eps"]), g["act_sample"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ag.predict_action(st), g["act_predict"], rtol=1e-4, atol=1e-5)


def test_policy_evaluate_matches_recorded_sample():
    g, pre, _ = load_full("full_rkl_ll_nonsac")
    mean, log_std, _, _, _ = okl.policy_forward(g["s"][0], pre["pi"])
    _, logp, z, _ = okl.policy_evaluate(mean, log_std, g["eps"][0], float(g["action_max"]))
    np.testing.assert_allclose(z, g["z"][0], rtol=0, atol=2e-6)
    np.testing.assert_allclose(logp, g["logp_sample"][0], rtol=2e-5, atol=2e-5)


def test_quadrature_1d_is_the_reference_grid():
    cc = golden("cc.npz")
    x, w = quadrature.clenshaw_curtis(64)
    np.testing.assert_allclose(x, cc["x64"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(w, cc["w64"], rtol=0, atol=1e-15)
    g = golden("full_fkl_intg_nonsac.npz")
    a, w = quadrature.grid_1d(int(g["n_param"]), float(g["action_max"]))
    assert a.dtype == np.float32 and w.dtype == np.float32
    np.testing.assert_array_equal(a, g["grid_a"])          # bit-exact with the reference's fp32 tensors
    np.testing.assert_array_equal(w, g["grid_w"])


@pytest.mark.parametrize("A,l", [(2, 6), (3, 4)])
def test_quadrature_smolyak_is_the_reference_grid(A, l):
    g = golden("smolyak.npz")
    a, w = quadrature.grid_smolyak(l, A, 1.5)
    ra, rw = g["a_%d_%d" % (A, l)], g["w_%d_%d" % (A, l)]
    assert a.shape == ra.shape
    np.testing.assert_allclose(a, ra, rtol=0, atol=1e-7)   # the reference holds float64(float32 node) * 1.5
    np.testing.assert_array_equal(w, rw)
    assert quadrature.integration_grid(A, 1.5, l_param=l)[0].shape == ra.shape


def test_clenshaw_curtis_exactness():
    """n-point CC integrates polynomials of degree < n exactly; weights are positive and sum to 2."""
    for n in (3, 5, 9, 17, 64, 1026):
        x, w = quadrature.clenshaw_curtis(n)
        assert np.all(w > 0) and abs(w.sum() - 2) < 1e-12 and np.all(np.diff(x) > 0)
        for p in [p for p in (2, 4, 10) if p < n]:
            assert abs((w * x ** p).sum() - 2.0 / (p + 1)) < 1e-12
    with pytest.raises(ValueError):
        quadrature.clenshaw_curtis(1)
