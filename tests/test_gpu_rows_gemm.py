"""GPU parity of the tensor-core GEMM of the B-row training path (csrc/rows_gemm_tc.cu: tcgen05 kind::tf32, 3 x TF32 operand
split) against fp64 numpy, through the C-ABI (rlc_rows_gemm), and of the training entry points that run on it at dense
minibatch sizes (rlc_critic_grads, rlc_mlp_forward / rlc_mlp_grads, rlc_critic_grad_action) against the oracle and against
the CUDA-core path.  The reference computes all of these in fp32 (torch.nn.Linear + autograd, forwardkl_network.py:
263-268,196-209): the bar is fp32-class agreement, stated per test."""
import numpy as np
import pytest

from oracle import oracle_np as onp
from test_gpu_parity import _rand_tin, _rand_tmid, _tin

pytestmark = pytest.mark.gpu

# |C - ref| <= TOL * (|op(A)| |op(B)|)[m,n]  (the natural scale of a dot product's rounding error); fp32 accumulation of
# K terms alone gives ~K^0.5 * 6e-8, the dropped lo.lo term 2^-22 = 2.4e-7
TOL_TC = 2e-6


def _gemm(eng, A, Bm, ta, tb, bias=None, Z=None, relu_a=False, alpha=1.0, split_k=False, path=2, pad=0):
    import torch
    from rlcontrol_b200._lib import check
    from rlcontrol_b200.engine import _ptr, _stream
    dev = eng.device
    M, K = (A.shape[1], A.shape[0]) if ta else A.shape
    N = Bm.shape[0] if tb else Bm.shape[1]

    def padded(x):                                       # leading dimension != width, and a misaligned base when pad is odd
        buf = torch.zeros(x.shape[0] * (x.shape[1] + pad) + pad, dtype=torch.float32, device=dev)
        v = buf[pad:].view(x.shape[0], x.shape[1] + pad)[:, : x.shape[1]]
        v.copy_(torch.as_tensor(x))
        return buf, v
    Ab, Av = padded(A)
    Bb, Bv = padded(Bm)
    Cb, Cv = padded(np.full((M, N), np.nan, np.float32)) if not split_k else (None, torch.full((M, N), float("nan"), device=dev))
    bt = None if bias is None else torch.as_tensor(bias, device=dev)
    Zb, Zv = padded(Z) if Z is not None else (None, None)
    check(eng.lib.rlc_rows_gemm(eng.h, int(ta), int(tb), M, N, K, _ptr(Av), Av.stride(0), _ptr(Bv), Bv.stride(0),
                                _ptr(Cv), Cv.stride(0), _ptr(bt), _ptr(Zv), 0 if Zv is None else Zv.stride(0),
                                int(relu_a), float(alpha), int(split_k), path, _stream()))
    torch.cuda.synchronize()
    assert eng.umma_error() == 0
    return Cv.cpu().numpy()


def _ref(A, Bm, ta, tb, bias=None, Z=None, relu_a=False, alpha=1.0):
    a = (A.T if ta else A).astype(np.float64)
    b = (Bm.T if tb else Bm).astype(np.float64)
    if relu_a:
        a = np.maximum(a, 0)
    c = alpha * (a @ b)
    scale = abs(alpha) * (np.abs(a) @ np.abs(b)) + 1e-30
    if bias is not None:
        c = c + bias
        scale = scale + np.abs(bias)
    if Z is not None:
        c = c * (Z > 0)
    return c, scale


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (1, 1, 1), (130, 300, 23), (257, 129, 100), (64, 16, 400), (300, 400, 97),
                                   (129, 257, 33), (5, 7, 1000)])
def test_tc_gemm_shapes(eng, ta, tb, M, N, K):
    """Every operand orientation on ragged shapes (tile edges, K tails, single rows), tensor path forced."""
    rng = np.random.RandomState(M * 7 + N * 3 + K + 2 * ta + tb)
    A = rng.randn(*((K, M) if ta else (M, K))).astype(np.float32)
    Bm = rng.randn(*((N, K) if tb else (K, N))).astype(np.float32)
    for pad in (0, 3):
        c = _gemm(eng, A, Bm, ta, tb, pad=pad)
        ref, scale = _ref(A, Bm, ta, tb)
        err = (np.abs(c - ref) / scale).max()
        assert err < TOL_TC, f"pad={pad}: {err:.3e}"


def test_tc_gemm_epilogue_and_split_k(eng):
    rng = np.random.RandomState(5)
    M, N, K = 515, 300, 400
    A = rng.randn(M, K).astype(np.float32)
    W = (rng.randn(K, N) * 0.05).astype(np.float32)
    bias = rng.randn(N).astype(np.float32)
    Z = rng.randn(M, N).astype(np.float32)
    for pad in (0, 1, 4):
        c = _gemm(eng, A, W, 0, 0, bias=bias, Z=Z, relu_a=True, alpha=0.37, pad=pad)
        ref, scale = _ref(A, W, 0, 0, bias=bias, Z=Z, relu_a=True, alpha=0.37)
        assert (np.abs(c - ref) / scale).max() < TOL_TC
    # weight-gradient form: X^T G over a long batch dimension, split-K slabs summed in a fixed order
    for Kb, Mw, Nw in [(4096, 400, 300), (4099, 23, 400), (1000, 130, 140), (70, 40, 50)]:
        X = rng.randn(Kb, Mw).astype(np.float32)
        G = (rng.randn(Kb, Nw) * 1e-4).astype(np.float32)          # gradient-sized magnitudes: no scaling needed
        c1 = _gemm(eng, X, G, 1, 0, relu_a=True, split_k=True)
        c2 = _gemm(eng, X, G, 1, 0, relu_a=True, split_k=True)
        assert np.array_equal(c1, c2), "split-K sum must be deterministic"
        ref, scale = _ref(X, G, 1, 0, relu_a=True)
        assert (np.abs(c1 - ref) / scale).max() < TOL_TC
        c0 = _gemm(eng, X, G, 1, 0, relu_a=True, split_k=True, path=1)       # CUDA cores: same contract
        assert (np.abs(c0 - ref) / scale).max() < TOL_TC


def test_tc_gemm_dynamic_range(eng):
    """fp32 exponent range survives the split (an fp16 split would not): 1e-30 .. 1e+30 operands, exact powers of two."""
    M = N = 128
    K = 64
    A = np.zeros((M, K), np.float32)
    Bm = np.zeros((K, N), np.float32)
    A[:, 0] = 2.0 ** -100
    Bm[0, :] = 2.0 ** 90
    A[:, 1] = 2.0 ** 60
    Bm[1, :] = 2.0 ** -70
    A[3, 2] = 1.0 + 2.0 ** -20                        # needs the lo part
    Bm[2, 5] = 1.0 + 2.0 ** -19
    c = _gemm(eng, A, Bm, 0, 0)
    ref, _ = _ref(A, Bm, 0, 0)
    np.testing.assert_allclose(c, ref, rtol=3e-7, atol=0)


@pytest.mark.parametrize("topo", ["tin", "tmid"])
def test_critic_grads_dense_batch_on_tensor_cores(eng, topo):
    """rlc_critic_grads at a dense minibatch (B=2048, 17+6 -> 400 -> 300): the tcgen05 path against the fp64 oracle
    backprop and against the CUDA-core path (RLC dispatcher forced per call)."""
    import rlcontrol_b200 as rb
    from rlcontrol_b200 import _lib
    rng = np.random.RandomState(21)
    S, A, H1, H2, B = 17, 6, 400, 300, 2048
    s, a = rng.randn(B, S).astype(np.float32), rng.uniform(-1, 1, (B, A)).astype(np.float32)
    y = rng.randn(B).astype(np.float32)
    if topo == "tin":
        p = _rand_tin(rng, S, A, H1, H2)
        cr = _tin(eng, p, S, A, H1, H2)
        loss, grads = onp.tin_mse_grads(s, a, y, p)
        flat = np.concatenate([grads[0].T.ravel(), grads[1], grads[2].T.ravel(), grads[3], grads[4].ravel(), grads[5]])
    else:
        p = _rand_tmid(rng, S, A, H1, H2)
        cr = rb.Critic(eng, rb.TMID, S, A, H1, H2)
        cr.load(*p, rb.LAYOUT_IN_OUT)
        loss, grads = onp.tmid_mse_grads(s, a, y, p)
        flat = np.concatenate([x.ravel() for x in grads])
    n0 = eng.launches
    grad, gl, q = cr.grads(s, a, y)
    assert eng.umma_error() == 0
    g = grad.cpu().numpy().astype(np.float64)
    scale = np.abs(flat).max()
    assert abs(float(gl) - loss) < 1e-5 * max(1, loss)
    assert np.abs(g - flat).max() < 1e-5 * scale + 1e-9, np.abs(g - flat).max() / scale   # incl. the fp32 column sums of the bias gradients
    # the same call with the dispatcher's tensor path disabled: fp32 CUDA cores, same contract
    eng.lib.rlc_rows_gemm_force(0)
    try:
        grad0, _, q0 = cr.grads(s, a, y)
    finally:
        eng.lib.rlc_rows_gemm_force(-1)
    g0 = grad0.cpu().numpy().astype(np.float64)
    assert np.abs(g0 - flat).max() < 1e-5 * scale + 1e-9
    assert np.abs(g - g0).max() < 1e-5 * scale + 1e-9
    qn, q0n = q.cpu().numpy(), q0.cpu().numpy()
    assert np.abs(qn - q0n).max() < 5e-6 * np.abs(q0n).max()


def test_mlp_forward_grads_dense_batch(eng):
    """rlc_mlp_forward / rlc_mlp_grads (value and policy networks of the KL agents) at B=4096 rows on the tensor path
    against an fp64 numpy MLP."""
    import torch
    import rlcontrol_b200 as rb
    rng = np.random.RandomState(22)
    inp, H1, H2, O, B = 17, 400, 300, 12, 4096
    W1, b1 = (rng.randn(H1, inp) * 0.2).astype(np.float32), (rng.randn(H1) * 0.1).astype(np.float32)
    W2, b2 = (rng.randn(H2, H1) * 0.05).astype(np.float32), (rng.randn(H2) * 0.1).astype(np.float32)
    W3, b3 = (rng.randn(O, H2) * 0.05).astype(np.float32), (rng.randn(O) * 0.1).astype(np.float32)
    x = rng.randn(B, inp).astype(np.float32)
    dout = (rng.randn(B, O) / B).astype(np.float32)
    m = rb.Mlp(eng, inp, H1, H2, O).load_torch(W1, b1, W2, b2, W3, b3)
    xt = torch.as_tensor(x, device=eng.device)
    act = m.act_buffer(B)
    out = m.forward(xt, act=act)
    grad, dx = m.grads(xt, torch.as_tensor(dout, device=eng.device), act=act, want_dx=True)
    assert eng.umma_error() == 0
    f = np.float64
    z1 = x.astype(f) @ W1.T.astype(f) + b1
    h1 = np.maximum(z1, 0)
    z2 = h1 @ W2.T.astype(f) + b2
    h2 = np.maximum(z2, 0)
    o = h2 @ W3.T.astype(f) + b3
    assert np.abs(out.cpu().numpy() - o).max() < 3e-6 * np.abs(o).max()
    # ReLU masks from the device's own pre-activations: a unit within rounding of 0 may sit on either side in fp32 and
    # fp64, and one flipped unit moves a gradient element by a whole term (~1e-5 here)
    a = act.cpu().numpy()
    nz1 = (B * H1 + 3) // 4 * 4
    m1, m2 = a[: B * H1].reshape(B, H1) > 0, a[nz1: nz1 + B * H2].reshape(B, H2) > 0
    assert (m1 != (z1 > 0)).mean() < 1e-5 and (m2 != (z2 > 0)).mean() < 1e-5
    h1, h2 = z1 * m1, z2 * m2
    g2 = (dout.astype(f) @ W3.astype(f)) * m2
    g1 = (g2 @ W2.astype(f)) * m1
    ref = np.concatenate([(x.T.astype(f) @ g1).ravel(), g1.sum(0), (h1.T @ g2).ravel(), g2.sum(0),
                          (h2.T @ dout.astype(f)).ravel(), dout.astype(f).sum(0)])
    g = grad.cpu().numpy().astype(f)
    assert np.abs(g - ref).max() < 3e-6 * np.abs(ref).max()
    np.testing.assert_allclose(dx.cpu().numpy(), g1 @ W1.astype(f), rtol=1e-4, atol=3e-6 * np.abs(g1 @ W1.astype(f)).max())


def test_grad_action_dense_rows(eng):
    """dQ/da on 16 384 stacked rows (T-in: forward + input gradient through both layers on the tensor path)."""
    rng = np.random.RandomState(23)
    S, A, H1, H2, R = 17, 6, 400, 300, 16384 + 77
    p = _rand_tin(rng, S, A, H1, H2)
    cr = _tin(eng, p, S, A, H1, H2)
    s = rng.randn(R, S).astype(np.float32)
    a = rng.uniform(-1, 1, (R, A)).astype(np.float32)
    g, q = cr.grad_action(s, a)
    assert eng.umma_error() == 0
    idx = rng.choice(R, 512, replace=False)
    gref, qref = onp.tin_dq_da(s[idx], a[idx], p), onp.tin_forward(s[idx], a[idx], *p, dtype=np.float64)
    gn = g.cpu().numpy()[idx]
    assert np.abs(gn - gref).max() < 5e-6 * np.abs(gref).max() + 1e-7
    np.testing.assert_allclose(q.cpu().numpy()[idx], qref.reshape(-1), rtol=1e-5, atol=2e-6)


def test_rows_gemm_rejects_bad_arguments(eng):
    """Leading dimensions smaller than the row length, a split-K request that is not a weight-gradient form: RLC_ERR_INVALID
    before anything is launched."""
    import torch
    from rlcontrol_b200._lib import RlcError, check
    from rlcontrol_b200.engine import _ptr, _stream
    x = torch.zeros(64 * 64, device=eng.device)
    n0 = eng.launches
    for args in ((0, 0, 8, 8, 8, 4, 8, 8, 0, 0),        # lda < K
                 (0, 1, 8, 8, 8, 8, 4, 8, 0, 0),        # ldb < K for a transposed B
                 (0, 0, 8, 8, 8, 8, 8, 4, 0, 0),        # ldc < N
                 (0, 0, 8, 8, 8, 8, 8, 8, 1, 0),        # split-K needs trans_a
                 (0, 0, 8, 8, 8, 8, 8, 8, 0, 3)):       # unknown path
        ta, tb, M, N, K, lda, ldb, ldc, split, path = args
        with pytest.raises(RlcError):
            check(eng.lib.rlc_rows_gemm(eng.h, ta, tb, M, N, K, _ptr(x), lda, _ptr(x), ldb, _ptr(x), ldc, None, None, 0, 0, 1.0,
                                        split, path, _stream()))
    assert eng.launches == n0


def test_tc_gemm_random_shapes(eng):
    """Seeded sweep over ragged shapes, operand orientations, leading-dimension pads / misaligned bases and epilogue
    options (tile edges in M, N and K, K tails of 1..31, vector and scalar load/store paths)."""
    rng = np.random.RandomState(77)
    for it in range(48):
        M, N, K = int(rng.randint(1, 400)), int(rng.randint(1, 400)), int(rng.randint(1, 300))
        ta, tb = int(rng.randint(2)), int(rng.randint(2))
        pad = int(rng.choice([0, 0, 1, 2, 4]))
        A = rng.randn(*((K, M) if ta else (M, K))).astype(np.float32)
        Bm = rng.randn(*((N, K) if tb else (K, N))).astype(np.float32)
        bias = rng.randn(N).astype(np.float32) if rng.randint(2) else None
        Z = rng.randn(M, N).astype(np.float32) if rng.randint(2) else None
        relu_a, alpha = bool(rng.randint(2)), float(rng.choice([1.0, -0.5, 3.25]))
        c = _gemm(eng, A, Bm, ta, tb, bias=bias, Z=Z, relu_a=relu_a, alpha=alpha, pad=pad)
        ref, scale = _ref(A, Bm, ta, tb, bias=bias, Z=Z, relu_a=relu_a, alpha=alpha)
        err = (np.abs(c - ref) / scale).max()
        assert err < TOL_TC, (it, M, N, K, ta, tb, pad, err)
    for it in range(12):                                 # split-K weight-gradient form, incl. K shorter than one slice
        K, M, N = int(rng.randint(1, 5000)), int(rng.randint(1, 420)), int(rng.randint(1, 420))
        X, G = rng.randn(K, M).astype(np.float32), rng.randn(K, N).astype(np.float32)
        c = _gemm(eng, X, G, 1, 0, split_k=True)
        ref, scale = _ref(X, G, 1, 0)
        assert (np.abs(c - ref) / scale).max() < TOL_TC, (it, K, M, N)
