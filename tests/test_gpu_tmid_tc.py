"""GPU parity of the tensor-core T-mid stack evaluation (csrc/tmid_rows_tc.cu: 128-row tiles of one state as tcgen05
kind::tf32 GEMMs, 3 x TF32 split, state term as two "ones" K columns, head folded by sign partition) against the fp64
oracle of the reference's T-mid critic (critic_network.py:77-99) and against the fp32 CUDA-core kernels, through
rlc_critic_eval.  Bar: the same TOL_FP32 the CUDA-core path is held to; top-k indices bit-identical to the oracle's
wherever the oracle's gap exceeds 1e-4 of the state's rms Q."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import oracle_np as onp
from test_gpu_parity import TOL_FP32, _rand_tmid

pytestmark = pytest.mark.gpu


def _critic(eng, rng, S, A, H1, H2, last=1.0):
    import rlcontrol_b200 as rb
    p = _rand_tmid(rng, S, A, H1, H2, last=last)
    smin, smax = -np.ones(S) * 1.5, np.ones(S) * 1.5
    cr = rb.Critic(eng, rb.TMID, S, A, H1, H2, smin, smax)
    cr.load(*p, rb.LAYOUT_IN_OUT)
    return cr, p, smin, smax


def _eval_forced(eng, cr, s, a, mode):
    prev = eng.lib.rlc_tmid_tc_force(mode)
    prev_g = eng.lib.rlc_rows_gemm_force(0 if mode == 0 else -1)      # mode 0: the state term on CUDA cores as well
    try:
        q = cr.eval(s, a, "fp32").cpu().numpy()
    finally:
        eng.lib.rlc_tmid_tc_force(prev)
        eng.lib.rlc_rows_gemm_force(prev_g)
    assert eng.umma_error() == 0
    return q


@pytest.mark.parametrize("S,A,H1,H2,B,N", [(1, 1, 200, 200, 32, 120), (17, 6, 400, 300, 16, 200), (3, 2, 50, 33, 5, 77),
                                           (4, 8, 64, 64, 3, 40), (17, 6, 400, 300, 7, 1024), (2, 3, 16, 16, 1, 1),
                                           (5, 5, 40, 500, 9, 129), (6, 7, 30, 304, 300, 128), (3, 4, 20, 150, 2, 383)])
@pytest.mark.parametrize("per_state", [True, False])
def test_tmid_tc_matches_oracle(eng, S, A, H1, H2, B, N, per_state):
    rng = np.random.RandomState(S * 100 + A + N)
    cr, p, smin, smax = _critic(eng, rng, S, A, H1, H2)
    s = (rng.randn(B, S) * 2).astype(np.float32)
    a = rng.uniform(-1, 1, (B, N, A) if per_state else (N, A)).astype(np.float32)
    ref = onp.tmid_eval(s, a, p, smin, smax, dtype=np.float64)
    q = _eval_forced(eng, cr, s, a, 2)
    assert q.shape == (B, N)
    assert rel_err(q, ref).max() < TOL_FP32
    q0 = _eval_forced(eng, cr, s, a, 0)
    assert rel_err(q0, ref).max() < TOL_FP32
    assert rel_err(q, q0).max() < TOL_FP32


def test_tmid_tc_one_sided_heads_and_zero_weights(eng):
    """Sign partition edge cases: all-positive, all-negative and partly zero output weights."""
    rng = np.random.RandomState(4)
    S, A, H1, H2, B, N = 3, 2, 24, 47, 4, 200
    for kind in ("pos", "neg", "zeros"):
        cr, p, smin, smax = _critic(eng, rng, S, A, H1, H2)
        if kind == "pos":
            p[4] = np.abs(p[4]) + 0.01
        elif kind == "neg":
            p[4] = -np.abs(p[4]) - 0.01
        else:
            p[4][::3] = 0.0
        import rlcontrol_b200 as rb
        cr.load(*p, rb.LAYOUT_IN_OUT)
        s = rng.randn(B, S).astype(np.float32)
        a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
        ref = onp.tmid_eval(s, a, p, smin, smax, dtype=np.float64)
        assert rel_err(_eval_forced(eng, cr, s, a, 2), ref).max() < TOL_FP32, kind


def test_tmid_tc_large_stack_default_dispatch_and_topk(eng):
    """At stack size (161 x 1024 rows >= 148 x 1024) the dispatcher itself takes the tensor path; q within TOL_FP32 of the
    oracle on sampled states, deterministic, and the top-6 indices equal the oracle's wherever its gap allows."""
    rng = np.random.RandomState(9)
    S, A, H1, H2, B, N = 17, 6, 400, 300, 161, 1024
    cr, p, smin, smax = _critic(eng, rng, S, A, H1, H2)
    s = (rng.randn(B, S) * 2).astype(np.float32)
    a = rng.uniform(-1, 1, (B, N, A)).astype(np.float32)
    q = cr.eval(s, a, "fp32")
    q2 = cr.eval(s, a, "fp32")
    assert eng.umma_error() == 0
    assert np.array_equal(q.cpu().numpy(), q2.cpu().numpy())
    q0 = _eval_forced(eng, cr, s, a, 0)
    qn = q.cpu().numpy()
    assert not np.array_equal(qn, q0), "the dispatcher should have taken the tensor path at this size"
    full = onp.tmid_eval(s, a, p, smin, smax, dtype=np.float64)          # every row of the stack against the fp64 oracle
    assert rel_err(qn, full).max() < TOL_FP32
    assert rel_err(q0, full).max() < TOL_FP32
    rows = np.array([0, 1, 77, B - 1])
    ref = full[rows]
    idx, _, _ = eng.topk(q, 6)
    idx = idx.cpu().numpy()[rows]
    order = np.argsort(-ref, axis=1, kind="stable")
    rms = np.sqrt((ref ** 2).mean(1))
    for r in range(len(rows)):
        srt = ref[r, order[r]]
        gaps = srt[:6] - srt[1:7]
        if (gaps > 1e-4 * rms[r]).all():
            assert np.array_equal(idx[r], order[r, :6])


def test_tmid_eval_grad_large_stack_rows4(eng):
    """rlc_tmid_eval_grad on a stack that fills the machine takes the 4-rows-per-thread gradient kernel: q bit-identical to
    the forward rows4 kernel (same FMA sequence), dq/da against the oracle on sampled states, ragged tail included."""
    rng = np.random.RandomState(31)
    S, A, H1, H2, B, N = 17, 6, 400, 300, 161, 997
    cr, p, smin, smax = _critic(eng, rng, S, A, H1, H2)
    s = (rng.randn(B, S) * 2).astype(np.float32)
    a = rng.uniform(-1, 1, (B, N, A)).astype(np.float32)
    assert B * N >= 148 * 1024
    prev_g = eng.lib.rlc_rows_gemm_force(0)          # state term on the CUDA cores for both calls
    try:
        qg, g = cr.eval_grad(s, a)
    finally:
        eng.lib.rlc_rows_gemm_force(prev_g)
    q0 = _eval_forced(eng, cr, s, a, 0)
    np.testing.assert_array_equal(qg.cpu().numpy(), q0)
    qd, gd = cr.eval_grad(s, a)                      # default dispatch: the state term of 161 states runs as tensor-core GEMMs
    assert eng.umma_error() == 0
    assert rel_err(qd.cpu().numpy(), q0).max() < 2 * TOL_FP32        # two fp32-class results against each other
    rows = np.array([0, 3, 77, B - 1])
    assert rel_err(qd.cpu().numpy()[rows], onp.tmid_eval(s[rows], a[rows], p, smin, smax, dtype=np.float64)).max() < TOL_FP32
    gref = onp.tmid_dq_da(onp.stack_state_major(s[rows], N), a[rows].reshape(-1, A), p, smin, smax)
    gn = g.cpu().numpy()[rows].reshape(-1, A)
    assert np.abs(gn - gref).max() < 2e-5 * np.abs(gref).max() + 1e-6


def test_tmid_tc_random_shapes(eng):
    """Seeded sweep over (S, A <= 8, H1, H2 <= 480, B, N) with shared and per-state actions, tensor path forced: ragged tiles
    (N not a multiple of 128), single states, hidden widths that are not multiples of 16, every action width."""
    rng = np.random.RandomState(123)
    for it in range(30):
        S, A = int(rng.randint(1, 20)), int(rng.randint(1, 9))
        H1, H2 = int(rng.randint(8, 200)), int(rng.randint(1, 481))
        if it < 6:
            H2 = [345, 400, 480, 161, 256, 257][it]          # accumulator ring of 2 / 3 / 4 buffers, one and two parts
        B, N = int(rng.randint(1, 40)), int(rng.randint(1, 700))
        per_state = bool(rng.randint(2))
        cr, p, smin, smax = _critic(eng, rng, S, A, H1, H2)
        s = (rng.randn(B, S) * 2).astype(np.float32)
        a = rng.uniform(-1, 1, (B, N, A) if per_state else (N, A)).astype(np.float32)
        ref = onp.tmid_eval(s, a, p, smin, smax, dtype=np.float64)
        try:
            q = _eval_forced(eng, cr, s, a, 2)
        except AssertionError as ex:
            raise AssertionError(f"kernel error flag for {(it, S, A, H1, H2, B, N, per_state)}: {ex}")
        e = rel_err(q, ref).max()
        assert e < TOL_FP32, (it, S, A, H1, H2, B, N, per_state, e)
