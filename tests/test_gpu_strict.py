"""GPU parity of the STRICT tensor-core mode (precision "fp16x3", RLC_PREC_FP16X3; csrc/critic_umma_grid3.cuh):
north_star's bar is 1e-3 relative against the reference's fp32 arithmetic (forwardkl_network.py:263-268); this mode
is held to 2e-5 -- the same tolerance as the fp32 CUDA-core path -- on the reference's own golden vectors, on random
networks / ragged shapes, and on the bench workload.  Every call goes through the C-ABI."""
import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import oracle_np as onp
from test_gpu_parity import _p, _rand_tin, _tin

pytestmark = pytest.mark.gpu

TOL_STRICT = 2e-5          # max of |dq| / max(|q|, rms_state q) against the exact oracle / the reference's q
TOL_STATED = 2e-5          # max against the fp64 restatement of the kernel's own arithmetic (head="grid3")
# "fp16c8" (RLC_PREC_FP16C8): the same split with the two correction terms on the FP8 pipe.  Held to 6e-4 max against the
# exact oracle (north_star: 1e-3; the fp64 emulation of its arithmetic gives 1.7e-4 .. 3.5e-4 on these inputs) and to the
# same 2e-5 against ITS stated arithmetic (head="grid3c8") -- the kernel computes exactly what it says.
TOL_C8 = 6e-4
MODES = {"fp16x3": ("grid3", TOL_STRICT), "fp16c8": ("grid3c8", TOL_C8)}


def _check(cr, eng, s, a, p, exact=None, prec="fp16x3"):
    head, tol = MODES[prec]
    q = cr.eval(s, a, prec).cpu().numpy()
    assert eng.umma_error() == 0
    assert cr.tensor_arithmetic(True, prec) == head
    exact = onp.tin_eval(s, a, p, dtype=np.float64) if exact is None else exact
    e = rel_err(q, exact).max()
    assert e < tol, f"{prec} vs exact: {e:.3e}"
    d = rel_err(q, onp.tin_eval_rounded(s, a, p, head=head)).max()
    assert d < TOL_STATED, f"{prec} vs stated arithmetic: {d:.3e}"
    return q


@pytest.mark.parametrize("name,dims", [("tin_cfg1.npz", (3, 1, 200, 200)),
                                       ("tin_400_300.npz", (17, 6, 400, 300)),
                                       ("tin_cfg4_exact.npz", (3, 1, 400, 300))])
@pytest.mark.parametrize("prec", ["fp16x3", "fp16c8"])
def test_strict_golden(eng, name, dims, prec):
    """q of the reference's own SoftQNetwork (oracle/make_golden.py) within 2e-5 (fp16x3) / 6e-4 (fp16c8) on the tensor path."""
    g = golden(name)
    p = _p(g)
    _check(_tin(eng, p, *dims), eng, g["s"], g["a"], p, exact=g["q"].astype(np.float64), prec=prec)


@pytest.mark.parametrize("B,N", [(1, 1), (5, 257), (64, 62), (1, 513), (3, 31), (4, 32), (9, 1024), (130, 33)])
@pytest.mark.parametrize("prec", ["fp16x3", "fp16c8"])
def test_strict_ragged_shapes(eng, B, N, prec):
    rng = np.random.RandomState(B * 1000 + N)
    S, A, H1, H2 = 5, 2, 72, 40
    p = _rand_tin(rng, S, A, H1, H2, last=3.0)
    s = rng.randn(B, S).astype(np.float32)
    a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
    _check(_tin(eng, p, S, A, H1, H2), eng, s, a, p, prec=prec)


@pytest.mark.parametrize("S,A,H1,H2", [(3, 1, 200, 200), (17, 6, 400, 300), (1, 1, 32, 32), (30, 8, 256, 256),
                                       (11, 3, 100, 480), (2, 2, 16, 304), (4, 2, 95, 33), (6, 3, 191, 300)])
@pytest.mark.parametrize("prec", ["fp16x3", "fp16c8"])
def test_strict_network_shapes(eng, S, A, H1, H2, prec):
    """Chunking, slot count, accumulator parts and weight-ring geometry (resident / streamed) across widths."""
    rng = np.random.RandomState(S * 7 + H1)
    p = _rand_tin(rng, S, A, H1, H2, last=5.0)
    s = rng.randn(9, S).astype(np.float32)
    a = rng.uniform(-1, 1, (190, A)).astype(np.float32)
    cr = _tin(eng, p, S, A, H1, H2)
    if prec == "fp16c8" and H2 > 448:
        # 480 accumulator columns leave 32 TMEM columns: no room for two activation slots of one K=32 step each
        import rlcontrol_b200 as rb
        assert cr.tensor_arithmetic(True, prec) == "unsupported"
        with pytest.raises(rb.RlcError):
            cr.eval(s, a, prec)
        return
    _check(cr, eng, s, a, p, prec=prec)


def test_strict_state_clip_and_scales(eng):
    """Clipped states (TF T-in critic) and weights far from unit scale (the power-of-two normalisation of W2 and of the
    head must keep both hi and lo parts in fp16's normal range)."""
    import rlcontrol_b200 as rb
    rng = np.random.RandomState(5)
    S, A, H1, H2 = 6, 2, 128, 96
    for w2_scale, w3_scale in ((1.0, 1.0), (1e-3, 50.0), (30.0, 1e-3)):
        p = _rand_tin(rng, S, A, H1, H2, last=2.0)
        p[2] = (p[2] * w2_scale).astype(np.float32)
        p[3] = (p[3] * w2_scale).astype(np.float32)
        p[4] = (p[4] * w3_scale).astype(np.float32)
        s = (rng.randn(20, S) * 3).astype(np.float32)
        a = rng.uniform(-1, 1, (100, A)).astype(np.float32)
        smin, smax = -np.ones(S, np.float32), np.ones(S, np.float32) * 1.5
        cr = rb.Critic(eng, rb.TIN, S, A, H1, H2, state_min=smin, state_max=smax).load(*p, rb.LAYOUT_OUT_IN)
        sc = np.clip(s, smin, smax)
        q = cr.eval(s, a, "fp16x3").cpu().numpy()
        assert eng.umma_error() == 0
        assert rel_err(q, onp.tin_eval(sc, a, p, dtype=np.float64)).max() < TOL_STRICT


def test_strict_bench_workload_and_index_parity(eng):
    """cfg4 workload (S=17, A=6, 400-300, N=1024) on a 64-state slice: 2e-5 of the exact oracle, and argmax / top-6
    indices identical to the oracle's wherever the oracle's gap exceeds 1e-4 of the state's rms Q (north_star: bit-exact
    indices where the gap exceeds the tolerance)."""
    import bench
    W = bench.WORKLOAD
    params = bench.make_params(np.random.RandomState(0), W["S"], W["A"], W["H1"], W["H2"])
    s, a, _, _ = bench.make_inputs(np.random.RandomState(1000), 64, W["N"], W["S"], W["A"])
    cr = _tin(eng, params, W["S"], W["A"], W["H1"], W["H2"])
    ref = onp.tin_eval(s, a, params, dtype=np.float64)
    _check(cr, eng, s, a, params, exact=ref, prec="fp16c8")
    q = _check(cr, eng, s, a, params, exact=ref)
    import torch
    idx = eng.topk(torch.as_tensor(q, device=eng.device), 6)[0].cpu().numpy()
    ridx = onp.topk_desc(ref, 6)
    rms = np.sqrt((ref ** 2).mean(1))
    srt = -np.sort(-ref, axis=1)[:, :7]
    gaps = (srt[:, :-1] - srt[:, 1:]) / rms[:, None]           # gap below each of the top-6 ranks
    decided = gaps > 1e-4
    assert (idx == ridx)[decided].all(), "index mismatch at a decided rank"
    assert decided.mean() > 0.9


def test_strict_full_size_consistency(eng):
    """Full cfg4 size (B=4096 x N=1024): a row's result does not depend on where its tile sits -- the full launch must
    reproduce, bit for bit, the 64-state launch that the oracle checked above -- and the result is deterministic."""
    import bench
    import torch
    W = bench.WORKLOAD
    params = bench.make_params(np.random.RandomState(0), W["S"], W["A"], W["H1"], W["H2"])
    s, a, _, _ = bench.make_inputs(np.random.RandomState(1000), W["B_per_gpu"], W["N"], W["S"], W["A"])
    cr = _tin(eng, params, W["S"], W["A"], W["H1"], W["H2"])
    q_full = cr.eval(s, a, "fp16x3")
    q_again = cr.eval(s, a, "fp16x3")
    assert eng.umma_error() == 0
    assert torch.equal(q_full, q_again)
    rows = np.r_[0:32, 2000:2016, 4080:4096]
    q_part = cr.eval(s[rows], a, "fp16x3")
    assert torch.equal(q_full[torch.as_tensor(rows, device=eng.device)], q_part)
    ref = onp.tin_eval(s[rows], a, params, dtype=np.float64)
    assert rel_err(q_part.cpu().numpy(), ref).max() < TOL_STRICT
    # against the fp32 CUDA-core path on a larger slice (both are within 2e-5 of exact)
    q32 = cr.eval(s[:512], a, "fp32").cpu().numpy()
    assert rel_err(q_full[:512].cpu().numpy(), q32).max() < 2 * TOL_STRICT


def test_strict_unsupported_and_auto(eng):
    """Per-state action stacks are not a shared grid: the strict mode refuses (no silent downgrade); "auto" picks the
    strict mode for large shared-grid evaluations and the fp32 path otherwise -- never the single-rounding fp16 mode."""
    import rlcontrol_b200 as rb
    import torch
    rng = np.random.RandomState(3)
    S, A, H1, H2 = 17, 6, 400, 300
    p = _rand_tin(rng, S, A, H1, H2, last=3.0)
    cr = _tin(eng, p, S, A, H1, H2)
    s = rng.randn(32, S).astype(np.float32)
    a = rng.uniform(-1, 1, (1024, A)).astype(np.float32)
    with pytest.raises(rb.RlcError):
        cr.eval(s[:2], np.tile(a[None, :8], (2, 1, 1)), "fp16x3")
    assert torch.equal(cr.eval(s, a, "auto"), cr.eval(s, a, "fp16x3"))                 # 32768 rows, shared grid
    ap = np.tile(a[None], (32, 1, 1))
    assert torch.equal(cr.eval(s, ap, "auto"), cr.eval(s, ap, "fp32"))                 # per-state actions
    assert torch.equal(cr.eval(s[:4], a, "auto"), cr.eval(s[:4], a, "fp32"))           # 4096 rows: launch-bound


def test_strict_range_flag(eng):
    """Layer-1 pre-activations beyond fp16's range cannot be split: the pre-pass raises the handle's error flag (91)
    instead of returning saturated values silently."""
    rng = np.random.RandomState(4)
    S, A, H1, H2 = 3, 1, 64, 64
    p = _rand_tin(rng, S, A, H1, H2)
    cr = _tin(eng, p, S, A, H1, H2)
    s = (rng.randn(8, S) * 1e6).astype(np.float32)
    a = rng.uniform(-1, 1, (64, A)).astype(np.float32)
    for prec in ("fp16x3", "fp16c8", "fp16"):
        cr.eval(s, a, prec)
        assert eng.umma_error() == 91, prec
        assert eng.umma_error() == 0            # cleared on read
    q = cr.eval(s, a, "bf16")                   # bf16 has fp32's range: no flag, finite results
    assert eng.umma_error() == 0 and np.isfinite(q.cpu().numpy()).all()


# ---------------------------------------------------------------------------------------------
# fused evaluation + per-state policy reduction (rlc_critic_eval_reduce_policy): the reduction inside K1's epilogue
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["fkl", "rkl"])
@pytest.mark.parametrize("prec", ["fp16x3", "fp16c8"])
@pytest.mark.parametrize("S,A,H1,H2,B,N", [(17, 6, 400, 300, 1280, 1024), (3, 1, 200, 200, 1200, 62), (5, 2, 72, 40, 1333, 257)])
def test_fused_eval_reduce_matches_composition_and_oracle(eng, kind, prec, S, A, H1, H2, B, N):
    """B >= 8 x SMs: the reduction runs in the evaluation kernel (online softmax across a state's action blocks, state-major
    tiles).  It must agree with the two-kernel composition on the same q (1e-5) and with the fp64 oracle evaluated on the
    EXACT q (loss / gradients to the precision the mode's q carries), with and without the q output."""
    import torch
    rng = np.random.RandomState(B + N + A)
    p = _rand_tin(rng, S, A, H1, H2, last=3.0)
    cr = _tin(eng, p, S, A, H1, H2)
    s = rng.randn(B, S).astype(np.float32)
    a = rng.uniform(-0.95, 0.95, (N, A)).astype(np.float32)
    w = rng.uniform(0.5, 1.5, N).astype(np.float32) / N
    mean = (rng.randn(B, A) * 0.5).astype(np.float32)
    lstd = (rng.randn(B, A) * 0.3 - 0.5).astype(np.float32)
    v = rng.randn(B).astype(np.float32)
    alpha = 0.2
    lb, dm, ds, q = cr.eval_reduce_policy(s, a, w, 1.0, mean, lstd, alpha, kind=kind, v=v, precision=prec, want_q=True,
                                          b_total=2 * B, fuse=True)
    assert eng.umma_error() == 0
    lb2, dm2, ds2, q_none = cr.eval_reduce_policy(s, a, w, 1.0, mean, lstd, alpha, kind=kind, v=v, precision=prec, b_total=2 * B,
                                                  fuse=True)
    assert q_none is None
    for x, y in ((lb, lb2), (dm, dm2), (ds, ds2)):
        assert torch.equal(x, y)                                   # the q output does not change the reduction
    assert torch.equal(q, cr.eval(s, a, prec))                     # same q as the plain evaluation, bit for bit
    t = lambda z: torch.as_tensor(z, device=eng.device)
    if kind == "fkl":
        ref = eng.fkl_policy(q, t(w), t(a), 1.0, t(mean), t(lstd), alpha, b_total=2 * B)
    else:
        ref = eng.rkl_policy(q, t(v), t(w), t(a), 1.0, t(mean), t(lstd), alpha, b_total=2 * B)
    for mine, r in zip((lb, dm, ds), ref[:3]):
        r = r.cpu().numpy()
        np.testing.assert_allclose(mine.cpu().numpy(), r, rtol=2e-4, atol=2e-5 * max(1e-6, np.abs(r).max()))
    rows = np.r_[0:8, B - 8:B]
    q64 = onp.tin_eval(s[rows], a, p, dtype=np.float64)
    if kind == "fkl":
        o = onp.fkl_policy_reduce(q64, w, a, mean[rows], lstd[rows], 1.0, alpha)
    else:
        o = onp.rkl_policy_reduce(q64, v[rows], w, a, mean[rows], lstd[rows], 1.0, alpha)
    scale = len(rows) / (2 * B)          # the oracle's gradients are for the mean over `rows`, the kernel's for 1 / b_total
    tol = 5e-3 if prec == "fp16c8" else 5e-4
    np.testing.assert_allclose(lb.cpu().numpy()[rows], o[0], rtol=tol, atol=tol * np.abs(o[0]).max())
    np.testing.assert_allclose(dm.cpu().numpy()[rows], o[1] * scale, rtol=tol, atol=tol * np.abs(o[1] * scale).max())
    np.testing.assert_allclose(ds.cpu().numpy()[rows], o[2] * scale, rtol=tol, atol=tol * np.abs(o[2] * scale).max())


def test_fused_eval_reduce_small_batch_composes(eng):
    """Below 8 x SMs states the call composes the evaluation and the reduction kernel (same API, same results); fp32 too."""
    import torch
    rng = np.random.RandomState(2)
    S, A, H1, H2, B, N = 5, 2, 72, 40, 96, 200
    p = _rand_tin(rng, S, A, H1, H2, last=3.0)
    cr = _tin(eng, p, S, A, H1, H2)
    s = rng.randn(B, S).astype(np.float32)
    a = rng.uniform(-0.95, 0.95, (N, A)).astype(np.float32)
    w = np.full(N, 2.0 / N, np.float32)
    mean, lstd = (rng.randn(B, A) * 0.5).astype(np.float32), (rng.randn(B, A) * 0.3 - 0.5).astype(np.float32)
    t = lambda z: torch.as_tensor(z, device=eng.device)
    for prec in ("fp32", "fp16x3", "auto"):
        for fuse in (True, False):
            lb, dm, ds, q = cr.eval_reduce_policy(s, a, w, 1.0, mean, lstd, 0.3, precision=prec, want_q=True, fuse=fuse)
            ref = eng.fkl_policy(cr.eval(s, a, prec), t(w), t(a), 1.0, t(mean), t(lstd), 0.3)
            assert torch.equal(lb, ref[0]) and torch.equal(dm, ref[1]) and torch.equal(ds, ref[2])
            lb2 = cr.eval_reduce_policy(s, a, w, 1.0, mean, lstd, 0.3, precision=prec, fuse=fuse)[0]
            assert torch.equal(lb, lb2)
