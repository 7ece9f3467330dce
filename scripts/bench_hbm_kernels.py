"""Achieved bytes/s of the memory-bound kernels against the measured HBM copy bandwidth (MEASURED_PEAKS.json),
on inputs larger than L2 (rotating rings).  One JSON line per kernel: algorithmic bytes / CUDA-event time."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rlcontrol_b200 as rb

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
RING = 10


ONLY = os.environ.get("HBM_ONLY", "")          # substring filter on the kernel names, e.g. HBM_ONLY=K6
NCU = os.environ.get("HBM_NCU", "") == "1"      # under ncu: two eager calls per kernel, no graph, no timing


def timeit(fn, reps=20):
    """Device time per call: the calls are captured into one CUDA graph (RING calls over the rotating buffers) and the
    graph is replayed, so the number is kernel time, not the Python/ctypes issue rate (a few us per call, which is
    longer than several of these kernels)."""
    st = torch.cuda.Stream()
    if NCU:
        with torch.cuda.stream(st):
            fn(0); fn(1)
        st.synchronize()
        return 1.0
    with torch.cuda.stream(st):
        for i in range(RING):
            fn(i)
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for i in range(RING):
            fn(i)
    with torch.cuda.stream(st):
        g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
    st.synchronize()
    return e0.elapsed_time(e1) / (reps * RING) * 1e-3


def report(name, nbytes, sec, note=""):
    if ONLY and ONLY not in name:
        return
    sec = sec() if callable(sec) else sec
    note = note(sec) if callable(note) else note
    gbs = nbytes / sec / 1e9
    print(json.dumps({"kernel": name, "algorithmic_bytes": int(nbytes), "us": round(sec * 1e6, 2), "GB_per_s": round(gbs, 1),
                      "hbm_peak_GB_per_s": PEAK, "frac": round(gbs / PEAK, 3), "note": note}), flush=True)


def main():
    eng = rb.Engine(0)
    dev = eng.device
    B, N, A, k = 4096, 1024, 6, 6
    g = torch.Generator(device="cuda").manual_seed(0)
    qs = [torch.randn(B, N, device=dev, generator=g) for _ in range(RING)]            # 10 x 16.8 MB > 126 MB L2
    lps = [torch.randn(B, N, device=dev, generator=g) for _ in range(RING)]
    w = torch.rand(N, device=dev, generator=g)
    v = torch.randn(B, device=dev, generator=g)
    report("K3 rlc_reduce_topk k=6 (idx + q_sel, no gather) q[4096,1024]", B * N * 4 + B * k * 12,
           lambda: timeit(lambda i: eng.topk(qs[i], k)), "k selection rounds over a register-resident row")
    report("K3 rlc_reduce_stats (argmax, max, mean)", B * N * 4 + B * 16, lambda: timeit(lambda i: eng.stats(qs[i])))
    report("K3 rlc_reduce_lse (SQL soft value)", B * N * 4 + B * 4, lambda: timeit(lambda i: eng.soft_value(qs[i], A)))
    report("K3 rlc_reduce_fkl (q, logp in; loss_b out)", 2 * B * N * 4 + B * 4,
           lambda: timeit(lambda i: eng.fkl(qs[i], w, lps[i], 0.1, want_boltz=False, want_grad=False)))
    report("K3 rlc_reduce_fkl (q, logp in; loss_b, boltz, dlogp out)", 4 * B * N * 4 + B * 4,
           lambda: timeit(lambda i: eng.fkl(qs[i], w, lps[i], 0.1)))
    report("K3 rlc_reduce_rkl (q, logp in; loss_b, dlogp out)", 3 * B * N * 4 + B * 8,
           lambda: timeit(lambda i: eng.rkl(qs[i], v, w, lps[i], 0.1)))
    grid = torch.rand(N, A, device=dev, generator=g) * 1.9 - 0.95
    mean = torch.randn(B, A, device=dev, generator=g) * 0.5
    lstd = torch.randn(B, A, device=dev, generator=g) * 0.3 - 0.5
    outp = (torch.empty(B, device=dev), torch.empty(B, A, device=dev), torch.empty(B, A, device=dev))
    report("K3 rlc_reduce_fkl_policy (q in; loss_b, dmean, dlog_std out; log-density in place, A=6)", B * N * 4 + B * (1 + 2 * A) * 4,
           lambda: timeit(lambda i: eng.fkl_policy(qs[i], w, grid, 1.0, mean, lstd, 0.1, out=outp)), "k_grid_logterms + k_policy_reduce")
    report("K3 rlc_reduce_rkl_policy (same, ReverseKL)", B * N * 4 + B * (2 + 2 * A) * 4,
           lambda: timeit(lambda i: eng.rkl_policy(qs[i], v, w, grid, 1.0, mean, lstd, 0.1, out=outp)), "k_grid_logterms + k_policy_reduce")
    # K2: hoisted T-mid evaluation with per-state actions (cfg3 critic at cfg4 size): actions in, q out
    S, H1, H2 = 17, 400, 300
    rng = np.random.RandomState(0)
    u = lambda kk, *sh: rng.uniform(-kk, kk, sh).astype(np.float32)
    cr = rb.Critic(eng, rb.TMID, S, A, H1, H2).load(u(.4, S, H1), u(.4, H1), u(.09, H1 + A, H2), u(.09, H2), u(.3, H2, 1), u(.3, 1),
                                                     rb.LAYOUT_IN_OUT)
    s = torch.randn(B, S, device=dev, generator=g)
    acts = [torch.rand(B, N, A, device=dev, generator=g) * 2 - 1 for _ in range(RING)]
    report("K2 T-mid hoisted evaluation, per-state actions [4096,1024,6] (state term + rows)", B * N * (A + 1) * 4,
           lambda: timeit(lambda i: cr.eval_into(s, acts[i], qs[i], "fp32"), 5),
           lambda sec: "also %.1f TFLOP/s fp32 on 2(A+1)H2 flop/row: issue-bound, not HBM-bound" % (B * N * 2 * (A + 1) * H2 / sec / 1e12))
    # K5 input gradient dQ/da on the whole B x N stack (AE+ ascent ae_plus_network.py:310-343, SQL sql_network.py:101-107):
    # T-mid fused evaluation + gradient (hoisted, per-state actions), and the T-in row kernel on a stacked slice
    gs = [torch.empty(B, N, A, device=dev) for _ in range(2)]

    def tmid_grad(i):
        from rlcontrol_b200._lib import check as _chk
        from rlcontrol_b200.engine import _ptr as _pp, _stream as _ss
        import ctypes as _C
        _chk(eng.lib.rlc_tmid_eval_grad(eng.h, _C.byref(cr._desc), _pp(s), B, _pp(acts[i]), N, rb.ACT_PER_STATE, _pp(qs[i]),
                                        _pp(gs[i % 2]), _ss()))
    report("K5 dQ/da T-mid on the B x N stack [4096,1024,6] (rlc_tmid_eval_grad: q and dq/da out)", B * N * (2 * A + 1) * 4,
           lambda: timeit(tmid_grad, 3),
           lambda sec: "%.2f G rows/s; %.1f TFLOP/s fp32 on 4(A+1)H2 flop/row" % (B * N / sec / 1e9, B * N * 4 * (A + 1) * H2 / sec / 1e12))
    crin = rb.Critic(eng, rb.TIN, S, A, H1, H2).load(u(.2, H1, S + A), u(.2, H1), u(.05, H2, H1), u(.05, H2), u(.3, 1, H2), u(.3, 1),
                                                      rb.LAYOUT_OUT_IN)
    R = B * N // 8                                        # 524 288 stacked rows (an eighth of the cfg4 stack; fp32 CUDA cores)
    s_rows = torch.randn(R, S, device=dev, generator=g)
    a_rows = torch.rand(R, A, device=dev, generator=g) * 2 - 1
    g_rows, q_rows = torch.empty(R, A, device=dev), torch.empty(R, device=dev)

    def tin_grad(i):
        from rlcontrol_b200._lib import check as _chk
        from rlcontrol_b200.engine import _ptr as _pp, _stream as _ss
        import ctypes as _C
        _chk(eng.lib.rlc_critic_grad_action(eng.h, _C.byref(crin._desc), _pp(s_rows), _pp(a_rows), R, _pp(g_rows), _pp(q_rows), _ss()))
    fl_row = 2 * ((S + A) * H1 + H1 * H2 + H2) + 2 * (H2 + H1 * H2 + A * H1)
    report("K5 dQ/da T-in on %d stacked rows (rlc_critic_grad_action, row GEMMs on tcgen05 3xTF32)" % R, R * (S + 2 * A + 1) * 4,
           lambda: timeit(tin_grad, 2),
           lambda sec: "%.3f G rows/s = %.1f TFLOP/s fp32 (forward + input-gradient, %d flop/row); the full 4.19M-row stack takes %.1f ms"
                       % (R / sec / 1e9, R * fl_row / sec / 1e12, fl_row, sec * 8 * 1e3))
    # K6: replay gather on a 4M-slot ring (S=17, A=6): 168 B per sampled transition in, the same out
    from rlcontrol_b200.replaybuffer import ReplayBuffer
    cap, nb = 1 << 22, 1 << 20
    st_ = torch.randn(cap, S, device=dev, generator=g); ac_ = torch.randn(cap, A, device=dev, generator=g)
    rw_ = torch.randn(cap, device=dev, generator=g); s2_ = torch.randn(cap, S, device=dev, generator=g); gm_ = torch.rand(cap, device=dev, generator=g)
    idxs = [torch.randint(0, cap, (nb,), device=dev, generator=g) for _ in range(4)]
    outs = (torch.empty(nb, S, device=dev), torch.empty(nb, A, device=dev), torch.empty(nb, device=dev),
            torch.empty(nb, S, device=dev), torch.empty(nb, device=dev))
    import ctypes as C
    from rlcontrol_b200.engine import _ptr, _stream
    from rlcontrol_b200._lib import check

    def gather(i):
        check(eng.lib.rlc_replay_gather(eng.h, _ptr(st_), _ptr(ac_), _ptr(rw_), _ptr(s2_), _ptr(gm_), cap, S, A, _ptr(idxs[i % 4]), nb,
                                        *[_ptr(o) for o in outs], _stream()))
    row = (2 * S + A + 2) * 4
    report("K6 rlc_replay_gather, 1M random transitions from a 4M-slot ring (S=17, A=6)", nb * (2 * row + 8), lambda: timeit(gather, 5),
           "random 68/24/4-byte segments: sector-granular reads, so the DRAM traffic exceeds the algorithmic bytes")
    # the same gather from the record layout (one 64-byte-aligned 192-byte record per transition)
    lib = eng.lib
    stride = lib.rlc_replay_rec_stride(S, A)
    rec = torch.randn(cap, stride, device=dev, generator=g)

    def gather_rec(i):
        check(lib.rlc_replay_gather_rec(eng.h, _ptr(rec), cap, stride, S, A, _ptr(idxs[i % 4]), nb, *[_ptr(o) for o in outs], _stream()))
    report("K6 rlc_replay_gather_rec, same gather from 64-byte-aligned records (stride %d floats)" % stride, nb * (2 * row + 8),
           lambda: timeit(gather_rec, 5), "one contiguous %d-byte DRAM read per transition" % (stride * 4))


if __name__ == "__main__":
    main()
