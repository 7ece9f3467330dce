"""Device time of the B-row training GEMMs at cfg4 size (B=4096, 23 -> 400 -> 300) on both paths of rlc_rows_gemm
(1 = fp32 CUDA cores, 2 = tcgen05 3xTF32), and of the entry points built from them.  CUDA events around `reps` back-to-back
launches on the current stream, after warm-up.  One JSON line per measurement."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import rlcontrol_b200 as rb  # noqa: E402
from rlcontrol_b200._lib import check  # noqa: E402
from rlcontrol_b200.engine import _ptr, _stream  # noqa: E402


def timed(fn, reps=20, warm=3):
    """us per call: `reps` calls captured in one CUDA graph (no host launch cost in the number), 5 replays timed."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(warm):
            fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            g.replay()
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)


def main():
    eng = rb.Engine(0)
    dev = eng.device
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    g = torch.Generator(device="cpu").manual_seed(0)
    shapes = [  # name, ta, tb, M, N, K, split
        ("fwd L1  X[B,23] W1[23,400]", 0, 0, B, 400, 23, 0),
        ("fwd L2  H1[B,400] W2[400,300]", 0, 0, B, 300, 400, 0),
        ("bwd G1  G2[B,300] W2^T", 0, 1, B, 400, 300, 0),
        ("gW2     H1^T[400,B] G2[B,300] (split-K)", 1, 0, 400, 300, B, 1),
        ("gW1     X^T[23,B] G1[B,400] (split-K)", 1, 0, 23, 400, B, 1),
    ]
    for name, ta, tb, M, N, K, split in shapes:
        A = torch.randn((K, M) if ta else (M, K), generator=g).to(dev)
        Bm = torch.randn((N, K) if tb else (K, N), generator=g).to(dev)
        Cc = torch.empty((M, N), device=dev)
        row = {"gemm": name, "M": M, "N": N, "K": K}
        for path, label in ((1, "cuda_core_us"), (2, "tcgen05_us")):
            def call():
                check(eng.lib.rlc_rows_gemm(eng.h, ta, tb, M, N, K, _ptr(A), A.stride(0), _ptr(Bm), Bm.stride(0), _ptr(Cc), N,
                                            None, None, 0, 0, 1.0, split, path, _stream()))
            row[label] = round(timed(call), 2)
        row["tcgen05_tflops_alg"] = round(2.0 * M * N * K / row["tcgen05_us"] * 1e-6, 1)
        print(json.dumps(row), flush=True)
    assert eng.umma_error() == 0

    # entry points at cfg4 size
    rng = np.random.RandomState(0)
    S, A_, H1, H2 = 17, 6, 400, 300
    k1, k2 = 1 / np.sqrt(S + A_), 1 / np.sqrt(H1)
    p = [rng.uniform(-k1, k1, (H1, S + A_)).astype(np.float32), rng.uniform(-k1, k1, H1).astype(np.float32),
         rng.uniform(-k2, k2, (H2, H1)).astype(np.float32), rng.uniform(-k2, k2, H2).astype(np.float32),
         rng.uniform(-.003, .003, (1, H2)).astype(np.float32), rng.uniform(-.003, .003, 1).astype(np.float32)]
    cr = rb.Critic(eng, rb.TIN, S, A_, H1, H2).load(*p, rb.LAYOUT_OUT_IN)
    s = torch.randn((B, S), generator=g).to(dev)
    a = (torch.rand((B, A_), generator=g) * 2 - 1).to(dev)
    y = torch.randn((B,), generator=g).to(dev)
    grad = torch.empty_like(cr.theta)
    loss = torch.empty((1,), device=dev)
    q = torch.empty((B,), device=dev)
    m = rb.Mlp(eng, S, H1, H2, 2 * A_)
    m.theta.copy_(torch.randn(m.theta.shape, generator=g) * 0.05)
    act = m.act_buffer(B)
    out = torch.empty((B, 2 * A_), device=dev)
    dout = (torch.randn((B, 2 * A_), generator=g) / B).to(dev)
    mg = torch.empty_like(m.theta)
    R = 16384
    sr = torch.randn((R, S), generator=g).to(dev)
    ar = (torch.rand((R, A_), generator=g) * 2 - 1).to(dev)
    for force, label in ((0, "cuda_core_us"), (-1, "dispatcher_us")):
        eng.lib.rlc_rows_gemm_force(force)
        res = {
            "rlc_critic_grads B=%d" % B: timed(lambda: cr.grads_into(s, a, y, grad, loss, q), 30),
            "rlc_mlp_forward B=%d (policy net, O=12)" % B: timed(lambda: m.forward(s, out=out, act=act), 30),
            "rlc_mlp_grads B=%d (policy net, O=12)" % B: timed(lambda: m.grads(s, dout, act=act, grad_out=mg), 30),
            "rlc_critic_grad_action R=%d (T-in dQ/da)" % R: timed(lambda: cr.grad_action(sr, ar), 5),
        }
        for k, v in res.items():
            print(json.dumps({"entry": k, label: round(v, 2)}), flush=True)
    eng.lib.rlc_rows_gemm_force(-1)
    assert eng.umma_error() == 0


if __name__ == "__main__":
    main()
