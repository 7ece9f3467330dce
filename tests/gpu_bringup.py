"""GPU bring-up of the tcgen05 critic kernel (run under gpurun; not a pytest file).

For each RLC_UMMA_VARIANT (bit0: LBO/SBO roles, bit1: rank->B-half mapping) a child process runs
(a) a 64-64 net with identity W2 and one-hot w3 so q exposes layer-1 features one at a time,
(b) random 64-64 and 400-300 nets, and dumps everything to gpurun_out/ for offline analysis."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def child(variant):
    import torch
    import rlcontrol_b200 as rb
    from oracle import oracle_np as onp
    eng = rb.Engine(0)
    rng = np.random.RandomState(0)
    res = {}

    def run(name, S, A, H1, H2, B, N, W1, b1, W2, b2, W3, b3, per_state=False, prec="fp16"):
        cr = rb.Critic(eng, rb.TIN, S, A, H1, H2)
        cr.load(W1, b1, W2, b2, W3, b3, rb.LAYOUT_OUT_IN)
        s = rng.randn(B, S).astype(np.float32)
        a = rng.uniform(-1, 1, size=(B, N, A) if per_state else (N, A)).astype(np.float32)
        ref = onp.tin_eval(s, a, (W1, b1, W2, b2, W3, b3), dtype=np.float64)
        q32 = cr.eval(s, a, "fp32").cpu().numpy()
        t0 = time.time()
        q = cr.eval(s, a, prec)
        torch.cuda.synchronize()
        err = eng.umma_error()
        q = q.cpu().numpy()
        scale = max(np.abs(ref).max(), 1e-9)
        e32 = np.abs(q32 - ref).max() / scale
        e16 = np.abs(q - ref).max() / scale
        print(f"[v{variant}] {name}: fp32 relerr {e32:.2e}  {prec} relerr {e16:.3e}  umma_err {err} "
              f"({time.time()-t0:.2f}s)", flush=True)
        res[name + "_q"] = q
        res[name + "_ref"] = ref
        return q, ref

    # (a) layer-1 probe: H1=H2=64, W2=I, b2=0, w3=e_j
    S, A, H1, H2 = 3, 1, 64, 64
    W1 = rng.randn(H1, S + A).astype(np.float32) * 0.5
    b1 = rng.randn(H1).astype(np.float32) * 0.1
    for j in (0, 1, 8, 31, 32, 33, 63):
        W3 = np.zeros((1, H2), np.float32)
        W3[0, j] = 1.0
        run(f"probe_j{j}", S, A, H1, H2, 4, 128, W1, b1, np.eye(H2, H1, dtype=np.float32),
            np.zeros(H2, np.float32), W3, np.zeros(1, np.float32))
    # (b) random nets
    def rnd(S, A, H1, H2):
        return (rng.randn(H1, S + A).astype(np.float32) / np.sqrt(S + A),
                rng.randn(H1).astype(np.float32) * 0.1,
                rng.randn(H2, H1).astype(np.float32) / np.sqrt(H1),
                rng.randn(H2).astype(np.float32) * 0.1,
                rng.randn(1, H2).astype(np.float32) / np.sqrt(H2),
                rng.randn(1).astype(np.float32))
    run("rand64", 3, 1, 64, 64, 4, 128, *rnd(3, 1, 64, 64))
    run("rand200", 3, 1, 200, 200, 32, 62, *rnd(3, 1, 200, 200))
    run("rand400_300", 17, 6, 400, 300, 16, 1024, *rnd(17, 6, 400, 300))
    run("rand400_300_ps", 17, 6, 400, 300, 8, 300, *rnd(17, 6, 400, 300), per_state=True)
    run("rand400_300_bf16", 17, 6, 400, 300, 16, 1024, *rnd(17, 6, 400, 300), prec="bf16")
    np.savez_compressed(os.path.join(OUT, f"bringup_v{variant}.npz"), **res)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
        return
    for v in (0, 1, 2, 3):
        env = dict(os.environ, RLC_UMMA_VARIANT=str(v))
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), str(v)], env=env,
                               timeout=240, capture_output=True, text=True)
            print(p.stdout[-6000:])
            if p.returncode != 0:
                print(f"[v{v}] exit {p.returncode}\n{p.stderr[-3000:]}")
        except subprocess.TimeoutExpired:
            print(f"[v{v}] TIMEOUT")


if __name__ == "__main__":
    main()
