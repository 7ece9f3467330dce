"""Kernel-level timing of the critic evaluation at cfg4 (CUDA events on the launching stream)."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlcontrol_b200 as rb

def main():
    B = int(os.environ.get("B", 4096)); N = int(os.environ.get("N", 1024))
    S, A, H1, H2 = 17, 6, 400, 300
    eng = rb.Engine(0)
    rng = np.random.RandomState(0)
    k1, k2 = 1 / np.sqrt(S + A), 1 / np.sqrt(H1)
    p = [rng.uniform(-k1, k1, (H1, S + A)).astype(np.float32), rng.uniform(-k1, k1, H1).astype(np.float32),
         rng.uniform(-k2, k2, (H2, H1)).astype(np.float32), rng.uniform(-k2, k2, H2).astype(np.float32),
         rng.uniform(-.3, .3, (1, H2)).astype(np.float32), rng.uniform(-.3, .3, 1).astype(np.float32)]
    cr = rb.Critic(eng, rb.TIN, S, A, H1, H2).load(*p, rb.LAYOUT_OUT_IN)
    s = torch.randn(B, S, device="cuda").clamp_(-10, 10)
    a = torch.rand(N, A, device="cuda") * 2 - 1
    a_ps = torch.rand(B, N, A, device="cuda") * 2 - 1
    flops = B * 2 * S * H1 + B * N * 2 * (A * H1 + H1 * H2 + H2)
    cases = (("fp16c8 shared", a, "fp16c8"), ("fp16x3 shared", a, "fp16x3"), ("fp16 shared", a, "fp16"), ("fp16 per-state", a_ps, "fp16"), ("bf16 shared", a, "bf16"),
             ("fp32 shared (B/8)", a, "fp32"))
    if os.environ.get("ONLY"):
        cases = [c for c in cases if c[0].startswith(os.environ["ONLY"] + " shared")]
    for name, act, prec in cases:
        ss = s if prec != "fp32" else s[: B // 8]
        for _ in range(3):
            q = cr.eval(ss, act if act.dim() == 2 else act[: ss.shape[0]], prec)
        torch.cuda.synchronize()
        reps = 10
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            q = cr.eval(ss, act if act.dim() == 2 else act[: ss.shape[0]], prec)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
        ms = float(np.median(ts))
        fl = flops * ss.shape[0] / B
        chk = ""
        if os.environ.get("CHECK"):        # parity of this build/knob setting against the fp32-class split mode (same inputs)
            qr = cr.eval(ss, act if act.dim() == 2 else act[: ss.shape[0]], "fp16x3" if prec != "fp16x3" else "fp32")
            d = (torch.as_tensor(q) - torch.as_tensor(qr)).abs().max().item()
            chk = f"  max|dq|/max|q| vs {'fp16x3' if prec != 'fp16x3' else 'fp32'} = {d / torch.as_tensor(qr).abs().max().item():.2e}"
        print(f"{name:22s} {ms:8.3f} ms  (min {min(ts):.3f})  {fl/ms/1e9:8.1f} TFLOP/s  {ss.shape[0]*N/ms/1e6:8.2f} G Q-evals/s  err={eng.umma_error()}{chk}",
              flush=True)

if __name__ == "__main__":
    main()
