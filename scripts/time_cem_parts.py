"""Where the CEM kernel's time goes: iterations x mixture size (device time via graph replay)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import rlcontrol_b200 as rb
from bench_configs import tmid_params
eng = rb.Engine(0)
rng = np.random.RandomState(0)
S, A, H1, H2, B, N, top_m = 17, 6, 400, 300, 256, 1024, 6
p = tmid_params(rng, S, A, H1, H2)
cr = rb.Critic(eng, rb.TMID, S, A, H1, H2, -10 * np.ones(S), 10 * np.ones(S)).load(*p, rb.LAYOUT_IN_OUT)
t = lambda x: torch.as_tensor(x, device="cuda")
sd = t(rng.randn(B, S).astype(np.float32)); u0 = t(rng.uniform(size=(B, N, A)).astype(np.float32))
def dev_time(fn, reps=20):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); fn()
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(5): fn()
    with torch.cuda.stream(st):
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): g.replay()
        e1.record()
    st.synchronize()
    return e0.elapsed_time(e1) / (reps * 5) * 1e3
amin, amax = t(-np.ones(A, np.float32)), t(np.ones(A, np.float32))
if os.environ.get("CEM_ONCE") == "1":          # under ncu (scripts/ncu_cem.sh): two eager cfg3 calls, nothing else
    noise = t(rng.randn(2, B, N, A).astype(np.float32)); cu = t(rng.uniform(size=(2, B, N)).astype(np.float32))
    for _ in range(2):
        cr.cem(sd, u0, noise, cu, top_m, 2, -np.ones(A), np.ones(A))
    torch.cuda.synchronize()
    sys.exit(0)
for iters in (1, 2, 3):
    noise = t(rng.randn(iters - 1, B, N, A).astype(np.float32)) if iters > 1 else None
    cu = t(rng.uniform(size=(iters - 1, B, N)).astype(np.float32)) if iters > 1 else None
    for M in (1, 2):
        print("iters %d M %d: %.1f us" % (iters, M, dev_time(lambda: cr.cem(sd, u0, noise, cu, top_m, M, -np.ones(A), np.ones(A)))))
q = torch.empty(B, N, device="cuda")
print("eval alone (state term + rows kernel) B=256,N=1024: %.1f us" % dev_time(lambda: cr.eval_into(sd, u0, q, "fp32")))
print("topk alone: %.1f us" % dev_time(lambda: eng.topk(q, top_m)))
