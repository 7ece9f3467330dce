"""Device time of the T-mid stack evaluation (rlc_critic_eval, RLC_TMID) at cfg4 stack size, per path, and of the state
term alone (N = 1).  CUDA graph of 10 calls over rotating action buffers, CUDA events."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import rlcontrol_b200 as rb

def timed(fn, reps=10):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3):
            fn(i)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(reps):
                fn(i)
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            g.replay()
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)

eng = rb.Engine(0)
dev = eng.device
B, N, A, S, H1, H2 = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 6, 17, 400, 300
rng = np.random.RandomState(0)
u = lambda kk, *sh: rng.uniform(-kk, kk, sh).astype(np.float32)
cr = rb.Critic(eng, rb.TMID, S, A, H1, H2).load(u(.4, S, H1), u(.4, H1), u(.09, H1 + A, H2), u(.09, H2), u(.3, H2, 1), u(.3, 1), rb.LAYOUT_IN_OUT)
s = torch.randn(B, S, device=dev)
acts = [torch.rand(B, N, A, device=dev) * 2 - 1 for _ in range(10)]
qs = [torch.empty(B, N, device=dev) for _ in range(10)]
a1 = torch.rand(B, 1, A, device=dev)
q1 = torch.empty(B, 1, device=dev)
t_state = timed(lambda i: cr.eval_into(s, a1, q1, "fp32"))
out = {"B": B, "N": N, "state_term_plus_4096_rows_us": round(t_state, 1)}
for mode, name in ((0, "cuda_core_us"), (2, "tcgen05_us")):
    eng.lib.rlc_tmid_tc_force(mode)
    out[name] = round(timed(lambda i: cr.eval_into(s, acts[i], qs[i], "fp32")), 1)
eng.lib.rlc_tmid_tc_force(-1)
assert eng.umma_error() == 0
out["rows_per_s_tcgen05"] = round(B * N / out["tcgen05_us"] * 1e6 / 1e9, 2)
print(json.dumps(out))
