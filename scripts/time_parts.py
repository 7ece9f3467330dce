"""CUDA-event timing of the individual calls of one bench step (cfg4), 50 reps each."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlcontrol_b200 as rb
import bench

def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

W = bench.WORKLOAD
B, N, S, A, H1, H2 = W["B_per_gpu"], W["N"], W["S"], W["A"], W["H1"], W["H2"]
eng = rb.Engine(0)
rng = np.random.RandomState(0)
p = bench.make_params(rng, S, A, H1, H2)
s, a, w, (m, ls) = bench.make_inputs(rng, B, N, S, A)
cr = rb.Critic(eng, rb.TIN, S, A, H1, H2).load(*p, rb.LAYOUT_OUT_IN)
t = lambda x: torch.as_tensor(x, device="cuda")
sd, ad, wd, md, lsd = t(s), t(a), t(w), t(m), t(ls)
q = torch.empty((B, N), device="cuda")
print("K1 eval_into (k_grid_parts + k_critic_umma_grid): %.1f us" % timeit(lambda: cr.eval_into(sd, ad, q, "fp16")))
print("K3 fkl_policy (k_grid_logterms + k_policy_reduce): %.1f us" % timeit(lambda: eng.fkl_policy(q, wd, ad, 1.0, md, lsd, 0.1)))
lp = torch.randn(B, N, device="cuda")
print("K3 fkl (unfused, logp tensor): %.1f us" % timeit(lambda: eng.fkl(q, wd, lp, 0.1, want_boltz=False, want_grad=False)))
print("K3 topk k=6: %.1f us" % timeit(lambda: eng.topk(q, 6)))
print("K3 stats: %.1f us" % timeit(lambda: eng.stats(q)))
print("K3 lse: %.1f us" % timeit(lambda: eng.soft_value(q, A)))
def repack():
    cr.invalidate(); cr.eval_into(sd, ad, q, "fp16")
print("K1 with operand repack: %.1f us" % timeit(repack))
aps = torch.rand(B, N, A, device="cuda") * 2 - 1
print("K1 per-state actions (TS kernel): %.1f us" % timeit(lambda: cr.eval_into(sd, aps, q, "fp16"), 20))
