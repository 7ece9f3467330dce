"""CUDA-event timing of the pieces of one cfg1 / cfg4-exact update (back-to-back launches of each piece)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlcontrol_b200 as rb
from rlcontrol_b200 import quadrature

def timeit(fn, reps=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

eng = rb.Engine(0)
rng = np.random.RandomState(0)
t = lambda x: torch.as_tensor(np.asarray(x, np.float32), device="cuda")
for name, B, npar, H1, H2 in (("cfg1", 32, 64, 200, 200), ("cfg4x", 4096, 1026, 400, 300)):
    S, A = 3, 1
    acts, w = quadrature.grid_1d(npar, 2.0)
    N = acts.shape[0]
    u = lambda k, *sh: rng.uniform(-k, k, sh).astype(np.float32)
    cr = rb.Critic(eng, rb.TIN, S, A, H1, H2).load(u(.5, H1, S + A), u(.5, H1), u(.07, H2, H1), u(.07, H2), u(.3, 1, H2), u(.3, 1), rb.LAYOUT_OUT_IN)
    s, a, y = t(rng.randn(B, S)), t(rng.uniform(-2, 2, (B, A))), t(rng.randn(B))
    grid, wd = t(acts), t(w)
    q = torch.empty((B, N), device="cuda")
    print(name, "grid eval fp32: %.1f us" % timeit(lambda: cr.eval_into(s, grid, q, "fp32")))
    try:
        print(name, "grid eval fp16 (pack cached): %.1f us" % timeit(lambda: cr.eval_into(s, grid, q, "fp16")))
        def rp():
            cr.invalidate(); cr.eval_into(s, grid, q, "fp16")
        print(name, "grid eval fp16 + repack: %.1f us" % timeit(rp))
    except Exception as e:
        print(name, "fp16 path:", e)
    a1 = a.view(B, 1, A); q1 = torch.empty((B, 1), device="cuda")
    print(name, "Q(s,a_new) B rows: %.1f us" % timeit(lambda: cr.eval_into(s, a1, q1, "fp32")))
    g = torch.empty_like(cr.theta); loss = torch.empty(1, device="cuda"); qo = torch.empty(B, device="cuda")
    print(name, "critic grads: %.1f us" % timeit(lambda: cr.grads_into(s, a, y, g, loss, qo)))
    m = rb.Mlp(eng, S, H1, H2, 2).load_torch(u(.5, H1, S), u(.5, H1), u(.07, H2, H1), u(.07, H2), u(.3, 2, H2), u(.3, 2))
    act = m.act_buffer(B); out = torch.empty((B, 2), device="cuda"); dout = t(rng.randn(B, 2)); gm = torch.empty_like(m.theta)
    print(name, "mlp forward: %.1f us" % timeit(lambda: m.forward(s, out=out, act=act)))
    print(name, "mlp grads: %.1f us" % timeit(lambda: m.grads(s, dout, act=act, grad_out=gm)))
    mm, vv, sd = torch.zeros_like(g), torch.zeros_like(g), torch.zeros(4, dtype=torch.int32, device="cuda")
    print(name, "adam_dev: %.1f us" % timeit(lambda: eng.adam_step_dev(cr.theta, g, mm, vv, sd, 1e-9)))
    mean, ls, v = t(rng.randn(B, A) * .3), t(rng.randn(B, A) * .3 - 1), t(rng.randn(B))
    print(name, "rkl_policy reduce: %.1f us" % timeit(lambda: eng.rkl_policy(q, v, wd, grid, 2.0, mean, ls, 0.1)))
