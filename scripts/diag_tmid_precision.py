import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import rlcontrol_b200 as rb
from oracle import oracle_np as onp
from conftest import rel_err
from test_gpu_parity import _rand_tmid
eng = rb.Engine(0)
rng = np.random.RandomState(9)
S, A, H1, H2, B, N = 17, 6, 400, 300, 161, 1024
p = _rand_tmid(rng, S, A, H1, H2)
smin, smax = -np.ones(S) * 1.5, np.ones(S) * 1.5
cr = rb.Critic(eng, rb.TMID, S, A, H1, H2, smin, smax); cr.load(*p, rb.LAYOUT_IN_OUT)
s = (rng.randn(B, S) * 2).astype(np.float32)
a = rng.uniform(-1, 1, (B, N, A)).astype(np.float32)
full = onp.tmid_eval(s, a, p, smin, smax, dtype=np.float64)
def ev(tc, gemm):
    eng.lib.rlc_tmid_tc_force(tc); eng.lib.rlc_rows_gemm_force(gemm)
    q = cr.eval(s, a, "fp32").cpu().numpy()
    eng.lib.rlc_tmid_tc_force(-1); eng.lib.rlc_rows_gemm_force(-1)
    return q
for name, tc, gm in (("rows cuda / state cuda", 0, 0), ("rows cuda / state TC-gemm", 0, -1), ("rows TC / state cuda", 2, 0), ("rows TC / state TC-gemm", 2, -1)):
    q = ev(tc, gm)
    e = rel_err(q, full)
    print(name, "max %.2e  rms %.2e  p99.9 %.2e" % (e.max(), np.sqrt((e**2).mean()), np.quantile(e, 0.999)), "worst state", np.unravel_index(e.argmax(), e.shape))
