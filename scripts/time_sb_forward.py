"""rlc_sb_forward timing sweep: hidden sizes, rows, and the aligned (cp.async.bulk) vs misaligned (fallback) weight path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import rlcontrol_b200 as rb
from rlcontrol_b200 import _lib
from rlcontrol_b200._lib import check
eng = rb.Engine(0)
dev = eng.device


def timeit(fn, reps=50):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(s); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn(s)
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s); s.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (5 * reps)


for (inp, H1, H2, O, B, mis) in [(3, 200, 200, 2, 1, 0), (3, 200, 200, 2, 1, 1), (3, 64, 64, 2, 1, 0), (3, 200, 200, 2, 32, 0),
                                 (3, 400, 400, 2, 1, 0), (3, 200, 8, 2, 1, 0), (3, 8, 200, 2, 1, 0)]:
    numel = inp * H1 + H1 + H1 * H2 + H2 + H2 * O + O
    big = torch.randn(numel + 8, device=dev) * 0.1
    theta = big[mis:mis + numel]
    x, out = torch.randn(B, inp, device=dev), torch.zeros(B, O, device=dev)
    net = (_lib.RlcSbNet * 1)()
    n = net[0]
    n.theta, n.inp, n.H1, n.H2, n.O, n.x0, n.n0, n.n1, n.out = theta.data_ptr(), inp, H1, H2, O, x.data_ptr(), inp, 0, out.data_ptr()
    us = timeit(lambda s: check(eng.lib.rlc_sb_forward(eng.h, net, 1, B, C.c_void_p(s.cuda_stream))))
    print(json.dumps(dict(inp=inp, H1=H1, H2=H2, O=O, B=B, misaligned=mis, us=round(us, 2))))

# in-kernel %globaltimer trace of one launch (CTA 0, thread 0): slots 0 start, 1 inputs staged, 2 layer 1 done,
# 8+2c / 9+2c / 10+2c = layer-2 chunk c: data arrived / computed / CTA past the barrier, 3 layer 2 done, 4 output layer done
dbg = torch.zeros(128, dtype=torch.int64, device=dev)
os.environ["RLC_SB_DEBUG"] = hex(dbg.data_ptr())
inp, H1, H2, O, B = 3, 200, 200, 2, int(os.environ.get('TRACE_B', 1))
numel = inp * H1 + H1 + H1 * H2 + H2 + H2 * O + O
theta = torch.randn(numel, device=dev) * 0.1
x, out = torch.randn(B, inp, device=dev), torch.zeros(B, O, device=dev)
net = (_lib.RlcSbNet * 1)()
n = net[0]
n.theta, n.inp, n.H1, n.H2, n.O, n.x0, n.n0, n.n1, n.out = theta.data_ptr(), inp, H1, H2, O, x.data_ptr(), inp, 0, out.data_ptr()
for _ in range(3):
    check(eng.lib.rlc_sb_forward(eng.h, net, 1, B, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
t = dbg.cpu().numpy()
t0 = t[0]
print("trace ns:", {k: int(t[k] - t0) for k in (1, 2, 3, 4)})
print("layer-2 chunks (arrived, computed, past barrier):", [(int(t[8 + 2 * c] - t0), int(t[9 + 2 * c] - t0), int(t[10 + 2 * c] - t0)) for c in range(13)])
