"""Where a device-resident run spends host time: feed drawing, graph launches, log draining (per chunk)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.argv = [sys.argv[0], "1000", "1"]
import importlib.util
spec_ = importlib.util.spec_from_file_location("bdl", os.path.join(os.path.dirname(__file__), "bench_device_loop.py"))
src = open(spec_.origin).read().split("dt1, res1 = timed")[0]
ns = {"__file__": spec_.origin}
exec(compile(src, spec_.origin, "exec"), ns)
make = ns["make"]
e = make(0)
e._build()
e.begin()
torch.cuda.synchronize()
K = e.K
for rep in range(3):
    t0 = time.perf_counter(); learn = e._fill_feeds(rep & 1, K); t1 = time.perf_counter()
    with torch.cuda.stream(e.stream):
        for i in range(K):
            e.g_learn.replay()
    t2 = time.perf_counter()
    e.stream.synchronize(); t3 = time.perf_counter()
    with torch.cuda.stream(e.stream):
        e.g_eval.replay()
    t4 = time.perf_counter(); e.stream.synchronize(); t5 = time.perf_counter()
    print(f"chunk of {K}: fill_feeds {1e6*(t1-t0)/K:.1f} us/step, enqueue {1e6*(t2-t1)/K:.1f} us/step, "
          f"device (after enqueue) {1e3*(t3-t2):.1f} ms -> total {1e6*(t3-t1)/K:.1f} us/step; eval session launch {1e6*(t4-t3):.0f} us, run {1e3*(t5-t4):.2f} ms")
