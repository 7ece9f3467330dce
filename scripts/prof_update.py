"""Eager (no graph) cfg1 / cfg4-exact full updates for an ncu launch list: python scripts/prof_update.py [cfg1|cfg4]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import rlcontrol_b200 as rb
from rlcontrol_b200 import kl_networks
from bench_updates import cfg
which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
eng = rb.Engine(0)
c = cfg(eng, 3, 1, 2.0, 32, 64, 200, 200, use_cuda_graph=False) if which == "cfg1" else \
    cfg(eng, 3, 1, 2.0, 4096, 1026, 400, 300, use_cuda_graph=False)
net = kl_networks.ReverseKLNetwork(None, None, c)
rng = np.random.RandomState(0)
B, S, A = c.batch_size, 3, 1
for i in range(3):
    net.update_network(rng.randn(B, S), rng.uniform(-2, 2, (B, A)), rng.randn(B, S), rng.randn(B), np.full(B, 0.99))
    net.update_target_network()
torch.cuda.synchronize()
print("ok", net.last_losses)
