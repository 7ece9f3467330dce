import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlcontrol_b200 as rb
eng = rb.Engine(0); dev = eng.device
B, N, A = 4096, 1024, 6
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(B, N, device=dev, generator=g); w = torch.rand(N, device=dev, generator=g)
grid = torch.rand(N, A, device=dev, generator=g) * 1.9 - 0.95
mean = torch.randn(B, A, device=dev, generator=g) * 0.5; lstd = torch.randn(B, A, device=dev, generator=g) * 0.3 - 0.5
for _ in range(6):
    eng.fkl_policy(q, w, grid, 1.0, mean, lstd, 0.1)
torch.cuda.synchronize(); print("ok")
