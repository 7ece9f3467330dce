"""ncu target: a few launches of the cfg4-size training GEMMs on the tensor path (rlc_rows_gemm path=2)."""
import sys
import torch
sys.path.insert(0, ".")
import rlcontrol_b200 as rb
from rlcontrol_b200._lib import check
from rlcontrol_b200.engine import _ptr, _stream
eng = rb.Engine(0)
dev = eng.device
B = 4096
for ta, tb, M, N, K, split in [(0, 0, B, 300, 400, 0), (0, 1, B, 400, 300, 0), (1, 0, 400, 300, B, 1), (0, 0, B, 400, 23, 0)]:
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    Bm = torch.randn((N, K) if tb else (K, N), device=dev)
    Cc = torch.empty((M, N), device=dev)
    for _ in range(3):
        check(eng.lib.rlc_rows_gemm(eng.h, ta, tb, M, N, K, _ptr(A), A.stride(0), _ptr(Bm), Bm.stride(0), _ptr(Cc), N,
                                    None, None, 0, 0, 1.0, split, 2, _stream()))
    torch.cuda.synchronize()
assert eng.umma_error() == 0
