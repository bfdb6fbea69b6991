"""One-shape driver for ncu captures of the tensor-core GEMM."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops  # noqa: E402

M, N, K = 34445, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 768
act = int(sys.argv[2]) if len(sys.argv) > 2 else 1
A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
ops.f16_split(W, hi, lo)
C = torch.empty(M, N, device="cuda")
for _ in range(5):
    ops.tc_gemm(A, hi, lo, C, act=act)
torch.cuda.synchronize()
