"""One-shape driver for ncu captures of the tensor-core GEMM (edge GEMM 1 with its gather epilogue)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops  # noqa: E402

M, N, K = 34445, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 768
A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
hi, lo = torch.empty_like(W), torch.empty_like(W)
ops.tf32_split(W, hi, lo)
C = torch.empty(M, N, device="cuda")
P = torch.randn(2643, 1024, device="cuda")
Cb = torch.randn(256, 512, device="cuda")
i1 = torch.randint(0, 2643, (M,), device="cuda", dtype=torch.int32)
i2 = torch.randint(0, 2643, (M,), device="cuda", dtype=torch.int32)
i3 = torch.randint(0, 256, (M,), device="cuda", dtype=torch.int32)
for _ in range(5):
    ops.tc_gemm(A, hi, lo, C, gathers=[(P[:, :512], i1), (P[:, 512:], i2), (Cb, i3)], act=ops.ACT_SILU)
torch.cuda.synchronize()
