"""Micro-benchmark of the two GEMM kernels on the hot shapes (CUDA events, L2 flushed between launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops  # noqa: E402

shapes = [("edge1", 34445, 512, 768), ("edge2", 34445, 512, 512), ("pq", 2643, 1024, 512), ("node1", 2643, 512, 1024),
          ("node2", 2643, 512, 512), ("edge1@32", 4400, 512, 768)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, M, N, K in shapes:
    A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
    ops.f16_split(W, hi, lo)
    C = torch.empty(M, N, device="cuda")
    for kind in ("ffma", "tc"):
        ts = []
        for it in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if kind == "ffma":
                ops.sgemm(A, W, C, act=ops.ACT_SILU)
            else:
                ops.tc_gemm(A, hi, lo, C, act=ops.ACT_SILU)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = sorted(ts[2:])[len(ts[2:]) // 2]
        print("%-9s %-5s M=%6d N=%5d K=%5d  %8.1f us  %7.1f TFLOP/s (fp32-equivalent)" % (name, kind, M, N, K, t * 1e3, 2.0 * M * N * K / t / 1e9))
