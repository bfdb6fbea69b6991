import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(M, N, K, act):
    A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
    ops.f16_split(W, hi, lo)
    C = torch.empty(M, N, device="cuda")
    ts = []
    for it in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.tc_gemm(A, hi, lo, C, act=act); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts[2:])[2] * 1e3
for M in (34445, 18944):
    for K in (32, 128, 256, 512, 768):
        print("M=%d N=512 K=%4d  none %7.1f us   silu %7.1f us" % (M, K, run(M, 512, K, 0), run(M, 512, K, 1)))
