"""What does writing a 70 MB fp32 output cost right after an L2 flush (dirty L2), and with a clean L2?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
M, N = 34445, 512
C = torch.empty(M, N, device="cuda")
src = torch.randn(M, N, device="cuda")


def timeit(fn, pre):
    ts = []
    for it in range(9):
        pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts[2:])[3] * 1e3

dirty = lambda: flush.zero_()
clean = lambda: flush.sum()
for name, pre in (("dirty-L2", dirty), ("clean-L2", clean)):
    print(name, "fill 70MB      %.1f us" % timeit(lambda: C.fill_(1.0), pre))
    print(name, "copy 70MB->70MB %.1f us" % timeit(lambda: C.copy_(src), pre))
    for K in (32, 128, 512):
        A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / K ** 0.5
        amax = A.abs().amax(dim=1).contiguous()
        hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
        ops.f16_split(W, hi, lo)
        print(name, "tc sep K=%d act=0  %.1f us" % (K, timeit(lambda: ops.tc_gemm(A, hi, lo, C, a_amax=amax), pre)))
        s = ops.merged_scale(W)
        ops.f16_split(W, hi, lo, s, 1.0)
        print(name, "tc mrg K=%d act=0  %.1f us" % (K, timeit(lambda: ops.tc_gemm(A, hi, lo, C, a_amax=amax, alpha=1 / s, flags=1), pre)))
        print(name, "tc mrg K=%d act=1  %.1f us" % (K, timeit(lambda: ops.tc_gemm(A, hi, lo, C, a_amax=amax, alpha=1 / s, flags=1, act=1), pre)))
# small M: fixed latency
for M2 in (128, 148 * 128, 2 * 148 * 128):
    A, W = torch.randn(M2, 512, device="cuda"), torch.randn(N, 512, device="cuda") / 512 ** 0.5
    amax = A.abs().amax(dim=1).contiguous()
    hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
    s = ops.merged_scale(W)
    ops.f16_split(W, hi, lo, s, 1.0)
    C2 = torch.empty(M2, N, device="cuda")
    print("M=%d K=512 merged %.1f us" % (M2, timeit(lambda: ops.tc_gemm(A, hi, lo, C2, a_amax=amax, alpha=1 / s, flags=1), dirty)))
    ops.f16_split(W, hi, lo)
    print("M=%d K=512 sep    %.1f us" % (M2, timeit(lambda: ops.tc_gemm(A, hi, lo, C2, a_amax=amax), dirty)))
