import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops
M, N, K = 2643, 512, 100
A, W, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.randn(N, device="cuda")
C = torch.empty(M, N, device="cuda")
am = torch.zeros(M, device="cuda")
def t(fn):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        g.replay()
        e.record(st); torch.cuda.synchronize()
    return a.elapsed_time(e) / 20 * 1e3
print("sgemm K=100 bias+amax %.1f us" % t(lambda: ops.sgemm(A, W, C, bias=b, amax_out=am)))
print("sgemm K=100 bias      %.1f us" % t(lambda: ops.sgemm(A, W, C, bias=b)))
A2, W2 = torch.randn(M, 112, device="cuda"), torch.randn(N, 112, device="cuda")
print("sgemm K=112 bias      %.1f us" % t(lambda: ops.sgemm(A2, W2, C, bias=b)))
A3, W3 = torch.randn(M, 128, device="cuda"), torch.randn(N, 128, device="cuda")
print("sgemm K=128 bias      %.1f us" % t(lambda: ops.sgemm(A3, W3, C, bias=b)))
hi, lo = torch.empty_like(W3, dtype=torch.float16), torch.empty_like(W3, dtype=torch.float16)
ops.f16_split(W3, hi, lo)
amx = A3.abs().amax(dim=1).contiguous()
print("tc    K=128 bias      %.1f us" % t(lambda: ops.tc_gemm(A3, hi, lo, C, bias=b, a_amax=amx)))

for (M2, N2, K2) in ((2643, 512, 512), (2643, 512, 1024), (2643, 1024, 512)):
    A4, W4 = torch.randn(M2, K2, device="cuda"), torch.randn(N2, K2, device="cuda")
    h4, l4 = torch.empty_like(W4, dtype=torch.float16), torch.empty_like(W4, dtype=torch.float16)
    ops.f16_split(W4, h4, l4)
    am4 = A4.abs().amax(dim=1).contiguous()
    C4 = torch.empty(M2, N2, device="cuda")
    print("tc M=%d N=%d K=%d  %.1f us" % (M2, N2, K2, t(lambda: ops.tc_gemm(A4, h4, l4, C4, act=1, a_amax=am4))))
    os.environ["MI_TC_TN"] = "64"
    print("tc M=%d N=%d K=%d  TN=64 %.1f us" % (M2, N2, K2, t(lambda: ops.tc_gemm(A4, h4, l4, C4, act=1, a_amax=am4))))
    del os.environ["MI_TC_TN"]
