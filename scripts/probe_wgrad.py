import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops
def t(fn, n=20):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): fn()
        g.replay(); torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); g.replay(); e.record(st); torch.cuda.synchronize()
    return a.elapsed_time(e) / n * 1e3
for (M, N, K) in ((512, 768, 2786), (512, 512, 2786), (512, 1024, 204), (512, 512, 204), (1024, 512, 204), (512, 9, 18), (100, 512, 204)):
    dY, X = torch.randn(K, M, device="cuda"), torch.randn(K, N, device="cuda")
    G = torch.zeros(M, N, device="cuda")
    for sk in (1, 2, 4, 11, 22):
        us = t(lambda: ops.sgemm(dY, X, G, transA=True, transB=False, M=M, N=N, K=K, beta=1.0, splitk=sk))
        print("wgrad M=%d N=%d K=%d splitk=%d: %.1f us" % (M, N, K, sk, us), flush=True)
for (M, N, K) in ((2786, 512, 512), (204, 512, 512), (204, 1024, 512), (204, 512, 1024)):
    dY, W = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
    C = torch.zeros(M, N, device="cuda")
    print("dgrad NN M=%d N=%d K=%d: %.1f us" % (M, N, K, t(lambda: ops.sgemm(dY, W, C, transB=False, M=M, N=N, K=K))))
