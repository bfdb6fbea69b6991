import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from matinvent_b200 import ops
from matinvent_b200.models.diffcsp.graph import CrystalGraph
g = CrystalGraph(bench.atom_counts(256), "cuda")
M, N, K = g.E, 512, 768
A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
ops.f16_split(W, hi, lo)
C = torch.empty(M, N, device="cuda")
P = torch.randn(g.N, 1024, device="cuda"); Cb = torch.randn(g.B, 512, device="cuda")
am = torch.zeros(M, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(**kw):
    ts = []
    for it in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.tc_gemm(A, hi, lo, C, act=1, **kw); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts[2:])[2] * 1e3
gs = [(P[:, :512], g.edge_src), (P[:, 512:], g.edge_dst), (Cb, g.edge_graph)]
print("plain            %.1f us" % run())
print("amax             %.1f us" % run(amax_out=am))
print("1 gather (src)   %.1f us" % run(gathers=gs[:1]))
print("1 gather (dst)   %.1f us" % run(gathers=gs[1:2]))
print("3 gathers        %.1f us" % run(gathers=gs))
print("3 gathers + amax %.1f us" % run(gathers=gs, amax_out=am))
