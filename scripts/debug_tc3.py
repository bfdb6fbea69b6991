import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops
torch.manual_seed(0)
for M, N, K in ((300, 512, 512), (2643, 512, 512)):
    for sc in (1e-3, 1.0, 100.0, 3e4, 1e6):
        A = torch.randn(M, K, device="cuda") * sc
        W = torch.randn(N, K, device="cuda") * 0.05
        hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
        ops.f16_split(W, hi, lo)
        C = torch.empty(M, N, device="cuda")
        amax = A.abs().amax(dim=1).contiguous()
        ops.tc_gemm(A, hi, lo, C, a_amax=amax)
        ref = A.double() @ W.double().t()
        bad = int((~torch.isfinite(C)).sum())
        err = float((C.double() - ref).abs().max() / ref.abs().max()) if bad == 0 else float("nan")
        print("M=%d scale %.0e: nonfinite %d rel err %.2e   C[0,:4]=%s ref=%s" % (M, sc, bad, err, C[0, :4].tolist(), ref[0, :4].tolist()))
