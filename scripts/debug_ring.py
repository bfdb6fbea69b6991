import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops
torch.manual_seed(0)
def probe(M, N, K, merged, use_amax=True):
    A = torch.randn(M, K, device="cuda")
    amax = A.abs().amax(dim=1).contiguous() if use_amax else None
    blocks = A.view(M, K // 32, 32).sum(-1)          # [M, nkb]
    out = []
    for kb0 in range(K // 32):
        W = torch.zeros(N, K, device="cuda")
        W[:, kb0 * 32:(kb0 + 1) * 32] = 1.0
        hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
        C = torch.full((M, N), -1.0, device="cuda")
        if merged:
            ops.f16_split(W, hi, lo, 2.0 ** 14, 1.0)
            ops.tc_gemm(A, hi, lo, C, a_amax=amax, alpha=2.0 ** -14, flags=1)
        else:
            ops.f16_split(W, hi, lo)
            ops.tc_gemm(A, hi, lo, C, a_amax=amax)
        torch.cuda.synchronize()
        # which A block does column 0 of C correspond to?
        d = (C[:, :1] - blocks).abs().max(dim=0).values          # [nkb]
        j = int(d.argmin())
        out.append("%d->%d(%.1e)" % (kb0, j, float(d[j])))
    print("M=%d N=%d K=%d merged=%d amax=%d:" % (M, N, K, merged, use_amax), " ".join(out))
for merged in (0, 1):
    for K in (128, 256):
        probe(128, 256, K, merged)
probe(128, 256, 256, 0, False)
