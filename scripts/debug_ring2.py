import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops
torch.manual_seed(0)
def run(M, N, K, merged, A=None, W=None):
    A = torch.randn(M, K, device="cuda") if A is None else A
    W = torch.randn(N, K, device="cuda") if W is None else W
    amax = A.abs().amax(dim=1).contiguous()
    hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
    C = torch.full((M, N), -1.0, device="cuda")
    if merged:
        s = ops.merged_scale(W)
        ops.f16_split(W, hi, lo, s, 1.0)
        ops.tc_gemm(A, hi, lo, C, a_amax=amax, alpha=1 / s, flags=1)
    else:
        ops.f16_split(W, hi, lo)
        ops.tc_gemm(A, hi, lo, C, a_amax=amax)
    ref = A.double() @ W.double().t()
    err = (C.double() - ref).abs()
    return float(err.max() / ref.abs().max()), err
for merged in (0, 1):
    for K in (64, 96, 128, 160, 192, 256, 384, 512, 768):
        e, err = run(128, 256, K, merged)
        bad_rows = (err.max(dim=1).values > 1e-3).sum().item()
        bad_cols = (err.max(dim=0).values > 1e-3).sum().item()
        print("merged=%d K=%d err %.2e bad rows %d cols %d" % (merged, K, e, bad_rows, bad_cols))
# which operand matters
K = 256
for merged in (0, 1):
    e1, _ = run(128, 256, K, merged, A=torch.ones(128, K, device="cuda"))
    e2, _ = run(128, 256, K, merged, W=torch.ones(256, K, device="cuda"))
    print("merged=%d K=%d  A=ones err %.2e   W=ones err %.2e" % (merged, K, e1, e2))
