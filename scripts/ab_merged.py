"""A/B: separate-accumulator 128x128 tiles vs merged single-accumulator 128x256 tiles (time + error vs float64)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matinvent_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.manual_seed(0)


def timeit(fn):
    ts = []
    for it in range(9):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts[2:])[3] * 1e3


def run(M, N, K, act, positive=False):
    A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / K ** 0.5
    if positive:
        A, W = A.abs(), W.abs()
    amax = A.abs().amax(dim=1).contiguous()
    ref = (A.double() @ W.double().t())
    if act:
        ref = torch.nn.functional.silu(ref)
    hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
    out = {}
    ops.f16_split(W, hi, lo)
    C = torch.empty(M, N, device="cuda")
    t = timeit(lambda: ops.tc_gemm(A, hi, lo, C, act=act, a_amax=amax))
    out["sep"] = (t, float((C.double() - ref).abs().max() / ref.abs().max()))
    s = ops.merged_scale(W)
    ops.f16_split(W, hi, lo, s, 1.0)
    C2 = torch.empty(M, N, device="cuda")
    t = timeit(lambda: ops.tc_gemm(A, hi, lo, C2, act=act, a_amax=amax, alpha=1.0 / s, flags=ops.TC_MERGED))
    out["mrg"] = (t, float((C2.double() - ref).abs().max() / ref.abs().max()))
    return out


for M in (34445, 18944):
    for K in (128, 512, 768):
        for pos in (False, True):
            r = run(M, 512, K, 1, pos)
            print("M=%d N=512 K=%4d pos=%d  sep %7.1f us err %.2e | merged %7.1f us err %.2e" %
                  (M, K, pos, r["sep"][0], r["sep"][1], r["mrg"][0], r["mrg"][1]), flush=True)
