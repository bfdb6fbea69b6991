"""Developer tool: build libmatinvent_b200 with -DMI_TC_TRACE into lib/libmi_trace.so (build step, CPU) or run a traced
tc GEMM and print the per-role SM-clock timeline of a few CTAs (GPU).   python scripts/trace_tc.py build|run"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "matinvent_b200", "lib", os.environ.get("TRACE_LIB", "libmi_trace.so"))
if sys.argv[1] == "build":
    from matinvent_b200.csrc import build as b
    srcs = [os.path.join(b.HERE, s) for s in b.SOURCES]
    subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ["-DMI_TC_TRACE"] + os.environ.get("DEFS", "").split() + ["-shared", "-o", TRACE_LIB] + srcs)
    print(TRACE_LIB)
    sys.exit(0)
os.environ["MATINVENT_B200_LIB"] = TRACE_LIB
import torch
from matinvent_b200 import ops, _lib
M, N, K = int(os.environ.get("M", "34445")), int(os.environ.get("N", "512")), int(os.environ.get("K", "512"))
merged = int(os.environ.get("MERGED", "1"))
act = int(os.environ.get("ACT", "1"))
A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / K ** 0.5
amax = A.abs().amax(dim=1).contiguous()
hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
s = ops.merged_scale(W) if merged else 1.0
ops.f16_split(W, hi, lo, s, 1.0 if merged else 2048.0)
C = torch.empty(M, N, device="cuda")
flush = torch.empty((256 << 20) if int(os.environ.get("FLUSH", "1")) else 16, dtype=torch.uint8, device="cuda")
gemm1 = int(os.environ.get("GEMM1", "0"))      # the first per-edge block: pre-split A, two gathers, row maxima
if gemm1:
    Ahi = torch.empty(M, K, device="cuda", dtype=torch.float16)
    Alo = torch.empty_like(Ahi)
    Asc = torch.rand(M, K, device="cuda") * 2 - 1
    ops.f16_split(Asc, Ahi, Alo, 2.0 ** 14 if merged else 1.0, 1.0 if merged else 2048.0)
    Nn = 2643
    PQ = torch.randn(Nn, 2 * N, device="cuda")
    src = torch.randint(0, Nn, (M,), device="cuda").sort().values.int()
    dst = torch.randint(0, Nn, (M,), device="cuda").int()
    am = torch.zeros(M, device="cuda")
evs = []
for _ in range(5):
    flush.zero_()
    evs.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
    evs[-1][0].record()
    if gemm1:
        ops.tc_gemm_presplit(Ahi, Alo, hi, lo, C, gathers=[(PQ[:, :N], src), (PQ[:, N:], dst)], act=act, amax_out=am,
                             alpha=(2.0 ** -14 if merged else 1.0) / s, flags=merged)
    else:
        ops.tc_gemm(A, hi, lo, C, act=act, a_amax=amax, alpha=1.0 / s, flags=merged)
    evs[-1][1].record()
torch.cuda.synchronize()
print("launch times (us):", " ".join("%.1f" % (a.elapsed_time(b) * 1e3) for a, b in evs))
TT, TS = 8, 24
buf = (ctypes.c_longlong * (160 * TT * TS))()
lib = _lib.load()
lib.mi_tc_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.mi_tc_trace_read(buf, 160 * TT * TS) == 0
names = ["tma0", "mma_accempty", "mma_full0", "mma_commit", "epi_accfull", "ld0", "ld1", "ld2", "ld3", "ch0", "ch1", "ch2",
         "ch3", "split0", "start", "end", "c0_sts", "c0_b0ld", "c0_b0st", "c0_b1ld", "c0_b1st", "", "", ""]
for cta in (0, 77):
    base = buf[(cta * TT + 0) * TS + 14]
    print("CTA %d  (cycles relative to producer start; end = %d)" % (cta, buf[(cta * TT) * TS + 15] - base))
    for t in range(TT):
        row = [buf[(cta * TT + t) * TS + k] for k in range(TS)]
        if row[0] == 0:
            continue
        print("  tile %d: " % t + " ".join("%s=%d" % (names[k], row[k] - base) for k in list(range(14)) + list(range(16, 21)) if row[k]))
