"""Developer tool (GPU, trace build): where do the ~24 us of a one-tile-per-CTA node-level tensor-core GEMM go?
In-kernel wall clock (globaltimer at entry / exit of every CTA) against the per-launch time of a CUDA graph of back-to-back
launches, alternating with a small ordinary kernel like the real step does.   python scripts/trace_tc.py build first."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MATINVENT_B200_LIB"] = os.path.join(ROOT, "matinvent_b200", "lib", "libmi_trace.so")
import torch
from matinvent_b200 import ops, _lib
lib = _lib.load()
lib.mi_tc_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
TT, TS = 8, 24
for (M, N, K) in ((2643, 512, 512), (2643, 512, 1024), (1334, 512, 512), (2643, 3, 512)):
    A, W = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / K ** 0.5
    amax = A.abs().amax(dim=1).contiguous()
    hi, lo = torch.empty_like(W, dtype=torch.float16), torch.empty_like(W, dtype=torch.float16)
    ops.f16_split(W, hi, lo)
    Cs = [torch.empty(M, N, device="cuda") for _ in range(8)]
    x = torch.randn(M, 512, device="cuda"); y = torch.empty_like(x); g = torch.ones(512, device="cuda"); b = torch.zeros(512, device="cuda")
    for mode in ("gemm only", "gemm + layernorm alternating"):
        def body():
            for C in Cs:
                ops.tc_gemm(A, hi, lo, C, act=1, a_amax=amax)
                if mode != "gemm only":
                    ops.layernorm_fwd(x, g, b, y, M, 512)
        body(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            body()
        for _ in range(3): gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): gr.replay()
        e1.record(); torch.cuda.synchronize()
        per = e0.elapsed_time(e1) * 1e3 / (20 * len(Cs))
        buf = (ctypes.c_longlong * (160 * TT * TS))()
        assert lib.mi_tc_trace_read(buf, 160 * TT * TS) == 0
        ctas = min(148, ((M + 127) // 128) * ((N + 127) // 128 if N > 64 else 1))
        ent = [buf[(c * TT) * TS + 21] for c in range(ctas)]
        ext = [buf[(c * TT) * TS + 22] for c in range(ctas)]
        cyc = [buf[(c * TT) * TS + 15] - buf[(c * TT) * TS + 14] for c in range(ctas)]
        print("M=%d N=%d K=%d  %-30s per pair of launches %.1f us | in-kernel span (min entry -> max exit) %.1f us, per-CTA "
              "entry->exit median %.1f us, entry spread %.1f us, producer-start->end median %d cycles"
              % (M, N, K, mode, per, (max(ext) - min(ent)) / 1e3, sorted(e - s for s, e in zip(ent, ext))[len(ent) // 2] / 1e3,
                 (max(ent) - min(ent)) / 1e3, sorted(cyc)[len(cyc) // 2]))
