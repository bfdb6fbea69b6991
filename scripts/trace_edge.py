"""Developer tool: per-tile SM-clock timeline of the CTA-pair per-edge kernels (csrc/mi_edge.cu built with -DMI_EDGE_TRACE
into lib/libmi_etrace.so).   python scripts/trace_edge.py build   (CPU)   |   MODE=0|1 python scripts/trace_edge.py run   (GPU)"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "matinvent_b200", "lib", "libmi_etrace.so")
if sys.argv[1] == "build":
    from matinvent_b200.csrc import build as b
    srcs = [os.path.join(b.HERE, s) for s in b.SOURCES]
    subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ["-DMI_EDGE_TRACE", "-shared", "-o", TRACE_LIB] + srcs)
    print(TRACE_LIB)
    sys.exit(0)
os.environ["MATINVENT_B200_LIB"] = TRACE_LIB
import torch
import bench
from matinvent_b200 import _lib
from matinvent_b200.models.diffcsp import PhiloxNoise
from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData
dev = torch.device("cuda", 0)
m = bench.build_model(dev)
na = bench.atom_counts(int(os.environ.get("B", "256")))
batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])
m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=2, use_cuda_graph=False)
dec = m.decoder
g = dec.graph_for(batch.num_atoms)
ws = dec.workspace(g, False)
presplit, merged = dec.edge_mode(g.E)
mode = int(os.environ.get("MODE", "0"))
H = dec.hidden_dim
torch.cuda.synchronize()
ws.cat[0][:, H:].zero_()
if mode == 0:
    dec.edge_gemm1(0, ws, g, g.E, ws.a1[0], False, presplit, merged)
else:
    dec.edge_gemm2(0, ws, g, g.E, ws.a1[0], ws.cat[0][:, H:], False, merged)
TT, TS = 6, 8
buf = (ctypes.c_longlong * (160 * TT * TS))()
lib = _lib.load()
lib.mi_edge_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.mi_edge_trace_read(buf, 160 * TT * TS) == 0
names = ["tma_first", "tma_last", "mma_at_accwait", "mma_acc_free", "mma_first_full", "mma_committed", "epi_accfull", "epi_done"]
for cta in (0, 1, 76, 77, 146):
    base = buf[(cta * TT) * TS + 0] or buf[(cta * TT) * TS + 6]
    print("CTA %d" % cta)
    for t in range(TT):
        row = [buf[(cta * TT + t) * TS + k] for k in range(TS)]
        if not any(row):
            continue
        print("  tile %d: " % t + " ".join("%s=%d" % (names[k], row[k] - base) for k in range(TS) if row[k]))
