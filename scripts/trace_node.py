"""Developer tool: per-phase SM-clock timeline of mi_node_chain (csrc/mi_node.cu built with -DMI_NODE_TRACE into
lib/libmi_ntrace.so).   python scripts/trace_node.py build   (CPU)   |   python scripts/trace_node.py run   (GPU)"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "matinvent_b200", "lib", "libmi_ntrace.so")
if sys.argv[1] == "build":
    from matinvent_b200.csrc import build as b
    srcs = [os.path.join(b.HERE, s) for s in b.SOURCES]
    subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ["-DMI_NODE_TRACE", "-shared", "-o", TRACE_LIB] + srcs)
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
m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=3)       # the last chain launch is a 2-phase one
dec = m.decoder
g = dec.graph_for(batch.num_atoms)
ws = dec.workspace(g, False)
# one more 3-phase launch so that the trace holds the full chain: rerun a forward and stop after layer 0's chain
import matinvent_b200.ops as ops
calls = []
orig = ops.node_chain
def once(*a, **k):
    if not calls:
        orig(*a, **k)
        torch.cuda.synchronize()
        calls.append(1)
        raise StopIteration
ops.node_chain = once
try:
    x = torch.rand(g.N, 3, device=dev); l = torch.randn(g.B, 3, 3, device=dev) + 4 * torch.eye(3, device=dev)
    dec.forward_graph(g, torch.randn(g.B, 256, device=dev), torch.randn(g.N, 100, device=dev), x, l)
except StopIteration:
    pass
NS = 32
buf = (ctypes.c_longlong * (160 * NS))()
lib = _lib.load()
lib.mi_node_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.mi_node_trace_read(buf, 160 * NS) == 0
names = ["start", "setup", "prologue", "B0", "p0_accfull", "p0_epi", "B1", "p1_accfull", "p1_store", "B2", "ln_out", "B3",
         "p2_t0", "p2_t1", "p2_t2", "p2_done", "end", "m0_w", "m0_a", "m0_commit", "m1_w", "m1_a", "m1_commit", "m2_w", "m2_a", "m2_commit"]
for cta in (0, 1, 2, 3, 40, 41, 80, 83):
    base = buf[cta * NS]
    order = [0, 1, 17, 2, 3, 18, 19, 4, 5, 6, 20, 21, 22, 7, 8, 9, 10, 11, 23, 24, 12, 13, 14, 25, 15, 16]
    print("CTA %2d: " % cta + " ".join("%s=%d" % (names[k], buf[cta * NS + k] - base) for k in order if buf[cta * NS + k]))
