"""Edge scatter (mi_segment_reduce) against the HBM roofline at the benchmark sizes: (a) one launch at a time, L2 flushed by a
256 MB memset before it; (b) launches back to back over rotating inputs that together exceed the L2 (6 x E x H floats), the
way the kernel runs inside a step: launch latency and ramp overlap the previous launch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from matinvent_b200 import ops
from matinvent_b200.models.diffcsp.graph import CrystalGraph
for B in (256, 1024):
    g = CrystalGraph(bench.atom_counts(B), "cuda")
    H = 512
    NB = 6 if B == 256 else 3
    Xs = [torch.randn(g.E, H, device="cuda") for _ in range(NB)]
    out = torch.empty(g.N, 2 * H, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    nbytes = 4 * g.E * H + 4 * (g.N + 1) + 4 * g.N * H
    for am in (None, torch.zeros(g.N, device="cuda")):
        ts = []
        for it in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.segment_reduce(Xs[0], g.seg_ptr, out[:, H:], g.N, H, mean=True, amax_out=am); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = sorted(ts[2:])[3] * 1e-3
        reps = 10
        for _ in range(2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            a.record()
            for r in range(reps):
                for X in Xs:
                    ops.segment_reduce(X, g.seg_ptr, out[:, H:], g.N, H, mean=True, amax_out=am)
            b.record(); torch.cuda.synchronize()
        tp = a.elapsed_time(b) * 1e-3 / (reps * NB)
        print("B=%d E=%d amax=%s: isolated %.1f us %.0f GB/s (%.1f%%) | back to back over %d x %d MB: %.1f us %.0f GB/s (%.1f%% of 6443)"
              % (B, g.E, am is not None, t * 1e6, nbytes / t / 1e9, 100 * nbytes / t / 1e9 / 6442.9, NB, 4 * g.E * H >> 20,
                 tp * 1e6, nbytes / tp / 1e9, 100 * nbytes / tp / 1e9 / 6442.9))
