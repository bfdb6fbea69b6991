import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from matinvent_b200 import ops
from matinvent_b200.models.diffcsp.graph import CrystalGraph
for B in (256, 1024):
    g = CrystalGraph(bench.atom_counts(B), "cuda")
    H = 512
    X = torch.randn(g.E, H, device="cuda"); out = torch.empty(g.N, 2 * H, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for am, rows in ((None, 0), (torch.zeros(g.N, device="cuda"), 0), (None, g.E), (torch.zeros(g.N, device="cuda"), g.E)):
        ts = []
        for it in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.segment_reduce(X, g.seg_ptr, out[:, H:], g.N, H, mean=True, amax_out=am, rows=rows); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = sorted(ts[2:])[3] * 1e-3
        nbytes = 4 * g.E * H + 4 * (g.N + 1) + 4 * g.N * H
        print("B=%d E=%d amax=%s streaming=%s: %.1f us  %.0f GB/s  (%.1f%% of 6443)" % (B, g.E, am is not None, rows > 0, t * 1e6, nbytes / t / 1e9, 100 * nbytes / t / 1e9 / 6442.9))
