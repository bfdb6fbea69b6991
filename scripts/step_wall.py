"""Wall-clock time per reverse step (CUDA-graph replays, Philox noise) at a few batch sizes: what bench.py's value is made of."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from matinvent_b200.models.diffcsp import PhiloxNoise
from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData
dev = torch.device("cuda", 0)
m = bench.build_model(dev)
for B in [int(b) for b in os.environ.get("BS", "256,128").split(",")]:
    na = bench.atom_counts(B)
    batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])
    m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=30)
    T = int(os.environ.get("T", "600"))
    best = 1e9
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=2 + rep), timesteps=T)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    per = best / T                        # includes the eager first step and one graph capture (~15 ms per run)
    print("B=%d: %.1f us per reverse step (%d steps) -> %.1f crystals/s" % (B, per * 1e6, T, B / (per * 1000)))
