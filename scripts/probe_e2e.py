"""Where does DiffCSPSampler.generate spend its time beyond the 1000 graph replays?"""
import os, sys, time, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from matinvent_b200.models.diffcsp import DiffCSPSampler, PhiloxNoise
from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData
dev = torch.device("cuda", 0)
m = bench.build_model(dev)
sampler = DiffCSPSampler(batch_size=256, num_batches=1)
np.random.seed(0); torch.manual_seed(0)
sampler.generate(m)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    data, _ = sampler.generate(m)
    torch.cuda.synchronize(); t1 = time.time()
    print("generate: %.1f ms for %d crystals, %d atoms" % ((t1 - t0) * 1e3, len(data), sum(int(d.num_atoms) for d in data)))
# same shape twice through sample(): first call builds graph + workspace + CUDA graph
na = bench.atom_counts(256)
for T in (1000, 1000):
    batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])
    torch.cuda.synchronize(); t0 = time.time()
    m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=T)
    torch.cuda.synchronize(); t1 = time.time()
    print("sample(T=%d) fixed shape: %.1f ms" % (T, (t1 - t0) * 1e3))
na2 = list(reversed(na))
batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na2])
torch.cuda.synchronize(); t0 = time.time()
m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=2)
torch.cuda.synchronize(); t1 = time.time()
print("sample(T=2) new shape (graph + workspace + capture): %.1f ms" % ((t1 - t0) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
data, _ = sampler.generate(m)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
