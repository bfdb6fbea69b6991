"""Per-kernel GPU time of reverse-diffusion steps replayed from the CUDA graph (torch profiler / CUPTI), B=256."""
import os, sys, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from matinvent_b200.models.diffcsp import PhiloxNoise
from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData
dev = torch.device("cuda", 0)
m = bench.build_model(dev)
B = int(os.environ.get("B", "256"))
na = bench.atom_counts(B)
batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])
T = 40
m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=T)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=2), timesteps=T)
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = ev.name[:90]
        tot[k][0] += ev.device_time
        tot[k][1] += 1
allt = sum(v[0] for v in tot.values())
print("total GPU time %.1f us per reverse step (%d steps)" % (allt / T, T))
for k, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%8.1f us/step  %6.1f us/launch  x%5.1f/step  %s" % (t / T, t / n, n / T, k))
