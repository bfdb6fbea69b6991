"""Fine-tune timestep latency at the reference's working point (<= 18 crystals; 3 epochs x 1000 timesteps per RL
iteration) and the kernel breakdown of one captured timestep."""
import os, sys, time, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from matinvent_b200.models.diffcsp import PhiloxNoise
from matinvent_b200.models.diffcsp.finetune import FineTuner
from test_gpu_pipeline import _ft_batch
dev = torch.device("cuda", 0)
agent, prior = bench.build_model(dev), bench.build_model(dev)
for p in prior.parameters():
    p.requires_grad = False
B = int(os.environ.get("B", "18"))
na = bench.atom_counts(B)
na = [max(1, n) for n in na]
data, batch = _ft_batch(na, 4)
tuner = FineTuner(agent, prior, lr=1e-4, accum_steps=50, sigma=0.025, noise=PhiloxNoise(dev, seed=1),
                  group=int(os.environ["G"]) if "G" in os.environ else None)
tuner.run_batch(batch, 100)
torch.cuda.synchronize(); t0 = time.time()
tuner.run_batch(batch, 500)
torch.cuda.synchronize(); t1 = time.time()
print("B=%d crystals, %d atoms, %d edges: %.3f ms per fine-tune timestep (graph replay, Adam every 50)" %
      (B, sum(na), sum(n * n for n in na), (t1 - t0) / 500 * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tuner.run_batch(batch, 50)
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        tot[ev.name[:80]][0] += ev.device_time; tot[ev.name[:80]][1] += 1
allt = sum(v[0] for v in tot.values())
print("GPU busy %.1f us per timestep, %d launches per timestep" % (allt / 50, sum(v[1] for v in tot.values()) / 50))
for k, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:14]:
    print("%8.1f us/step %6.1f us/launch x%5.1f  %s" % (t / 50, t / n, n / 50, k))
