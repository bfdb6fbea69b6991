import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from matinvent_b200.models.diffcsp import PhiloxNoise
from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData
dev = torch.device("cuda")
m = bench.build_model(dev)
m.decoder.use_tc = (len(sys.argv) > 1 and sys.argv[1] == "tc")
na = bench.atom_counts(64)
batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])
out, traj = m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), return_traj=True)
for t in sorted(traj.keys(), reverse=True):
    if t % 100 == 0 or t < 3:
        d = traj[t]
        print(t, "max|l| %.3e  max|a| %.3e finite %s" % (float(d["lattices"].abs().max()), float(d["atom_types"].abs().max()), bool(torch.isfinite(d["lattices"]).all())))
ws = m.decoder.workspace(m.decoder.graph_for(batch.num_atoms), False)
print("max|a1| %.3e max|a2| %.3e max|cat| %.3e max|h| %.3e max|cb| %.3e" % tuple(float(x.abs().max()) for x in (ws.a1[0], ws.a2, ws.cat[0], ws.h[0], ws.cb)))
