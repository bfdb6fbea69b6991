import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from matinvent_b200.models.diffcsp import PhiloxNoise
from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData
dev = torch.device("cuda")
m = bench.build_model(dev)
B = int(sys.argv[1]); T = int(sys.argv[2])
na = bench.atom_counts(B)
batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])
res = {}
for mode, graph in (("ffma", True), ("tc", False), ("tc", True)):
    m.decoder.use_tc = mode == "tc"
    out, traj = m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=T, use_cuda_graph=graph, return_traj=True)
    res[(mode, graph)] = traj
    first_bad = None
    for t in sorted(traj.keys(), reverse=True):
        if not torch.isfinite(traj[t]["lattices"]).all() or not torch.isfinite(traj[t]["atom_types"]).all() or not torch.isfinite(traj[t]["frac_coords"]).all():
            first_bad = t; break
    print(mode, "graph" if graph else "eager", "max|l| %.3e" % float(out["lattices"].abs().max()), "first non-finite t:", first_bad)
ref = res[("ffma", True)]
for key in (("tc", False), ("tc", True)):
    for t in sorted(ref.keys(), reverse=True)[:12]:
        a, b = res[key][t]["lattices"], ref[t]["lattices"]
        print(key, t, "rel diff l %.2e" % float((a - b).abs().max() / b.abs().max()))
