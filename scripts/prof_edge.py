"""Driver for ncu captures of the two per-edge GEMMs exactly as CSPNet issues them at the benchmark batch
(256 crystals, 34 445 edges): one score-network evaluation to fill the workspace, then the two launches of layer 0
a few times.   ncu --set full -k regex:edge_pair_kernel ... python scripts/prof_edge.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from matinvent_b200.models.diffcsp import PhiloxNoise  # noqa: E402
from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData  # noqa: E402

dev = torch.device("cuda", 0)
m = bench.build_model(dev)
na = bench.atom_counts(256)
batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])
m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=2, use_cuda_graph=False)
dec = m.decoder
g = dec.graph_for(batch.num_atoms)
ws = dec.workspace(g, False)
presplit, merged = dec.edge_mode(g.E)
torch.cuda.synchronize()
print("E =", g.E, "presplit", presplit, "merged", merged)
H = dec.hidden_dim
for rep in range(3):
    ws.cat[0][:, H:].zero_()
    dec.edge_gemm1(0, ws, g, g.E, ws.a1[0], False, presplit, merged)
    dec.edge_gemm2(0, ws, g, g.E, ws.a1[0], ws.cat[0][:, H:], False, merged)
torch.cuda.synchronize()
