"""Driver for ncu captures of the per-layer kernels in situ (no CUDA graph): two reverse steps at B crystals (env B, default
256).   ncu --set full --clock-control none -k regex:"node_chain|edge_pair" -s 12 -c 3 ... python scripts/prof_node.py"""
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
na = bench.atom_counts(int(os.environ.get("B", "256")))
batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])
m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=2, use_cuda_graph=False)
torch.cuda.synchronize()
