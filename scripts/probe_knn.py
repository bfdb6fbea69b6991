"""Throughput of the knn sampler (edge_style='knn', cutoff 7, max_neighbors 20: radius_graph_pbc rebuilt on the device every
forward, one D2H read of the edge count per forward, no CUDA graph) next to the fc sampler on the same batch: STEPS reverse steps
of B crystals (env B, STEPS; defaults 256, 100), CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from matinvent_b200.models.diffcsp import DiffCSPModule, PhiloxNoise  # noqa: E402
from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData  # noqa: E402

dev = torch.device("cuda", 0)
B, STEPS = int(os.environ.get("B", "256")), int(os.environ.get("STEPS", "100"))
na = bench.atom_counts(B)
HP = bench.HP
for style in ("fc", "knn"):
    m = DiffCSPModule(
        decoder=dict(hidden_dim=HP["hidden_dim"], num_layers=HP["num_layers"], max_atoms=HP["max_atoms"],
                     num_freqs=HP["num_freqs"], edge_style=style, cutoff=7.0, max_neighbors=20, ln=True, ip=True),
        beta_scheduler=dict(timesteps=HP["timesteps"], scheduler_mode="cosine"),
        sigma_scheduler=dict(timesteps=HP["timesteps"], sigma_begin=HP["sigma_begin"], sigma_end=HP["sigma_end"]),
        cost_lattice=HP["costs"][0], cost_coord=HP["costs"][1], cost_type=HP["costs"][2],
        time_dim=HP["time_dim"], latent_dim=HP["latent_dim"], device=dev, sigmas_norm=bench.sigmas_norm())
    m.decoder.reset_parameters(seed=0)
    for k in ("coord_w", "lattice_w", "type_w", "type_b"):
        m.decoder.w(k).mul_(bench.HEAD_SCALE)
    m.decoder.weights_changed()
    batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])
    m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=1), timesteps=5)      # warm-up: workspaces, tensor maps
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    out, _ = m.sample(batch, step_lr=bench.STEP_LR, noise=PhiloxNoise(dev, seed=2), timesteps=STEPS)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    g = m.decoder.graph_for(batch.num_atoms)
    print("%-3s  B=%d  edges=%d  %.3f ms per reverse step  -> %.1f crystals/s at 1000 steps  finite=%s" %
          (style, B, g.E, ms / STEPS, B / (ms / STEPS), bool(torch.isfinite(out["frac_coords"]).all())))
