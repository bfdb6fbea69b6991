import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


@pytest.fixture(scope="session")
def gold_small():
    return load_gold("small_net.pt")


@pytest.fixture(scope="session")
def gold_full():
    return load_gold("full_net.pt")


@pytest.fixture(scope="session")
def gold_rg():
    return load_gold("radius_graph.pt")


@pytest.fixture(scope="session")
def gold_baseline():
    """one forward of the unmodified reference on the benchmark batch (oracle/make_golden.py --baseline)"""
    return load_gold("baseline_forward.pt")


def baseline_inputs(gb):
    """the inputs oracle/make_golden.py forward_case drew for that fixture, regenerated from its seed"""
    from oracle import diffcsp_oracle as O
    na = gb["num_atoms"]
    B, N = len(na), int(na.sum())
    g = torch.Generator().manual_seed(gb["seed"])
    t = O.time_embedding(torch.full((B,), gb["t_int"]), gb["hp"]["time_dim"])
    a = torch.randn(N, 100, generator=g)
    x = torch.rand(N, 3, generator=g)
    l = torch.randn(B, 3, 3, generator=g)
    return na, t, a, x, l, torch.repeat_interleave(torch.arange(B), na)


def build_module(hp, sd, sigmas_norm, device="cuda"):
    """matinvent_b200 DiffCSPModule from an oracle-style hparam dict + reference-named decoder weights."""
    from matinvent_b200.models.diffcsp import DiffCSPModule
    m = DiffCSPModule(
        decoder=dict(hidden_dim=hp["hidden_dim"], num_layers=hp["num_layers"], max_atoms=hp["max_atoms"],
                     act_fn="silu", dis_emb="sin", num_freqs=hp["num_freqs"], edge_style=hp["edge_style"],
                     cutoff=hp["cutoff"], max_neighbors=hp["max_neighbors"], ln=hp["ln"], ip=hp["ip"]),
        beta_scheduler=dict(timesteps=hp["timesteps"], scheduler_mode=hp["beta_mode"]),
        sigma_scheduler=dict(timesteps=hp["timesteps"], sigma_begin=hp["sigma_begin"], sigma_end=hp["sigma_end"]),
        cost_lattice=hp["cost_lattice"], cost_coord=hp["cost_coord"], cost_type=hp["cost_type"],
        time_dim=hp["time_dim"], latent_dim=hp["latent_dim"], device=device, sigmas_norm=sigmas_norm)
    m.decoder.load_state_dict(sd)
    return m


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def wrapped_err(a, b):
    d = (a.detach().double().cpu() - b.detach().double().cpu()).abs()
    return float(torch.minimum(d, 1 - d).abs().max())
