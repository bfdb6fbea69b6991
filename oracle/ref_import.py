"""TEST INFRASTRUCTURE (build container only): import the UNMODIFIED reference DiffCSP path.

`/root/reference/models/diffcsp/{cspnet,diffusion,scheduler,utils}.py` need torch_scatter,
torch_geometric, pytorch_lightning and hydra, none of which is installed; `oracle/shims/` restates
the few leaf functions they use.  Nothing here is shipped or imported by the
product path; `/root/reference` does not exist on the GPU box, so only `oracle/make_golden.py`
(run here, outputs committed under `tests/golden/`) and the `not gpu` pinning tests that skip when
the reference is absent may call this module.
"""
import os
import sys
import types

import torch

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# staged copy of the four reference files the sampler needs (git-ignored, shipped to the GPU box by gpurun, written by
# stage_reference() in the build container): lets bench.py time the UNMODIFIED reference where /root/reference is absent
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")
STAGED_FILES = ("cspnet.py", "diffusion.py", "scheduler.py", "utils.py")


def _pick_root():
    env = os.environ.get("MATINVENT_REFERENCE")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "models", "diffcsp", "cspnet.py")):
        return "/root/reference"
    return STAGED_ROOT


REF_ROOT = _pick_root()


def stage_reference():
    """copy models/diffcsp/{cspnet,diffusion,scheduler,utils}.py, unmodified, from the reference tree into
    baseline/_ref/ (build container only; returns the staged root, or None where the reference tree is absent)"""
    import filecmp
    import shutil
    src = os.path.join("/root/reference", "models", "diffcsp")
    if not os.path.isfile(os.path.join(src, "cspnet.py")):
        return None
    dst = os.path.join(STAGED_ROOT, "models", "diffcsp")
    os.makedirs(dst, exist_ok=True)
    for f in STAGED_FILES:
        if not os.path.isfile(os.path.join(dst, f)) or not filecmp.cmp(os.path.join(src, f), os.path.join(dst, f), shallow=False):
            shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    return STAGED_ROOT


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "diffcsp", "cspnet.py"))


def import_reference():
    """Returns (cspnet, diffusion, scheduler, utils) modules of the reference."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    for p in (REF_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, _SHIMS)
    import importlib
    utils = importlib.import_module("models.diffcsp.utils")
    scheduler = importlib.import_module("models.diffcsp.scheduler")
    cspnet = importlib.import_module("models.diffcsp.cspnet")
    diffusion = importlib.import_module("models.diffcsp.diffusion")
    return cspnet, diffusion, scheduler, utils


def build_reference_module(hp, sigmas_norm=None):
    """Hand-construct the reference `DiffCSPModule` (its own __init__ needs hydra.instantiate;
    models/diffcsp/diffusion.py:70-79) with the decoder/scheduler arguments in `hp` (a dict with the
    keys of oracle.diffcsp_oracle.default_hparams)."""
    cspnet, diffusion, scheduler, _ = import_reference()
    import pytorch_lightning as pl

    class RefModule(diffusion.DiffCSPModule):
        def __init__(self):
            pl.LightningModule.__init__(self)
            self.hparams.update(dict(cost_lattice=hp["cost_lattice"], cost_coord=hp["cost_coord"],
                                     cost_type=hp["cost_type"], time_dim=hp["time_dim"],
                                     latent_dim=hp["latent_dim"]))
            self.decoder = cspnet.CSPNet(
                hidden_dim=hp["hidden_dim"], latent_dim=hp["latent_dim"] + hp["time_dim"],
                num_layers=hp["num_layers"], max_atoms=hp["max_atoms"], act_fn="silu", dis_emb="sin",
                num_freqs=hp["num_freqs"], edge_style=hp["edge_style"], cutoff=hp["cutoff"],
                max_neighbors=hp["max_neighbors"], ln=hp["ln"], ip=hp["ip"], smooth=True,
                pred_type=True)
            self.beta_scheduler = scheduler.BetaScheduler(hp["timesteps"], hp["beta_mode"])
            self.sigma_scheduler = scheduler.SigmaScheduler(hp["timesteps"], hp["sigma_begin"],
                                                            hp["sigma_end"])
            self.time_dim = hp["time_dim"]
            self.time_embedding = diffusion.SinusoidalTimeEmbeddings(self.time_dim)
            self.keep_lattice = hp["cost_lattice"] < 1e-5
            self.keep_coords = hp["cost_coord"] < 1e-5

    m = RefModule()
    if sigmas_norm is not None:
        m.sigma_scheduler.sigmas_norm.copy_(torch.as_tensor(sigmas_norm))
    return m


def make_batch(num_atoms, **extra):
    """PyG-batch stand-in with the attributes the reference reads (num_graphs, num_nodes, num_atoms,
    batch, and optionally lengths/angles/frac_coords/atom_types/reward)."""
    num_atoms = torch.as_tensor(num_atoms, dtype=torch.long)
    b = types.SimpleNamespace(
        num_graphs=int(num_atoms.numel()), num_nodes=int(num_atoms.sum()), num_atoms=num_atoms,
        batch=torch.repeat_interleave(torch.arange(num_atoms.numel()), num_atoms))
    for k, v in extra.items():
        setattr(b, k, v)
    return b


def _import_memory_module(fname, modname):
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    for p in (REF_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, _SHIMS)
    import importlib.util
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, "memory", fname))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def import_reference_replay_buffer():
    """The reference's `memory.replay_buffer.ReplayBuffer`, unmodified (same two type-hint imports as ltm.py)."""
    return _import_memory_module("replay_buffer.py", "_ref_memory_replay_buffer").ReplayBuffer


def import_reference_ltm():
    """The reference's `memory.ltm.LongTimeMem`, unmodified (pandas is installed; its pymatgen / PyG imports are type
    hints, stubbed under oracle/shims)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    for p in (REF_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, _SHIMS)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_memory_ltm", os.path.join(REF_ROOT, "memory", "ltm.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.LongTimeMem


def import_reference_reward():
    """The reference's `rewards.reward` module, unmodified (numpy only; omegaconf / pymatgen names stubbed under shims)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    for p in (REF_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, _SHIMS)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_rewards_reward", os.path.join(REF_ROOT, "rewards", "reward.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def import_reference_mattergen_adapter():
    """The reference's in-tree MatterGen adapter, unmodified: (models.mattergen.pl_module, models.mattergen.loss) under the
    stub `mattergen` leaves of oracle/shims/mattergen."""
    if not os.path.isfile(os.path.join(REF_ROOT, "models", "mattergen", "pl_module.py")):
        raise RuntimeError("reference MatterGen adapter not found under %s" % REF_ROOT)
    for p in (REF_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, _SHIMS)
    import importlib
    loss = importlib.import_module("models.mattergen.loss")
    plm = importlib.import_module("models.mattergen.pl_module")
    return plm, loss


def mattergen_adapter_available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "mattergen", "pl_module.py"))
