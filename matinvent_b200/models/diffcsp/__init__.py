from .cspnet import CSPNet  # noqa: F401
from .diffusion import DiffCSPModule, PhiloxNoise, TapeNoise, TorchNoise  # noqa: F401
from .sample import ATOM_DIST, CrystalBatch, CrystalData, DiffCSPSampler, SampleDataset  # noqa: F401
