from .pl_module import MatterGenModule, SampleLoss  # noqa: F401
from .sample import MatterGenSampler, draw_samples_from_sampler  # noqa: F401
