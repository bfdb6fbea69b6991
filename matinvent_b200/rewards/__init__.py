from .composition import CompositionReward, synthetic_table  # noqa: F401
