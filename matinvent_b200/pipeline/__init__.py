from .base import ReinL  # noqa: F401
from .mat_invent import MatInvent  # noqa: F401
