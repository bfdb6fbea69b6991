from .base import ModelSuite  # noqa: F401
from .diffcsp import DiffCSPSuite  # noqa: F401
