from .ltm import LongTimeMem  # noqa: F401
from .replay_buffer import ReplayBuffer  # noqa: F401
