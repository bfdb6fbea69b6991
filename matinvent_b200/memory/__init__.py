from .replay_buffer import ReplayBuffer  # noqa: F401
