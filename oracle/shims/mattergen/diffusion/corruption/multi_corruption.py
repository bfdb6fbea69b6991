from typing import Generic, TypeVar

T = TypeVar("T")


class MultiCorruption(Generic[T]):
    pass


def apply(fns, broadcast=None, **kwargs):
    """mattergen.diffusion.corruption.multi_corruption.apply (published helper, restated): call fns[field] with, for every
    keyword, the entry of that keyword's per-field mapping, plus the broadcast arguments"""
    broadcast = broadcast or {}
    return {field: fn(**{k: v[field] for k, v in kwargs.items() if field in v}, **broadcast) for field, fn in fns.items()}
