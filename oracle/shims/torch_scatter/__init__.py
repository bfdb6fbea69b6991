"""Shim: torch_scatter leaf semantics (sum / mean with count clamp >= 1) via index_add_."""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce='sum'):
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res.index_add_(0, index, src)
    if reduce in ('sum', 'add'):
        return res
    if reduce == 'mean':
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        cnt.clamp_(min=1)
        return res / cnt.view((-1,) + (1,) * (src.dim() - 1))
    raise NotImplementedError(reduce)


def segment_coo(src, index, out=None, dim_size=None, reduce='sum'):
    return scatter(src, index, 0, None, dim_size, reduce)


def segment_csr(src, indptr, out=None, reduce='sum'):
    n = indptr.numel() - 1
    idx = torch.repeat_interleave(torch.arange(n, device=src.device), indptr[1:] - indptr[:-1])
    return scatter(src, idx, 0, None, n, reduce)
