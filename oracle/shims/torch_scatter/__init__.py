"""Pure-PyTorch stand-in for the two `torch_scatter` calls the reference makes on the path -- TEST INFRASTRUCTURE ONLY.
scatter_sum / scatter_mean(src, index, dim, dim_size=None) with a 1-D index along `dim` (droid_net.py:67, geom/ba.py:14-28);
semantics and the known answers of thirdparty/pytorch_scatter/test/test_scatter.py:12-24 (checked in tests/test_shims_cpu.py)."""
import torch

__all__ = ["scatter_sum", "scatter_mean", "scatter_add"]


def _index_along(src, index, dim):
    if index.dim() == src.dim():
        return index
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.dim()
    idx = _index_along(src, index, dim)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.zeros(shape, dtype=src.dtype, device=src.device) if out is None else out
    return res.scatter_add_(dim, idx, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.dim()
    s = scatter_sum(src, index, dim, None, dim_size)
    idx = _index_along(src, index, dim)
    cnt = torch.zeros_like(s).scatter_add_(dim, idx, torch.ones_like(src)).clamp(min=1)
    return s / cnt
