"""Elastic-net ball helpers -- mirror of `modl.utils.math.enet`
[ref: modl/utils/math/enet.pxd:10-16; enet.pyx:38-168], computed by sm_100a kernels.

Same call signatures as the reference (1-D vectors, in-place outputs).  Arguments may be
NumPy arrays (uploaded, results copied back in place) or CUDA tensors.
"""
import torch

from . import _lib
from ._util import as_device, ctx_of, ptr, stream_of

__all__ = ["enet_norm", "enet_projection", "enet_scale"]


def enet_norm(v, l1_ratio):
    t, _ = as_device(v)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    out = torch.empty(1, dtype=t.dtype, device=t.device)
    fn = getattr(_lib.lib(), "modl_enet_norm_" + _lib.sfx_of(t.dtype))
    _lib.check(fn(ctx_of(t).handle, ptr(t), 1, t.numel(), t.numel(), float(l1_ratio), ptr(out), stream_of(t.device)))
    return float(out.item())


def enet_projection(v, out, radius, l1_ratio):
    """out <- projection of v on {x : l1_ratio |x|_1 + (1 - l1_ratio) |x|_2^2 <= radius}."""
    t, _ = as_device(v)
    o, o_np = as_device(out, dtype=t.dtype)
    rad = torch.full((1,), float(radius), dtype=t.dtype, device=t.device)
    fn = getattr(_lib.lib(), "modl_enet_projection_" + _lib.sfx_of(t.dtype))
    _lib.check(fn(ctx_of(t).handle, ptr(t), ptr(o), 1, t.numel(), t.numel(), ptr(rad), float(l1_ratio),
                  stream_of(t.device)))
    if o_np:
        out[...] = o.cpu().numpy()
    elif o.data_ptr() != out.data_ptr():
        out.copy_(o)
    return out


def enet_scale(X, l1_ratio, radius=1.0):
    """Rescale the vector (or every row of a 2-D array) to elastic-net norm `radius`, in place."""
    t, was_np = as_device(X)
    rows, n = (1, t.numel()) if t.dim() == 1 else (t.shape[0], t.shape[1])
    fn = getattr(_lib.lib(), "modl_enet_scale_" + _lib.sfx_of(t.dtype))
    _lib.check(fn(ctx_of(t).handle, ptr(t), rows, n, n, float(l1_ratio), float(radius), stream_of(t.device)))
    if was_np:
        X[...] = t.cpu().numpy()
    elif t.data_ptr() != X.data_ptr():
        X.copy_(t)
    return X
