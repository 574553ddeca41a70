"""Batched code solvers -- mirror of `modl.decomposition.dict_fact_fast`
[ref: modl/decomposition/dict_fact_fast.pyx:33-228], executed by the sm_100a kernels in
modl_b200/csrc (cd_kernels.cuh, ridge_kernels.cuh, basic_kernels.cuh).

Function names, argument order and in-place behaviour follow the reference so that its call
sites (dict_fact.py:77-90, 605-648) can bind to these directly.  Arrays may be NumPy (uploaded
for the call, `code` / `G_average` copied back in place) or CUDA tensors (zero-copy).
"""
import numpy as np
import torch

from . import _lib
from ._util import as_device, ctx_of, ptr, stream_of

__all__ = ["_enet_regression_single_gram", "_enet_regression_multi_gram", "_update_G_average",
           "_batch_weight"]


def _batch_weight(count, batch_size, learning_rate, offset):
    """[ref: dict_fact_fast.pyx:115-122]"""
    return float(_lib.lib().modl_batch_weight(int(count), int(batch_size), float(learning_rate), float(offset)))


def _regression(kind, G, Dx, X, code, indices, l1_ratio, alpha, positive, tol, max_iter, sweeps):
    g, _ = as_device(G)
    dx, dx_np = as_device(Dx, dtype=g.dtype)
    x, _ = as_device(X, dtype=g.dtype)
    c, c_np = as_device(code, dtype=g.dtype)
    idx = None
    if indices is not None:
        idx, _ = as_device(np.asarray(indices, dtype=np.int64) if not isinstance(indices, torch.Tensor) else indices,
                           dtype=torch.int64)
    b, k = dx.shape
    sw = None
    if sweeps is not None:
        sw = torch.zeros(b, dtype=torch.int32, device=g.device)
    fn = getattr(_lib.lib(), "modl_enet_regression_%s_gram_%s" % (kind, _lib.sfx_of(g.dtype)))
    _lib.check(fn(ctx_of(g).handle, ptr(g), ptr(dx), ptr(x), x.shape[1], x.shape[1], None, ptr(c), ptr(idx),
                  b, k, float(l1_ratio), float(alpha), int(bool(positive)), float(tol), int(max_iter), ptr(sw),
                  stream_of(g.device)))
    if float(l1_ratio) == 0:
        ctx_of(g).check_info(torch.cuda.current_stream(g.device).cuda_stream)
    if c_np:
        code[...] = c.cpu().numpy()
    elif c.data_ptr() != code.data_ptr():
        code.copy_(c)
    if dx_np and float(l1_ratio) == 0:
        Dx[...] = dx.cpu().numpy()      # the ridge branch overwrites Dx [ref: :188-197]
    if sweeps is not None:
        sweeps[...] = sw.cpu().numpy() if isinstance(sweeps, np.ndarray) else sw
    return code


def _enet_regression_single_gram(G, Dx, X, code, indices, l1_ratio, alpha, positive, tol, max_iter,
                                 sweeps=None):
    """For all ii: code[indices[ii]] <- argmin_w 1/2 w'Gw - Dx[ii]'w + alpha(l1 |w|_1 + (1-l1)/2 |w|^2),
    warm-started from code[indices[ii]]  [ref: dict_fact_fast.pyx:125-215]."""
    return _regression("single", G, Dx, X, code, indices, l1_ratio, alpha, positive, tol, max_iter, sweeps)


def _enet_regression_multi_gram(G, Dx, X, code, indices, l1_ratio, alpha, positive, tol, max_iter,
                                sweeps=None):
    """Same with one Gram matrix per sample, G[ii]  [ref: dict_fact_fast.pyx:33-113]."""
    return _regression("multi", G, Dx, X, code, indices, l1_ratio, alpha, positive, tol, max_iter, sweeps)


def _update_G_average(G_average, G, w_sample):
    """G_average[ii] = (1 - w[ii]) G_average[ii] + w[ii] G  [ref: dict_fact_fast.pyx:217-228]."""
    ga, ga_np = as_device(G_average)
    g, _ = as_device(G, dtype=ga.dtype)
    w, _ = as_device(w_sample, dtype=ga.dtype)
    fn = getattr(_lib.lib(), "modl_update_G_average_" + _lib.sfx_of(ga.dtype))
    _lib.check(fn(ctx_of(ga).handle, ptr(ga), ptr(g), ptr(w), None, w.shape[0], g.shape[0], stream_of(ga.device)))
    if ga_np:
        G_average[...] = ga.cpu().numpy()
    elif ga.data_ptr() != G_average.data_ptr():
        G_average.copy_(ga)
    return G_average
