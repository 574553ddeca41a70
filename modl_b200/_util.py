"""Small marshalling helpers shared by the host-side mirrors."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("modl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_device(x, dtype=None, device=None):
    """-> (CUDA tensor, was_numpy).  NumPy input is uploaded; tensors must already be CUDA."""
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            x = x.to(device or default_device())
        if dtype is not None and x.dtype != dtype:
            x = x.to(dtype)
        return x.contiguous(), False
    a = np.ascontiguousarray(x)
    t = torch.from_numpy(a).to(device or default_device())
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t, True


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream_of(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ctx_of(t):
    return _lib.get_context(t.device.index if t.device.index is not None else torch.cuda.current_device())


def torch_dtype(np_dtype):
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.float32:
        return torch.float32
    if np_dtype == np.float64:
        return torch.float64
    raise TypeError("float32 or float64 expected, got %s" % np_dtype)
