"""Host-side random streams, bit-exact with the reference's compiled helpers.

Mirrors `modl.utils.randomkit` [ref: modl/utils/randomkit/__init__.py:1-2]:
`RandomState` [ref: random_fast.pyx:49-150] and `Sampler` [ref: sampler.pyx:10-69].
The arithmetic lives in C++ (modl_b200/csrc/host_rng.cpp) behind the C ABI; these classes
only marshal NumPy int64 buffers.
"""
import ctypes as C

import numpy as np

from ._lib import lib

__all__ = ["RandomState", "Sampler"]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RandomState(object):
    def __init__(self, seed=None):
        if seed is None:
            # the reference seeds from /dev/urandom here [ref: random_fast.pyx:66-68]
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        elif not isinstance(seed, (int, np.integer)):
            raise ValueError("Wrong seed")
        self.initial_seed = int(seed)
        self._h = lib().modl_rs_create(C.c_uint64(self.initial_seed & 0xFFFFFFFFFFFFFFFF))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().modl_rs_destroy(h)
            except Exception:
                pass

    def __reduce__(self):
        # like the reference, only the initial seed is pickled, not the stream position
        # [ref: random_fast.pyx:56-57, 149-150]
        return (RandomState, (self.initial_seed,))

    def seed(self, seed):
        lib().modl_rs_seed(self._h, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF))

    def randint(self, high):
        return int(lib().modl_rs_randint(self._h, C.c_uint64(int(high))))

    def binomial(self, n, p):
        return int(lib().modl_rs_binomial(self._h, int(n), float(p)))

    def permutation(self, size):
        out = np.empty(int(size), dtype=np.int64)
        lib().modl_rs_permutation(self._h, _ptr(out), int(size))
        return out

    def shuffle(self, x):
        """In-place shuffle along the first axis."""
        if isinstance(x, np.ndarray) and x.ndim == 1 and x.dtype == np.int64 and x.flags["C_CONTIGUOUS"]:
            lib().modl_rs_shuffle(self._h, _ptr(x), x.shape[0])
            return
        n = len(x)
        trace = np.empty(n, dtype=np.int64)
        lib().modl_rs_shuffle_with_trace(self._h, n, None, _ptr(trace))
        x[:] = x[trace] if isinstance(x, np.ndarray) else [x[i] for i in trace]

    def shuffle_with_trace(self, arrays):
        """Shuffle every array of the list with ONE swap sequence; returns the permutation
        such that new == old[perm] [ref: random_fast.pyx:127-144].  Arrays may be NumPy
        arrays or torch tensors (shuffled along dim 0)."""
        n = len(arrays[0])
        trace = np.empty(n, dtype=np.int64)
        lib().modl_rs_shuffle_with_trace(self._h, n, None, _ptr(trace))
        for x in arrays:
            if isinstance(x, np.ndarray):
                x[:] = x[trace]
            else:
                import torch
                idx = torch.as_tensor(trace, device=x.device)
                x.copy_(x.index_select(0, idx))
        return trace


class Sampler(object):
    """Feature-subset generator [ref: sampler.pyx:10-69]."""

    def __init__(self, range, rand_size, replacement, random_seed):
        self.range = int(range)
        self.rand_size = bool(rand_size)
        self.replacement = bool(replacement)
        self._h = lib().modl_sampler_create(self.range, int(self.rand_size), int(self.replacement),
                                            C.c_uint64(int(random_seed) & 0xFFFFFFFFFFFFFFFF))
        if not self._h:
            raise ValueError("invalid sampler range")
        self._buf = np.empty(max(self.range, 1), dtype=np.int64)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().modl_sampler_destroy(h)
            except Exception:
                pass

    def yield_subset(self, reduction):
        n = lib().modl_sampler_yield_subset(self._h, float(reduction), _ptr(self._buf))
        return self._buf[:n].copy()
