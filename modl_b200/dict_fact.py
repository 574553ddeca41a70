"""DictFact / Coder -- the reference's scikit-learn-style estimators with the per-minibatch
inner loop executed on a B200.

Host-side mirror of `modl.decomposition.dict_fact` [ref: modl/decomposition/dict_fact.py]:
same constructor keywords (:128-153), same public methods (`fit`, `partial_fit`, `transform`,
`score`, `prepare`, `shuffle`, `set_params`) and the same fitted attributes (:223-250).  The
state lives in device memory (torch CUDA tensors); the NumPy attributes of the reference
(`components_`, `code_`, `C_`, `B_`, `G_`, ...) are exposed as properties that copy to / from the
device, and the tensors themselves as `<name>dev` (e.g. `components_dev`).

What runs where
  * host (Python + C++ behind the C ABI): the integer bookkeeping of a step -- feature subset
    (bit-exact MT19937 sampler), `n_iter_`, `sample_n_iter_`, batch weight, atom order -- exactly
    the statements of `_single_batch_fit` [ref: :507-515, :672];
  * device (hand-written sm_100a kernels, one C call `modl_batch_fit_*` per minibatch): gathers,
    Gram products, code solve, statistics, dictionary update [ref: :517-533].
There is no CPU fallback: without the built extension or without a CUDA device this raises.
"""
import ctypes as C
import time
from math import ceil

import numpy as np
import torch
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils import check_array, check_random_state, gen_batches
from sklearn.utils.validation import check_is_fitted

from . import _lib
from ._util import default_device, ptr, stream_of, torch_dtype
from .dict_fact_fast import _batch_weight
from .randomkit import RandomState, Sampler

MAX_INT = np.iinfo(np.int64).max

__all__ = ["DictFact", "Coder", "CodingMixin", "get_sub_slice"]


def get_sub_slice(indices, sub_indices):
    """Safe indexer with nested slices [ref: modl/utils/__init__.py:4-27]."""
    if indices is None:
        if isinstance(sub_indices, slice):
            return np.arange(sub_indices.start, sub_indices.stop)
        return sub_indices
    if isinstance(indices, slice):
        return np.arange(indices.start + sub_indices.start, indices.start + sub_indices.stop)
    return indices[sub_indices]


class _DeviceArray(object):
    """NumPy-facing view of a device tensor stored as `_d_<name>` on the instance."""

    def __init__(self, name):
        self.slot = "_d_" + name
        self.name = name

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        t = obj.__dict__.get(self.slot)
        if t is None:
            raise AttributeError(self.name)
        settle = getattr(obj, "_settle", None)
        if settle is not None:
            settle()                # work enqueued on side streams lands before the state is read
        return t.detach().cpu().numpy()

    def __set__(self, obj, value):
        if value is None:
            obj.__dict__[self.slot] = None
            return
        dev = obj.__dict__.get("_device") or default_device()
        cur = obj.__dict__.get(self.slot)
        if isinstance(value, torch.Tensor):
            t = value.to(dev)
        else:
            t = torch.from_numpy(np.ascontiguousarray(value)).to(dev)
        obj.__dict__["_state_dirty"] = True     # the next partial_fit orders its own streams behind this write
        if cur is not None and cur.shape == t.shape:
            cur.copy_(t)            # keep the storage (and dtype) the kernels already point at
        else:
            obj.__dict__[self.slot] = t.contiguous()

    def __delete__(self, obj):
        obj.__dict__.pop(self.slot, None)


def _to_host_rows(X, dtype):
    X = check_array(X, order='C', dtype=[np.float32, np.float64])
    if dtype is not None and X.dtype != dtype:
        X = X.astype(dtype)
    return X


class CodingMixin(TransformerMixin):
    """[ref: dict_fact.py:23-124]"""

    components_ = _DeviceArray("components_")

    def _set_coding_params(self, n_components, code_alpha=1, code_l1_ratio=1, tol=1e-2, max_iter=100,
                           code_pos=False, random_state=None, n_threads=1):
        self.n_components = n_components
        self.code_l1_ratio = code_l1_ratio
        self.code_alpha = code_alpha
        self.code_pos = code_pos
        self.random_state = random_state
        self.tol = tol
        self.max_iter = max_iter
        # kept for API compatibility; the device kernels parallelise over samples themselves
        self.n_threads = n_threads

    @property
    def components_dev(self):
        return self.__dict__.get("_d_components_")

    def _ctx(self):
        dev = self._d_components_.device
        return _lib.get_context(dev.index if dev.index is not None else torch.cuda.current_device())

    def _gram_dx(self, D, X, G_out, Dx_out, xnorm2):
        k, p = D.shape
        b = X.shape[0] if X is not None else 0
        fn = getattr(_lib.lib(), "modl_gram_dx_" + _lib.sfx_of(D.dtype))
        _lib.check(fn(self._ctx().handle, ptr(D), D.stride(0), ptr(X), X.stride(0) if X is not None else p,
                      None, 0, k, b, p, 1.0, ptr(G_out), ptr(Dx_out), ptr(xnorm2), stream_of(D.device)))

    def transform(self, X, chunk_rows=16384):
        """Codes of the rows of X on the current dictionary [ref: dict_fact.py:47-92].
        NumPy in -> NumPy out; CUDA tensor in -> CUDA tensor out."""
        check_is_fitted(self, 'components_')
        D = self._d_components_
        tensor_in = isinstance(X, torch.Tensor)
        if not tensor_in:
            X = _to_host_rows(X, np.dtype(str(D.dtype).replace('torch.', '')))
        n = X.shape[0]
        k = self.n_components
        dev = D.device
        if getattr(self, 'G_agg', None) == 'full' and self.__dict__.get('_d_G_') is not None:
            G = self._d_G_
        else:
            G = torch.empty((k, k), dtype=D.dtype, device=dev)
            self._gram_dx(D, None, G, None, None)
        code = torch.ones((n, k), dtype=D.dtype, device=dev)
        reg = getattr(_lib.lib(), "modl_enet_regression_single_gram_" + _lib.sfx_of(D.dtype))
        for sl in gen_batches(n, int(chunk_rows)):
            Xc = X[sl]
            Xc = (Xc.to(dev, D.dtype) if tensor_in else torch.from_numpy(np.ascontiguousarray(Xc)).to(dev)).contiguous()
            bc = Xc.shape[0]
            Dx = torch.empty((bc, k), dtype=D.dtype, device=dev)
            xn = torch.empty((bc,), dtype=D.dtype, device=dev)
            self._gram_dx(D, Xc, None, Dx, xn)
            cview = code[sl]
            _lib.check(reg(self._ctx().handle, ptr(G), ptr(Dx), None, 0, 0, ptr(xn), ptr(cview), None, bc, k,
                           float(self.code_l1_ratio), float(self.code_alpha), int(bool(self.code_pos)),
                           float(self.tol), int(self.max_iter), None, stream_of(dev)))
        return code if tensor_in else code.cpu().numpy()

    def score(self, X):
        """Objective value on test data [ref: dict_fact.py:94-114]."""
        check_is_fitted(self, 'components_')
        D = self._d_components_
        Xd = X if isinstance(X, torch.Tensor) else torch.from_numpy(
            _to_host_rows(X, np.dtype(str(D.dtype).replace('torch.', ''))))
        Xd = Xd.to(D.device, D.dtype)
        code = self.transform(Xd)
        loss = torch.sum((Xd - code @ D) ** 2) / 2
        norm1 = torch.sum(torch.abs(code))
        norm2 = torch.sum(code ** 2)
        regul = self.code_alpha * (norm1 * self.code_l1_ratio + (1 - self.code_l1_ratio) * norm2 / 2)
        return float((loss + regul).item()) / Xd.shape[0]

    # pickling: device tensors travel as NumPy arrays, host RNG objects are dropped
    def __getstate__(self):
        state = {}
        for key, val in self.__dict__.items():
            if key in ("_pipeline", "_time_events", "_keepalive", "_ovl", "_prm", "_step_fn", "_fit_loop", "_fit_keep"):
                continue
            if isinstance(val, torch.Tensor):
                state[key] = ("__tensor__", val.detach().cpu().numpy())
            elif isinstance(val, Sampler):
                continue       # like the reference, the sampler stream is not picklable
            else:
                state[key] = val
        return state

    def __setstate__(self, state):
        for key, val in state.items():
            if isinstance(val, tuple) and len(val) == 2 and isinstance(val[0], str) and val[0] == "__tensor__":
                dev = default_device()
                self.__dict__[key] = torch.from_numpy(val[1]).to(dev)
            else:
                self.__dict__[key] = val


class DictFact(CodingMixin, BaseEstimator):
    """Stochastic-subsampled online matrix factorisation [ref: dict_fact.py:127-250].

    Parameters are those of the reference (see its docstring, :154-222).  `n_threads` is
    accepted and ignored.  Two extra keywords: `device` (torch device or index, default: the
    current CUDA device) and `async_host_copy` (default False).  `partial_fit` takes one optional extra
    argument, `code_out` (a pinned host tensor / array of n_rows x n_components that receives the codes of the
    rows just seen, copied out on a stream of its own).  With `async_host_copy=True`,
    `partial_fit` on PINNED host rows returns as soon as the copies and kernels are enqueued --
    the usual contract of an asynchronous pinned-memory copy: the caller must leave those rows
    untouched until `synchronize()` (or any later read of a fitted attribute, which
    synchronises) -- so that the copy of the next batch overlaps the kernels of this one even
    when every call carries a single batch.
    """

    code_ = _DeviceArray("code_")
    C_ = _DeviceArray("C_")
    B_ = _DeviceArray("B_")
    G_ = _DeviceArray("G_")
    comp_norm_ = _DeviceArray("comp_norm_")
    Dx_average_ = _DeviceArray("Dx_average_")
    G_average_ = _DeviceArray("G_average_")

    def __init__(self, reduction=1, learning_rate=1, sample_learning_rate=0.76, Dx_agg='masked',
                 G_agg='masked', optimizer='variational', dict_init=None, code_alpha=1, code_l1_ratio=1,
                 comp_l1_ratio=0, step_size=1, tol=1e-2, max_iter=100, code_pos=False, comp_pos=False,
                 random_state=None, n_epochs=1, n_components=10, batch_size=10, verbose=0, callback=None,
                 n_threads=1, rand_size=True, replacement=True, device=None, async_host_copy=False):
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.sample_learning_rate = sample_learning_rate
        self.Dx_agg = Dx_agg
        self.G_agg = G_agg
        self.reduction = reduction
        self.dict_init = dict_init
        self._set_coding_params(n_components, code_l1_ratio=code_l1_ratio, code_alpha=code_alpha,
                                code_pos=code_pos, random_state=random_state, tol=tol, max_iter=max_iter,
                                n_threads=n_threads)
        self.comp_l1_ratio = comp_l1_ratio
        self.comp_pos = comp_pos
        self.optimizer = optimizer
        self.step_size = step_size
        self.n_epochs = n_epochs
        self.verbose = verbose
        self.callback = callback
        self.n_threads = n_threads
        self.rand_size = rand_size
        self.replacement = replacement
        self.device = device
        self.async_host_copy = async_host_copy

    # ------------------------------------------------------------------ device views
    code_dev = property(lambda self: self.__dict__.get("_d_code_"))
    C_dev = property(lambda self: self.__dict__.get("_d_C_"))
    B_dev = property(lambda self: self.__dict__.get("_d_B_"))
    G_dev = property(lambda self: self.__dict__.get("_d_G_"))
    comp_norm_dev = property(lambda self: self.__dict__.get("_d_comp_norm_"))

    @property
    def gradient_(self):
        """The reference keeps `gradient_[:, subset] = B_[:, subset]` as scratch for the dictionary
        update [ref: :446, :532]; here that panel lives in the kernel workspace, so the attribute
        is reconstructed as a copy of B_ (its value on every column touched so far)."""
        return self.B_

    @property
    def time_(self):
        """Seconds of device time spent in partial_fit so far [ref: `time_`, :488, :505, :526].
        Reading it synchronises with the device."""
        ev = self.__dict__.setdefault("_time_events", [])
        tot = self.__dict__.get("_time_acc", 0.0)
        for start, end in ev:
            end.synchronize()
            tot += start.elapsed_time(end) * 1e-3
        self.__dict__["_time_events"] = []
        self.__dict__["_time_acc"] = tot
        return tot

    @time_.setter
    def time_(self, value):
        self.__dict__["_time_events"] = []
        self.__dict__["_time_acc"] = float(value)

    # ------------------------------------------------------------------ fit / partial_fit
    def fit(self, X):
        """[ref: dict_fact.py:286-311]"""
        tensor_in = isinstance(X, torch.Tensor)
        if not tensor_in:
            X = check_array(X, order='C', dtype=[np.float32, np.float64])
        if self.dict_init is None:
            dict_init = X
        else:
            dict_init = self.dict_init
            if not isinstance(dict_init, torch.Tensor):
                dict_init = check_array(dict_init, dtype=X.dtype.type if not tensor_in else None)
        self.prepare(n_samples=X.shape[0], X=dict_init,
                     dtype=None if not tensor_in else str(X.dtype).replace('torch.', ''))
        for _ in range(self.n_epochs):
            self.partial_fit(X)
            permutation = self.shuffle()
            X = X[torch.as_tensor(permutation, device=X.device)] if tensor_in else X[permutation]
        return self

    def partial_fit(self, X, sample_indices=None, code_out=None):
        """Update the factorisation with the rows of X, `batch_size` rows per step
        [ref: dict_fact.py:313-337].  X: NumPy array (host; pinned memory is copied without
        staging), CPU tensor or CUDA tensor.  The whole call is ONE call into the C ABI
        (`modl_partial_fit_*`: the loop over the batches, the host bookkeeping of every step and the
        two-stream schedule); with `verbose` / `callback` set, or in a subclass that overrides
        `_single_batch_fit`, the batches are walked here instead, one `modl_batch_fit_*` call each."""
        if self.__dict__.get("_d_components_") is None:
            raise AttributeError("call prepare() (or fit()) before partial_fit()")
        dev, dt = self._device, self._d_components_.dtype
        if isinstance(X, torch.Tensor):
            Xt = X if X.dtype == dt else X.to(dt)
        else:
            Xt = torch.from_numpy(_to_host_rows(X, self._np_dtype))
        n_samples, n_features = Xt.shape
        if n_features != self._d_components_.shape[1]:
            raise ValueError("X has %d features, the dictionary %d" % (n_features, self._d_components_.shape[1]))
        stream = torch.cuda.current_stream(dev)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        if self._c_loop_ok():
            self._partial_fit_c(Xt, sample_indices, code_out, stream)
        else:
            if code_out is not None:
                raise ValueError("code_out needs the C minibatch loop (no verbose / callback / overridden step)")
            batches = list(gen_batches(n_samples, self.batch_size))
            if Xt.is_cuda:
                for batch in batches:
                    self._single_batch_fit(Xt[batch], get_sub_slice(sample_indices, batch))
            else:
                self._partial_fit_host(Xt, batches, sample_indices, stream)
        end.record(stream)
        self.__dict__.setdefault("_time_events", []).append((start, end))
        if len(self._time_events) > 256:
            _ = self.time_
        return self

    # ------------------------------------------------------------------ the C minibatch loop
    def _c_loop_ok(self):
        """The C loop runs the step of THIS class; anything that hooks into the Python step (verbose printing, the
        callback, a subclass with its own `_single_batch_fit`) keeps the per-batch Python walk."""
        if self.verbose or self.callback is not None or self.__dict__.get("python_loop", False):
            return False
        return type(self)._single_batch_fit is DictFact._single_batch_fit

    def _fit_loop_handle(self):
        loop = self.__dict__.get("_fit_loop")
        if loop is None or loop.ctx is not self._ctx():
            loop = self.__dict__["_fit_loop"] = _lib.FitLoop(self._ctx())
        return loop

    def _fit_params(self, batch_rows):
        D = self._d_components_
        k, p = D.shape
        prm = _lib.FitParams()
        prm.n_samples, prm.n_features, prm.n_components, prm.batch_size = self._d_code_.shape[0], p, k, int(batch_rows)
        prm.components, prm.code = D.data_ptr(), self._d_code_.data_ptr()
        prm.C, prm.B, prm.comp_norm = self._d_C_.data_ptr(), self._d_B_.data_ptr(), self._d_comp_norm_.data_ptr()
        for name, field in (("_d_G_", "G_full"), ("_d_Dx_average_", "Dx_average"), ("_d_G_average_", "G_average")):
            t = self.__dict__.get(name)
            setattr(prm, field, t.data_ptr() if t is not None else None)
        prm.reduction, prm.learning_rate = float(self.reduction), float(self.learning_rate)
        prm.code_alpha, prm.code_l1_ratio = float(self.code_alpha), float(self.code_l1_ratio)
        prm.comp_l1_ratio, prm.tol, prm.step_size = float(self.comp_l1_ratio), float(self.tol), float(self.step_size)
        prm.max_iter, prm.code_pos, prm.comp_pos = int(self.max_iter), int(bool(self.code_pos)), int(bool(self.comp_pos))
        prm.Dx_agg, prm.G_agg = _lib.AGG[self.Dx_agg], _lib.AGG[self.G_agg]
        prm.optimizer_sgd = int(self.optimizer == 'sgd')
        return prm

    def _partial_fit_c(self, Xt, sample_indices, code_out, stream, world=1):
        """[ref: dict_fact.py:313-337 + :495-526] -- one `modl_partial_fit_*` call (per batch in the 'average'
        modes, whose per-sample weights are NumPy expressions of the counters, :513)."""
        dev = self._device
        n, p = Xt.shape
        k = self.n_components
        bs = int(self.batch_size)
        if n == 0:
            return
        if Xt.stride(1) != 1 or (Xt.is_cuda and Xt.stride(0) < p):
            Xt = Xt.contiguous()
        if sample_indices is None:
            idx = None
        else:
            idx = get_sub_slice(sample_indices, slice(0, n))
            idx = np.ascontiguousarray(idx, dtype=np.int64)
            if idx.shape[0] != n:
                raise ValueError("sample_indices and X do not match")
            if n and (idx.min() < 0 or idx.max() >= self._d_code_.shape[0]):
                raise IndexError("sample_indices out of range")
        n_batches = -(-n // bs)
        # the atom orders of this call, in batch order: the only draws the estimator's NumPy RandomState makes
        # inside partial_fit [ref: :672]
        orders = np.empty((n_batches, k), dtype=np.int64)
        for i in range(n_batches):
            orders[i] = self.random_state.permutation(k)
        loop = self._fit_loop_handle()
        fn = getattr(_lib.lib(), "modl_partial_fit_" + _lib.sfx_of(self._d_components_.dtype))
        io = _lib.FitBatches()
        io.ldx = Xt.stride(0)
        io.x_location = 0 if Xt.is_cuda else (1 if Xt.is_pinned() else 2)
        io.sampler = self.feature_sampler_._h
        n_iter = C.c_int64(int(self.n_iter_))
        io.h_n_iter = C.addressof(n_iter)
        counters = self.sample_n_iter_
        if counters.dtype != np.int64 or not counters.flags["C_CONTIGUOUS"]:
            counters = self.sample_n_iter_ = np.ascontiguousarray(counters, dtype=np.int64)
        io.h_sample_n_iter = counters.ctypes.data
        last_subset = np.empty(max(p, 1), dtype=np.int64)
        last_len = C.c_int64(0)
        io.h_last_subset, io.h_last_subset_len = last_subset.ctypes.data, C.addressof(last_len)
        sw = None
        if self.__dict__.get("record_sweeps", False):
            sw = self.__dict__.get("_d_sweeps")
            if sw is None or sw.shape[0] < bs:
                sw = self.__dict__["_d_sweeps"] = torch.zeros(bs, dtype=torch.int32, device=dev)
            io.sweeps = sw.data_ptr()
        keep = [Xt, idx, orders, counters, last_subset, sw]
        co = None
        if code_out is not None:
            co = code_out if isinstance(code_out, torch.Tensor) else torch.from_numpy(code_out)
            if co.is_cuda or co.dtype != self._d_components_.dtype or tuple(co.shape) != (n, k) or not co.is_contiguous():
                raise ValueError("code_out must be a contiguous host array of shape (n_rows, n_components) and the estimator dtype")
            io.h_code_out = co.data_ptr()
            keep.append(co)
        wait_host = 0 if getattr(self, "async_host_copy", False) else 1
        io.wait_host = wait_host
        # stream ordering against the caller (modl_fit_batches.fence): state written since the last call -> wait;
        # device rows declared final (`device_rows_final = True`) -> do not; otherwise automatic
        if self.__dict__.pop("_state_dirty", False):
            io.fence = 1
        elif Xt.is_cuda and self.__dict__.get("device_rows_final", False):
            io.fence = 2
        else:
            io.fence = 0
        itemsize = Xt.element_size()
        average = self.G_agg == 'average' or self.Dx_agg == 'average'
        st = C.c_void_p(stream.cuda_stream)
        handle = loop.handle

        def run(r0, r1, o0, upd, w_sample):
            prm = self._fit_params(bs)
            io.X = Xt.data_ptr() + r0 * Xt.stride(0) * itemsize
            io.n_rows = r1 - r0
            io.h_sample_indices = idx[r0:].ctypes.data if idx is not None else None
            if idx is None and r0:
                # rows r0.. of an un-indexed call are samples r0..: make that explicit for a partial call
                ids = np.arange(r0, r1, dtype=np.int64)
                keep.append(ids)
                io.h_sample_indices = ids.ctypes.data
            io.h_orders = orders[o0:].ctypes.data
            io.update_counters = int(upd)
            io.h_w_sample = w_sample.ctypes.data if w_sample is not None else None
            if co is not None:
                io.h_code_out = co.data_ptr() + r0 * k * itemsize
            _lib.check(fn(handle, C.byref(prm), C.byref(io), st))

        if not average:
            run(0, n, 0, True, None)
        else:
            for i in range(n_batches):
                r0, r1 = i * bs, min(n, (i + 1) * bs)
                rows = idx[r0:r1] if idx is not None else np.arange(r0, r1)
                n_iter.value += (r1 - r0) * world
                counters[rows] += 1
                w_sample = np.power(counters[rows], -self.sample_learning_rate).astype(self._np_dtype)   # [ref: :513]
                keep.append(w_sample)
                run(r0, r1, i, False, w_sample)
        self.n_iter_ = int(n_iter.value)
        self.__dict__["last_subset_"] = last_subset[:int(last_len.value)].copy()
        self.__dict__["last_order_"] = orders[-1].copy()
        # host buffers the asynchronous copies still read stay alive until the next synchronisation point
        if wait_host:
            keep = [Xt] if Xt.is_cuda else []
        self.__dict__.setdefault("_fit_keep", []).append(keep)
        del self._fit_keep[:-4]

    def _partial_fit_host(self, Xt, batches, sample_indices, stream):
        """Host rows -> device through two device staging slots on a side stream, so that the copy
        of batch i+1 overlaps the kernels of batch i -- also ACROSS calls: the slot counter and
        the events persist, and the call returns as soon as the host buffer has been consumed (its
        last copy has landed), not when the kernels are done."""
        dev = self._device
        pipe = self.__dict__.get("_pipeline")
        bs, p = self.batch_size, Xt.shape[1]
        if pipe is None or pipe["dev"][0].shape != (bs, p) or pipe["dev"][0].dtype != Xt.dtype:
            pipe = {
                "dev": [torch.empty((bs, p), dtype=Xt.dtype, device=dev) for _ in range(2)],
                "pin": None,
                "copy_stream": torch.cuda.Stream(device=dev),
                "copied": [torch.cuda.Event() for _ in range(2)],
                "done": [torch.cuda.Event() for _ in range(2)],
                "n": 0,
            }
            self.__dict__["_pipeline"] = pipe
        pinned = Xt.is_pinned()
        if not pinned and pipe["pin"] is None:
            pipe["pin"] = [torch.empty((bs, p), dtype=Xt.dtype, pin_memory=True) for _ in range(2)]
        cs = pipe["copy_stream"]
        for batch in batches:
            n = pipe["n"]
            slot = n & 1
            rows = Xt[batch]
            nb = rows.shape[0]
            if pinned:
                src = rows                                # DMA straight from the caller's pinned rows
            else:
                if n >= 2:
                    pipe["copied"][slot].synchronize()    # the previous DMA out of this staging buffer is over
                src = pipe["pin"][slot][:nb]
                src.copy_(rows)
            dst = pipe["dev"][slot][:nb]
            with torch.cuda.stream(cs):
                if n >= 2:
                    cs.wait_event(pipe["done"][slot])     # kernels of batch n-2 no longer read this slot
                dst.copy_(src, non_blocking=True)
                pipe["copied"][slot].record(cs)
            stream.wait_event(pipe["copied"][slot])
            self._single_batch_fit(dst, get_sub_slice(sample_indices, batch))
            pipe["done"][slot].record(stream)
            pipe["n"] = n + 1
        if pinned and not getattr(self, "async_host_copy", False):
            # the caller may reuse its buffer once the copies have read it
            for ev in pipe["copied"]:
                ev.synchronize()

    def synchronize(self):
        """Block until every enqueued copy and kernel of this estimator has finished."""
        pipe = self.__dict__.get("_pipeline")
        if pipe is not None:
            pipe["copy_stream"].synchronize()
        loop = self.__dict__.get("_fit_loop")
        if loop is not None:
            loop.synchronize()
        torch.cuda.current_stream(self._device).synchronize()
        self.__dict__["_fit_keep"] = []
        return self

    def set_params(self, **params):
        """[ref: dict_fact.py:339-357] -- including its quirk: only a switch to G_agg='full' is
        honoured for G_agg (other values are dropped), and nothing is returned."""
        G_agg = params.pop('G_agg', None)
        if G_agg == 'full' and self.G_agg != 'full':
            if self.__dict__.get("_d_components_") is not None:
                D = self._d_components_
                G = torch.empty((D.shape[0], D.shape[0]), dtype=D.dtype, device=D.device)
                self._gram_dx(D, None, G, None, None)
                self.__dict__["_d_G_"] = G
            self.G_agg = 'full'
        self.__dict__["_state_dirty"] = True
        BaseEstimator.set_params(self, **params)
        # A switch to Dx_agg='average' after prepare() (the 'gram' schedules of the front-ends,
        # image.py:142-144, fmri.py:497-499) finds no Dx_average_ in the reference and raises there;
        # here the running average starts from zero, as prepare() would have left it.
        if self.Dx_agg == 'average' and self.__dict__.get("_d_code_") is not None \
                and self.__dict__.get("_d_Dx_average_") is None:
            code = self._d_code_
            self.__dict__["_d_Dx_average_"] = torch.zeros(code.shape, dtype=code.dtype, device=code.device)

    def shuffle(self):
        """Shuffle the per-sample state with one permutation and return it
        [ref: dict_fact.py:359-379]."""
        random_seed = self.random_state.randint(MAX_INT)
        random_state = RandomState(random_seed)
        arrays = [self._d_code_]
        if self.G_agg == 'average':
            arrays.append(self._d_G_average_)
        if self.Dx_agg == 'average':
            arrays.append(self._d_Dx_average_)
        perm = random_state.shuffle_with_trace(arrays)
        self.__dict__["_state_dirty"] = True
        self.labels_ = self.labels_[perm]
        return perm

    def prepare(self, n_samples=None, n_features=None, dtype=None, X=None):
        """Allocate and initialise the estimator state [ref: dict_fact.py:381-489]."""
        dev = torch.device(self.device) if self.device is not None else default_device()
        if dev.type != 'cuda':
            raise RuntimeError("modl_b200 runs on CUDA devices only")
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        self.__dict__["_device"] = dev
        tensor_in = isinstance(X, torch.Tensor)
        if X is not None:
            if not tensor_in:
                X = check_array(X, order='C', dtype=[np.float32, np.float64])
            if dtype is None:
                dtype = X.dtype if not tensor_in else str(X.dtype).replace('torch.', '')
            this_n_samples = X.shape[0]
            if n_samples is None:
                n_samples = this_n_samples
            if n_features is None:
                n_features = X.shape[1]
            elif n_features != X.shape[1]:
                raise ValueError('n_features and X does not match')
        else:
            if n_features is None or n_samples is None:
                raise ValueError('Either provide shape or data to function prepare.')
            if dtype is None:
                dtype = np.float64
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError('dtype should be float32 or float64')
        if self.optimizer not in ['variational', 'sgd']:
            raise ValueError("optimizer should be 'variational' or 'sgd'")
        if self.optimizer == 'sgd':
            self.reduction = 1
            self.G_agg = 'full'
            self.Dx_agg = 'full'
        self.__dict__["_np_dtype"] = dtype
        tdt = torch_dtype(dtype)
        k = self.n_components
        kw = dict(dtype=tdt, device=dev)

        for name in ("G_average_", "Dx_average_", "G_"):
            self.__dict__["_d_" + name] = None
        if self.G_agg == 'average':
            self.__dict__["_d_G_average_"] = torch.zeros((n_samples, k, k), **kw)
        if self.Dx_agg == 'average':
            self.__dict__["_d_Dx_average_"] = torch.zeros((n_samples, k), **kw)
        self.__dict__["_d_C_"] = torch.zeros((k, k), **kw)
        self.__dict__["_d_B_"] = torch.zeros((k, n_features), **kw)

        self.random_state = check_random_state(self.random_state)
        if X is None:
            comp = torch.from_numpy(self.random_state.randn(k, n_features).astype(dtype)).to(dev)
        elif tensor_in:
            comp = X[:k].to(dev, tdt).clone()
        else:
            comp = torch.from_numpy(np.array(X[:k], dtype=dtype, order='C', copy=True)).to(dev)
        if comp.shape[0] != k:
            raise ValueError("need at least n_components=%d rows to initialise the dictionary" % k)
        comp = comp.contiguous()
        if self.comp_pos:
            comp = torch.where(comp <= 0, -comp, comp)
        self.__dict__["_d_components_"] = comp
        fn = getattr(_lib.lib(), "modl_enet_scale_" + _lib.sfx_of(tdt))
        _lib.check(fn(self._ctx().handle, ptr(comp), k, n_features, comp.stride(0), float(self.comp_l1_ratio),
                      1.0, stream_of(dev)))

        self.__dict__["_d_code_"] = torch.ones((n_samples, k), **kw)
        self.labels_ = np.arange(n_samples)
        self.__dict__["_d_comp_norm_"] = torch.zeros((k,), **kw)
        if self.G_agg == 'full':
            G = torch.empty((k, k), **kw)
            self._gram_dx(comp, None, G, None, None)
            self.__dict__["_d_G_"] = G
        self.n_iter_ = 0
        self.sample_n_iter_ = np.zeros(n_samples, dtype='int')
        self.random_state = check_random_state(self.random_state)
        random_seed = self.random_state.randint(MAX_INT)
        self.feature_sampler_ = Sampler(n_features, self.rand_size, self.replacement, random_seed)
        if self.verbose:
            self.verbose_iter_ = np.linspace(0, n_samples * self.n_epochs, self.verbose).tolist()
        self.time_ = 0
        # nothing cached for the previous fit (kernel entry points are dtype-specific, streams and scratch tensors
        # are device-specific) survives a new prepare()
        for name in ("_d_sweeps", "_pipeline", "_step_fn", "_prm", "_ovl", "_keepalive", "_d_inc_sub", "_d_inc", "_fit_keep"):
            self.__dict__.pop(name, None)
        self.__dict__["_d_sweeps"] = None
        self.__dict__["_pipeline"] = None
        self.__dict__["_state_dirty"] = True
        return self

    def _callback(self):
        if self.callback is not None:
            self.callback(self)

    # ------------------------------------------------------------------ the hot path
    def _host_bookkeeping(self, batch_rows, sample_indices):
        """The integer / scalar statements of a step, on the host, in the reference's order
        [ref: dict_fact.py:507-515, :672].  `batch_rows` is the size of the WHOLE minibatch
        (it differs from len(sample_indices) only under sample sharding)."""
        subset = self.feature_sampler_.yield_subset(self.reduction)            # [ref: :507]
        sample_indices = np.ascontiguousarray(sample_indices, dtype=np.int64)
        self.n_iter_ += batch_rows
        self.sample_n_iter_[sample_indices] += 1
        w = _batch_weight(self.n_iter_, batch_rows, self.learning_rate, 0)    # [ref: :515]
        order = np.ascontiguousarray(self.random_state.permutation(self.n_components), dtype=np.int64)  # [ref: :672]
        w_sample = None
        if self.G_agg == 'average' or self.Dx_agg == 'average':
            this_n_iter = self.sample_n_iter_[sample_indices]
            w_sample = np.power(this_n_iter, -self.sample_learning_rate).astype(self._np_dtype)   # [ref: :513]
        return subset, sample_indices, w, order, w_sample

    def _step_params(self, X, sample_indices, subset, order, w, w_sample, stats_inc=None, global_batch=0,
                     inc_sub=None):
        """Fill the C-ABI parameter block of one minibatch step (reused by every phase call of the step)."""
        dev = self._device
        D = self._d_components_
        k, p = D.shape
        b = X.shape[0]
        if X.stride(1) != 1:
            X = X.contiguous()
        w_dev = torch.from_numpy(w_sample).to(dev) if w_sample is not None else None
        itemsize = D.element_size()
        n_state = self._d_code_.shape[0]
        # contiguous ascending rows (the common case: `partial_fit(X)` / a slice of sample ids): address the
        # rows of the per-sample state through a pointer offset instead of uploading an index vector
        base = int(sample_indices[0]) if b > 0 else 0
        contiguous = b > 0 and int(sample_indices[-1]) - base == b - 1 and (b < 3 or bool(
            (np.diff(sample_indices) == 1).all()))
        idx_dev = None
        if not contiguous:
            idx_dev = torch.from_numpy(sample_indices).to(dev, non_blocking=True)
            base = 0
        prm = self.__dict__.get("_prm")
        if prm is None:
            prm = self.__dict__["_prm"] = _lib.StepParams()
        prm.n_samples, prm.n_features, prm.n_components, prm.batch_size = n_state - base, p, k, b
        prm.X, prm.ldx = X.data_ptr(), X.stride(0)
        prm.indices = idx_dev.data_ptr() if idx_dev is not None else None
        prm.h_subset, prm.subset_len = subset.ctypes.data, subset.shape[0]
        prm.h_order = order.ctypes.data
        prm.w_sample = w_dev.data_ptr() if w_dev is not None else None
        prm.w = w
        prm.components, prm.code = D.data_ptr(), self._d_code_.data_ptr() + base * k * itemsize
        prm.C, prm.B, prm.comp_norm = self._d_C_.data_ptr(), self._d_B_.data_ptr(), self._d_comp_norm_.data_ptr()
        t = self.__dict__.get("_d_G_")
        prm.G_full = t.data_ptr() if t is not None else None
        t = self.__dict__.get("_d_Dx_average_")
        prm.Dx_average = t.data_ptr() + base * k * itemsize if t is not None else None
        t = self.__dict__.get("_d_G_average_")
        prm.G_average = t.data_ptr() + base * k * k * itemsize if t is not None else None
        prm.reduction, prm.code_alpha = float(self.reduction), float(self.code_alpha)
        prm.code_l1_ratio, prm.comp_l1_ratio = float(self.code_l1_ratio), float(self.comp_l1_ratio)
        prm.tol, prm.step_size, prm.max_iter = float(self.tol), float(self.step_size), int(self.max_iter)
        prm.code_pos, prm.comp_pos = int(bool(self.code_pos)), int(bool(self.comp_pos))
        prm.Dx_agg, prm.G_agg = _lib.AGG[self.Dx_agg], _lib.AGG[self.G_agg]
        prm.optimizer_sgd = int(self.optimizer == 'sgd')
        prm.phases, prm.global_batch = 0, int(global_batch)
        prm.stats_inc = stats_inc.data_ptr() if stats_inc is not None else None
        prm.inc_sub = inc_sub.data_ptr() if inc_sub is not None else None
        prm.ev_after_apply_sub = None
        sw = self.__dict__.get("_d_sweeps")
        if self.__dict__.get("record_sweeps", False):
            if sw is None or sw.shape[0] < b:
                sw = self.__dict__["_d_sweeps"] = torch.zeros(max(b, self.batch_size), dtype=torch.int32, device=dev)
            prm.sweeps = sw.data_ptr()
        else:
            prm.sweeps = None
        # keep the step's inputs alive until the (asynchronous) kernels have consumed them
        self.__dict__.setdefault("_keepalive", []).append((idx_dev, w_dev, X, subset, order))
        del self._keepalive[:-8]
        return prm

    def _run_phases(self, prm, phases=0, stream=None):
        """One call into the C ABI (`modl_batch_fit_*`): the selected device phases of the step."""
        prm.phases = int(phases)
        fn = self.__dict__.get("_step_fn")
        if fn is None:
            fn = self.__dict__["_step_fn"] = getattr(_lib.lib(), "modl_batch_fit_" + _lib.sfx_of(self._d_components_.dtype))
        st = stream_of(self._device) if stream is None else C.c_void_p(stream.cuda_stream)
        _lib.check(fn(self._ctx().handle, C.byref(prm), st))

    def _launch_step(self, X, sample_indices, subset, order, w, w_sample, phases=0, stats_inc=None,
                     global_batch=0, inc_sub=None, stream=None):
        prm = self._step_params(X, sample_indices, subset, order, w, w_sample, stats_inc, global_batch, inc_sub)
        self._run_phases(prm, phases, stream)

    def _single_batch_fit(self, X, sample_indices):
        """One minibatch step; X is a (batch x n_features) CUDA tensor of the estimator dtype
        [ref: dict_fact.py:495-526]."""
        if self.verbose and self.verbose_iter_ and self.n_iter_ >= self.verbose_iter_[0]:
            print('Iteration %i' % self.n_iter_)
            self.verbose_iter_ = self.verbose_iter_[1:]
            self._callback()
        subset, sample_indices, w, order, w_sample = self._host_bookkeeping(X.shape[0], sample_indices)
        if self.__dict__.get("overlap_stats", False):
            self._overlapped_local_step(X, sample_indices, subset, order, w, w_sample)
        else:
            self._launch_step(X, sample_indices, subset, order, w, w_sample)
        self.__dict__["last_subset_"] = subset
        self.__dict__["last_order_"] = order

    # The dictionary update (sequential, 16 SMs) only needs C_ and the SUBSET columns of B_
    # [ref: dict_fact.py:532], while `_update_B` touches all p columns [ref: :559-566].  With
    # `est.overlap_stats = True` the step computes the k x s slice first (MODL_PHASE_STATS_SUB), starts the
    # dictionary update, and runs the full-width product B_ = (1-w) B_ + w/b code^T X on a second stream.
    # Off by default on one GPU: measured 0.519 ms/step vs 0.484 ms for the fused call -- the 16-CTA
    # cluster of the dictionary update cannot be placed while the 140 one-per-SM GEMM CTAs are resident,
    # so the two serialise and the split only adds launches.  The sharded estimator uses the same phases
    # to hide its all-reduce (distributed.py).
    def _side_state(self):
        st = self.__dict__.get("_ovl")
        if st is None:
            st = self.__dict__["_ovl"] = {"stream": torch.cuda.Stream(device=self._device), "ev_code": torch.cuda.Event(),
                                          "ev_sub": torch.cuda.Event(), "ev_applied": torch.cuda.Event()}
            st["ev_sub"].record(torch.cuda.current_stream(self._device))   # creates the CUDA event handle the C side records
        return st

    def _inc_sub_buffer(self, s):
        D = self._d_components_
        k = D.shape[0]
        need = k * k + k * (4 * ((s + 3) // 4) if s > 0 else 4)
        buf = self.__dict__.get("_d_inc_sub")
        if buf is None or buf.numel() < need or buf.dtype != D.dtype:
            buf = self.__dict__["_d_inc_sub"] = torch.zeros(need + need // 4, dtype=D.dtype, device=D.device)
        return buf[:need]

    def _overlapped_local_step(self, X, sample_indices, subset, order, w, w_sample):
        st = self._side_state()
        main, side = torch.cuda.current_stream(self._device), st["stream"]
        prm = self._step_params(X, sample_indices, subset, order, w, w_sample,
                                inc_sub=self._inc_sub_buffer(subset.shape[0]))
        self._run_phases(prm, _lib.PHASE_CODE | _lib.PHASE_STATS_SUB)
        st["ev_code"].record(main)
        prm.ev_after_apply_sub = st["ev_sub"].cuda_event     # recorded inside the call: B_[:, subset] has been read
        self._run_phases(prm, _lib.PHASE_APPLY_SUB | _lib.PHASE_DICT | _lib.PHASE_REUSE_SUBSET)
        side.wait_event(st["ev_sub"])
        self._run_phases(prm, _lib.PHASE_STATS_B, stream=side)
        st["ev_applied"].record(side)
        main.wait_event(st["ev_applied"])           # ends long before the dictionary update: costs nothing, orders everything

    @property
    def last_sweeps_(self):
        """Per-sample CD sweep counts of the last minibatch (needs `est.record_sweeps = True`)."""
        sw = self.__dict__.get("_d_sweeps")
        return None if sw is None else sw.cpu().numpy()


class Coder(CodingMixin, BaseEstimator):
    """Sparse coder over a fixed dictionary [ref: dict_fact.py:724-745]."""

    def __init__(self, dictionary, code_alpha=1, code_l1_ratio=1, tol=1e-2, max_iter=100, code_pos=False,
                 random_state=None, n_threads=1, device=None):
        self.dictionary = dictionary
        self.device = device
        if device is not None:           # the dictionary lands on this device, not on whichever one is current
            dev = torch.device(device)
            self.__dict__["_device"] = dev if dev.index is not None else torch.device('cuda', torch.cuda.current_device())
        self._set_coding_params(dictionary.shape[0], code_l1_ratio=code_l1_ratio, code_alpha=code_alpha,
                                code_pos=code_pos, random_state=random_state, tol=tol, max_iter=max_iter,
                                n_threads=n_threads)
        self.components_ = dictionary

    def fit(self, X=None):
        return self
