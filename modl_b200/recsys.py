"""RecsysDictFact -- the reference's matrix-completion estimator on the B200 hot path
[ref: modl/decomposition/recsys.py:16-330, modl/decomposition/recsys_fast.pyx:10-37].

"Next" row 3 of SURVEY section 8f: the missing-value path.  Same constructor keywords, fitted attributes
(`components_`, `code_`, `C_`, `B_`, `comp_norm_`, `feature_n_iter_`, `feature_freq_`, `n_iter_`,
`row_mean_`, `col_mean_`) and NumPy random stream (initial dictionary, epoch permutations, atom orders) as
the reference, plus one keyword, `device`.

The reference visits the rows of a minibatch one at a time (`_single_sample_update`, recsys.py:164-185).
Here a minibatch is a handful of launches over the CSR matrix, which stays resident in HBM:

  1. per-row Gram / correlation over the row's observed columns  (modl_recsys_gram_dx_*, one CTA per row)
  2. per-row Cholesky solve                                      (modl_enet_regression_multi_gram_*, ridge branch)
  3. the batch's entries ordered by (column, position in batch)  (index bookkeeping: torch sort / unique)
  4. per-column B_ recurrence with the feature counters          (modl_recsys_update_B_*)
  5. C_ update                                                   (modl_recsys_update_C_*)
  6. dictionary update on the union of observed columns          (modl_update_dict_*, the DictFact kernel,
     comp_l1_ratio = 0: the plain L2-ball projection of recsys.py:201-206)
  7. refresh of the transposed dictionary copy                   (modl_recsys_sync_transposed_*)

`linalg.solve` (LU) of the reference becomes a Cholesky factorisation: the matrices are SPD by construction
(Gram + alpha/reduction I); results agree to rounding.  There is no CPU fallback.
"""
from math import ceil, log

import numpy as np
import scipy.sparse as sp
import torch
from sklearn.base import BaseEstimator
from sklearn.utils import check_array, check_random_state, gen_batches

from . import _lib
from ._util import default_device, ptr, stream_of, torch_dtype
from .dict_fact_fast import _batch_weight

__all__ = ["RecsysDictFact", "compute_biases", "rmse"]


class _DeviceKernels(object):
    """The C-ABI calls of the path (include/modl_b200.h, "Recsys" block), on one device."""

    def __init__(self, device, tdt):
        self.device = device
        self.sfx = _lib.sfx_of(tdt)
        self.ctx = _lib.get_context(device.index)
        self.lib = _lib.lib()

    def _call(self, name, *args):
        _lib.check(getattr(self.lib, name + self.sfx)(self.ctx.handle, *args, stream_of(self.device)))

    def gram_dx(self, Dt, indptr, indices, data, rows, row0, b, p, alpha, G, Dx):
        k = Dt.shape[1]
        self._call("modl_recsys_gram_dx_", ptr(Dt), Dt.stride(0), ptr(indptr), ptr(indices), ptr(data), ptr(rows),
                   int(row0), int(b), k, int(p), float(alpha), ptr(G), ptr(Dx))

    def solve(self, G, Dx, code, rows, b):
        """code[rows] = G[ii]^-1 Dx[ii] (Cholesky; Dx is overwritten).  rows None: code is the b x k batch."""
        k = code.shape[1]
        self._call("modl_enet_regression_multi_gram_", ptr(G), ptr(Dx), None, 0, 0, None, ptr(code), ptr(rows),
                   int(b), k, 0.0, 0.0, 0, 0.0, 0, None)

    def check_solves(self):
        """Raise if any Cholesky since the last check met a non-positive pivot (sticky device flag, one sync)."""
        self.ctx.check_info(torch.cuda.current_stream(self.device).cuda_stream)

    def update_B(self, B, code, subset, col_ptr, entry_row, entry_val, feature_n_iter, w, n_iter):
        self._call("modl_recsys_update_B_", ptr(B), B.stride(0), ptr(code), ptr(subset), ptr(col_ptr), ptr(entry_row),
                   ptr(entry_val), ptr(feature_n_iter), subset.shape[0], code.shape[1], float(w), int(n_iter))

    def update_C(self, C, code, rows, w):
        self._call("modl_recsys_update_C_", ptr(C), ptr(code), ptr(rows), rows.shape[0], code.shape[1], float(w))

    def update_dict(self, D, B, C, comp_norm, subset, order):
        k, p = D.shape
        self._call("modl_update_dict_", ptr(D), D.stride(0), ptr(B), B.stride(0), ptr(C), ptr(comp_norm), None,
                   ptr(subset), subset.shape[0], order.ctypes.data, k, p, 0.0, 0, 0, 0.0, 0.0)

    def sync_transposed(self, D, Dt, subset):
        s = D.shape[1] if subset is None else subset.shape[0]
        self._call("modl_recsys_sync_transposed_", ptr(D), D.stride(0), ptr(Dt), Dt.stride(0), ptr(subset), s,
                   D.shape[0])

    def predict(self, code, Dt, indptr, indices, out):
        self._call("modl_recsys_predict_", ptr(code), ptr(Dt), Dt.stride(0), ptr(indptr), ptr(indices),
                   indptr.shape[0] - 1, code.shape[1], ptr(out))


def _kernels_for(device, tdt):
    """Factory of the kernel set (one indirection, so that the CPU tests of the host logic can stand a
    NumPy restatement of the kernels' contracts in its place)."""
    if device.type != 'cuda':
        raise RuntimeError("modl_b200 runs on CUDA devices only; there is no CPU fallback")
    return _DeviceKernels(device, tdt)


class _DeviceCSR(object):
    """A CSR matrix resident on the device: indptr int64, indices int32, data in the estimator dtype.
    `indptr_host` stays on the host for the per-batch sizes (no device round trip for them)."""

    def __init__(self, X, tdt, device):
        self.shape = X.shape
        self.indptr_host = np.ascontiguousarray(X.indptr, dtype=np.int64)
        self.indptr = torch.from_numpy(self.indptr_host).to(device)
        self.indices = torch.from_numpy(np.ascontiguousarray(X.indices, dtype=np.int32)).to(device)
        self.data = torch.from_numpy(np.ascontiguousarray(X.data)).to(device=device, dtype=tdt)


def _as_csr(X, dtype=None):
    if not sp.issparse(X):
        X = sp.csr_matrix(X)
    if dtype is None:
        return check_array(X, accept_sparse='csr')
    return check_array(X, accept_sparse='csr', dtype=dtype, copy=True)


class RecsysDictFact(BaseEstimator):
    """Matrix factorisation of a sparse ratings matrix by masked online dictionary learning
    [ref: recsys.py:16-84].  alpha: ridge penalty of the codes; batch_size None = ceil(1 / density)."""

    def __init__(self, alpha=1.0, beta=.0, n_components=30, learning_rate=1., batch_size=1, dict_init=None,
                 l1_ratio=0, n_epochs=1, random_state=None, verbose=0, detrend=False, crop=None, callback=None,
                 device=None, bookkeeping='batch'):
        self.callback = callback
        self.verbose = verbose
        self.random_state = random_state
        self.n_epochs = n_epochs
        self.l1_ratio = l1_ratio
        self.dict_init = dict_init
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.n_components = n_components
        self.alpha = alpha
        self.beta = beta
        self.detrend = detrend
        self.crop = crop
        self.device = device
        # 'batch': the entry ordering of a minibatch is built when the minibatch is visited (one device sort and
        # one host synchronisation per minibatch); 'epoch': built for many minibatches at once (one sort and one
        # synchronisation per chunk of up to 2**26 entries), which takes the index work off the per-batch path
        self.bookkeeping = bookkeeping

    # ---------------------------------------------------------------- state: device tensors, NumPy views
    def _state(self, name):
        t = self.__dict__.get("_d_" + name)
        if t is None:
            raise AttributeError(name)
        return t.detach().cpu().numpy()

    components_ = property(lambda self: self._state("components_"))
    code_ = property(lambda self: self._state("code_"))
    C_ = property(lambda self: self._state("C_"))
    B_ = property(lambda self: self._state("B_"))
    comp_norm_ = property(lambda self: self._state("comp_norm_"))
    feature_n_iter_ = property(lambda self: self._state("feature_n_iter_"))

    def _callback(self):
        if self.callback is not None:
            self.callback(self)

    # ---------------------------------------------------------------- fit
    def fit(self, X, y=None):
        """Learn a dictionary from the sparse matrix X (n_samples, n_features) [ref: recsys.py:86-145]."""
        X = _as_csr(X, dtype=[np.float32, np.float64])
        dtype = X.dtype
        n_samples, n_features = X.shape
        if np.any(np.diff(X.indptr) == 0):
            # the reference divides by the number of observed entries of every row in _refit (:259)
            raise ZeroDivisionError("RecsysDictFact.fit: a row of X has no observed entry")
        dev = torch.device(self.device) if self.device is not None else default_device()
        if dev.type == 'cuda' and dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        tdt = torch_dtype(dtype)
        self.__dict__["_device"] = dev
        self.__dict__["_kernels"] = kern = _kernels_for(dev, tdt)

        self.random_state = check_random_state(self.random_state)

        if self.detrend:
            self.row_mean_, self.col_mean_ = compute_biases(X, beta=self.beta, inplace=False)
            X.data -= np.repeat(self.row_mean_, np.diff(X.indptr))
            X.data -= self.col_mean_.take(X.indices, mode='clip')

        k = self.n_components
        components = self.random_state.randn(k, n_features).astype(dtype)
        S = np.sqrt(np.sum(components ** 2, axis=1))
        components /= S[:, np.newaxis]
        kw = dict(dtype=tdt, device=dev)
        self.__dict__["_d_components_"] = D = torch.from_numpy(components).to(dev)
        self.__dict__["_d_Dt"] = Dt = torch.empty((n_features, k), **kw)
        kern.sync_transposed(D, Dt, None)
        self.__dict__["_d_code_"] = torch.zeros((n_samples, k), **kw)
        Xd = _DeviceCSR(X, tdt, dev)
        self._refit(Xd)

        self.feature_freq_ = np.bincount(X.indices) / n_samples
        self.__dict__["_d_feature_n_iter_"] = torch.zeros((n_features,), dtype=torch.int64, device=dev)

        sparsity = X.nnz / n_samples / n_features
        if self.batch_size is None:
            batch_size = int(ceil(1. / sparsity))
        else:
            batch_size = self.batch_size

        self.__dict__["_d_comp_norm_"] = torch.zeros((k,), **kw)
        self.__dict__["_d_C_"] = torch.zeros((k, k), **kw)
        self.__dict__["_d_B_"] = torch.zeros((k, n_features), **kw)
        self.n_iter_ = 0

        if self.verbose:
            log_lim = log(n_samples * self.n_epochs / batch_size, 10)
            self.verbose_iter_ = ((np.logspace(0, log_lim, self.verbose, base=10) - 1) * batch_size).tolist()

        if self.bookkeeping not in ('batch', 'epoch'):
            raise ValueError("bookkeeping should be 'batch' or 'epoch'")
        for _ in range(self.n_epochs):
            permutation = self.random_state.permutation(n_samples)
            if self.bookkeeping == 'epoch':
                for entries in self._epoch_entries(Xd, permutation, batch_size):
                    self._batch_step(Xd, *entries)
            else:
                for batch in gen_batches(n_samples, batch_size):
                    self._single_batch_fit(Xd, permutation[batch])
            self._check_solves()
        self._refit(Xd)
        return self

    def _refit(self, Xd, chunk_rows=4096):
        """code_[i] for EVERY row on the current dictionary [ref: recsys.py:254-265]."""
        kern = self._kernels
        code, Dt = self._d_code_, self._d_Dt
        n_samples, n_features = Xd.shape
        k = code.shape[1]
        rows_per_call = max(1, min(int(chunk_rows), n_samples))
        G = torch.empty((rows_per_call, k, k), dtype=code.dtype, device=code.device)
        Dx = torch.empty((rows_per_call, k), dtype=code.dtype, device=code.device)
        for sl in gen_batches(n_samples, rows_per_call):
            b = sl.stop - sl.start
            kern.gram_dx(Dt, Xd.indptr, Xd.indices, Xd.data, None, sl.start, b, n_features, self.alpha, G, Dx)
            kern.solve(G, Dx, code[sl], None, b)
        self._check_solves()

    def _check_solves(self):
        """The reference solves every row with `linalg.solve`, which raises on a singular system [ref: recsys.py:178,
        :265]; the device Cholesky records a non-positive pivot in a sticky flag instead, read here (one
        synchronisation per refit / per epoch)."""
        self._kernels.check_solves()

    def _batch_entries(self, Xd, batch):
        """The stored entries of the rows `batch`, ordered by (column, position of the row in the batch):
        -> rows (device int64[b]), subset (sorted distinct columns, recsys.py:159-161), col_ptr, entry_row,
        entry_val.  Index bookkeeping only; sizes come from the host copy of indptr."""
        dev = self._device
        batch = np.ascontiguousarray(batch, dtype=np.int64)
        starts = Xd.indptr_host[batch]
        lens = Xd.indptr_host[batch + 1] - starts
        total = int(lens.sum())
        rows = torch.from_numpy(batch).to(dev)
        lens_d = torch.from_numpy(lens).to(dev)
        first = torch.from_numpy(starts - (np.cumsum(lens) - lens)).to(dev)        # entry e of row ii sits at first[ii] + e
        pos = torch.repeat_interleave(torch.arange(batch.shape[0], device=dev), lens_d, output_size=total)
        src = first[pos] + torch.arange(total, device=dev)
        cols, by_col = torch.sort(Xd.indices[src].to(torch.int64), stable=True)
        subset, counts = torch.unique_consecutive(cols, return_counts=True)
        col_ptr = torch.zeros((subset.shape[0] + 1,), dtype=torch.int64, device=dev)
        torch.cumsum(counts, dim=0, out=col_ptr[1:])
        src = src[by_col]
        return rows, subset, col_ptr, rows[pos[by_col]], Xd.data[src]

    def _epoch_entries(self, Xd, permutation, batch_size, max_entries=1 << 26):
        """`_batch_entries` for every minibatch of an epoch, the index work done for MANY minibatches at once:
        the stored entries of a chunk of consecutive minibatches are ordered by (minibatch, column, position)
        with one stable sort of the key `minibatch * n_features + column`; the distinct keys are the subsets
        of all the minibatches back to back.  Yields the same tuples as `_batch_entries` (the entry arrays
        are shared by the chunk, `col_ptr` holds offsets into them).  One host synchronisation per chunk."""
        dev = self._device
        n_features = Xd.shape[1]
        permutation = np.ascontiguousarray(permutation, dtype=np.int64)
        n = permutation.shape[0]
        starts_all = Xd.indptr_host[permutation]
        lens_all = Xd.indptr_host[permutation + 1] - starts_all
        # chunks of whole minibatches holding at most max_entries stored entries (at least one minibatch)
        per_batch = np.add.reduceat(lens_all, np.arange(0, n, batch_size))
        chunks, lo, acc = [], 0, 0
        for j, cnt in enumerate(per_batch):
            if j > lo and acc + cnt > max_entries:
                chunks.append((lo, j))
                lo, acc = j, 0
            acc += cnt
        chunks.append((lo, per_batch.shape[0]))
        for b0, b1 in chunks:
            p0, p1 = b0 * batch_size, min(n, b1 * batch_size)
            starts, lens = starts_all[p0:p1], lens_all[p0:p1]
            total = int(lens.sum())
            rows = torch.from_numpy(permutation[p0:p1]).to(dev)
            lens_d = torch.from_numpy(lens).to(dev)
            first = torch.from_numpy(starts - (np.cumsum(lens) - lens)).to(dev)
            pos = torch.repeat_interleave(torch.arange(p1 - p0, device=dev), lens_d, output_size=total)
            src = first[pos] + torch.arange(total, device=dev)
            key = torch.div(pos, batch_size, rounding_mode='floor') * n_features + Xd.indices[src].to(torch.int64)
            key, by_key = torch.sort(key, stable=True)
            ukey, counts = torch.unique_consecutive(key, return_counts=True)
            ubatch = torch.div(ukey, n_features, rounding_mode='floor')
            subset_all = ukey - ubatch * n_features
            col_ptr_all = torch.zeros((ukey.shape[0] + 1,), dtype=torch.int64, device=dev)
            torch.cumsum(counts, dim=0, out=col_ptr_all[1:])
            per = torch.bincount(ubatch, minlength=b1 - b0).cpu().numpy()        # the one synchronisation
            offs = np.concatenate([[0], np.cumsum(per)])
            src = src[by_key]
            entry_row, entry_val = rows[pos[by_key]], Xd.data[src]
            for j in range(b1 - b0):
                r0, r1 = j * batch_size, min(p1 - p0, (j + 1) * batch_size)
                yield (rows[r0:r1], subset_all[offs[j]:offs[j + 1]], col_ptr_all[offs[j]:offs[j + 1] + 1],
                       entry_row, entry_val)

    def _single_batch_fit(self, Xd, batch):
        """One minibatch [ref: recsys.py:151-213]."""
        self._batch_step(Xd, *self._batch_entries(Xd, batch))

    def _batch_step(self, Xd, rows, subset, col_ptr, entry_row, entry_val):
        if self.verbose and self.verbose_iter_ and self.n_iter_ >= self.verbose_iter_[0]:
            print('Iteration %i' % self.n_iter_)
            self.verbose_iter_ = self.verbose_iter_[1:]
            self._callback()
        kern = self._kernels
        n_features = Xd.shape[1]
        code, D, Dt = self._d_code_, self._d_components_, self._d_Dt
        k = code.shape[1]
        batch_size = rows.shape[0]
        self.n_iter_ += batch_size
        w = _batch_weight(self.n_iter_, batch_size, self.learning_rate, 0)

        # codes of the batch on the current dictionary (:169-178); independent across rows
        G = torch.empty((batch_size, k, k), dtype=code.dtype, device=code.device)
        Dx = torch.empty((batch_size, k), dtype=code.dtype, device=code.device)
        kern.gram_dx(Dt, Xd.indptr, Xd.indices, Xd.data, rows, 0, batch_size, n_features, self.alpha, G, Dx)
        kern.solve(G, Dx, code, rows, batch_size)
        # B_ and the feature counters (:168, :182-185), then C_ (:156-157)
        kern.update_B(self._d_B_, code, subset, col_ptr, entry_row, entry_val, self._d_feature_n_iter_, w, self.n_iter_)
        kern.update_C(self._d_C_, code, rows, w)
        # dictionary update on the union of the observed columns (:159-162, :187-213)
        order = np.ascontiguousarray(self.random_state.permutation(k), dtype=np.int64)
        kern.update_dict(D, self._d_B_, self._d_C_, self._d_comp_norm_, subset, order)
        kern.sync_transposed(D, Dt, subset)

    # ---------------------------------------------------------------- predict / score
    def predict(self, X):
        """Values of the factorisation at the stored entries of X, as a CSR matrix with the same structure
        [ref: recsys.py:215-244].  The values are floating point whatever the dtype of X.data (the reference computes them
        in a double buffer, recsys_fast.pyx:10-37)."""
        X = _as_csr(X)
        if X.shape != (self._d_code_.shape[0], self._d_components_.shape[1]):
            raise ValueError("X has shape %s, the model %s" % (X.shape, (self._d_code_.shape[0],
                                                                          self._d_components_.shape[1])))
        dev = self._device
        indptr = torch.from_numpy(np.ascontiguousarray(X.indptr, dtype=np.int64)).to(dev)
        indices = torch.from_numpy(np.ascontiguousarray(X.indices, dtype=np.int32)).to(dev)
        out_d = torch.zeros((X.nnz,), dtype=torch.float64, device=dev)
        self._kernels.predict(self._d_code_, self._d_Dt, indptr, indices, out_d)
        out_dtype = X.data.dtype if X.data.dtype in (np.float32, np.float64) else np.float64
        out = out_d.cpu().numpy().astype(out_dtype, copy=False)
        if self.detrend:
            out += np.repeat(self.row_mean_, np.diff(X.indptr))
            out += self.col_mean_.take(X.indices, mode='clip')
        if self.crop is not None:
            out[out > self.crop[1]] = self.crop[1]
            out[out < self.crop[0]] = self.crop[0]
        return sp.csr_matrix((out, X.indices, X.indptr), shape=X.shape)

    def score(self, X):
        """Root mean squared error of the predictions at the stored entries of X [ref: recsys.py:246-252]."""
        X = _as_csr(X)
        return rmse(X, self.predict(X))


def compute_biases(X, beta=0, inplace=False):
    """Row and column offsets of a CSR matrix by two rounds of alternating centring
    [ref: recsys.py:268-303]; with `inplace` the data of X is centred as a side effect."""
    if not inplace:
        X = X.copy()
    X = sp.csr_matrix(X)
    acc_u = np.zeros(X.shape[0])
    acc_m = np.zeros(X.shape[1])
    n_u = X.getnnz(axis=1)
    n_m = X.getnnz(axis=0)
    n_u[n_u == 0] = 1
    n_m[n_m == 0] = 1
    row_len = np.diff(X.indptr)
    average_rating = np.mean(X.data)
    for _ in range(2):
        w_u = (np.asarray(X.sum(axis=1))[:, 0] + average_rating * beta) / (n_u + beta)
        X.data -= np.repeat(w_u, row_len)
        w_m = np.asarray(X.sum(axis=0))[0] / (n_m + beta)
        X.data -= w_m.take(X.indices, mode='clip')
        acc_u += w_u
        acc_m += w_m
    return acc_u, acc_m


def rmse(X_true, X_pred):
    """Root mean squared error between two sparse matrices of identical structure [ref: recsys.py:306-311]."""
    X_true = check_array(X_true, accept_sparse='csr')
    X_pred = check_array(X_pred, accept_sparse='csr')
    mse = np.mean((X_true.data - X_pred.data) ** 2)
    return np.sqrt(mse)
