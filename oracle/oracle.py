"""CPU oracle for the MODL minibatch hot path -- TEST INFRASTRUCTURE, not product.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module; nothing under modl_b200/ does.  Parity status: PINNED (see
oracle/modl_oracle.c header and tests/test_oracle.py).

Two layers:
  * thin ctypes/numpy wrappers over oracle/modl_oracle.c (compiled with gcc into
    oracle/_build/libmodl_oracle.so on first use), mirroring the names the reference's
    estimator imports (modl/decomposition/dict_fact.py:13-18);
  * `OracleDictFact`, a NumPy restatement of the estimator state machine
    (modl/decomposition/dict_fact.py:286-715): prepare / partial_fit / the
    per-minibatch step / transform / score / shuffle;
  * `OracleRecsysDictFact`, the same for the matrix-completion estimator
    (modl/decomposition/recsys.py, recsys_fast.pyx).
"""
import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libmodl_oracle.so")
_SRC = [os.path.join(_HERE, "modl_oracle.c"), os.path.join(_HERE, "modl_oracle_real.inc")]

MAX_INT = np.iinfo(np.int64).max


def build(force=False):
    """Compile the C restatement (gcc, -O2, no fast-math so float semantics hold)."""
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in _SRC)):
        return _SO
    os.makedirs(_BUILD, exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
           "-o", _SO, _SRC[0], "-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i64, f64, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
        L.orc_rng_new.restype = vp
        L.orc_rng_new.argtypes = [C.c_uint64]
        L.orc_rng_free.argtypes = [vp]
        L.orc_interval.restype = C.c_uint64
        L.orc_interval.argtypes = [vp, C.c_uint64]
        L.orc_double.restype = f64
        L.orc_double.argtypes = [vp]
        L.orc_binomial.restype = C.c_long
        L.orc_binomial.argtypes = [vp, C.c_long, f64]
        L.orc_shuffle_i64.argtypes = [vp, vp, C.c_long]
        L.orc_shuffle_trace.argtypes = [vp, C.c_long, vp, vp]
        L.orc_permutation.argtypes = [vp, vp, C.c_long]
        L.orc_sampler_new.restype = vp
        L.orc_sampler_new.argtypes = [C.c_long, i32, i32, C.c_uint64]
        L.orc_sampler_free.argtypes = [vp]
        L.orc_sampler_yield.restype = C.c_long
        L.orc_sampler_yield.argtypes = [vp, f64, vp]
        L.orc_batch_weight.restype = f64
        L.orc_batch_weight.argtypes = [C.c_long, C.c_long, f64, f64]
        for sfx, real in (("f32", C.c_float), ("f64", C.c_double)):
            f = getattr(L, "orc_enet_norm_" + sfx)
            f.restype = real
            f.argtypes = [vp, C.c_long, real]
            getattr(L, "orc_enet_scale_" + sfx).argtypes = [vp, C.c_long, real, real]
            getattr(L, "orc_enet_projection_" + sfx).argtypes = [vp, vp, C.c_long, real, real]
            f = getattr(L, "orc_cd_gram_" + sfx)
            f.restype = i32
            f.argtypes = [vp, real, real, vp, vp, vp, C.c_long, C.c_long, vp, vp, i32, real, i32]
            for nm in ("orc_regression_single_gram_", "orc_regression_multi_gram_"):
                f = getattr(L, nm + sfx)
                f.restype = i32
                f.argtypes = [vp, vp, vp, vp, vp, C.c_long, C.c_long, C.c_long,
                              real, real, i32, real, i32, vp]
            getattr(L, "orc_update_G_average_" + sfx).argtypes = [vp, vp, vp, C.c_long, C.c_long]
            getattr(L, "orc_update_dict_panel_" + sfx).argtypes = [
                vp, vp, vp, vp, vp, C.c_long, C.c_long, real, i32]
        _lib = L
    return _lib


def _sfx(a):
    if a.dtype == np.float32:
        return "f32"
    if a.dtype == np.float64:
        return "f64"
    raise TypeError("float32 or float64 expected, got %s" % a.dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _chk(a, dtype=None):
    assert a.flags["C_CONTIGUOUS"], "C-contiguous array required"
    if dtype is not None:
        assert a.dtype == dtype
    return a


# --------------------------------------------------------------------------- #
# random stream (modl/utils/randomkit/random_fast.pyx, sampler.pyx)
# --------------------------------------------------------------------------- #
class RandomState(object):
    def __init__(self, seed):
        self.initial_seed = int(seed)
        self._h = lib().orc_rng_new(C.c_uint64(self.initial_seed & 0xFFFFFFFFFFFFFFFF))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_rng_free(self._h)
            self._h = None

    def randint(self, high):
        return int(lib().orc_interval(self._h, C.c_uint64(int(high))))

    def random_double(self):
        return float(lib().orc_double(self._h))

    def binomial(self, n, p):
        return int(lib().orc_binomial(self._h, int(n), float(p)))

    def permutation(self, size):
        out = np.empty(int(size), dtype=np.int64)
        lib().orc_permutation(self._h, _p(out), int(size))
        return out

    def shuffle(self, x):
        """In-place Fisher-Yates of a 1-D int64 array (random_fast.pyx:87-125)."""
        assert x.dtype == np.int64 and x.ndim == 1 and x.flags["C_CONTIGUOUS"]
        lib().orc_shuffle_i64(self._h, _p(x), x.shape[0])

    def shuffle_with_trace(self, arrays):
        """Apply one random swap sequence to every array of the list (along axis 0) and
        return the permutation it realises (random_fast.pyx:127-144)."""
        n = len(arrays[0])
        swap = np.zeros(n, dtype=np.int64)
        trace = np.empty(n, dtype=np.int64)
        lib().orc_shuffle_trace(self._h, n, _p(swap), _p(trace))
        for x in arrays:
            x[:] = x[trace]          # replaying the swaps == gathering by the trace
        return trace


class Sampler(object):
    def __init__(self, range, rand_size, replacement, random_seed):
        self.range = int(range)
        self._h = lib().orc_sampler_new(self.range, int(bool(rand_size)), int(bool(replacement)),
                                        C.c_uint64(int(random_seed) & 0xFFFFFFFFFFFFFFFF))
        self._buf = np.empty(max(self.range, 1), dtype=np.int64)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_sampler_free(self._h)
            self._h = None

    def yield_subset(self, reduction):
        n = lib().orc_sampler_yield(self._h, float(reduction), _p(self._buf))
        return self._buf[:n].copy()


def batch_weight(count, batch_size, learning_rate, offset=0.0):
    return float(lib().orc_batch_weight(int(count), int(batch_size), float(learning_rate), float(offset)))


# --------------------------------------------------------------------------- #
# numerics
# --------------------------------------------------------------------------- #
def enet_norm(v, l1_ratio):
    v = np.ascontiguousarray(v)
    return float(getattr(lib(), "orc_enet_norm_" + _sfx(v))(_p(v), v.shape[0], l1_ratio))


def enet_scale(x, l1_ratio, radius=1.0):
    _chk(x)
    getattr(lib(), "orc_enet_scale_" + _sfx(x))(_p(x), x.shape[0], l1_ratio, radius)
    return x


def enet_projection(v, out, radius, l1_ratio):
    _chk(v), _chk(out, v.dtype)
    getattr(lib(), "orc_enet_projection_" + _sfx(v))(_p(v), _p(out), v.shape[0], radius, l1_ratio)
    return out


def cd_gram(w, alpha, beta, Q, q, y, max_iter, tol, positive):
    """One sample; returns the sweep count (dict_fact_fast.pyx:270-427)."""
    k = Q.shape[0]
    H = np.empty(k, dtype=Q.dtype)
    XtA = np.empty(k, dtype=Q.dtype)
    for a in (w, Q, q, y):
        _chk(a, Q.dtype)
    return int(getattr(lib(), "orc_cd_gram_" + _sfx(Q))(
        _p(w), alpha, beta, _p(Q), _p(q), _p(y), y.shape[0], k, _p(H), _p(XtA),
        int(max_iter), tol, int(bool(positive))))


def enet_regression_single_gram(G, Dx, X, code, indices, l1_ratio, alpha, positive,
                                tol, max_iter, return_sweeps=False):
    b, k = Dx.shape
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    for a in (G, Dx, X, code):
        _chk(a, G.dtype)
    sweeps = np.zeros(b, dtype=np.int32)
    info = getattr(lib(), "orc_regression_single_gram_" + _sfx(G))(
        _p(G), _p(Dx), _p(X), _p(code), _p(indices), b, k, X.shape[1],
        l1_ratio, alpha, int(bool(positive)), tol, int(max_iter), _p(sweeps))
    if return_sweeps:
        return code, sweeps, info
    return code


def enet_regression_multi_gram(G, Dx, X, code, indices, l1_ratio, alpha, positive,
                               tol, max_iter, return_sweeps=False):
    b, k = Dx.shape
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    for a in (G, Dx, X, code):
        _chk(a, G.dtype)
    sweeps = np.zeros(b, dtype=np.int32)
    info = getattr(lib(), "orc_regression_multi_gram_" + _sfx(G))(
        _p(G), _p(Dx), _p(X), _p(code), _p(indices), b, k, X.shape[1],
        l1_ratio, alpha, int(bool(positive)), tol, int(max_iter), _p(sweeps))
    if return_sweeps:
        return code, sweeps, info
    return code


def update_G_average(G_average, G, w_sample):
    for a in (G_average, G, w_sample):
        _chk(a, G.dtype)
    getattr(lib(), "orc_update_G_average_" + _sfx(G))(
        _p(G_average), _p(G), _p(w_sample), w_sample.shape[0], G.shape[0])
    return G_average


def update_dict_panel(D_sub, grad_sub, C_, comp_norm, order, l1_ratio, positive):
    """BCD over the gathered panels (dict_fact.py:663-694); all arrays modified in place."""
    k, s = D_sub.shape
    order = np.ascontiguousarray(order, dtype=np.int64)
    for a in (D_sub, grad_sub, C_, comp_norm):
        _chk(a, D_sub.dtype)
    getattr(lib(), "orc_update_dict_panel_" + _sfx(D_sub))(
        _p(D_sub), _p(grad_sub), _p(C_), _p(comp_norm), _p(order), k, s, l1_ratio,
        int(bool(positive)))
    return D_sub


def get_sub_slice(indices, sub):
    """modl/utils/__init__.py:4-27."""
    if indices is None:
        if isinstance(sub, slice):
            return np.arange(sub.start, sub.stop)
        return sub
    if isinstance(indices, slice):
        return np.arange(indices.start + sub.start, indices.start + sub.stop)
    return indices[sub]


# --------------------------------------------------------------------------- #
# estimator state machine
# --------------------------------------------------------------------------- #
def _check_random_state(seed):
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, (int, np.integer)):
        return np.random.RandomState(seed)
    return seed


class OracleDictFact(object):
    """NumPy restatement of DictFact (dict_fact.py:127-715); same kwargs, same
    fitted attributes.  Single-threaded."""

    def __init__(self, reduction=1, learning_rate=1, sample_learning_rate=0.76,
                 Dx_agg='masked', G_agg='masked', optimizer='variational', dict_init=None,
                 code_alpha=1, code_l1_ratio=1, comp_l1_ratio=0, step_size=1, tol=1e-2,
                 max_iter=100, code_pos=False, comp_pos=False, random_state=None,
                 n_epochs=1, n_components=10, batch_size=10, verbose=0, callback=None,
                 n_threads=1, rand_size=True, replacement=True):
        self.__dict__.update({k: v for k, v in locals().items() if k != 'self'})
        self.sweep_log_ = []          # oracle-only diagnostics: per-batch sweep counts
        self.subset_log_ = []         # oracle-only: subsets drawn
        self.order_log_ = []          # oracle-only: atom orders drawn

    # -- prepare: dict_fact.py:381-489 --
    def prepare(self, n_samples=None, n_features=None, dtype=None, X=None):
        k = self.n_components
        if X is not None:
            X = np.ascontiguousarray(X)
            if X.dtype not in (np.float32, np.float64):
                X = X.astype(np.float64)
            if dtype is None:
                dtype = X.dtype
            if n_samples is None:
                n_samples = X.shape[0]
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
        if self.optimizer == 'sgd':
            self.reduction, self.G_agg, self.Dx_agg = 1, 'full', 'full'
        if self.G_agg == 'average':
            self.G_average_ = np.zeros((n_samples, k, k), dtype=dtype)
        if self.Dx_agg == 'average':
            self.Dx_average_ = np.zeros((n_samples, k), dtype=dtype)
        self.C_ = np.zeros((k, k), dtype=dtype)
        self.B_ = np.zeros((k, n_features), dtype=dtype)
        self.gradient_ = np.zeros((k, n_features), dtype=dtype)
        self.random_state = _check_random_state(self.random_state)
        if X is None:
            self.components_ = np.empty((k, n_features), dtype=dtype)
            self.components_[:, :] = self.random_state.randn(k, n_features)
        else:
            self.components_ = np.array(X[:k], dtype=dtype, order='C', copy=True)
        if self.comp_pos:
            neg = self.components_ <= 0
            self.components_[neg] = -self.components_[neg]
        for i in range(k):
            enet_scale(self.components_[i], self.comp_l1_ratio, 1.0)
        self.code_ = np.ones((n_samples, k), dtype=dtype)
        self.labels_ = np.arange(n_samples)
        self.comp_norm_ = np.zeros(k, dtype=dtype)
        if self.G_agg == 'full':
            self.G_ = self.components_.dot(self.components_.T)
        self.n_iter_ = 0
        self.sample_n_iter_ = np.zeros(n_samples, dtype=np.int64)
        seed = self.random_state.randint(MAX_INT)
        self.feature_sampler_ = Sampler(n_features, self.rand_size, self.replacement, seed)
        if self.verbose:
            self.verbose_iter_ = np.linspace(0, n_samples * self.n_epochs, self.verbose).tolist()
        self.time_ = 0
        return self

    # -- fit / partial_fit: dict_fact.py:286-337 --
    def fit(self, X):
        X = np.ascontiguousarray(X)
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        init = X if self.dict_init is None else np.asarray(self.dict_init, dtype=X.dtype)
        self.prepare(n_samples=X.shape[0], X=init)
        for _ in range(self.n_epochs):
            self.partial_fit(X)
            perm = self.shuffle()
            X = X[perm]
        return self

    def partial_fit(self, X, sample_indices=None):
        X = np.ascontiguousarray(X)
        n = X.shape[0]
        for start in range(0, n, self.batch_size):
            sl = slice(start, min(start + self.batch_size, n))
            self._single_batch_fit(X[sl], get_sub_slice(sample_indices, sl))
        return self

    # -- set_params: dict_fact.py:339-357 (only a switch to G_agg='full' is honoured for G_agg) --
    def set_params(self, **params):
        G_agg = params.pop('G_agg', None)
        if G_agg == 'full' and self.G_agg != 'full':
            if hasattr(self, 'components_'):
                self.G_ = self.components_.dot(self.components_.T)
            self.G_agg = 'full'
        for key, value in params.items():
            if key not in self.__dict__:
                raise ValueError('Invalid parameter %s' % key)
            setattr(self, key, value)

    def shuffle(self):
        seed = self.random_state.randint(MAX_INT)
        rs = RandomState(seed)
        arrays = [self.code_]
        if self.G_agg == 'average':
            arrays.append(self.G_average_)
        if self.Dx_agg == 'average':
            arrays.append(self.Dx_average_)
        perm = rs.shuffle_with_trace(arrays)
        self.labels_ = self.labels_[perm]
        return perm

    # -- the hot path: dict_fact.py:495-526 --
    def _single_batch_fit(self, X, idx):
        if self.verbose and self.verbose_iter_ and self.n_iter_ >= self.verbose_iter_[0]:
            print('Iteration %i' % self.n_iter_)
            self.verbose_iter_ = self.verbose_iter_[1:]
            if self.callback is not None:
                self.callback(self)
        t0 = time.perf_counter()
        subset = self.feature_sampler_.yield_subset(self.reduction)
        self.subset_log_.append(subset)
        b = X.shape[0]
        self.n_iter_ += b
        self.sample_n_iter_[idx] += 1
        w_sample = np.power(self.sample_n_iter_[idx], -self.sample_learning_rate).astype(
            self.components_.dtype)
        w = batch_weight(self.n_iter_, b, self.learning_rate, 0)
        self._compute_code(X, idx, w_sample, subset)
        code = self.code_[idx]
        self._update_C(code, w)
        self._update_B(X, code, w)
        self.gradient_[:, subset] = self.B_[:, subset]
        self._update_dict(subset, w)
        self.time_ += time.perf_counter() - t0

    # dict_fact.py:568-575
    def _update_C(self, code, w):
        b = code.shape[0]
        if self.optimizer == 'variational':
            self.C_ *= 1 - w
            self.C_ += w * code.T.dot(code) / b
        else:
            self.C_ = code.T.dot(code) / b

    # dict_fact.py:559-566
    def _update_B(self, X, code, w):
        b = X.shape[0]
        if self.optimizer == 'variational':
            self.B_ *= 1 - w
            self.B_ += w * code.T.dot(X) / b
        else:
            self.B_ = code.T.dot(X) / b

    # dict_fact.py:577-648
    def _compute_code(self, X, idx, w_sample, subset):
        r = self.reduction
        idx = np.asarray(idx)
        if self.Dx_agg != 'full' or self.G_agg != 'full':
            D_sub = self.components_[:, subset]
        if self.Dx_agg == 'full':
            Dx = X.dot(self.components_.T)
        else:
            Dx = X[:, subset].dot(D_sub.T) * r
            if self.Dx_agg == 'average':
                self.Dx_average_[idx] *= 1 - w_sample[:, None]
                self.Dx_average_[idx] += Dx * w_sample[:, None]
                Dx = self.Dx_average_[idx]
        Dx = np.ascontiguousarray(Dx, dtype=self.components_.dtype)
        if self.G_agg != 'full':
            G = np.ascontiguousarray(D_sub.dot(D_sub.T) * r, dtype=self.components_.dtype)
            if self.G_agg == 'average':
                G_av = np.array(self.G_average_[idx], copy=True)
                update_G_average(G_av, G, w_sample)
                self.G_average_[idx] = G_av
        else:
            G = self.G_
        X = np.ascontiguousarray(X)
        if self.G_agg == 'average':
            _, sw, _ = enet_regression_multi_gram(G_av, Dx, X, self.code_, idx, self.code_l1_ratio,
                                                  self.code_alpha, self.code_pos, self.tol,
                                                  self.max_iter, return_sweeps=True)
        else:
            _, sw, _ = enet_regression_single_gram(G, Dx, X, self.code_, idx, self.code_l1_ratio,
                                                   self.code_alpha, self.code_pos, self.tol,
                                                   self.max_iter, return_sweeps=True)
        self.sweep_log_.append(sw)

    # dict_fact.py:650-715
    def _update_dict(self, subset, w):
        k, p = self.components_.shape
        s = subset.shape[0]
        D_sub = np.ascontiguousarray(self.components_[:, subset])
        grad = np.ascontiguousarray(self.gradient_[:, subset])
        if self.G_agg == 'full' and s < p / 2.:
            self.G_ -= D_sub.dot(D_sub.T)
        order = self.random_state.permutation(k)
        self.order_log_.append(np.asarray(order, dtype=np.int64))
        if self.optimizer == 'variational':
            update_dict_panel(D_sub, grad, self.C_, self.comp_norm_, order,
                              self.comp_l1_ratio, self.comp_pos)
        else:
            grad -= self.C_.dot(D_sub)
            for a in order:
                self.comp_norm_[a] += enet_norm(D_sub[a], self.comp_l1_ratio)
            D_sub += w * self.step_size * grad
            tmp = np.zeros(s, dtype=D_sub.dtype)
            for a in range(k):
                enet_projection(D_sub[a], tmp, self.comp_norm_[a], self.comp_l1_ratio)
                D_sub[a] = tmp
                self.comp_norm_[a] -= enet_norm(D_sub[a], self.comp_l1_ratio)
        self.components_[:, subset] = D_sub
        if self.G_agg == 'full':
            if s < p / 2.:
                self.G_ += D_sub.dot(D_sub.T)
            else:
                self.G_[:] = self.components_.dot(self.components_.T)

    # -- inference: dict_fact.py:47-114 --
    def transform(self, X):
        dtype = self.components_.dtype
        X = np.ascontiguousarray(X, dtype=dtype)
        n = X.shape[0]
        if getattr(self, 'G_agg', None) != 'full':
            G = self.components_.dot(self.components_.T)
        else:
            G = self.G_
        G = np.ascontiguousarray(G)
        Dx = np.ascontiguousarray(X.dot(self.components_.T))
        code = np.ones((n, self.n_components), dtype=dtype)
        enet_regression_single_gram(G, Dx, X, code, np.arange(n), self.code_l1_ratio,
                                    self.code_alpha, self.code_pos, self.tol, self.max_iter)
        return code

    def score(self, X):
        code = self.transform(X)
        loss = np.sum((X - code.dot(self.components_)) ** 2) / 2
        n1 = np.sum(np.abs(code))
        n2 = np.sum(code ** 2)
        regul = self.code_alpha * (n1 * self.code_l1_ratio + (1 - self.code_l1_ratio) * n2 / 2)
        return (loss + regul) / X.shape[0]


# ---------------------------------------------------------------------------------------
# Matrix completion: RecsysDictFact (modl/decomposition/recsys.py:16-265,
# modl/decomposition/recsys_fast.pyx:10-37).  A different algorithm from the estimator above:
# per-row ridge solves over the row's observed columns, per-row scaled rank-1 B_ updates with
# per-feature counters, dictionary update with a plain L2-ball projection on the union of the
# batch's observed columns.  Pinned by tests/golden/recsys.npz (tests/test_oracle.py).
# ---------------------------------------------------------------------------------------
class OracleRecsysDictFact(object):
    def __init__(self, alpha=1.0, beta=.0, n_components=30, learning_rate=1., batch_size=1,
                 n_epochs=1, random_state=None, detrend=False, crop=None):
        self.__dict__.update({k: v for k, v in locals().items() if k != 'self'})

    @staticmethod
    def _row(X, i):
        lo, hi = X.indptr[i], X.indptr[i + 1]
        return X.indices[lo:hi], X.data[lo:hi]

    def _solve_row(self, X, i):
        """recsys.py:169-178 (and :256-265): ridge code of row i over its observed columns."""
        cols, vals = self._row(X, i)
        D_obs = self.components_[:, cols]
        gram = D_obs.dot(D_obs.T)
        gram[np.diag_indices_from(gram)] += self.alpha / (X.shape[1] / cols.shape[0])
        self.code_[i] = np.linalg.solve(gram, D_obs.dot(vals))

    def fit(self, X):
        import scipy.sparse as sp
        X = sp.csr_matrix(X, copy=True)
        n, p = X.shape
        k = self.n_components
        dtype = X.dtype
        rng = _check_random_state(self.random_state)
        if self.detrend:                                           # recsys.py:105-111
            self.row_mean_, self.col_mean_ = recsys_biases(X, self.beta)
            X.data -= np.repeat(self.row_mean_, np.diff(X.indptr))
            X.data -= self.col_mean_[X.indices]
        D = rng.randn(k, p).astype(dtype)                          # :113-116
        D /= np.sqrt(np.sum(D ** 2, axis=1))[:, None]
        self.components_ = D
        self.code_ = np.zeros((n, k), dtype=dtype)
        for i in range(n):                                         # _refit, :254-265
            self._solve_row(X, i)
        self.feature_n_iter_ = np.zeros(p, dtype=np.int64)
        bs = self.batch_size
        if bs is None:                                             # :123-127
            bs = int(np.ceil(1. / (X.nnz / n / p)))
        self.comp_norm_ = np.zeros(k, dtype=dtype)
        self.C_ = np.zeros((k, k), dtype=dtype)
        self.B_ = np.zeros((k, p), dtype=dtype)
        self.n_iter_ = 0
        for _ in range(self.n_epochs):                             # :139-143
            perm = rng.permutation(n)
            for lo in range(0, n, bs):
                self._step(X, perm[lo:lo + bs], rng)
        for i in range(n):
            self._solve_row(X, i)
        return self

    def _step(self, X, rows, rng):
        """One minibatch, recsys.py:151-213."""
        b = rows.shape[0]
        self.n_iter_ += b
        w = batch_weight(self.n_iter_, b, self.learning_rate, 0)
        for i in rows:                                             # _single_sample_update, :164-185
            cols, vals = self._row(X, i)
            self.feature_n_iter_[cols] += 1
            self._solve_row(X, i)
            w_B = np.minimum(1, w * self.n_iter_ / self.feature_n_iter_[cols])
            self.B_[:, cols] *= 1 - w_B
            self.B_[:, cols] += np.outer(self.code_[i], vals * w_B)
        batch_code = self.code_[rows]
        self.C_ *= 1 - w
        self.C_ += w / b * batch_code.T.dot(batch_code)
        union = np.unique(np.concatenate([self._row(X, i)[0] for i in rows]))
        # _update_dict, :187-213: sequential atoms, L2 ball of squared radius comp_norm_[a]
        D_sub = self.components_[:, union]
        resid = self.B_[:, union] - self.C_.dot(D_sub)
        order = rng.permutation(self.n_components)
        self.comp_norm_ += np.sum(D_sub ** 2, axis=1)
        for a in order:
            resid += np.outer(self.C_[a], D_sub[a])
            if self.C_[a, a] > 1e-20:
                D_sub[a] = resid[a] / self.C_[a, a]
            length, limit = np.sqrt(np.sum(D_sub[a] ** 2)), np.sqrt(self.comp_norm_[a])
            if length > limit:
                D_sub[a] /= length / limit
            resid -= np.outer(self.C_[a], D_sub[a])
        self.comp_norm_ -= np.sum(D_sub ** 2, axis=1)
        self.components_[:, union] = D_sub

    def predict_data(self, X):
        """Values at the stored entries of X (recsys.py:215-244 with recsys_fast.pyx:10-37), as a data array."""
        import scipy.sparse as sp
        X = sp.csr_matrix(X)
        out = np.zeros(X.nnz, dtype=np.float64)
        for u in range(X.shape[0]):
            for e in range(X.indptr[u], X.indptr[u + 1]):
                dot = 0.
                for a in range(self.n_components):
                    dot += float(self.code_[u, a]) * float(self.components_[a, X.indices[e]])
                out[e] = dot
        if self.detrend:
            out += np.repeat(self.row_mean_, np.diff(X.indptr))
            out += self.col_mean_[X.indices]
        if self.crop is not None:
            np.clip(out, self.crop[0], self.crop[1], out=out)
        return out


def recsys_biases(X, beta=0):
    """Row / column offsets by two rounds of alternating centring (recsys.py:268-303)."""
    import scipy.sparse as sp
    X = sp.csr_matrix(X, copy=True)
    lens = np.diff(X.indptr)
    n_row = np.maximum(X.getnnz(axis=1), 1)
    n_col = np.maximum(X.getnnz(axis=0), 1)
    row_off, col_off = np.zeros(X.shape[0]), np.zeros(X.shape[1])
    mean = np.mean(X.data)
    for _ in range(2):
        r = (np.asarray(X.sum(axis=1))[:, 0] + mean * beta) / (n_row + beta)
        X.data -= np.repeat(r, lens)
        c = np.asarray(X.sum(axis=0))[0] / (n_col + beta)
        X.data -= c[X.indices]
        row_off += r
        col_off += c
    return row_off, col_off
