"""RecsysDictFact against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py::gold_recsys) and against the reference's own test properties
[ref: modl/decomposition/tests/test_recsys.py].

On CPU: `compute_biases` / `rmse` (host code), and the estimator's HOST logic -- random stream, batch cutting,
the (column, batch position) ordering of a minibatch's entries, counters, detrending, cropping -- with a NumPy
restatement of the device kernels' contracts standing in for them (tests may do that; the product never does).
On the GPU: the same cases through the CUDA kernels, plus kernel-level checks at larger shapes."""
import json
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rel_err

torch = pytest.importorskip("torch")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def gold(golden):
    return golden("recsys.npz")


def _data(name, dtype):
    """Same seeded construction as make_golden.recsys_data."""
    rng = np.random.RandomState(0)
    if name == "dense":
        X = sp.csr_matrix(np.dot(rng.rand(50, 3), rng.rand(3, 20)))
        return X.astype(dtype), X.astype(dtype)
    if name == "missing":
        X = np.dot(rng.rand(100, 4), rng.rand(4, 20))
        keep = rng.rand(100, 20) < 0.85
        keep[np.arange(100), rng.randint(20, size=100)] = True
        return sp.csr_matrix(X * keep).astype(dtype), sp.csr_matrix(X * ~keep).astype(dtype)
    U, V = rng.rand(120, 4), rng.rand(4, 60)
    R = np.clip(np.round(1 + 4 * U.dot(V) / 2.), 1, 5)
    seen = rng.rand(120, 60) < 0.25
    seen[np.arange(120), rng.randint(60, size=120)] = True
    test = ~seen & (rng.rand(120, 60) < 0.1)
    return sp.csr_matrix(R * seen).astype(dtype), sp.csr_matrix(R * test).astype(dtype)


class NumpyKernels(object):
    """What each device kernel is specified to compute (include/modl_b200.h, "Recsys" block), in NumPy on CPU
    tensors.  Used ONLY to exercise the estimator's host logic without a GPU."""

    def __init__(self, device, tdt):
        pass

    def gram_dx(self, Dt, indptr, indices, data, rows, row0, b, p, alpha, G, Dx):
        Dt, indptr, indices, data = Dt.numpy(), indptr.numpy(), indices.numpy(), data.numpy()
        k = Dt.shape[1]
        for ii in range(b):
            r = int(rows[ii]) if rows is not None else row0 + ii
            cols = indices[indptr[r]:indptr[r + 1]]
            Ds = Dt[cols].T
            g = Ds.dot(Ds.T)
            g.flat[::k + 1] += alpha / (p / cols.shape[0])
            G.numpy()[ii] = g
            Dx.numpy()[ii] = Ds.dot(data[indptr[r]:indptr[r + 1]])

    def check_solves(self):
        pass                                # np.linalg.solve below raises by itself

    def solve(self, G, Dx, code, rows, b):
        out = code.numpy()
        for ii in range(b):
            c = np.linalg.solve(G.numpy()[ii].astype(np.float64), Dx.numpy()[ii].astype(np.float64))
            out[int(rows[ii]) if rows is not None else ii] = c

    def update_B(self, B, code, subset, col_ptr, entry_row, entry_val, feature_n_iter, w, n_iter):
        B, code, cnt = B.numpy(), code.numpy(), feature_n_iter.numpy()
        for j, col in enumerate(subset.numpy()):
            for e in range(int(col_ptr[j]), int(col_ptr[j + 1])):
                cnt[col] += 1
                wB = min(1., w * n_iter / cnt[col])
                scaled = (B[:, col].astype(np.float64) * (1 - wB)).astype(B.dtype)
                B[:, col] = scaled.astype(np.float64) + code[int(entry_row[e])].astype(np.float64) * (float(entry_val[e]) * wB)

    def update_C(self, C, code, rows, w):
        cb = code.numpy()[rows.numpy()]
        C.numpy()[:] = (1 - w) * C.numpy() + w / cb.shape[0] * cb.T.dot(cb)

    def update_dict(self, D, B, C, comp_norm, subset, order):
        D, B, C, norm, subset = D.numpy(), B.numpy(), C.numpy(), comp_norm.numpy(), subset.numpy()
        Ds = D[:, subset]
        grad = B[:, subset] - C.dot(Ds)
        for kk in order:
            norm[kk] += np.sum(Ds[kk] ** 2)
            grad += np.outer(C[kk], Ds[kk])
            if C[kk, kk] > 1e-20:
                Ds[kk] = grad[kk] / C[kk, kk]
            ssq = np.sum(Ds[kk] ** 2)
            if ssq > norm[kk]:
                Ds[kk] /= np.sqrt(ssq / norm[kk]) if norm[kk] > 0 else np.inf
            norm[kk] -= np.sum(Ds[kk] ** 2)
            grad -= np.outer(C[kk], Ds[kk])
        D[:, subset] = Ds

    def sync_transposed(self, D, Dt, subset):
        if subset is None:
            Dt.numpy()[:] = D.numpy().T
        else:
            Dt.numpy()[subset.numpy()] = D.numpy()[:, subset.numpy()].T

    def predict(self, code, Dt, indptr, indices, out):
        code, Dt, indptr, indices, o = code.numpy(), Dt.numpy(), indptr.numpy(), indices.numpy(), out.numpy()
        for u in range(indptr.shape[0] - 1):
            for e in range(indptr[u], indptr[u + 1]):
                dot = 0.
                for a in range(code.shape[1]):
                    dot += float(code[u, a]) * float(Dt[indices[e], a])
                o[e] = dot


def _check_case(recsys, gold, ci, case, device, tol64, tol32, bookkeeping='batch'):
    X, X_te = _data(case["data"], case["dtype"])
    kw = dict(case["kw"])
    if "crop" in kw:
        kw["crop"] = tuple(kw["crop"])
    est = recsys.RecsysDictFact(random_state=0, device=device, bookkeeping=bookkeeping, **kw).fit(X)
    tag = "fit_%d_" % ci
    tol = tol64 if case["dtype"] == "float64" else tol32
    assert est.n_iter_ == int(gold[tag + "n_iter_"])
    np.testing.assert_array_equal(est.feature_n_iter_, gold[tag + "feature_n_iter_"])        # integer: exact
    np.testing.assert_array_equal(est.feature_freq_, gold[tag + "feature_freq_"])
    errs = {}
    for name in ("components_", "code_", "C_", "B_"):
        got = getattr(est, name)
        assert got.dtype == gold[tag + name].dtype and got.shape == gold[tag + name].shape
        errs[name] = rel_err(got, gold[tag + name])
    errs["comp_norm_"] = float(np.abs(est.comp_norm_ - gold[tag + "comp_norm_"]).max())
    if kw.get("detrend"):
        np.testing.assert_allclose(est.row_mean_, gold[tag + "row_mean_"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(est.col_mean_, gold[tag + "col_mean_"], rtol=1e-13, atol=1e-15)
    if case["dtype"] == "float64":
        errs["pred_train"] = rel_err(est.predict(X).data, gold[tag + "pred_train"])
        errs["pred_test"] = rel_err(est.predict(X_te).data, gold[tag + "pred_test"])
        errs["score"] = abs(est.score(X_te) - float(gold[tag + "score_test"])) / float(gold[tag + "score_test"])
    print("case %d %s %s: %s" % (ci, case["data"], case["dtype"], ", ".join("%s %.2g" % kv for kv in errs.items())))
    assert all(v < tol for v in errs.values()), (ci, case, errs)
    return est, X, X_te


def test_compute_biases_and_rmse(gold):
    from modl_b200.recsys import compute_biases, rmse
    X, X_te = _data("ratings", "float64")
    before = X.data.copy()
    row, col = compute_biases(X, beta=3., inplace=False)
    np.testing.assert_array_equal(X.data, before)                                   # not in place
    np.testing.assert_allclose(row, gold["bias_row"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(col, gold["bias_col"], rtol=1e-13, atol=1e-15)
    Xc = X.copy()
    compute_biases(Xc, inplace=True)
    assert abs(Xc.data.mean()) < 0.05 and np.abs(Xc.data).max() < np.abs(X.data).max()
    Y = X.copy()
    Y.data = Y.data + 0.5
    assert abs(rmse(X, Y) - 0.5) < 1e-15


def test_empty_row_is_rejected():
    from modl_b200 import recsys
    X = sp.csr_matrix(np.array([[1., 0., 2.], [0., 0., 0.], [0., 3., 0.]]))
    with pytest.raises(ZeroDivisionError):
        recsys.RecsysDictFact(n_components=2, random_state=0, device="cpu").fit(X)


def test_no_cpu_fallback():
    from modl_b200 import recsys
    X, _ = _data("dense", "float64")
    with pytest.raises(RuntimeError):
        recsys.RecsysDictFact(n_components=2, random_state=0, device="cpu").fit(X)


@pytest.mark.parametrize("bookkeeping", ["batch", "epoch"])
def test_host_logic_cpu(gold, monkeypatch, bookkeeping):
    """Whole fits with the NumPy stand-in for the kernels reproduce the reference: the host side (random
    stream, batches, entry ordering -- per minibatch or for a whole epoch at once --, counters, detrend / crop)
    is right."""
    from modl_b200 import recsys
    monkeypatch.setattr(recsys, "_kernels_for", lambda device, tdt: NumpyKernels(device, tdt))
    for ci, case in enumerate(json.loads(str(gold["cases"]))):
        _check_case(recsys, gold, ci, case, "cpu", 1e-9, 2e-3, bookkeeping)


@pytest.mark.parametrize("max_entries", [1 << 26, 700, 1])
def test_epoch_entries_equal_batch_entries_cpu(max_entries):
    """The epoch-level bookkeeping yields, minibatch by minibatch, exactly what the per-minibatch one builds --
    also when the epoch is cut into several chunks (small `max_entries`; 1 = one minibatch per chunk)."""
    from modl_b200 import recsys
    X, _ = _data("ratings", "float64")
    est = recsys.RecsysDictFact(n_components=2, random_state=0, device="cpu")
    est.__dict__["_device"] = torch.device("cpu")
    Xd = recsys._DeviceCSR(X, torch.float64, torch.device("cpu"))
    perm = np.random.RandomState(3).permutation(X.shape[0])
    bs = 7                                                              # 120 rows: 17 minibatches + a ragged one of 1
    got = list(est._epoch_entries(Xd, perm, bs, max_entries=max_entries))
    assert len(got) == -(-X.shape[0] // bs)
    for j, (rows, subset, col_ptr, entry_row, entry_val) in enumerate(got):
        w_rows, w_subset, w_ptr, w_erow, w_eval = est._batch_entries(Xd, perm[j * bs:(j + 1) * bs])
        np.testing.assert_array_equal(rows.numpy(), w_rows.numpy())
        np.testing.assert_array_equal(subset.numpy(), w_subset.numpy())
        lo, hi = int(col_ptr[0]), int(col_ptr[-1])
        np.testing.assert_array_equal((col_ptr - lo).numpy(), w_ptr.numpy())
        np.testing.assert_array_equal(entry_row.numpy()[lo:hi], w_erow.numpy())
        np.testing.assert_array_equal(entry_val.numpy()[lo:hi], w_eval.numpy())


def test_batch_entries_ordering_cpu(monkeypatch):
    """Entries of a minibatch come out grouped by column, and inside a column in batch order."""
    from modl_b200 import recsys
    monkeypatch.setattr(recsys, "_kernels_for", lambda device, tdt: NumpyKernels(device, tdt))
    X, _ = _data("ratings", "float64")
    est = recsys.RecsysDictFact(n_components=2, random_state=0, device="cpu")
    est.__dict__["_device"] = torch.device("cpu")
    Xd = recsys._DeviceCSR(X, torch.float64, torch.device("cpu"))
    batch = np.array([17, 3, 99, 42, 8])
    rows, subset, col_ptr, entry_row, entry_val = est._batch_entries(Xd, batch)
    np.testing.assert_array_equal(rows.numpy(), batch)
    want_subset = np.unique(np.concatenate([X.indices[X.indptr[i]:X.indptr[i + 1]] for i in batch]))
    np.testing.assert_array_equal(subset.numpy(), want_subset)                     # [ref: recsys.py:159-161]
    where = {r: i for i, r in enumerate(batch)}
    for j, col in enumerate(subset.numpy()):
        seg = slice(int(col_ptr[j]), int(col_ptr[j + 1]))
        seg_rows = entry_row.numpy()[seg]
        assert list(seg_rows) == sorted(seg_rows, key=lambda r: where[r]) and len(set(seg_rows)) == len(seg_rows)
        for r, v in zip(seg_rows, entry_val.numpy()[seg]):
            assert X[r, col] == v
    assert int(col_ptr[-1]) == sum(X.indptr[i + 1] - X.indptr[i] for i in batch)


# ------------------------------------------------------------------------------------------------ GPU
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


@pytest.mark.gpu
def test_recsys_kernels_against_numpy():
    """Each recsys kernel against the NumPy statement of its contract, at shapes that exercise the tiling:
    k not a multiple of 4, k > 64 (four blocks per thread), rows longer than one tile, float32 and float64."""
    dev = _gpu()
    from modl_b200 import recsys
    rng = np.random.RandomState(5)
    for dt, k, n, p, dens in ((np.float64, 7, 40, 300, 0.2), (np.float32, 50, 64, 1000, 0.1),
                              (np.float64, 96, 24, 500, 0.3), (np.float32, 128, 16, 700, 0.05)):
        tdt = torch.float32 if dt == np.float32 else torch.float64
        tol = 1e-12 if dt == np.float64 else 2e-5
        M = (rng.rand(n, p) < dens) * (1 + rng.rand(n, p))
        M[np.arange(n), rng.randint(p, size=n)] = 1.5
        X = sp.csr_matrix(M).astype(dt)
        D = rng.randn(k, p).astype(dt)
        kd, kn = recsys._DeviceKernels(dev, tdt), NumpyKernels(None, None)
        Xd, Xh = recsys._DeviceCSR(X, tdt, dev), recsys._DeviceCSR(X, tdt, torch.device("cpu"))
        Dd, Dh = torch.from_numpy(D).to(dev), torch.from_numpy(D.copy())
        # transposed copy: full, then a subset refresh
        Dtd, Dth = torch.zeros((p, k), dtype=tdt, device=dev), torch.zeros((p, k), dtype=tdt)
        kd.sync_transposed(Dd, Dtd, None)
        np.testing.assert_array_equal(Dtd.cpu().numpy(), D.T)
        sub = np.sort(rng.permutation(p)[:p // 3]).astype(np.int64)
        Dd[:, torch.from_numpy(sub).to(dev)] *= 2
        kd.sync_transposed(Dd, Dtd, torch.from_numpy(sub).to(dev))
        np.testing.assert_array_equal(Dtd.cpu().numpy(), Dd.cpu().numpy().T)
        Dh.copy_(Dd.cpu())
        kn.sync_transposed(Dh, Dth, None)
        # Gram / correlation, by row list and by range
        rows = rng.permutation(n)[:n // 2].astype(np.int64)
        for rws, row0, b in ((rows, 0, rows.shape[0]), (None, 3, n - 3)):
            Gd, Dxd = torch.zeros((b, k, k), dtype=tdt, device=dev), torch.zeros((b, k), dtype=tdt, device=dev)
            Gh, Dxh = torch.zeros((b, k, k), dtype=tdt), torch.zeros((b, k), dtype=tdt)
            kd.gram_dx(Dtd, Xd.indptr, Xd.indices, Xd.data, None if rws is None else torch.from_numpy(rws).to(dev),
                       row0, b, p, 0.7, Gd, Dxd)
            kn.gram_dx(Dth, Xh.indptr, Xh.indices, Xh.data, None if rws is None else torch.from_numpy(rws), row0, b, p,
                       0.7, Gh, Dxh)
            assert rel_err(Gd.cpu().numpy(), Gh.numpy()) < tol and rel_err(Dxd.cpu().numpy(), Dxh.numpy()) < tol
        # solve through the shared ridge kernel
        coded = torch.zeros((n, k), dtype=tdt, device=dev)
        Gd2, Dxd2 = torch.zeros((n, k, k), dtype=tdt, device=dev), torch.zeros((n, k), dtype=tdt, device=dev)
        kd.gram_dx(Dtd, Xd.indptr, Xd.indices, Xd.data, None, 0, n, p, 50., Gd2, Dxd2)
        want = np.stack([np.linalg.solve(g.astype(np.float64), d.astype(np.float64))
                         for g, d in zip(Gd2.cpu().numpy(), Dxd2.cpu().numpy())])
        perm = torch.from_numpy(rng.permutation(n).astype(np.int64)).to(dev)
        kd.solve(Gd2, Dxd2, coded, perm, n)
        assert rel_err(coded.cpu().numpy()[perm.cpu().numpy()], want) < (1e-10 if dt == np.float64 else 1e-4)
        # B_ recurrence + counters, C_
        est = recsys.RecsysDictFact(device=dev)
        est.__dict__["_device"] = dev
        batch = rng.permutation(n)[:n // 2]
        rws, subset, col_ptr, entry_row, entry_val = est._batch_entries(Xd, batch)
        code = rng.randn(n, k).astype(dt)
        B0 = rng.randn(k, p).astype(dt)
        cnt0 = rng.randint(0, 4, size=p).astype(np.int64)
        Bd, cntd, cd = torch.from_numpy(B0.copy()).to(dev), torch.from_numpy(cnt0.copy()).to(dev), torch.from_numpy(code).to(dev)
        Bh, cnth = torch.from_numpy(B0.copy()), torch.from_numpy(cnt0.copy())
        kd.update_B(Bd, cd, subset, col_ptr, entry_row, entry_val, cntd, 0.31, 57)
        kn.update_B(Bh, torch.from_numpy(code), subset.cpu(), col_ptr.cpu(), entry_row.cpu(), entry_val.cpu(), cnth, 0.31, 57)
        np.testing.assert_array_equal(cntd.cpu().numpy(), cnth.numpy())
        np.testing.assert_array_equal(Bd.cpu().numpy(), Bh.numpy())                 # same roundings as NumPy: exact
        C0 = rng.randn(k, k).astype(dt)
        Cd, Ch = torch.from_numpy(C0.copy()).to(dev), torch.from_numpy(C0.copy())
        kd.update_C(Cd, cd, rws, 0.31)
        kn.update_C(Ch, torch.from_numpy(code), rws.cpu(), 0.31)
        assert rel_err(Cd.cpu().numpy(), Ch.numpy()) < tol
        # predictions
        outd, outh = torch.zeros(X.nnz, dtype=torch.float64, device=dev), torch.zeros(X.nnz, dtype=torch.float64)
        kd.predict(cd, Dtd, Xd.indptr, Xd.indices, outd)
        kn.predict(torch.from_numpy(code), Dth, Xh.indptr, Xh.indices, outh)
        np.testing.assert_array_equal(outd.cpu().numpy(), outh.numpy())             # unfused float64 chain: exact
        torch.cuda.synchronize()


@pytest.mark.gpu
def test_recsys_fits_match_the_reference(gold):
    """Whole fits on the device against the unmodified reference (its own test problems, minibatches, several
    epochs, detrending, cropping, automatic batch size, float32)."""
    _gpu()
    from modl_b200 import recsys
    for ci, case in enumerate(json.loads(str(gold["cases"]))):
        est, X, X_te = _check_case(recsys, gold, ci, case, None, 1e-8, 5e-3)
        if ci == 0:
            # the reference's own assertions [ref: tests/test_recsys.py:13-35]
            Y = np.dot(est.code_, est.components_)
            np.testing.assert_array_almost_equal(Y, est.predict(X).toarray())
            assert abs(np.sqrt(np.mean((X.toarray() - Y) ** 2)) - est.score(X)) < 1e-7


@pytest.mark.gpu
def test_recsys_fits_epoch_bookkeeping(gold):
    _gpu()
    from modl_b200 import recsys
    for ci, case in enumerate(json.loads(str(gold["cases"]))):
        _check_case(recsys, gold, ci, case, None, 1e-8, 5e-3, bookkeeping='epoch')


@pytest.mark.gpu
def test_recsys_config5_width_against_the_oracle(oracle):
    """BASELINE configs[4] width (10^5 items at Movielens-10M density, k = 50, alpha = 1, batch 10 as in the
    reference's examples/predict_recsys.py:41-45) on a bounded number of rows: device fit against the CPU oracle
    (itself pinned to the reference, tests/test_oracle.py::test_golden_recsys); rows hold ~1 340 ratings, i.e.
    84 tiles of the Gram kernel, and a minibatch touches ~12 000 distinct columns."""
    _gpu()
    from modl_b200.recsys import RecsysDictFact
    rng = np.random.RandomState(7)
    n, p, k = 48, 100000, 50
    seen = rng.rand(n, p) < 0.0134
    U, V = rng.rand(n, 8), rng.rand(8, p)
    X = sp.csr_matrix(np.where(seen, np.clip(np.round(1 + U.dot(V) + 0.3 * rng.randn(n, p)), 1, 5), 0.))
    kw = dict(n_components=k, alpha=1., batch_size=10, n_epochs=1, random_state=0)
    est = RecsysDictFact(**kw).fit(X)
    orc = oracle.OracleRecsysDictFact(**kw).fit(X)
    np.testing.assert_array_equal(est.feature_n_iter_, orc.feature_n_iter_)
    for name in ("components_", "code_", "C_", "B_"):
        err = rel_err(getattr(est, name), getattr(orc, name))
        assert err < 1e-9, (name, err)
    assert rel_err(est.predict(X).data, orc.predict_data(X)) < 1e-9


@pytest.mark.gpu
def test_recsys_completion_beats_centering():
    """[ref: tests/test_recsys.py:65-92] held-out error below that of the bias-only predictor."""
    _gpu()
    from modl_b200.recsys import RecsysDictFact, compute_biases
    X, X_te = _data("missing", "float64")
    mf = RecsysDictFact(n_components=4, n_epochs=1, alpha=1, random_state=0, detrend=True).fit(X)
    pred = mf.predict(X_te)
    err = np.sqrt(np.sum((X_te.data - pred.data) ** 2) / X_te.data.shape[0])
    centred = X_te.copy()
    compute_biases(centred, inplace=True)
    err_c = np.sqrt(np.sum((X_te.data - centred.data) ** 2) / X_te.data.shape[0])
    assert err < err_c
