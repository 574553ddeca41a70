"""Pins the CPU oracle (oracle/) against the reference's known-answer tests and the
golden vectors generated from the unmodified reference (tests/golden/make_golden.py).
CPU only."""
import numpy as np
import pytest

from conftest import rel_err

GOLD_FIT_CASES = None


def _fit_cases():
    import importlib.util
    import os
    p = os.path.join(os.path.dirname(__file__), "golden", "make_golden.py")
    src = open(p).read()
    # FIT_CASES / fit_data are pure data: extract without importing the reference
    start = src.index("FIT_CASES = (")
    end = src.index("def gold_fit")
    ns = {"np": np}
    exec(src[start:end], ns)
    return ns["FIT_CASES"], ns["fit_data"]


# ---- reference KATs: modl/utils/randomkit/tests/test_random.py:10-38 ----
def test_kat_random(oracle):
    rs = oracle.RandomState(0)
    vals = [rs.randint(10) for _ in range(10000)]
    np.testing.assert_almost_equal(np.mean(vals), 5.018)
    vals = [rs.binomial(1000, 0.8) for _ in range(10000)]
    np.testing.assert_almost_equal(np.mean(vals), 799.8564)


def test_kat_shuffle_permutation(oracle):
    ind = np.arange(10)
    oracle.RandomState(0).shuffle(ind)
    np.testing.assert_array_equal(ind, [2, 8, 4, 9, 1, 6, 7, 3, 0, 5])
    np.testing.assert_array_equal(oracle.RandomState(0).permutation(10), [2, 8, 4, 9, 1, 6, 7, 3, 0, 5])
    a, b = np.arange(10), np.arange(9, -1, -1)
    perm = oracle.RandomState(0).shuffle_with_trace([a, b])
    np.testing.assert_array_equal(a, [2, 8, 4, 9, 1, 6, 7, 3, 0, 5])
    np.testing.assert_array_equal(b, [7, 1, 5, 0, 8, 3, 2, 6, 9, 4])
    np.testing.assert_array_equal(a, perm)


# ---- reference KATs: modl/utils/randomkit/tests/test_sampler.py:6-45 ----
def test_kat_sampler(oracle):
    s = oracle.Sampler(100, True, True, 0)
    np.testing.assert_array_equal(s.yield_subset(10), [14, 58, 11, 49, 36, 62, 87, 45, 72, 47, 48, 13, 98, 97, 25, 93])
    assert np.mean([s.yield_subset(10).shape[0] for _ in range(100)]) == 10.19
    s = oracle.Sampler(100, False, False, 0)
    A = np.concatenate([s.yield_subset(10) for _ in range(10)])
    np.testing.assert_array_equal(np.sort(A), np.arange(100))
    s = oracle.Sampler(100, False, True, 0)
    np.testing.assert_array_equal(s.yield_subset(10), [6, 55, 1, 25, 87, 49, 69, 63, 13, 8])
    s = oracle.Sampler(100, True, False, 0)
    A = np.concatenate([s.yield_subset(10) for _ in range(20)])
    np.testing.assert_array_equal(np.sort(A[:100]), np.arange(100))


def test_golden_rng(oracle, golden):
    g = golden("rng.npz")
    for seed in (0, 42, 2 ** 35 + 11):
        rs = oracle.RandomState(seed)
        np.testing.assert_array_equal([rs.randint(10) for _ in range(64)], g["randint10_%d" % seed])
        np.testing.assert_array_equal([rs.randint(2 ** 40) for _ in range(32)], g["randint_big_%d" % seed])
        for n, p in ((1000, 0.8), (100, 0.1), (10000, 0.125), (200000, 1. / 12), (50, 0.3)):
            np.testing.assert_array_equal([rs.binomial(n, p) for _ in range(48)],
                                          g["binom_%d_%d_%g" % (seed, n, p)])
        np.testing.assert_array_equal(rs.permutation(37), g["perm_%d" % seed])
        a, b = np.arange(23), np.arange(22, -1, -1)
        np.testing.assert_array_equal(rs.shuffle_with_trace([a, b]), g["trace_%d" % seed])
        np.testing.assert_array_equal(a, g["trace_a_%d" % seed])
        np.testing.assert_array_equal(b, g["trace_b_%d" % seed])
    for rand_size in (0, 1):
        for repl in (0, 1):
            for rng_, red in ((100, 10), (1000, 8), (57, 3.5), (20, 1), (20, 2)):
                s = oracle.Sampler(rng_, rand_size, repl, 5)
                subs = [s.yield_subset(red) for _ in range(40)]
                tag = "samp_%d_%d_%d_%g" % (rand_size, repl, rng_, red)
                np.testing.assert_array_equal([len(x) for x in subs], g[tag + "_len"])
                np.testing.assert_array_equal(np.concatenate(subs), g[tag + "_cat"])
    bw = [oracle.batch_weight(c, b, lr, off) for c, b, lr, off in
          ((512, 512, 1., 0.), (1024, 512, .9, 0.), (30, 10, .95, 0.), (5000, 7, .92, 0.), (70, 10, 1., 3.))]
    np.testing.assert_array_equal(bw, g["batch_weight"])


def test_golden_enet(oracle, golden):
    g = golden("enet.npz")
    for dt in (np.float32, np.float64):
        v = g["v_%s" % dt.__name__]
        for l1 in (0., 0.15, 0.5, 1.):
            for radius in (0.5, 1., 40.):
                want = g["proj_%s_%g_%g" % (dt.__name__, l1, radius)]
                for i in range(v.shape[0]):
                    out = np.zeros(v.shape[1], dt)
                    oracle.enet_projection(v[i].copy(), out, radius, l1)
                    np.testing.assert_array_equal(out, want[i])      # restatement is bit-exact
            want = g["norm_%s_%g" % (dt.__name__, l1)]
            got = np.array([oracle.enet_norm(v[i], l1) for i in range(v.shape[0])], dtype=dt)
            np.testing.assert_array_equal(got, want)
            sc = v.copy()
            for i in range(v.shape[0]):
                oracle.enet_scale(sc[i], l1, 1.)
            np.testing.assert_array_equal(sc, g["scale_%s_%g" % (dt.__name__, l1)])


# reference's own enet tests restated on the oracle: modl/utils/math/tests/test_enet.py:99-156
def test_enet_properties(oracle):
    rng = np.random.RandomState(0)
    for _ in range(5):
        a = rng.randn(20000)
        a /= np.sqrt(np.sum(a ** 2))
        c = np.zeros(20000)
        oracle.enet_projection(a, c, 1, 0.15)
        np.testing.assert_almost_equal(oracle.enet_norm(c, 0.15), 1.0)
    for _ in range(5):
        a = rng.randn(100)
        c = np.zeros(100)
        oracle.enet_projection(a, c, 2, 0.0)
        np.testing.assert_almost_equal(np.sqrt(np.sum(c ** 2)), np.sqrt(2))
        b = np.zeros(100)
        oracle.enet_projection(a, b, 1, 1.0)
        np.testing.assert_almost_equal(np.sum(np.abs(b)), 1.0)
    a = rng.randn(100)
    for r in (1., 2.):
        for l1 in (0., 0.5, 1.):
            oracle.enet_scale(a, l1, r)
            np.testing.assert_almost_equal(oracle.enet_norm(a, l1), r)


def test_golden_regression(oracle, golden):
    g = golden("regression.npz")
    cases = g["cases"]
    for dt, tol_cmp in ((np.float32, 5e-6), (np.float64, 1e-13)):
        tag = dt.__name__
        G, Dx, X, idx, code0, Gm = (g[n + "_" + tag] for n in ("G", "Dx", "X", "idx", "code0", "Gm"))
        for ci, (l1, alpha, pos, tol) in enumerate(cases):
            c = code0.copy()
            oracle.enet_regression_single_gram(G, Dx.copy(), X, c, idx, l1, alpha, bool(pos), tol, 100)
            assert rel_err(c, g["single_%s_%d" % (tag, ci)]) < tol_cmp
            # rows not in idx untouched
            mask = np.ones(c.shape[0], bool)
            mask[idx] = False
            np.testing.assert_array_equal(c[mask], code0[mask])
            c = code0.copy()
            oracle.enet_regression_multi_gram(Gm.copy(), Dx.copy(), X, c, idx, l1, alpha, bool(pos), tol, 100)
            assert rel_err(c, g["multi_%s_%d" % (tag, ci)]) < tol_cmp
        ga = Gm.copy()
        oracle.update_G_average(ga, G, g["ws_" + tag])
        np.testing.assert_array_equal(ga, g["gavg_" + tag])


def test_golden_update_dict(oracle, golden):
    g = golden("update_dict.npz")
    for dt, tol_cmp in ((np.float32, 2e-5), (np.float64, 1e-12)):
        for ci, (l1, pos, full) in enumerate(g["cases"]):
            tag = "%s_%d" % (dt.__name__, ci)
            D = g["D0_" + tag].copy()
            subset, order = g["subset_" + tag], g["order_" + tag]
            D_sub = np.ascontiguousarray(D[:, subset])
            grad = np.ascontiguousarray(g["B_" + tag][:, subset])
            norm = g["norm0_" + tag].copy()
            oracle.update_dict_panel(D_sub, grad, g["C_" + tag], norm, order, l1, bool(pos))
            D[:, subset] = D_sub
            assert rel_err(D, g["D1_" + tag]) < tol_cmp
            assert np.abs(norm - g["norm1_" + tag]).max() < tol_cmp * 10


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_golden_fit(oracle, golden, dt):
    g = golden("fit.npz")
    cases, fit_data = _fit_cases()
    X, k = fit_data(dt)
    # f32: rounding-order differences (BLAS vs plain loops) flip a few CD stop decisions and
    # compound over ~12 minibatches; single-step tests above are the tight ones
    tol = 2e-3 if dt == np.float32 else 1e-10
    for ci, kw in enumerate(cases):
        est = oracle.OracleDictFact(n_components=k, random_state=0, **kw).fit(X)
        tag = "%s_%d" % (dt.__name__, ci)
        assert rel_err(est.components_, g["D_" + tag]) < tol, (ci, kw)
        assert rel_err(est.code_, g["code_" + tag]) < tol, (ci, kw)
        assert rel_err(est.C_, g["C_" + tag]) < tol
        assert rel_err(est.B_, g["B_" + tag]) < tol
        assert np.abs(est.comp_norm_ - g["norm_" + tag]).max() < tol
        np.testing.assert_array_equal(est.labels_, g["labels_" + tag])
        np.testing.assert_array_equal(est.sample_n_iter_, g["sni_" + tag])
        assert rel_err(est.transform(X), g["T_" + tag]) < tol
        assert abs(est.score(X) - float(g["score_" + tag])) < tol * max(1., abs(float(g["score_" + tag])))


def test_oracle_vs_compiled_reference(oracle, reference):
    """When oracle/_ref is present, check the restatement live on fresh random inputs."""
    rng = np.random.RandomState(3)
    n, p, k = 200, 80, 12
    X = (rng.randn(n, k) @ rng.randn(k, p) + 0.1 * rng.randn(n, p)).astype(np.float32)
    kw = dict(n_components=k, random_state=1, reduction=4, batch_size=25, n_epochs=2, code_alpha=0.05)
    a = reference.DictFact(**kw).fit(X)
    b = oracle.OracleDictFact(**kw).fit(X)
    assert rel_err(b.components_, a.components_) < 1e-4
    assert rel_err(b.code_, a.code_) < 1e-4
    np.testing.assert_array_equal(a.labels_, b.labels_)


def test_golden_recsys(oracle, golden):
    """OracleRecsysDictFact against whole fits of the unmodified reference (tests/golden/recsys.npz)."""
    import json
    from test_recsys import _data
    g = golden("recsys.npz")
    for ci, case in enumerate(json.loads(str(g["cases"]))):
        X, X_te = _data(case["data"], case["dtype"])
        kw = {k: v for k, v in case["kw"].items()}
        if "crop" in kw:
            kw["crop"] = tuple(kw["crop"])
        est = oracle.OracleRecsysDictFact(random_state=0, **kw).fit(X)
        tag = "fit_%d_" % ci
        tol = 1e-10 if case["dtype"] == "float64" else 1e-4
        assert est.n_iter_ == int(g[tag + "n_iter_"])
        np.testing.assert_array_equal(est.feature_n_iter_, g[tag + "feature_n_iter_"])
        for name in ("components_", "code_", "C_", "B_"):
            assert rel_err(getattr(est, name), g[tag + name]) < tol, (ci, name)
        if case["dtype"] == "float64":
            assert rel_err(est.predict_data(X_te), g[tag + "pred_test"]) < tol
    row, col = oracle.recsys_biases(_data("ratings", "float64")[0], beta=3.)
    np.testing.assert_allclose(row, g["bias_row"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(col, g["bias_col"], rtol=1e-13, atol=1e-15)
