"""GPU parity of the individual C-ABI entry points against the golden vectors (generated from
the unmodified reference) and the CPU oracle.  Bit-exact for indices; floating tolerances are
written next to each check."""
import numpy as np
import pytest

from conftest import rel_err

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

F32_TOL = 2e-5      # one step, float32: different summation order than BLAS, same algorithm
F64_TOL = 1e-11


def _tol(dt):
    return F32_TOL if dt == np.float32 else F64_TOL


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def test_native_library_is_loaded(dev):
    from modl_b200 import _lib
    ctx = _lib.get_context(0)
    assert ctx.sm_count > 0
    maps = open("/proc/self/maps").read()
    assert "libmodl_b200.so" in maps


def test_enet_golden(dev, golden):
    from modl_b200.enet import enet_norm, enet_projection, enet_scale
    g = golden("enet.npz")
    for dt in (np.float32, np.float64):
        v = g["v_%s" % dt.__name__]
        for l1 in (0., 0.15, 0.5, 1.):
            for radius in (0.5, 1., 40.):
                want = g["proj_%s_%g_%g" % (dt.__name__, l1, radius)]
                for i in range(v.shape[0]):
                    out = np.zeros(v.shape[1], dt)
                    enet_projection(v[i].copy(), out, radius, l1)
                    assert rel_err(out, want[i]) < (2e-6 if dt == np.float32 else 1e-13), (dt, l1, radius, i)
            got = np.array([enet_norm(v[i], l1) for i in range(v.shape[0])])
            assert rel_err(got, g["norm_%s_%g" % (dt.__name__, l1)]) < (1e-6 if dt == np.float32 else 1e-14)
            sc = v.copy()
            enet_scale(sc, l1, 1.)
            assert rel_err(sc, g["scale_%s_%g" % (dt.__name__, l1)]) < (1e-6 if dt == np.float32 else 1e-14)


def test_enet_projection_large_vectors(dev, oracle):
    from modl_b200.enet import enet_norm, enet_projection
    rng = np.random.RandomState(0)
    for n, l1, radius in ((20000, 0.15, 1.), (16667, 1.0, 3.), (57344, 0.5, 10.), (1250, 0., 0.7)):
        a = rng.randn(n).astype(np.float32)
        want = np.zeros_like(a)
        oracle.enet_projection(a, want, radius, l1)
        got = np.zeros_like(a)
        enet_projection(a, got, radius, l1)
        assert rel_err(got, want) < 2e-6, (n, l1)
        np.testing.assert_allclose(enet_norm(got, l1), radius, rtol=2e-5)


def test_gram_dx(dev):
    from modl_b200 import _lib
    from modl_b200._util import ptr, stream_of
    rng = np.random.RandomState(1)
    for dt in (np.float32, np.float64):
        for (k, b, p, s) in ((16, 10, 50, 20), (70, 33, 301, 77), (256, 64, 1000, 130)):
            D = rng.randn(k, p).astype(dt)
            X = rng.randn(b, p).astype(dt)
            subset = rng.permutation(p)[:s].astype(np.int64)
            Dd, Xd, sd = (torch.from_numpy(a).to(dev) for a in (D, X, subset))
            G = torch.empty((k, k), dtype=Dd.dtype, device=dev)
            Dx = torch.empty((b, k), dtype=Dd.dtype, device=dev)
            xn = torch.empty((b,), dtype=Dd.dtype, device=dev)
            fn = getattr(_lib.lib(), "modl_gram_dx_" + _lib.sfx_of(Dd.dtype))
            _lib.check(fn(_lib.get_context(0).handle, ptr(Dd), p, ptr(Xd), p, ptr(sd), s, k, b, p, 3.0,
                          ptr(G), ptr(Dx), ptr(xn), stream_of(dev)))
            Ds, Xs = D[:, subset].astype(np.float64), X[:, subset].astype(np.float64)
            tol = 2e-6 if dt == np.float32 else 1e-14
            assert rel_err(G.cpu().numpy(), 3 * Ds @ Ds.T) < tol
            assert rel_err(Dx.cpu().numpy(), 3 * Xs @ Ds.T) < tol
            assert rel_err(xn.cpu().numpy(), (X.astype(np.float64) ** 2).sum(1)) < tol
            Gh = G.cpu().numpy()
            if dt == np.float64:
                np.testing.assert_array_equal(Gh, Gh.T)      # exactly symmetric, like syrk
            else:
                # tensor-core path: the two cross terms of the 3xTF32 split enter (i, j) and (j, i)
                # in a different order; every consumer reads the lower triangle only
                assert rel_err(Gh, Gh.T) < 1e-6
            _lib.check(fn(_lib.get_context(0).handle, ptr(Dd), p, ptr(Xd), p, None, 0, k, b, p, 1.0,
                          ptr(G), ptr(Dx), ptr(xn), stream_of(dev)))
            assert rel_err(G.cpu().numpy(), D.astype(np.float64) @ D.T.astype(np.float64)) < tol
            assert rel_err(Dx.cpu().numpy(), X.astype(np.float64) @ D.T.astype(np.float64)) < tol


def test_regression_golden(dev, golden, oracle):
    from modl_b200.dict_fact_fast import _enet_regression_multi_gram, _enet_regression_single_gram, _update_G_average
    g = golden("regression.npz")
    for dt in (np.float32, np.float64):
        tag = dt.__name__
        G, Dx, X, idx, code0, Gm = (g[n + "_" + tag] for n in ("G", "Dx", "X", "idx", "code0", "Gm"))
        for ci, (l1, alpha, pos, tol) in enumerate(g["cases"]):
            for kind, fn, Gin in (("single", _enet_regression_single_gram, G), ("multi", _enet_regression_multi_gram, Gm)):
                c = code0.copy()
                sw = np.zeros(len(idx), np.int32)
                fn(Gin.copy(), Dx.copy(), X, c, idx, l1, alpha, bool(pos), tol, 100, sweeps=sw)
                want = g["%s_%s_%d" % (kind, tag, ci)]
                assert rel_err(c, want) < _tol(dt), (kind, tag, ci, rel_err(c, want))
                mask = np.ones(c.shape[0], bool)
                mask[idx] = False
                np.testing.assert_array_equal(c[mask], code0[mask])        # other rows untouched
                if l1 > 0:    # the stop decision must agree sample by sample
                    ofn = oracle.enet_regression_single_gram if kind == "single" else oracle.enet_regression_multi_gram
                    _, osw, _ = ofn(Gin.copy(), Dx.copy(), X, code0.copy(), idx, l1, alpha, bool(pos), tol, 100,
                                    return_sweeps=True)
                    np.testing.assert_array_equal(sw, osw)
        ga = Gm.copy()
        _update_G_average(ga, G, g["ws_" + tag])
        assert rel_err(ga, g["gavg_" + tag]) < (2e-7 if dt == np.float32 else 1e-15)


@pytest.mark.parametrize("k", [5, 32, 33, 100, 256, 300, 330, 520])
def test_cd_all_tile_counts(dev, oracle, k):
    """Every register-tile instantiation of the CD kernel (shared-memory Gram up to k = 320,
    streamed Gram beyond), ragged k, zero diagonal entries, positivity."""
    from modl_b200.dict_fact_fast import _enet_regression_single_gram
    rng = np.random.RandomState(k)
    b, p = 37, 2 * k + 7
    D = (rng.randn(k, p) / np.sqrt(p)).astype(np.float32)
    if k > 3:
        D[3] = 0          # zero Gram row/diagonal -> coordinate skipped [ref: dict_fact_fast.pyx:357-358]
    X = rng.randn(b, p).astype(np.float32)
    G, Dx = np.ascontiguousarray(D @ D.T), np.ascontiguousarray(X @ D.T)
    for pos in (False, True):
        c0 = np.ones((b, k), np.float32)
        want, osw, _ = oracle.enet_regression_single_gram(G, Dx.copy(), X, c0.copy(), np.arange(b), 0.9, 0.05, pos,
                                                          1e-3, 100, return_sweeps=True)
        got = c0.copy()
        sw = np.zeros(b, np.int32)
        _enet_regression_single_gram(G, Dx.copy(), X, got, np.arange(b), 0.9, 0.05, pos, 1e-3, 100, sweeps=sw)
        assert rel_err(got, want) < 1e-4, (k, pos, rel_err(got, want), (sw != osw).sum())
        assert (sw != osw).sum() <= 1, (sw, osw)


@pytest.mark.parametrize("k", [4, 40, 70, 256, 300])
def test_ridge(dev, oracle, k):
    from modl_b200.dict_fact_fast import _enet_regression_multi_gram, _enet_regression_single_gram
    rng = np.random.RandomState(k)
    b, p = 29, 3 * k
    for dt in (np.float32, np.float64):
        D = (rng.randn(k, p) / np.sqrt(p)).astype(dt)
        X = rng.randn(b, p).astype(dt)
        G, Dx = np.ascontiguousarray(D @ D.T), np.ascontiguousarray(X @ D.T)
        want = np.linalg.solve(G.astype(np.float64) + 0.1 * np.eye(k), Dx.T.astype(np.float64)).T
        got = np.ones((b, k), dt)
        dx = Dx.copy()
        _enet_regression_single_gram(G, dx, X, got, np.arange(b), 0., 0.1, False, 1e-2, 100)
        tol = 3e-5 if dt == np.float32 else 1e-11
        assert rel_err(got, want) < tol, (k, dt, rel_err(got, want))
        np.testing.assert_array_equal(dx, got)          # Dx is overwritten with the solution
        Gm = (np.tile(G, (b, 1, 1)) * (1 + 0.02 * np.arange(b)[:, None, None])).astype(dt)
        wantm = np.stack([np.linalg.solve(Gm[i].astype(np.float64) + 0.1 * np.eye(k), Dx[i].astype(np.float64))
                          for i in range(b)])
        gotm = np.ones((b, k), dt)
        _enet_regression_multi_gram(Gm, Dx.copy(), X, gotm, np.arange(b), 0., 0.1, False, 1e-2, 100)
        assert rel_err(gotm, wantm) < tol


def test_ridge_reports_non_spd(dev):
    from modl_b200 import ModlError
    from modl_b200.dict_fact_fast import _enet_regression_single_gram
    G = -np.eye(8, dtype=np.float32)
    with pytest.raises(ModlError):
        _enet_regression_single_gram(G, np.ones((3, 8), np.float32), np.ones((3, 5), np.float32),
                                     np.ones((3, 8), np.float32), np.arange(3), 0., 0.1, False, 1e-2, 10)


VARIANTS = {"blocked": dict(bcd_blocked=1, bcd_pilot=1, bcd_pipeline=1),
            "pipelined": dict(bcd_blocked=0, bcd_pilot=1, bcd_pipeline=1),
            "pilot": dict(bcd_blocked=0, bcd_pilot=1, bcd_pipeline=0),
            "plain": dict(bcd_blocked=0, bcd_pilot=0, bcd_pipeline=0)}


def _run_update_dict(dev, D0, B, C, norm0, subset, order, l1, pos, G0=None, mode=0, w=0.5, step=1.0, cluster=None,
                     variant="blocked"):
    from modl_b200 import _lib
    from modl_b200._util import ptr, stream_of
    k, p = D0.shape
    ctx = _lib.get_context(0)
    if cluster is not None:
        ctx.set_option("bcd_cluster", cluster)
    for name, val in VARIANTS[variant].items():
        ctx.set_option(name, val)
    try:
        Dd, Bd, Cd, nd, sd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (D0, B, C, norm0, subset))
        Gd = torch.from_numpy(G0.copy()).to(dev) if G0 is not None else None
        order = np.ascontiguousarray(order, dtype=np.int64)
        fn = getattr(_lib.lib(), "modl_update_dict_" + _lib.sfx_of(Dd.dtype))
        _lib.check(fn(ctx.handle, ptr(Dd), p, ptr(Bd), p, ptr(Cd), ptr(nd), ptr(Gd), ptr(sd), len(subset),
                      order.ctypes.data, k, p, float(l1), int(pos), mode, w, step, stream_of(dev)))
        torch.cuda.synchronize()
    finally:
        if cluster is not None:
            ctx.set_option("bcd_cluster", 16)
        for name, val in VARIANTS["blocked"].items():
            ctx.set_option(name, val)
    return Dd.cpu().numpy(), nd.cpu().numpy(), (Gd.cpu().numpy() if Gd is not None else None)


@pytest.mark.parametrize("variant", ["blocked", "pipelined", "pilot", "plain"])
@pytest.mark.parametrize("cluster", [16, 8, 2, 0])
def test_update_dict_golden(dev, golden, cluster, variant):
    """One BCD dictionary step vs the reference's _update_dict (golden), for every barrier
    flavour (16/8/2-CTA clusters, cooperative global barrier) and every kernel variant (block-wise
    coefficient-space solve, look-ahead pilot kernel with and without the pipelined norm exchange, plain
    per-atom kernel)."""
    g = golden("update_dict.npz")
    for dt in (np.float32, np.float64):
        for ci, (l1, pos, full) in enumerate(g["cases"]):
            tag = "%s_%d" % (dt.__name__, ci)
            G0 = g["G0_" + tag] if full else None
            D1, n1, G1 = _run_update_dict(dev, g["D0_" + tag], g["B_" + tag], g["C_" + tag], g["norm0_" + tag],
                                          g["subset_" + tag], g["order_" + tag], l1, pos, G0, cluster=cluster,
                                          variant=variant)
            tol = 2e-5 if dt == np.float32 else 1e-11
            assert rel_err(D1, g["D1_" + tag]) < tol, (tag, cluster, rel_err(D1, g["D1_" + tag]))
            assert np.abs(n1 - g["norm1_" + tag]).max() < 10 * tol, (tag, cluster)
            if full:
                assert rel_err(G1, g["G1_" + tag]) < 10 * tol
            # columns outside the subset are untouched
            mask = np.ones(D1.shape[1], bool)
            mask[g["subset_" + tag]] = False
            np.testing.assert_array_equal(D1[:, mask], g["D0_" + tag][:, mask])


@pytest.mark.parametrize("variant", ["blocked", "pipelined", "pilot"])
@pytest.mark.parametrize("shape", [(256, 10000, 1250), (70, 30000, 2500), (64, 4000, 4000), (256, 6000, 300), (37, 900, 333),
                                   (70, 40000, 17000), (24, 9000, 9000, "f64")])
def test_update_dict_bench_shapes(dev, oracle, shape, variant):
    """Config-2 / config-4-like panels against the oracle's BCD, L2 and L1 balls.  The last two shapes are the
    grid-wide (cooperative) kernel with the candidate row of the L1 projection held in registers: s = 17 000 is
    the subset of BASELINE configs[3] (p = 2e5, reduction 12); the float64 one takes the 48-per-thread variant."""
    k, p, s = shape[:3]
    dt = np.float64 if len(shape) > 3 else np.float32
    if variant != "blocked" and (len(shape) > 3 or shape[2] >= 4000):
        pytest.skip("panels beyond one cluster take the grid-wide kernel whatever the variant: one pass is enough")
    rng = np.random.RandomState(5)
    for l1, pos in ((0., False), (1., False), (0.3, True)):
        D0 = rng.randn(k, p).astype(dt)
        if pos:
            D0 = np.abs(D0)
        for i in range(k):
            oracle.enet_scale(D0[i], l1, 1.)
        A = (rng.randn(512, k) * (rng.rand(512, k) < 0.3)).astype(dt)
        C = np.ascontiguousarray((A.T @ A / 512).astype(dt))
        B = np.ascontiguousarray((A.T @ (A @ D0 + 0.1 * rng.randn(512, p).astype(dt)) / 512).astype(dt))
        subset = rng.permutation(p)[:s].astype(np.int64)
        order = rng.permutation(k).astype(np.int64)
        norm0 = np.zeros(k, dt)
        Dw = np.ascontiguousarray(D0[:, subset])
        gw = np.ascontiguousarray(B[:, subset])
        nw = norm0.copy()
        oracle.update_dict_panel(Dw, gw, C, nw, order, l1, pos)
        want = D0.copy()
        want[:, subset] = Dw
        D1, n1, _ = _run_update_dict(dev, D0, B, C, norm0, subset, order, l1, pos, variant=variant)
        tol = 5e-5 if dt == np.float32 else 1e-11
        assert rel_err(D1, want) < tol, (shape, l1, pos, rel_err(D1, want))
        assert np.abs(n1 - nw).max() < 10 * tol, (shape, l1, np.abs(n1 - nw).max())


def test_update_stats(dev):
    from modl_b200 import _lib
    from modl_b200._util import ptr, stream_of
    rng = np.random.RandomState(2)
    for dt in (np.float32, np.float64):
        n, b, k, p = 50, 24, 40, 333
        code = rng.randn(n, k).astype(dt)
        idx = rng.permutation(n)[:b].astype(np.int64)
        X = rng.randn(b, p).astype(dt)
        C0, B0 = rng.randn(k, k).astype(dt), rng.randn(k, p).astype(dt)
        cd, id_, Xd, Cd, Bd = (torch.from_numpy(a).to(dev) for a in (code, idx, X, C0, B0))
        fn = getattr(_lib.lib(), "modl_update_stats_" + _lib.sfx_of(cd.dtype))
        _lib.check(fn(_lib.get_context(0).handle, ptr(cd), ptr(id_), ptr(Xd), p, ptr(Cd), ptr(Bd), p, 0.3, b, k, p, 0,
                      stream_of(dev)))
        cb = code[idx].astype(np.float64)
        tol = 2e-6 if dt == np.float32 else 1e-14
        assert rel_err(Cd.cpu().numpy(), 0.7 * C0 + 0.3 / b * cb.T @ cb) < tol
        assert rel_err(Bd.cpu().numpy(), 0.7 * B0 + 0.3 / b * cb.T @ X) < tol
        Ch = Cd.cpu().numpy()
        _lib.check(fn(_lib.get_context(0).handle, ptr(cd), ptr(id_), ptr(Xd), p, ptr(Cd), ptr(Bd), p, 0.3, b, k, p, 1,
                      stream_of(dev)))
        assert rel_err(Cd.cpu().numpy(), cb.T @ cb / b) < tol
        assert rel_err(Bd.cpu().numpy(), cb.T @ X / b) < tol
        del Ch
