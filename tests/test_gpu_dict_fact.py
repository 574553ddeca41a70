"""GPU parity of the estimator (DictFact / Coder on the sm_100a path) against golden fits of
the unmodified reference, the CPU oracle and -- when oracle/_ref travelled with the repo -- the
compiled reference itself.  Also the reference's own end-to-end tests
[ref: modl/decomposition/tests/test_dict_fact.py:55-154] run on the GPU estimator."""
import numpy as np
import pytest

from conftest import rel_err
from test_oracle import _fit_cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _state_close(est, g, tag, tol, X):
    assert rel_err(est.components_, g["D_" + tag]) < tol, ("D", tag, rel_err(est.components_, g["D_" + tag]))
    assert rel_err(est.code_, g["code_" + tag]) < tol, ("code", tag, rel_err(est.code_, g["code_" + tag]))
    assert rel_err(est.C_, g["C_" + tag]) < tol
    assert rel_err(est.B_, g["B_" + tag]) < tol
    assert np.abs(est.comp_norm_ - g["norm_" + tag]).max() < tol
    np.testing.assert_array_equal(est.labels_, g["labels_" + tag])
    np.testing.assert_array_equal(est.sample_n_iter_, g["sni_" + tag])
    assert rel_err(est.transform(X), g["T_" + tag]) < tol
    want = float(g["score_" + tag])
    assert abs(est.score(X) - want) < tol * max(1., abs(want))
    if "G_" + tag in g.files:
        assert rel_err(est.G_, g["G_" + tag]) < 10 * tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_fit_matches_reference_golden(golden, dt):
    """Whole fits (2 epochs, all aggregation modes, ridge / lasso / positive / sgd) against the
    state the unmodified reference reaches from the same seed.  float32: 2e-3 over ~12
    chained minibatches (rounding-order noise compounds; the oracle itself sits at
    3e-4 from the reference here); float64: 1e-9."""
    from modl_b200 import DictFact
    g = golden("fit.npz")
    cases, fit_data = _fit_cases()
    X, k = fit_data(dt)
    tol = 2e-3 if dt == np.float32 else 1e-9
    for ci, kw in enumerate(cases):
        est = DictFact(n_components=k, random_state=0, **kw).fit(X)
        _state_close(est, g, "%s_%d" % (dt.__name__, ci), tol, X)


def _planted(n, p, k, seed=0, density=0.3, noise=0.1):
    rng = np.random.RandomState(seed)
    D0 = rng.randn(k, p)
    D0 /= np.linalg.norm(D0, axis=1, keepdims=True)
    A = rng.randn(n, k) * (rng.rand(n, k) < density)
    return (A @ D0 + noise * rng.randn(n, p)).astype(np.float32)


def test_single_step_config2_shape_vs_oracle(oracle):
    """BASELINE config 2 shape (k=256, p=10000, batch=512, reduction=8, l1 coding), two
    minibatches from identical state: bookkeeping bit-exact, code within 1e-4 relative
    (north_star tolerance), sweep counts compared sample by sample."""
    from modl_b200 import DictFact
    n, p, k, b = 1024, 10000, 256, 512
    X = _planted(n, p, k)
    kw = dict(n_components=k, batch_size=b, reduction=8, code_l1_ratio=1., code_alpha=1., tol=1e-2, max_iter=100,
              random_state=0)
    orc = oracle.OracleDictFact(**kw)
    orc.prepare(n_samples=n, X=X[:k])
    est = DictFact(**kw)
    est.record_sweeps = True
    est.prepare(n_samples=n, X=X[:k])
    assert rel_err(est.components_, orc.components_) < 1e-6
    for step in range(2):
        sl = slice(step * b, (step + 1) * b)
        idx = np.arange(sl.start, sl.stop)
        orc.partial_fit(X[sl], idx)
        est.partial_fit(X[sl], idx)
        np.testing.assert_array_equal(est.last_subset_, orc.subset_log_[-1])      # bit-exact
        np.testing.assert_array_equal(est.last_order_, orc.order_log_[-1])
        assert est.n_iter_ == orc.n_iter_
        np.testing.assert_array_equal(est.sample_n_iter_, orc.sample_n_iter_)
        sw, osw = est.last_sweeps_[:b], orc.sweep_log_[-1]
        n_diff = int((sw != osw).sum())
        err_code = rel_err(est.code_[sl], orc.code_[sl])
        print("step %d: code rel err %.3g, sweeps differ on %d/%d samples, mean sweeps %.2f, density %.3f"
              % (step, err_code, n_diff, b, osw.mean(), (orc.code_[sl] != 0).mean()))
        assert err_code < 1e-4
        assert n_diff <= b // 100
        assert rel_err(est.C_, orc.C_) < 1e-4
        assert rel_err(est.B_, orc.B_) < 1e-4
        assert rel_err(est.components_, orc.components_) < 1e-4
        assert np.abs(est.comp_norm_ - orc.comp_norm_).max() < 1e-4


def test_config2_vs_compiled_reference(reference):
    """Same as above against the UNMODIFIED reference (oracle/_ref), 3 minibatches."""
    from modl_b200 import DictFact
    n, p, k, b = 1536, 10000, 256, 512
    X = _planted(n, p, k, seed=1)
    kw = dict(n_components=k, batch_size=b, reduction=8, code_l1_ratio=1., code_alpha=1., tol=1e-2, max_iter=100,
              random_state=0)
    ref = reference.DictFact(**kw)
    ref.prepare(n_samples=n, X=X[:k])
    est = DictFact(**kw)
    est.prepare(n_samples=n, X=X[:k])
    for step in range(3):
        sl = slice(step * b, (step + 1) * b)
        idx = np.arange(sl.start, sl.stop)
        ref.partial_fit(X[sl], idx)
        est.partial_fit(X[sl], idx)
        e = rel_err(est.code_[sl], ref.code_[sl])
        print("step %d vs reference: code %.3g  D %.3g" % (step, e, rel_err(est.components_, ref.components_)))
        assert e < 1e-4
    assert rel_err(est.components_, ref.components_) < 2e-4
    assert est.n_iter_ == ref.n_iter_


def test_config1_cpu_runnable_case(oracle):
    """BASELINE config 0: 2000 x 500, n_components=16, reduction=1, ridge coding, float64."""
    from modl_b200 import DictFact
    rng = np.random.RandomState(0)
    X = rng.randn(2000, 500)
    kw = dict(n_components=16, reduction=1, code_l1_ratio=0, random_state=0)
    a = oracle.OracleDictFact(**kw).fit(X)
    b = DictFact(**kw).fit(X)
    assert rel_err(b.components_, a.components_) < 1e-9
    assert rel_err(b.code_, a.code_) < 1e-9


# ---- the reference's own end-to-end tests on the GPU estimator ------------------------------
solver_dict = {
    'masked': {'Dx_agg': 'masked', 'G_agg': 'masked'},
    'gram': {'Dx_agg': 'masked', 'G_agg': 'full'},
    'average': {'Dx_agg': 'average', 'G_agg': 'average'},
    'full': {'Dx_agg': 'full', 'G_agg': 'full'},
}


def generate_synthetic(n_samples=200, n_components=4, n_features=16, dictionary_rank=None):
    rng = np.random.RandomState(0)
    if dictionary_rank is None:
        Q = rng.randn(n_components, n_features)
    else:
        Q = rng.randn(n_components, dictionary_rank).dot(rng.randn(dictionary_rank, n_features))
    code = rng.randn(n_samples, n_components)
    return code.dot(Q), Q


def generate_sparse_synthetic(n_samples=200, square_size=4):
    rng = np.random.RandomState(0)
    half = square_size // 2
    Q = np.zeros((4, square_size ** 2))
    for i in range(2):
        for j in range(2):
            atom = np.zeros((square_size, square_size))
            atom[half * i:half * (i + 1), half * j:half * (j + 1)] = 1
            Q[2 * i + j] = atom.ravel()
    return rng.randn(n_samples, 4).dot(Q), Q


@pytest.mark.parametrize("solver", list(solver_dict))
def test_dict_mf_reconstruction(solver):
    from modl_b200 import DictFact
    X, Q = generate_synthetic()
    est = DictFact(n_components=4, code_alpha=1e-4, n_epochs=5, comp_l1_ratio=0, random_state=0, reduction=1,
                   **solver_dict[solver]).fit(X)
    Y = est.transform(X).dot(est.components_)
    assert np.sum((X - Y) ** 2) / np.sum(X ** 2) < 0.02


@pytest.mark.parametrize("solver", list(solver_dict))
def test_dict_mf_reconstruction_reduction(solver):
    from modl_b200 import DictFact
    X, Q = generate_synthetic(n_features=20, n_samples=400, dictionary_rank=4)
    est = DictFact(n_components=4, code_alpha=1e-4, n_epochs=2, comp_l1_ratio=0, random_state=0, reduction=2,
                   **solver_dict[solver]).fit(X)
    Y = est.transform(X).dot(est.components_)
    assert np.sum((X - Y) ** 2) / np.sum(X ** 2) < 0.02


@pytest.mark.parametrize("solver", list(solver_dict))
def test_dict_mf_reconstruction_reproductible(solver):
    """Two fits from the same seed are BITWISE identical (no atomics anywhere on the path)."""
    from modl_b200 import DictFact
    X, Q = generate_synthetic(n_features=20, n_samples=400, dictionary_rank=4)
    est = DictFact(n_components=4, code_alpha=1e-4, n_epochs=2, comp_l1_ratio=0, random_state=0, reduction=2,
                   **solver_dict[solver])
    est.fit(X)
    D1, P1 = est.components_.copy(), est.transform(X)
    est.random_state = 0
    est.fit(X)
    np.testing.assert_array_equal(D1, est.components_)
    np.testing.assert_array_equal(P1, est.transform(X))


@pytest.mark.parametrize("solver", list(solver_dict))
def test_dict_mf_reconstruction_reduction_batch(solver):
    from modl_b200 import DictFact
    X, Q = generate_synthetic(n_features=20, n_samples=400, dictionary_rank=4)
    est = DictFact(n_components=4, code_alpha=1e-4, n_epochs=2, comp_l1_ratio=0, random_state=0, reduction=2,
                   batch_size=10, **solver_dict[solver]).fit(X)
    Y = est.transform(X).dot(est.components_)
    assert np.sum((X - Y) ** 2) / np.sum(X ** 2) < 0.06


@pytest.mark.parametrize("solver", list(solver_dict))
def test_dict_mf_reconstruction_sparse_dict(solver):
    from modl_b200 import DictFact
    X, Q = generate_sparse_synthetic(500, 4)
    rng = np.random.RandomState(0)
    dict_init = Q + rng.randn(*Q.shape) * 0.2
    est = DictFact(n_components=4, code_alpha=1e-2, n_epochs=2, code_l1_ratio=0, comp_l1_ratio=1,
                   dict_init=dict_init, random_state=0, **solver_dict[solver]).fit(X)
    Q_rec = est.components_
    Q_rec /= np.sqrt(np.sum(Q_rec ** 2, axis=1))[:, np.newaxis]
    Qn = Q / np.sqrt(np.sum(Q ** 2, axis=1))[:, np.newaxis]
    G = np.abs(Q_rec.dot(Qn.T))
    assert min(np.sum(np.any(G > 0.95, axis=1)), np.sum(np.any(G > 0.95, axis=0))) >= 4


def test_bitwise_reproducible_at_config2_shape():
    from modl_b200 import DictFact
    X = _planted(1024, 10000, 256, seed=3)
    outs = []
    for _ in range(2):
        est = DictFact(n_components=256, batch_size=512, reduction=8, random_state=0)
        est.prepare(n_samples=1024, X=X[:256])
        est.partial_fit(X)
        outs.append((est.components_, est.code_))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


def test_input_kinds_agree():
    """NumPy rows, pinned host tensor and CUDA tensor inputs give identical results."""
    from modl_b200 import DictFact
    X = _planted(300, 600, 20, seed=4)
    res = []
    for kind in ("numpy", "pinned", "cuda"):
        est = DictFact(n_components=20, batch_size=64, reduction=3, random_state=0)
        est.prepare(n_samples=300, X=X[:20])
        data = X if kind == "numpy" else (torch.from_numpy(X).pin_memory() if kind == "pinned"
                                           else torch.from_numpy(X).cuda())
        est.partial_fit(data)
        res.append(est.components_)
    np.testing.assert_array_equal(res[0], res[1])
    np.testing.assert_array_equal(res[0], res[2])


def test_coder_matches_dictfact_transform():
    from modl_b200 import Coder, DictFact
    X = _planted(200, 300, 10, seed=5)
    est = DictFact(n_components=10, batch_size=50, reduction=2, random_state=0, code_alpha=0.1).fit(X)
    coder = Coder(est.components_, code_alpha=0.1)
    np.testing.assert_array_equal(coder.transform(X), est.transform(X))


@pytest.mark.parametrize("kw", [dict(), dict(Dx_agg='full', G_agg='full'), dict(Dx_agg='average', G_agg='average'),
                                dict(code_l1_ratio=0., comp_l1_ratio=1.), dict(optimizer='sgd', step_size=0.1)])
def test_overlapped_statistics_equal_the_fused_step(kw):
    """With overlap_stats the step computes the k x s slice of the B_ update first and runs the full
    k x p product on a second stream behind the dictionary update (the phases the sharded estimator
    uses to hide its all-reduce); it must reach the same state as the single fused call (same
    mathematics, different GEMM splits: 1e-5)."""
    from modl_b200 import DictFact
    X = _planted(768, 3000, 64, seed=8)
    res = []
    for overlap in (True, False):
        est = DictFact(n_components=64, batch_size=128, reduction=4, random_state=0, **kw)
        est.prepare(n_samples=768, X=X[:64])
        est.overlap_stats = overlap
        est.partial_fit(X)
        res.append((est.components_, est.B_, est.C_, est.code_, est.comp_norm_))
    for a, c in zip(*res):
        assert rel_err(a, c.astype(np.float64)) < 1e-5


def test_async_host_copy_matches_synchronous():
    from modl_b200 import DictFact
    X = _planted(1024, 2000, 32, seed=9)
    Xp = torch.from_numpy(X).pin_memory()
    res = []
    for mode in (False, True):
        est = DictFact(n_components=32, batch_size=128, reduction=4, random_state=0, async_host_copy=mode)
        est.prepare(n_samples=1024, X=X[:32])
        for i in range(0, 1024, 128):                       # one batch per call
            est.partial_fit(Xp[i:i + 128], np.arange(i, i + 128))
        est.synchronize()
        res.append(est.components_)
    np.testing.assert_array_equal(res[0], res[1])


@pytest.mark.parametrize("name,cfg", [
    # BASELINE configs[2]: ImageDictFact's DictFact (16 x 16 x 224 patches, NMF setting: positive codes and atoms)
    ("config3_image_nmf", dict(p=57344, k=256, b=512, kw=dict(reduction=8, code_l1_ratio=1., code_alpha=0.1, comp_l1_ratio=0.,
                                                                 code_pos=True, comp_pos=True, tol=1e-2))),
    # BASELINE configs[3]: fMRIDictFact's DictFact (ridge code, L1-ball atoms), p ~ 2e5, k = 70, reduction 12
    ("config4_fmri", dict(p=200000, k=70, b=200, kw=dict(reduction=12, code_l1_ratio=0., code_alpha=1., comp_l1_ratio=1.))),
])
def test_single_step_other_baseline_shapes_vs_oracle(oracle, name, cfg):
    """One minibatch at the full feature width of BASELINE configs 3 and 4 (panels that do NOT fit one
    cluster: the cooperative dictionary-update kernel; positive / L1-ball projections; ridge coding)."""
    from modl_b200 import DictFact
    p, k, b = cfg["p"], cfg["k"], cfg["b"]
    n = b + k
    rng = np.random.RandomState(12)
    # sparse non-negative atoms with little overlap: a well-conditioned Gram matrix, so that the test measures
    # the kernels and not the chaos of a slowly converging coordinate descent (dense positive atoms are so
    # coherent that float32 and float64 REFERENCE codes already differ by percents)
    D0 = (np.abs(rng.randn(k, p)) * (rng.rand(k, p) < 0.03)).astype(np.float32)
    D0 /= np.linalg.norm(D0, axis=1, keepdims=True)
    A = (np.abs(rng.randn(n, k)) * (rng.rand(n, k) < 0.1)).astype(np.float32)
    X = A @ D0 + 0.001 * np.abs(rng.randn(n, p)).astype(np.float32)
    Dinit = (D0 + 0.05 * np.abs(rng.randn(k, p)).astype(np.float32) * (D0 > 0)).astype(np.float32)
    kw = dict(n_components=k, batch_size=b, random_state=0, max_iter=100, **cfg["kw"])
    orc = oracle.OracleDictFact(**kw)
    orc.prepare(n_samples=n, X=Dinit)
    est = DictFact(**kw)
    est.prepare(n_samples=n, X=Dinit)
    idx = np.arange(k, k + b)
    orc.partial_fit(X[k:], idx)
    est.partial_fit(X[k:], idx)
    np.testing.assert_array_equal(est.last_subset_, orc.subset_log_[-1])
    np.testing.assert_array_equal(est.last_order_, orc.order_log_[-1])
    errs = dict(code=rel_err(est.code_[idx], orc.code_[idx]), C=rel_err(est.C_, orc.C_), B=rel_err(est.B_, orc.B_),
                D=rel_err(est.components_, orc.components_))
    print(name, errs)
    assert errs["code"] < 1e-4 and errs["C"] < 1e-4 and errs["B"] < 1e-4 and errs["D"] < 1e-4, errs


# ---- the C minibatch loop (modl_partial_fit_*: two streams, fused subset statistics) --------------------
LOOP_CASES = [
    dict(code_l1_ratio=1., code_alpha=0.5, reduction=4),
    dict(code_l1_ratio=0., code_alpha=0.1, reduction=3, comp_l1_ratio=1.),
    dict(code_l1_ratio=0.5, code_alpha=0.3, reduction=2, Dx_agg='average', G_agg='average'),
    dict(code_l1_ratio=1., code_alpha=0.5, reduction=5, G_agg='full', Dx_agg='full'),
    dict(code_l1_ratio=1., code_alpha=0.5, reduction=2, code_pos=True, comp_pos=True, Dx_agg='full', G_agg='masked'),
    dict(code_l1_ratio=1., code_alpha=0.5, reduction=1, optimizer='sgd', step_size=0.1),
    dict(code_l1_ratio=1., code_alpha=0.5, reduction=6, rand_size=False, replacement=False),
]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("case", range(len(LOOP_CASES)))
def test_c_loop_equals_per_batch_walk(dt, case):
    """`partial_fit` through ONE C call (loop, host bookkeeping, two streams, subset statistics folded straight into
    C_ and the B_[:, subset] panel) against the per-batch Python walk over `modl_batch_fit_*` on one stream: same
    integer bookkeeping bit for bit, same state up to the rounding of two equivalent products."""
    from modl_b200 import DictFact
    kw = LOOP_CASES[case]
    n, p, k, b = 460, 700, 24, 96                       # ragged last batch (76 rows)
    X = _planted(n, p, k, seed=case).astype(dt)
    if kw.get("code_pos"):
        X = np.abs(X)
    rng = np.random.RandomState(1)
    idx = rng.permutation(n)                            # scattered rows of the per-sample state
    ests = []
    for python_loop in (False, True):
        est = DictFact(n_components=k, batch_size=b, random_state=0, **kw)
        est.python_loop = python_loop
        est.prepare(n_samples=n, X=X[:k])
        est.partial_fit(X, idx)                         # non-contiguous indices, 5 batches
        est.partial_fit(X[:2 * b], np.arange(2 * b))    # contiguous indices
        est.partial_fit(X[b:2 * b + 7])                 # no indices
        ests.append(est)
    a, c = ests
    # float32: the two paths round the B_[:, subset] panel differently (1e-7), which can tip a coordinate-descent stop test
    tol = 5e-4 if dt == np.float32 else 1e-9
    assert a.n_iter_ == c.n_iter_
    np.testing.assert_array_equal(a.sample_n_iter_, c.sample_n_iter_)
    np.testing.assert_array_equal(a.last_subset_, c.last_subset_)
    np.testing.assert_array_equal(a.last_order_, c.last_order_)
    for name in ("components_", "code_", "C_", "B_", "comp_norm_"):
        e = rel_err(getattr(a, name), getattr(c, name))
        assert e < tol, (name, kw, e)
    if kw.get("G_agg") == 'full':
        assert rel_err(a.G_, c.G_) < 10 * tol
    if kw.get("G_agg") == 'average':
        assert rel_err(a.G_average_, c.G_average_) < tol
        assert rel_err(a.Dx_average_, c.Dx_average_) < tol


@pytest.mark.parametrize("source", ["numpy", "pinned", "pinned_async", "cuda"])
def test_c_loop_host_rows_and_code_out(source):
    """Host rows (pageable through the bounce buffers, pinned straight from the caller's memory, asynchronous) and the
    per-batch code read-back: same result as device rows, and code_out holds exactly code_[sample_indices]."""
    from modl_b200 import DictFact
    n, p, k, b = 700, 900, 32, 100
    X = _planted(n, p, k, seed=3)
    kw = dict(n_components=k, batch_size=b, reduction=3, code_l1_ratio=1., code_alpha=0.4, random_state=0)
    half = 3 * b + 17
    Xd = torch.from_numpy(X).cuda()
    want = DictFact(**kw)
    want.prepare(n_samples=n, X=X[:k])
    want.partial_fit(Xd[:half], np.arange(half))
    want.partial_fit(Xd[half:], np.arange(half, n))
    est = DictFact(async_host_copy=(source == "pinned_async"), **kw)
    est.prepare(n_samples=n, X=X[:k])
    rows = {"numpy": X, "cuda": torch.from_numpy(X).cuda()}.get(source)
    if rows is None:
        rows = torch.from_numpy(X).pin_memory()
    out = torch.zeros((n, k), dtype=torch.float32).pin_memory()
    est.partial_fit(rows[:half], np.arange(half), code_out=out[:half])       # several calls, ragged tails
    est.partial_fit(rows[half:], np.arange(half, n), code_out=out[half:])
    est.synchronize()
    np.testing.assert_array_equal(est.code_, want.code_)
    np.testing.assert_array_equal(est.components_, want.components_)
    np.testing.assert_array_equal(est.B_, want.B_)
    # a row's code is final when its batch has been solved (no row is visited twice here)
    np.testing.assert_array_equal(out.numpy(), est.code_)


@pytest.mark.parametrize("option", ["overlap", "gate"])
def test_c_loop_schedules_agree(option):
    """One stream vs two streams, and the second stream with / without waiting for the dictionary kernel to be
    resident: scheduling choices only, bit-identical state (the overlapped schedule uses the same kernels on the
    subset statistics whichever stream runs the full-width product)."""
    from modl_b200 import DictFact
    n, p, k, b = 600, 2000, 64, 120
    X = torch.from_numpy(_planted(n, p, k, seed=4)).cuda()
    outs = []
    for value in (1, 0):
        est = DictFact(n_components=k, batch_size=b, reduction=4, code_l1_ratio=1., code_alpha=0.4, random_state=0)
        est.prepare(n_samples=n, X=X[:k])
        est._fit_loop_handle().set_option(option, value)
        est.partial_fit(X)
        outs.append(est)
    a, c = outs
    tol = 0 if option == "gate" else 2e-5
    for name in ("components_", "code_", "C_", "B_", "comp_norm_"):
        if tol == 0:
            np.testing.assert_array_equal(getattr(a, name), getattr(c, name))
        else:
            assert rel_err(getattr(a, name), getattr(c, name)) < tol, name


@pytest.mark.parametrize("mode", [2, 1])
@pytest.mark.parametrize("rows", ["device", "pinned"])
def test_c_loop_graph_replay_is_bit_identical(rows, mode):
    """From its fifth block on the loop replays a block as CUDA graphs updated in place (modl_fit_set_option "graph":
    2 = one graph per block with the full-width product forked inside it, 1 = the three calls of the two-stream
    schedule as three graphs): same kernels, same inputs, so every state array, the codes read back and the bookkeeping
    are bit-identical to plain stream launches -- across calls, ragged tails and host rows."""
    from modl_b200 import DictFact
    n, p, k, b = 1900, 2000, 64, 120                  # 16 blocks, the last one ragged
    Xh = torch.from_numpy(_planted(n, p, k, seed=6))
    X = Xh.cuda() if rows == "device" else Xh.pin_memory()
    outs, codes, stats = [], [], []
    for value in (mode, 0):
        est = DictFact(n_components=k, batch_size=b, reduction=4, code_l1_ratio=1., code_alpha=0.4, random_state=0,
                       async_host_copy=True)
        est.prepare(n_samples=n, X=Xh[:k].numpy())
        loop = est._fit_loop_handle()
        loop.set_option("graph", value)
        out = torch.empty((n, k), dtype=torch.float32).pin_memory()
        cuts = [0, 120, 240, 360, 480, 600, 720, 1320, n]         # one block per call, then several, then the tail
        for r0, r1 in zip(cuts[:-1], cuts[1:]):
            est.partial_fit(X[r0:r1], np.arange(r0, r1), code_out=out[r0:r1])
        est.synchronize()
        codes.append(out.numpy().copy())
        # the scratch slots are given back in mid-fit: they must grow while a call is being captured -> those calls fall
        # back to plain launches (MODL_EGROW), the next ones replay again
        before = loop.graph_stats()
        est._ctx().set_option("drop_workspace", 1)
        est.partial_fit(X[:4 * b], np.arange(4 * b))
        est.synchronize()
        outs.append(est)
        stats.append(loop.graph_stats())
        if value:
            assert stats[-1]["fallbacks"] > before["fallbacks"], (before, stats[-1])
            assert stats[-1]["launches"] > before["launches"], (before, stats[-1])
    assert stats[1]["launches"] == 0
    assert stats[0]["gate_timeouts"] == 0, stats[0]
    assert stats[0]["launches"] >= (4 - mode) * (16 - 4) - stats[0]["fallbacks"] > 0, stats[0]
    a, c = outs
    for name in ("components_", "code_", "C_", "B_", "comp_norm_", "sample_n_iter_", "last_subset_", "last_order_"):
        np.testing.assert_array_equal(getattr(a, name), getattr(c, name), err_msg=name)
    assert a.n_iter_ == c.n_iter_
    np.testing.assert_array_equal(codes[0], codes[1])


def test_refit_with_another_dtype():
    """prepare() drops every dtype- / device-bound cache: the same estimator fits float32 then float64 data."""
    from modl_b200 import DictFact
    X = _planted(300, 400, 12, seed=5)
    est = DictFact(n_components=12, batch_size=50, reduction=2, random_state=0, code_alpha=0.3)
    for python_loop in (True, False):
        est.python_loop = python_loop
        est.random_state = 0
        est.fit(X)
        d32 = est.components_.copy()
        assert d32.dtype == np.float32
        est.random_state = 0
        est.fit(X.astype(np.float64))
        assert est.components_.dtype == np.float64
        assert rel_err(est.components_, d32) < 5e-3


def test_long_run_drift_at_config2(reference):
    """50 minibatches at the BASELINE config-2 shape (k=256, p=10000, batch=512, reduction=8, l1 codes) against the
    UNMODIFIED reference:
      (a) per step, from the reference's state: the batch code within 1e-4 relative (north_star), and the number of
          samples whose code differs by more than 1e-4 (a coordinate-descent stop decision that fell the other way);
      (b) free-running for the 50 steps: the cumulative drift of the dictionary and of the surrogate statistics.
    Bounds asserted below are what DESIGN.md states."""
    from modl_b200 import DictFact
    steps, p, k, b = 50, 10000, 256, 512
    n = steps * b
    X = _planted(n, p, k, seed=7)
    kw = dict(n_components=k, batch_size=b, reduction=8, code_l1_ratio=1., code_alpha=1., tol=1e-2, max_iter=100,
              random_state=0)
    ref = reference.DictFact(**kw)
    ref.prepare(n_samples=n, X=X[:k])
    free = DictFact(**kw)
    free.prepare(n_samples=n, X=X[:k])
    same = DictFact(**kw)                       # re-seated on the reference's state before every step
    same.prepare(n_samples=n, X=X[:k])
    Xd = torch.from_numpy(X).cuda()
    worst_code, worst_step, flipped_total, d_same_worst = 0., -1, 0, 0.
    for t in range(steps):
        sl = slice(t * b, (t + 1) * b)
        idx = np.arange(sl.start, sl.stop)
        for name in ("components_", "C_", "B_", "comp_norm_"):
            setattr(same, name, getattr(ref, name))
        same.code_dev[sl] = torch.from_numpy(np.ascontiguousarray(ref.code_[sl])).cuda()
        ref.partial_fit(X[sl], idx)
        same.partial_fit(Xd[sl], idx)
        free.partial_fit(Xd[sl], idx)
        np.testing.assert_array_equal(same.last_subset_, free.last_subset_)
        got, want = same.code_dev[sl].cpu().numpy().astype(np.float64), ref.code_[sl].astype(np.float64)
        e = rel_err(got, want)
        per_sample = np.linalg.norm(got - want, axis=1) / np.maximum(np.linalg.norm(want, axis=1), 1e-30)
        flipped_total += int((per_sample > 1e-4).sum())
        if e > worst_code:
            worst_code, worst_step = e, t
        d_same_worst = max(d_same_worst, rel_err(same.components_, ref.components_))
    drift = {name: rel_err(getattr(free, name), getattr(ref, name)) for name in ("components_", "C_", "B_", "code_")}
    print("drift over %d steps at config 2: worst per-step code error from the same state %.3g (step %d); samples beyond "
          "1e-4: %d of %d; worst one-step dictionary error %.3g; free-running drift %s"
          % (steps, worst_code, worst_step, flipped_total, steps * b, d_same_worst,
             {k_: "%.3g" % v for k_, v in drift.items()}))
    assert free.n_iter_ == ref.n_iter_
    assert worst_code < 1e-4
    assert d_same_worst < 1e-4
    assert flipped_total <= steps * b // 200
    assert drift["components_"] < 2e-3 and drift["B_"] < 2e-3 and drift["C_"] < 2e-3
