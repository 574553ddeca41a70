"""fMRIDictFact front-end against golden vectors produced by the UNMODIFIED reference's
`_compute_components` (tests/golden/make_golden.py::gold_fmri; nilearn / nibabel stubbed there, fmri.py itself
untouched).  On CPU the learning loop's HOST logic -- one shared RandomState feeding the sampler seed, the
record order, the row permutations and the atom orders, the epoch schedules, n_samples = total + 1, the sign
flip -- is checked with the CPU oracle standing in for the device estimator (tests may do that; the product
never does).  On the GPU the same cases run through the real device path."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import rel_err

torch = pytest.importorskip("torch")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def gold(golden):
    return golden("fmri.npz")


def _records(dtype):
    """The generator's records, regenerated (same seeded construction as make_golden.fmri_records)."""
    rng = np.random.RandomState(11)
    n_voxels, k = 150, 5
    maps = rng.randn(k, n_voxels) * (rng.rand(k, n_voxels) < 0.25)
    lengths = (23, 17, 30, 20)
    return [(rng.randn(n, k) @ maps + 0.05 * rng.randn(n, n_voxels)).astype(dtype) for n in lengths], n_voxels


def _run_case(fmri, gold, ci, case, common, tmp_path, device, as_paths):
    records, n_voxels = _records(case["dtype"])
    masker = fmri.RecordMasker(mask=np.ones(n_voxels, dtype=bool)).fit()
    if as_paths:
        imgs = []
        for ri, rec in enumerate(records):
            imgs.append(str(tmp_path / ("case%d_record_%d.npy" % (ci, ri))))
            np.save(imgs[-1], rec)
    else:
        imgs = records
    kw = dict(common)
    kw.update(case["kw"])
    dict_init = np.random.RandomState(5).randn(7, n_voxels) if case.get("init") else None
    comp = fmri._compute_components(masker, imgs, dict_init=dict_init, device=device, **kw)
    want = gold["fit_%d_components" % ci]
    assert comp.shape == want.shape and comp.dtype == want.dtype
    return comp, want, records, kw


def test_flip_and_scan(gold, tmp_path):
    from modl_b200 import fmri
    np.testing.assert_array_equal(fmri._flip(gold["flip_in"]), gold["flip_out"])
    np.testing.assert_array_equal(fmri._flip(torch.from_numpy(gold["flip_in"])).numpy(), gold["flip_out"])
    a, b = np.zeros((5, 7), dtype=np.float64), np.zeros((3, 7), dtype=np.float32)
    np.save(tmp_path / "a.npy", a)
    lengths, dtype = fmri._lazy_scan([str(tmp_path / "a.npy"), b])
    assert lengths == [5, 3] and dtype == np.float32          # dtype of the LAST record [ref: fmri.py:559-575]


def test_record_masker(tmp_path):
    from modl_b200.fmri import RecordMasker
    rng = np.random.RandomState(0)
    mask = rng.rand(4, 5, 3) < 0.5
    vol = rng.randn(9, 4, 5, 3)
    m = RecordMasker(mask=mask).fit()
    flat = m.transform(vol)
    assert flat.shape == (9, int(mask.sum()))
    np.testing.assert_array_equal(flat, vol[:, mask])
    np.testing.assert_array_equal(m.transform(flat), flat)                 # already masked: passes through
    back = m.inverse_transform(flat[:2])
    assert back.shape == (2, 4, 5, 3) and np.all(back[:, ~mask] == 0)
    np.testing.assert_array_equal(back[:, mask], flat[:2])
    with pytest.raises(ValueError):
        m.transform(rng.randn(9, 7))
    with pytest.raises(ValueError):
        RecordMasker()._check_fitted()
    # inferred mask; cleaning
    m2 = RecordMasker(standardize=True, detrend=True).fit([flat])
    assert m2.n_voxels_ == flat.shape[1]
    x = flat + np.linspace(0, 3, 9)[:, None]
    x[:, 0] = 2.0                                                          # constant voxel
    z = m2.transform(x)
    assert np.allclose(z.mean(axis=0), 0, atol=1e-12) and np.all(z[:, 0] == 0)
    assert np.allclose((z[:, 1:] ** 2).mean(axis=0), 1)
    assert np.allclose(np.arange(9) @ z, 0, atol=1e-10)                    # no linear trend left
    with pytest.raises(ValueError):
        RecordMasker().fit([])


def test_estimator_shell_parameters():
    from modl_b200.fmri import fMRICoder, fMRIDictFact
    est = fMRIDictFact(n_components=7, method='gram', reduction=12, random_state=3)
    params = est.get_params()
    for name in ("alpha", "dict_init", "mask", "smoothing_fwhm", "standardize", "detrend", "low_pass", "high_pass",
                 "t_r", "target_affine", "target_shape", "mask_strategy", "mask_args", "memory", "memory_level",
                 "n_jobs", "verbose", "method", "n_epochs", "batch_size", "reduction", "step_size", "positive",
                 "learning_rate", "random_state", "callback", "transform_batch_size"):
        assert name in params, name
    assert params["method"] == 'gram' and params["reduction"] == 12
    est.set_params(reduction=4)
    assert est.reduction == 4
    with pytest.raises(ValueError):
        est.fit(None)
    assert "dictionary" in fMRICoder(np.zeros((2, 3))).get_params()


def test_learning_loop_host_logic_cpu(gold, oracle, tmp_path, monkeypatch):
    """`_compute_components` with the CPU oracle in place of the device estimator reproduces the reference."""
    from modl_b200 import fmri

    class Stand_in(oracle.OracleDictFact):
        def __init__(self, device=None, **kw):
            oracle.OracleDictFact.__init__(self, **kw)

        def partial_fit(self, X, sample_indices=None):
            assert isinstance(X, torch.Tensor)                 # staged + permuted by the front-end
            return oracle.OracleDictFact.partial_fit(self, X.numpy(), sample_indices)

    monkeypatch.setattr(fmri, "DictFact", Stand_in)
    spec = json.loads(str(gold["cases"]))
    for ci, case in enumerate(spec["cases"]):
        comp, want, _, _ = _run_case(fmri, gold, ci, case, spec["common"], tmp_path, "cpu", as_paths=(ci % 2 == 0))
        err = rel_err(comp, want)
        print("case %d %s: components %.3g" % (ci, case["kw"]["method"], err))
        assert err < (1e-9 if case["dtype"] == "float64" else 2e-3), (ci, case, err)


@pytest.mark.gpu
def test_fmri_fits_match_the_reference(gold, tmp_path):
    """Every method of `_compute_components` (ridge codes, L1-ball atoms, positive maps, fewer initial atoms than
    asked, float32) on the device, against the unmodified reference; then the Coder the estimator builds."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from modl_b200 import fmri
    from modl_b200.dict_fact import Coder
    spec = json.loads(str(gold["cases"]))
    for ci, case in enumerate(spec["cases"]):
        comp, want, records, kw = _run_case(fmri, gold, ci, case, spec["common"], tmp_path, None, as_paths=(ci % 2 == 0))
        tol = 1e-9 if case["dtype"] == "float64" else 2e-3
        err = rel_err(comp, want)
        coder = Coder(dictionary=comp, code_alpha=kw["alpha"], code_l1_ratio=0).fit()
        e_code = rel_err(coder.transform(records[0]), gold["fit_%d_code" % ci])
        e_score = abs(coder.score(records[0]) - float(gold["fit_%d_score" % ci])) / abs(float(gold["fit_%d_score" % ci]))
        print("case %d %s: components %.3g, code %.3g, score %.3g" % (ci, case["kw"]["method"], err, e_code, e_score))
        assert err < tol and e_code < 10 * tol and e_score < 10 * tol, (ci, case, err, e_code, e_score)


@pytest.mark.gpu
def test_fmri_estimator_end_to_end(tmp_path):
    """fMRIDictFact.fit / transform / score on volume-shaped records with a 3-D mask: the fitted maps equal the
    learning loop's, land inside the mask, and the length-weighted score matches its definition."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from modl_b200.fmri import RecordMasker, _compute_components, fMRICoder, fMRIDictFact
    rng = np.random.RandomState(4)
    mask = rng.rand(6, 7, 5) < 0.6
    nv = int(mask.sum())
    maps = rng.randn(4, nv) * (rng.rand(4, nv) < 0.3)
    vols = []
    for n in (21, 16, 12):
        flat = rng.randn(n, 4) @ maps + 0.05 * rng.randn(n, nv)
        v = np.zeros((n,) + mask.shape)
        v[:, mask] = flat
        vols.append(v)
    kw = dict(n_components=5, alpha=0.05, method='masked', reduction=2, n_epochs=2, batch_size=8, random_state=1)
    est = fMRIDictFact(mask=mask, standardize=False, detrend=False, **kw).fit(vols)
    want = _compute_components(RecordMasker(mask=mask).fit(), vols, **kw)
    np.testing.assert_array_equal(est.components_, want)
    assert est.components_img_.shape == (5,) + mask.shape and np.all(est.components_img_[:, ~mask] == 0)
    codes = est.transform(vols)
    assert [c.shape for c in codes] == [(21, 5), (16, 5), (12, 5)]
    per = [est.coder_.score(v[:, mask]) for v in vols]
    assert abs(est.score(vols) - np.average(per, weights=[21, 16, 12])) < 1e-12 * abs(per[0])
    # a coder over the learnt maps gives the same loadings
    coder = fMRICoder(est.components_, alpha=0.05, mask=mask).fit()
    np.testing.assert_allclose(coder.transform(vols[1])[0], codes[1], rtol=1e-12, atol=1e-14)
