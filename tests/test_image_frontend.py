"""ImageDictFact front-end against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py::gold_image).  The patch bookkeeping (which patches, in which order, after
which shuffle) is integer work: bit-exact, checked on CPU.  The fits run on the GPU."""
import json

import numpy as np
import pytest

from conftest import rel_err

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gold(golden):
    return golden("image.npz")


EXTRACTOR_CASES = (("a", "img_a", dict(patch_size=(4, 4), max_patches=None)),
                   ("b", "img_b", dict(patch_size=(4, 4), max_patches=120)),      # missing values (-1)
                   ("c", "img_a", dict(patch_size=None, max_patches=50)))         # default patch size


def _check_extractor(gold, device):
    from modl_b200.image import LazyCleanPatchExtractor
    for tag, img, kw in EXTRACTOR_CASES:
        ex = LazyCleanPatchExtractor(random_state=5, device=device, **kw).fit(gold[img].copy())
        np.testing.assert_array_equal(ex.indices_3d, gold["ex_%s_idx" % tag])
        np.testing.assert_array_equal(ex.partial_transform(batch=7).cpu().numpy(), gold["ex_%s_first" % tag])
        ex.shuffle()
        np.testing.assert_array_equal(ex.indices_3d, gold["ex_%s_idx_shuffled" % tag])
        np.testing.assert_array_equal(ex.partial_transform(batch=slice(3, 12)).cpu().numpy(), gold["ex_%s_slice" % tag])
        assert tuple(ex.patch_shape_) == gold["ex_%s_first" % tag].shape[1:]


def test_patch_extractor_bookkeeping_cpu(gold):
    _check_extractor(gold, "cpu")


def test_scale_patches_cpu(gold):
    from modl_b200.image import LazyCleanPatchExtractor, scale_patches
    pat = LazyCleanPatchExtractor(random_state=5, patch_size=(4, 4), device="cpu").fit(gold["img_a"].copy()).partial_transform(batch=9)
    for mean in (True, False):
        got = scale_patches(pat.numpy(), with_mean=mean, with_std=True, copy=True)
        np.testing.assert_allclose(got, gold["scaled_mean%d" % mean], rtol=1e-12, atol=1e-15)
    # a constant patch has zero spread: left at zero, not divided by zero [ref: modl/input_data/image.py:15-16]
    flat = np.ones((2, 4, 4, 3))
    assert np.all(scale_patches(flat, with_mean=True, with_std=True) == 0)


def test_method_and_setting_tables():
    """The lookup tables the reference exposes as class attributes [ref: modl/decomposition/image.py:14-32]."""
    from modl_b200.image import ImageDictFact
    assert ImageDictFact.methods == {'masked': {'G_agg': 'masked', 'Dx_agg': 'masked'},
                                     'dictionary only': {'G_agg': 'full', 'Dx_agg': 'full'},
                                     'gram': {'G_agg': 'masked', 'Dx_agg': 'masked'},
                                     'average': {'G_agg': 'average', 'Dx_agg': 'average'},
                                     'reducing ratio': {'G_agg': 'masked', 'Dx_agg': 'masked'}}
    assert ImageDictFact.settings['NMF'] == {'comp_l1_ratio': 0, 'code_l1_ratio': 1, 'comp_pos': True, 'code_pos': True,
                                             'with_std': True, 'with_mean': False}
    assert ImageDictFact.settings['dictionary learning']['with_mean'] is True


@pytest.mark.gpu
def test_patch_extractor_bookkeeping_gpu(gold):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    _check_extractor(gold, "cuda")


@pytest.mark.gpu
def test_image_fits_match_the_reference(gold):
    """Whole ImageDictFact fits (every method, both settings, missing values, float32) against the state the
    unmodified reference reaches from the same seed: the shared RandomState stream must be consumed in the same
    order (patch selection, sampler seed, atom orders, epoch shuffles), the buffers cut at the same places."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from modl_b200.image import ImageDictFact
    spec = json.loads(str(gold["cases"]))
    for ci, case in enumerate(spec["cases"]):
        dt = case.get("dtype", "float64")
        img = gold["img_" + case["img"]].astype(dt)
        kw = dict(spec["common"])
        kw["patch_size"] = tuple(kw["patch_size"])
        est = ImageDictFact(**kw, **case["kw"]).fit(img.copy())
        tol = 1e-9 if dt == "float64" else 2e-3
        assert est.n_iter_ == int(gold["fit_%d_n_iter" % ci])
        err = rel_err(est.components_, gold["fit_%d_components" % ci])
        test = gold["fit_%d_test" % ci]
        e_code = rel_err(est.transform(test), gold["fit_%d_code" % ci])
        e_score = abs(est.score(test) - float(gold["fit_%d_score" % ci])) / abs(float(gold["fit_%d_score" % ci]))
        print("case %d %s: components %.3g, code %.3g, score %.3g" % (ci, case["kw"]["method"], err, e_code, e_score))
        assert err < tol and e_code < 10 * tol and e_score < 10 * tol, (ci, case, err, e_code, e_score)
