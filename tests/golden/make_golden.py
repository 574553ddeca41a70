"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container, where /root/reference exists:
    python oracle/build_ref.py && python tests/golden/make_golden.py
It imports the reference compiled into oracle/_ref/ (never the oracle restatement,
never modl_b200) and records inputs -> outputs of the hot-path functions for fixed
seeds.  The .npz files are committed; the GPU box has no /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from modl.decomposition.dict_fact import DictFact  # noqa: E402
from modl.decomposition.dict_fact_fast import (  # noqa: E402
    _batch_weight, _enet_regression_multi_gram, _enet_regression_single_gram,
    _update_G_average)
from modl.utils.math.enet import enet_norm, enet_projection, enet_scale  # noqa: E402
from modl.utils.randomkit import RandomState, Sampler  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024.))


def gold_rng():
    out = {}
    for seed in (0, 42, 2 ** 35 + 11):
        rs = RandomState(seed)
        out["randint10_%d" % seed] = np.array([rs.randint(10) for _ in range(64)])
        out["randint_big_%d" % seed] = np.array([rs.randint(2 ** 40) for _ in range(32)])
        for n, p in ((1000, 0.8), (100, 0.1), (10000, 0.125), (200000, 1. / 12), (50, 0.3)):
            out["binom_%d_%d_%g" % (seed, n, p)] = np.array(
                [rs.binomial(n, p) for _ in range(48)])
        out["perm_%d" % seed] = np.asarray(rs.permutation(37)).copy()
        a = np.arange(23)
        b = np.arange(22, -1, -1)
        out["trace_%d" % seed] = rs.shuffle_with_trace([a, b])
        out["trace_a_%d" % seed] = a
        out["trace_b_%d" % seed] = b
    for rand_size in (0, 1):
        for repl in (0, 1):
            for rng_, red in ((100, 10), (1000, 8), (57, 3.5), (20, 1), (20, 2)):
                s = Sampler(rng_, bool(rand_size), bool(repl), 5)
                subs = [np.asarray(s.yield_subset(red)).copy() for _ in range(40)]
                tag = "samp_%d_%d_%d_%g" % (rand_size, repl, rng_, red)
                out[tag + "_len"] = np.array([len(x) for x in subs])
                out[tag + "_cat"] = np.concatenate(subs) if subs else np.zeros(0, int)
    out["batch_weight"] = np.array(
        [_batch_weight(c, b, lr, off) for c, b, lr, off in
         ((512, 512, 1., 0.), (1024, 512, .9, 0.), (30, 10, .95, 0.), (5000, 7, .92, 0.), (70, 10, 1., 3.))])
    save("rng.npz", **out)


def gold_enet():
    rng = np.random.RandomState(0)
    out = {}
    for dt in (np.float32, np.float64):
        v = rng.randn(6, 300).astype(dt)
        v[3] *= 0.01
        out["v_%s" % dt.__name__] = v
        for l1 in (0., 0.15, 0.5, 1.):
            for radius in (0.5, 1., 40.):
                res = np.zeros_like(v)
                for i in range(v.shape[0]):
                    enet_projection(v[i], res[i], radius, l1)
                out["proj_%s_%g_%g" % (dt.__name__, l1, radius)] = res
            out["norm_%s_%g" % (dt.__name__, l1)] = np.array(
                [enet_norm(v[i], l1) for i in range(v.shape[0])], dtype=dt)
            sc = v.copy()
            for i in range(v.shape[0]):
                enet_scale(sc[i], l1, 1.)
            out["scale_%s_%g" % (dt.__name__, l1)] = sc
    save("enet.npz", **out)


def gold_regression():
    rng = np.random.RandomState(1)
    out = {}
    k, b, p, n = 48, 20, 150, 30
    for dt in (np.float32, np.float64):
        D = (rng.randn(k, p) / np.sqrt(p)).astype(dt)
        X = rng.randn(b, p).astype(dt)
        G = np.ascontiguousarray(D @ D.T)
        Dx = np.ascontiguousarray(X @ D.T)
        idx = rng.permutation(n)[:b].astype(np.int64)
        code0 = (rng.randn(n, k) * (rng.rand(n, k) < 0.5)).astype(dt)
        Gm = (np.tile(G, (b, 1, 1)) * (1 + 0.02 * np.arange(b)[:, None, None])).astype(dt)
        w_s = rng.rand(b).astype(dt)
        tag = dt.__name__
        out.update({"G_" + tag: G, "Dx_" + tag: Dx, "X_" + tag: X, "idx_" + tag: idx,
                    "code0_" + tag: code0, "Gm_" + tag: Gm, "ws_" + tag: w_s})
        cases = ((1., .1, False, 1e-3), (1., .05, True, 1e-2), (.5, .05, False, 1e-3),
                 (.9, .3, False, 1e-2), (0., .1, False, 1e-2))
        for ci, (l1, alpha, pos, tol) in enumerate(cases):
            c = code0.copy()
            _enet_regression_single_gram(G, Dx.copy(), X, c, idx, l1, alpha, pos, tol, 100)
            out["single_%s_%d" % (tag, ci)] = c
            c = code0.copy()
            _enet_regression_multi_gram(Gm.copy(), Dx.copy(), X, c, idx, l1, alpha, pos, tol, 100)
            out["multi_%s_%d" % (tag, ci)] = c
        out["cases"] = np.array([(l1, a, float(pos), tol) for l1, a, pos, tol in cases])
        ga = Gm.copy()
        _update_G_average(ga, G, w_s)
        out["gavg_" + tag] = ga
    save("regression.npz", **out)


def gold_update_dict():
    """One dictionary BCD step with hand-set statistics (dict_fact.py:650-715)."""
    out = {}
    k, p, n = 24, 120, 40
    cases = (dict(comp_l1_ratio=0., comp_pos=False), dict(comp_l1_ratio=1., comp_pos=False),
             dict(comp_l1_ratio=.3, comp_pos=False), dict(comp_l1_ratio=0., comp_pos=True),
             dict(comp_l1_ratio=.5, comp_pos=True, G_agg='full'))
    for dt in (np.float32, np.float64):
        for ci, kw in enumerate(cases):
            rng = np.random.RandomState(10 + ci)
            X = rng.randn(n, p).astype(dt)
            est = DictFact(n_components=k, random_state=3, reduction=3, **kw)
            est.prepare(n_samples=n, X=X)
            A = rng.randn(60, k) * (rng.rand(60, k) < 0.6)
            est.C_[:] = (A.T @ A / 60).astype(dt)
            est.C_[5, :] = 0
            est.C_[:, 5] = 0          # exercises the C[k,k] <= 1e-20 skip rule
            est.B_[:] = (A.T @ (A @ rng.randn(k, p) * .3 + rng.randn(60, p)) / 60).astype(dt)
            est.comp_norm_[:] = (rng.rand(k) * .05).astype(dt)
            subset = np.asarray(est.feature_sampler_.yield_subset(3)).copy()
            est.gradient_[:, subset] = est.B_[:, subset]
            state = est.random_state.get_state()
            order = np.random.RandomState(0)
            order.set_state(state)
            order = order.permutation(k)
            tag = "%s_%d" % (dt.__name__, ci)
            out.update({"D0_" + tag: est.components_.copy(), "C_" + tag: est.C_.copy(),
                        "B_" + tag: est.B_.copy(), "norm0_" + tag: est.comp_norm_.copy(),
                        "subset_" + tag: subset, "order_" + tag: order})
            if kw.get('G_agg') == 'full':
                out["G0_" + tag] = est.G_.copy()
            est._update_dict(subset, 0.5)
            out["D1_" + tag] = est.components_.copy()
            out["norm1_" + tag] = est.comp_norm_.copy()
            if kw.get('G_agg') == 'full':
                out["G1_" + tag] = est.G_.copy()
    out["cases"] = np.array([(c['comp_l1_ratio'], float(c['comp_pos']), float(c.get('G_agg') == 'full'))
                             for c in cases])
    save("update_dict.npz", **out)


FIT_CASES = (
    dict(reduction=2, batch_size=16, n_epochs=2),
    dict(reduction=3, batch_size=16, n_epochs=2, code_l1_ratio=0., comp_l1_ratio=1., code_alpha=.1),
    dict(reduction=2, batch_size=16, n_epochs=2, comp_pos=True, code_pos=True, code_alpha=.01),
    dict(reduction=2, batch_size=16, n_epochs=2, G_agg='full', Dx_agg='full', comp_l1_ratio=.5),
    dict(reduction=2, batch_size=16, n_epochs=2, G_agg='average', Dx_agg='average'),
    dict(reduction=2, batch_size=16, n_epochs=2, G_agg='full', Dx_agg='average',
         rand_size=False, replacement=False),
    dict(batch_size=16, n_epochs=2, optimizer='sgd', step_size=.1),
    dict(reduction=1, batch_size=10, n_epochs=1, code_l1_ratio=0.),
)


def fit_data(dt):
    rng = np.random.RandomState(7)
    n, p, k = 96, 40, 6
    X = (rng.randn(n, k) @ rng.randn(k, p) + 0.1 * rng.randn(n, p)).astype(dt)
    return X, k


def gold_fit():
    """Whole fits on a tiny problem: final state after n_epochs (dict_fact.py:286-311)."""
    out = {}
    for dt in (np.float32, np.float64):
        X, k = fit_data(dt)
        for ci, kw in enumerate(FIT_CASES):
            est = DictFact(n_components=k, random_state=0, **kw).fit(X)
            tag = "%s_%d" % (dt.__name__, ci)
            out["D_" + tag] = est.components_
            out["code_" + tag] = est.code_
            out["C_" + tag] = est.C_
            out["B_" + tag] = est.B_
            out["norm_" + tag] = est.comp_norm_
            out["labels_" + tag] = est.labels_
            out["sni_" + tag] = est.sample_n_iter_
            out["T_" + tag] = est.transform(X)
            out["score_" + tag] = np.array(est.score(X))
            if hasattr(est, 'G_'):
                out["G_" + tag] = est.G_
    save("fit.npz", **out)


def gold_image():
    """ImageDictFact / LazyCleanPatchExtractor / scale_patches of the reference.
    The reference's patch extractor imports `sklearn.feature_extraction.image.extract_patches`, which newer
    scikit-learn only ships as `_extract_patches`; the alias below is the one-line shim that lets the
    UNMODIFIED reference modules import."""
    import json
    import sklearn.feature_extraction.image as skimg
    if not hasattr(skimg, "extract_patches"):
        skimg.extract_patches = skimg._extract_patches
    from modl.decomposition.image import ImageDictFact
    from modl.feature_extraction.image import LazyCleanPatchExtractor
    from modl.input_data.image import scale_patches

    rng = np.random.RandomState(3)
    img_a = rng.rand(26, 22, 3)                               # clean, float64
    img_b = rng.rand(20, 20, 6)                               # missing values, 6 channels > patch width 4
    for (r, c, ch) in ((3, 4, 0), (10, 11, 2), (15, 2, 3), (7, 16, 5), (18, 18, 4)):
        img_b[r, c, ch] = -1
    out = {"img_a": img_a, "img_b": img_b}
    # extractor: selection, shuffles, patches
    for tag, img, kw in (("a", img_a, dict(patch_size=(4, 4), max_patches=None)),
                         ("b", img_b, dict(patch_size=(4, 4), max_patches=120)),
                         ("c", img_a, dict(patch_size=None, max_patches=50))):
        ex = LazyCleanPatchExtractor(random_state=5, **kw).fit(img.copy())
        out["ex_%s_idx" % tag] = np.asarray(ex.indices_3d).copy()
        out["ex_%s_first" % tag] = np.asarray(ex.partial_transform(batch=7)).copy()
        ex.shuffle()
        out["ex_%s_idx_shuffled" % tag] = np.asarray(ex.indices_3d).copy()
        out["ex_%s_slice" % tag] = np.asarray(ex.partial_transform(batch=slice(3, 12))).copy()
    pat = np.asarray(LazyCleanPatchExtractor(random_state=5, patch_size=(4, 4)).fit(img_a.copy()).partial_transform(batch=9))
    for mean in (True, False):
        out["scaled_mean%d" % mean] = scale_patches(pat, with_mean=mean, with_std=True, copy=True)
    # whole fits
    cases = [
        dict(img="a", kw=dict(method="masked", setting="dictionary learning", reduction=2, n_epochs=2)),
        dict(img="a", kw=dict(method="gram", setting="NMF", reduction=3, n_epochs=2)),
        dict(img="a", kw=dict(method="reducing ratio", setting="dictionary learning", reduction=4, n_epochs=3)),
        dict(img="a", kw=dict(method="average", setting="dictionary learning", reduction=2, n_epochs=2)),
        dict(img="a", kw=dict(method="dictionary only", setting="NMF", reduction=2, n_epochs=1)),
        dict(img="a", kw=dict(method="sgd", setting="dictionary learning", n_epochs=1, step_size=1e-2)),
        dict(img="b", kw=dict(method="masked", setting="dictionary learning", reduction=2, n_epochs=2, max_patches=150)),
        dict(img="a", dtype="float32", kw=dict(method="masked", setting="dictionary learning", reduction=2, n_epochs=2,
                                                buffer_size=64)),
    ]
    common = dict(patch_size=(4, 4), batch_size=10, n_components=8, alpha=0.1, random_state=0)
    for ci, case in enumerate(cases):
        img = out["img_" + case["img"]].astype(case.get("dtype", "float64"))
        est = ImageDictFact(**common, **case["kw"]).fit(img.copy())
        out["fit_%d_components" % ci] = est.components_.copy()
        out["fit_%d_n_iter" % ci] = np.array(est.n_iter_)
        test = np.asarray(LazyCleanPatchExtractor(random_state=9, patch_size=(4, 4), max_patches=30).fit(img.copy()).transform())
        out["fit_%d_test" % ci] = test
        out["fit_%d_code" % ci] = est.transform(test)
        out["fit_%d_score" % ci] = np.array(est.score(test))
    out["cases"] = np.array(json.dumps(dict(common=common, cases=cases)))
    save("image.npz", **out)


def _stub_neuroimaging_stack():
    """`modl/decomposition/fmri.py` imports nibabel, nilearn and `sklearn.externals.joblib`, none of which this
    image has.  The fMRI row is the learning loop `_compute_components` (fmri.py:423-546),
    which only needs: `ImageFileError` (raised by `check_niimg` for a `.npy` record so that `_lazy_scan` takes
    its `np.load` branch, fmri.py:564-572), `check_niimg(mask).get_data()`, and a fitted masker with
    `transform(img, confounds)`.  The stubs below provide exactly that much; fmri.py itself runs UNMODIFIED."""
    import types
    import joblib

    class ImageFileError(Exception):
        pass

    class MaskImg(object):
        def __init__(self, mask):
            self._mask = np.asarray(mask)

        def get_data(self):
            return self._mask

    def check_niimg(img):
        if isinstance(img, MaskImg):
            return img
        raise ImageFileError("not a Niimg: %r" % (img,))

    class CacheMixin(object):
        def _cache(self, func, **kwargs):
            return func

    class Memory(object):
        def __init__(self, *args, **kwargs):
            pass

    class NiftiMasker(object):
        """Masker over pre-masked records: a record is a `.npy` file (n_samples, n_voxels)."""
        def __init__(self, mask_img=None, **kwargs):
            self.mask_img = mask_img

        def fit(self, imgs=None):
            self.mask_img_ = self.mask_img
            return self

        def _check_fitted(self):
            assert hasattr(self, "mask_img_")

        def transform(self, img, confounds=None):
            return np.load(img)

    class BaseNilearnEstimator(object):
        pass

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("nibabel"); mod("nibabel.filebasedimages", ImageFileError=ImageFileError)
    mod("nilearn"); mod("nilearn._utils", CacheMixin=CacheMixin, check_niimg=check_niimg)
    mod("nilearn.input_data", NiftiMasker=NiftiMasker)
    import sklearn
    ext = mod("sklearn.externals")
    ext.joblib = mod("sklearn.externals.joblib", Memory=Memory, Parallel=joblib.Parallel, delayed=joblib.delayed)
    sklearn.externals = ext
    mod("modl.input_data.fmri"); mod("modl.input_data.fmri.base", BaseNilearnEstimator=BaseNilearnEstimator)
    return MaskImg, NiftiMasker


FMRI_CASES = [
    dict(dtype="float64", kw=dict(method="masked", reduction=3, n_epochs=2, alpha=0.05)),
    dict(dtype="float64", kw=dict(method="gram", reduction=2, n_epochs=7, alpha=0.05)),          # switches at epoch 5
    dict(dtype="float64", kw=dict(method="reducing ratio", reduction=4, n_epochs=3, alpha=0.05)),
    dict(dtype="float64", kw=dict(method="average", reduction=2, n_epochs=2, alpha=0.05)),
    dict(dtype="float64", kw=dict(method="dictionary only", reduction=2, n_epochs=1, alpha=0.05, positive=True)),
    dict(dtype="float64", kw=dict(method="sgd", n_epochs=1, alpha=0.05, step_size=1e-2)),
    dict(dtype="float64", init=True, kw=dict(method="masked", reduction=2, n_epochs=1, alpha=0.02, n_components=9)),
    dict(dtype="float32", kw=dict(method="masked", reduction=3, n_epochs=2, alpha=0.05)),
]


def fmri_records(dtype):
    """Four 'subjects' of pre-masked records with ragged lengths, from spatially sparse planted maps."""
    rng = np.random.RandomState(11)
    n_voxels, k = 150, 5
    maps = rng.randn(k, n_voxels) * (rng.rand(k, n_voxels) < 0.25)
    lengths = (23, 17, 30, 20)
    return [(rng.randn(n, k) @ maps + 0.05 * rng.randn(n, n_voxels)).astype(dtype) for n in lengths], n_voxels


def gold_fmri():
    """`_compute_components` + `_flip` of the reference (fmri.py:423-556) on `.npy` records."""
    import json
    import tempfile
    MaskImg, NiftiMasker = _stub_neuroimaging_stack()
    from modl.decomposition.fmri import _compute_components, _flip
    from modl.decomposition.dict_fact import Coder
    out = {}
    common = dict(n_components=6, batch_size=8, learning_rate=0.9, random_state=0)
    for ci, case in enumerate(FMRI_CASES):
        records, n_voxels = fmri_records(case["dtype"])
        masker = NiftiMasker(mask_img=MaskImg(np.ones(n_voxels, dtype=bool))).fit()
        with tempfile.TemporaryDirectory() as tmp:
            paths = []
            for ri, rec in enumerate(records):
                paths.append(os.path.join(tmp, "record_%d.npy" % ri))
                np.save(paths[-1], rec)
            kw = dict(common)
            kw.update(case["kw"])
            dict_init = None
            if case.get("init"):
                dict_init = np.random.RandomState(5).randn(7, n_voxels)     # fewer atoms than n_components asked
            comp = _compute_components(masker, paths, dict_init=dict_init, **kw)
        out["fit_%d_components" % ci] = comp
        coder = Coder(dictionary=comp, code_alpha=kw["alpha"], code_l1_ratio=0).fit()
        out["fit_%d_code" % ci] = coder.transform(records[0])
        out["fit_%d_score" % ci] = np.array(coder.score(records[0]))
    v = np.random.RandomState(2).randn(6, 31)
    v[3] = 0
    out["flip_in"], out["flip_out"] = v, _flip(v)
    out["cases"] = np.array(json.dumps(dict(common=common, cases=FMRI_CASES)))
    save("fmri.npz", **out)


RECSYS_CASES = [
    # the reference's own test problems [ref: modl/decomposition/tests/test_recsys.py:13-35, :38-62, :65-92]
    dict(data="dense", dtype="float64", kw=dict(n_components=3, n_epochs=1, alpha=1e-3, detrend=False)),
    dict(data="dense", dtype="float64", kw=dict(n_components=3, n_epochs=1, alpha=1e-3, detrend=True)),
    dict(data="missing", dtype="float64", kw=dict(n_components=4, n_epochs=1, alpha=1, detrend=True)),
    # minibatches, several epochs, forgetting, cropping, automatic batch size
    dict(data="ratings", dtype="float64", kw=dict(n_components=5, n_epochs=3, alpha=0.5, batch_size=7,
                                                   learning_rate=0.8, detrend=True, crop=(1., 5.), beta=2.)),
    dict(data="ratings", dtype="float64", kw=dict(n_components=6, n_epochs=2, alpha=0.1, batch_size=None)),
    dict(data="ratings", dtype="float32", kw=dict(n_components=5, n_epochs=2, alpha=1., batch_size=16,
                                                   learning_rate=0.9)),
]


def recsys_data(name, dtype):
    import scipy.sparse as sp
    rng = np.random.RandomState(0)
    if name == "dense":
        X = sp.csr_matrix(np.dot(rng.rand(50, 3), rng.rand(3, 20)))
        return X.astype(dtype), X.astype(dtype)
    if name == "missing":
        X = np.dot(rng.rand(100, 4), rng.rand(4, 20))
        keep = rng.rand(100, 20) < 0.85
        keep[np.arange(100), rng.randint(20, size=100)] = True          # no empty row
        return sp.csr_matrix(X * keep).astype(dtype), sp.csr_matrix(X * ~keep).astype(dtype)
    # ratings-like: 120 users x 60 items, 1..5, 25 % observed, every user rates at least one item
    U, V = rng.rand(120, 4), rng.rand(4, 60)
    R = np.clip(np.round(1 + 4 * U.dot(V) / 2.), 1, 5)
    seen = rng.rand(120, 60) < 0.25
    seen[np.arange(120), rng.randint(60, size=120)] = True
    test = ~seen & (rng.rand(120, 60) < 0.1)
    return sp.csr_matrix(R * seen).astype(dtype), sp.csr_matrix(R * test).astype(dtype)


def gold_recsys():
    """RecsysDictFact of the reference (recsys.py), whole fits + predictions."""
    import json
    from modl.decomposition.recsys import RecsysDictFact, compute_biases
    out = {}
    for ci, case in enumerate(RECSYS_CASES):
        X, X_te = recsys_data(case["data"], case["dtype"])
        kw = dict(case["kw"])
        if "crop" in kw:
            kw["crop"] = tuple(kw["crop"])
        est = RecsysDictFact(random_state=0, **kw).fit(X)
        tag = "fit_%d_" % ci
        for name in ("components_", "code_", "C_", "B_", "comp_norm_", "feature_n_iter_", "feature_freq_"):
            out[tag + name] = np.asarray(getattr(est, name))
        out[tag + "n_iter_"] = np.array(est.n_iter_)
        if kw.get("detrend"):
            out[tag + "row_mean_"], out[tag + "col_mean_"] = est.row_mean_, est.col_mean_
        if case["dtype"] == "float64":          # the reference's _predict only takes float64 buffers
            out[tag + "pred_train"] = est.predict(X.astype(np.float64)).data
            out[tag + "pred_test"] = est.predict(X_te.astype(np.float64)).data
            out[tag + "score_test"] = np.array(est.score(X_te.astype(np.float64)))
    X, _ = recsys_data("ratings", "float64")
    out["bias_row"], out["bias_col"] = compute_biases(X, beta=3., inplace=False)
    out["cases"] = np.array(json.dumps(RECSYS_CASES))
    save("recsys.npz", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "image":
        gold_image()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "recsys":
        gold_recsys()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fmri":
        gold_fmri()
        sys.exit(0)
    gold_rng()
    gold_enet()
    gold_regression()
    gold_update_dict()
    gold_fit()
    gold_image()
    gold_fmri()
    gold_recsys()
