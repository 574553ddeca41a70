"""ImageDictFact -- the reference's image front-end over the B200 hot path
[ref: modl/decomposition/image.py:13-199], with the patch extraction and scaling it needs kept on the
device [ref: modl/feature_extraction/image.py:8-83, modl/input_data/image.py:4-23,
modl/input_data/image_fast.pyx:12-74].

"Next" row 1 of SURVEY section 8f: pure plumbing over `DictFact` -- `method` / `setting` -> estimator
keywords, the first `n_components` patches as the initial dictionary, buffers of `10 x batch_size` scaled
patches streamed into `partial_fit(patches, buffer_slice)`, the per-epoch schedules of the 'gram' and
'reducing ratio' methods -- with the same shared NumPy `RandomState` stream as the reference (patch
selection, sampler seed, atom orders, epoch shuffles all draw from ONE generator, in the reference's order).
The image stays in HBM; a buffer of patches is one gather (`image[p + dy, q + dx, :]`) plus the
per-patch, per-channel centring / scaling, and never visits the host.
"""
import time
from math import sqrt

import numpy as np
import torch
from sklearn.base import BaseEstimator
from sklearn.utils import check_random_state, gen_batches

from ._util import default_device
from .dict_fact import DictFact

__all__ = ["ImageDictFact", "LazyCleanPatchExtractor", "scale_patches", "DictionaryScorer"]


def scale_patches(X, with_mean=True, with_std=True, channel_wise=True, copy=True):
    """Centre / normalise patches (n, h, w, c), per channel by default
    [ref: modl/input_data/image.py:4-23].  torch tensors (any device) or NumPy arrays."""
    numpy_in = not isinstance(X, torch.Tensor)
    if numpy_in:
        X = torch.from_numpy(np.array(X, copy=True))
    elif copy:
        X = X.clone()
    if with_mean:
        if channel_wise:
            X -= X.mean(dim=(1, 2), keepdim=True)
        else:
            X -= X.mean(dim=(1, 2, 3), keepdim=True)
    if with_std:
        if channel_wise:
            n_channel = X.shape[3]
            std = torch.sqrt((X ** 2).sum(dim=(1, 2), keepdim=True))
            std[std == 0] = 1
            X /= std * sqrt(n_channel)
        else:
            std = torch.sqrt((X ** 2).sum(dim=(1, 2, 3), keepdim=True))
            std[std == 0] = 1
            X /= std
    return X.numpy() if numpy_in else X


def _flatten_patches(patches, with_mean=True, with_std=True, copy=False):
    """[ref: modl/decomposition/image.py:193-199]"""
    n_patches = patches.shape[0]
    patches = scale_patches(patches, with_mean=with_mean, with_std=with_std, copy=copy)
    return patches.reshape((n_patches, -1))


class LazyCleanPatchExtractor(BaseEstimator):
    """Patch extractor for images with partial data (missing values are -1): only fully known
    patches are listed; patches are materialised on demand, on the device
    [ref: modl/feature_extraction/image.py:8-83]."""

    def __init__(self, patch_size=None, random_state=None, max_patches=None, device=None):
        self.patch_size = patch_size
        self.max_patches = max_patches
        self.random_state = random_state
        self.device = device

    def fit(self, X, y=None):
        self.random_state = check_random_state(self.random_state)
        dev = torch.device(self.device) if self.device is not None else default_device()
        img = X if isinstance(X, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(X))
        img = img.to(dev)
        i_h, i_w, n_channels = img.shape
        if self.patch_size is None:
            patch_size = i_h // 10, i_w // 10
        else:
            patch_size = self.patch_size
        ph, pw = int(patch_size[0]), int(patch_size[1])
        self.image_ = img
        self._patch_shape = (ph, pw, int(n_channels))
        n_p, n_q = i_h - ph + 1, i_w - pw + 1
        missing = img == -1
        if not bool(missing.any()):
            # fill(p, q, 1): every position, row-major [ref: image_fast.pyx:59-74].  Position j is (j // n_q, j % n_q),
            # so only the selected ones are materialised (there are 16.6 M for a 4096 x 4096 image).
            selection = self.random_state.permutation(n_p * n_q)[:self.max_patches]
            pp, qq = np.divmod(selection.astype(np.int64), n_q)
        else:
            # clean_mask [ref: image_fast.pyx:12-57]: a patch is dropped when a missing value falls inside it.
            # The reference's channel loop bounds its window with the patch WIDTH (`rr - y + 1`, :45), so only
            # channels rr < patch width ever invalidate a patch; reproduced as is.
            bad = missing[:, :, :min(int(n_channels), pw)].any(dim=2).to(torch.float32)[None, None]
            dirty = torch.nn.functional.max_pool2d(bad, kernel_size=(ph, pw), stride=1)[0, 0] > 0
            keep = (~dirty).cpu().numpy()
            pp, qq = np.nonzero(keep)                       # row-major, like the nested loops
            selection = self.random_state.permutation(pp.shape[0])[:self.max_patches]
            pp, qq = pp[selection], qq[selection]
        self.indices_3d = np.stack([pp, qq, np.zeros_like(pp)], axis=1).astype(np.int64)
        ar_h = torch.arange(ph, device=dev)
        ar_w = torch.arange(pw, device=dev)
        self._offsets = (ar_h[None, :, None], ar_w[None, None, :])
        return self

    def _gather(self, idx):
        idx = torch.from_numpy(np.ascontiguousarray(idx)).to(self.image_.device)
        pp, qq = idx[:, 0], idx[:, 1]
        oh, ow = self._offsets
        return self.image_[pp[:, None, None] + oh, qq[:, None, None] + ow]      # (n, ph, pw, c), a fresh tensor

    def partial_transform(self, X=None, batch=None):
        if X is not None:
            self.fit(X)
        if batch is None:
            return self.transform()
        elif isinstance(batch, int):
            batch = slice(0, batch)
        return self._gather(self.indices_3d[batch])

    def transform(self, X=None):
        if X is not None:
            self.fit(X)
        return self._gather(self.indices_3d)

    def shuffle(self, permutation=None):
        if permutation is None:
            n_samples = self.indices_3d.shape[0]
            permutation = self.random_state.permutation(n_samples)
        self.indices_3d = self.indices_3d[permutation]

    @property
    def n_patches_(self):
        return self.indices_3d.shape[0]

    @property
    def patch_shape_(self):
        return self._patch_shape


# method -> (G_agg, Dx_agg) of the first epochs; 'sgd' is handled apart (full statistics, reduction 1)
_AGGREGATION = dict([('masked', ('masked', 'masked')), ('dictionary only', ('full', 'full')),
                     ('gram', ('masked', 'masked')), ('average', ('average', 'average')),
                     ('reducing ratio', ('masked', 'masked'))])
# setting -> (non-negative codes and atoms?, centre the patches?); both settings use Lasso codes
# (code_l1_ratio 1), L2-ball atoms (comp_l1_ratio 0) and unit-norm patches
_CONSTRAINTS = {'dictionary learning': (False, True), 'NMF': (True, False)}


class ImageDictFact(BaseEstimator):
    """Dictionary learning / NMF on the patches of one image [ref: modl/decomposition/image.py:13-191].
    Same constructor keywords, fitted attributes (`dict_fact_`, `patch_shape_`, `components_`, `n_iter_`,
    `time_`) and random stream as the reference; one extra keyword, `device`."""

    # the reference exposes its two lookup tables as class attributes (:14-32)
    methods = {name: {'G_agg': g, 'Dx_agg': dx} for name, (g, dx) in _AGGREGATION.items()}
    settings = {name: {'comp_l1_ratio': 0, 'code_l1_ratio': 1, 'comp_pos': pos, 'code_pos': pos,
                       'with_std': True, 'with_mean': centre} for name, (pos, centre) in _CONSTRAINTS.items()}

    def __init__(self, method='masked', setting='dictionary learning', patch_size=(8, 8), batch_size=100,
                 buffer_size=None, step_size=1e-3, n_components=50, alpha=0.1, learning_rate=0.92, reduction=10,
                 n_epochs=1, random_state=None, callback=None, max_patches=None, verbose=0, n_threads=1,
                 device=None):
        for name, value in list(locals().items()):
            if name != 'self':
                setattr(self, name, value)

    # ---------------------------------------------------------------- pieces of fit()
    def _estimator_keywords(self):
        """`method` / `setting` -> DictFact keywords [ref: :71-113]."""
        if self.method == 'sgd':
            kw = dict(optimizer='sgd', reduction=1, G_agg='full', Dx_agg='full')
        else:
            g, dx = _AGGREGATION[self.method]
            kw = dict(optimizer='variational', reduction=self.reduction, G_agg=g, Dx_agg=dx)
        pos, _ = _CONSTRAINTS[self.setting]
        kw.update(comp_l1_ratio=0, code_l1_ratio=1, comp_pos=pos, code_pos=pos, tol=1e-2,
                  code_alpha=self.alpha, n_components=self.n_components, batch_size=self.batch_size,
                  learning_rate=self.learning_rate, step_size=self.step_size, n_epochs=self.n_epochs,
                  n_threads=self.n_threads, verbose=self.verbose, callback=self._callback,
                  random_state=self.random_state, device=self.device)
        return kw

    def _rows(self, patches, copy):
        """Scaled, flattened patches: one row per patch."""
        _, centre = _CONSTRAINTS[self.setting]
        return _flatten_patches(patches, with_mean=centre, with_std=True, copy=copy)

    def _epoch_schedule(self, epoch):
        """What changes at the start of epoch `epoch` (0-based) [ref: :142-146]."""
        if self.method == 'gram' and epoch == 4:
            self.dict_fact_.set_params(G_agg='full', Dx_agg='average')
        if self.method == 'reducing ratio':
            self.dict_fact_.set_params(reduction=1 + (self.reduction - 1) / sqrt(epoch + 1))

    def fit(self, image, y=None):
        # ONE generator feeds the patch selection, the sampler seed, the atom orders and the epoch shuffles
        self.random_state = check_random_state(self.random_state)
        self.dict_fact_ = DictFact(**self._estimator_keywords())
        extractor = LazyCleanPatchExtractor(patch_size=self.patch_size, max_patches=self.max_patches,
                                            random_state=self.random_state, device=self.device)
        if self.verbose:
            print('Preparing patch extraction')
        extractor.fit(image)
        n_patches = extractor.n_patches_
        self.patch_shape_ = extractor.patch_shape_

        if self.verbose:
            print('Fitting dictionary')
        # the first n_components patches initialise the atoms [ref: :128-132]
        self.dict_fact_.prepare(n_samples=n_patches,
                                X=self._rows(extractor.partial_transform(batch=self.n_components), copy=False))
        span = self.batch_size * 10 if self.buffer_size is None else self.buffer_size
        for epoch in range(self.n_epochs):
            if self.verbose:
                print('Epoch %i' % (epoch + 1))
            if epoch > 0:
                if self.verbose:
                    print('Shuffling dataset')
                extractor.shuffle(self.dict_fact_.shuffle())       # same permutation for the codes and the patches
            self._epoch_schedule(epoch)
            for window in gen_batches(n_patches, span):
                # The reference reuses its `buffer_size` name inside this loop (:148), so from the second epoch on
                # the windows have the length of the LAST window of the previous epoch -- which moves the minibatch
                # boundaries, hence the results.  Reproduced.
                span = window.stop - window.start
                self.dict_fact_.partial_fit(self._rows(extractor.partial_transform(batch=window), copy=False), window)
        return self

    def transform(self, patches):
        return self.dict_fact_.transform(self._rows(patches, copy=True))

    def score(self, patches):
        return self.dict_fact_.score(self._rows(patches, copy=True))

    # attributes the callbacks read [ref: :170-187]
    n_iter_ = property(lambda self: self.dict_fact_.n_iter_)
    time_ = property(lambda self: self.dict_fact_.time_)

    @property
    def components_(self):
        return self.dict_fact_.components_.reshape((self.n_components,) + tuple(self.patch_shape_))

    def _callback(self, *args):
        if self.callback is not None:
            self.callback(self)


class DictionaryScorer:
    """Callback recording the test objective along the fit [ref: modl/decomposition/image.py:202-224]
    (`time.clock`, removed from Python, is `time.perf_counter` here)."""

    def __init__(self, test_data, info=None):
        self.start_time = time.perf_counter()
        self.test_data = test_data
        self.test_time = 0
        self.time = []
        self.cpu_time = []
        self.score = []
        self.iter = []
        self.info = info

    def __call__(self, dict_fact):
        test_time = time.perf_counter()
        score = dict_fact.score(self.test_data)
        self.test_time += time.perf_counter() - test_time
        this_time = time.perf_counter() - self.start_time - self.test_time
        self.time.append(this_time)
        self.score.append(score)
        self.iter.append(dict_fact.n_iter_)
        self.cpu_time.append(dict_fact.time_)
        if self.info is not None:
            self.info['time'] = self.cpu_time
            self.info['score'] = self.score
