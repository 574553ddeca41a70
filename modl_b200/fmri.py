"""fMRIDictFact -- the reference's fMRI front-end over the B200 hot path, without the neuro-imaging stack
[ref: modl/decomposition/fmri.py:40-680].

"Next" row 2 of SURVEY section 8f.  What this file keeps is the LEARNING LOOP of the reference,
`_compute_components` (fmri.py:423-546): `method` -> aggregation keywords, ridge codes + L1-ball atoms
(`code_l1_ratio=0, comp_l1_ratio=1`), `n_samples = total + 1` (the reference's off-by-one, :469), one
shared NumPy `RandomState` that feeds the sampler seed, the record order of every epoch, the row
permutation of every record and the atom orders, the epoch schedules exactly as the reference executes
them (see `_compute_components`), the final sign flip (`_flip`, :549-556) -- plus the estimator shells
(`fMRIDictFact`, `fMRICoder`, their `fit` / `transform` / `score`) and the scoring callback.

What it replaces is nilearn / nibabel (absent here, and out of the hot path): records are arrays or
`.npy` files -- the branch the reference itself takes when `check_niimg` raises `ImageFileError`
(fmri.py:564-572) -- and the masker is any object with `transform(img, confounds=None)`,
`inverse_transform(components)` and `mask_img_`; `RecordMasker` below is the built-in one.

Device side: a record is uploaded ONCE, its row permutation is a device gather, and the permuted rows
are handed to `DictFact.partial_fit` as a CUDA tensor -- at the shape of BASELINE configs[3]
(p ~ 2e5 voxels, k = 70, reduction 12) a record of 1 200 volumes is 0.96 GB and never returns to the host.
"""
import itertools
import os
import time
import warnings
from math import sqrt

import numpy as np
import torch
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils import check_random_state

from ._util import default_device
from .dict_fact import Coder, DictFact
from .distributed import ShardedDictFact

__all__ = ["fMRIDictFact", "fMRICoder", "fMRICoderMixin", "RecordMasker", "rfMRIDictionaryScorer",
           "_compute_components", "_flip", "_lazy_scan", "_check_dict_init"]

# method -> (G_agg, Dx_agg) [ref: fmri.py:440-445]
_AGGREGATION = {'masked': ('masked', 'masked'),
                'dictionary only': ('full', 'full'),
                'gram': ('masked', 'masked'),
                'average': ('average', 'average'),
                'reducing ratio': ('masked', 'masked')}


def _load_record(img, mmap=False):
    """A record is an array-like (n_samples, ...) or the path of a `.npy` file holding one."""
    if isinstance(img, (str, os.PathLike)):
        return np.load(img, mmap_mode='r' if mmap else None)
    if isinstance(img, torch.Tensor):
        return img
    return np.asarray(img)


class _MaskImage(object):
    """The part of a Niimg the learning loop reads: `get_data()` [ref: fmri.py:470]."""

    def __init__(self, mask):
        self._mask = np.asarray(mask, dtype=bool)
        self.shape = self._mask.shape

    def get_data(self):
        return self._mask


class RecordMasker(BaseEstimator):
    """Masker over array records -- the stand-in for `nilearn.input_data.NiftiMasker` /
    `MultiNiftiMasker` in the reference's pipeline [ref: modl/input_data/fmri/base.py:40-64].

    mask: boolean array of any shape (the volume grid), or None: every voxel of the first record.
    A record is (n_samples, n_voxels) -- already masked -- or (n_samples,) + mask.shape, in which case the
    mask is applied here.  `detrend` removes the per-voxel linear trend, `standardize` centres each voxel's
    time series and scales it to unit variance (constant voxels are left at zero); both default to off and
    run on `device` when one is given.  They are conveniences, NOT pinned to nilearn's `signal.clean`."""

    def __init__(self, mask=None, standardize=False, detrend=False, device=None):
        self.mask = mask
        self.standardize = standardize
        self.detrend = detrend
        self.device = device

    def fit(self, imgs=None, y=None):
        if self.mask is not None:
            mask = np.asarray(self.mask.get_data() if hasattr(self.mask, 'get_data') else self.mask) != 0
        else:
            if imgs is None:
                raise ValueError('RecordMasker needs a mask or records to infer one from')
            if isinstance(imgs, (str, os.PathLike)) or not isinstance(imgs, (list, tuple)):
                imgs = [imgs]
            if len(imgs) == 0:
                raise ValueError('Need one or more records as input, an empty list was given.')
            first = _load_record(imgs[0], mmap=True)
            mask = np.ones(tuple(first.shape[1:]), dtype=bool)
        self.mask_img_ = _MaskImage(mask)
        self.n_voxels_ = int(mask.sum())
        return self

    def _check_fitted(self):
        if not hasattr(self, 'mask_img_'):
            raise ValueError('It seems that %s has not been fitted.' % self.__class__.__name__)

    def _clean(self, data):
        if not (self.detrend or self.standardize):
            return data
        numpy_in = not isinstance(data, torch.Tensor)
        x = torch.as_tensor(np.ascontiguousarray(data) if numpy_in else data)
        if self.device is not None:
            x = x.to(self.device)
        x = x.clone() if x.is_floating_point() else x.to(torch.float64)
        n = x.shape[0]
        if self.detrend and n > 1:
            t = torch.arange(n, dtype=x.dtype, device=x.device)
            t = (t - t.mean()) / torch.sqrt(((t - t.mean()) ** 2).sum())
            x -= x.mean(dim=0, keepdim=True)
            x -= t[:, None] * (t[:, None] * x).sum(dim=0, keepdim=True)
        if self.standardize:
            x -= x.mean(dim=0, keepdim=True)
            std = torch.sqrt((x ** 2).mean(dim=0, keepdim=True))
            std[std < torch.finfo(x.dtype).eps] = 1
            x /= std
        return x.cpu().numpy() if numpy_in and self.device is None else x

    def transform(self, imgs, confounds=None):
        """One record -> (n_samples, n_voxels); a list of records -> a list of them."""
        self._check_fitted()
        if isinstance(imgs, (list, tuple)):
            return [self.transform(img) for img in imgs]
        # a `.npy` record is memory-mapped when no cleaning is asked for: under sharding a rank then reads only its rows
        data = _load_record(imgs, mmap=not (self.detrend or self.standardize))
        mask = self.mask_img_.get_data()
        if data.ndim == 2 and data.shape[1] == self.n_voxels_:
            pass                                                   # already masked
        elif tuple(data.shape[1:]) == mask.shape:
            data = data[:, torch.as_tensor(mask)] if isinstance(data, torch.Tensor) else data[:, mask]
        else:
            raise ValueError('record of shape %s matches neither the mask %s nor its %d voxels'
                             % (tuple(data.shape), mask.shape, self.n_voxels_))
        return self._clean(data)

    def inverse_transform(self, components):
        """(k, n_voxels) -> (k,) + mask.shape, zeros outside the mask."""
        self._check_fitted()
        mask = self.mask_img_.get_data()
        components = np.asarray(components)
        out = np.zeros((components.shape[0],) + mask.shape, dtype=components.dtype)
        out[:, mask] = components
        return out


def _check_dict_init(dict_init, mask_img, n_components=None, masker=None):
    """[ref: fmri.py:405-420]  An array must have one column per voxel of the mask; anything else goes
    through the masker."""
    if dict_init is None:
        return None
    if isinstance(dict_init, (np.ndarray, torch.Tensor)) and dict_init.ndim == 2:
        assert dict_init.shape[1] == mask_img.get_data().sum()
        components = dict_init
    else:
        if masker is None:
            masker = RecordMasker(mask=mask_img).fit()
        components = RecordMasker(mask=masker.mask_img_).fit().transform(dict_init)   # no cleaning of atoms
    if n_components is not None:
        return components[:n_components]
    return components


def _lazy_scan(imgs):
    """Number of samples of every record and the dtype of the LAST one, without loading the data
    [ref: fmri.py:559-575; the reference also prints every record, not reproduced]."""
    n_samples_list = []
    dtype = None
    for img in imgs:
        rec = _load_record(img, mmap=True)
        n_samples_list.append(int(rec.shape[0]))
        dtype = np.dtype(str(rec.dtype).replace('torch.', '')) if isinstance(rec, torch.Tensor) else rec.dtype
    return n_samples_list, dtype


def _flip(components):
    """Flip the sign of every map with more negative than positive entries [ref: fmri.py:549-556]."""
    if isinstance(components, torch.Tensor):
        components = components.clone()
        flip = (components < 0).sum(dim=1) > (components > 0).sum(dim=1)
        components[flip] *= -1
        return components
    components = components.copy()
    for component in components:
        if np.sum(component < 0) > np.sum(component > 0):
            component *= -1
    return components


def _stage_record(masked_data, permutation, dtype, device):
    """`masked_data.astype(dtype)[permutation]` on the device [ref: fmri.py:519, 532].  A whole record is one H2D
    copy plus one device gather; a partial row set (one rank's share under sharding) is gathered on the host
    first, so that only those rows cross PCIe."""
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    permutation = np.asarray(permutation, dtype=np.int64)
    if isinstance(masked_data, torch.Tensor):
        rows = masked_data.to(device=device, dtype=tdt)
    elif permutation.shape[0] < masked_data.shape[0]:
        return torch.from_numpy(np.ascontiguousarray(masked_data[permutation], dtype=dtype)).to(device)
    else:
        host = np.ascontiguousarray(masked_data, dtype=dtype)
        with warnings.catch_warnings():                  # a memory-mapped record is read-only; it is only read here
            warnings.simplefilter("ignore", UserWarning)
            rows = torch.from_numpy(host).to(device)
    return rows.index_select(0, torch.from_numpy(permutation).to(device))


def _sharded_record_fit(dict_fact, masked_data, permutation, sample_indices, batch_size, world, rank, dtype, device):
    """One record under sample sharding: minibatch m of the permuted record is rows
    [m b, m b + t); rank g solves the g-th of `world` equal shares and the statistics increments are summed over
    ranks (`ShardedDictFact`).  A ragged last minibatch whose size does not divide by `world` is run REPLICATED
    -- every rank processes all of its rows, no exchange -- which keeps the replicas bit-identical without
    unequal shards.  Sample index = position in the permuted record (fmri.py:528-531 as executed).  Which rank solves
    a position depends on the record's length (a shorter record re-partitions its tail minibatch), so the rows of
    `code_` a rank holds are NOT a fixed set across records: harmless here because the fMRI codes are ridge solutions
    (code_l1_ratio = 0: no warm start is read back), and the reason `code_` must not be compared across ranks."""
    n = permutation.shape[0]
    plan, mine = [], []
    for lo in range(0, n, batch_size):
        t = min(batch_size, n - lo)
        if t % world == 0:
            share = t // world
            pos = np.arange(lo + rank * share, lo + (rank + 1) * share)
            plan.append((True, pos))
        else:
            pos = np.arange(lo, lo + t)
            plan.append((False, pos))
        mine.append(pos)
    mine = np.concatenate(mine)
    rows = _stage_record(masked_data, permutation[mine], dtype, device)      # this rank's rows only, one upload
    at = 0
    for shard, pos in plan:
        X = rows[at:at + pos.shape[0]]
        at += pos.shape[0]
        idx = pos if sample_indices is None else sample_indices[pos]
        if shard:
            dict_fact.partial_fit(X, sample_indices=idx)                     # <= local batch rows: exactly one step
        else:
            dict_fact._replicated_step(X, idx)


def _compute_components(masker, imgs, step_size=1, confounds=None, dict_init=None, alpha=1, positive=False,
                        reduction=1, learning_rate=1, n_components=20, batch_size=20, n_epochs=1,
                        method='masked', verbose=0, random_state=None, callback=None, n_jobs=1, device=None,
                        return_estimator=False, intended_schedules=False, sharded=False, process_group=None):
    """The learning loop [ref: fmri.py:423-546], same signature plus `device`, `return_estimator` (also hand
    back the fitted `DictFact`), `intended_schedules`, and `sharded` / `process_group`: with
    `torch.distributed` initialised, every rank calls this function with the SAME arguments and seed; each
    minibatch of `batch_size` volumes is split over the ranks (`_sharded_record_fit`), every rank returns the
    same maps, and the result equals the single-process one up to the summation order of the all-reduce.

    The reference rebinds `method` to its table entry (`method = methods[method]`, fmri.py:459), so the three
    string tests that follow (:497 'gram' switch, :500 'reducing ratio' schedule, :528 per-record
    `sample_indices` for 'average' / 'gram') never fire: every method keeps its first-epoch keywords and every
    record addresses rows 0..len-1 of the per-sample state.  That is what the reference computes, so it is the
    default here (and what the golden vectors pin); `intended_schedules=True` runs the schedules as written."""
    masker._check_fitted()
    dict_init = _check_dict_init(dict_init, mask_img=masker.mask_img_, n_components=n_components, masker=masker)
    if dict_init is not None:
        n_components = dict_init.shape[0]            # dict_init might have fewer components than asked for
    random_state = check_random_state(random_state)
    if method == 'sgd':
        optimizer, G_agg, Dx_agg, reduction = 'sgd', 'full', 'full', 1
    else:
        G_agg, Dx_agg = _AGGREGATION[method]
        optimizer = 'variational'
    schedule = method if intended_schedules else None

    if verbose:
        print("Scanning data")
    n_records = len(imgs)
    if confounds is None:
        confounds = itertools.repeat(None)
    data_list = list(zip(imgs, confounds))
    n_samples_list, dtype = _lazy_scan(imgs)
    indices_list = np.zeros(len(imgs) + 1, dtype='int')
    indices_list[1:] = np.cumsum(n_samples_list)
    n_samples = indices_list[-1] + 1                 # sic [ref: fmri.py:469]
    n_voxels = int(np.sum(masker.mask_img_.get_data() != 0))

    if verbose:
        print("Learning...")
    kw = dict(n_components=n_components, code_alpha=alpha, code_l1_ratio=0, comp_l1_ratio=1, comp_pos=positive,
              reduction=reduction, Dx_agg=Dx_agg, optimizer=optimizer, step_size=step_size, G_agg=G_agg,
              learning_rate=learning_rate, batch_size=batch_size, random_state=random_state, n_threads=n_jobs,
              verbose=0, device=device)
    world, rank = 1, 0
    if sharded:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("sharded=True needs an initialised torch.distributed process group")
        world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        if batch_size % world:
            raise ValueError("batch_size=%d does not divide over %d ranks" % (batch_size, world))
        if 'average' in (G_agg, Dx_agg) or intended_schedules:
            raise NotImplementedError("per-sample running averages are owned by row; not available under sharding")
        kw.update(batch_size=batch_size // world, process_group=process_group)
        dict_fact = ShardedDictFact(**kw)
    else:
        dict_fact = DictFact(**kw)
    dict_fact.prepare(n_samples=n_samples, n_features=n_voxels, X=dict_init, dtype=dtype)
    stage_device = getattr(dict_fact, '_device', None) or (torch.device(device) if device is not None else default_device())
    cpu_time = 0
    io_time = 0
    if n_records > 0:
        if verbose:
            verbose_iter_ = np.linspace(0, n_records * n_epochs, verbose).tolist()
        current_n_records = 0
        for i in range(n_epochs):
            if verbose:
                print('Epoch %i' % (i + 1))
            if schedule == 'gram' and i == 5:
                dict_fact.set_params(G_agg='full', Dx_agg='average')
            if schedule == 'reducing ratio':
                reduction = 1 + (reduction - 1) / sqrt(i + 1)        # compounds from epoch to epoch [ref: :500]
                dict_fact.set_params(reduction=reduction)
            record_list = random_state.permutation(n_records)
            for record in record_list:
                if verbose and verbose_iter_ and current_n_records >= verbose_iter_[0]:
                    print('Record %i' % current_n_records)
                    if callback is not None:
                        callback(masker, dict_fact, cpu_time, io_time)
                    verbose_iter_ = verbose_iter_[1:]

                # IO bound: load + mask the record
                t0 = time.perf_counter()
                img, these_confounds = data_list[record]
                masked_data = masker.transform(img, confounds=these_confounds)
                io_time += time.perf_counter() - t0

                # device bound
                t0 = time.perf_counter()
                permutation = random_state.permutation(masked_data.shape[0])
                if schedule in ['average', 'gram']:
                    sample_indices = np.arange(indices_list[record], indices_list[record + 1])[permutation]
                else:
                    sample_indices = None
                if sharded:
                    _sharded_record_fit(dict_fact, masked_data, permutation, sample_indices, batch_size, world, rank,
                                        dtype, stage_device)
                else:
                    dict_fact.partial_fit(_stage_record(masked_data, permutation, dtype, stage_device),
                                          sample_indices=sample_indices)
                current_n_records += 1
                cpu_time += time.perf_counter() - t0
    components = _flip(dict_fact.components_)
    if return_estimator:
        return components, dict_fact
    return components


def _transform_img(coder, masker, img, confounds):
    return coder.transform(masker.transform(img, confounds=confounds))


def _score_img(coder, masker, img, confounds):
    return coder.score(masker.transform(img, confounds=confounds))


def _as_list(imgs):
    if isinstance(imgs, (str, os.PathLike)) or not isinstance(imgs, (list, tuple)):
        return [imgs]
    return imgs


class fMRICoderMixin(BaseEstimator, TransformerMixin):
    """Masking + ridge coding against a fixed set of maps [ref: fmri.py:40-180].  The nilearn masker keywords
    of the reference are accepted and stored (so that parameter grids written for it keep working); the ones
    `RecordMasker` implements -- `mask`, `standardize`, `detrend` -- take effect, the rest need a real masker
    passed as `mask`."""

    def __init__(self, n_components=20, alpha=0.1, dict_init=None, transform_batch_size=None, mask=None,
                 smoothing_fwhm=None, standardize=True, detrend=True, low_pass=None, high_pass=None, t_r=None,
                 target_affine=None, target_shape=None, mask_strategy='background', mask_args=None, memory=None,
                 memory_level=2, n_jobs=1, verbose=0, device=None):
        for name, value in list(locals().items()):
            if name != 'self':
                setattr(self, name, value)

    def _fit_masker(self, imgs):
        """[ref: modl/input_data/fmri/base.py:40-64]"""
        if hasattr(self.mask, 'transform'):
            self.masker_ = self.mask
            if not hasattr(self.masker_, 'mask_img_'):
                self.masker_.fit(imgs)
        else:
            self.masker_ = RecordMasker(mask=self.mask, standardize=self.standardize, detrend=self.detrend,
                                        device=self.device).fit(imgs)
        self.mask_img_ = self.masker_.mask_img_

    def _set_components(self, components):
        self.components_ = components
        self.components_img_ = self.masker_.inverse_transform(components)
        self.coder_ = Coder(dictionary=components, code_alpha=self.alpha, code_l1_ratio=0, n_threads=self.n_jobs,
                            device=self.device).fit()

    def _base_fit(self, imgs=None, confounds=None):
        if imgs is not None:
            imgs = _as_list(imgs)
            if len(imgs) == 0:
                raise ValueError('Need one or more Niimg-like objects as input, an empty list was given.')
            self._fit_masker(imgs)
        elif self.dict_init is not None:
            self._fit_masker([self.dict_init])
        else:
            self._fit_masker(None)
        self.components_ = _check_dict_init(self.dict_init, mask_img=self.mask_img_, n_components=self.n_components,
                                            masker=self.masker_)
        if self.components_ is not None:
            self._set_components(np.asarray(self.components_))

    def fit(self, imgs=None, y=None, confounds=None):
        self._base_fit(imgs, confounds)
        return self

    def score(self, imgs, confounds=None):
        """Length-weighted mean objective over the records; lower is better [ref: fmri.py:96-130]."""
        imgs = _as_list(imgs)
        if confounds is None:
            confounds = itertools.repeat(None)
        scores, len_imgs = [], []
        for img, these_confounds in zip(imgs, confounds):
            data = self.masker_.transform(img, confounds=these_confounds)
            scores.append(self.coder_.score(data))
            len_imgs.append(data.shape[0])
        scores, len_imgs = np.array(scores), np.array(len_imgs)
        return np.sum(scores * len_imgs) / np.sum(len_imgs)

    def transform(self, imgs, confounds=None):
        """One (n_samples, n_components) array of loadings per record [ref: fmri.py:132-159]."""
        imgs = _as_list(imgs)
        if confounds is None:
            confounds = itertools.repeat(None)
        return [_transform_img(self.coder_, self.masker_, img, these_confounds)
                for img, these_confounds in zip(imgs, confounds)]


class fMRIDictFact(fMRICoderMixin):
    """Sparse spatial maps from a list of records [ref: fmri.py:183-366]; same keywords plus `device`."""

    def __init__(self, n_components=20, alpha=0.1, dict_init=None, transform_batch_size=None, mask=None,
                 smoothing_fwhm=None, standardize=True, detrend=True, low_pass=None, high_pass=None, t_r=None,
                 target_affine=None, target_shape=None, mask_strategy='background', mask_args=None, memory=None,
                 memory_level=0, n_jobs=1, verbose=0, method='masked', n_epochs=1, batch_size=20, reduction=1,
                 step_size=1, positive=False, learning_rate=1, random_state=None, callback=None, device=None):
        fMRICoderMixin.__init__(self, n_components=n_components, alpha=alpha, dict_init=dict_init,
                                transform_batch_size=transform_batch_size, mask=mask, smoothing_fwhm=smoothing_fwhm,
                                standardize=standardize, detrend=detrend, low_pass=low_pass, high_pass=high_pass,
                                t_r=t_r, target_affine=target_affine, target_shape=target_shape,
                                mask_strategy=mask_strategy, mask_args=mask_args, memory=memory,
                                memory_level=memory_level, n_jobs=n_jobs, verbose=verbose, device=device)
        self.method = method
        self.n_epochs = n_epochs
        self.batch_size = batch_size
        self.reduction = reduction
        self.step_size = step_size
        self.positive = positive
        self.learning_rate = learning_rate
        self.random_state = random_state
        self.callback = callback

    def fit(self, imgs=None, y=None, confounds=None):
        if imgs is None:
            raise ValueError('imgs is None, use fMRICoder instead')
        imgs = _as_list(imgs)
        self._base_fit(imgs, confounds)
        components, self.dict_fact_ = _compute_components(
            self.masker_, imgs, step_size=self.step_size, confounds=confounds, dict_init=self.components_,
            alpha=self.alpha, reduction=self.reduction, learning_rate=self.learning_rate,
            n_components=self.n_components, batch_size=self.batch_size, positive=self.positive,
            n_epochs=self.n_epochs, method=self.method, verbose=self.verbose, random_state=self.random_state,
            callback=self.callback, n_jobs=self.n_jobs, device=self.device, return_estimator=True)
        self._set_components(components)
        return self


class fMRICoder(fMRICoderMixin):
    """Coder over a given dictionary of maps [ref: fmri.py:369-402]."""

    def __init__(self, dictionary, alpha=0.1, transform_batch_size=None, mask=None, smoothing_fwhm=None,
                 standardize=False, detrend=False, low_pass=None, high_pass=None, t_r=None, target_affine=None,
                 target_shape=None, mask_strategy='background', mask_args=None, memory=None, memory_level=2,
                 n_jobs=1, verbose=0, device=None):
        self.dictionary = dictionary
        fMRICoderMixin.__init__(self, n_components=None, alpha=alpha, dict_init=dictionary,
                                transform_batch_size=transform_batch_size, mask=mask, smoothing_fwhm=smoothing_fwhm,
                                standardize=standardize, detrend=detrend, low_pass=low_pass, high_pass=high_pass,
                                t_r=t_r, target_affine=target_affine, target_shape=target_shape,
                                mask_strategy=mask_strategy, mask_args=mask_args, memory=memory,
                                memory_level=memory_level, n_jobs=n_jobs, verbose=verbose, device=device)


class rfMRIDictionaryScorer:
    """Callback recording the test objective along the fit [ref: fmri.py:590-646].  Artifacts are `.npy`
    arrays of the flipped maps (the reference writes NIfTI images through nibabel)."""

    def __init__(self, test_imgs, test_confounds=None, info=None, artifact_dir=None):
        self.start_time = time.perf_counter()
        self.test_imgs = test_imgs
        if test_confounds is None:
            test_confounds = itertools.repeat(None)
        self.test_confounds = test_confounds
        self.test_time = 0
        self.score = []
        self.iter = []
        self.time = []
        self.cpu_time = []
        self.io_time = []
        self.info = info
        self.artifact_dir = artifact_dir

    def __call__(self, masker, dict_fact, cpu_time, io_time):
        test_time = time.perf_counter()
        if not hasattr(self, 'data'):
            self.data = [masker.transform(img, confounds=c)
                         for img, c in zip(_as_list(self.test_imgs), self.test_confounds)]
        scores = np.array([dict_fact.score(data) for data in self.data])
        len_imgs = np.array([data.shape[0] for data in self.data])
        score = np.sum(scores * len_imgs) / np.sum(len_imgs)
        self.test_time += time.perf_counter() - test_time
        this_time = time.perf_counter() - self.start_time - self.test_time
        self.score.append(score)
        self.time.append(this_time)
        self.cpu_time.append(cpu_time)
        self.io_time.append(io_time)
        self.iter.append(dict_fact.n_iter_)
        if self.info is not None:
            self.info['time'] = self.cpu_time
            self.info['score'] = self.score
            self.info['iter'] = self.iter
        if self.artifact_dir is not None:
            components = masker.inverse_transform(_flip(np.asarray(dict_fact.components_)))
            np.save(os.path.join(self.artifact_dir, 'components_%i.npy' % dict_fact.n_iter_), components)
