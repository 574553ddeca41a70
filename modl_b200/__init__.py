"""modl_b200 -- B200-native implementation of MODL's per-minibatch inner loop
(`DictFact._single_batch_fit` of arthurmensch/modl) behind the reference's estimator API.

    from modl_b200 import DictFact, Coder, ImageDictFact, fMRIDictFact, RecsysDictFact, ShardedDictFact

Host code is Python; the hot path is hand-written CUDA for sm_100a in
`modl_b200/csrc`, reached through the C ABI of `include/modl_b200.h`
(`libmodl_b200.so`, built in-tree by `python -m modl_b200.build`).
"""
from ._lib import LIB_PATH, ModlError, get_context, lib  # noqa: F401
from .randomkit import RandomState, Sampler  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # the estimators need torch + sklearn; import them lazily so that the host-only helpers
    # (sampler, RNG) stay importable in minimal environments
    if name in ("DictFact", "Coder", "CodingMixin", "get_sub_slice"):
        from . import dict_fact
        return getattr(dict_fact, name)
    if name in ("ImageDictFact", "LazyCleanPatchExtractor", "scale_patches", "DictionaryScorer"):
        from . import image
        return getattr(image, name)
    if name in ("fMRIDictFact", "fMRICoder", "RecordMasker", "rfMRIDictionaryScorer"):
        from . import fmri
        return getattr(fmri, name)
    if name in ("RecsysDictFact", "compute_biases", "rmse"):
        from . import recsys
        return getattr(recsys, name)
    if name in ("ShardedDictFact",):
        from . import distributed
        return getattr(distributed, name)
    if name in ("enet_norm", "enet_projection", "enet_scale"):
        from . import enet
        return getattr(enet, name)
    raise AttributeError(name)
