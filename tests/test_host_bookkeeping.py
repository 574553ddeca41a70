"""Host-side integer bookkeeping of the product (C++ behind the C ABI) against the
reference's known-answer tests and golden vectors.  CPU only: these entry points do not touch
the GPU.  [ref: modl/utils/randomkit/tests/test_random.py, test_sampler.py,
modl/utils/tests/test_utils.py]"""
import ctypes
import pickle
import re
import os

import numpy as np
import pytest

from conftest import ROOT

from modl_b200 import _lib
from modl_b200.randomkit import RandomState, Sampler


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    assert L.modl_version() >= 100
    header = open(os.path.join(ROOT, "include", "modl_b200.h")).read()
    declared = set(re.findall(r"\b(modl_[a-z0-9_A-Z]+)\s*\(", header))
    declared -= {"modl_step_params"}
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), "libmodl_b200.so does not export %s" % name
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)


def test_ctypes_structs_match_the_compiled_layout():
    """The ctypes mirrors of the three parameter structs have the size the library was compiled with (a field added on one
    side only would shift every later member silently)."""
    sizes = (ctypes.c_int64 * 3)()
    _lib.lib().modl_struct_sizes(sizes)
    assert list(sizes) == [ctypes.sizeof(_lib.StepParams), ctypes.sizeof(_lib.FitParams), ctypes.sizeof(_lib.FitBatches)]


def test_random_kat():
    rs = RandomState(seed=0)
    vals = [rs.randint(10) for _ in range(10000)]
    np.testing.assert_almost_equal(np.mean(vals), 5.018)
    vals = [rs.binomial(1000, 0.8) for _ in range(10000)]
    np.testing.assert_almost_equal(np.mean(vals), 799.8564)


def test_shuffle_kat():
    ind = np.arange(10)
    RandomState(seed=0).shuffle(ind)
    np.testing.assert_array_equal(ind, [2, 8, 4, 9, 1, 6, 7, 3, 0, 5])


def test_shuffle_with_trace_kat():
    ind, ind2 = np.arange(10), np.arange(9, -1, -1)
    perm = RandomState(seed=0).shuffle_with_trace([ind, ind2])
    np.testing.assert_array_equal(ind, [2, 8, 4, 9, 1, 6, 7, 3, 0, 5])
    np.testing.assert_array_equal(ind2, [7, 1, 5, 0, 8, 3, 2, 6, 9, 4])
    np.testing.assert_array_equal(ind, perm)


def test_permutation_kat():
    np.testing.assert_array_equal(RandomState(seed=0).permutation(10), [2, 8, 4, 9, 1, 6, 7, 3, 0, 5])


def test_random_state_pickle():
    rs = RandomState(seed=0)
    a = rs.randint(5)
    rs2 = pickle.loads(pickle.dumps(rs))
    assert a == rs2.randint(5)


def test_sampler_kat():
    s = Sampler(100, rand_size=True, replacement=True, random_seed=0)
    np.testing.assert_array_equal(s.yield_subset(10),
                                  [14, 58, 11, 49, 36, 62, 87, 45, 72, 47, 48, 13, 98, 97, 25, 93])
    assert np.mean([s.yield_subset(10).shape[0] for _ in range(100)]) == 10.19
    s = Sampler(100, rand_size=False, replacement=False, random_seed=0)
    A = np.concatenate([s.yield_subset(10) for _ in range(10)])
    np.testing.assert_array_equal(np.sort(A), np.arange(100))
    s = Sampler(100, rand_size=False, replacement=True, random_seed=0)
    np.testing.assert_array_equal(s.yield_subset(10), [6, 55, 1, 25, 87, 49, 69, 63, 13, 8])
    assert np.mean([s.yield_subset(10).shape[0] for _ in range(100)]) == 10
    A = np.concatenate([s.yield_subset(10) for _ in range(100)])
    assert np.mean(np.bincount(A)) == 10
    s = Sampler(100, rand_size=True, replacement=False, random_seed=0)
    A = np.concatenate([s.yield_subset(10) for _ in range(20)])
    np.testing.assert_array_equal(np.sort(A[:100]), np.arange(100))


def test_golden_streams(golden):
    g = golden("rng.npz")
    for seed in (0, 42, 2 ** 35 + 11):
        rs = RandomState(seed)
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
                s = Sampler(rng_, rand_size, repl, 5)
                subs = [s.yield_subset(red) for _ in range(40)]
                tag = "samp_%d_%d_%d_%g" % (rand_size, repl, rng_, red)
                np.testing.assert_array_equal([len(x) for x in subs], g[tag + "_len"])
                np.testing.assert_array_equal(np.concatenate(subs), g[tag + "_cat"])
    bw = [_lib.lib().modl_batch_weight(c, b, lr, off) for c, b, lr, off in
          ((512, 512, 1., 0.), (1024, 512, .9, 0.), (30, 10, .95, 0.), (5000, 7, .92, 0.), (70, 10, 1., 3.))]
    np.testing.assert_array_equal(bw, g["batch_weight"])


def test_streams_match_oracle_at_bench_sizes(oracle):
    """config-2 / config-4 sized samplers: product C++ vs oracle C, long streams, bit-exact."""
    for rng_, red, rs_, rp in ((10000, 8, True, True), (200000, 12, True, True), (57344, 6, False, False)):
        a, b = Sampler(rng_, rs_, rp, 123), oracle.Sampler(rng_, rs_, rp, 123)
        for _ in range(12):
            np.testing.assert_array_equal(a.yield_subset(red), b.yield_subset(red))


def test_get_sub_slice():
    from modl_b200.dict_fact import get_sub_slice
    # [ref: modl/utils/tests/test_utils.py:8-13]
    np.testing.assert_array_equal(get_sub_slice(slice(5, 15), slice(2, 4)), np.arange(7, 9))
    np.testing.assert_array_equal(get_sub_slice(None, slice(2, 4)), np.arange(2, 4))
    np.testing.assert_array_equal(get_sub_slice(np.arange(10, 20), slice(2, 4)), np.arange(12, 14))


def test_device_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    st = _lib.lib().modl_ctx_create(0, ctypes.byref(h))
    assert st != 0 and _lib.lib().modl_last_error()
    from modl_b200 import DictFact
    with pytest.raises(RuntimeError):
        DictFact(n_components=2).prepare(n_samples=4, n_features=3)


def test_sampler_lookahead_keeps_the_stream(oracle):
    """The helper thread draws the next subset speculatively on a copy of the state; changing
    `reduction` between calls must discard the speculation without disturbing the stream."""
    from modl_b200 import Sampler
    reds = [4., 4., 8., 2., 2., 2., 5., 5., 3.]
    for rand_size in (True, False):
        for replacement in (True, False):
            mine = Sampler(997, rand_size, replacement, 7)
            ref = oracle.Sampler(997, rand_size, replacement, 7)
            for r in reds:
                np.testing.assert_array_equal(mine.yield_subset(r), ref.yield_subset(r))
