"""World-size-2 test of the sample-sharded step's HOST logic on CPU (gloo backend).

The device phases are replaced by the CPU oracle (tests may do that; the product never does):
what is under test is modl_b200.distributed.ShardedStepMixin -- shard bookkeeping, global batch
weight, identical RNG streams on every rank, the all-reduce of the statistics increments, and
that the replicas stay identical -- against the single-process oracle run on the concatenated
batch."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_problem(dtype):
    rng = np.random.RandomState(11)
    n, p, k = 96, 60, 8
    X = (rng.randn(n, k) @ rng.randn(k, p) + 0.1 * rng.randn(n, p)).astype(dtype)
    return X, k


KW = dict(reduction=3, code_alpha=0.05, code_l1_ratio=0.9, random_state=0)


def _cpu_sharded_class():
    """ShardedStepMixin driving the oracle's numerics instead of the CUDA phases."""
    import oracle as orc
    from modl_b200.dict_fact import DictFact
    from modl_b200.distributed import ShardedStepMixin

    class CpuSharded(ShardedStepMixin, orc.OracleDictFact):
        _host_bookkeeping = DictFact._host_bookkeeping

        def __init__(self, process_group=None, device=None, **kw):
            orc.OracleDictFact.__init__(self, **kw)
            self.process_group = process_group

        def prepare(self, **kw):
            orc.OracleDictFact.prepare(self, **kw)
            self._np_dtype = self.components_.dtype
            return self

        def partial_fit(self, X, sample_indices=None):
            X = X.numpy() if isinstance(X, torch.Tensor) else X
            return orc.OracleDictFact.partial_fit(self, X, sample_indices)

        def _replicated_step(self, X, sample_indices):
            X = X.numpy() if isinstance(X, torch.Tensor) else X
            return orc.OracleDictFact._single_batch_fit(self, np.ascontiguousarray(X), sample_indices)

        def _callback(self):
            pass

        def _phase_code_and_increments(self, X, idx, subset, order, w, w_sample, b_global):
            X = np.ascontiguousarray(X)
            self._compute_code(X, idx, w_sample, subset)
            code = self.code_[idx]
            inc = np.concatenate([(w * code.T.dot(code) / b_global).ravel(),
                                  (w * code.T.dot(X) / b_global).ravel()])
            return torch.from_numpy(np.ascontiguousarray(inc.astype(self.components_.dtype)))

        def _phase_apply_and_dict(self, X, idx, subset, order, w, w_sample, inc, b_global):
            k, p = self.components_.shape
            inc = inc.numpy()
            self.C_ *= 1 - w
            self.C_ += inc[:k * k].reshape(k, k)
            self.B_ *= 1 - w
            self.B_ += inc[k * k:].reshape(k, p)
            D_sub = np.ascontiguousarray(self.components_[:, subset])
            grad = np.ascontiguousarray(self.B_[:, subset])
            full, small = self.G_agg == 'full', subset.shape[0] < p / 2.
            if full and small:                                # G_ down/up-date [ref: dict_fact.py:667-668, 711-715]
                self.G_ -= D_sub.dot(D_sub.T)
            orc.update_dict_panel(D_sub, grad, self.C_, self.comp_norm_, order, self.comp_l1_ratio, self.comp_pos)
            self.components_[:, subset] = D_sub
            if full:
                if small:
                    self.G_ += D_sub.dot(D_sub.T)
                else:
                    self.G_[:] = self.components_.dot(self.components_.T)

    return CpuSharded


def _init(rank, world, port):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _worker(rank, world, port, dtype_name, out_dir):
    _init(rank, world, port)
    CpuSharded = _cpu_sharded_class()
    dtype = np.dtype(dtype_name)

    X, k = _make_problem(dtype)
    b_local = 8
    est = CpuSharded(n_components=k, batch_size=b_local, **KW)
    est.prepare(n_samples=X.shape[0], X=X[:k])
    n_steps = X.shape[0] // (b_local * world)
    for t in range(n_steps):
        lo = t * b_local * world + rank * b_local
        idx = np.arange(lo, lo + b_local)
        est._single_batch_fit(X[idx], idx)
    # replicas identical?
    D = torch.from_numpy(est.components_.copy())
    lo_, hi_ = D.clone(), D.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    spread = float((hi_ - lo_).abs().max())
    # gather the codes (each rank owns the rows it solved)
    code = torch.from_numpy(est.code_.copy())
    owner = torch.zeros(X.shape[0], dtype=torch.int64)
    for t in range(n_steps):
        lo = t * b_local * world + rank * b_local
        owner[lo:lo + b_local] = 1
    code = code * owner[:, None].to(code.dtype)
    dist.all_reduce(code)
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded_%s.npz" % dtype_name), D=est.components_, C=est.C_, B=est.B_,
                 code=code.numpy(), norm=est.comp_norm_, n_iter=est.n_iter_, spread=spread)
    dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name", ["float64", "float32"])
def test_sharded_step_equals_single_process(tmp_path, oracle, dtype_name):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dtype_name, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded_%s.npz" % dtype_name))
    assert float(got["spread"]) == 0.0                      # replicas bit-identical
    X, k = _make_problem(np.dtype(dtype_name))
    ref = oracle.OracleDictFact(n_components=k, batch_size=16, **KW)
    ref.prepare(n_samples=X.shape[0], X=X[:k])
    ref.partial_fit(X)
    tol = 1e-11 if dtype_name == "float64" else 2e-4
    err = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
    assert int(got["n_iter"]) == ref.n_iter_
    assert err(got["D"], ref.components_) < tol
    assert err(got["C"], ref.C_) < tol
    assert err(got["B"], ref.B_) < tol
    solved = slice(0, (X.shape[0] // 16) * 16)
    assert err(got["code"][solved], ref.code_[solved]) < tol


# ------------------------------------------------------------------------------------------------ fMRI loop
def _fmri_worker(rank, world, port, case_index, out_dir):
    _init(rank, world, port)
    import json
    from modl_b200 import fmri
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_fmri_frontend import _records
    fmri.ShardedDictFact = _cpu_sharded_class()            # the oracle's numerics under the real sharding logic
    gold = np.load(os.path.join(ROOT, "tests", "golden", "fmri.npz"))
    spec = json.loads(str(gold["cases"]))
    case = spec["cases"][case_index]
    records, n_voxels = _records(case["dtype"])
    masker = fmri.RecordMasker(mask=np.ones(n_voxels, dtype=bool)).fit()
    kw = dict(spec["common"])
    kw.update(case["kw"])
    comp = fmri._compute_components(masker, records, device="cpu", sharded=True, **kw)
    D = torch.from_numpy(np.ascontiguousarray(comp))
    lo_, hi_ = D.clone(), D.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.savez(os.path.join(out_dir, "fmri_%d.npz" % case_index), D=comp, spread=float((hi_ - lo_).abs().max()))
    dist.destroy_process_group()


@pytest.mark.parametrize("case_index", [0, 4, 7])       # masked f64, dictionary only + positive maps, masked f32
def test_sharded_fmri_loop_equals_the_reference(tmp_path, oracle, case_index):
    """`_compute_components(sharded=True)` on two ranks: minibatches of 8 volumes split 4 + 4, the ragged tails of
    the 23- and 17-volume records (7 and 1 rows) replicated, those of the 30- and 20-volume ones (6 and 4) split;
    every rank ends with the same maps, equal to what the unmodified reference computes in one process."""
    world = 2
    mp.spawn(_fmri_worker, args=(world, _free_port(), case_index, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "fmri_%d.npz" % case_index))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "fmri.npz"))
    want = gold["fit_%d_components" % case_index]
    assert float(got["spread"]) == 0.0
    err = float(np.linalg.norm(got["D"].astype(np.float64) - want) / np.linalg.norm(want))
    assert err < (1e-9 if want.dtype == np.float64 else 2e-3), err
