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


def _worker(rank, world, port, dtype_name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    from modl_b200.dict_fact import DictFact
    from modl_b200.distributed import ShardedStepMixin

    dtype = np.dtype(dtype_name)

    class CpuSharded(ShardedStepMixin, orc.OracleDictFact):
        """ShardedStepMixin driving the oracle's numerics instead of the CUDA phases."""
        _host_bookkeeping = DictFact._host_bookkeeping

        def prepare(self, **kw):
            orc.OracleDictFact.prepare(self, **kw)
            self._np_dtype = self.components_.dtype
            return self

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
            orc.update_dict_panel(D_sub, grad, self.C_, self.comp_norm_, order, self.comp_l1_ratio, self.comp_pos)
            self.components_[:, subset] = D_sub

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
