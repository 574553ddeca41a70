"""Sample-sharded data parallelism of the minibatch step: one process per GPU.

A minibatch shards naturally over its samples (the reference's own thread
pool splits the batch the same way, dict_fact.py:584-586, 621-635): every rank holds replicas
of `components_`, `C_`, `B_`, `comp_norm_` and identically seeded host RNG streams (same
subset, same atom order on every rank), solves the codes of ITS rows, and contributes its
share of the statistics increments.  The one exchange step is a sum of those increments over
ranks -- `torch.distributed.all_reduce` (NCCL over NVLink 5 / NVSwitch on B200) of a single
k x (k + p) buffer -- after which every rank applies the decay and runs the same deterministic
dictionary update on bit-identical inputs, so the replicas never diverge and no dictionary
broadcast is needed.

    dist.init_process_group("nccl")
    est = ShardedDictFact(n_components=256, batch_size=512, ...)   # batch_size = rows PER RANK
    est.prepare(n_samples=n, X=X0)            # same X0 on every rank
    est.partial_fit(X_local, sample_indices_local)

The result equals the single-process estimator run with batch_size * world_size and the ranks'
rows concatenated in rank order (up to floating-point summation order of the all-reduce).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .dict_fact import DictFact

__all__ = ["ShardedStepMixin", "ShardedDictFact"]


class ShardedStepMixin(object):
    """Host-side orchestration of the sharded step; the two device phases are provided by the
    concrete class (`_phase_code_and_increments`, `_phase_apply_and_dict`)."""

    process_group = None

    def _world(self):
        if not dist.is_available() or not dist.is_initialized():
            return 1, 0
        return dist.get_world_size(self.process_group), dist.get_rank(self.process_group)

    def _single_batch_fit(self, X, sample_indices):
        world, rank = self._world()
        if self.G_agg == 'full' and self.optimizer != 'sgd' and world > 1:
            pass   # G_ is updated inside the (replicated) dictionary phase: nothing to exchange
        if self.verbose and self.verbose_iter_ and self.n_iter_ >= self.verbose_iter_[0]:
            if rank == 0:
                print('Iteration %i' % self.n_iter_)
            self.verbose_iter_ = self.verbose_iter_[1:]
            self._callback()
        b_local = X.shape[0]
        b_global = self._global_rows(b_local, world)
        subset, sample_indices, w, order, w_sample = self._host_bookkeeping(b_global, sample_indices)
        if world > 1 and getattr(self, "overlap_exchange", False) and hasattr(self, "_overlapped_step"):
            self._overlapped_step(X, sample_indices, subset, order, w, w_sample, b_global)
        else:
            inc = self._phase_code_and_increments(X, sample_indices, subset, order, w, w_sample, b_global)
            if world > 1:
                dist.all_reduce(inc, op=dist.ReduceOp.SUM, group=self.process_group)
            self._phase_apply_and_dict(X, sample_indices, subset, order, w, w_sample, inc, b_global)
        self.__dict__["last_subset_"] = subset
        self.__dict__["last_order_"] = order

    def _global_rows(self, b_local, world):
        """Rows of the whole minibatch.  Shards must be equal (the weights w and 1/b assume it, and n_iter_ must advance
        identically on every rank): a full `batch_size` block is equal by construction; any other row count is
        verified with one small collective, and unequal shards raise instead of silently mis-weighting the step."""
        if world == 1 or b_local == self.batch_size:
            return b_local * world
        dev = getattr(self, "_device", None)
        on_gpu = dev is not None and dist.get_backend(self.process_group) == "nccl"
        t = torch.tensor([b_local, -b_local], dtype=torch.int64, device=dev if on_gpu else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.process_group)
        hi, lo = int(t[0].item()), -int(t[1].item())
        if hi != lo:
            raise ValueError("sample-sharded step: ranks hold between %d and %d rows of this minibatch; shards must be "
                             "equal (run a ragged tail through _replicated_step)" % (lo, hi))
        return b_local * world

    def _replicated_step(self, X, sample_indices):
        """One minibatch processed WHOLE by every rank, without exchange (identical inputs and kernels on every
        rank, so the replicas stay bit-identical): the un-sharded `_single_batch_fit` of the estimator below this
        mixin.  Used for ragged minibatches that do not divide over the ranks."""
        return super(ShardedStepMixin, self)._single_batch_fit(X, sample_indices)

    def check_replicas(self):
        """Debug helper: max abs difference of the dictionary across ranks (should be 0.0)."""
        world, _ = self._world()
        D = self.components_dev.clone()
        lo, hi = D.clone(), D.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.process_group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.process_group)
        return float((hi - lo).abs().max().item())


class ShardedDictFact(ShardedStepMixin, DictFact):
    """DictFact whose minibatches are sharded over the ranks of a torch.distributed group."""

    def __init__(self, process_group=None, overlap_exchange=True, **kwargs):
        DictFact.__init__(self, **kwargs)
        self.process_group = process_group
        self.overlap_exchange = overlap_exchange

    @classmethod
    def _get_param_names(cls):
        return sorted(set(DictFact._get_param_names()) | {"process_group", "overlap_exchange"})

    # -- the C minibatch loop with its own NCCL communicators --------------------------------------
    # Full `batch_size` blocks go through modl_partial_fit_* (one C call per partial_fit: the loop, the host
    # bookkeeping, both streams and both all-reduces live in the library; include/modl_b200.h,
    # modl_fit_set_comm); a ragged tail takes the Python walk below, which verifies that the shards are equal.
    def _c_loop_ok(self):
        if self.verbose or self.callback is not None or self.__dict__.get("python_loop", False):
            return False
        if type(self)._single_batch_fit is not ShardedStepMixin._single_batch_fit:
            return False
        world, _ = self._world()
        if world == 1:
            return True
        return (dist.get_backend(self.process_group) == "nccl" and self.G_agg != 'average' and self.Dx_agg != 'average'
                and self.optimizer != 'sgd')

    def _comm_loop(self):
        world, rank = self._world()
        loop = self._fit_loop_handle()
        if loop.world != world:
            ids = [_lib.nccl_unique_id(), _lib.nccl_unique_id()] if rank == 0 else [None, None]
            src = dist.get_global_rank(self.process_group, 0) if self.process_group is not None else 0
            dist.broadcast_object_list(ids, src=src, group=self.process_group, device=self._device)
            loop.set_comm(world, rank, ids[0], ids[1])       # collective: every rank reaches it at its first step
            loop.set_option("overlap", int(bool(self.overlap_exchange)))
        return loop

    def _partial_fit_c(self, Xt, sample_indices, code_out, stream, world=1):
        world, _ = self._world()
        if world == 1:
            return DictFact._partial_fit_c(self, Xt, sample_indices, code_out, stream)
        self._comm_loop()
        n, bs = Xt.shape[0], int(self.batch_size)
        n_full = (n // bs) * bs
        idx = None
        if sample_indices is not None:
            from .dict_fact import get_sub_slice
            idx = np.ascontiguousarray(get_sub_slice(sample_indices, slice(0, n)), dtype=np.int64)
        if n_full:
            DictFact._partial_fit_c(self, Xt[:n_full], idx[:n_full] if idx is not None else None,
                                    code_out[:n_full] if code_out is not None else None, stream, world=world)
        if n_full < n:
            if code_out is not None:
                raise ValueError("code_out with a ragged sharded tail is not supported")
            tail = Xt[n_full:]
            if not tail.is_cuda:
                tail = tail.to(self._device, non_blocking=False)
            self._single_batch_fit(tail, idx[n_full:] if idx is not None else np.arange(n_full, n))

    # -- overlapped exchange ---------------------------------------------------------------------
    # The dictionary update needs C_ and only the SUBSET columns of B_ (dict_fact.py:532), so the
    # exchange is split: a small all-reduce of [C inc | B inc[:, subset]] (1.5 MB at the benchmark
    # shape) on the compute stream, and the all-reduce of the full B increment (10 MB) plus its
    # fold-in on a side stream with its own NCCL communicator, hidden behind the sequential
    # dictionary update.
    def _overlap_state(self):
        st = self._side_state()
        if "group" not in st:
            ranks = dist.get_process_group_ranks(self.process_group) if self.process_group is not None else None
            st["group"] = dist.new_group(ranks=ranks, backend="nccl")   # collective: every rank creates it at its first step
        return st

    def _overlapped_step(self, X, sample_indices, subset, order, w, w_sample, b_global):
        st = self._overlap_state()
        main, side = torch.cuda.current_stream(self._device), st["stream"]
        k = self._d_components_.shape[0]
        inc = self._inc_buffer()
        inc_sub = self._inc_sub_buffer(subset.shape[0])
        prm = self._step_params(X, sample_indices, subset, order, w, w_sample, stats_inc=inc, global_batch=b_global,
                                inc_sub=inc_sub)
        # compute stream: codes, then this rank's share of [C inc | B inc[:, subset]] only
        self._run_phases(prm, _lib.PHASE_CODE | _lib.PHASE_STATS_SUB)
        st["ev_code"].record(main)
        dist.all_reduce(inc_sub, op=dist.ReduceOp.SUM, group=self.process_group)
        prm.ev_after_apply_sub = st["ev_sub"].cuda_event     # recorded inside the call: B_[:, subset] has been read
        self._run_phases(prm, _lib.PHASE_APPLY_SUB | _lib.PHASE_DICT | _lib.PHASE_REUSE_SUBSET)
        # side stream: the full-width product, its all-reduce and its fold-in, behind the dictionary update
        side.wait_event(st["ev_code"])
        self._run_phases(prm, _lib.PHASE_STATS_B, stream=side)
        with torch.cuda.stream(side):
            dist.all_reduce(inc[k * k:], op=dist.ReduceOp.SUM, group=st["group"])
        side.wait_event(st["ev_sub"])
        self._run_phases(prm, _lib.PHASE_APPLY_B, stream=side)
        st["ev_applied"].record(side)
        main.wait_event(st["ev_applied"])

    def _inc_buffer(self):
        D = self._d_components_
        k, p = D.shape
        buf = self.__dict__.get("_d_inc")
        if buf is None or buf.numel() != k * (k + p) or buf.dtype != D.dtype:
            buf = self.__dict__["_d_inc"] = torch.empty(k * (k + p), dtype=D.dtype, device=D.device)
        return buf

    def _phase_code_and_increments(self, X, sample_indices, subset, order, w, w_sample, b_global):
        inc = self._inc_buffer()
        self._launch_step(X, sample_indices, subset, order, w, w_sample,
                          phases=_lib.PHASE_CODE | _lib.PHASE_STATS, stats_inc=inc, global_batch=b_global)
        return inc

    def _phase_apply_and_dict(self, X, sample_indices, subset, order, w, w_sample, inc, b_global):
        self._launch_step(X, sample_indices, subset, order, w, w_sample,
                          phases=_lib.PHASE_APPLY | _lib.PHASE_DICT, stats_inc=inc, global_batch=b_global)
