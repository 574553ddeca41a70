"""torchrun worker for tests/test_gpu_distributed.py: sample-sharded ShardedDictFact on N GPUs
against the single-GPU DictFact run on the concatenated batch (rank 0 computes the latter)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from modl_b200 import DictFact
    from modl_b200.distributed import ShardedDictFact

    rng = np.random.RandomState(0)
    k, p, b_local, steps = 64, 2000, 96, 8      # the C loop replays CUDA graphs from its fifth step on
    n = b_local * world * steps
    D0 = rng.randn(k, p)
    D0 /= np.linalg.norm(D0, axis=1, keepdims=True)
    A = rng.randn(n, k) * (rng.rand(n, k) < 0.3)
    X = (A @ D0 + 0.1 * rng.randn(n, p)).astype(np.float32)
    out = {}
    # overlap=True: split exchange (small all-reduce first, the full B increment on a side stream)
    for name, kw, overlap in (("lasso_l2", dict(code_l1_ratio=1., code_alpha=0.5, comp_l1_ratio=0.), True),
                              ("lasso_l2_plain", dict(code_l1_ratio=1., code_alpha=0.5, comp_l1_ratio=0.), False),
                              ("ridge_l1", dict(code_l1_ratio=0., code_alpha=0.1, comp_l1_ratio=1.), True)):
        base = dict(n_components=k, reduction=4, random_state=0, **kw)
        est = ShardedDictFact(batch_size=b_local, overlap_exchange=overlap, **base)
        est.prepare(n_samples=n, X=X[:k])
        for t in range(steps):
            lo = t * b_local * world + rank * b_local
            idx = np.arange(lo, lo + b_local)
            est.partial_fit(X[idx], idx)
        spread = est.check_replicas()
        if rank == 0:
            ref = DictFact(batch_size=b_local * world, **base)
            ref.prepare(n_samples=n, X=X[:k])
            ref.partial_fit(X)
            err = lambda a, c: float(np.linalg.norm(a.astype(np.float64) - c) / np.linalg.norm(c))
            mine = slice(0, b_local)
            out[name] = {"spread": spread, "D": err(est.components_, ref.components_), "C": err(est.C_, ref.C_),
                         "B": err(est.B_, ref.B_), "code_rank0_rows": err(est.code_[mine], ref.code_[mine]),
                         "n_iter": [int(est.n_iter_), int(ref.n_iter_)]}
            # ... and against the UNMODIFIED reference (oracle/_ref, compiled from /root/reference) on the same rows
            ref_dir = os.path.join(ROOT, "oracle", "_ref")
            if os.path.isdir(os.path.join(ref_dir, "modl")):
                if ref_dir not in sys.path:
                    sys.path.insert(0, ref_dir)
                import modl.decomposition.dict_fact as ref_df
                gold = ref_df.DictFact(batch_size=b_local * world, **base)
                gold.prepare(n_samples=n, X=X[:k])
                gold.partial_fit(X)
                out[name].update({"ref_D": err(est.components_, gold.components_), "ref_C": err(est.C_, gold.C_),
                                  "ref_B": err(est.B_, gold.B_), "ref_code_rank0_rows": err(est.code_[mine], gold.code_[mine]),
                                  "ref_n_iter": int(gold.n_iter_)})
    if rank == 0:
        print("DIST_RESULT " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
