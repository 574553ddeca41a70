"""torchrun worker for tests/test_gpu_distributed.py::test_sharded_fmri_loop: the fMRI learning loop sharded over
N GPUs (`_compute_components(sharded=True)`) against the golden maps of the unmodified reference."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from modl_b200 import fmri
    from test_fmri_frontend import _records

    gold = np.load(os.path.join(ROOT, "tests", "golden", "fmri.npz"))
    spec = json.loads(str(gold["cases"]))
    out = {}
    for ci in (0, 4, 5, 7):            # masked f64, dictionary only + positive, sgd, masked f32
        case = spec["cases"][ci]
        records, n_voxels = _records(case["dtype"])
        masker = fmri.RecordMasker(mask=np.ones(n_voxels, dtype=bool)).fit()
        kw = dict(spec["common"])
        kw.update(case["kw"])
        comp, est = fmri._compute_components(masker, records, sharded=True, return_estimator=True, **kw)
        spread = est.check_replicas()
        want = gold["fit_%d_components" % ci]
        err = float(np.linalg.norm(comp.astype(np.float64) - want) / np.linalg.norm(want))
        out["case_%d" % ci] = {"spread": spread, "err": err, "dtype": case["dtype"]}
    if rank == 0:
        print("DIST_RESULT " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
