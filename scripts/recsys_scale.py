"""BASELINE configs[4] at its stated size: RecsysDictFact on a synthetic 1e6 x 1e5 ratings matrix at Movielens-10M
density (1.34 % observed: ~1.34e9 ratings), n_components = 50, alpha = 1, one epoch at batch 512 + the two refits.

    python scripts/recsys_scale.py [--rows 1000000] [--out file.json]

The matrix is generated on the device 1000 rows at a time (scripts/next_rows_bench.py::ratings_matrix) and handed to
`fit` as a SciPy CSR matrix, like a user would.  If the host does not have the memory for the CSR arrays and the copies
`fit` makes (~12 bytes per rating x 3), the row count is halved until it does; the row count actually run is reported."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import next_rows_bench as nrb
    from modl_b200.recsys import RecsysDictFact
    p, k, dens = 100000, 50, 0.0134
    n = args.rows
    need = lambda rows: rows * p * dens * 12 * 3.2 / 1e9
    avail = mem_available_gb()
    while n > 20000 and need(n) > 0.8 * avail:
        n //= 2
    out = dict(row="RecsysDictFact.fit at BASELINE configs[4]", requested_rows=args.rows, rows=n, n_features=p, n_components=k,
               density=dens, host_mem_available_gb=avail, dtype="f64")
    t = time.perf_counter()
    X = nrb.ratings_matrix(n, p, dens, 3, device=torch.device("cuda", 0))
    out["nnz"] = int(X.nnz)
    out["generate_seconds"] = time.perf_counter() - t
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    t = time.perf_counter()
    est = RecsysDictFact(n_components=k, alpha=1, batch_size=512, n_epochs=1, random_state=0).fit(X)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    out.update(value=n / dt, unit="samples/s", seconds=dt, device_peak_gb=torch.cuda.max_memory_allocated() / 1e9,
               note="whole fit: CSR upload, refit of every row, one epoch at batch 512, refit; wall clock, device synchronised")
    print(json.dumps(out), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
    # sanity: the factorisation predicts the observed ratings better than their mean
    try:
        pred = est.predict(X)
        hi = int(X.indptr[2000])
        rmse = float(np.sqrt(np.mean((pred.data[:hi] - X.data[:hi]) ** 2)))
        base = float(np.sqrt(np.mean((X.data[:hi] - X.data[:hi].mean()) ** 2)))
        out["train_rmse_first_rows"] = rmse
        out["rmse_of_the_mean"] = base
    except Exception as exc:      # noqa: BLE001
        out["predict_error"] = repr(exc)
    print(json.dumps(out), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
