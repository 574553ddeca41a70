"""Throughput of the three "next" rows at the shapes of BASELINE configs[2..4], next to the
unmodified reference (oracle/_ref) timed on the box's host cores on a bounded sample of the same data.

    python scripts/next_rows_bench.py --out gpurun_out/<tag>/next_rows.json [--budget 150] [--dry]

One JSON object per row.  Not the headline benchmark (bench.py is): these are first measurements of the
front-ends, wall-clock around whole `fit` calls with a device synchronize on both sides (the calls include
their host bookkeeping -- that is what a user sees).  `--dry` runs only the host-side pieces at toy sizes.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

T0 = time.perf_counter()


def elapsed():
    return time.perf_counter() - T0


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def sync():
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ fMRI
def bench_fmri(dry):
    """configs[3]: p ~ 2e5 voxels, k = 70, reduction 12; batch 100, alpha 1e-4, learning_rate 0.92 as in the
    reference's exps/exp_decompose_fmri.py:32-38."""
    import torch
    from modl_b200.fmri import RecordMasker, _compute_components
    p, k, n_rec, n_vol = (200000, 70, 4, 600) if not dry else (300, 5, 2, 30)
    kw = dict(n_components=k, batch_size=100 if not dry else 10, reduction=12, alpha=1e-4, learning_rate=0.92,
              method='masked', random_state=0)
    out = dict(row="fMRIDictFact._compute_components", shape=dict(n_voxels=p, n_components=k, reduction=12,
                                                                   batch_size=kw["batch_size"], records=n_rec,
                                                                   volumes_per_record=n_vol), dtype="f32")
    if dry:
        return out
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    maps = torch.randn(40, p, device=dev, generator=g) * (torch.rand(40, p, device=dev, generator=g) < 0.05)
    records = [torch.randn(n_vol, 40, device=dev, generator=g) @ maps + 0.1 * torch.randn(n_vol, p, device=dev, generator=g)
               for _ in range(n_rec)]
    masker = RecordMasker(mask=np.ones(p, dtype=bool)).fit()
    _compute_components(masker, [records[0][:200]], **kw)            # warm-up: workspaces, function attributes
    sync()
    t = time.perf_counter()
    comp, est = _compute_components(masker, records, return_estimator=True, **kw)
    sync()
    dt = time.perf_counter() - t
    n = n_rec * n_vol
    dev_s = est.time_
    out.update(value=n / dt, unit="samples/s", seconds=dt, samples=n, partial_fit_device_seconds=dev_s,
               value_partial_fit_only=n / dev_s if dev_s else None, nonzero_fraction=float((comp != 0).mean()))
    try:
        out["cuda_graphs"] = est._fit_loop_handle().graph_stats()
    except Exception as exc:
        out["cuda_graphs"] = repr(exc)
    # reference arm: same estimator keywords, 1 warm-up + 3 timed minibatches of host rows
    try:
        from modl.decomposition.dict_fact import DictFact as RefDictFact
        Xh = records[0][:400].cpu().numpy()
        ref = RefDictFact(n_components=k, code_alpha=1e-4, code_l1_ratio=0, comp_l1_ratio=1, reduction=12,
                          learning_rate=0.92, batch_size=100, random_state=0)
        ref.prepare(n_samples=401, n_features=p, dtype=np.float32)
        ref.partial_fit(Xh[:100])
        t = time.perf_counter()
        ref.partial_fit(Xh[100:400])
        dtr = time.perf_counter() - t
        out["cpu_reference"] = dict(value=300 / dtr, unit="samples/s", sample="3 minibatches of 100 volumes after 1 warm-up",
                                    cores=os.cpu_count())
    except Exception as exc:      # noqa: BLE001
        out["cpu_reference"] = dict(error=repr(exc))
    return out


# ------------------------------------------------------------------------------------------------ image
def bench_image(dry):
    """configs[2]: 16 x 16 patches of a 4096 x 4096 x 224 cube, k = 256; batch 200, reduction 10, alpha 0.1,
    method 'gram', setting 'NMF' as in the reference's exps/multi_decompose_images.py:34-50."""
    import torch
    from modl_b200.image import ImageDictFact
    side, ch, k, n_patches = (4096, 224, 256, 40000) if not dry else (40, 3, 4, 50)
    kw = dict(method='gram', setting='NMF', patch_size=(16, 16) if not dry else (4, 4), batch_size=200 if not dry else 10,
              reduction=10, alpha=0.1, n_components=k, max_patches=n_patches, n_epochs=1, random_state=0)
    out = dict(row="ImageDictFact.fit", shape=dict(image=[side, side, ch], patch=list(kw["patch_size"]),
                                                    n_features=kw["patch_size"][0] * kw["patch_size"][1] * ch,
                                                    n_components=k, batch_size=kw["batch_size"], reduction=10,
                                                    patches=n_patches), dtype="f32")
    if dry:
        return out
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    image = torch.rand(side, side, ch, device=dev, generator=g)
    small = image[:64, :64].contiguous()
    ImageDictFact(**dict(kw, max_patches=1000)).fit(small)            # warm-up at the same feature width
    sync()
    t = time.perf_counter()
    est = ImageDictFact(**kw).fit(image)
    sync()
    dt = time.perf_counter() - t
    dev_s = est.time_
    out.update(value=n_patches / dt, unit="samples/s", seconds=dt, samples=n_patches,
               partial_fit_device_seconds=dev_s, value_partial_fit_only=n_patches / dev_s if dev_s else None)
    try:
        out["cuda_graphs"] = est._fit_loop_handle().graph_stats()
    except Exception as exc:
        out["cuda_graphs"] = repr(exc)
    try:
        from modl.decomposition.dict_fact import DictFact as RefDictFact
        from modl_b200.image import LazyCleanPatchExtractor, _flatten_patches
        ex = LazyCleanPatchExtractor(patch_size=(16, 16), max_patches=800, random_state=0).fit(image[:256, :256].contiguous())
        rows = _flatten_patches(ex.partial_transform(batch=800), with_mean=False, with_std=True).cpu().numpy()
        ref = RefDictFact(n_components=k, code_alpha=0.1, code_l1_ratio=1, comp_l1_ratio=0, code_pos=True, comp_pos=True,
                          reduction=10, learning_rate=0.92, batch_size=200, tol=1e-2, random_state=0)
        ref.prepare(n_samples=800, X=rows[:k])
        ref.partial_fit(rows[:200])
        t = time.perf_counter()
        ref.partial_fit(rows[200:800])
        dtr = time.perf_counter() - t
        out["cpu_reference"] = dict(value=600 / dtr, unit="samples/s", sample="3 minibatches of 200 patches after 1 warm-up",
                                    cores=os.cpu_count())
    except Exception as exc:      # noqa: BLE001
        out["cpu_reference"] = dict(error=repr(exc))
    return out


# ------------------------------------------------------------------------------------------------ recsys
def ratings_matrix(n, p, density, seed, device=None):
    """Ratings 1..5 from a rank-8 model, `density` of the entries observed, no empty row; built 1000 rows at a
    time (on `device` when given) straight into CSR arrays."""
    import scipy.sparse as sp
    import torch
    device = device or torch.device("cpu")
    g = torch.Generator(device=device).manual_seed(seed)
    U = torch.rand(n, 8, device=device, generator=g, dtype=torch.float64)
    V = torch.rand(8, p, device=device, generator=g, dtype=torch.float64)
    counts, cols, vals = [], [], []
    for r0 in range(0, n, 1000):
        rows = min(1000, n - r0)
        m = torch.rand(rows, p, device=device, generator=g) < density
        m[torch.arange(rows, device=device), torch.randint(p, (rows,), device=device, generator=g)] = True
        score = U[r0:r0 + rows] @ V + 0.3 * torch.randn(rows, p, device=device, generator=g, dtype=torch.float64)
        r, c = torch.nonzero(m, as_tuple=True)                  # row-major: already CSR order
        counts.append(m.sum(dim=1).cpu().numpy())
        cols.append(c.to(torch.int32).cpu().numpy())
        vals.append(torch.clamp(torch.round(1 + score[r, c]), 1, 5).cpu().numpy())
    indptr = np.concatenate([[0], np.cumsum(np.concatenate(counts))]).astype(np.int64)
    return sp.csr_matrix((np.concatenate(vals), np.concatenate(cols), indptr), shape=(n, p))


def bench_recsys(dry):
    """configs[4] scaled in the number of rows: p = 1e5 items at Movielens-10M density (1.34 %), k = 50, alpha 1
    (examples/predict_recsys.py:41-45); one epoch, batch sizes 10 (the example's) and 512."""
    import torch
    from modl_b200.recsys import RecsysDictFact
    n, p, k, dens = (20000, 100000, 50, 0.0134) if not dry else (60, 200, 4, 0.1)
    out = dict(row="RecsysDictFact.fit", shape=dict(n_samples=n, n_features=p, density=dens, n_components=k), dtype="f64")
    X = ratings_matrix(n, p, dens, 3, device=None if dry else torch.device("cuda", 0))
    out["shape"]["nnz"] = int(X.nnz)
    n_ref = 96 if not dry else 20
    try:
        from modl.decomposition.recsys import RecsysDictFact as RefRecsys
        t = time.perf_counter()
        ref = RefRecsys(n_components=k, alpha=1, batch_size=10, n_epochs=1, random_state=0).fit(X[:n_ref])
        dtr = time.perf_counter() - t
        out["cpu_reference"] = dict(value=n_ref / dtr, unit="samples/s", cores=os.cpu_count(),
                                    sample="whole fit (refit, one epoch at batch 10, refit) of the first %d rows" % n_ref)
    except Exception as exc:      # noqa: BLE001
        ref = None
        out["cpu_reference"] = dict(error=repr(exc))
    if dry:
        return out
    # parity at this width against the reference run above
    est = RecsysDictFact(n_components=k, alpha=1, batch_size=10, n_epochs=1, random_state=0).fit(X[:n_ref])
    if ref is not None:
        out["parity_first_rows"] = dict(components=rel_err(est.components_, ref.components_),
                                        code=rel_err(est.code_, ref.code_), B=rel_err(est.B_, ref.B_))
    for bs, bk in ((512, 'batch'), (10, 'batch'), (512, 'epoch'), (10, 'epoch')):
        if elapsed() > ARGS.budget:
            break
        name = "batch_%d" % bs + ("" if bk == 'batch' else "_epoch_bookkeeping")
        try:
            sync()
            t = time.perf_counter()
            RecsysDictFact(n_components=k, alpha=1, batch_size=bs, n_epochs=1, random_state=0, bookkeeping=bk).fit(X)
            sync()
            dt = time.perf_counter() - t
            out[name] = dict(value=n / dt, unit="samples/s", seconds=dt,
                             note="whole fit: upload, refit of every row, one epoch, refit")
        except Exception as exc:      # noqa: BLE001
            out[name] = dict(error=repr(exc))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--budget", type=float, default=150., help="stop starting new sections after this many seconds")
    ap.add_argument("--dry", action="store_true")
    ap.add_argument("--only", default="")
    ARGS = ap.parse_args()
    results = []
    for name, fn in (("recsys", bench_recsys), ("fmri", bench_fmri), ("image", bench_image)):
        if ARGS.only and name not in ARGS.only.split(","):
            continue
        if elapsed() > ARGS.budget:
            results.append(dict(row=name, skipped="time budget"))
            continue
        try:
            res = fn(ARGS.dry)
        except Exception as exc:      # noqa: BLE001
            import traceback
            res = dict(row=name, error=repr(exc), traceback=traceback.format_exc()[-1500:])
        res["t_end_s"] = round(elapsed(), 1)
        results.append(res)
        print(json.dumps(res), flush=True)
        if ARGS.out:
            os.makedirs(os.path.dirname(os.path.abspath(ARGS.out)), exist_ok=True)
            with open(ARGS.out, "w") as f:
                json.dump(results, f, indent=1)
