"""Plain launches vs graph-replayed calls (modl_fit_set_option "graph" 0 / 1 / 2) of the device-resident minibatch loop at
shapes and settings other than the bench's: ms per step (CUDA events around 30 steps after 8 warm-up steps)."""
import os
import sys
import json

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from modl_b200 import DictFact

CASES = [
    ("config2 f32 (bench)", dict(k=256, p=10000, b=512, r=8, dtype=np.float32, kw=dict(code_l1_ratio=1., code_alpha=1., tol=1e-2))),
    ("config2 f64", dict(k=256, p=10000, b=512, r=8, dtype=np.float64, kw=dict(code_l1_ratio=1., code_alpha=1., tol=1e-2))),
    ("config2 f32, L1-ball atoms", dict(k=256, p=10000, b=512, r=8, dtype=np.float32,
                                        kw=dict(code_l1_ratio=1., code_alpha=1., tol=1e-2, comp_l1_ratio=1.))),
    ("config2 f32, non-negative", dict(k=256, p=10000, b=512, r=8, dtype=np.float32,
                                       kw=dict(code_l1_ratio=1., code_alpha=1., tol=1e-2, code_pos=True, comp_pos=True))),
    ("config2 f32, ridge codes", dict(k=256, p=10000, b=512, r=8, dtype=np.float32, kw=dict(code_l1_ratio=0., code_alpha=1.))),
    ("config2 f32, average modes", dict(k=64, p=4000, b=256, r=4, dtype=np.float32,
                                        kw=dict(code_l1_ratio=1., code_alpha=0.5, Dx_agg='average', G_agg='average'))),
    ("config1 ridge r=1", dict(k=16, p=500, b=100, r=1, dtype=np.float64, kw=dict(code_l1_ratio=0., code_alpha=1.))),
    ("small f32", dict(k=64, p=2000, b=120, r=4, dtype=np.float32, kw=dict(code_l1_ratio=1., code_alpha=0.4))),
]


def run(case, graph):
    k, p, b, r, dt = case["k"], case["p"], case["b"], case["r"], case["dtype"]
    steps, warm = 30, 8
    rng = np.random.RandomState(0)
    n = (steps + warm) * b
    D0 = rng.randn(k, p); D0 /= np.linalg.norm(D0, axis=1, keepdims=True)
    A = rng.randn(n, k) * (rng.rand(n, k) < 0.1)
    X = (A @ D0 + 0.1 * rng.randn(n, p)).astype(dt)
    if case["kw"].get("code_pos"):
        X = np.abs(X)
    Xd = torch.from_numpy(X).cuda()
    est = DictFact(n_components=k, batch_size=b, reduction=r, random_state=0, **case["kw"])
    est.prepare(n_samples=n, X=X[:k])
    est.device_rows_final = True
    est._fit_loop_handle().set_option("graph", graph)
    for i in range(warm):
        est.partial_fit(Xd[i * b:(i + 1) * b], np.arange(i * b, (i + 1) * b))
    est.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + steps):
        est.partial_fit(Xd[i * b:(i + 1) * b], np.arange(i * b, (i + 1) * b))
    e1.record()
    est.synchronize()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, est._fit_loop_handle().graph_stats()


out = []
for name, case in CASES:
    row = {"case": name}
    for g in (0, 1, 2):
        try:
            ms, st = run(case, g)
            row["graph%d_ms" % g] = round(ms, 4)
            row["graph%d_stats" % g] = st
        except Exception as exc:          # a setting the estimator refuses is not a scheduling result
            row["graph%d_ms" % g] = repr(exc)[:120]
    print(json.dumps(row), flush=True)
    out.append(row)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
