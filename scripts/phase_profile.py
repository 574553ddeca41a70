"""Per-phase device times of the minibatch step (in-library CUDA-event profiler, `modl_ctx_profile`) at the
shapes of the "next" rows: python scripts/phase_profile.py [fmri] [image] [--out file.json] [--coop-min-cols N] [--flag-barrier]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modl_b200 import DictFact, _lib  # noqa: E402

SHAPES = {
    # fmri.py:481-495 keywords at BASELINE configs[3]; image.py:96-114 ('NMF') at configs[2]
    "fmri": dict(p=200000, b=100, steps=12, kw=dict(n_components=70, code_alpha=1e-4, code_l1_ratio=0, comp_l1_ratio=1,
                                                   reduction=12, learning_rate=0.92, batch_size=100)),
    "image": dict(p=57344, b=200, steps=12, kw=dict(n_components=256, code_alpha=0.1, code_l1_ratio=1, comp_l1_ratio=0,
                                                   code_pos=True, comp_pos=True, reduction=10, learning_rate=0.92,
                                                   batch_size=200, tol=1e-2)),
}


def run(name):
    cfg = SHAPES[name]
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    p, b, k = cfg["p"], cfg["b"], cfg["kw"]["n_components"]
    maps = torch.randn(40, p, device=dev, generator=g) * (torch.rand(40, p, device=dev, generator=g) < 0.05)
    n = b * (cfg["steps"] + 3)
    X = torch.randn(n, 40, device=dev, generator=g) @ maps + 0.1 * torch.randn(n, p, device=dev, generator=g)
    if name == "image":
        X = X.abs()
        X /= X.norm(dim=1, keepdim=True)
    est = DictFact(random_state=0, **cfg["kw"])
    est.prepare(n_samples=n, X=X[:k] if name == "image" else None, n_features=p, dtype=np.float32)
    est.partial_fit(X[:3 * b])
    torch.cuda.synchronize()
    ctx = _lib.get_context(0)
    if "--flag-barrier" in sys.argv:         # per-CTA epoch flags instead of the atomic counter (round-2 validation)
        ctx.set_option("bcd_flag_barrier", 1)
        est.partial_fit(X[:b])
        torch.cuda.synchronize()
    if "--coop-min-cols" in sys.argv:        # sweep: fewer, wider CTAs at the grid barrier of the dictionary update
        ctx.set_option("bcd_coop_min_cols", int(sys.argv[sys.argv.index("--coop-min-cols") + 1]))
        est.partial_fit(X[:b])
        torch.cuda.synchronize()
    ctx.profile(True)
    t = time.perf_counter()
    est.partial_fit(X[3 * b:])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t
    tot, nst = ctx.profile_read()
    ctx.profile(False)
    return dict(shape=name, p=p, k=k, batch=b, steps=nst, wall_ms_per_step=1e3 * wall / cfg["steps"],
                phase_ms_per_step={ph: ms / max(nst, 1) for ph, ms in tot.items()})


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in SHAPES] or list(SHAPES)
    out = [run(nm) for nm in names]
    for o in out:
        print(json.dumps(o))
    if "--out" in sys.argv:
        with open(sys.argv[sys.argv.index("--out") + 1], "w") as f:
            json.dump(out, f, indent=1)
