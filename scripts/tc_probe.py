#!/usr/bin/env python
"""Bring-up probe of the tcgen05 GEMM: runs modl_gram_dx_f32 / modl_update_stats_f32 on a few
shapes in a child process per descriptor mode and prints relative errors against float64."""
import os
import subprocess
import sys

CHILD = r'''
import numpy as np, torch, sys
sys.path.insert(0, ".")
from modl_b200 import _lib
from modl_b200._util import ptr, stream_of
dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)
def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-300))
ctx = _lib.get_context(0)
for (k, b, p, s) in ((128, 128, 64, 32), (16, 10, 50, 20), (70, 33, 301, 77), (256, 512, 10000, 1250)):
    D = rng.randn(k, p).astype(np.float32); X = rng.randn(b, p).astype(np.float32)
    subset = rng.permutation(p)[:s].astype(np.int64)
    Dd, Xd, sd = (torch.from_numpy(a).to(dev) for a in (D, X, subset))
    G = torch.zeros((k, k), dtype=torch.float32, device=dev); Dx = torch.zeros((b, k), dtype=torch.float32, device=dev)
    xn = torch.zeros((b,), dtype=torch.float32, device=dev)
    fn = _lib.lib().modl_gram_dx_f32
    _lib.check(fn(ctx.handle, ptr(Dd), p, ptr(Xd), p, ptr(sd), s, k, b, p, 3.0, ptr(G), ptr(Dx), ptr(xn), stream_of(dev)))
    torch.cuda.synchronize()
    Ds, Xs = D[:, subset].astype(np.float64), X[:, subset].astype(np.float64)
    print("gram_dx k=%d b=%d p=%d s=%d: G %.3g  Dx %.3g  xn %.3g" % (k, b, p, s, rel(G.cpu().numpy(), 3 * Ds @ Ds.T),
          rel(Dx.cpu().numpy(), 3 * Xs @ Ds.T), rel(xn.cpu().numpy(), (X.astype(np.float64) ** 2).sum(1))), flush=True)
    if k == 128 and rel(G.cpu().numpy(), 3 * Ds @ Ds.T) > 1e-3:
        Gh = G.cpu().numpy(); W = 3 * Ds @ Ds.T
        print("  G[0,:4]", Gh[0, :4], "want", W[0, :4]); print("  G[:4,0]", Gh[:4, 0], "want", W[:4, 0])
        print("  diag", np.diag(Gh)[:6], "want", np.diag(W)[:6])
for (n, b, k, p) in ((50, 24, 40, 333), (1000, 512, 256, 10000)):
    code = rng.randn(n, k).astype(np.float32); idx = rng.permutation(n)[:b].astype(np.int64)
    X = rng.randn(b, p).astype(np.float32); C0 = rng.randn(k, k).astype(np.float32); B0 = rng.randn(k, p).astype(np.float32)
    cd, id_, Xd, Cd, Bd = (torch.from_numpy(a).to(dev) for a in (code, idx, X, C0, B0))
    _lib.check(_lib.lib().modl_update_stats_f32(ctx.handle, ptr(cd), ptr(id_), ptr(Xd), p, ptr(Cd), ptr(Bd), p, 0.3, b, k, p, 0, stream_of(dev)))
    torch.cuda.synchronize()
    cb = code[idx].astype(np.float64)
    print("stats b=%d k=%d p=%d: C %.3g  B %.3g" % (b, k, p, rel(Cd.cpu().numpy(), 0.7 * C0 + 0.3 / b * cb.T @ cb),
          rel(Bd.cpu().numpy(), 0.7 * B0 + 0.3 / b * cb.T @ X)), flush=True)
'''
for mode in (0,):
    env = dict(os.environ, MODL_TC_GEMM="1")
    try:
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, timeout=120, capture_output=True, text=True)
        print(r.stdout[-3000:])
        print(r.stderr[-1500:])
    except subprocess.TimeoutExpired as e:
        print("TIMEOUT", (e.stdout or b"")[-2000:])
