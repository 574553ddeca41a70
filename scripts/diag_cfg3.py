import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import oracle
from modl_b200 import _lib
from modl_b200._util import ptr, stream_of
from modl_b200.dict_fact_fast import _enet_regression_single_gram
dev = torch.device("cuda", 0)
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-300))
p, k, b = 57344, 256, 512
n = b + k
rng = np.random.RandomState(12)
D0 = np.abs(rng.randn(k, p)).astype(np.float32); D0 /= np.linalg.norm(D0, axis=1, keepdims=True)
A = (np.abs(rng.randn(n, k)) * (rng.rand(n, k) < 0.2)).astype(np.float32)
X = (A @ D0 + 0.01 * np.abs(rng.randn(n, p)).astype(np.float32))[k:]
D = X[:0]
D = (A @ D0 + 0.01 * np.abs(rng.randn(n, p)).astype(np.float32))[:k] if False else D0.copy()
subset = rng.permutation(p)[:7168].astype(np.int64)
ctx = _lib.get_context(0)
for tc in (1, 0):
    ctx.set_option("tc_gemm", tc)
    Dd, Xd, sd = (torch.from_numpy(a).to(dev) for a in (D, X, subset))
    G = torch.zeros((k, k), dtype=torch.float32, device=dev); Dx = torch.zeros((b, k), dtype=torch.float32, device=dev)
    xn = torch.zeros((b,), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().modl_gram_dx_f32(ctx.handle, ptr(Dd), p, ptr(Xd), p, ptr(sd), len(subset), k, b, p, 8.0, ptr(G), ptr(Dx), ptr(xn), stream_of(dev)))
    torch.cuda.synchronize()
    Ds, Xs = D[:, subset].astype(np.float64), X[:, subset].astype(np.float64)
    G64, Dx64 = 8 * Ds @ Ds.T, 8 * Xs @ Ds.T
    print("tc=%d  G %.3g  Dx %.3g  xn %.3g" % (tc, rel(G.cpu().numpy(), G64), rel(Dx.cpu().numpy(), Dx64), rel(xn.cpu().numpy(), (X.astype(np.float64) ** 2).sum(1))))
    Gh, Dxh = G.cpu().numpy(), Dx.cpu().numpy()
    # CD from the device's G / Dx vs the oracle on the SAME inputs
    for tag, (Gi, Dxi) in (("device G,Dx", (Gh, Dxh)), ("f64-rounded G,Dx", (G64.astype(np.float32), Dx64.astype(np.float32)))):
        code_o = np.ones((b, k), np.float32); code_d = np.ones((b, k), np.float32)
        oracle.enet_regression_single_gram(Gi.copy(), Dxi.copy(), X, code_o, np.arange(b), 1., 0.1, True, 1e-2, 100)
        _enet_regression_single_gram(Gi.copy(), Dxi.copy(), X, code_d, np.arange(b), 1., 0.1, True, 1e-2, 100)
        print("   CD on %-18s: device vs oracle %.3g" % (tag, rel(code_d, code_o.astype(np.float64))))
    if tc == 1:
        keepG, keepDx = Gh, Dxh
    else:
        co1 = np.ones((b, k), np.float32); co2 = np.ones((b, k), np.float32)
        oracle.enet_regression_single_gram(keepG.copy(), keepDx.copy(), X, co1, np.arange(b), 1., 0.1, True, 1e-2, 100)
        oracle.enet_regression_single_gram(Gh.copy(), Dxh.copy(), X, co2, np.arange(b), 1., 0.1, True, 1e-2, 100)
        print("   oracle CD: tensor-core inputs vs FFMA inputs %.3g" % rel(co1, co2.astype(np.float64)))
