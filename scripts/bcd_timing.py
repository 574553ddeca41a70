"""Debug: clock cycles of the cluster dictionary-update kernels at the config-2 shape (CTA 0 stamps): the block-wise
coefficient-space kernel (per-phase breakdown of a block of 16 atoms), the look-ahead pilot with the pipelined norm
exchange, and the pilot without it (per-stage breakdown: 0-1 candidate, 1-2 warp sums, 2-3 send, 3-4 wait, 4-5 sums,
5-6 projection)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from modl_b200 import _lib
from bench import make_data, EST_KW, K, B
from modl_b200 import DictFact

X = make_data(4 * B)
ctx = _lib.get_context(0)
L = _lib.lib()
L.modl_debug_bcd_timing.argtypes = [C.c_void_p, C.c_void_p]
L.modl_debug_bcd_timing.restype = C.c_int
L.modl_debug_bcd_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.modl_debug_bcd_stamps.restype = C.c_int
print("library:", _lib.LIB_PATH)
Xd = torch.from_numpy(X).cuda()
for name, blocked, pipe in (("blocked", 1, 1), ("pipelined pilot", 0, 1), ("pilot", 0, 0)):
    ctx.set_option("bcd_blocked", blocked)
    ctx.set_option("bcd_pipeline", pipe)
    ctx.set_option("bcd_timing", 1)
    est = DictFact(**EST_KW)
    est.python_loop = True          # one stream, one fused call per step
    est.prepare(n_samples=4 * B, X=X[:K])
    for i in range(3):
        est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
    torch.cuda.synchronize()
    NST = 8 * K + 16 + 2 * ((K + 7) // 8) + 8
    st = (C.c_longlong * NST)()
    L.modl_debug_bcd_stamps(ctx.handle, st, NST)
    g = [st[8 * K + i] for i in range(8)]
    if blocked:
        nbk = (K + 15) // 16
        ph = np.array([[st[8 * b + i] for i in range(8)] for b in range(nbk)], dtype=np.int64)
        nxt = np.concatenate([ph[1:, 0], [g[3]]])              # start of the next block = end of S4 (after its barrier)
        iv = np.column_stack([np.diff(ph[:, :8], axis=1), nxt - ph[:, 6]])[1:-1]     # blocks 1 .. nbk-2
        names = ["S0a tables+apply", "S0b repair+basis", "S1 loads+gram+push", "S2a reduce-scatter wait+sum+push",
                 "S2b all-gather wait", "S3 unpack", "S4 solver", "S4 + wait for the workers"]
        W = 8 * K + 16
        la = np.array([st[W + 4 * b + 1] - st[W + 4 * b] for b in range(1, nbk - 1)])
        print(name, "| kernel %d cycles: loads %d, first product %d, blocks %d, last apply %d, norms+write-back %d, "
              "fence+cluster barrier %d, comp_norm %d" % (g[7] - g[0], g[1] - g[0], g[2] - g[1], g[3] - g[2], g[4] - g[3],
                                                          g[5] - g[4], g[6] - g[5], g[7] - g[6]))
        per_block = (ph[-1, 0] - ph[1, 0]) / (nbk - 2)
        print("   cycles per block of 16: %.0f  (%.0f per atom)" % (per_block, per_block / 16))
        print("   phases (mean over blocks 1..%d):" % (nbk - 2), {n: int(round(v)) for n, v in zip(names, list(iv.mean(axis=0)))},
              "look-ahead of the next block (worker warps) %d" % la.mean())
    else:
        t0 = [st[8 * t] for t in range(K)]
        per = (t0[K - 1] - t0[8]) / (K - 9)
        print(name, "| cycles per atom (atoms 8..%d): %.0f" % (K - 1, per),
              " kernel prologue %d, loop %d, epilogue %d cycles" % (g[1] - g[0], g[2] - g[1], g[4] - g[2]))
        if not pipe:
            gaps = (C.c_double * 7)()
            L.modl_debug_bcd_timing(ctx.handle, gaps)
            names = ["0-1", "1-2", "2-3", "3-4", "4-5", "5-6", "6-7"]
            print("   stages", {n: round(g_) for n, g_ in zip(names, gaps)}, "total", round(sum(gaps)))
ctx.set_option("bcd_timing", 0)
ctx.set_option("bcd_pipeline", 1)
ctx.set_option("bcd_blocked", 1)
