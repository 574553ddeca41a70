"""Debug: clock cycles per atom of the cluster dictionary-update kernel at the config-2 shape (CTA 0 stamps), for the
kernel variants: look-ahead pilot with the pipelined norm exchange, without it, and the per-stage breakdown of the
latter (0-1 candidate, 1-2 warp sums, 2-3 send, 3-4 wait, 4-5 sums, 5-6 projection)."""
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
for pipe in (1, 0):
    ctx.set_option("bcd_pipeline", pipe)
    ctx.set_option("bcd_timing", 1)
    est = DictFact(**EST_KW)
    est.python_loop = True          # one stream, one fused call per step
    est.prepare(n_samples=4 * B, X=X[:K])
    Xd = torch.from_numpy(X).cuda()
    for i in range(3):
        est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
    torch.cuda.synchronize()
    st = (C.c_longlong * (8 * K + 16))()
    L.modl_debug_bcd_stamps(ctx.handle, st, 8 * K + 16)
    t0 = [st[8 * t] for t in range(K)]
    per = (t0[K - 1] - t0[8]) / (K - 9)
    print("pipeline", pipe, "cycles per atom (atoms 8..%d): %.0f" % (K - 1, per),
          " kernel prologue %d, loop %d cycles" % (st[8 * K + 1] - st[8 * K], st[8 * K + 2] - st[8 * K + 1]))
    if not pipe:
        gaps = (C.c_double * 7)()
        L.modl_debug_bcd_timing(ctx.handle, gaps)
        names = ["0-1", "1-2", "2-3", "3-4", "4-5", "5-6", "6-7"]
        print("   stages", {n: round(g) for n, g in zip(names, gaps)}, "total", round(sum(gaps)))
ctx.set_option("bcd_timing", 0)
ctx.set_option("bcd_pipeline", 1)
