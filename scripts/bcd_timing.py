"""Debug: per-phase clock cycles inside the dictionary-update kernel (CTA 0) at config-2 shape."""
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
for cluster in (16, 8):
    ctx.set_option("bcd_cluster", cluster)
    ctx.set_option("bcd_timing", 1)
    est = DictFact(**EST_KW)
    est.prepare(n_samples=4 * B, X=X[:K])
    Xd = torch.from_numpy(X).cuda()
    for i in range(3):
        est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
    gaps = (C.c_double * 7)()
    L = _lib.lib()
    L.modl_debug_bcd_timing.argtypes = [C.c_void_p, C.c_void_p]
    L.modl_debug_bcd_timing.restype = C.c_int
    st = L.modl_debug_bcd_timing(ctx.handle, gaps)
    names = ["0-1", "1-2", "2-3", "3-4", "4-5", "5-6", "6-7"]
    print("cluster", cluster, "status", st, {n: round(g) for n, g in zip(names, gaps)}, "total", round(sum(gaps)), "(pilot: 0-1 candidate, 1-2 warp sums, 2-3 send, 3-4 wait, 4-5 sums, 5-6 projection; stamp-to-next-atom gap not included)")
