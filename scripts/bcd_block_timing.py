"""Debug: per-phase clock cycles inside bcd_block_kernel (CTA 0) at config-2 shape."""
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
ctx.set_option("bcd_timing", 1)
est = DictFact(**EST_KW)
est.prepare(n_samples=4 * B, X=X[:K])
Xd = torch.from_numpy(X).cuda()
for i in range(3):
    est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
nbk = K // 8
buf = (C.c_longlong * (nbk * 10))()
L = _lib.lib()
L.modl_debug_bcd_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.modl_debug_bcd_stamps.restype = C.c_int
print("status", L.modl_debug_bcd_stamps(ctx.handle, buf, nbk * 10))
t = np.array(list(buf), dtype=np.int64).reshape(nbk, 10)
blk = t[2:-2]
print("block period (cycles): mean %.0f" % np.diff(t[:, 0]).mean())
for name, a, b in (("wait+tables+product", 0, 1), ("A basis rows", 1, 2), ("B gram", 2, 3), ("cluster barrier", 3, 4),
                   ("dsmem reduce", 4, 5), ("solve", 5, 6), ("apply", 6, 7)):
    print("%-22s %6.0f" % (name, (blk[:, b] - blk[:, a]).mean()))
print("  B: load_block issue %6.0f, tiles (warp 0) %6.0f, rest (sync + gloc reduce + sync) %6.0f" % (
    (blk[:, 8] - blk[:, 2]).mean(), (blk[:, 9] - blk[:, 8]).mean(), (blk[:, 3] - blk[:, 9]).mean()))
