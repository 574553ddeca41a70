"""Debug: timeline of the two-stream minibatch loop (modl_fit_trace) at the bench shape.
    python scripts/loop_trace.py [device|pinned] [steps]
Prints, per step, when each point of the schedule was reached (ms, relative to the step's critical-path start)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import make_data, EST_KW, K, B
from modl_b200 import DictFact

mode = sys.argv[1] if len(sys.argv) > 1 else "device"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
X = make_data((steps + 6) * B)
rows = torch.from_numpy(X).cuda() if mode == "device" else torch.from_numpy(X).pin_memory()
est = DictFact(async_host_copy=True, **EST_KW)
est.prepare(n_samples=X.shape[0], X=X[:K])
est.device_rows_final = os.environ.get("MODL_TRACE_FENCE", "0") != "1"     # rows uploaded once, above
out = torch.empty((B, K), dtype=torch.float32).pin_memory() if mode != "device" else None
for i in range(6):
    est.partial_fit(rows[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B), code_out=out)
est.synchronize()
loop = est._fit_loop_handle()
loop.trace(steps)
for i in range(6, 6 + steps):
    est.partial_fit(rows[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B), code_out=out)
est.synchronize()
tr = loop.trace_read()
loop.trace(0)
names = loop.TRACE_POINTS
print("mode", mode, "| columns: absolute main_start, then each point relative to it (ms)")
prev = None
for i, row in enumerate(tr):
    base = row.get("main_start", float("nan"))
    period = "" if prev is None else " period %.3f" % (base - prev)
    prev = base
    print("step %2d  main_start %8.3f%s | " % (i, base, period) +
          "  ".join("%s %+.3f" % (n, row[n] - base) for n in names if n in row and n != "main_start"))
