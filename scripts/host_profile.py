"""Host-side profile of the end-to-end loop (pinned rows -> partial_fit), one batch per call."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_data, EST_KW, K, B, N_SAMPLES_STATE
from modl_b200 import DictFact
steps = 60
X = make_data((steps + 3) * B)
est = DictFact(async_host_copy=True, **EST_KW)
est.prepare(n_samples=N_SAMPLES_STATE, X=X[:K])
Xp = torch.from_numpy(X).pin_memory()
for i in range(3):
    est.partial_fit(Xp[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
torch.cuda.synchronize()
def loop():
    for i in range(3, 3 + steps):
        est.partial_fit(Xp[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable(); loop(); pr.disable()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host loop %.3f ms/step, +sync %.3f ms/step" % ((t1 - t0) / steps * 1e3, (t2 - t0) / steps * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
# variants without the profiler: (a) partial_fit only, (b) + D2H read-back of the batch code
dev = torch.device("cuda", 0)
code_host = torch.empty((B, K), dtype=torch.float32).pin_memory()
for variant in ("fit only", "fit + d2h", "fit + d2h, 2 batches per call"):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t0 = time.perf_counter()
    if variant.endswith("per call"):
        for i in range(3, 3 + steps, 2):
            est.partial_fit(Xp[i * B:(i + 2) * B], np.arange(i * B, (i + 2) * B))
            code_host.copy_(est.code_dev[(i + 1) * B:(i + 2) * B], non_blocking=True)
    else:
        for i in range(3, 3 + steps):
            est.partial_fit(Xp[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
            if variant != "fit only":
                code_host.copy_(est.code_dev[i * B:(i + 1) * B], non_blocking=True)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print("%-32s host %.3f ms/step, device events %.3f ms/step" % (variant, (t1 - t0) / steps * 1e3, e0.elapsed_time(e1) / steps))

