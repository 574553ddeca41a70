"""Where does the end-to-end step go?  (A) the H2D pipeline alone, (B) kernels alone through the same
call pattern, (C) both."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_data, EST_KW, K, B, N_SAMPLES_STATE
from modl_b200 import DictFact
steps = 40
X = make_data((steps + 3) * B)
Xp = torch.from_numpy(X).pin_memory()
Xd = torch.from_numpy(X).cuda()
dev = torch.device("cuda", 0)

def timed(fn, label):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    print("%-44s host %.3f ms/step   device %.3f ms/step" % (label, (t1 - t0) / steps * 1e3, e0.elapsed_time(e1) / steps), flush=True)

# (A) copies alone, two slots, side stream, events like DictFact._partial_fit_host
cs = torch.cuda.Stream()
slots = [torch.empty((B, X.shape[1]), device=dev) for _ in range(2)]
def copies():
    main = torch.cuda.current_stream()
    for i in range(3, 3 + steps):
        with torch.cuda.stream(cs):
            slots[i & 1].copy_(Xp[i * B:(i + 1) * B], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(cs)
        main.wait_event(ev)
timed(copies, "(A) H2D pipeline alone")

def make():
    est = DictFact(async_host_copy=True, **EST_KW)
    est.prepare(n_samples=N_SAMPLES_STATE, X=X[:K])
    return est
est = make()
for i in range(3):
    est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
def dev_loop():
    for i in range(3, 3 + steps):
        est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
timed(dev_loop, "(B) device-resident rows")
est = make()
for i in range(3):
    est.partial_fit(Xp[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
def host_loop():
    for i in range(3, 3 + steps):
        est.partial_fit(Xp[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
timed(host_loop, "(C) pinned host rows, async copy")
# (D) like (C) but the copy stream has HIGH priority and the kernels run while ANOTHER process-wide copy is active
est = make()
est.partial_fit(Xp[:B], np.arange(B))
est._pipeline["copy_stream"] = torch.cuda.Stream(priority=-1)
for i in range(1, 3):
    est.partial_fit(Xp[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
timed(host_loop, "(D) same, high-priority copy stream")
