"""Why is the step slower when host rows stream in?  The device-resident loop timed (1) alone, and with an UNRELATED
copy loop saturating (2) PCIe host->device, (3) PCIe device->host, (4) the copy engine device->device; then the
one-stream per-phase profile (CUDA events inside the library) with and without the host->device background."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_data, EST_KW, K, B, N_SAMPLES_STATE
from modl_b200 import DictFact
import modl_b200

steps = 40
X = make_data((steps + 6) * B)
Xd = torch.from_numpy(X).cuda()
dev = torch.device("cuda", 0)
nbytes = B * X.shape[1] * 4
hbuf = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
dbuf = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
dbuf2 = torch.empty_like(dbuf)
bg = torch.cuda.Stream()


def background(kind, n):
    with torch.cuda.stream(bg):
        for _ in range(n):
            if kind == "h2d":
                dbuf.copy_(hbuf, non_blocking=True)
            elif kind == "d2h":
                hbuf.copy_(dbuf, non_blocking=True)
            elif kind == "d2d":
                for _ in range(40):
                    dbuf2.copy_(dbuf, non_blocking=True)


est = DictFact(async_host_copy=True, **EST_KW)
est.prepare(n_samples=N_SAMPLES_STATE, X=X[:K])
est.device_rows_final = True
for i in range(6):
    est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
est.synchronize()


def loop():
    for i in range(6, 6 + steps):
        est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))


def timed(label, kind):
    torch.cuda.synchronize()
    if kind:
        background(kind, 60)            # 60 x 0.37 ms queued ahead: covers the whole loop
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loop(); e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t0 = time.perf_counter(); torch.cuda.synchronize(); left = time.perf_counter() - t0
    print("%-46s %.4f ms/step   (background still running %.1f ms after the loop)" % (label, ms, left * 1e3), flush=True)


for rep in range(2):
    timed("(1) device-resident loop alone", None)
    timed("(2) + unrelated pinned host->device copies", "h2d")
    timed("(3) + unrelated device->pinned host copies", "d2h")
    timed("(4) + unrelated device->device copies", "d2d")

ctx = modl_b200._lib.get_context(0)
for kind in (None, "h2d", None, "h2d"):
    for i in range(2):
        ctx.profile(True)
        est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
        torch.cuda.synchronize()
        ctx.profile(False)
    torch.cuda.synchronize()
    if kind:
        background(kind, 60)
    ctx.profile(True)
    for i in range(20):
        est.partial_fit(Xd[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
    est.synchronize()
    tot, _ = ctx.profile_read()
    ctx.profile(False)
    torch.cuda.synchronize()
    print("phases (us/step) background=%s: %s  sum %.1f" % (kind, {k_: round(v / 20 * 1e3, 1) for k_, v in tot.items()},
                                                          sum(tot.values()) / 20 * 1e3), flush=True)
