#!/usr/bin/env python
"""bench.py -- samples/s of one `partial_fit` minibatch step of MODL's hot path on B200.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
    DictFact, synthetic dense planted-model X (float32), n_components=256, n_features=10000,
    batch_size=512 per GPU, reduction=8, l1 coding (code_l1_ratio=1, code_alpha=1, tol=1e-2,
    max_iter=100), state sized for n_samples=100000.
A "step" is one pass of the hot path (`DictFact._single_batch_fit`) over one batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scaling weak|strong]

N > 1 is launched by torchrun, one rank per GPU; the batch is sharded over samples, the
per-batch statistics increments are summed with one NCCL all-reduce and every rank applies the
(deterministic) dictionary update.  `--scaling weak` (default) keeps 512 samples per GPU per
step; `--scaling strong` splits one 512-sample batch over the ranks.

Rank 0 prints ONE JSON line (see the keys below).  `--impl reference` times the UNMODIFIED
reference (oracle/_ref, its own Cython + NumPy path) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K, P, B, R = 256, 10000, 512, 8
N_SAMPLES_STATE = 100000
METRIC = "samples/sec per partial_fit (k=256, p=10k, batch=512)"
EST_KW = dict(n_components=K, batch_size=B, reduction=R, code_l1_ratio=1., code_alpha=1., tol=1e-2,
              max_iter=100, Dx_agg='masked', G_agg='masked', rand_size=True, replacement=True,
              random_state=0)


def make_data(n_rows, seed=0):
    """Planted sparse model of BASELINE.md section 3 (rows beyond the first K are i.i.d.)."""
    rng = np.random.RandomState(seed)
    D0 = rng.randn(K, P).astype(np.float32)
    D0 /= np.linalg.norm(D0, axis=1, keepdims=True)
    A = (rng.randn(n_rows, K) * (rng.rand(n_rows, K) < 0.3)).astype(np.float32)
    X = A @ D0
    X += 0.1 * rng.randn(n_rows, P).astype(np.float32)
    return np.ascontiguousarray(X, dtype=np.float32)


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm
def _reference_dictfact():
    ref = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref, "modl", "decomposition", "dict_fact.py")):
        sys.path.insert(0, ref)
        from modl.decomposition.dict_fact import DictFact
        return DictFact, "reference"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    return oracle.OracleDictFact, "port"


def time_reference(X, steps, warmup, n_threads=1):
    """samples/s of the reference's own partial_fit on the host cores; X holds (warmup+steps)*B rows."""
    DictFact, kind = _reference_dictfact()
    kw = dict(EST_KW)
    if kind == "reference":
        kw["n_threads"] = n_threads
    est = DictFact(**kw)
    est.prepare(n_samples=N_SAMPLES_STATE, X=X[:K])
    for i in range(warmup):
        est.partial_fit(X[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        est.partial_fit(X[i * B:(i + 1) * B], np.arange(i * B, (i + 1) * B))
    dt = time.perf_counter() - t0
    return steps * B / dt, dt / steps * 1e3, kind


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([d.get("num_threads", 1) for d in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    X = make_data((steps + warmup) * B)
    best = None
    for nt in (1, min(32, os.cpu_count() or 1)):
        v_, ms_, kind = time_reference(X, steps, warmup, n_threads=nt)
        if best is None or v_ > best[0]:
            best = (v_, ms_, nt)
        if kind != "reference":
            break
    val, ms, nt_best = best
    cores = max(blas_threads(), nt_best)
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DictFact k=256 p=10000 batch=512 reduction=8 l1 coding (BASELINE configs[1])",
                   "n_components": K, "n_features": P, "batch_size": B, "reduction": R},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": "%d timed minibatches of 512 rows after %d warm-up, reference DictFact.partial_fit, "
                                   "best of n_threads in {1, 32} (n_threads=%d), BLAS threads=%d, %d host cores"
                                   % (steps, warmup, nt_best, blas_threads(), os.cpu_count() or 1)},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_json(out)


# --------------------------------------------------------------------------- B200 arm
def algorithmic_work(s_mean, sweeps_mean, b=B):
    """Algorithmic bytes / flops of each phase for ONE step, float32."""
    f = 4
    return {
        "gather": {"bytes": f * (b * P + b * s_mean + 2 * K * s_mean), "flops": 2 * b * P},
        "gram": {"bytes": f * ((K + b) * s_mean + K * K + b * K), "flops": 2 * b * s_mean * K + 2 * K * K * s_mean},
        "code": {"bytes": f * (K * K + 4 * b * K + b), "flops": 4 * K * K * b * sweeps_mean},
        "stats": {"bytes": f * (b * P + 2 * K * P + b * K + 2 * K * K), "flops": 2 * K * b * P + 2 * K * K * b},
        "dict_prep": {"bytes": f * (2 * K * s_mean), "flops": 0},
        "dict_bcd": {"bytes": f * (3 * K * s_mean + K * K), "flops": 2 * K * K * s_mean},
        "dict_post": {"bytes": f * (2 * K * s_mean), "flops": 0},
    }


def run_b200(args, rank, world):
    import torch
    import torch.distributed as dist
    from modl_b200 import _lib
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from modl_b200.distributed import ShardedDictFact as Est
    else:
        from modl_b200 import DictFact as Est
    steps, warmup = args.steps, args.warmup
    b_local = B if args.scaling == "weak" else B // world
    b_global = b_local * world
    total = steps + warmup
    X = make_data(total * b_local, seed=rank if world > 1 else 0)
    if world > 1:
        # the dictionary initialisation must be identical on every rank
        X0 = make_data(K, seed=0)
    else:
        X0 = X
    kw = dict(EST_KW)
    kw["batch_size"] = b_local if world > 1 else B

    if world > 1 and os.environ.get("MODL_OVERLAP") is not None:
        kw["overlap_exchange"] = bool(int(os.environ["MODL_OVERLAP"]))

    def new_est():
        est = Est(device=dev, **kw)
        est.prepare(n_samples=N_SAMPLES_STATE, X=X0[:K])
        return est

    def idx_of(i):
        # global row ids of this rank's shard of step i
        base = i * b_global + rank * b_local
        return np.arange(base, base + b_local) % N_SAMPLES_STATE

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    ctx = _lib.get_context(local_rank)

    # ---- (1) device-resident pass: inputs in HBM before the timed region -----------------
    est = new_est()
    est.record_sweeps = True
    Xd = torch.from_numpy(X).to(dev)          # (steps+warmup)*b rows: 20 MB per step, > L2 over the run
    for i in range(warmup):
        est.partial_fit(Xd[i * b_local:(i + 1) * b_local], idx_of(i))
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host0 = time.perf_counter()
    for i in range(warmup, total):
        est.partial_fit(Xd[i * b_local:(i + 1) * b_local], idx_of(i))
    host_ms = (time.perf_counter() - host0) * 1e3 / steps      # host time to ENQUEUE a step (no synchronisation)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = steps * b_global / (ms_total * 1e-3)
    sw = est.last_sweeps_
    sweeps_mean = float(sw[:b_local].mean()) if sw is not None else float("nan")
    last_idx = idx_of(total - 1)
    density = float((est.code_dev[torch.as_tensor(last_idx, device=dev)] != 0).float().mean().item())
    s_mean = float(est.last_subset_.shape[0])

    # ---- (2) end-to-end pass: host (pinned) rows in, batch code out, copies timed ----------
    if args.no_e2e:      # profiler runs only (ncu): skip the end-to-end and CPU legs
        args.no_cpu = True
    est2 = new_est()
    est2.async_host_copy = True      # pinned rows stay untouched until the final synchronisation below
    Xp = torch.from_numpy(X).pin_memory()
    code_host = torch.empty((b_local, K), dtype=torch.float32).pin_memory()
    for i in range(warmup):
        est2.partial_fit(Xp[i * b_local:(i + 1) * b_local], idx_of(i))
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    wall0 = time.perf_counter()
    for i in range(warmup, warmup + 1 if args.no_e2e else total):
        ids = idx_of(i)
        est2.partial_fit(Xp[i * b_local:(i + 1) * b_local], ids)
        code_host.copy_(est2.code_dev[ids[0]:ids[0] + b_local] if ids[-1] - ids[0] == b_local - 1
                        else est2.code_dev[torch.as_tensor(ids, device=dev)], non_blocking=True)
    t1.record()
    barrier()
    e2e_ms = max(t0.elapsed_time(t1), 0.0)
    wall_ms = (time.perf_counter() - wall0) * 1e3
    e2e_ms = max(e2e_ms, wall_ms * 0.0)   # device time on the launching stream; wall kept for reference
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = steps * b_global / (e2e_ms * 1e-3)

    # pinned host -> device bandwidth of this box (what bounds the end-to-end number once the kernels are fast)
    h2d_gbps = h2d_fresh_gbps = None
    if rank == 0:
        buf_h = Xp[:b_local]
        buf_d = torch.empty_like(buf_h, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            buf_d.copy_(buf_h, non_blocking=True)
        ev0.record()
        for _ in range(10):
            buf_d.copy_(buf_h, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize(dev)
        h2d_gbps = 10 * buf_h.numel() * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
        # the same copy over DISTINCT batches (what the end-to-end loop does: every step reads fresh host pages)
        nfresh = min(total, 16)
        ev0.record()
        for i in range(nfresh):
            buf_d.copy_(Xp[i * b_local:(i + 1) * b_local], non_blocking=True)
        ev1.record()
        torch.cuda.synchronize(dev)
        h2d_fresh_gbps = nfresh * buf_h.numel() * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9

    # ---- (3) per-phase device times (separate profiled pass, CUDA events on the stream) ----
    phases_ms, roof = None, None
    # every rank runs the pass (the sharded step contains a collective); rank 0 reports
    ctx.profile(True)
    nprof = min(steps, 10)
    for i in range(nprof):
        j = warmup + i
        est.partial_fit(Xd[j * b_local:(j + 1) * b_local], idx_of(j))
    torch.cuda.synchronize(dev)
    tot, nst = ctx.profile_read()
    ctx.profile(False)
    barrier()
    if rank == 0:
        phases_ms = {k_: v / max(nprof, 1) for k_, v in tot.items()}
        work = algorithmic_work(s_mean, sweeps_mean, b_local)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            pk = json.load(open(peaks_path))
            hbm_peak, tf_peak, which = float(pk["hbm_gbs"]), float(pk["bf16_tflops"]), "measured"
        else:
            hbm_peak, tf_peak, which = 6650.0, 1590.0, "fallback"
        dom = max((p_ for p_ in phases_ms if p_ in work), key=lambda p_: phases_ms[p_])
        dom_ms = phases_ms[dom]
        achieved = work[dom]["bytes"] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        prof_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(prof_json):
            traffic = json.load(open(prof_json)).get(dom)
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": which,
                "ms_per_launch": dom_ms,
                "note": "the two dominant phases (code = warp-per-sample CD, dict_bcd = sequential atoms) are "
                        "dependency-chain bound, not HBM or tensor bound (SURVEY H4); per-phase table in `phases`",
                "phases": {p_: {"ms": phases_ms[p_],
                                "GBps": work[p_]["bytes"] / (phases_ms[p_] * 1e-3) / 1e9 if phases_ms[p_] > 0 else None,
                                "TFLOPs": work[p_]["flops"] / (phases_ms[p_] * 1e-3) / 1e12 if phases_ms[p_] > 0 else None,
                                "alg_bytes": work[p_]["bytes"], "alg_flops": work[p_]["flops"]}
                           for p_ in phases_ms if p_ in work},
                "tensor_peak_TFLOPs": tf_peak}

    # ---- (4) CPU baseline: the reference's own path on this box's host cores (rank 0, N=1) --
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nb_cpu, warm_cpu = args.cpu_steps, 1
        Xc = X[:(nb_cpu + warm_cpu) * B] if X.shape[0] >= (nb_cpu + warm_cpu) * B else make_data((nb_cpu + warm_cpu) * B)
        best = None
        for nt in (1, min(32, os.cpu_count() or 1)):
            v_, ms_, kind = time_reference(Xc, nb_cpu, warm_cpu, n_threads=nt)
            if best is None or v_ > best[0]:
                best = (v_, ms_, nt)
            if kind != "reference":
                break
        v, ms_c, nt_best = best
        cpu = {"value": v, "unit": "samples/s", "cores": max(blas_threads(), nt_best), "kind": kind, "ms_per_step": ms_c,
               "sample": "%d timed minibatches of 512 rows (same data, same seeds) after %d warm-up; reference "
                         "DictFact.partial_fit, best of n_threads in {1, 32} (n_threads=%d), BLAS threads=%d, "
                         "%d host cores" % (nb_cpu, warm_cpu, nt_best, blas_threads(), os.cpu_count() or 1)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "host_enqueue_ms_per_step": host_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "DictFact k=256 p=10000 batch=512 reduction=8 l1 coding (BASELINE configs[1])",
                       "n_components": K, "n_features": P, "batch_size_per_gpu": b_local, "global_batch": b_global,
                       "reduction": R, "code_alpha": 1.0, "tol": 1e-2, "n_samples_state": N_SAMPLES_STATE,
                       "parallelism": "sample-sharded dp%d" % world,
                       "l2": "inputs larger than L2: a fresh 20.5 MB batch per step, %d MB over the timed region"
                             % int(steps * b_local * P * 4 / 1e6),
                       "mean_cd_sweeps": sweeps_mean, "code_density": density, "subset_len_last": s_mean},
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms / steps,
                    "h2d_bytes_per_step": int(b_local * P * 4 + b_local * 8), "d2h_bytes_per_step": int(b_local * K * 4),
                    "api": "DictFact(async_host_copy=True).partial_fit(pinned host rows, sample_indices), one batch per "
                           "call, + read-back of the batch code into pinned memory every step",
                    "pinned_h2d_GBps": h2d_gbps, "pinned_h2d_distinct_batches_GBps": h2d_fresh_gbps},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print_json(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner, ...)
    # goes to stderr instead
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global print_json

    def print_json(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-steps", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiler runs: shortest possible e2e leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        if world != args.gpus and rank == 0:
            sys.stderr.write("note: WORLD_SIZE=%d, --gpus=%d; using WORLD_SIZE\n" % (world, args.gpus))
        run_b200(args, rank, world)


if __name__ == "__main__":
    main()
