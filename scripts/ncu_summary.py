#!/usr/bin/env python
"""Summarise one gpurun_out/<tag>/ directory (scripts/gpu_round.sh) into profiles/:
   profiles/<tag>_launches.csv   the ncu launch list (gpu__time_duration per launch), verbatim
   profiles/<tag>_summary.md     per-kernel share of the step + the --set full metrics we cite
   profiles/<tag>_bench.json, <tag>_bench_ref.json
usage: python scripts/ncu_summary.py r01_v10"""
import collections
import csv
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__cluster_size']
out = ["# ncu summary %s" % tag, ""]
lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(dst, tag + "_launches.csv"))
    lines = [l for l in open(lp) if l.startswith('"')]
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for x in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', x['Kernel Name'])
        tot[name] = tot.get(name, 0) + float(x['Metric Value'])
        cnt[name] += 1
    T = sum(tot.values())
    out += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache serialised: shares only)",
            "", "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, cnt[k], v / 1e3, 100 * v / T, v / 1e3 / cnt[k]))
    out.append("")
for f in sorted(os.listdir(src)):
    if not f.endswith(".raw.csv"):
        continue
    rows = list(csv.reader(open(os.path.join(src, f))))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out += ["## `ncu --set full` %s" % f.replace(".raw.csv", ""), ""]
    for r in rows[2:]:
        out.append("**%s**" % r[idx['Kernel Name']][:120])
        out.append("")
        out += ["| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in idx:
                out.append("| %s | %s | %s |" % (w, r[idx[w]], units[idx[w]]))
        if 'dram__bytes_read.sum' in idx:
            def val(name):
                v, u = float(r[idx[name]].replace(',', '')), units[idx[name]]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            out.append("| traffic (dram read + write) | %.3f | MB |" % ((val('dram__bytes_read.sum') + val('dram__bytes_write.sum')) / 1e6))
        out.append("")
for f in ("bench.json", "bench_ref.json"):
    if os.path.exists(os.path.join(src, f)) and os.path.getsize(os.path.join(src, f)) > 0:
        shutil.copy(os.path.join(src, f), os.path.join(dst, tag + "_" + f))
# per-phase DRAM traffic of one step (bytes), summed over the captured launches of the phase's kernels
import json
PHASE_OF = [("bcd_blocked", "dict_bcd"), ("bcd_pilot", "dict_bcd"), ("cd_regression", "code"), ("tc_pack_rows", "gather"),
            ("tc_pack_cols", "stats_sub"), ("tc_gemm", None)]
GEMM_ORDER = ["gram", "stats_sub", "stats"]     # launches of one step of the two-stream loop, in host order
traffic, gemm_seen = {}, 0
for f in sorted(os.listdir(src)):
    if not f.endswith(".raw.csv"):
        continue
    rows = list(csv.reader(open(os.path.join(src, f))))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    if 'dram__bytes_read.sum' not in idx:
        continue
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        def val(n):
            v, u = float(r[idx[n]].replace(',', '')), units[idx[n]]
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        byt = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
        for key, ph in PHASE_OF:
            if key in name:
                if key == "tc_gemm":      # [G ; Dx], [B_[:, subset] | C_] (critical path), B_ full width (second stream)
                    ph = GEMM_ORDER[gemm_seen % 3]
                    gemm_seen += 1
                traffic[ph] = traffic.get(ph, 0) + byt
                break
if traffic:
    traffic["_source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, %s" % tag
    json.dump(traffic, open(os.path.join(dst, "ncu_traffic.json"), "w"), indent=1)
open(os.path.join(dst, tag + "_summary.md"), "w").write("\n".join(out) + "\n")
print("wrote profiles/%s_summary.md" % tag)
