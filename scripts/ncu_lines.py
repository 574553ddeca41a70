#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python scripts/ncu_lines.py gpurun_out/<tag>/prof_X.ncu-rep [top_n]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg, text = None, None, collections.Counter(), {}
inst = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name':
        print("==", r[1]); continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 6 or r[0] == '':
        continue
    try:
        key = (cur, int(r[0]))
        n = int(r[-2] or 0) if len(r) > 6 else int(r[4] or 0)
        n = int(r[4] or 0)
    except ValueError:
        # source text containing commas: the numeric columns are the last two
        try:
            key = (cur, int(r[0])); n = int(r[-2] or 0)
        except ValueError:
            continue
    agg[key] += n
    text[key] = r[1]
tot = sum(agg.values())
print("total samples", tot)
for (f, l), v in agg.most_common(topn):
    print("%5d %5.1f%%  %s:%d  %s" % (v, 100 * v / max(tot, 1), f, l, text[(f, l)].strip()[:120]))
