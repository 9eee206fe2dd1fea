#!/usr/bin/env python
"""Dynamic instruction / stall-sample share per source function from an ncu source-page CSV
(ncu -i X.ncu-rep --page source --csv --print-source cuda,sass): every source line is attributed to the function of
its file whose definition starts last before it.

    python tools/ncu_func_breakdown.py /tmp/src.csv <bases per launch> [--lines file.cuh]"""
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_lines import func_at  # noqa: E402

csv_path, bases = sys.argv[1], float(sys.argv[2])
want_lines = sys.argv[sys.argv.index("--lines") + 1] if "--lines" in sys.argv else None
rows = list(csv.reader(open(csv_path)))
cur, hdr, ci = "", None, {}
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
per_line = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {}
        for i, h in enumerate(hdr):
            ci.setdefault(h, i)
        continue
    if hdr is None or not r[0].isdigit():
        continue

    def g(k):
        v = r[ci[k]] if k in ci else "0"
        return int(v) if v.lstrip("-").isdigit() else 0

    key = (os.path.basename(cur), func_at(cur, int(r[0])))
    a = agg[key]
    a[0] += g("Instructions Executed")
    a[1] += g("Thread Instructions Executed")
    a[2] += g("# Samples")
    for k in ("stall_no_inst", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_branch_resolving",
              "stall_mio", "stall_lg", "stall_barrier", "stall_not_selected", "stall_selected", "stall_dispatch"):
        a[3][k] += g(k)
    if want_lines and os.path.basename(cur) == want_lines:
        per_line.append((int(r[0]), g("Instructions Executed"), g("# Samples"), r[1][:90]))
tot_i = sum(a[0] for a in agg.values())
tot_s = sum(a[2] for a in agg.values())
print("total warp-instr %.4e  thread-instr/base %.1f  warp-instr/base %.2f  samples %d" % (
    tot_i, sum(a[1] for a in agg.values()) / bases, tot_i / bases, tot_s))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    top = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(a[2], 1)) for k, v in a[3].most_common(4))
    print("%5.1f%% ins %5.1f%% smp %7.2f warp-instr/base  %-20s %-24s %s" % (
        100 * a[0] / tot_i, 100 * a[2] / max(tot_s, 1), a[0] / bases, key[0], key[1], top))
if per_line:
    print("--- lines of %s" % want_lines)
    for ln, ins, smp, src in sorted(per_line):
        if ins or smp:
            print("%5d %6.2f%% ins %6.2f%% smp  %s" % (ln, 100 * ins / tot_i, 100 * smp / max(tot_s, 1), src))
