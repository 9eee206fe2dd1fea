"""Instruction / stall-sample share per kernel phase (line ranges delimited by '---- phase' comments
and function starts) from an ncu cuda,sass source CSV."""
import csv, os, re, sys
src_path, csv_path, bases = sys.argv[1], sys.argv[2], float(sys.argv[3])
lines = open(src_path).read().split("\n")
marks = [(1, "file head")]
for i, l in enumerate(lines, 1):
    m = re.search(r"// ---- (phase \d[^-]*|sub-tile walk)", l)
    if m: marks.append((i, m.group(1).strip()[:40]))
    m = re.match(r"^(__device__|template|__global__).*?(\w+)\(", l)
    if m and "forceinline" not in l and i > 60: marks.append((i, "fn " + m.group(2)))
    if "-- k-mer windows" in l: marks.append((i, "4a k-mer fetch"))
    if "-- propensities Pv" in l: marks.append((i, "4b propensities"))
    if "-- pairwise window sums" in l: marks.append((i, "4c wp tree"))
    if "-- smoothed window count" in l: marks.append((i, "4d trimmed sum"))
    if "-- expected count: fast" in l: marks.append((i, "4e expected"))
    if "-- strand combine" in l: marks.append((i, "4f combine/lut"))
    if "-- stores: one 256" in l: marks.append((i, "4g stores"))
marks.sort()
rows = list(csv.reader(open(csv_path)))
data, cur, hdr = [], "", None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; ci = {}
        for i, h in enumerate(hdr): ci.setdefault(h, i)
        continue
    if hdr is None or not r[0].isdigit(): continue
    g = lambda k: int(r[ci[k]]) if r[ci[k]].lstrip("-").isdigit() else 0
    data.append((cur, int(r[0]), g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed")))
tot_i = sum(d[3] for d in data); tot_s = sum(d[2] for d in data)
base = os.path.basename(src_path)
print("total warp-instr %.3e, thread-instr/base %.1f" % (tot_i, sum(d[4] for d in data) / bases))
for j, (ln, name) in enumerate(marks):
    end = marks[j + 1][0] - 1 if j + 1 < len(marks) else 10 ** 9
    sel = [d for d in data if d[0] == base and ln <= d[1] <= end]
    ins = sum(d[3] for d in sel); smp = sum(d[2] for d in sel); th = sum(d[4] for d in sel)
    if ins or smp:
        print("%4d-%-5s %-42s %5.1f%% ins %5.1f%% smp %7.1f thr-instr/base" % (ln, end if end < 10**9 else "", name, 100 * ins / tot_i, 100 * smp / tot_s, th / bases))
oth = {}
for d in data:
    if d[0] != base: oth[d[0]] = oth.get(d[0], 0) + d[4]
for k, v in sorted(oth.items(), key=lambda kv: -kv[1])[:6]: print("other file %-28s %7.1f thr-instr/base" % (k, v / bases))
