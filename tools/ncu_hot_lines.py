"""Top source lines of `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` output:
share of warp-stall samples and of executed warp instructions per CUDA source line."""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
data, cur, hdr = [], "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; ci = {}
        for i, h in enumerate(hdr):
            ci.setdefault(h, i)
        continue
    if hdr is None or not r[0].isdigit():
        continue
    g = lambda k: int(r[ci[k]]) if r[ci[k]].lstrip("-").isdigit() else 0
    data.append((cur, int(r[0]), r[1].strip()[:80], g("# Samples"), g("Instructions Executed"), r))
tot_s = sum(d[3] for d in data) or 1
tot_i = sum(d[4] for d in data) or 1
print("total samples", tot_s, "total warp-instr", tot_i)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
key = 4 if (len(sys.argv) > 3 and sys.argv[3] == "inst") else 3
for d in sorted(data, key=lambda d: -d[key])[:n]:
    st = sorted(((int(d[5][ci[h]]) if d[5][ci[h]].isdigit() else 0, h[6:]) for h in stall_cols), reverse=True)[:3]
    print("%-13s %4d %5.1f%%smp %5.1f%%ins  %-80s %s" % (d[0][:13], d[1], 100 * d[3] / tot_s, 100 * d[4] / tot_i, d[2],
                                                       " ".join("%s:%d" % (h, v) for v, h in st if v)))
