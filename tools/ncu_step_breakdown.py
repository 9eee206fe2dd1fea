"""Share of executed warp-instructions, shared-memory wavefronts and stall samples per STEP of the warp-autonomous
scoring kernel, from an ncu source-page export (every SASS instruction counted once, attributed to the source line ncu
maps it to; inlined helpers are listed under their own file):

    ncu -i capture.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_step_breakdown.py src.csv footprint-tools_b200/csrc/fpt_warp_core.cuh

The steps are the line ranges of the functions of fpt_warp_core.cuh (looked up by name in the file given, which must be
the version the capture was built from)."""
import collections, csv, re, sys

csv_path, core_path = sys.argv[1], sys.argv[2]
src = open(core_path).read().split("\n")


def line_of(pattern, after=1):
    for i in range(after - 1, len(src)):
        if re.search(pattern, src[i]):
            return i + 1
    raise SystemExit("pattern not found: " + pattern)


score = line_of(r"FPT_HD void step_score\(")
marks = [
    (1, "planning / geometry / sub_of"),
    (line_of(r"FPT_HD unsigned lo16"), "helpers (lo16 / hi16 / lds128 / agg)"),
    (line_of(r"FPT_HD void stage_issue\("), "A  stage_issue (cp.async)"),
    (line_of(r"FPT_HD unsigned stage_pack\("), "A  pack / fix mask"),
    (line_of(r"FPT_HD void step_sums\("), "B  window sums"),
    (line_of(r"constexpr int kWSmoothMax"), "exact replicas (guard band, cold)"),
    (line_of(r"FPT_HD void store_partial\("), "store_partial"),
    (score, "D  sequence window"),
    (line_of(r"trimmed window sums T", score), "D  trimmed sums"),
    (line_of(r"the 13 k-mers starting", score), "D  k-mer look-ups + estimate"),
    (line_of(r"observed counts, p-value table", score), "D  table gather / stores / z"),
    (line_of(r"FPT_HD void step_direct\("), "D2 direct evaluation"),
    (line_of(r"FPT_HD void step_windows\("), "E  window sums"),
    (line_of(r"for \(int k = 0; k < ns; \+\+k\)"), "E  rolled loop / edge rule / stores"),
    (line_of(r"FPT_HD bool process_item\("), "item glue (incl. the two loops of step C)"),
]
marks.sort()


def step(fname, line):
    if fname != core_path.split("/")[-1]:
        return {"fpt_tile.cuh": "E  normal tail (ndtr4c, fpt_tile.cuh)", "fpt_warp.cu": "kernel loop / Env (fpt_warp.cu)",
                "fpt_portable.cuh": "SIMD-in-word intrinsics (fpt_portable.cuh)"}.get(fname, "other (" + fname + ")")
    name = marks[0][1]
    for ln, nm in marks:
        if line is not None and line >= ln:
            name = nm
    return name


rows = list(csv.reader(open(csv_path)))
sass, cur, hdr, line = {}, None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in reversed(list(enumerate(r)))}; continue
    if hdr is None:
        continue
    if r[0] != "":
        line = int(r[0]); continue
    if len(r) > 2 and r[2].startswith("0x"):
        def g(k):
            v = r[hdr[k]]
            return int(v) if v.lstrip("-").isdigit() else 0
        sass[int(r[2], 16)] = (g("Instructions Executed"), g("L1 Wavefronts Shared"), g("# Samples"), cur, line)
tot = [sum(v[i] for v in sass.values()) or 1 for i in range(3)]
hot = sum(1 for v in sass.values() if v[0] * 100 >= max(x[0] for x in sass.values()))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for n, w, s, f, l in sass.values():
    a = agg[step(f, l)]
    a[0] += n; a[1] += w; a[2] += s
    a[3] += 1 if n * 100 >= max(1, tot[0] // len(sass)) * 5 else 0
print("warp-instructions %.4e, shared-memory wavefronts %.4e, SASS instructions %d" % (tot[0], tot[1], len(sass)))
print("%-46s %8s %10s %9s %9s" % ("step", "instr %", "smem wf %", "samples %", "hot SASS"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-46s %8.1f %10.1f %9.1f %9d" % (k, 100 * v[0] / tot[0], 100 * v[1] / tot[1], 100 * v[2] / tot[2], v[3]))
print("%-46s %8s %10s %9s %9d" % ("all", "", "", "", sum(v[3] for v in agg.values())))
