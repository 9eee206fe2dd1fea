#!/usr/bin/env python
"""Static SASS instruction counts of a kernel attributed to source functions (innermost inlined frame).

    python tools/sass_lines.py footprint-tools_b200/lib/fpt_warp.o 'score_warp_kernelILb1' [--lines]

Extracts the cubin (cuobjdump -xelf), disassembles with line info (nvdisasm -gi) and, for the kernel whose mangled name
contains the pattern, counts instructions per source function: a `//## File f, line n` marker is mapped to the
function of file f whose definition starts last before line n (definitions found by a regex over the source).
Loop bodies are straight-line here, so the per-function count is the per-round warp-instruction count."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def functions_of(path, cache={}):
    if path in cache:
        return cache[path]
    out = []
    try:
        with open(path) as f:
            for i, line in enumerate(f, 1):
                m = re.match(r"^(?:template.*>\s*)?(?:static\s+)?(?:FPT_HD|FPT_NOINLINE_HD|__device__|__global__|inline)[^;=]*?\b([A-Za-z_0-9]+)\s*\(", line)
                if m and not line.strip().startswith("//"):
                    out.append((i, m.group(1)))
    except OSError:
        pass
    cache[path] = out
    return out


def func_at(path, line):
    name = "?"
    for start, fn in functions_of(path):
        if start <= line:
            name = fn
        else:
            break
    return name


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    by_line = "--lines" in sys.argv
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    counts = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    inside = False
    cur = ("?", 0)
    fresh = True
    total = 0
    for ln in dis:
        if ln.startswith(".text."):
            inside = pat in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                cur = (m.group(1), int(m.group(2)))
                fresh = False
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            fresh = True
            key = (os.path.basename(cur[0]), cur[1]) if by_line else (os.path.basename(cur[0]), func_at(cur[0], cur[1]))
            counts[key] += 1
            ops[key][m.group(1).split(".")[0]] += 1
            total += 1
    print("total static instructions: %d" % total)
    for key, n in sorted(counts.items(), key=lambda kv: -kv[1])[:60]:
        top = ", ".join("%s %d" % kv for kv in ops[key].most_common(8))
        print("%6d  %-22s %-28s %s" % (n, key[0], key[1], top))


if __name__ == "__main__":
    main()
