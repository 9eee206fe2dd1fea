"""Print the per-kernel avg_ms of a bench.py JSON line (stdin or file args)."""
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable", e); continue
    k = d["roofline"]["kernels"]
    print("%-28s step %.3f ms  %.3e bases/s  path frac %.3f | %s" % (
        path.split("/")[-1], d["ms_per_step"], d["value"], d["roofline"]["path"]["frac"],
        "  ".join("%s %.3f" % (n, v["avg_ms"]) for n, v in k.items() if v["avg_ms"] > 0.03)))
