"""print the essentials of the last JSON line of bench.py's output: python tools/benchline.py FILE   (or on stdin: ... | benchline.py -)"""
import json, sys
if len(sys.argv) < 2:
    sys.exit(__doc__)               # never sit waiting on a terminal's stdin
text = sys.stdin.read() if sys.argv[1] == "-" else open(sys.argv[1]).read()
lines = [l for l in text.splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
out = "value %.1f GS/s  e2e %.2f GS/s  ms/step %.3f" % (d["value"] / 1e3, d["e2e"]["value"] / 1e3, d["ms_per_step"])
if "replicas" in d:
    out += "  | replicas %.1f GS/s e2e %.2f" % (d["replicas"]["value"] / 1e3, d["replicas"]["e2e"] / 1e3)
if "kernels_ms_per_step" in d:
    out += "  kernels %s" % {k: round(v, 3) for k, v in d["kernels_ms_per_step"].items() if k != "note"}
for k in ("config64", "tx"):
    if k in d:
        out += "  | %s %.1f GS/s" % (k, d[k]["value"] / 1e3)
print(out)
