"""Print the headline solve time and the 4096^2 full-step kernel times of bench.py JSON lines.
usage: python tools/show_step.py a.json b.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        line = [f.split("/")[-1], f"solve {d['ms_per_step']:.4f} ms"]
        for k, x in (d.get("full_step") or {}).items():
            if isinstance(x, dict) and k.startswith("4096"):
                line.append(f"step {x['ms_per_step']:.4f} ms")
                line += [f"{kk} {vv['ms']:.4f} ({vv['frac']:.2f})" for kk, vv in (x.get("kernels") or {}).items()]
        print(" | ".join(line))
    except Exception as e:
        print(f, "failed:", e)
