"""One-line summary of bench.py JSON lines: python tools/benchline.py FILE [FILE ...]  (never reads stdin)."""
import json
import sys

for path in sys.argv[1:]:
    for l in open(path):
        l = l.strip()
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        r = d.get("roofline") or {}
        v = r.get("v128256") or {}
        print(path, d.get("impl", "lac_b200"), "n_gpus", d.get("n_gpus"), "tok/s %.3fM" % (d["value"] / 1e6),
              "step %.2f ms" % d["ms_per_step"],
              "| encode %.3f ms (%.1f%%)" % (r.get("ms_per_launch", 0), 100 * r.get("frac", 0)) if r else "",
              "decode %.3f ms (%.1f%%)" % (r["decode"]["ms_per_call"], 100 * r["decode"]["frac"]) if r else "",
              "| v128k enc %.1f%% dec %.1f%% clk %s" % (100 * v["encode"]["frac"], 100 * v["decode"]["frac"],
                                                        v["clocks"]) if v else "",
              "| e2e %.0f" % d["e2e"]["value"] if d.get("e2e") else "", "| clocks", d.get("clocks"))
