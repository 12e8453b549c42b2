"""Pretty-print bench.py JSON lines (development helper)."""
import json
import sys

for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith("{"):
        if ln:
            print(ln[:300])
        continue
    d = json.loads(ln)
    r = d.get("roofline", {})
    print(d.get("config", {}).get("workload"), "ms", round(d["ms_per_step"], 2), "Gpts/s", round(d["value"] / 1e9, 2),
          "outGB/s", round(d.get("output_GBps", 0), 1), "frac", round(r.get("frac") or 0, 4), r.get("kernel"),
          r.get("stage_ms"), "| e2e Gpts/s", round(d["e2e"]["value"] / 1e9, 2), "ms", round(d["e2e"].get("ms_per_step", 0), 1),
          "| clocks", d.get("clocks"), "| gen_s", d.get("config", {}).get("generate_s"), "|", r.get("note"),
          "| cpu", d.get("cpu_baseline"))
