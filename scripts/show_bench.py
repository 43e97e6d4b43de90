"""Condensed view of a bench.py JSON line: headline, e2e, roofline, every config."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("C2 value %.0f GB/s  ms/step %.3f  frac %.3f  e2e %.1f GB/s  pageable %s  parity %s  rejected %s  launches %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], (d["e2e"].get("pageable") or {}).get("value"), d["parity_checked_vs_oracle"], d.get("rejected"), d.get("gpu_launches")))
print("cpu", d.get("cpu_baseline") and {k: d["cpu_baseline"].get(k) for k in ("value", "mean", "stdev")})
for k, v in d.get("configs", {}).items():
    if "error" in v:
        print(k, "ERROR", v["error"]); continue
    r = v.get("roofline", {})
    print("%s value %.2f %s  frac %s  kernel %s  parity %s  %s" % (k, v["value"], v["unit"].split()[0], r.get("frac") and round(r["frac"], 4), v.get("kernel"), v.get("parity_checked_vs_oracle"),
                                                                   {x: v[x] for x in ("passes", "seconds", "ms_per_step", "matches") if x in v}))
