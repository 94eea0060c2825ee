"""One line per bench JSON: the numbers that are compared from run to run."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "ERR", e)
        continue
    r = d["roofline"]
    print(f.split("/")[-1], "t_step", round(d["ms_per_step"], 4), "dev", round(d["device_ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3),
          "frac", round(r["frac"], 3), "frac_t_step", round(r.get("frac_on_t_step", 0), 3), {k: round(v, 4) for k, v in d["stage_ms"].items()},
          {k: int(v) for k, v in d["per_step"].items()}, "fallbacks", d["e2e"].get("event_sort_fallbacks"))
    for k in ("parity", "config1", "config2", "config4", "cpu_baseline"):
        if k in d:
            print("   ", k, json.dumps(d[k])[:700])
