import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f.split('/')[-1], "wall", round(d["ms_per_step"],4), "dev", round(d["device_ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms"].items()}, {k:int(v) for k,v in d["per_step"].items()}, "e2e", round(d["e2e"]["ms_per_step"],3), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "whole", round(d["roofline"]["whole_step"]["frac"],3))
