import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
        print(f, round(d["value"],1), "steps/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["kernel_ms_per_step"].items() if v}, "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "sweeps", d["sweeps_per_step"])
    except Exception as e:
        print(f, "ERR", e)
