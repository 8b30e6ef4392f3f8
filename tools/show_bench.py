import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); print(f, "value %.0f"%d["value"], {k:round(v,2) for k,v in d["kernels_ms_per_step"].items()}, "e2e", d["e2e"] and round(d["e2e"]["value"]), "B", d["config"]["frames_per_step_per_gpu"], "retries", d["retries"])
    except Exception as ex: print(f, "ERR", ex)
