import json, sys
for l in sys.stdin:
    try:
        d = json.loads(l)
    except Exception:
        continue
    print("value %.1f e2e %.1f ms/step %.2f stage_ms %s clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["stage_ms"].items()}, d.get("clocks")))
    if "cpu_baseline" in d:
        print("cpu_baseline", d["cpu_baseline"])
