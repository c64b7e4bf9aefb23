"""prints a one-line summary of bench.py's JSON line read from stdin"""
import json
import sys

txt = sys.stdin.read().strip().splitlines()
d = json.loads(txt[-1])
print("%s value=%.4g ms/step=%.3f e2e=%.4g turns/step=%.0f roofline=%.3f clocks=%s cpu=%s" % (
    " ".join(sys.argv[1:]), d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("turns_per_step", 0),
    d["roofline"]["frac"], d.get("clocks", {}).get("sm_mhz"), (d.get("cpu_baseline") or {}).get("value")))
