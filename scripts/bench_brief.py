"""Print the headline numbers and the level-0..2 breakdown of a bench.py JSON line (argv[1])."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cycles/s %.2f  ms %.3f  profiled %.3f  launches %d" % (d["value"], d["ms_per_step"], d["ms_per_step_profiled_pass"], d["gpu_launches"]))
for r in d["breakdown_ms_per_cycle"][:3]:
    print({k: (round(v["ms"], 3) if isinstance(v, dict) else v) for k, v in r.items()})
