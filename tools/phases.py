import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("%.3f ms/step " % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["phases_ms_per_step"].items()})
