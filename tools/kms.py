import json,sys
d=json.loads(sys.stdin.read().strip().split("\n")[-1])
print(round(d["value"],1), {k: round(v,1) for k,v in d["roofline"]["kernel_ms"].items()}, {k: round(v,1) for k,v in d["host_breakdown_last_step_ms"].items()})
