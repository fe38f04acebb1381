import sys,json
for l in sys.stdin.read().strip().splitlines():
    if l.startswith('{'):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"].get("step_frac_of_peak"))
