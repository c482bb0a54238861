import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["roofline"]["stage_ms"], d["config"]["bit_exact_vs_oracle"], d["clocks"])
