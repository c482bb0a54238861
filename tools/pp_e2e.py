import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["e2e"]["value"], d["e2e"].get("host_buffers_bit_exact_vs_oracle"), d["e2e"].get("host_numa_node"))
