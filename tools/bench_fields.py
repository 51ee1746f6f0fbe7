import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"): continue
    d = json.loads(line)
    print({k: d.get(k) for k in ("value", "ms_per_step", "clocks")}, "e2e_ms", d["e2e"].get("ms_per_step"), "kernel_ms", d["roofline"].get("kernel_launch_ms"))
