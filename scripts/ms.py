"""stdin: bench.py output; prints value and ms_per_step of the JSON line (development helper)."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        print(f"{d['value']:.1f} {d['unit']}  {d['ms_per_step']:.2f} ms/step  kernel {d['config'].get('kernel')}")
