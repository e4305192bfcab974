#!/bin/bash
# usage: scripts/bench_sweep.sh "1 2 3 4" [extra bench args]
for s in $1; do
  python bench.py --steps 12 --warmup 3 --no-cpu-baseline --inflight $s ${@:2} 2>&1 | tail -1 > /tmp/b.json
  python - "$s" <<'PY'
import sys, json
try:
    d = json.loads(open('/tmp/b.json').read())
    print("inflight", sys.argv[1], "value %.0f e2e %.0f loglike %.0f" % (d["value"], d["e2e"]["value"], d["loglike"]["value"]), {k: round(v, 3) for k, v in d["roofline_fp64"]["ms_per_step"].items()}, d["gpu_launches"], d["clocks"], flush=True)
except Exception as e:
    print("failed", e, open('/tmp/b.json').read()[-2000:])
PY
done
