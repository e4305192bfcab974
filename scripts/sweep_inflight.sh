#!/bin/bash
# usage: scripts/sweep_inflight.sh "1 2 4 8" [extra bench args]   (prints FS value, e2e and loglike per setting)
for s in $1; do
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline --inflight $s ${@:2} 2>&1 | tail -1 > /tmp/b.json
  python - "$s" <<'PY'
import sys, json
try:
    d = json.loads(open('/tmp/b.json').read())
    print("inflight", sys.argv[1], "value %.0f (%.2f ms) e2e %.0f loglike %.0f (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["loglike"]["value"], d["loglike"]["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline_fp64"]["ms_per_step"].items()}, flush=True)
except Exception as e:
    print("failed", e, open('/tmp/b.json').read()[-2000:])
PY
done
