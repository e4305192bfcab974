#!/bin/bash
# GPU-box check used between kernel passes: parity tier, then a short bench line (stage times included)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bcur.json
  python -c "
import json; d=json.load(open('gpurun_out/bcur.json')); print('value %.0f  %.3f ms  e2e %.0f  loglike %.0f  %.3f ms  ens %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['loglike']['value'], d['loglike']['ms_per_step'], d['loglike']['ensemble_4096']['ms_per_ensemble']), {k: round(v, 3) for k, v in d['roofline_fp64']['ms_per_step'].items()})"
done
