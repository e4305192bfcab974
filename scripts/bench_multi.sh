#!/bin/bash
# bench.py on N GPUs of one box as the driver launches it: scripts/bench_multi.sh N TAG [extra bench args]
N=$1; TAG=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench_${N}gpu_line.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -c 400 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${TAG}_bench_${N}gpu_line.json") if l.startswith("{")][-1])
print("N=$N value %.4g e2e %.4g loglike %.4g ens %.3f ms" % (d["value"], d["e2e"]["value"], d["loglike"]["value"], d["loglike"]["ensemble_4096"]["ms_per_ensemble"]))
PY
