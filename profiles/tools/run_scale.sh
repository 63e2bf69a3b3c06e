#!/bin/bash
# usage (on the GPU box): bash profiles/tools/run_scale.sh N [config] [extra bench args] -> gpurun_out/r2_bench_<config>_n<N>.json
cd /root/repo
N=${1:-2}; C=${2:-c2}; shift; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29514 bench.py --gpus $N --config $C --steps 50 --warmup 5 --no-cpu "$@" > gpurun_out/r2_bench_${C}_n${N}.json 2> gpurun_out/r2_bench_${C}_n${N}.err; echo rc=$?
tail -n 2 gpurun_out/r2_bench_${C}_n${N}.err
python - <<PY
import json
for l in open('gpurun_out/r2_bench_${C}_n${N}.json'):
    if l.startswith('{'):
        d=json.loads(l); print("$C", d["value"], d["ms_per_step"], d["scaling"], d["parity"]["parity_checked"], d["step_ms_rank0"]); print({k:round(v['us_per_step'],1) for k,v in d['kernels_rank0'].items()})
PY
