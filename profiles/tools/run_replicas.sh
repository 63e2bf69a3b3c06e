#!/bin/bash
# usage (on the GPU box): bash profiles/tools/run_replicas.sh N config -> gpurun_out/r2_bench_<config>_n<N>.json (batch / replica sharding)
cd /root/repo
N=${1:-2}; C=${2:-c4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29516 bench.py --gpus $N --config $C --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_${C}_n${N}.json 2> gpurun_out/r2_bench_${C}_n${N}.err; echo rc=$?
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/r2_bench_${C}_n${N}.err | tail -n 5
python - <<PY
import json
for l in open('gpurun_out/r2_bench_${C}_n${N}.json'):
    if l.startswith('{'):
        d=json.loads(l); print("$C", d["n_gpus"], d["value"], d["unit"], d["ms_per_step"], d["scaling"], d["config"]["parallelism"], d.get("parity"))
PY
