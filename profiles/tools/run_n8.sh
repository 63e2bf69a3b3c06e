#!/bin/bash
# usage (on an 8-GPU box): bash profiles/tools/run_n8.sh -> weak-scaling C2 line and the 64 M-sphere config-3 line
cd /root/repo
bash profiles/tools/run_scale.sh 8 c2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29515 bench.py --gpus 8 --config c3 --n-total 67108864 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_c3_64m_n8.json 2> gpurun_out/r2_bench_c3_64m_n8.err; echo rc=$?
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_c3_64m_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print("c3 64M", d["value"], d["ms_per_step"], d["scaling"], d["parity"]["parity_checked"]); print({k:round(v['us_per_step'],1) for k,v in d['kernels_rank0'].items()})
PY
