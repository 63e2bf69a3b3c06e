cd /root/repo
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
#timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_c2_n${N}.json 2> gpurun_out/r2_bench_c2_n${N}.err; echo rc=$?
tail -n 2 gpurun_out/r2_bench_c2_n${N}.err
if [ "$N" = "8" ]; then
timeout 500 $TR --master-port 29515 bench.py --gpus $N --config c3 --n-total 67108864 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_c3_64m_n${N}.json 2> gpurun_out/r2_bench_c3_64m_n${N}.err; echo rc=$?
tail -n 2 gpurun_out/r2_bench_c3_64m_n${N}.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_c*_n$N.json')):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); print(f, d["value"], d["ms_per_step"], d["scaling"], d["step_ms_rank0"], [r[:4]+r[6:] for r in d["per_rank_ms"]["rows"]]); print({k:round(v['us_per_step'],1) for k,v in d['kernels_rank0'].items()})
PY
