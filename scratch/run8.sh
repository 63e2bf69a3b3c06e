cd /root/repo
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
timeout 300 $TR --master-port 29512 tests/slab_worker.py 200000 10 cundallstrack peer > gpurun_out/r2_s2.log 2>&1; echo rc=$?; tail -n 2 gpurun_out/r2_s2.log
timeout 300 $TR --master-port 29513 tests/slab_worker.py 200000 10 cundallstrack sendrecv > gpurun_out/r2_s3.log 2>&1; echo rc=$?; tail -n 2 gpurun_out/r2_s3.log
fi
for c in c2 c3; do
timeout 400 $TR --master-port 29514 bench.py --gpus $N --config $c --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_${c}_n${N}.json 2> gpurun_out/r2_bench_${c}_n${N}.err; echo rc=$?
tail -n 3 gpurun_out/r2_bench_${c}_n${N}.err
done
python - <<PY
import json
for c in ("c2","c3"):
  for l in open('gpurun_out/r2_bench_%s_n$N.json' % c):
    if l.startswith('{'):
        d=json.loads(l); print(c, d["value"], d["ms_per_step"], d["step_ms_rank0"], d["per_rank_ms"]["rows"]); print({k:round(v['us_per_step'],1) for k,v in d['kernels_rank0'].items()})
PY
