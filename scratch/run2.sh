set -x
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slab" 2>&1 | tail -n 15
timeout 300 $TR --master-port 29511 tests/slab_worker.py 200000 10 spring peer > gpurun_out/r2_s1.log 2>&1; echo rc=$? >> gpurun_out/r2_s1.log
timeout 300 $TR --master-port 29512 tests/slab_worker.py 200000 10 cundallstrack peer > gpurun_out/r2_s2.log 2>&1; echo rc=$? >> gpurun_out/r2_s2.log
timeout 300 $TR --master-port 29513 tests/slab_worker.py 200000 10 spring sendrecv > gpurun_out/r2_s3.log 2>&1; echo rc=$? >> gpurun_out/r2_s3.log
timeout 400 $TR --master-port 29514 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_c2_n2c.json 2> gpurun_out/r2_bench_c2_n2c.err; echo rc=$?
for f in gpurun_out/r2_s1.log gpurun_out/r2_s2.log gpurun_out/r2_s3.log; do tail -n 3 $f; done
tail -5 gpurun_out/r2_bench_c2_n2c.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_c2_n2c.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["step_ms_rank0"], d["per_rank_ms"]); print({k:round(v['us_per_step'],1) for k,v in d['kernels_rank0'].items()})
PY
