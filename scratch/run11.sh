cd /root/repo
timeout 900 python -m pytest tests/test_gpu_minimize.py tests/test_gpu_parity.py -x -q -m gpu -k "minimize or energy or force" 2>&1 | tail -n 4
for c in c4 c2; do
timeout 400 python bench.py --config $c --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_${c}_b.json 2> gpurun_out/r2_bench_${c}_b.err; echo rc=$?; tail -n 2 gpurun_out/r2_bench_${c}_b.err
done
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_c4_b.json','gpurun_out/r2_bench_c2_b.json'):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["step_roofline"]["frac"], d.get("parity")); print({k:round(v['us_per_step'],1) for k,v in d['kernels'].items()})
PY
