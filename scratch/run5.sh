cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 8
timeout 300 python bench.py --config c3 --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_c3_b.json 2> gpurun_out/r2_bench_c3_b.err; echo rc=$?
tail -n 3 gpurun_out/r2_bench_c3_b.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_c3_b.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d.get("parity")); print({k:round(v['us_per_step'],1) for k,v in d['kernels'].items()})
PY
