cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 6
for c in c4 c5 c2nl; do
timeout 400 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_${c}_b.json 2> gpurun_out/r2_bench_${c}_b.err; echo rc=$?; tail -n 2 gpurun_out/r2_bench_${c}_b.err
done
python - <<'PY'
import json
for c in ("c4","c5","c2nl"):
  for l in open('gpurun_out/r2_bench_%s_b.json' % c):
    if l.startswith('{'):
        d=json.loads(l); print(c, d["value"], d["ms_per_step"], d["step_roofline"]["frac"], d["parity"]["parity_checked"]); print({k:round(v['us_per_step'],1) for k,v in d['kernels'].items()})
PY
