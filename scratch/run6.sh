cd /root/repo
timeout 600 python -m pytest tests/test_gpu_system.py tests/test_gpu_parity.py -x -q -m gpu -k "readme or elementwise or full_step" 2>&1 | tail -n 4
timeout 300 python bench.py --config c3 --steps 30 --warmup 5 --no-cpu > gpurun_out/r2_bench_c3_b.json 2> gpurun_out/r2_bench_c3_b.err; echo rc=$?
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_c2_b.json 2> gpurun_out/r2_bench_c2_b.err; echo rc=$?
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_c3_b.json','gpurun_out/r2_bench_c2_b.json'):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["step_roofline"]["frac"]); print({k:round(v['us_per_step'],1) for k,v in d['kernels'].items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pair_rows|k_finalize|k_after|k_rotation|k_hash' -s 18 -c 6 -o gpurun_out/r2_c3 python profiles/tools/prof_step.py c3 1 6 > gpurun_out/r2_c3_ncu.log 2>&1; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 27 --csv --log-file gpurun_out/r2_launches_c2.csv python profiles/tools/prof_step.py c2 1 6 > /dev/null 2>&1; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/r2_launches_c3.csv python profiles/tools/prof_step.py c3 1 6 > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out/r2_c3.ncu-rep
