#!/bin/bash
# One-GPU evidence refresh (on the GPU box): full GPU test suite, every bench configuration, the reference arm,
# ncu --set full captures and launch lists of the C2 and C3 steps.  Outputs under gpurun_out/.
cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 4 > gpurun_out/r2_gputests.log; cat gpurun_out/r2_gputests.log
for c in c2 c3 c4 c5 c2nl; do
  timeout 600 python bench.py --config $c --steps 50 --warmup 5 > gpurun_out/r2_final_${c}_n1.json 2> gpurun_out/r2_final_${c}_n1.err; echo "$c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err; echo "ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_' -s 24 -c 8 -o gpurun_out/r2_c2_full python profiles/tools/prof_step.py c2 1 5 > gpurun_out/r2_c2_ncu.log 2>&1; echo "ncu c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_' -s 27 -c 9 -o gpurun_out/r2_c3_full python profiles/tools/prof_step.py c3 1 5 > gpurun_out/r2_c3_ncu.log 2>&1; echo "ncu c3 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 24 --csv --log-file gpurun_out/r2_launches_c2.csv python profiles/tools/prof_step.py c2 1 6 > /dev/null 2>&1; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 27 --csv --log-file gpurun_out/r2_launches_c3.csv python profiles/tools/prof_step.py c3 1 6 > /dev/null 2>&1; echo rc=$?
python - <<'PY'
import json
for c in ("c2","c3","c4","c5","c2nl"):
  for l in open('gpurun_out/r2_final_%s_n1.json' % c):
    if l.startswith('{'):
        d=json.loads(l); print(c, d["value"], d["ms_per_step"], d["step_roofline"]["frac"], d["parity"]["parity_checked"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for l in open('gpurun_out/r2_final_reference.json'):
    if l.startswith('{'):
        d=json.loads(l); print("reference", d["value"], d["cpu_baseline"]["cores"])
PY
ls -la gpurun_out/*.ncu-rep
# the 64 M-sphere single-GPU baseline of the strong-scaling claim
timeout 1200 python bench.py --config c3 --n-total 67108864 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_final_c3_64m_n1.json 2> gpurun_out/r2_final_c3_64m_n1.err; echo "64m rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_final_c3_64m_n1.json'):
    if l.startswith('{'):
        d=json.loads(l); print("c3 64M n1", d["value"], d["ms_per_step"], {k:round(v['us_per_step'],1) for k,v in d['kernels'].items()})
PY
