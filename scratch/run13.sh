cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 15
timeout 600 python profiles/tools/sorted_fallback.py > gpurun_out/sorted_fallback.json 2> gpurun_out/sorted_fallback.err; echo rc=$?; tail -n 3 gpurun_out/sorted_fallback.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/sorted_fallback.json'))
for k,v in d.items(): print(k, v["ms_per_step"], v["used_dense"], v["kernels_us_per_step"])
PY
