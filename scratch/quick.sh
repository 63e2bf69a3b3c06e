set -e
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err || tail -5 gpurun_out/bench_q.err
