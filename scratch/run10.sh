cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "order_id or slab or partition" 2>&1 | tail -n 5
for law in spring cundallstrack; do
timeout 300 $TR --master-port 29512 tests/slab_worker.py 200000 10 $law peer > gpurun_out/r2_s_$law.log 2>&1; echo rc=$?; grep "slab\]\|SLAB" gpurun_out/r2_s_$law.log
done
