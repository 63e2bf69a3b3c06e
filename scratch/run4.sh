cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slab" 2>&1 | tail -n 5
timeout 400 $TR --master-port 29520 scratch/slab_prof2.py peer 2>&1 | grep -v "^\*\|OMP_NUM\|^$"
