"""Step a bench.py configuration a few times (for ncu): python profiles/tools/prof_step.py [cfg] [n_steps_per_call] [calls]"""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench, jaxdem_b200 as jd
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
npart = (1 << 22) if cfg == "c5" else (1 << 20)
wl = bench.workload_for(cfg, npart, seed=1)
st, sy = bench.build_gpu_config(jd, torch, cfg, wl, torch.device("cuda", 0))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    flush.fill_(1)
    jd.System.step(st, sy, n=n)
torch.cuda.synchronize()
