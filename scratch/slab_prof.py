import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import bench
from jaxdem_b200.slab import create_slab_system
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
wl = bench.make_workload(seed=1, stack=world)
slab = create_slab_system(dict(pos=wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"]), box_size=wl["box"], dt=1e-3,
                          rotation_integrator_type="", dtype=torch.float32, device=dev, capacity_factor=1.35)
slab.compute_force()
for _ in range(3): slab.step(1)
T = {"before": 0, "exchange": 0, "force": 0, "after": 0}
def sync():
    torch.cuda.synchronize()
    return time.perf_counter()
K = 10
for _ in range(K):
    t0 = sync(); slab.engine.before_force(slab.view(slab.n_own))
    t1 = sync(); slab.exchange()
    t2 = sync(); slab.engine.compute_force(slab.view(slab.n_own + slab.n_ghost))
    t3 = sync(); slab.engine.after_force(slab.view(slab.n_own))
    t4 = sync()
    for k, d in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)): T[k] += d
if rank == 0: print("world", world, {k: round(1e3 * v / K, 3) for k, v in T.items()}, "ms")
if rank == 0 and world > 1:
    import torch.profiler as tp
    with tp.profile(activities=[tp.ProfilerActivity.CPU, tp.ProfilerActivity.CUDA]) as prof:
        slab.exchange(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25))
elif world > 1:
    slab.exchange(); torch.cuda.synchronize()
if world > 1: dist.barrier(); dist.destroy_process_group()
