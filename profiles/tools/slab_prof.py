import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import bench
from jaxdem_b200 import _lib
from jaxdem_b200.slab import create_slab_system
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1 << 20
for transport in sys.argv[1:] or ["peer", "sendrecv"]:
    wl = bench.workload_for("c2", n, seed=1 + rank)
    L = float(wl["box"][0]); wl["pos"][:, 2] += np.float32(rank * L)
    box = np.array([L, L, L * world], dtype=np.float32)
    slab = create_slab_system(dict(pos=wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"]), box_size=box, dt=1e-3,
                              rotation_integrator_type="", dtype=torch.float32, device=dev, capacity_factor=1.35,
                              local_gid=rank * n + np.arange(n), n_total=n * world, rad_range=(0.5, 0.5), transport=transport)
    slab.compute_force()
    for _ in range(5): slab.step(1)
    torch.cuda.synchronize(); dist.barrier()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    evs = []
    for _ in range(20):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); slab.step(1); b.record(); evs.append((a, b))
    torch.cuda.synchronize(); dist.barrier()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    _lib.kernel_timing(rank == 0)
    for _ in range(5):
        flush.fill_(1); slab.step(1)
    torch.cuda.synchronize()
    kt = _lib.kernel_timing_collect() if rank == 0 else {}
    _lib.kernel_timing(False)
    dist.barrier()
    if rank == 0:
        print(transport, "bound", slab.bound, "cap", slab.cap, "caps", slab.ghost_cap, slab.migrant_cap, "step ms min/med/max", ts[0], ts[len(ts)//2], ts[-1])
        print({k: round(1e3 * v[0] / v[1], 1) for k, v in kt.items()})
    del slab
dist.barrier(); dist.destroy_process_group()
