import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1 << 20, dtype=torch.uint8, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "rendezvous ok; ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs], flush=True)
peer = (rank + 1) % world
remote = hdl.get_buffer(peer, (1 << 20,), torch.uint8)
t.fill_(rank + 1)
hdl.barrier(channel=0)
remote[:16].fill_(100 + rank)       # P2P store into the peer's buffer
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "local buffer head after peer wrote:", t[:4].tolist(), "rest", t[16:20].tolist(), flush=True)
# timing of the barrier
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100): hdl.barrier(channel=0)
torch.cuda.synchronize(); print(rank, "barrier us", (time.perf_counter() - t0) * 1e4, flush=True)
dist.destroy_process_group()
