"""ms / step of config 3's physics on one GPU: 2**20 spheres, cundallstrack + verlet + verletspiral
(hook-by-hook inside jdb200_system_step), L2 flushed before every step."""
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench, jaxdem_b200 as jd
wl = bench.make_workload()
n = wl["pos"].shape[0]
rng = np.random.default_rng(3)
mats = [jd.Material.create("elasticfrict", young=1.0e4, poisson=0.3, density=1.0, mu=0.5, e=0.8, mu_r=0.05)]
mt = jd.MaterialTable.from_materials(mats, matcher=jd.MaterialMatchmaker.create("harmonic"))
st = jd.State.create(wl["pos"], vel=wl["vel"], ang_vel=rng.uniform(-1, 1, (n, 3)), rad=wl["rad"], mass=wl["mass"],
                     dtype=torch.float32, device="cuda")
sy = jd.System.create(st.shape, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="verletspiral",
                      collider_type="CellList", collider_kw=dict(state=st), domain_type="periodic",
                      domain_kw=dict(box_size=wl["box"]), force_model_type="cundallstrack", mat_table=mt,
                      dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); rd = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5): jd.System.step(st, sy, n=1)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in ev:
    flush.fill_(1); rd.sum(); a.record(); jd.System.step(st, sy, n=1); b.record()
torch.cuda.synchronize()
ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
from jaxdem_b200 import _lib
_lib.kernel_timing(True)
for _ in range(3):
    flush.fill_(1); rd.sum(); jd.System.step(st, sy, n=1)
torch.cuda.synchronize()
kt = _lib.kernel_timing_collect(); _lib.kernel_timing(False)
print(json.dumps({"config": "C3 physics at 2**20 spheres on 1 GPU: cundallstrack + verlet + verletspiral, periodic",
                  "ms_per_step": ms, "particle_steps_per_s": n / (ms * 1e-3), "alg_bytes_per_particle_step": 352,
                  "hbm_frac": 352 * n / (ms * 1e-3) / 1e9 / 6551.4,
                  "kernels_us_per_step": {k: round(1e3 * v[0] / 3, 1) for k, v in kt.items()}}))
