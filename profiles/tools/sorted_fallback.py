"""The sorted strategy of the partition (radix sort + binary-searched stencil walk) measured next to the dense
one: (a) the C2 workload with grid_mode forced to "sorted"; (b) the reference's own
benchmark recipe (benchmarks/base.py:10-34: N spheres, radii U(0.08, 0.12), uniform in a 200^3 free box), whose
grid (833^3 = 578 M cells >> 4 N + 1024) cannot use the dense table.
    python profiles/tools/sorted_fallback.py > gpurun_out/sorted_fallback.json"""
import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench, jaxdem_b200 as jd
from jaxdem_b200 import _lib

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(st, sy, steps=20, warm=4):
    for _ in range(warm):
        flush.fill_(1); jd.System.step(st, sy, n=1)
    ev = []
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); jd.System.step(st, sy, n=1); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    _lib.kernel_timing(True)
    for _ in range(3):
        flush.fill_(1); jd.System.step(st, sy, n=1)
    torch.cuda.synchronize()
    kt = _lib.kernel_timing_collect(); _lib.kernel_timing(False)
    return ms, {k: round(1e3 * v[0] / 3.0, 1) for k, v in kt.items()}


out = {}
n = 1 << 20
wl = bench.make_workload(n=n, seed=1)
for mode in ("auto", "sorted"):
    st = jd.State.create(wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"], dtype=torch.float32, device=dev)
    sy = jd.System.create(st.shape, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                          collider_type="CellList", collider_kw=dict(state=st, grid_mode=mode), domain_type="periodic",
                          domain_kw=dict(box_size=wl["box"]), force_model_type="spring", dtype=torch.float32, device=dev)
    ms, k = timed(st, sy)
    out[f"c2_{mode}"] = dict(ms_per_step=ms, particle_steps_per_s=n / (ms * 1e-3), used_dense=bool(sy.collider.partition(st, sy)[3]),
                             kernels_us_per_step=k)
# the reference's benchmark input: dilute gas in a free 200^3 box (hook-by-hook flow: FreeDomain.apply every step)
rng = np.random.default_rng(0)
pos = rng.uniform(0.0, 200.0, (n, 3)).astype(np.float32)
vel = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
rad = rng.uniform(0.08, 0.12, n).astype(np.float32)
st = jd.State.create(pos, vel=vel, rad=rad, mass=np.ones(n, np.float32), dtype=torch.float32, device=dev)
sy = jd.System.create(st.shape, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                      collider_type="CellList", collider_kw=dict(state=st), domain_type="free",
                      force_model_type="spring", dtype=torch.float32, device=dev)
ms, k = timed(st, sy)
out["reference_benchmark_input_free_200cube"] = dict(ms_per_step=ms, particle_steps_per_s=n / (ms * 1e-3),
                                                     used_dense=bool(sy.collider.partition(st, sy)[3]), kernels_us_per_step=k)
print(json.dumps(out, indent=1))
