import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench, jaxdem_b200 as jd
wl = bench.make_workload()
st = jd.State.create(wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"], dtype=torch.float32, device="cuda")
sy = jd.System.create(st.shape, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                      collider_type="CellList", collider_kw=dict(state=st), domain_type="periodic",
                      domain_kw=dict(box_size=wl["box"]), force_model_type="spring", dtype=torch.float32, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for _ in range(reps):
    jd.System.step(st, sy, n=n)
torch.cuda.synchronize()
