import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import jaxdem_b200 as jd
which = sys.argv[1]
dtype = torch.float64 if "64" in which else torch.float32
mt = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1000.0, poisson=0.3, density=1.0)],
                                     matcher=jd.MaterialMatchmaker.create("linear"))
if "bond" in which:
    st = jd.State.create([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]], rad=[1.1] * 3, bond_id=[[1], [0, 2], [1]], dtype=dtype)
elif "clump" in which:
    st = jd.State.create([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]], rad=[1.1] * 3, clump_id=[0, 0, 1], dtype=dtype)
else:
    st = jd.State.create([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]], rad=[1.1] * 3, dtype=dtype)
print("nmat", len(mt), "W", st.bond_id.shape)
sy = jd.System.create(st.shape, dt=1e-3, collider_type="CellList", collider_kw=dict(state=st), mat_table=mt, dtype=dtype)
sy.collider.compute_force(st, sy)
torch.cuda.synchronize()
print(which, st.force.cpu().numpy())
