"""Build the SAME system twice from numpy inputs: once for the CPU oracle (the checker)
and once for jaxdem_b200 (the CUDA product under test)."""

from __future__ import annotations

import numpy as np
import torch

import oracle
from oracle import colliders as ocol

TORCH_F = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


def make_inputs(n, dim, *, seed=0, box=None, phi=0.5, poly=1.0, dtype=np.float64, clumps=False,
                bonds=False, nmat=1, fixed_frac=0.0, spread=1.0):
    """Random dense packing (overlaps guaranteed).  Returns a dict of numpy arrays."""
    rng = np.random.default_rng(seed)
    rad = rng.uniform(0.5 / poly, 0.5, n)
    vol = (np.pi * rad**2).sum() if dim == 2 else (4 / 3 * np.pi * rad**3).sum()
    L = (vol / phi) ** (1.0 / dim) if box is None else box
    pos = rng.uniform(0, L, (n, dim)) * spread
    A = 1 if dim == 2 else 3
    d = dict(
        pos=pos, rad=rad, vel=rng.uniform(-1, 1, (n, dim)), ang_vel=rng.uniform(-1, 1, (n, A)),
        mass=rng.uniform(0.5, 1.5, n), box=np.full(dim, L),
    )
    if clumps:
        cid = np.arange(n)
        k = 0
        sizes = [3, 1, 4, 2, 5]
        s = 0
        while k < n:
            m = sizes[s % len(sizes)]
            cid[k : k + m] = k
            k += m
            s += 1
        d["clump_id"] = cid
        # clump members share pos_c / vel / ang_vel / q / mass / inertia (state.py:169-181)
        first = np.unique(cid, return_index=True)[1]
        rep = first[np.searchsorted(np.unique(cid), cid)]
        d["pos_p"] = rng.normal(0, 0.3, (n, dim)) * (cid != np.arange(n))[:, None]
        d["pos_p"][first] = rng.normal(0, 0.3, (len(first), dim)) * 0.5
        for key in ("pos", "vel", "ang_vel", "mass"):
            d[key] = d[key][rep]
        q = rng.normal(size=(n, 4))
        if dim == 2:
            q[:, 1:3] = 0
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        d["q"] = q[rep]
        d["inertia"] = rng.uniform(0.05, 0.2, (n, A))[rep]
    if bonds:
        d["bond_id"] = [[(i + 1) % n] if i % 3 == 0 else [] for i in range(n)]
    if nmat > 1:
        d["mat_id"] = rng.integers(0, nmat, n)
    if fixed_frac > 0:
        d["fixed"] = rng.uniform(size=n) < fixed_frac
    return {k: (v.astype(dtype) if isinstance(v, np.ndarray) and v.dtype.kind == "f" else v) for k, v in d.items()}


MATS = [
    dict(young=1.0e3, poisson=0.3, density=1.0, mu=0.5, e=0.8, mu_r=0.05),
    dict(young=3.0e3, poisson=0.25, density=2.0, mu=0.3, e=0.6, mu_r=0.1),
    dict(young=0.5e3, poisson=0.35, density=0.5, mu=0.7, e=0.0, mu_r=0.0),
]


def build_oracle(inp, *, dtype, domain="periodic", law="spring", collider="celllist", nmat=1, dt=1e-3,
                 lin="verlet", rot="verletspiral", anchor=None, gravity=None, collider_kw=None,
                 restitution=1.0, interact=False):
    st = oracle.create_state(
        inp["pos"], vel=inp.get("vel"), ang_vel=inp.get("ang_vel"), rad=inp["rad"], mass=inp.get("mass"),
        clump_id=inp.get("clump_id"), pos_p=inp.get("pos_p"), q=inp.get("q"), inertia=inp.get("inertia"),
        bond_id=inp.get("bond_id"), mat_id=inp.get("mat_id"), fixed=inp.get("fixed"), dtype=dtype)
    mt = oracle.make_material_table(MATS[:nmat], "harmonic")
    dkw = dict(box_size=inp["box"], anchor=anchor)
    if domain == "reflect":
        dkw["restitution_coefficient"] = restitution
    sy = oracle.create_system(
        st, dt=dt, linear_integrator_type=lin, rotation_integrator_type=rot, collider_type=collider,
        collider_kw=dict(collider_kw or {}), domain_type=domain, domain_kw=dkw, force_model_type=law,
        mat_table=mt, gravity=gravity, interact_same_bond_id=interact)
    return st, sy


def build_gpu(inp, *, dtype, domain="periodic", law="spring", collider="CellList", nmat=1, dt=1e-3,
              lin="verlet", rot="verletspiral", anchor=None, gravity=None, collider_kw=None,
              restitution=1.0, interact=False, grid_mode="auto", max_cells=None):
    import jaxdem_b200 as jd
    F = TORCH_F[np.dtype(dtype)]
    st = jd.State.create(
        inp["pos"], vel=inp.get("vel"), ang_vel=inp.get("ang_vel"), rad=inp["rad"], mass=inp.get("mass"),
        clump_id=inp.get("clump_id"), pos_p=inp.get("pos_p"), q=inp.get("q"), inertia=inp.get("inertia"),
        bond_id=inp.get("bond_id"), mat_id=inp.get("mat_id"), fixed=inp.get("fixed"), dtype=F, device="cuda")
    mats = [jd.Material.create("elasticfrict", **m) for m in MATS[:nmat]]
    mt = jd.MaterialTable.from_materials(mats, matcher=jd.MaterialMatchmaker.create("harmonic"))
    dkw = dict(box_size=inp["box"], anchor=anchor)
    if domain == "reflect":
        dkw["restitution_coefficient"] = restitution
    ckw = dict(collider_kw or {})
    if collider.lower() in ("celllist", "multicelllist"):
        ckw.update(state=st, grid_mode=grid_mode)
        if max_cells is not None:
            ckw["max_cells"] = max_cells
    elif collider.lower() == "neighborlist":
        ckw.update(state=st)
    sy = jd.System.create(
        st.shape, dt=dt, linear_integrator_type=lin, rotation_integrator_type=rot, collider_type=collider,
        collider_kw=ckw, domain_type=domain, domain_kw=dkw, force_model_type=law, mat_table=mt,
        force_manager_kw=dict(gravity=gravity), interact_same_bond_id=interact, dtype=F, device="cuda")
    return st, sy


def tol(dtype):
    """north_star tolerances: rel 1e-5 (f32) / 1e-12 (f64) per step."""
    return 1e-5 if np.dtype(dtype) == np.float32 else 1e-12


def assert_close(got, want, dtype, name="", scale=None, factor=1.0):
    """|got - want| <= tol * (|want| + scale), scale = the field's own magnitude (max |want|)."""
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    s = float(np.max(np.abs(want))) if scale is None else scale
    t = tol(dtype) * factor
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bound = t * (np.abs(want) + s) + 1e-300
    bad = err > bound
    assert not bad.any(), f"{name}: {bad.sum()} / {bad.size} off, max err {err.max():.3e}, scale {s:.3e}, tol {t:.1e}"


def compare_states(gst, ost, dtype, factor=1.0, fields=("pos_c", "vel", "force", "torque", "ang_vel")):
    for f in fields:
        assert_close(getattr(gst, f), getattr(ost, f), dtype, f, factor=factor)
    assert_close(gst.q.w, ost.q_w, dtype, "q_w", factor=factor)
    assert_close(gst.q.xyz, ost.q_xyz, dtype, "q_xyz", factor=factor, scale=1.0)
    assert_close(gst._pos_p_rot, ost._pos_p_rot, dtype, "_pos_p_rot", factor=factor, scale=1.0)


# ---- committed fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle) ----
def golden_cases():
    import glob
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(glob.glob(os.path.join(here, "*.npz")))


def load_golden(path):
    """-> (inputs dict as make_inputs returns it, build kwargs, stored outputs, meta)."""
    z = np.load(path)
    meta = eval(str(z["meta"]), {"__builtins__": {}}, {"True": True, "False": False})
    dtype = np.dtype(meta["dtype"]).type
    inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    kw = dict(dtype=dtype, domain=meta["domain"], law=meta["law"], lin=meta["lin"], rot=meta["rot"], dt=1e-3,
              nmat=meta["nmat"])
    out = {k: z[k] for k in z.files if not k.startswith("in_") and k != "meta"}
    return inp, kw, out, meta


def golden_close(got, want, dtype, name):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    if want.dtype.kind in "iub":
        assert np.array_equal(got, want), name
        return
    tol = 1e-5 if np.dtype(dtype) == np.float32 else 1e-11
    scale = max(1.0, float(np.abs(want).max()))
    assert float(np.abs(got.astype(np.float64) - want).max()) <= 8 * tol * scale, (name, float(np.abs(got - want).max()), scale)
