"""Parity against outputs of the REFERENCE ITSELF (tests/golden/extras/*.npz, tests/golden/ref_*.npz).

The fixtures are produced in the build container by tests/golden/make_reference_extras.py /
make_reference_golden.py: the unmodified sources under /root/reference run through their public API on a numpy
stand-in for JAX (tests/golden/jaxshim), float64.  Here
  * the CPU tests pin the ORACLE on them (integers bit for bit, floats to a few ulp of the field scale), and
  * the ``gpu`` tests compare the CUDA path, through the C ABI, with the same reference numbers
for the rows beyond the plain cell-list step: FIRE (f1), NeighborList (f2), MultiCellList (f3),
trajectory_rollout (f4), naive collider with clumps + bonds, ForceManager buffers on clumps, reflecting box with
restitution, batched step.  Neither needs /root/reference at run time."""
import os

import numpy as np
import pytest

import oracle
from oracle import colliders as ocol, minimizers as omin
from helpers import MATS, build_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
F64 = np.float64


_DT = [F64]  # dtype of the fixture loaded last: float32 fixtures widen the f64 tolerances by 1e7 (<= 1e-4)
BOTH = pytest.mark.parametrize("sfx", ["", "_f32"], ids=["f64", "f32"])


def load(name):
    z = np.load(os.path.join(HERE, "golden", "extras", name + ".npz"))
    meta = eval(str(z["meta"]), {"__builtins__": {}}, {})
    _DT[0] = np.dtype(meta["dtype"]).type
    return {k: z[k] for k in z.files if k != "meta"}, meta


def sub(z, tag):
    """make_inputs-style dict of the arrays stored as ``<tag>_in_*``."""
    p = f"{tag}_in_" if tag else "in_"
    return {k[len(p):]: v for k, v in z.items() if k.startswith(p)}


def close(got, want, name, tol=1e-12, scale=None):
    """|got - want| <= tol * max(field scale, 1e-300); ``tol`` 1e-12 is the north-star f64 bound."""
    if hasattr(got, "detach"):
        got = got.detach().cpu().numpy()
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    s = float(np.abs(want).max()) if scale is None else scale
    err = float(np.abs(got - want).max()) if got.size else 0.0
    if _DT[0] == np.float32:
        tol = min(1e-4, tol * 1e7)
    assert err <= tol * max(s, 1e-300), f"{name}: max err {err:.3e}, scale {s:.3e}, tol {tol:.1e}"


def state_close(st, z, prefix, tol=1e-12, q=True, fscale=None):
    get = (lambda f: getattr(st, f))
    for f in ("pos_c", "vel", "force", "ang_vel"):
        close(get(f), z[prefix + f], prefix + f, tol)
    fs = float(np.abs(z[prefix + "force"]).max()) if fscale is None else fscale
    close(get("torque"), z[prefix + "torque"], prefix + "torque", tol, scale=max(fs, float(np.abs(z[prefix + "torque"]).max())))
    if q:
        if hasattr(st, "q_w"):
            qq = np.concatenate([st.q_w, st.q_xyz], axis=-1)
        else:
            import torch
            qq = torch.cat([st.q.w, st.q.xyz], dim=-1)
        close(qq, z[prefix + "q"], prefix + "q", tol, scale=1.0)


# ------------------------------------------------------------------------------------------- CPU: the oracle
def test_oracle_fire_matches_reference():
    z, meta = load("fire")
    for tag, law in (("a", "spring"), ("b", "hertz")):
        inp = sub(z, tag)
        ost, osy = build_oracle(inp, dtype=_DT[0], law=law, dt=1e-2, nmat=2 if law != "spring" else 1)
        steps, pe, _ = omin.minimize(ost, osy, omin.FireConfig(1e-2), max_steps=meta["K"], pe_tol=0.0,
                                     pe_diff_tol=0.0, force_tol=-1.0)
        assert steps == int(z[f"{tag}_steps"]) == meta["K"]
        assert abs(pe - float(z[f"{tag}_pe"])) <= 1e-12 * abs(float(z[f"{tag}_pe"]))
        for f in ("pos_c", "force", "torque"):
            close(getattr(ost, f), z[f"{tag}_{f}"], f, 1e-11, scale=float(np.abs(z[f"{tag}_force"]).max()) if f == "torque" else None)
        close(np.concatenate([ost.q_w, ost.q_xyz], axis=1), z[f"{tag}_q"], "q", 1e-11, scale=1.0)
        close(ost._pos_p_rot, z[f"{tag}_pos_p_rot"], "_pos_p_rot", 1e-11, scale=1.0)
    ost, osy = _jam_oracle(z)
    steps, pe, _ = omin.minimize(ost, osy, omin.FireConfig(1e-2), max_steps=5000)
    assert steps == int(z["c_steps"]) and 10 < steps < 5000
    assert pe <= 1e-16 and float(z["c_pe"]) <= 1e-16
    close(ost.pos_c, z["c_pos_c"], "pos_c", 1e-9)


def _jam_oracle(z):
    N = len(z["c_in_rad"])
    ost = oracle.create_state(z["c_in_pos"], rad=z["c_in_rad"], mass=np.ones(N))
    mt = oracle.make_material_table([dict(young=1.0, poisson=0.5, density=1.0)], "harmonic")
    osy = oracle.create_system(ost, dt=1e-2, collider_type="celllist", collider_kw=dict(state=ost),
                               domain_type="periodic", domain_kw=dict(box_size=z["c_in_box"]), mat_table=mt)
    return ost, osy


def test_oracle_naive_matches_reference():
    z, _ = load("naive")
    for tag, law in (("a", "cundallstrack"), ("b", "hertz")):
        inp = sub(z, tag)
        inp["bond_id"] = inp.pop("bond")
        ost, osy = build_oracle(inp, dtype=_DT[0], law=law, collider="naive", nmat=2)
        assert np.array_equal(ost.bond_id, z[f"{tag}_bond_id"])  # State.create's symmetrisation + padding
        ocol.naive_compute_force(ost, osy)
        close(ost.force, z[f"{tag}_force"], "force")
        close(ost.torque, z[f"{tag}_torque"], "torque", scale=float(np.abs(z[f"{tag}_force"]).max()))
        e = ocol.naive_compute_potential_energy(ost, osy)
        close(np.asarray(e), z[f"{tag}_energy"], "energy")


@BOTH
def test_oracle_neighborlist_matches_reference(sfx):
    z, meta = load("nlist" + sfx)
    inp = sub(z, "")
    ost, osy = build_oracle(inp, dtype=_DT[0], law="hertz", collider="neighborlist", nmat=2, dt=2e-3,
                            collider_kw=dict(cutoff=meta["cutoff"], skin=meta["skin"]))
    assert osy.collider.max_neighbors == int(z["max_neighbors"])
    ocol.compute_force(ost, osy)
    assert np.array_equal(osy.collider.neighbor_list, z["nl0"]) and osy.collider.n_build_times == int(z["builds0"])
    assert bool(osy.collider.overflow) == bool(z["overflow0"])
    close(ost.force, z["force0"], "force0")
    close(ost.torque, z["torque0"], "torque0", scale=float(np.abs(z["force0"]).max()))
    close(np.asarray(ocol.compute_potential_energy(ost, osy)), z["energy0"], "energy0")
    builds = []
    for _ in range(6):
        oracle.step(ost, osy, 5)
        builds.append(int(osy.collider.n_build_times))
    assert builds == list(z["builds"]) and builds[-1] > 2  # the thin skin forced rebuilds on the way
    assert np.array_equal(osy.collider.neighbor_list, z["nl_after"])
    state_close(ost, z, "after_", 1e-11)


@BOTH
def test_oracle_multicelllist_matches_reference(sfx):
    z, _ = load("multicell" + sfx)
    for tag, law, domain in (("a", "spring", "periodic"), ("b", "hertz", "reflect")):
        ost, osy = build_oracle(sub(z, tag), dtype=_DT[0], law=law, domain=domain, collider="multicelllist", nmat=2)
        close(np.asarray(osy.collider.cell_size), z[f"{tag}_cell_size"], "cell_size", 1e-15)
        ocol.compute_force(ost, osy)
        close(ost.force, z[f"{tag}_force"], "force")
        close(np.asarray(ocol.compute_potential_energy(ost, osy)), z[f"{tag}_energy"], "energy")
        oracle.step(ost, osy, 2)
        state_close(ost, z, f"{tag}_after_", 1e-11)


def _member_count(clump_id):
    cid = np.unique(np.asarray(clump_id), return_inverse=True)[1]
    return np.bincount(cid)[cid].astype(np.float64)


@BOTH
def test_oracle_force_manager_matches_reference(sfx):
    z, _ = load("force_manager" + sfx)
    inp = sub(z, "")
    g, fe, fc, te = (inp.pop(k) for k in ("gravity", "fe", "fc", "te"))
    ost, osy = build_oracle(inp, dtype=_DT[0], law="spring", gravity=g)
    cnt = _member_count(inp["clump_id"]).astype(_DT[0])[:, None]
    fm = osy.force_manager  # ForceManager.add_force / add_torque (force_manager.py:196-303): COM shares
    fm.external_force = fm.external_force + fe
    fm.external_force_com = fm.external_force_com + fc / cnt
    fm.external_torque = fm.external_torque + te / cnt
    oracle.step(ost, osy, 1)
    state_close(ost, z, "s1_", 1e-12)
    oracle.step(ost, osy, 1)
    state_close(ost, z, "s2_", 1e-11)


@BOTH
def test_oracle_reflect_restitution_matches_reference(sfx):
    z, meta = load("reflect" + sfx)
    for tag in ("a", "b"):
        ost, osy = build_oracle(sub(z, tag), dtype=_DT[0], law="spring", domain="reflect", dt=meta["dt"],
                                restitution=meta["restitution"])
        oracle.step(ost, osy, meta["steps"])
        state_close(ost, z, f"{tag}_", 1e-10)


def test_oracle_rollout_frames_match_reference():
    z, meta = load("rollout")
    ost, osy = build_oracle(sub(z, ""), dtype=_DT[0], law="spring")
    for k in range(meta["n"]):  # a frame is saved AFTER its `stride` steps; the initial state is not a frame
        oracle.step(ost, osy, meta["stride"])
        close(ost.pos_c, z["frames_pos_c"][k], f"frame {k} pos_c")
        close(ost.vel, z["frames_vel"][k], f"frame {k} vel")
        assert int(z["frames_step_count"][k]) == (k + 1) * meta["stride"] == osy.step_count
        assert abs(float(z["frames_time"][k]) - float(osy.time)) <= 1e-15
    close(ost.pos_c, z["final_pos_c"], "final")


@BOTH
def test_oracle_batched_step_matches_reference(sfx):
    z, meta = load("batched" + sfx)
    for b in range(meta["B"]):
        ost, osy = build_oracle(sub(z, f"b{b}"), dtype=_DT[0], law="spring")
        oracle.step(ost, osy, meta["steps"])
        for f in ("pos_c", "vel", "force", "ang_vel"):
            close(getattr(ost, f), z[f][b], f"{f}[{b}]")


def test_material_tables_match_reference():
    """MaterialTable.from_materials under the harmonic and the linear matchmaker: the oracle's tables AND the
    product's host-side tables (jaxdem_b200/materials.py, plain torch on CPU) against the reference's."""
    import jaxdem_b200 as jd
    z, _ = load("materials")
    for matcher in ("harmonic", "linear"):
        omt = oracle.make_material_table(MATS, matcher)
        pmt = jd.MaterialTable.from_materials([jd.Material.create("elasticfrict", **m) for m in MATS],
                                              matcher=jd.MaterialMatchmaker.create(matcher))
        for k in ("young", "poisson", "density", "mu", "e", "mu_r"):
            for name in (k, k + "_eff"):
                want = z[f"{matcher}_{name}"]
                close(getattr(omt, name), want, f"oracle {matcher} {name}", 1e-15)
                close(getattr(pmt, name), want, f"product {matcher} {name}", 1e-15)


@BOTH
def test_oracle_cross_neighbor_list_matches_reference(sfx):
    z, meta = load("cross" + sfx)
    for tag, domain in (("a", "periodic"), ("b", "reflect")):
        inp = sub(z, tag)
        pos_a = inp.pop("pos_a")
        ost, osy = build_oracle(inp, dtype=_DT[0], law="spring", domain=domain)
        for K in (40, 4):
            nl, ovf = ocol.celllist_create_cross_neighbor_list(pos_a, ost.pos, osy, meta["cutoff"], K)
            assert np.array_equal(nl, z[f"{tag}_nl{K}"]) and nl.dtype == z[f"{tag}_nl{K}"].dtype, (tag, K)
            assert bool(ovf) == bool(z[f"{tag}_ovf{K}"]) == (K == 4)


# ------------------------------------------------------------------------------------------- GPU: the CUDA path
gpu = pytest.mark.gpu


def _gpu(inp, **kw):
    from helpers import build_gpu
    return build_gpu(inp, dtype=_DT[0], **kw)


@gpu
def test_cuda_fire_matches_reference():
    import jaxdem_b200 as jd
    import torch
    z, meta = load("fire")
    for tag, law in (("a", "spring"), ("b", "hertz")):
        gst, gsy = _gpu(sub(z, tag), law=law, dt=1e-2, nmat=2 if law != "spring" else 1, collider="CellList")
        gst, gsy, steps, pe = jd.minimizers.minimize(gst, gsy, max_steps=meta["K"], pe_tol=0.0, pe_diff_tol=0.0,
                                                     force_tol=-1.0)
        assert int(steps) == int(z[f"{tag}_steps"])
        assert abs(float(pe) - float(z[f"{tag}_pe"])) <= 1e-10 * abs(float(z[f"{tag}_pe"]))
        fs = float(np.abs(z[f"{tag}_force"]).max())
        close(gst.pos_c, z[f"{tag}_pos_c"], "pos_c", 1e-10)
        close(gst.force, z[f"{tag}_force"], "force", 1e-9)
        close(gst.torque, z[f"{tag}_torque"], "torque", 1e-9, scale=fs)
        close(torch.cat([gst.q.w, gst.q.xyz], dim=-1), z[f"{tag}_q"], "q", 1e-10, scale=1.0)
    N = len(z["c_in_rad"])
    L = [float(v) for v in z["c_in_box"]]
    gst = jd.State.create(z["c_in_pos"], rad=z["c_in_rad"], mass=np.ones(N), dtype=torch.float64)
    gsy = jd.System.create(gst.shape, dt=1e-2, minimizer=jd.minimizers.fire, minimizer_kw=dict(dt=1e-2),
                           collider_type="CellList", collider_kw=dict(state=gst), domain_type="periodic",
                           domain_kw=dict(box_size=L), force_model_type="spring",
                           mat_table=jd.MaterialTable.from_materials(
                               [jd.Material.create("elastic", young=1.0, poisson=0.5, density=1.0)]),
                           dtype=torch.float64)
    gst, gsy, steps, pe = jd.System.minimize(gst, gsy, max_steps=5000)
    assert abs(int(steps) - int(z["c_steps"])) <= 2 and float(pe) <= 1e-16  # the reference's own stop iteration


@gpu
def test_cuda_naive_matches_reference():
    z, _ = load("naive")
    for tag, law in (("a", "cundallstrack"), ("b", "hertz")):
        inp = sub(z, tag)
        inp["bond_id"] = inp.pop("bond")
        gst, gsy = _gpu(inp, law=law, collider="naive", nmat=2)
        assert np.array_equal(gst.bond_id.cpu().numpy(), z[f"{tag}_bond_id"])
        gsy.collider.compute_force(gst, gsy)
        fs = float(np.abs(z[f"{tag}_force"]).max())
        close(gst.force, z[f"{tag}_force"], "force", 1e-11)
        close(gst.torque, z[f"{tag}_torque"], "torque", 1e-11, scale=fs)
        _, _, e = gsy.collider.compute_potential_energy(gst, gsy)
        close(e.cpu().numpy().reshape(()), z[f"{tag}_energy"], "energy", 1e-11)


@gpu
@BOTH
def test_cuda_neighborlist_matches_reference(sfx):
    import jaxdem_b200 as jd
    z, meta = load("nlist" + sfx)
    gst, gsy = _gpu(sub(z, ""), law="hertz", collider="NeighborList", nmat=2, dt=2e-3,
                    collider_kw=dict(cutoff=meta["cutoff"], skin=meta["skin"]))
    assert gsy.collider.max_neighbors == int(z["max_neighbors"])
    gsy.collider.compute_force(gst, gsy)
    assert np.array_equal(gsy.collider.neighbor_list.cpu().numpy(), z["nl0"])
    assert int(gsy.collider.n_build_times) == int(z["builds0"])
    close(gst.force, z["force0"], "force0", 1e-11)
    _, _, e = gsy.collider.compute_potential_energy(gst, gsy)
    close(e.cpu().numpy().reshape(()), z["energy0"], "energy0", 1e-11)
    builds = []
    for _ in range(6):
        jd.System.step(gst, gsy, n=5)
        builds.append(int(gsy.collider.n_build_times))
    assert builds == list(z["builds"])  # the device-side rebuild gate fires on the reference's steps
    assert np.array_equal(gsy.collider.neighbor_list.cpu().numpy(), z["nl_after"])
    state_close(gst, z, "after_", 1e-9)


@gpu
@BOTH
def test_cuda_multicelllist_matches_reference(sfx):
    import jaxdem_b200 as jd
    z, _ = load("multicell" + sfx)
    for tag, law, domain in (("a", "spring", "periodic"), ("b", "hertz", "reflect")):
        gst, gsy = _gpu(sub(z, tag), law=law, domain=domain, collider="MultiCellList", nmat=2)
        gsy.collider.compute_force(gst, gsy)
        close(gst.force, z[f"{tag}_force"], "force", 1e-11)
        _, _, e = gsy.collider.compute_potential_energy(gst, gsy)
        close(e.cpu().numpy().reshape(()), z[f"{tag}_energy"], "energy", 1e-11)
        jd.System.step(gst, gsy, n=2)
        state_close(gst, z, f"{tag}_after_", 1e-10)


@gpu
@BOTH
def test_cuda_force_manager_matches_reference(sfx):
    import jaxdem_b200 as jd
    z, _ = load("force_manager" + sfx)
    inp = sub(z, "")
    g, fe, fc, te = (inp.pop(k) for k in ("gravity", "fe", "fc", "te"))
    gst, gsy = _gpu(inp, law="spring", gravity=g)
    gsy = gsy.force_manager.add_force(gst, gsy, fe)
    gsy = gsy.force_manager.add_force(gst, gsy, fc, is_com=True)
    gsy = gsy.force_manager.add_torque(gst, gsy, te)
    jd.System.step(gst, gsy, n=1)
    state_close(gst, z, "s1_", 1e-11)
    jd.System.step(gst, gsy, n=1)
    state_close(gst, z, "s2_", 1e-10)


@gpu
@BOTH
def test_cuda_reflect_restitution_matches_reference(sfx):
    import jaxdem_b200 as jd
    z, meta = load("reflect" + sfx)
    for tag in ("a", "b"):
        gst, gsy = _gpu(sub(z, tag), law="spring", domain="reflect", dt=meta["dt"], restitution=meta["restitution"])
        jd.System.step(gst, gsy, n=meta["steps"])
        state_close(gst, z, f"{tag}_", 1e-9)


@gpu
def test_cuda_rollout_frames_match_reference():
    import jaxdem_b200 as jd
    z, meta = load("rollout")
    gst, gsy = _gpu(sub(z, ""), law="spring")
    gst, gsy, traj = jd.System.trajectory_rollout(gst, gsy, n=meta["n"], stride=meta["stride"])
    close(traj["pos_c"], z["frames_pos_c"], "frames pos_c", 1e-11)
    close(traj["vel"], z["frames_vel"], "frames vel", 1e-11)
    assert [int(v) for v in traj.step_count.reshape(-1).tolist()] == [int(v) for v in z["frames_step_count"]]
    close(traj.time.reshape(-1), z["frames_time"], "frames time", 1e-13)
    close(gst.pos_c, z["final_pos_c"], "final", 1e-11)


@gpu
@BOTH
def test_cuda_batched_step_matches_reference(sfx):
    import jaxdem_b200 as jd
    import torch
    z, meta = load("batched" + sfx)
    B = meta["B"]
    singles = [_gpu(sub(z, f"b{b}"), law="spring") for b in range(B)]
    stb = jd.State.stack([s for s, _ in singles])
    big = max(singles, key=lambda p: float(p[0].rad.max()))[0]  # one cell size that serves every system
    box = [[float(v) for v in z[f"b{b}_in_box"]] for b in range(B)]
    mt = jd.MaterialTable.from_materials([jd.Material.create("elasticfrict", **MATS[0])],
                                         matcher=jd.MaterialMatchmaker.create("harmonic"))
    syb = jd.System.create(stb.shape, dt=1e-3, collider_type="CellList", collider_kw=dict(state=big),
                           domain_type="periodic", domain_kw=dict(box_size=np.asarray(box)), force_model_type="spring",
                           mat_table=mt, dtype=torch.float32 if _DT[0] == np.float32 else torch.float64, device="cuda")
    jd.System.step(stb, syb, n=meta["steps"])
    for f in ("pos_c", "vel", "force", "ang_vel"):
        close(getattr(stb, f), z[f], f, 1e-11)


@gpu
@BOTH
def test_cuda_cross_neighbor_list_matches_reference(sfx):
    import torch
    z, meta = load("cross" + sfx)
    for tag, domain in (("a", "periodic"), ("b", "reflect")):
        inp = sub(z, tag)
        pos_a = inp.pop("pos_a")
        gst, gsy = _gpu(inp, law="spring", domain=domain)
        for K in (40, 4):
            nl, ovf = gsy.collider.create_cross_neighbor_list(torch.as_tensor(pos_a, device="cuda"), gst.pos, gsy,
                                                              meta["cutoff"], K)
            assert np.array_equal(nl.cpu().numpy(), z[f"{tag}_nl{K}"]), (tag, K)
            assert bool(ovf) == bool(z[f"{tag}_ovf{K}"])


# ------------------------------------------------------------------------------------------- full size (2**20, f32)
def _sha(a):
    import hashlib
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _fullsize_ref():
    import json
    with open(os.path.join(HERE, "golden", "ref_fullsize_c2.json")) as f:
        return json.load(f)


def test_oracle_fullsize_partition_matches_reference_digests():
    """BASELINE config 2 at FULL size (bench.py's 2**20-sphere workload, float32): the numpy oracle's and the C
    oracle's cell permutation, sorted hashes and stencil hashes hash to the digests of the reference's own
    ``_get_spatial_partition`` output (tests/golden/make_reference_fullsize.py)."""
    import bench
    from oracle import c_oracle
    ref = _fullsize_ref()
    wl = bench.make_workload()
    ost = oracle.create_state(wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"], dtype=np.float32)
    osy = oracle.create_system(ost, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                               collider_type="celllist", domain_type="periodic", domain_kw=dict(box_size=wl["box"]),
                               force_model_type="spring")
    assert ost.N == ref["n"] and float(osy.collider.cell_size) == ref["cell_size"]
    perm, sh, nh, ovf, _ = ocol.get_spatial_partition(ost.pos, osy, osy.collider.cell_size,
                                                      osy.collider.neighbor_mask, ost.idtype)
    assert perm.dtype == np.int32 and [int(v) for v in perm[:8]] == ref["perm_head"]
    assert _sha(perm) == ref["perm_sha256"] and _sha(sh) == ref["sorted_hash_sha256"]
    assert _sha(nh) == ref["nbr_hash_sha256"] and bool(ovf) == ref["hash_overflow"]
    cperm, csh, _ = c_oracle.CStep(ost, osy).partition()
    assert _sha(cperm.astype(np.int32)) == ref["perm_sha256"] and _sha(csh.astype(np.int32)) == ref["sorted_hash_sha256"]


@gpu
def test_cuda_fullsize_partition_matches_reference_digests():
    """The CUDA partition of the full-size config-2 workload against the reference's digests: permutation, sorted
    hashes and the 27 stencil hashes of every particle, bit for bit."""
    import jaxdem_b200 as jd
    import torch
    import bench
    ref = _fullsize_ref()
    wl = bench.make_workload()
    gst = jd.State.create(wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"], dtype=torch.float32)
    gsy = jd.System.create(gst.shape, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                           collider_type="CellList", collider_kw=dict(state=gst), domain_type="periodic",
                           domain_kw=dict(box_size=wl["box"]), force_model_type="spring", dtype=torch.float32)
    perm, sh, nh, _ = gsy.collider.partition(gst, gsy)
    assert gst.N == ref["n"]
    assert _sha(perm.to(torch.int32)) == ref["perm_sha256"]
    assert _sha(sh.to(torch.int32)) == ref["sorted_hash_sha256"]
    assert _sha(nh.to(torch.int32)) == ref["nbr_hash_sha256"]


@pytest.mark.parametrize("name", ["ref_c2_like_f32", "ref_c2_like_f64"])
def test_c_oracle_matches_reference(name):
    """oracle/c (the full-size checker and the timed CPU baseline) directly against the reference's outputs:
    permutation / sorted hashes bit for bit, contact forces and the 3-step velocity-Verlet trajectory."""
    from helpers import load_golden
    from oracle import c_oracle
    inp, kw, want, meta = load_golden(os.path.join(HERE, "golden", name + ".npz"))
    _DT[0] = kw["dtype"]
    ost, osy = build_oracle(inp, **kw)
    cs = c_oracle.CStep(ost, osy)
    perm, sh, _ = cs.partition()
    assert np.array_equal(perm, want["perm"]) and np.array_equal(sh, want["sorted_hash"])
    cs.compute_force()
    close(ost.force, want["force0"], "force0", 1e-12)
    cs.step(meta["steps"])
    for f in ("pos_c", "vel", "force"):
        close(getattr(ost, f), want[f + "_after"], f, 1e-11)


def test_state_create_defaults_match_reference():
    """State.create (a2): every field of the State the reference builds — defaults, material-table masses, dense
    clump relabelling, bond symmetrisation + padding, the ``_pos_p_rot`` cache — from the oracle's create_state AND
    from the product's host-side State.create (torch on the CPU; no kernel involved)."""
    import torch
    import jaxdem_b200 as jd
    z, _ = load("state_create")
    fields = ("pos_c", "pos_p", "vel", "force", "ang_vel", "torque", "rad", "volume", "mass", "inertia", "clump_id",
              "bond_id", "mat_id", "species_id", "fixed", "_pos_p_rot")
    omt = oracle.make_material_table(MATS, "harmonic")
    pmt = jd.MaterialTable.from_materials([jd.Material.create("elasticfrict", **m) for m in MATS],
                                          matcher=jd.MaterialMatchmaker.create("harmonic"))
    cases = {
        "a": (dict(), dict()), "b": (dict(), dict()),
        "c": (dict(rad=z["c_in_rad"], mat_id=z["c_in_mat_id"], mat_table=omt),
              dict(rad=z["c_in_rad"], mat_id=z["c_in_mat_id"], mat_table=pmt)),
    }
    dkw = dict(pos_p=z["d_in_pos_p"], q=z["d_in_q"], clump_id=z["d_in_clump_id"], bond_id=z["d_in_bond"])
    cases["d"] = (dkw, dkw)
    for tag, (okw, pkw) in cases.items():
        ost = oracle.create_state(z[f"{tag}_in_pos"], dtype=F64, **okw)
        pst = jd.State.create(z[f"{tag}_in_pos"], dtype=torch.float64, device="cpu", **pkw)
        for f in fields:
            want = z[f"{tag}_{f}"]
            for who, got in (("oracle", getattr(ost, f)), ("product", getattr(pst, f).numpy())):
                if want.dtype.kind in "iub":
                    assert np.array_equal(np.asarray(got), want), (who, tag, f)
                else:
                    close(got, want, f"{who} {tag} {f}", 1e-15, scale=max(1.0, float(np.abs(want).max())))
        close(np.concatenate([ost.q_w, ost.q_xyz], axis=1), z[f"{tag}_q"], "oracle q", 1e-15, scale=1.0)
        close(torch.cat([pst.q.w, pst.q.xyz], dim=-1), z[f"{tag}_q"], "product q", 1e-15, scale=1.0)


NL_KW = (dict(cutoff=1.0), dict(cutoff=1.0, skin=0.3), dict(cutoff=1.5, skin_fraction=0.2),
         dict(cutoff=1.0, skin=0.1, max_neighbors=7), dict(cutoff=1.2, skin=0.1, number_density=0.6, safety_factor=1.5))


def test_collider_create_matches_reference():
    """Collider ``Create`` (a9, f2, f3): cell size and stencil of CellList / MultiCellList from the radius spread (with
    and without a box), skin and max_neighbors sizing of NeighborList — the oracle's constructors AND the product's
    host-side ones (torch on the CPU) against the reference's."""
    import torch
    import warnings
    import jaxdem_b200 as jd
    z, _ = load("collider_create")
    for i in range(4):
        pos, rad, box = z[f"k{i}_in_pos"], z[f"k{i}_in_rad"], z[f"k{i}_in_box"]
        box = None if box.size == 0 else box
        ost = oracle.create_state(pos, rad=rad, dtype=F64)
        pst = jd.State.create(pos, rad=rad, dtype=torch.float64, device="cpu")
        oc, om = ocol.celllist_create(ost, box_size=box), ocol.multicelllist_create(ost, box_size=box)
        pc = jd.Collider.create("CellList", state=pst, box_size=box)
        pm = jd.Collider.create("MultiCellList", state=pst, box_size=box)
        for who, c, m in (("oracle", oc, om), ("product", pc, pm)):
            assert float(c.cell_size) == float(z[f"k{i}_cell_size"]), (who, i)
            assert float(m.cell_size) == float(z[f"k{i}_multi_cell_size"]), (who, i)
            assert np.array_equal(np.asarray(c.neighbor_mask), z[f"k{i}_neighbor_mask"]), (who, i)
            assert np.array_equal(np.asarray(m.neighbor_mask), z[f"k{i}_multi_neighbor_mask"]), (who, i)
        for j, kw in enumerate(NL_KW):
            want = z[f"k{i}_nl{j}"]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                on = ocol.neighborlist_create(ost, **kw)
                pn = jd.Collider.create("NeighborList", state=pst, **kw)
            for who, n in (("oracle", on), ("product", pn)):
                got = [float(n.cutoff), float(n.skin), float(int(n.max_neighbors)), float(n.secondary_collider.cell_size)]
                assert got == [float(v) for v in want], (who, i, j, got, want)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("law", ["spring", "hertz", "cundallstrack"])
def test_pair_law_contract_matches_reference(dim, law):
    """The per-pair plugin contract ``ForceModel.force / energy(i, j, pos, state, system)`` (forces/__init__.py:55-150):
    the oracle's laws AND the product's torch implementation (jaxdem_b200/pair_laws.py, CPU tensors) against the
    reference's own calls on touching, random and self pairs."""
    import torch
    import jaxdem_b200 as jd
    from oracle import forces as oforces
    z, _ = load("pair_laws")
    inp = sub(z, f"d{dim}")
    i, j = z[f"d{dim}_i"], z[f"d{dim}_j"]
    wf, wt, we = (z[f"d{dim}_{law}_{k}"] for k in ("f", "t", "e"))
    assert np.abs(wf).max() > 0 and (law != "cundallstrack" or np.abs(wt).max() > 0)
    ost = oracle.create_state(inp["pos"], vel=inp["vel"], ang_vel=inp["ang_vel"], rad=inp["rad"], mass=inp["mass"],
                              mat_id=inp["mat_id"], dtype=F64)
    osy = oracle.create_system(ost, domain_type="periodic", domain_kw=dict(box_size=inp["box"]), force_model_type=law,
                               mat_table=oracle.make_material_table(MATS[:2], "harmonic"))
    of, ot = getattr(oforces, law + "_force")(i, j, ost.pos, ost, osy)
    oe = getattr(oforces, law + "_energy")(i, j, ost.pos, ost, osy)
    st = jd.State.create(inp["pos"], vel=inp["vel"], ang_vel=inp["ang_vel"], rad=inp["rad"], mass=inp["mass"],
                         mat_id=inp["mat_id"], dtype=torch.float64, device="cpu")
    mt = jd.MaterialTable.from_materials([jd.Material.create("elasticfrict", **m) for m in MATS[:2]],
                                         matcher=jd.MaterialMatchmaker.create("harmonic"))
    sy = jd.System.create(st.shape, domain_type="periodic", domain_kw=dict(box_size=inp["box"]), force_model_type=law,
                          mat_table=mt, collider_type="naive", dtype=torch.float64, device="cpu")
    pf, pt = sy.force_model.force(torch.as_tensor(i), torch.as_tensor(j), st.pos, st, sy)
    pe = sy.force_model.energy(torch.as_tensor(i), torch.as_tensor(j), st.pos, st, sy)
    fs = float(np.abs(wf).max())
    for who, (f, t, e) in (("oracle", (of, ot, oe)), ("product", (pf, pt, pe))):
        close(f, wf, f"{who} force", 1e-13)
        close(t, wt, f"{who} torque", 1e-13, scale=fs)
        close(e, we, f"{who} energy", 1e-13)


def test_oracle_energy_helpers_match_reference():
    """utils/thermal.py kinetic / potential / total energy of a clump system under gravity (what the energy-drift
    tests of test_oracle_pins.py and tests/test_gpu_system.py integrate)."""
    from oracle import force_manager as ofm, system as osys
    z, _ = load("energy")
    for tag in ("a", "b"):
        inp = sub(z, tag)
        g = inp.pop("gravity")
        ost, osy = build_oracle(inp, dtype=F64, law="hertz", gravity=g)
        ke = osys.kinetic_energy(ost)
        close(np.asarray(ke), z[f"{tag}_ke_t"] + z[f"{tag}_ke_r"], "kinetic", 1e-13)
        pe = ofm.compute_potential_energy(ost, osy) + ocol.compute_potential_energy(ost, osy)
        close(np.asarray(pe), z[f"{tag}_pe"], "potential", 1e-12, scale=max(abs(float(z[f"{tag}_pe"])), abs(float(z[f"{tag}_e"]))))
        close(np.asarray(osys.total_energy(ost, osy)), z[f"{tag}_e"], "total", 1e-12)


def test_error_behaviour_matches_reference():
    """Invalid calls raise what the reference raises (type and message; for the factories' "Unknown X" errors the
    message up to the list of registered keys, which differs by scope): recorded from the reference in
    tests/golden/extras/errors.npz."""
    import torch
    import jaxdem_b200 as jd
    z, _ = load("errors")
    st = jd.State.create(np.random.default_rng(0).uniform(0, 3, (6, 2)), dtype=torch.float64, device="cpu")
    mt_el = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1.0, poisson=0.3, density=1.0)])
    sy = jd.System.create(st.shape, device="cpu")
    calls = dict(
        missing_material_fields=lambda: jd.System.create(st.shape, force_model_type="cundallstrack", mat_table=mt_el,
                                                         device="cpu"),
        skin_and_fraction=lambda: jd.Collider.create("neighborlist", state=st, cutoff=1.0, skin=0.1, skin_fraction=0.1),
        unknown_collider=lambda: jd.Collider.create("nosuchcollider"),
        rollout_without_n=lambda: jd.System.trajectory_rollout(st, sy),
        unknown_material=lambda: jd.Material.create("nosuchmaterial"),
        unknown_force_model=lambda: jd.ForceModel.create("nosuchlaw"),
        unknown_domain=lambda: jd.Domain.create("nosuchdomain", dim=2),
    )
    for name, f in calls.items():
        want_type, want_msg = (str(v) for v in z[name])
        assert want_type, name  # the reference did raise
        with pytest.raises(Exception) as ei:
            f()
        assert type(ei.value).__name__ == want_type, (name, type(ei.value).__name__, want_type)
        got = str(ei.value)
        if name.startswith("unknown_"):
            assert got.split("Available:")[0] == want_msg.split("Available:")[0], (name, got, want_msg)
        else:
            assert got == want_msg, (name, got, want_msg)
