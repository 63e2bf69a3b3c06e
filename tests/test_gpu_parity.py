"""GPU parity tests: the CUDA path (through the C ABI, via the jaxdem_b200 plugin classes)
against the CPU oracle on the same seeded inputs.  Integer / index results must be
BIT-EXACT; floating point within rel 1e-5 (f32) / 1e-12 (f64) of the field scale."""

import numpy as np
import pytest
import torch

import oracle
from oracle import colliders as ocol, domains as odom, force_manager as ofm, integrators as oint
from helpers import assert_close, build_gpu, build_oracle, compare_states, make_inputs

pytestmark = pytest.mark.gpu

DT = [np.float32, np.float64]


def _partition_oracle(ost, osy):
    pos = ost.pos
    perm, sh, nh, ovf, _ = ocol.get_spatial_partition(pos, osy, osy.collider.cell_size,
                                                      osy.collider.neighbor_mask, ost.idtype)
    if osy.domain.periodic:
        nh = ocol.dedup_stencil_hashes(nh)
    return perm, sh, nh, ovf


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("domain", ["periodic", "free", "reflect"])
@pytest.mark.parametrize("grid_mode", ["auto", "sorted"])
def test_partition_bit_exact(dtype, dim, domain, grid_mode):
    n = 3000
    # spread 1.3: a third of the particles lie outside the box (un-wrapped periodic
    # coordinates / out-of-grid cells for bounded domains)
    inp = make_inputs(n, dim, seed=11, dtype=dtype, poly=1.6, spread=1.3)
    inp["pos"] -= 0.15 * inp["box"]
    ost, osy = build_oracle(inp, dtype=dtype, domain=domain)
    gst, gsy = build_gpu(inp, dtype=dtype, domain=domain, grid_mode=grid_mode)
    perm, sh, nh, ovf = _partition_oracle(ost, osy)
    gperm, gsh, gnh, dense = gsy.collider.partition(gst, gsy)
    assert np.array_equal(gsh.cpu().numpy(), sh)
    assert np.array_equal(gperm.cpu().numpy(), perm)
    assert np.array_equal(gnh.cpu().numpy(), nh)
    if grid_mode == "sorted":
        assert not bool(dense)
    elif domain == "periodic":
        assert bool(dense)  # periodic hashes always fit the dense table here


@pytest.mark.parametrize("dtype", DT)
def test_partition_dense_equals_sorted_and_tiny_grids(dtype):
    # box of 2 cells per axis: stencil rows collide after the wrap (de-dup path)
    for dim, n in ((2, 200), (3, 400)):
        inp = make_inputs(n, dim, seed=5, dtype=dtype, box=2.3)
        ost, osy = build_oracle(inp, dtype=dtype)
        perm, sh, nh, _ = _partition_oracle(ost, osy)
        assert (nh == -1).any()
        for mode in ("auto", "sorted"):
            gst, gsy = build_gpu(inp, dtype=dtype, grid_mode=mode)
            gperm, gsh, gnh, dense = gsy.collider.partition(gst, gsy)
            assert np.array_equal(gperm.cpu().numpy(), perm) and np.array_equal(gsh.cpu().numpy(), sh)
            assert np.array_equal(gnh.cpu().numpy(), nh)
        ocol.celllist_compute_force(ost, osy)
        gsy.collider.compute_force(gst, gsy)
        assert_close(gst.force, ost.force, dtype, "force")


def test_dense_falls_back_when_cell_is_overfull():
    # 500 particles in ONE cell (> JDB200_DENSE_MAX_OCC): auto must take the sorted path
    rng = np.random.default_rng(0)
    inp = make_inputs(500, 3, seed=1, dtype=np.float64, box=20.0)
    inp["pos"] = (10.0 + rng.uniform(0, 0.4, (500, 3)))
    ost, osy = build_oracle(inp, dtype=np.float64)
    gst, gsy = build_gpu(inp, dtype=np.float64)
    perm, sh, nh, _ = _partition_oracle(ost, osy)
    gperm, gsh, gnh, dense = gsy.collider.partition(gst, gsy)
    assert not bool(dense)
    assert np.array_equal(gperm.cpu().numpy(), perm) and np.array_equal(gsh.cpu().numpy(), sh)
    ocol.celllist_compute_force(ost, osy)
    gsy.collider.compute_force(gst, gsy)
    assert_close(gst.force, ost.force, np.float64, "force")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("law", ["spring", "hertz", "cundallstrack"])
@pytest.mark.parametrize("domain", ["periodic", "free"])
def test_compute_force_and_energy(dtype, dim, law, domain):
    n = 2500
    inp = make_inputs(n, dim, seed=3, dtype=dtype, poly=1.5, phi=0.6, clumps=True, bonds=True, nmat=3)
    kw = dict(dtype=dtype, domain=domain, law=law, nmat=3)
    ost, osy = build_oracle(inp, **kw)
    gst, gsy = build_gpu(inp, **kw)
    if domain == "free":
        odom.free_apply(ost, osy)
        osy.domain.inv_box_size = (1.0 / osy.domain.box_size).astype(dtype)
        gsy.domain.apply(gst, gsy)
        assert_close(gsy.domain.box_size, osy.domain.box_size, dtype, "box", factor=0)
        assert_close(gsy.domain.anchor, osy.domain.anchor, dtype, "anchor", factor=0)
    ocol.celllist_compute_force(ost, osy)
    gsy.collider.compute_force(gst, gsy)
    assert float(np.abs(ost.force).max()) > 0
    assert_close(gst.force, ost.force, dtype, "force", factor=4)
    assert_close(gst.torque, ost.torque, dtype, "torque", factor=4)
    assert bool(gsy.collider.overflow) == bool(osy.collider.overflow)
    e = ocol.celllist_compute_potential_energy(ost, osy)
    _, _, ge = gsy.collider.compute_potential_energy(gst, gsy)
    assert_close(ge, e, dtype, "energy", factor=8)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("law", ["spring", "cundallstrack"])
def test_naive_collider_matches_oracle_naive(dtype, law):
    inp = make_inputs(300, 3, seed=9, dtype=dtype, phi=0.6, clumps=True, bonds=True, nmat=2)
    kw = dict(dtype=dtype, law=law, nmat=2)
    ost, osy = build_oracle(inp, collider="naive", **kw)
    gst, gsy = build_gpu(inp, collider="naive", **kw)
    ocol.naive_compute_force(ost, osy)
    gsy.collider.compute_force(gst, gsy)
    assert_close(gst.force, ost.force, dtype, "force", factor=4)
    assert_close(gst.torque, ost.torque, dtype, "torque", factor=4)
    _, _, ge = gsy.collider.compute_potential_energy(gst, gsy)
    assert_close(ge, ocol.naive_compute_potential_energy(ost, osy), dtype, "energy", factor=8)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("domain", ["periodic", "reflect"])
@pytest.mark.parametrize("K", [64, 6])
def test_neighbor_list_bit_exact(dtype, dim, domain, K):
    inp = make_inputs(1500, dim, seed=21, dtype=dtype, phi=0.55, clumps=True, bonds=True)
    ost, osy = build_oracle(inp, dtype=dtype, domain=domain)
    gst, gsy = build_gpu(inp, dtype=dtype, domain=domain)
    cutoff = 1.35
    nl, ovf = ocol.celllist_create_neighbor_list(ost, osy, cutoff, K)
    _, _, gnl, govf = gsy.collider.create_neighbor_list(gst, gsy, cutoff, K)
    assert bool(govf) == bool(ovf)
    assert ovf == (K == 6)
    assert np.array_equal(gnl.cpu().numpy(), nl)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
def test_integrator_hooks(dtype, dim):
    inp = make_inputs(2000, dim, seed=2, dtype=dtype, clumps=True, fixed_frac=0.2)
    rng = np.random.default_rng(4)
    for lin, rot in (("verlet", "verletspiral"), ("euler", "spiral")):
        ost, osy = build_oracle(inp, dtype=dtype, lin=lin, rot=rot, dt=5e-3)
        gst, gsy = build_gpu(inp, dtype=dtype, lin=lin, rot=rot, dt=5e-3)
        A = ost.torque.shape[1]
        f = rng.normal(0, 5, ost.force.shape).astype(dtype)
        t = rng.normal(0, 2, (ost.N, A)).astype(dtype)
        ost.force, ost.torque = f.copy(), t.copy()
        gst.force.copy_(torch.as_tensor(f))
        gst.torque.copy_(torch.as_tensor(t))
        for hook in (0, 1):
            oint.LINEAR[lin][hook](ost, osy)
            oint.ROTATION[rot][hook](ost, osy)
            if hook == 0:
                gsy.linear_integrator.step_before_force(gst, gsy)
                gsy.rotation_integrator.step_before_force(gst, gsy)
            else:
                gsy.linear_integrator.step_after_force(gst, gsy)
                gsy.rotation_integrator.step_after_force(gst, gsy)
            compare_states(gst, ost, dtype, factor=0.1)  # unfused arithmetic: ~ulp agreement


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("clumps", [False, True])
def test_force_manager_apply(dtype, dim, clumps):
    inp = make_inputs(1800, dim, seed=8, dtype=dtype, clumps=clumps)
    g = [0.0, -9.81] if dim == 2 else [0.0, 0.3, -9.81]
    ost, osy = build_oracle(inp, dtype=dtype, gravity=g)
    gst, gsy = build_gpu(inp, dtype=dtype, gravity=g)
    rng = np.random.default_rng(1)
    A = ost.torque.shape[1]
    arrs = {k: rng.normal(0, 3, s).astype(dtype) for k, s in
            (("force", ost.force.shape), ("torque", (ost.N, A)), ("ef", ost.force.shape),
             ("efc", ost.force.shape), ("et", (ost.N, A)))}
    ost.force, ost.torque = arrs["force"].copy(), arrs["torque"].copy()
    osy.force_manager.external_force = arrs["ef"].copy()
    osy.force_manager.external_force_com = arrs["efc"].copy()
    osy.force_manager.external_torque = arrs["et"].copy()
    gst.force.copy_(torch.as_tensor(arrs["force"]))
    gst.torque.copy_(torch.as_tensor(arrs["torque"]))
    gsy.force_manager.external_force.copy_(torch.as_tensor(arrs["ef"]))
    gsy.force_manager.external_force_com.copy_(torch.as_tensor(arrs["efc"]))
    gsy.force_manager.external_torque.copy_(torch.as_tensor(arrs["et"]))
    ofm.apply(ost, osy)
    gsy.force_manager.apply(gst, gsy)
    assert_close(gst.force, ost.force, dtype, "force", factor=0.1)
    assert_close(gst.torque, ost.torque, dtype, "torque", factor=0.1)
    for k in ("external_force", "external_force_com", "external_torque"):
        assert float(getattr(gsy.force_manager, k).abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("clumps", [False, True])
def test_reflect_apply(dtype, dim, clumps):
    inp = make_inputs(1500, dim, seed=13, dtype=dtype, clumps=clumps, fixed_frac=0.1, spread=1.08)
    inp["pos"] -= 0.04 * inp["box"]  # a layer of particles pokes through every wall
    kw = dict(dtype=dtype, domain="reflect", restitution=0.8, dt=2e-3)
    ost, osy = build_oracle(inp, **kw)
    gst, gsy = build_gpu(inp, **kw)
    rng = np.random.default_rng(6)
    f = rng.normal(0, 20, ost.force.shape).astype(dtype)
    t = rng.normal(0, 2, ost.torque.shape).astype(dtype)
    ost.force, ost.torque = f.copy(), t.copy()
    gst.force.copy_(torch.as_tensor(f))
    gst.torque.copy_(torch.as_tensor(t))
    v0 = ost.vel.copy()
    odom.reflect_apply(ost, osy)
    assert np.abs(ost.vel - v0).max() > 0.1  # impulses did fire
    gsy.domain.apply(gst, gsy)
    compare_states(gst, ost, dtype, factor=1.0, fields=("pos_c", "vel", "ang_vel"))


CONFIGS = [
    # (dim, domain, law, lin, rot, clumps)
    (3, "periodic", "spring", "verlet", "", False),            # BASELINE config 2 (scaled down)
    (3, "periodic", "cundallstrack", "verlet", "verletspiral", False),  # config 3
    (3, "periodic", "cundallstrack", "verlet", "verletspiral", True),   # config 5
    (3, "reflect", "spring", "verlet", "verletspiral", False),  # config 1 (README), cell list
    (2, "periodic", "spring", "verlet", "verletspiral", False),  # config 4 geometry
    (2, "free", "hertz", "euler", "spiral", True),
    (3, "free", "spring", "euler", "spiral", False),
]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "-".join(map(str, c)))
def test_full_step_matches_oracle(dtype, cfg):
    dim, domain, law, lin, rot, clumps = cfg
    inp = make_inputs(1200, dim, seed=17, dtype=dtype, phi=0.55, clumps=clumps, fixed_frac=0.05, poly=1.3)
    g = None if domain == "periodic" else ([0.0, -1.0] if dim == 2 else [0.0, 0.0, -1.0])
    kw = dict(dtype=dtype, domain=domain, law=law, lin=lin, rot=rot, dt=1e-3, gravity=g)
    ost, osy = build_oracle(inp, **kw)
    gst, gsy = build_gpu(inp, **kw)
    import jaxdem_b200 as jd
    hst, hsy = build_gpu(inp, **kw)  # hook-by-hook twin of the fused driver
    for step in range(3):
        oracle.step(ost, osy, 1)
        jd.System.step(gst, gsy, n=1)
        jd.System.step(hst, hsy, n=1, fused=False)
        compare_states(gst, ost, dtype, factor=4.0 * (step + 1))
        for f in ("pos_c", "vel", "force", "torque", "ang_vel", "_pos_p_rot"):  # same kernels, same order
            assert torch.equal(getattr(gst, f), getattr(hst, f)), f
        assert torch.equal(gst.q.w, hst.q.w) and torch.equal(gst.q.xyz, hst.q.xyz)
    assert float(gsy.step_count) == 3


@pytest.mark.parametrize("dtype", DT)
def test_fused_step_with_sphere_reference_offsets(dtype):
    """Sphere system (clump_id == arange) whose spheres carry a non-zero pos_p: the rotation integrator's
    step_before_force moves them (_pos_p_rot = R(q) pos_p) before the collider hashes them: the fused driver
    (k_rotation in front of the hash kernel, rotation after-kick in the pair epilogue) == hook-by-hook bitwise,
    both on the oracle."""
    import jaxdem_b200 as jd
    inp = make_inputs(1500, 3, seed=29, dtype=dtype, phi=0.5, poly=1.3, nmat=2)
    inp["pos_p"] = (np.random.default_rng(3).normal(0, 0.15, inp["pos"].shape)).astype(dtype)
    kw = dict(dtype=dtype, domain="periodic", law="cundallstrack", lin="verlet", rot="verletspiral", dt=1e-3, nmat=2)
    ost, osy = build_oracle(inp, **kw)
    gst, gsy = build_gpu(inp, **kw)
    hst, hsy = build_gpu(inp, **kw)
    assert not gst.has_clumps and float(gst._pos_p_rot.abs().max()) > 0
    for step in range(3):
        oracle.step(ost, osy, 1)
        jd.System.step(gst, gsy, n=1)
        jd.System.step(hst, hsy, n=1, fused=False)
        compare_states(gst, ost, dtype, factor=4.0 * (step + 1))
        for f in ("pos_c", "vel", "force", "torque", "ang_vel", "_pos_p_rot"):
            assert torch.equal(getattr(gst, f), getattr(hst, f)), f
        assert torch.equal(gst.q.w, hst.q.w) and torch.equal(gst.q.xyz, hst.q.xyz)


def test_readme_config_naive_reflect():
    # BASELINE config 1: 10x10x10 grid, spacing 0.5, r 0.1, reflect box 20, naive collider,
    # defaults of System.create (dt 0.005, verlet + verletspiral, spring, young_eff 1e4)
    import jaxdem_b200 as jd
    ost = oracle.grid_state((10, 10, 10), 0.5, 0.1, seed=0, dtype=np.float32)
    osy = oracle.create_system(ost, domain_type="reflect", domain_kw=dict(box_size=[20.0] * 3))
    gst = jd.State.create(ost.pos_c, vel=ost.vel, rad=ost.rad, mass=ost.mass, dtype=torch.float32)
    gsy = jd.System.create(gst.shape, domain_type="reflect", domain_kw=dict(box_size=[20.0] * 3))
    oracle.step(ost, osy, 20)
    jd.System.step(gst, gsy, n=20)
    compare_states(gst, ost, np.float32, factor=40.0)


@pytest.mark.parametrize("dtype", DT)
def test_batched_systems_match_unbatched(dtype):
    import jaxdem_b200 as jd
    F = torch.float32 if dtype == np.float32 else torch.float64
    singles, batched = [], []
    B = 5
    for b in range(B):
        inp = make_inputs(400, 2, seed=100 + b, dtype=dtype, phi=0.5 + 0.05 * b)
        st, sy = build_gpu(inp, dtype=dtype, law="hertz")
        singles.append((st, sy, inp))
    st_b = jd.State.stack([s[0].clone() for s in singles])
    sy_b = jd.System.create(st_b.shape, dt=1e-3, collider_type="CellList", collider_kw=dict(state=singles[0][0]),
                            domain_type="periodic", domain_kw=dict(box_size=np.stack([s[2]["box"] for s in singles])),
                            force_model_type="hertz",
                            mat_table=jd.MaterialTable.from_materials(
                                [jd.Material.create("elasticfrict", young=1.0e3, poisson=0.3, density=1.0,
                                                    mu=0.5, e=0.8, mu_r=0.05)]), dtype=F)
    jd.System.step(st_b, sy_b, n=4)
    for b, (st, sy, _) in enumerate(singles):
        jd.System.step(st, sy, n=4)
        for f in ("pos_c", "vel", "force", "ang_vel"):
            assert torch.equal(getattr(st_b, f)[b], getattr(st, f)), (b, f)


def test_pins_through_the_gpu_path():
    # reference tests/test_clump_pair_friction.py:167-186 and tests/test_excluded_pairs.py:11-61
    import jaxdem_b200 as jd
    mt = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1.0, poisson=0.3, density=1.0)])
    st = jd.State.create([[9.8, 5.0], [0.2, 5.4]], rad=[0.5, 0.5], clump_id=[0, 1], dtype=torch.float64)
    for col in ("naive", "CellList"):
        sy = jd.System.create(st.shape, collider_type=col, collider_kw=dict(state=st) if col == "CellList" else {},
                              domain_type="periodic", domain_kw=dict(box_size=[10.0, 10.0]), mat_table=mt,
                              dtype=torch.float64)
        sy.collider.compute_force(st, sy)
        np.testing.assert_allclose(st.force[0].cpu().numpy(), [-0.3071067811865475] * 2, rtol=1e-12)
    mt = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1000.0, poisson=0.3, density=1.0)],
                                         matcher=jd.MaterialMatchmaker.create("linear"))
    st = jd.State.create([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]], rad=[1.1] * 3, bond_id=[[1], [0, 2], [1]],
                         dtype=torch.float64)
    for col in ("naive", "CellList"):
        sy = jd.System.create(st.shape, dt=1e-3, collider_type=col,
                              collider_kw=dict(state=st) if col == "CellList" else {}, mat_table=mt,
                              dtype=torch.float64)
        sy.collider.compute_force(st, sy)
        f = st.force.cpu().numpy()
        assert np.allclose(f[1], 0.0, atol=1e-5) and abs(f[0, 0]) > 0.1 and np.allclose(f[0], -f[2], atol=1e-5)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim,n", [(3, 70000), (2, 20000)])
@pytest.mark.parametrize("law", ["spring", "hertz", "cundallstrack"])
@pytest.mark.parametrize("domain,plain", [("periodic", True), ("periodic", False), ("reflect", True), ("reflect", False)])
def test_tile_kernel_midsize(dtype, dim, n, law, domain, plain):
    """Systems large enough for interior tiles (the TMA-staged k_pair_tile path) next to
    edge tiles (k_pair_flat work list): forces and torques against the C restatement,
    bitwise repeatable.  plain = sphere system (clump_id == arange, no bonds, one material)."""
    from oracle import c_oracle
    inp = make_inputs(n, dim, seed=11, dtype=dtype, poly=1.0 if plain else 1.4, phi=0.55,
                      clumps=not plain, bonds=not plain, nmat=1 if plain else 3)
    kw = dict(dtype=dtype, domain=domain, law=law, nmat=1 if plain else 3)
    ost, osy = build_oracle(inp, **kw)
    gst, gsy = build_gpu(inp, **kw)
    c_oracle.CStep(ost, osy).compute_force()
    gsy.collider.compute_force(gst, gsy)
    assert float(np.abs(ost.force).max()) > 0
    assert_close(gst.force, ost.force, dtype, "force", factor=4)
    assert_close(gst.torque, ost.torque, dtype, "torque", factor=4)
    f1, t1 = gst.force.clone(), gst.torque.clone()
    gsy.collider.compute_force(gst, gsy)
    assert torch.equal(f1, gst.force) and torch.equal(t1, gst.torque)
    assert not bool(gsy.collider.overflow)


def test_full_size_properties_1m():
    """BASELINE config 2 at full size (1M spheres): size-independent properties."""
    import jaxdem_b200 as jd
    n_side = 102
    N = n_side**3
    r, phi = 0.5, 0.5
    L = (N * 4 / 3 * np.pi * r**3 / phi) ** (1 / 3)
    st = jd.utils.grid_state(n_per_axis=(n_side,) * 3, spacing=L / n_side, radius=r, jitter=0.05, seed=1,
                             dtype=torch.float32)
    sy = jd.System.create(st.shape, dt=1e-3, collider_type="CellList", collider_kw=dict(state=st),
                          rotation_integrator_type="", domain_type="periodic", domain_kw=dict(box_size=[L] * 3),
                          dtype=torch.float32)
    perm, sh, _, dense = sy.collider.partition(st, sy)
    assert bool(dense)
    assert bool((sh[1:] >= sh[:-1]).all())  # sortedness
    assert torch.equal(torch.sort(perm).values, torch.arange(N, device=perm.device, dtype=perm.dtype))
    same = sh[1:] == sh[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all())  # stable: ties in index order
    # dense and sorted strategies give the same permutation
    sy2 = jd.System.create(st.shape, dt=1e-3, collider_type="CellList",
                           collider_kw=dict(state=st, grid_mode="sorted"), rotation_integrator_type="",
                           domain_type="periodic", domain_kw=dict(box_size=[L] * 3), dtype=torch.float32)
    perm2, sh2, _, dense2 = sy2.collider.partition(st, sy2)
    assert not bool(dense2) and torch.equal(perm, perm2) and torch.equal(sh, sh2)
    sy.collider.compute_force(st, sy)
    f = st.force.double()
    assert float(f.abs().max()) > 0
    assert float(f.sum(0).abs().max()) < 1e-3 * float(f.abs().sum(0).max())  # Newton's third law
    f1 = st.force.clone()
    sy.collider.compute_force(st, sy)
    assert torch.equal(f1, st.force)  # deterministic: bitwise repeatable
    p0 = (st.vel.double() * st.mass.double()[:, None]).sum(0)
    pabs = float((st.vel.double().abs() * st.mass.double()[:, None]).sum())
    jd.System.step(st, sy, n=5)
    p1 = (st.vel.double() * st.mass.double()[:, None]).sum(0)
    # momentum conserved up to float32 rounding of 1M velocity updates (the float64 CPU
    # oracle conserves it to 1e-11; the float32 oracle drifts by the same 0.58 as the kernels)
    assert float((p1 - p0).abs().max()) < 1e-5 * pabs


def test_full_size_parity_vs_c_oracle_1m():
    """BASELINE config 2 at FULL size (2**20 spheres, bench.py's workload): the CUDA path
    against the C/OpenMP restatement of the reference — cell permutation and sorted hashes
    bit for bit, forces / velocities / positions to rel 1e-5 per step (float32)."""
    import jaxdem_b200 as jd
    import bench
    from oracle import c_oracle
    wl = bench.make_workload()
    ost = oracle.create_state(wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"], dtype=np.float32)
    osy = oracle.create_system(ost, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                               collider_type="celllist", domain_type="periodic",
                               domain_kw=dict(box_size=wl["box"]), force_model_type="spring")
    gst = jd.State.create(wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"], dtype=torch.float32)
    gsy = jd.System.create(gst.shape, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                           collider_type="CellList", collider_kw=dict(state=gst), domain_type="periodic",
                           domain_kw=dict(box_size=wl["box"]), force_model_type="spring", dtype=torch.float32)
    cs = c_oracle.CStep(ost, osy)
    cperm, csh, _ = cs.partition()
    perm, sh, _, _ = gsy.collider.partition(gst, gsy)
    assert np.array_equal(perm.cpu().numpy(), cperm) and np.array_equal(sh.cpu().numpy(), csh)
    cs.compute_force()
    gsy.collider.compute_force(gst, gsy)
    assert float(gst.force.abs().max()) > 100.0
    assert_close(gst.force, ost.force, np.float32, "force")
    # per-step parity on the SAME inputs (north_star): after every step the CUDA state is
    # re-synchronised to the oracle's, because at 1M stiff contacts in float32 a 1-ulp
    # difference in one position (7.6e-6 at x ~ 100) changes that particle's contact forces
    # by k * ulp ~ 0.08 and the two trajectories then separate chaotically
    for _ in range(3):
        for f in ("pos_c", "vel", "force"):
            getattr(gst, f).copy_(torch.as_tensor(getattr(ost, f)))
        cs.step(1)
        jd.System.step(gst, gsy, n=1)
        for f in ("pos_c", "vel", "force"):
            assert_close(getattr(gst, f), getattr(ost, f), np.float32, f)
    assert not bool(gsy.collider.overflow)


def _run_slab_worker(world, *args, graph=False):
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SLAB_GRAPH="1" if graph else "0")
    worker = os.path.join(root, "tests", "slab_worker.py")
    if world == 1:
        cmd = [sys.executable, worker, *map(str, args)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", worker, *map(str, args)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "SLAB-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert not graph or "replayed from a CUDA graph" in r.stdout


@pytest.mark.parametrize("law", ["spring", "cundallstrack"])
def test_slab_system_single_rank(law):
    """The slab driver with world_size 1 (CUDA engine, hook by hook) against System.step."""
    _run_slab_worker(1, 60000, 8, law)


@pytest.mark.parametrize("law,transport", [("spring", "peer"), ("cundallstrack", "peer"), ("spring", "sendrecv")])
def test_slab_system_two_gpus(law, transport):
    """Two slabs against the single-GPU trajectory (skipped on 1-GPU boxes): messages stored
    straight into the neighbour's memory over NVLink ("peer"), or NCCL send/recv."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_slab_worker(2, 200000, 10, law, transport)


@pytest.mark.parametrize("law", ["spring", "cundallstrack"])
def test_slab_step_cuda_graph_two_gpus(law):
    """The whole decomposed step (hooks + exchange + partition + pair kernel + epilogue) replayed from a CUDA graph
    on two GPUs: same bitwise agreement with the single-GPU trajectory as the stream-launched step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_slab_worker(2, 200000, 10, law, "peer", graph=True)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
def test_slab_pack_unpack_kernels_bit_exact(dtype, dim):
    """csrc/slab.cu (classify / compact / pack / unpack) against the numpy-torch restatement the
    gloo tests use, byte for byte, on one GPU: the rank is given a sub-range of the cell layers
    so that particles fall into every category (stay, halo, leave down / up, stray)."""
    import copy
    from jaxdem_b200.slab import create_slab_system, message_views
    from test_slab_gloo import OracleEngine
    n = 20000
    inp = make_inputs(n, dim, seed=17, dtype=dtype, phi=0.5)
    F = torch.float32 if dtype == np.float32 else torch.float64
    slab = create_slab_system(dict(pos=inp["pos"], vel=inp["vel"], ang_vel=inp["ang_vel"], rad=inp["rad"],
                                   mass=inp["mass"]), box_size=inp["box"], dtype=F, device="cuda")
    G = slab.layout.n_layers
    lo, up = G // 3, (2 * G) // 3
    slab.layout.bounds = [lo, up]
    slab.set_capacities(4096, 4096)
    # CPU twin with identical rows
    twin = copy.copy(slab)
    twin.device = torch.device("cpu")
    twin.buf = {k: v.cpu().clone() for k, v in slab.buf.items()}
    twin.header_local = torch.zeros(8, dtype=torch.int64)
    twin.set_capacities(4096, 4096)
    ref = OracleEngine(twin, box=inp["box"], law="spring", lin="verlet", rot="", dt=1e-3, dtype=dtype)
    for arrivals_fewer in (False, True):
        slab.engine.pack(slab)
        ref.pack(twin)
        torch.cuda.synchronize()
        hl, hr = slab.header_local.cpu(), twin.header_local
        assert torch.equal(hl[:4], hr[:4]) and int(hl[1]) > 0 and int(hl[2]) > 0 and int(hl[3]) > 0, (hl, hr)
        k_lo, k_up = int(hl[1]), int(hl[2])
        cm = slab.migrant_cap
        assert torch.equal(slab.holes[:k_lo].cpu(), twin.holes[:k_lo])
        assert torch.equal(slab.holes[cm:cm + k_up].cpu(), twin.holes[cm:cm + k_up])
        for a, b in ((slab.send_lo, twin.send_lo), (slab.send_up, twin.send_up)):
            va, vb = message_views(a.cpu(), slab.msg_layout, F), message_views(b, slab.msg_layout, F)
            assert torch.equal(va["header"][:3], vb["header"][:3])
            nm, ng = int(va["header"][0]), int(va["header"][1])
            assert ng > 0
            assert torch.equal(va["mig_f"][:nm], vb["mig_f"][:nm]) and torch.equal(va["mig_i"][:nm], vb["mig_i"][:nm])
            assert torch.equal(va["gh_f"][:ng], vb["gh_f"][:ng]) and torch.equal(va["gh_i"][:ng], vb["gh_i"][:ng])
        # self-exchange: what went down comes back from above and vice versa; second round: pretend
        # that fewer particles arrive than left, so that rows from the tail must fill holes
        for s in (slab, twin):
            s.recv_up.copy_(s.send_lo)
            s.recv_lo.copy_(s.send_up)
        h = [message_views(twin.recv_lo, twin.msg_layout, F)["header"],
             message_views(twin.recv_up, twin.msg_layout, F)["header"]]
        a_lo, a_up = int(h[0][0]), int(h[1][0])
        if arrivals_fewer:
            a_lo, a_up = a_lo // 3, a_up // 2
        counts = (slab.n_own, a_lo, a_up, k_lo, k_up, int(h[0][1]), int(h[1][1]))
        slab.engine.unpack(slab, counts)
        ref.unpack(twin, counts)
        torch.cuda.synchronize()
        n_new = slab.n_own - k_lo - k_up + a_lo + a_up
        tot = n_new + k_lo + k_up + counts[5] + counts[6]
        assert tot <= slab.cap
        for k in slab.buf:
            assert torch.equal(slab.buf[k][:tot].cpu(), twin.buf[k][:tot]), (k, arrivals_fewer)
        slab.n_own = twin.n_own = n_new


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim,domain", [(3, "periodic"), (3, "free"), (2, "reflect")])
@pytest.mark.parametrize("law", ["spring", "cundallstrack"])
def test_hashed_cell_table_dilute_system(dtype, dim, domain, law):
    """A grid with far more cells than the cell table has rows (dilute gas, the reference's own benchmark
    recipe benchmarks/base.py:10-34 scaled down): the table is addressed by a hash of the cell key instead of
    falling back to the radix sort.  Forces, torques, energy and neighbour lists against the oracle; forces also
    against the sorted strategy of the same build; particles outside a non-periodic grid included."""
    inp = make_inputs(6000, dim, seed=41, dtype=dtype, phi=0.03 if dim == 3 else 0.05, poly=1.5)
    if domain != "periodic":
        inp["pos"][::97] -= 0.6 * inp["box"]  # a few particles outside the box / grid
    ost, osy = build_oracle(inp, dtype=dtype, domain=domain, law=law)
    gst, gsy = build_gpu(inp, dtype=dtype, domain=domain, law=law)
    assert int(gsy.collider.table_strategy(gst, gsy)) == 2
    ocol.celllist_compute_force(ost, osy)
    gsy.collider.compute_force(gst, gsy)
    assert float(np.abs(ost.force).max()) > 0
    assert_close(gst.force, ost.force, dtype, "force")
    assert_close(gst.torque, ost.torque, dtype, "torque")
    assert not bool(gsy.collider.overflow)
    _, _, e = gsy.collider.compute_potential_energy(gst, gsy)
    want_e = ocol.celllist_compute_potential_energy(ost, osy)
    assert abs(float(e) - float(want_e)) <= (1e-5 if dtype == np.float32 else 1e-12) * max(1.0, abs(float(want_e)))
    # neighbour list: rows, padding and overflow bit for bit
    K = 12
    want, wovf = ocol.celllist_create_neighbor_list(ost, osy, 1.5, K)
    _, _, got, govf = gsy.collider.create_neighbor_list(gst, gsy, 1.5, K)
    assert np.array_equal(got.cpu().numpy(), want) and bool(govf) == bool(wovf)
    # the sorted strategy (radix sort + binary search) on the same inputs: same contacts, same per-contact arithmetic
    sst, ssy = build_gpu(inp, dtype=dtype, domain=domain, law=law, grid_mode="sorted")
    assert int(ssy.collider.table_strategy(sst, ssy)) == 0
    ssy.collider.compute_force(sst, ssy)
    assert_close(gst.force, sst.force.cpu().numpy(), dtype, "force vs sorted", factor=4)
    # a full step through the driver stays on the oracle's trajectory
    import jaxdem_b200 as jd
    oracle.step(ost, osy, 2)
    jd.System.step(gst, gsy, n=2)
    compare_states(gst, ost, dtype, factor=8.0)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("law,rot", [("spring", ""), ("cundallstrack", "verletspiral")])
def test_ragged_rows_device_count(dtype, law, rot):
    """jdb200_state.n_rows: the live row count read on the device, State.N only the launch bound.  Forces and a
    few steps on the first m rows of a larger State equal, bit for bit, the same calls on a State of exactly those
    m rows; rows [m, N) are never touched."""
    import jaxdem_b200 as jd
    n, m = 6000, 4321
    inp = make_inputs(n, 3, seed=37, dtype=dtype, phi=0.5, poly=1.3, nmat=2 if law != "spring" else 1)
    kw = dict(dtype=dtype, domain="periodic", law=law, lin="verlet", rot=rot, dt=1e-3, nmat=2 if law != "spring" else 1)
    big, bsy = build_gpu(inp, **kw)
    cut = {k: (v[:m] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in inp.items()}
    small, ssy = build_gpu(cut, **kw)
    # same grid for both: the cell size of the full set
    ssy.collider.cell_size.copy_(bsy.collider.cell_size)
    big.n_rows = torch.tensor([m], dtype=torch.int64, device="cuda")
    before = {f: getattr(big, f).clone() for f in ("pos_c", "vel", "force", "torque", "ang_vel")}
    bsy.collider.compute_force(big, bsy)
    ssy.collider.compute_force(small, ssy)
    assert torch.equal(big.force[:m], small.force) and torch.equal(big.torque[:m], small.torque)
    assert float(small.force.abs().max()) > 0
    for fused in (True, False):
        jd.System.step(big, bsy, n=2, fused=fused)
        jd.System.step(small, ssy, n=2, fused=fused)
        for f in ("pos_c", "vel", "force", "torque", "ang_vel"):
            assert torch.equal(getattr(big, f)[:m], getattr(small, f)), (f, fused)
    for f, t in before.items():  # the rows past the live count were left alone
        assert torch.equal(getattr(big, f)[m:], t[m:]), f


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
def test_partition_order_id_bit_exact(dtype, dim):
    """jdb200_state.order_id: particles that share a cell are ordered by the given id instead of the row index
    — the stable sort of (hash, iota) of a system whose particle index is that id (colliders/_partition.py:91-93).
    perm and sorted hashes bit for bit against the oracle's lexsort; forces equal the ones of the same particles
    laid out in id order (same cells, same in-cell order => same contact-sum order => bitwise)."""
    inp = make_inputs(6000, dim, seed=31, dtype=dtype, phi=0.7, poly=1.6)  # dense: many multi-occupancy cells
    ost, osy = build_oracle(inp, dtype=dtype)
    gst, gsy = build_gpu(inp, dtype=dtype)
    n = gst.N
    ids = np.random.default_rng(5).permutation(n).astype(np.int64) * 3 + 7
    gst.order_id = torch.as_tensor(ids, device="cuda")
    ost.order_id = ids
    perm, sh, _, dense = gsy.collider.partition(gst, gsy)
    assert bool(dense)
    want_perm, want_sh, _, _, _ = ocol.get_spatial_partition(ost.pos, osy, osy.collider.cell_size,
                                                              osy.collider.neighbor_mask, ost.idtype, order_id=ids)
    assert np.array_equal(perm.cpu().numpy(), want_perm) and np.array_equal(sh.cpu().numpy(), want_sh)
    counts = np.bincount(want_sh - want_sh.min())
    assert counts.max() >= 3  # the in-cell order really matters here
    gsy.collider.compute_force(gst, gsy)
    # the same particles with rows in id order and no order_id
    order = np.argsort(ids)
    inp2 = {k: (v[order] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in inp.items()}
    gst2, gsy2 = build_gpu(inp2, dtype=dtype)
    gsy2.collider.compute_force(gst2, gsy2)
    assert torch.equal(gst.force[torch.as_tensor(order, device="cuda")], gst2.force)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
def test_slab_device_protocol_loopback(dtype, dim):
    """jdb200_slab_pack_dev / _unpack_dev (row counts, parities, "message complete" flags and status
    bits on the device, no host in the loop) on ONE GPU that is its own lower and upper neighbour:
    rows, counts and exchange number after two exchanges (both parities) equal the host-protocol
    restatement the gloo tests use, byte for byte; a message that never arrives sets the timeout bit
    and leaves the rows alone."""
    import copy
    import ctypes as C
    from jaxdem_b200.slab import create_slab_system, message_views
    from test_slab_gloo import OracleEngine
    n = 20000
    inp = make_inputs(n, dim, seed=19, dtype=dtype, phi=0.5)
    F = torch.float32 if dtype == np.float32 else torch.float64
    slab = create_slab_system(dict(pos=inp["pos"], vel=inp["vel"], ang_vel=inp["ang_vel"], rad=inp["rad"],
                                   mass=inp["mass"]), box_size=inp["box"], dtype=F, device="cuda")
    G = slab.layout.n_layers
    slab.layout.bounds = [G // 3, (2 * G) // 3]
    slab.set_capacities(4096, 4096)
    twin = copy.copy(slab)
    twin.device = torch.device("cpu")
    twin.buf = {k: v.cpu().clone() for k, v in slab.buf.items()}
    twin.header_local = torch.zeros(8, dtype=torch.int64)
    twin.set_capacities(4096, 4096)
    ref = OracleEngine(twin, box=inp["box"], law="spring", lin="verlet", rot="", dt=1e-3, dtype=dtype)
    # loopback ports: what goes down arrives "from above" and vice versa, [parity] buffers
    nb = slab.msg_layout["bytes"]
    pool = torch.zeros(4 * nb, dtype=torch.uint8, device="cuda")
    base = pool.data_ptr()
    arr = lambda f: (C.c_void_p * 2)(*[C.c_void_p(f(q)) for q in (0, 1)])
    slab.ports = dict(send_lo=arr(lambda q: base + (2 * q + 1) * nb), send_up=arr(lambda q: base + (2 * q) * nb),
                      recv_lo=arr(lambda q: base + (2 * q) * nb), recv_up=arr(lambda q: base + (2 * q + 1) * nb))
    slab.dev_state = torch.zeros(16, dtype=torch.int64, device="cuda")
    slab.bound = slab.cap
    slab._push_counts()
    for it in range(2):
        slab.engine.pack_dev(slab)
        slab.engine.unpack_dev(slab)
        ref.pack(twin)
        twin.recv_up.copy_(twin.send_lo)
        twin.recv_lo.copy_(twin.send_up)
        hl = twin.header_local
        h = [message_views(twin.recv_lo, twin.msg_layout, F)["header"], message_views(twin.recv_up, twin.msg_layout, F)["header"]]
        k_lo, k_up, a_lo, a_up, g_lo, g_up = int(hl[1]), int(hl[2]), int(h[0][0]), int(h[1][0]), int(h[0][1]), int(h[1][1])
        assert min(k_lo, k_up, g_lo, g_up) > 0
        ref.unpack(twin, (twin.n_own, a_lo, a_up, k_lo, k_up, g_lo, g_up))
        n_old = twin.n_own
        twin.n_own = n_old - k_lo - k_up + a_lo + a_up
        tot = twin.n_own + k_lo + k_up + g_lo + g_up
        torch.cuda.synchronize()
        ds = slab.dev_state.cpu().tolist()
        assert ds[0] == twin.n_own and ds[1] == tot and ds[2] == it + 1, (ds, twin.n_own, tot)
        assert ds[3] == 1  # strays exist in this set-up (reported, kept as owned rows), nothing else
        assert ds[4:11] == [n_old, a_lo, a_up, k_lo, k_up, g_lo, g_up]
        assert ds[11] == max(k_lo, k_up, a_lo, a_up) or it == 1
        for k in slab.buf:
            assert torch.equal(slab.buf[k][:tot].cpu(), twin.buf[k][:tot]), (k, it)
        # a second pair of exchanges moves the rows again: drift them a little
        for s_ in (slab, twin):
            s_.buf["pos_c"][:twin.n_own, -1] += 0.37
    # the neighbour never answers: unpack without a pack -> timeout bit, rows and counts untouched
    before = {k: v.clone() for k, v in slab.buf.items()}
    ds0 = slab.dev_state.cpu().tolist()
    slab.timeout_s = 0.02
    slab.engine.unpack_dev(slab)
    torch.cuda.synchronize()
    ds = slab.dev_state.cpu().tolist()
    assert ds[3] & 8 and ds[0] == ds0[0] and ds[1] == ds0[0]
    for k in ("pos_c", "vel", "gid"):
        assert torch.equal(slab.buf[k][:ds0[0]], before[k][:ds0[0]])
    with pytest.raises(RuntimeError, match="did not arrive"):
        slab.sync_counts()


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("domain", ["periodic", "reflect"])
@pytest.mark.parametrize("K", [48, 5])
def test_cross_neighbor_list_bit_exact(dtype, dim, domain, K):
    """create_cross_neighbor_list: rows, padding and overflow flag bit for bit against the oracle."""
    inp = make_inputs(3000, dim, seed=6, dtype=dtype, phi=0.55, poly=1.3)
    ost, osy = build_oracle(inp, dtype=dtype, domain=domain)
    gst, gsy = build_gpu(inp, dtype=dtype, domain=domain)
    rng = np.random.default_rng(12)
    pos_a = (rng.uniform(0.0, 0.999, (777, dim)) * inp["box"]).astype(dtype)
    if domain == "periodic":
        pos_a[::5] += inp["box"].astype(dtype)  # un-wrapped queries
    cutoff = 1.2
    want, wovf = ocol.celllist_create_cross_neighbor_list(pos_a, ost.pos, osy, cutoff, K)
    got, govf = gsy.collider.create_cross_neighbor_list(torch.as_tensor(pos_a, device="cuda"), gst.pos, gsy, cutoff, K)
    assert np.array_equal(got.cpu().numpy(), want)
    assert bool(govf) == bool(wovf)
    assert (K == 5) == bool(wovf)


@pytest.mark.parametrize("dtype", DT)
def test_dense_table_key_windows(dtype):
    """jdb200_params.key_window_*: zeroing / scanning only windows of the dense cell table gives
    the same partition and forces as the whole table (one window, and two windows scanned as one
    array); a hash outside the windows falls back to the sorted strategy (same results)."""
    inp = make_inputs(30000, 3, seed=23, dtype=dtype, phi=0.5)
    gst, gsy = build_gpu(inp, dtype=dtype)
    perm0, sh0, _, dense0 = gsy.collider.partition(gst, gsy)
    gsy.collider.compute_force(gst, gsy)
    f0 = gst.force.clone()
    assert bool(dense0)
    big = int(gsy.collider.max_cells) + 1  # the kernels clamp a window to the rows of the table in use
    for windows, expect_dense in ((((0, big), (0, 0)), True),
                                  (((0, 8192), (8192, big - 8192)), True),
                                  (((0, 4096), (8192, big - 8192)), False)):
        gsy.collider.key_windows = windows
        perm, sh, _, dense = gsy.collider.partition(gst, gsy)
        assert bool(dense) == expect_dense
        assert torch.equal(perm, perm0) and torch.equal(sh, sh0)
        gst.force.zero_()
        gsy.collider.compute_force(gst, gsy)
        if expect_dense:
            assert torch.equal(gst.force, f0)
        else:  # sorted fallback: another kernel, another summation order
            assert_close(gst.force, f0.cpu().numpy(), dtype, "force", factor=4)


@pytest.mark.parametrize("domain,rot", [("periodic", ""), ("reflect", "verletspiral")])
def test_system_clock_advances_on_device(domain, rot):
    """System.time / step_count (system.py:62-63) after the single-call driver: the fused sphere
    flow advances them in its setup kernel, the hook-by-hook flow in one tiny launch."""
    inp = make_inputs(2000, 3, seed=2, dtype=np.float32, phi=0.5)
    gst, gsy = build_gpu(inp, dtype=np.float32, domain=domain, rot=rot, dt=1e-3)
    jd_step = __import__("jaxdem_b200").System.step
    jd_step(gst, gsy, n=3)
    jd_step(gst, gsy, n=1)
    want = np.float32(0)
    for _ in range(4):
        want = np.float32(want + np.float32(1e-3))
    assert int(gsy.step_count) == 4
    assert float(gsy.time) == float(want)


@pytest.mark.parametrize("path", __import__("helpers").golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_matches_golden(path):
    """The CUDA path against the committed fixtures (tests/golden/*.npz: oracle outputs on small seeded inputs, see
    make_golden.py): permutation, sorted hashes, neighbour hashes and neighbour lists bit for bit; forces, torques,
    energy and a 3-step trajectory (positions, velocities, angular velocities, quaternions) within the tolerance."""
    import jaxdem_b200 as jd
    from helpers import golden_close, load_golden
    inp, kw, want, meta = load_golden(path)
    gst, gsy = build_gpu(inp, **kw)
    dt = kw["dtype"]
    perm, sh, nh, _ = gsy.collider.partition(gst, gsy)
    golden_close(perm.cpu().numpy(), want["perm"], dt, "perm")
    golden_close(sh.cpu().numpy(), want["sorted_hash"], dt, "sorted_hash")
    golden_close(nh.cpu().numpy(), want["nbr_hash"], dt, "nbr_hash")
    gsy.collider.compute_force(gst, gsy)
    assert_close(gst.force, want["force0"], dt, "force0")
    assert_close(gst.torque, want["torque0"], dt, "torque0")
    _, _, e = gsy.collider.compute_potential_energy(gst, gsy)
    golden_close(e.cpu().numpy(), want["energy0"], dt, "energy0")
    _, _, nl, ovf = gsy.collider.create_neighbor_list(gst, gsy, 1.1, 24)
    golden_close(nl.cpu().numpy(), want["nlist"], dt, "nlist")
    assert bool(ovf) == bool(want["nlist_overflow"])
    jd.System.step(gst, gsy, n=meta["steps"])
    for f in ("pos_c", "vel", "force", "torque", "ang_vel"):
        assert_close(getattr(gst, f), want[f + "_after"], dt, f, factor=4.0 * meta["steps"])
    q = torch.cat([gst.q.w, gst.q.xyz], dim=-1)
    assert_close(q, want["q_after"], dt, "q", factor=4.0 * meta["steps"])
