"""Pin the CPU oracle against every closed-form / scalar pin the reference's own
tests hold for the hot path.  CPU only."""

import numpy as np
import pytest

import oracle
from oracle import colliders, domains, force_manager, integrators, linalg as la


def _sys(state, **kw):
    return oracle.create_system(state, **kw)


def test_periodic_min_image_spring_force():
    # reference tests/test_clump_pair_friction.py:167-186 (k = 1, r = 0.5, box 10)
    st = oracle.create_state([[9.8, 5.0], [0.2, 5.4]], rad=[0.5, 0.5], mass=[1, 1], clump_id=[0, 1])
    mt = oracle.make_material_table([dict(young=1.0, poisson=0.3, density=1.0)])
    for col in ("naive", "celllist"):
        sy = _sys(st, collider_type=col, domain_type="periodic", domain_kw=dict(box_size=[10.0, 10.0]),
                  mat_table=mt)
        s = st.copy()
        colliders.compute_force(s, sy)
        np.testing.assert_allclose(s.force[0], [-0.3071067811865475] * 2, rtol=1e-13)
        np.testing.assert_allclose(s.force[1], [0.3071067811865475] * 2, rtol=1e-13)


def test_overlap_force():
    # reference tests/test_clump_pair_friction.py:190-217: overlap 0.2 -> [-0.2, 0]
    st = oracle.create_state([[0.0, 0.0], [0.8, 0.0]], rad=[0.5, 0.5])
    mt = oracle.make_material_table([dict(young=1.0, poisson=0.3, density=1.0)])
    sy = _sys(st, collider_type="celllist", domain_type="periodic", domain_kw=dict(box_size=[10.0, 10.0]),
              mat_table=mt)
    colliders.compute_force(st, sy)
    np.testing.assert_allclose(st.force[0], [-0.2, 0.0], atol=1e-14)


def test_excluded_pairs():
    # reference tests/test_excluded_pairs.py:11-61
    mt = oracle.make_material_table([dict(young=1000.0, poisson=0.3, density=1.0)], "linear")
    for col in ("naive", "celllist"):
        st = oracle.create_state([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]], rad=[1.1] * 3,
                                 bond_id=[[1], [0, 2], [1]])
        assert st.bond_id.tolist() == [[1, -1], [0, 2], [1, -1]]
        sy = _sys(st, dt=1e-3, collider_type=col, mat_table=mt)
        colliders.compute_force(st, sy)
        assert np.allclose(st.force[1], 0.0, atol=1e-5)
        assert abs(st.force[0, 0]) > 0.1 and abs(st.force[2, 0]) > 0.1
        assert np.allclose(st.force[0], -st.force[2], atol=1e-5)


@pytest.mark.parametrize("rot", ["spiral", "verletspiral"])
def test_aspherical_rotation(rot):
    # reference tests/test_rotation_integrators.py:24-97 (same constants); the
    # body-frame torque is a user force function there - here it is injected as
    # external_torque each step, which ForceManager.apply treats identically.
    rho, Rad, H = 7750.0, 0.05, 0.15
    Mass = rho * np.pi * Rad * Rad * H
    Ix = 0.5 * Mass * Rad * Rad
    Iy = Mass * H * H / 12.0 + 0.25 * Mass * Rad * Rad
    Iz = Iy
    tx, wx0, wy0, wz0 = Rad * 0.5, 0.3, -0.9, 0.6
    st = oracle.create_state([[0.0, 0.0, 0.0]], ang_vel=[[wx0, wy0, wz0]], inertia=[[Ix, Iy, Iz]],
                             mass=[Mass], rad=[Rad])
    dt, steps = 1.0e-5, 70000
    sy = _sys(st, dt=dt, rotation_integrator_type=rot, collider_type="")
    tb = np.array([[tx, 0.0, 0.0]])
    for _ in range(steps):
        # one _step_once with the torque function evaluated inside force_manager.apply
        sy.time += sy.dt
        integrators.LINEAR["verlet"][0](st, sy)
        integrators.ROTATION[rot][0](st, sy)
        colliders.compute_force(st, sy)
        sy.force_manager.external_torque = la.q_rotate(st.q_w, st.q_xyz, tb)
        force_manager.apply(st, sy)
        integrators.LINEAR["verlet"][1](st, sy)
        integrators.ROTATION[rot][1](st, sy)
    t = 0.7
    A = (Ix - Iy) * (Iz - Ix) / (Iy * Iz)
    B = Iy / (Iz - Ix)
    E = 2.0 * tx * B / Ix
    wx = wx0 + tx * t / Ix
    eta = 0.5 * Ix * np.sqrt(-A) / tx
    C = E * eta
    K1 = (C * wy0 * np.cos(eta * wx0**2) - wz0 * np.sin(eta * wx0**2)) / C
    K2 = (C * wy0 * np.sin(eta * wx0**2) + wz0 * np.cos(eta * wx0**2)) / C
    D = eta * wx**2
    ww = np.array([wx, K1 * np.cos(D) + K2 * np.sin(D), C * (K2 * np.cos(D) - K1 * np.sin(D))])
    w_body = la.q_rotate_back(st.q_w, st.q_xyz, st.ang_vel)[0]
    err = np.log10(np.linalg.norm(w_body - ww) / np.linalg.norm(ww))
    assert err < -4.0, err


def _spheres(dim, poly, domain, seed=0, n=6, dtype=np.float64):
    # fixture recipe of reference tests/test_colliders_invariance.py:50-110
    rng = np.random.default_rng(seed)
    spacing = 1.2
    st = oracle.grid_state((n,) * dim, spacing, seed=seed, dtype=dtype,
                           radius_range=(0.5 / poly, 0.5) if poly > 1 else None, radius=0.5)
    st.vel = (st.vel * 1.0).astype(dtype)
    box = np.full(dim, spacing * n)
    anchor = np.full(dim, -0.5)
    mt = oracle.make_material_table([dict(young=1.0e3, poisson=0.3, density=1.0)])
    return st, dict(dt=1e-3, domain_type=domain, domain_kw=dict(box_size=box, anchor=anchor), mat_table=mt)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("domain", ["periodic", "reflect", "free"])
@pytest.mark.parametrize("poly", [1, 3])
def test_celllist_matches_naive(dim, domain, poly):
    # reference tests/test_colliders_invariance.py:365-402 compares after 200 steps
    # at 5e-2; a single oracle implements both colliders so we can be strict.
    st0, kw = _spheres(dim, poly, domain, n=5 if dim == 3 else 8)
    a, b = st0.copy(), st0.copy()
    sa = oracle.create_system(a, collider_type="naive", **kw)
    sb = oracle.create_system(b, collider_type="celllist", collider_kw=dict(state=b), **kw)
    oracle.step(a, sa, 40)
    oracle.step(b, sb, 40)
    for f in ("pos_c", "vel", "force", "ang_vel", "torque", "q_w", "q_xyz"):
        np.testing.assert_allclose(getattr(a, f), getattr(b, f), rtol=1e-9, atol=1e-9, err_msg=f)
    ea = colliders.compute_potential_energy(a, sa)
    eb = colliders.compute_potential_energy(b, sb)
    np.testing.assert_allclose(ea, eb, rtol=1e-9, atol=1e-12)


def test_pos_p_rot_cache_after_steps():
    # reference tests/test_state_cache.py:65-92
    rng = np.random.default_rng(3)
    n = 12
    pos = rng.uniform(0, 6, (n, 3))
    st = oracle.create_state(pos, pos_p=rng.normal(0, 0.1, (n, 3)), ang_vel=rng.normal(0, 1, (n, 3)),
                             vel=rng.normal(0, 1, (n, 3)), rad=np.full(n, 0.4))
    sy = _sys(st, dt=1e-3, collider_type="naive", domain_type="free")
    oracle.step(st, sy, 300)
    np.testing.assert_allclose(st._pos_p_rot, la.q_rotate(st.q_w, st.q_xyz, st.pos_p), rtol=0, atol=0)


def test_neighbor_list_matches_bruteforce():
    rng = np.random.default_rng(5)
    n, L = 300, 8.0
    st = oracle.create_state(rng.uniform(0, L, (n, 3)), rad=np.full(n, 0.5))
    sy = _sys(st, collider_type="celllist", collider_kw=dict(state=st), domain_type="periodic",
              domain_kw=dict(box_size=[L] * 3))
    nl, ovf = colliders.celllist_create_neighbor_list(st, sy, cutoff=1.3, max_neighbors=40)
    assert not ovf
    pos = st.pos
    for i in range(n):
        d = sy.domain.displacement(pos[i], pos)
        want = set(np.nonzero((np.sum(d * d, -1) <= 1.3**2) & (np.arange(n) != i))[0].tolist())
        got = [x for x in nl[i].tolist() if x >= 0]
        assert len(got) == len(set(got)) and set(got) == want


def test_neighbor_list_overflow_flag():
    rng = np.random.default_rng(5)
    st = oracle.create_state(rng.uniform(0, 3, (200, 3)), rad=np.full(200, 0.5))
    sy = _sys(st, collider_type="celllist", collider_kw=dict(state=st), domain_type="periodic",
              domain_kw=dict(box_size=[3.0] * 3))
    nl, ovf = colliders.celllist_create_neighbor_list(st, sy, cutoff=1.0, max_neighbors=4)
    assert ovf and nl.shape == (200, 4)


@pytest.mark.parametrize("lin,rot,thresh", [("verlet", "verletspiral", 1.8), ("euler", "spiral", 0.7)])
def test_energy_drift_slope(lin, rot, thresh):
    # reference tests/test_energy_conservation.py:34-88 (shortened: 3 dts, 1.0 time units)
    dts = [2e-3, 1e-3, 5e-4]
    stds = []
    for dt in dts:
        st, kw = _spheres(2, 1, "periodic", n=6)
        kw["dt"] = dt
        sy = oracle.create_system(st, collider_type="naive", linear_integrator_type=lin,
                                  rotation_integrator_type=rot, **kw)
        es = []
        for _ in range(int(round(1.0 / dt / 20))):
            oracle.step(st, sy, 20)
            es.append(oracle.system.total_energy(st, sy))
        stds.append(np.std(es))
    slope = np.polyfit(np.log(dts), np.log(stds), 1)[0]
    assert slope > thresh, (slope, stds)


def test_reflect_keeps_particles_inside_and_conserves_energy_elastic():
    st, kw = _spheres(2, 1, "reflect", n=5)
    st.vel = st.vel * 3.0
    sy = oracle.create_system(st, collider_type="celllist", collider_kw=dict(state=st), **kw)
    e0 = oracle.system.total_energy(st, sy)
    oracle.step(st, sy, 1500)
    lo = sy.domain.anchor + st.rad[:, None] - 0.05
    hi = sy.domain.anchor + sy.domain.box_size - st.rad[:, None] + 0.05
    assert np.all(st.pos_c > lo) and np.all(st.pos_c < hi)
    e1 = oracle.system.total_energy(st, sy)
    assert abs(e1 - e0) / e0 < 1e-2


def test_f32_dtypes_preserved():
    st, kw = _spheres(3, 2, "periodic", n=4, dtype=np.float32)
    sy = oracle.create_system(st, collider_type="celllist", collider_kw=dict(state=st),
                              force_model_type="cundallstrack",
                              **{**kw, "mat_table": oracle.make_material_table(
                                  [dict(young=1e3, poisson=0.3, density=1.0, mu=0.5, e=0.8, mu_r=0.05)])})
    oracle.step(st, sy, 5)
    for f in ("pos_c", "vel", "force", "ang_vel", "torque", "q_w", "q_xyz", "_pos_p_rot"):
        assert getattr(st, f).dtype == np.float32, f
    assert st.clump_id.dtype == np.int32 and sy.collider.neighbor_mask.dtype == np.int32


@pytest.mark.parametrize("dim,domain", [(3, "periodic"), (2, "periodic"), (3, "free")])
def test_cross_neighbor_list_against_brute_force(dim, domain):
    """create_cross_neighbor_list (cell_list.py:600-715): as sets, the rows equal the brute-force
    answer ||displacement(a, b)||^2 <= cutoff^2; the empty-input conventions of :627-636 hold."""
    from oracle import colliders as ocol
    from helpers import build_oracle, make_inputs
    inp = make_inputs(500, dim, seed=4, dtype=np.float64, phi=0.5)
    ost, osy = build_oracle(inp, dtype=np.float64, domain=domain)
    rng = np.random.default_rng(8)
    pos_a = rng.uniform(-0.1, 1.1, (137, dim)) * inp["box"]  # some queries outside the box
    if domain != "periodic":
        pos_a = np.clip(pos_a, 0.0, inp["box"] * 0.999)
    cutoff = 1.3
    nl, ovf = ocol.celllist_create_cross_neighbor_list(pos_a, ost.pos, osy, cutoff, 64)
    assert not ovf and nl.shape == (137, 64)
    for i in range(137):
        d = osy.domain.displacement(pos_a[i][None, :], ost.pos)
        want = set(np.nonzero((d * d).sum(-1) <= cutoff**2)[0].tolist())
        got = [j for j in nl[i].tolist() if j >= 0]
        assert len(got) == len(set(got)) and set(got) == want, i
    nl2, ovf2 = ocol.celllist_create_cross_neighbor_list(pos_a, ost.pos, osy, cutoff, 2)
    assert ovf2 and nl2.shape == (137, 2)
    assert ocol.celllist_create_cross_neighbor_list(pos_a[:0], ost.pos, osy, cutoff, 8)[0].shape == (0, 8)
    e = ocol.celllist_create_cross_neighbor_list(pos_a, ost.pos[:0], osy, cutoff, 8)
    assert (e[0] == -1).all() and not e[1]


def test_neighborlist_oracle_pins():
    # reference tests/test_clump_pair_friction.py:190-217: a fresh NeighborList is built before it is read
    # (force [-0.2, 0] for overlap 0.2, n_build_times == 1); tests/test_excluded_pairs.py:11-61 through the
    # NeighborList over a CellList; tests/test_colliders_invariance.py: same forces as the naive collider.
    mt = oracle.make_material_table([dict(young=1.0, poisson=0.3, density=1.0)])
    st = oracle.create_state([[0.0, 0.0], [0.8, 0.0]], rad=[0.5, 0.5], mass=[1, 1], clump_id=[0, 1])
    sy = _sys(st, collider_type="neighborlist", collider_kw=dict(cutoff=1.5, skin=0.1, max_neighbors=10),
              domain_type="periodic", domain_kw=dict(box_size=[10.0, 10.0]), mat_table=mt)
    assert sy.collider.max_neighbors == 2 and (sy.collider.neighbor_list == -1).all()
    colliders.compute_force(st, sy)
    np.testing.assert_allclose(st.force[0], [-0.2, 0.0], atol=1e-14)
    assert sy.collider.n_build_times == 1
    colliders.compute_force(st, sy)
    assert sy.collider.n_build_times == 1  # nothing moved
    st.pos_c = st.pos_c + np.array([[0.051, 0.0], [0.0, 0.0]])  # further than skin / 2
    colliders.compute_force(st, sy)
    assert sy.collider.n_build_times == 2
    mt = oracle.make_material_table([dict(young=1000.0, poisson=0.3, density=1.0)], "linear")
    st = oracle.create_state([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]], rad=[1.1] * 3, bond_id=[[1], [0, 2], [1]])
    sy = _sys(st, dt=1e-3, collider_type="neighborlist", collider_kw=dict(cutoff=3.0, skin=0.5), mat_table=mt)
    colliders.compute_force(st, sy)
    assert np.allclose(st.force[1], 0.0, atol=1e-5) and abs(st.force[0, 0]) > 0.1
    assert np.allclose(st.force[0], -st.force[2], atol=1e-5)
    rng = np.random.default_rng(3)
    for dim in (2, 3):
        n = 300
        L = (n / 0.9) ** (1 / dim)
        pos = rng.uniform(0, L, (n, dim))
        a = oracle.create_state(pos, rad=rng.uniform(0.35, 0.5, n))
        b = a.copy()
        kw = dict(domain_type="periodic", domain_kw=dict(box_size=[L] * dim))
        sa = _sys(a, collider_type="neighborlist", collider_kw=dict(cutoff=1.0, skin=0.2), **kw)
        sb = _sys(b, collider_type="naive", **kw)
        colliders.compute_force(a, sa)
        colliders.compute_force(b, sb)
        assert not sa.collider.overflow
        np.testing.assert_allclose(a.force, b.force, atol=1e-9 * np.abs(b.force).max())
        np.testing.assert_allclose(colliders.compute_potential_energy(a, sa),
                                   colliders.compute_potential_energy(b, sb), rtol=1e-12)


@pytest.mark.parametrize("dim,domain", [(2, "periodic"), (3, "periodic"), (3, "reflect")])
def test_multicelllist_oracle_equals_celllist(dim, domain):
    # reference multi_cell_list.py:283-289: "forces are bit-identical to DynamicCellList" (the prune only skips
    # contact-free cells); neighbour lists likewise.  Polydisperse clumps, cells of 2 r_max.
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import build_oracle, make_inputs
    inp = make_inputs(900, dim, seed=31, dtype=np.float64, phi=0.5, poly=3.0, clumps=True)
    a, sa = build_oracle(inp, dtype=np.float64, collider="multicelllist", law="hertz", domain=domain)
    b, sb = build_oracle(inp, dtype=np.float64, collider="celllist", law="hertz", domain=domain)
    assert float(sa.collider.cell_size) != float(sb.collider.cell_size)  # the defaults differ (alpha = 3 > 2.5)
    sb.collider.cell_size, sb.collider.neighbor_mask = sa.collider.cell_size, sa.collider.neighbor_mask
    colliders.compute_force(a, sa)
    colliders.compute_force(b, sb)
    assert np.array_equal(a.force, b.force) and np.array_equal(a.torque, b.torque)
    assert colliders.compute_potential_energy(a, sa) == colliders.compute_potential_energy(b, sb)
    na, oa = colliders.celllist_create_neighbor_list(a, sa, 0.8, 30)
    nb, ob = colliders.celllist_create_neighbor_list(b, sb, 0.8, 30)
    assert np.array_equal(na, nb) and oa == ob
    # and the prune really skips cells: some stencil cells of some particles hold members out of reach
    pos, perm, sh, nh, _ = colliders._partition_for(a, sa, sa.collider.cell_size)
    rad = a._rad[:, None]
    cc, ch = colliders.loose_cell_aabbs((pos - rad)[perm], (pos + rad)[perm], sh)
    start = np.minimum(np.searchsorted(sh, nh[:, 0], side="left"), a.N - 1)
    hit = sh[start] == nh[:, 0]
    dr = sa.domain.displacement(pos, cc[start])
    skipped = hit & ~np.all(np.abs(dr) <= rad + ch[start], axis=-1)
    assert skipped.any()
