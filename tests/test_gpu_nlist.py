"""GPU parity tests of the Verlet NeighborList collider (reference
jaxdem/colliders/neighbor_list.py) against the numpy oracle: list rows bit-exact at every build, the device-side
rebuild decision (n_build_times) identical, forces / torques / energy within the float tolerance."""

import numpy as np
import pytest
import torch

import oracle
from oracle import colliders as ocol
from helpers import assert_close, build_gpu, build_oracle, compare_states, make_inputs

pytestmark = pytest.mark.gpu

DT = [np.float32, np.float64]


def _pair(inp, dtype, cutoff, skin, **kw):
    ckw = dict(cutoff=cutoff, skin=skin)
    if "K" in kw:
        ckw["max_neighbors"] = kw.pop("K")
    ost, osy = build_oracle(inp, dtype=dtype, collider="neighborlist", collider_kw=ckw, **kw)
    gst, gsy = build_gpu(inp, dtype=dtype, collider="NeighborList", collider_kw=ckw, **kw)
    return ost, osy, gst, gsy


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("law", ["spring", "hertz", "cundallstrack"])
@pytest.mark.parametrize("domain", ["periodic", "reflect"])
def test_nl_force_energy_and_rows(dtype, dim, law, domain):
    inp = make_inputs(2500, dim, seed=31, dtype=dtype, phi=0.6, poly=1.5, nmat=2, clumps=(law == "cundallstrack"),
                      bonds=(law == "hertz"))
    ost, osy, gst, gsy = _pair(inp, dtype, 1.0, 0.15, law=law, domain=domain, nmat=2)
    assert gsy.collider.max_neighbors == osy.collider.max_neighbors
    ocol.compute_force(ost, osy)
    gsy.collider.compute_force(gst, gsy)
    assert int(gsy.collider.n_build_times) == osy.collider.n_build_times == 1
    assert np.array_equal(gsy.collider.neighbor_list.cpu().numpy(), osy.collider.neighbor_list)
    assert bool(gsy.collider.overflow) == bool(osy.collider.overflow)
    assert torch.equal(gsy.collider.old_pos, gst.pos)  # old_pos <- state.pos, bit for bit
    assert_close(gsy.collider.old_pos, osy.collider.old_pos, dtype, "old_pos")
    assert_close(gst.force, ost.force, dtype, "force", factor=2.0)
    assert_close(gst.torque, ost.torque, dtype, "torque", factor=2.0, scale=float(np.abs(ost.force).max()))
    e_o = ocol.compute_potential_energy(ost, osy)
    _, _, e_g = gsy.collider.compute_potential_energy(gst, gsy)
    assert abs(float(e_g) - float(e_o)) <= 20 * (1e-5 if dtype == np.float32 else 1e-12) * abs(float(e_o))
    assert int(gsy.collider.n_build_times) == 1  # nothing moved: the second and third call kept the list


@pytest.mark.parametrize("dtype", DT)
def test_nl_equals_celllist_forces(dtype):
    # reference tests/test_colliders_invariance.py: NeighborList over CellList gives the cell list's forces
    inp = make_inputs(4000, 3, seed=5, dtype=dtype, phi=0.55)
    _, _, gst, gsy = _pair(inp, dtype, 1.0, 0.1, law="spring")
    cst, csy = build_gpu(inp, dtype=dtype, law="spring")
    gsy.collider.compute_force(gst, gsy)
    csy.collider.compute_force(cst, csy)
    assert_close(gst.force, cst.force.cpu().numpy(), dtype, "force", factor=2.0)


def test_nl_rebuild_decision_over_a_run_f64():
    # 60 steps of a hot gas with a thin skin: the list must be rebuilt several times, at exactly the steps the
    # oracle rebuilds (same arithmetic for max |pos - old_pos|^2 > skin^2 / 4), and the trajectory must follow.
    import jaxdem_b200 as jd
    dtype = np.float64
    inp = make_inputs(1500, 3, seed=8, dtype=dtype, phi=0.5)
    inp["vel"] *= 3.0
    ost, osy, gst, gsy = _pair(inp, dtype, 1.0, 0.06, law="spring", rot="", dt=2e-3)
    hst, hsy = build_gpu(inp, dtype=dtype, collider="NeighborList", collider_kw=dict(cutoff=1.0, skin=0.06),
                         law="spring", rot="", dt=2e-3)
    builds = []
    for step in range(60):
        oracle.step(ost, osy, 1)
        jd.System.step(gst, gsy, n=1)               # jdb200_system_step_nl
        jd.System.step(hst, hsy, n=1, fused=False)  # hook by hook
        builds.append((osy.collider.n_build_times, int(gsy.collider.n_build_times), int(hsy.collider.n_build_times)))
        assert builds[-1][0] == builds[-1][1] == builds[-1][2], (step, builds[-1])
    assert 3 <= builds[-1][0] < 60, builds[-1]
    assert np.array_equal(gsy.collider.neighbor_list.cpu().numpy(), osy.collider.neighbor_list)
    compare_states(gst, ost, dtype, factor=1000.0, fields=("pos_c", "vel", "force"))
    for f in ("pos_c", "vel", "force"):
        assert torch.equal(getattr(gst, f), getattr(hst, f)), f


def test_nl_batched_systems_decide_independently():
    # under a batch axis each system rebuilds when IT must (the reference's cond becomes a per-element select)
    import jaxdem_b200 as jd
    dtype = np.float64
    singles = []
    rng = np.random.default_rng(40)
    gx, gy = np.meshgrid(np.arange(25), np.arange(24), indexing="ij")
    for b in range(3):  # lattice gas, no initial overlaps; speeds 0.2 : 1 : 5
        n = 600
        inp = dict(pos=1.2 * np.stack([gx.ravel(), gy.ravel()], 1).astype(dtype) + 0.6, rad=rng.uniform(0.3, 0.5, n),
                   vel=rng.uniform(-1, 1, (n, 2)) * (0.2, 1.0, 5.0)[b], mass=np.ones(n), box=np.array([30.0, 28.8]))
        st, sy = build_gpu(inp, dtype=dtype, collider="NeighborList", collider_kw=dict(cutoff=1.0, skin=0.2, max_neighbors=11),
                           rot="", dt=2e-3)
        singles.append((st, sy, inp))
    stb = jd.State.stack([s[0].clone() for s in singles])
    syb = jd.System.create(stb.shape, dt=2e-3, collider_type="NeighborList",
                           collider_kw=dict(state=stb, cutoff=1.0, skin=0.2, max_neighbors=11),
                           domain_type="periodic", domain_kw=dict(box_size=np.stack([s[2]["box"] for s in singles])),
                           rotation_integrator_type="",
                           mat_table=singles[0][1].mat_table, dtype=torch.float64)
    jd.System.step(stb, syb, n=100)
    counts = []
    for b, (st, sy, _) in enumerate(singles):
        jd.System.step(st, sy, n=100)
        counts.append(int(sy.collider.n_build_times))
        assert int(syb.collider.n_build_times[b]) == counts[-1], (b, syb.collider.n_build_times, counts)
        for f in ("pos_c", "vel", "force"):
            assert torch.equal(getattr(stb, f)[b], getattr(st, f)), (b, f)
        assert torch.equal(syb.collider.neighbor_list[b], sy.collider.neighbor_list)
    assert counts[0] < counts[1] < counts[2], counts


def test_nl_reference_pins():
    # reference tests/test_clump_pair_friction.py:190-217 (overlap 0.2 -> force [-0.2, 0], n_build_times == 1) and
    # tests/test_excluded_pairs.py:11-61 (bond exclusion through the NeighborList over a CellList)
    import jaxdem_b200 as jd
    mt = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1.0, poisson=0.3, density=1.0)])
    st = jd.State.create([[0.0, 0.0], [0.8, 0.0]], rad=[0.5, 0.5], mass=[1.0, 1.0], clump_id=[0, 1], dtype=torch.float64)
    sy = jd.System.create(st.shape, collider_type="NeighborList",
                          collider_kw=dict(state=st, cutoff=1.5, skin=0.1, max_neighbors=10),
                          domain_type="periodic", domain_kw=dict(box_size=[10.0, 10.0]), mat_table=mt, dtype=torch.float64)
    assert bool((sy.collider.neighbor_list == -1).all())  # the fresh cache is all padding
    sy.collider.compute_force(st, sy)
    np.testing.assert_allclose(st.force[0].cpu().numpy(), [-0.2, 0.0], atol=1e-14)
    assert int(sy.collider.n_build_times) == 1
    mt = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1000.0, poisson=0.3, density=1.0)],
                                         matcher=jd.MaterialMatchmaker.create("linear"))
    st = jd.State.create([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]], rad=[1.1] * 3, bond_id=[[1], [0, 2], [1]], dtype=torch.float64)
    sy = jd.System.create(st.shape, dt=1e-3, collider_type="NeighborList",
                          collider_kw=dict(state=st, cutoff=3.0, skin=0.5, secondary_collider_type="CellList",
                                           secondary_collider_kw=dict(state=st)), mat_table=mt, dtype=torch.float64)
    sy.collider.compute_force(st, sy)
    f = st.force.cpu().numpy()
    assert np.allclose(f[1], 0.0, atol=1e-5) and abs(f[0, 0]) > 0.1 and np.allclose(f[0], -f[2], atol=1e-5)
    # create_neighbor_list returns the cached list; K clamps to N (neighbor_list.py:377-379)
    _, _, nl, ovf = sy.collider.create_neighbor_list(st, sy, 0.0, 0)
    assert nl.shape == (3, 3) and not bool(ovf)
    assert nl.tolist() == [[2, -1, -1], [-1, -1, -1], [0, -1, -1]]  # bonded pairs never enter the list (cell_list.py:560-568)
    with pytest.raises(ValueError):
        jd.Collider.create("NeighborList", state=st, cutoff=1.0, skin=0.1, skin_fraction=0.1)


def test_nl_overflow_flag_and_compile_step():
    import jaxdem_b200 as jd
    dtype = np.float32
    inp = make_inputs(3000, 3, seed=12, dtype=dtype, phi=0.7)
    ost, osy, gst, gsy = _pair(inp, dtype, 1.0, 0.2, K=6, rot="")
    ocol.compute_force(ost, osy)
    gsy.collider.compute_force(gst, gsy)
    assert osy.collider.overflow and bool(gsy.collider.overflow)
    assert np.array_equal(gsy.collider.neighbor_list.cpu().numpy(), osy.collider.neighbor_list)
    # graph capture of NL steps: replay == stream, including rebuilds decided inside the replay
    a_st, a_sy = build_gpu(inp, dtype=dtype, collider="NeighborList", collider_kw=dict(cutoff=1.0, skin=0.05), rot="", dt=2e-3)
    b_st, b_sy = build_gpu(inp, dtype=dtype, collider="NeighborList", collider_kw=dict(cutoff=1.0, skin=0.05), rot="", dt=2e-3)
    step = jd.System.compile_step(a_st, a_sy, n=5)
    for _ in range(6):
        step()
    jd.System.step(b_st, b_sy, n=30)
    for f in ("pos_c", "vel", "force"):
        assert torch.equal(getattr(a_st, f), getattr(b_st, f)), f
    assert int(a_sy.collider.n_build_times) == int(b_sy.collider.n_build_times) > 1
