"""GPU parity tests of the loose-grid MultiCellList collider (reference
jaxdem/colliders/multi_cell_list.py) against the numpy oracle, and its defining property: the AABB prune never
changes a result (forces / energies equal the cell list's; neighbour lists identical)."""

import numpy as np
import pytest
import torch

import oracle
from oracle import colliders as ocol
from helpers import assert_close, build_gpu, build_oracle, compare_states, make_inputs

pytestmark = pytest.mark.gpu

DT = [np.float32, np.float64]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("law", ["spring", "cundallstrack"])
@pytest.mark.parametrize("domain,mode", [("periodic", "auto"), ("reflect", "auto"), ("periodic", "sorted")])
def test_multicell_force_energy(dtype, dim, law, domain, mode):
    inp = make_inputs(3000, dim, seed=13, dtype=dtype, phi=0.5, poly=3.0, nmat=2, clumps=(law == "spring"))
    ost, osy = build_oracle(inp, dtype=dtype, collider="multicelllist", law=law, domain=domain, nmat=2)
    gst, gsy = build_gpu(inp, dtype=dtype, collider="MultiCellList", law=law, domain=domain, nmat=2, grid_mode=mode)
    assert float(gsy.collider.cell_size) == float(osy.collider.cell_size)  # 2 r_max although r_max / r_min = 3
    assert np.array_equal(gsy.collider.neighbor_mask.cpu().numpy(), osy.collider.neighbor_mask)
    ocol.compute_force(ost, osy)
    gsy.collider.compute_force(gst, gsy)
    assert_close(gst.force, ost.force, dtype, "force", factor=2.0)
    assert_close(gst.torque, ost.torque, dtype, "torque", factor=2.0, scale=float(np.abs(ost.force).max()))
    e_o = ocol.compute_potential_energy(ost, osy)
    _, _, e_g = gsy.collider.compute_potential_energy(gst, gsy)
    assert abs(float(e_g) - float(e_o)) <= 20 * (1e-5 if dtype == np.float32 else 1e-12) * abs(float(e_o))
    # the prune drops only contact-free cells: the cell list on the same grid gives the same sums
    cst, csy = build_gpu(inp, dtype=dtype, collider="CellList", law=law, domain=domain, nmat=2, grid_mode=mode,
                         collider_kw=dict(cell_size=float(gsy.collider.cell_size)))
    csy.collider.compute_force(cst, csy)
    assert_close(gst.force, cst.force.cpu().numpy(), dtype, "force vs cell list", factor=2.0)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("K", [40, 5])
def test_multicell_neighbor_lists_bit_exact(dtype, dim, K):
    inp = make_inputs(2000, dim, seed=5, dtype=dtype, phi=0.6, poly=2.0, bonds=True)
    ost, osy = build_oracle(inp, dtype=dtype, collider="multicelllist")
    gst, gsy = build_gpu(inp, dtype=dtype, collider="MultiCellList")
    nl_o, ovf_o = ocol.celllist_create_neighbor_list(ost, osy, 0.9, K)
    _, _, nl_g, ovf_g = gsy.collider.create_neighbor_list(gst, gsy, 0.9, K)
    assert np.array_equal(nl_g.cpu().numpy(), nl_o) and bool(ovf_g) == bool(ovf_o)
    rng = np.random.default_rng(2)
    qa = rng.uniform(0, float(inp["box"][0]), (700, dim)).astype(dtype)
    x_o, xo = ocol.celllist_create_cross_neighbor_list(qa, ost.pos, osy, 0.9, K, idtype=ost.idtype)
    x_g, xg = gsy.collider.create_cross_neighbor_list(torch.as_tensor(qa, device="cuda"), gst.pos, gsy, 0.9, K)
    assert np.array_equal(x_g.cpu().numpy(), x_o) and bool(xg) == bool(xo)


def test_multicell_steps_and_minimises():
    import jaxdem_b200 as jd
    from oracle import minimizers as omin
    dtype = np.float64
    inp = make_inputs(800, 3, seed=21, dtype=dtype, phi=0.5, poly=2.5, clumps=True)
    kw = dict(dtype=dtype, law="hertz", rot="verletspiral", dt=1e-3)
    ost, osy = build_oracle(inp, collider="multicelllist", **kw)
    gst, gsy = build_gpu(inp, collider="MultiCellList", **kw)
    oracle.step(ost, osy, 3)
    jd.System.step(gst, gsy, n=3)  # jdb200_system_step, hook by hook inside one call
    compare_states(gst, ost, dtype, factor=20.0)
    ost, osy = build_oracle(inp, collider="multicelllist", dtype=dtype, law="spring", dt=1e-2)
    gst, gsy = build_gpu(inp, collider="MultiCellList", dtype=dtype, law="spring", dt=1e-2)
    s_o, pe_o, _ = omin.minimize(ost, osy, omin.FireConfig(1e-2), max_steps=12, pe_tol=0.0, pe_diff_tol=0.0, force_tol=-1.0)
    _, _, s_g, pe_g = jd.System.minimize(gst, gsy, max_steps=12, pe_tol=0.0, pe_diff_tol=0.0, force_tol=-1.0)
    assert int(s_g) == s_o == 12 and abs(float(pe_g) - pe_o) <= 1e-9 * abs(pe_o)
