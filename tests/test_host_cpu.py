"""CPU-side tests: the C-ABI library loads and exports every declared symbol, the host
mirror of the plugin API behaves like the reference's, and the product path refuses to
run without CUDA.  No compute calls (there is no GPU here)."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

import jaxdem_b200 as jd
from jaxdem_b200 import _call, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "jaxdem_b200.h")).read()
    declared = set(re.findall(r"JDB200_API\s+[\w\s]+?\b(jdb200_\w+)\s*\(", header))
    assert declared and declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.jdb200_abi_version() == _lib.ABI_VERSION


def test_params_struct_layout_matches_header():
    header = open(os.path.join(ROOT, "include", "jaxdem_b200.h")).read()
    body = re.search(r"typedef struct jdb200_params \{(.*?)\} jdb200_params;", header, re.S).group(1)
    names = re.findall(r"int(?:32|64)_t\s+(\w+)(?:\[\d+\])?;", body)
    assert names == [f[0] for f in _lib.Params._fields_]
    for struct, tag in ((_lib.StateView, "jdb200_state"), (_lib.SystemView, "jdb200_system")):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (tag, tag), header, re.S).group(1)
        names = re.findall(r"void\*\s+(\w+);", body)
        assert names == [f[0] for f in struct._fields_], tag


def test_workspace_bytes_and_argument_errors():
    lib = _lib.lib()
    st = jd.utils.grid_state(n_per_axis=(6, 6, 6), spacing=1.0, radius=0.5, device="cpu")
    sy = jd.System.create(st.shape, collider_type="CellList", collider_kw=dict(state=st), domain_type="periodic",
                          domain_kw=dict(box_size=[6.0] * 3), device="cpu")
    p = _call.params_for(st, sy)
    small = lib.jdb200_workspace_bytes(ctypes.byref(p))
    assert small > 0
    p.n = 1 << 20
    p.max_cells = 4 << 20
    big = lib.jdb200_workspace_bytes(ctypes.byref(p))
    assert 64e6 < big < 1e9
    p.dim = 4
    assert lib.jdb200_workspace_bytes(ctypes.byref(p)) == 0
    sv, yv = _lib.StateView(), _lib.SystemView()
    assert lib.jdb200_celllist_compute_force(None, ctypes.byref(p), ctypes.byref(sv), ctypes.byref(yv), None, 0) == -1
    p.dim = 3
    assert lib.jdb200_celllist_compute_force(None, ctypes.byref(p), ctypes.byref(sv), ctypes.byref(yv), None, 0) == -2
    buf = ctypes.create_string_buffer(16)
    assert lib.jdb200_celllist_compute_force(None, ctypes.byref(p), ctypes.byref(sv), ctypes.byref(yv), buf, 16) == -3


def test_no_cpu_fallback():
    st = jd.utils.grid_state(n_per_axis=(4, 4), spacing=1.0, radius=0.5, device="cpu")
    sy = jd.System.create(st.shape, collider_type="CellList", collider_kw=dict(state=st), device="cpu")
    for fn in (lambda: jd.System.step(st, sy, n=1), lambda: sy.collider.compute_force(st, sy),
               lambda: sy.force_manager.apply(st, sy), lambda: sy.linear_integrator.step_before_force(st, sy)):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()


def test_registry_keys_and_normalisation():
    # reference tests/test_public_api.py:24-35
    assert type(jd.Collider.create("cell_list", state=jd.State.create(np.zeros((3, 2)), device="cpu"))).__name__ == "DynamicCellList"
    assert jd.Collider.create("Cell-List", state=jd.State.create(np.zeros((3, 2)), device="cpu")).type_name == "celllist"
    assert jd.LinearIntegrator.create("Verlet").type_name == "verlet"
    assert jd.RotationIntegrator.create("verlet_spiral").type_name == "verletspiral"
    assert jd.Domain.create("Periodic", dim=2).periodic and not jd.Domain.create("reflect", dim=2).periodic
    assert jd.ForceModel.create("cundall strack").type_name == "cundallstrack"
    with pytest.raises(KeyError):
        jd.Collider.create("nope")
    with pytest.warns(UserWarning):
        jd.LinearIntegrator.create("verlet", bogus=1)
    with pytest.raises(ValueError):
        jd.Domain.create("periodic", dim=3, box_size=[1.0, 1.0])
    # separate registries per root
    assert "verlet" in jd.LinearIntegrator._registry and "verlet" not in jd.RotationIntegrator._registry


def test_state_create_defaults():
    # reference State.create defaults (state.py:647-867)
    st = jd.State.create(np.zeros((5, 3)), rad=[1, 2, 1, 1, 1], clump_id=[7, 7, 3, 9, 9],
                         bond_id=[[1], [], [4], [], []], device="cpu")
    assert st.clump_id.tolist() == [1, 1, 0, 2, 2] and st.has_clumps
    assert st.bond_id.tolist() == [[1], [0], [4], [-1], [2]]
    assert st.q.w.shape == (5, 1) and st.q.xyz.shape == (5, 3) and st.ang_vel.shape == (5, 3)
    np.testing.assert_allclose(st.inertia[1].numpy(), 0.4 * 1.0 * 4.0)
    assert st.dtype == torch.float32 and st.clump_id.dtype == torch.int32 and st.fixed.dtype == torch.bool
    st2 = jd.State.create(np.zeros((4, 2)), dtype=torch.float64, device="cpu")
    assert st2.ang_vel.shape == (4, 1) and st2.clump_id.dtype == torch.int64 and not st2.has_clumps
    st2.q = jd.Quaternion(torch.full((4, 1), np.cos(0.3)), torch.tensor([[0, 0, np.sin(0.3)]] * 4))
    assert st2._pos_p_rot.shape == (4, 2)  # cache refreshed on assignment (state.py:264-273)


def test_celllist_create_matches_oracle_create():
    import oracle
    from oracle import colliders as ocol
    rng = np.random.default_rng(0)
    for poly, box in ((1.0, None), (3.0, None), (1.2, [2.5, 2.5, 2.5])):
        rad = rng.uniform(0.5 / poly, 0.5, 50)
        ost = oracle.create_state(rng.uniform(0, 5, (50, 3)), rad=rad, dtype=np.float32)
        oc = ocol.celllist_create(ost, box_size=box)
        gc = jd.Collider.create("CellList", state=jd.State.create(ost.pos_c, rad=rad, device="cpu"), box_size=box)
        assert np.array_equal(gc.neighbor_mask.numpy(), oc.neighbor_mask)
        assert float(gc.cell_size) == float(oc.cell_size)


def test_slab_layout_and_key_windows():
    """Slab layout (whole cell layers per rank) and the dense-table windows a rank may touch."""
    import numpy as np
    from jaxdem_b200.slab import SlabLayout, key_windows, message_layout
    lay = SlabLayout(412, 4, 1)
    assert lay.bounds == [0, 103, 206, 309, 412] and (np.bincount(lay.owner) == 103).all()
    with pytest.raises(ValueError):
        SlabLayout(10, 4, 1)  # slabs thinner than 2 R + 1 layers
    S, G, rows = 103 * 103, 412, 412 * 103 * 103 + 2 * 103 * 103 + 1025
    # interior rank: one window, 2 R + 1 layers of margin on both sides (+ the end-of-run entry)
    assert key_windows(103, 206, 1, G, S, rows) == (((103 - 3) * S, (103 + 7) * S + 1), (0, 0))
    # first / last rank: the margin wraps around the periodic box -> two windows, the first
    # a multiple of the scan tile (4096), disjoint and ascending
    for lo, up in ((0, 103), (309, 412)):
        (a_lo, a_len), (b_lo, b_len) = key_windows(lo, up, 1, G, S, rows)
        assert a_lo == 0 and a_len % 4096 == 0 and a_len < b_lo and b_lo + b_len == rows
    assert key_windows(0, 5, 1, 10, 100, 1100) is None  # the layers cover the whole grid
    # message layout: sections 16-byte aligned, sized by the capacities
    L = message_layout(3, 4, 1000, 5000)
    assert L["WF"] == 24 and L["WG"] == 11
    for k in ("mig_f", "mig_i", "gh_f", "gh_i", "bytes"):
        assert L[k] % 16 == 0
    assert L["bytes"] >= 64 + 1000 * (24 * 4 + 24) + 5000 * (11 * 4 + 16)


@pytest.mark.parametrize("dim,f64", [(3, False), (3, True), (2, False), (2, True)])
def test_slab_abi_layout_and_argument_errors(dim, f64):
    """The C side and the Python side agree on the exchange-message layout byte for byte
    (jdb200_slab_message_bytes / _kept_bytes vs slab.message_layout), and the slab entry points
    reject bad descriptors / NULL pointers without touching a device."""
    from jaxdem_b200.slab import message_layout
    lib = _lib.lib()
    d = _lib.SlabDesc()
    d.n, d.cap_mig, d.cap_ghost, d.dim = 1000, 300, 7000, dim
    d.dtype = _lib.JDB200_F64 if f64 else _lib.JDB200_F32
    d.n_layers, d.lo_layer, d.up_layer, d.search_range = 40, 10, 20, 1
    fb = 8 if f64 else 4
    assert lib.jdb200_slab_message_bytes(ctypes.byref(d)) == message_layout(dim, fb, 300, 7000)["bytes"]
    assert lib.jdb200_slab_kept_bytes(ctypes.byref(d)) == message_layout(dim, fb, 0, 600)["bytes"]
    assert lib.jdb200_slab_scratch_bytes(ctypes.byref(d)) >= 1000 + 4 * 8 * 4
    assert lib.jdb200_slab_holes_bytes(ctypes.byref(d)) >= 4 * 300 * 4
    rows = _lib.SlabRows()
    # NULL messages / buffers
    assert lib.jdb200_slab_pack(None, ctypes.byref(d), ctypes.byref(rows), None, None, None, None, None, None, 0) == -2
    cnt = (ctypes.c_int64 * 7)(0, 0, 0, 0, 0, 0, 0)
    assert lib.jdb200_slab_unpack(None, ctypes.byref(d), ctypes.byref(rows), cnt, None, None, None, None) == -2
    # bad descriptors
    d.lo_layer, d.up_layer = 20, 10
    assert lib.jdb200_slab_message_bytes(ctypes.byref(d)) == 0
    assert lib.jdb200_slab_pack(None, ctypes.byref(d), ctypes.byref(rows), None, None, None, None, None, None, 0) == -1
    d.lo_layer, d.up_layer, d.dim = 10, 20, 4
    assert lib.jdb200_slab_scratch_bytes(ctypes.byref(d)) == 0


def test_params_argument_errors_of_new_fields():
    """key_window_* validation (check_params) and the fused force+after entry point's
    configuration check run on the host, before any launch."""
    lib = _lib.lib()
    st = jd.utils.grid_state(n_per_axis=(6, 6, 6), spacing=1.0, radius=0.5, device="cpu")
    sy = jd.System.create(st.shape, collider_type="CellList", collider_kw=dict(state=st), domain_type="periodic",
                          domain_kw=dict(box_size=[6.0] * 3), device="cpu")
    p = _call.params_for(st, sy)
    assert lib.jdb200_workspace_bytes(ctypes.byref(p)) > 0
    p.key_window_lo[0], p.key_window_len[0] = 0, 100      # two windows need a first window of k * 4096 rows
    p.key_window_lo[1], p.key_window_len[1] = 4096, 50
    assert lib.jdb200_workspace_bytes(ctypes.byref(p)) == 0
    p.key_window_len[0] = 4096
    assert lib.jdb200_workspace_bytes(ctypes.byref(p)) > 0
    p.key_window_lo[1] = 100                              # overlapping / descending windows
    assert lib.jdb200_workspace_bytes(ctypes.byref(p)) == 0
    p.key_window_lo[1], p.key_window_len[0] = 4096, -1
    assert lib.jdb200_workspace_bytes(ctypes.byref(p)) == 0


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("law", ["spring", "hertz", "cundallstrack"])
def test_force_model_pair_contract_matches_oracle(dim, law):
    """ForceModel.force / energy (the reference's per-pair contract, forces/__init__.py:55-150) in torch
    (jaxdem_b200/pair_laws.py; pure host-side code, runs on CPU tensors) against the oracle's restatement of the
    same laws on random pairs — periodic minimum image, two materials, self pairs give zero."""
    import jaxdem_b200 as jd
    from helpers import MATS, make_inputs
    import oracle
    from oracle import forces as oforces
    inp = make_inputs(300, dim, seed=5, dtype=np.float64, phi=0.7, nmat=2)
    ost = oracle.create_state(inp["pos"], vel=inp["vel"], ang_vel=inp["ang_vel"], rad=inp["rad"], mass=inp["mass"],
                              mat_id=inp["mat_id"], dtype=np.float64)
    osy = oracle.create_system(ost, domain_type="periodic", domain_kw=dict(box_size=inp["box"]), force_model_type=law,
                               mat_table=oracle.make_material_table(MATS[:2], "harmonic"))
    st = jd.State.create(inp["pos"], vel=inp["vel"], ang_vel=inp["ang_vel"], rad=inp["rad"], mass=inp["mass"],
                         mat_id=inp["mat_id"], dtype=torch.float64, device="cpu")
    mt = jd.MaterialTable.from_materials([jd.Material.create("elasticfrict", **m) for m in MATS[:2]],
                                         matcher=jd.MaterialMatchmaker.create("harmonic"))
    sy = jd.System.create(st.shape, domain_type="periodic", domain_kw=dict(box_size=inp["box"]), force_model_type=law,
                          mat_table=mt, collider_type="naive", dtype=torch.float64, device="cpu")
    rng = np.random.default_rng(1)
    # pairs that really touch (nearest periodic neighbours) plus random ones and a few self pairs
    d = inp["pos"][:, None, :] - inp["pos"][None, :, :]
    d -= inp["box"] * np.round(d / inp["box"])
    near = np.argsort((d**2).sum(-1), axis=1)[:, 1]
    i = np.concatenate([np.arange(300), rng.integers(0, 300, 200), np.arange(5)])
    j = np.concatenate([near, rng.integers(0, 300, 200), np.arange(5)])
    ti, tj = torch.as_tensor(i), torch.as_tensor(j)
    f, t = sy.force_model.force(ti, tj, st.pos, st, sy)
    e = sy.force_model.energy(ti, tj, st.pos, st, sy)
    wf, wt = getattr(oforces, law + "_force")(i, j, ost.pos, ost, osy)
    we = getattr(oforces, law + "_energy")(i, j, ost.pos, ost, osy)
    assert np.abs(wf).max() > 0 and (law != "cundallstrack" or np.abs(wt).max() > 0)
    for got, want in ((f, wf), (t, wt), (e, we)):
        want = np.asarray(want)
        assert got.shape == want.shape
        assert np.abs(got.numpy() - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
    assert float(f[-5:].abs().max()) == 0.0 and float(e[-5:].abs().max()) == 0.0


def test_force_model_pair_reference_pin():
    """The reference's own pin of the per-pair call (tests/test_clump_pair_friction.py:167-186): two unit spheres
    across the periodic boundary of a box of 10, centres 0.3 and 9.2 on the diagonal plane."""
    import jaxdem_b200 as jd
    pos = np.array([[0.5, 0.5], [9.7, 9.7]])
    st = jd.State.create(pos, rad=np.array([1.0, 1.0]), dtype=torch.float64, device="cpu")
    sy = jd.System.create(st.shape, domain_type="periodic", domain_kw=dict(box_size=[10.0, 10.0]),
                          collider_type="naive", dtype=torch.float64, device="cpu")
    f, t = sy.force_model.force(0, 1, st.pos, st, sy)
    rij = np.array([0.8, 0.8])
    r = np.linalg.norm(rij)
    k = float(sy.mat_table.young_eff[0, 0])
    assert np.allclose(f.numpy(), k * (2.0 - r) * rij / r, rtol=1e-13)
    assert t.shape == (1,) and float(t.abs().max()) == 0.0


@pytest.mark.parametrize("path", __import__("helpers").golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_golden(path):
    """The oracle against its committed outputs (tests/golden/, see make_golden.py for what they are): partition,
    neighbour hashes and neighbour lists bit for bit, forces / torques / energy / a 3-step trajectory to rounding."""
    from helpers import build_oracle, golden_close, load_golden
    import oracle
    from oracle import colliders as ocol
    inp, kw, want, meta = load_golden(path)
    ost, osy = build_oracle(inp, **kw)
    perm, sh, nh, _, _ = ocol.get_spatial_partition(ost.pos, osy, osy.collider.cell_size, osy.collider.neighbor_mask,
                                                     ost.idtype)
    if osy.domain.periodic:
        nh = ocol.dedup_stencil_hashes(nh)
    dt = kw["dtype"]
    golden_close(perm, want["perm"], dt, "perm")
    golden_close(sh, want["sorted_hash"], dt, "sorted_hash")
    golden_close(nh, want["nbr_hash"], dt, "nbr_hash")
    ocol.celllist_compute_force(ost, osy)
    golden_close(ost.force, want["force0"], dt, "force0")
    golden_close(ost.torque, want["torque0"], dt, "torque0")
    golden_close(ocol.celllist_compute_potential_energy(ost, osy), want["energy0"], dt, "energy0")
    nl, ovf = ocol.celllist_create_neighbor_list(ost, osy, 1.1, 24)
    golden_close(nl, want["nlist"], dt, "nlist")
    assert bool(ovf) == bool(want["nlist_overflow"])
    oracle.step(ost, osy, meta["steps"])
    for f in ("pos_c", "vel", "force", "torque", "ang_vel"):
        golden_close(getattr(ost, f), want[f + "_after"], dt, f)
