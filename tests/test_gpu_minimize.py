"""GPU parity tests of the minimiser inner loop (reference jaxdem/minimizers/routines.py:151-383,
optimizers.py:127-340) against the numpy oracle: the FIRE carry after a fixed number of iterations, the
termination tests, and batched ensembles whose systems stop at different iterations."""

import numpy as np
import pytest
import torch

import oracle
from oracle import minimizers as omin
from helpers import assert_close, build_gpu, build_oracle, make_inputs

pytestmark = pytest.mark.gpu


def _build(inp, dtype, collider, **kw):
    ocol = dict(celllist="celllist", naive="naive", neighborlist="neighborlist")[collider.lower()]
    ckw = dict(cutoff=1.0, skin=0.1) if ocol == "neighborlist" else None
    ost, osy = build_oracle(inp, dtype=dtype, collider=ocol, collider_kw=ckw, **kw)
    gst, gsy = build_gpu(inp, dtype=dtype, collider=collider, collider_kw=ckw, **kw)
    return ost, osy, gst, gsy


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,collider,clumps,law", [(2, "CellList", False, "spring"), (3, "CellList", True, "hertz"),
                                                    (3, "naive", True, "spring"), (2, "NeighborList", False, "spring")])
def test_fire_carry_after_k_iterations(dtype, dim, collider, clumps, law):
    import jaxdem_b200 as jd
    inp = make_inputs(500, dim, seed=3, dtype=dtype, phi=0.7, poly=1.4, clumps=clumps, fixed_frac=0.05)
    kw = dict(law=law, dt=1e-2, gravity=None)
    ost, osy, gst, gsy = _build(inp, dtype, collider, **kw)
    K = 25
    steps_o, pe_o, hist = omin.minimize(ost, osy, omin.FireConfig(1e-2), max_steps=K, pe_tol=0.0, pe_diff_tol=0.0,
                                        force_tol=-1.0)
    gst, gsy, steps, pe, carry = jd.minimizers.minimize(gst, gsy, max_steps=K, pe_tol=0.0, pe_diff_tol=0.0,
                                                        force_tol=-1.0, poll=7, return_carry=True)
    assert int(steps) == steps_o == K
    f = 200.0  # 25 iterations of a dynamical system: rounding differences compound
    assert abs(float(pe) - float(pe_o)) <= f * (1e-5 if dtype == np.float32 else 1e-12) * abs(float(hist[0]) / 500)
    for name in ("pos_c", "force", "torque", "_pos_p_rot"):
        assert_close(getattr(gst, name), getattr(ost, name), dtype, name, factor=f,
                     scale=None if name != "torque" else float(np.abs(ost.force).max()))
    assert_close(gst.q.w, ost.q_w, dtype, "q_w", factor=f, scale=1.0)
    assert_close(gst.q.xyz, ost.q_xyz, dtype, "q_xyz", factor=f, scale=1.0)
    fixed = np.asarray(inp["fixed"])
    assert np.array_equal(gst.pos_c.cpu().numpy()[fixed], np.asarray(inp["pos"], dtype)[fixed])  # fixed rows never move


def test_fire_scalars_match_oracle_f64():
    # dt / alpha / N_good / N_bad of the optimiser state after every chunk (poll = 1: one iteration per call)
    import jaxdem_b200 as jd
    dtype = np.float64
    inp = make_inputs(300, 2, seed=11, dtype=dtype, phi=0.8)
    ost, osy, gst, gsy = _build(inp, dtype, "CellList", law="spring", dt=1e-2)
    cfg = omin.FireConfig(1e-2)
    for K in (1, 7, 30):
        o2, s2 = ost.copy(), osy
        fs = omin.FireState(cfg, o2)
        g2 = gst.clone()
        _, _, steps, pe, carry = jd.minimizers.minimize(g2, gsy, max_steps=K, pe_tol=0.0, pe_diff_tol=0.0, force_tol=-1.0,
                                                        return_carry=True)
        # replay the oracle loop by hand to read its optimiser state
        pe_o = omin.eval_step(o2, s2)
        for _ in range(K):
            mask = (~o2.fixed)[:, None]
            upd = omin.fire_update(cfg, fs, {"pos_c": -o2.force * mask, "rotvec": -o2.torque * mask})
            omin.apply_delta_params(o2, o2.pos_c + upd["pos_c"] * mask, upd["rotvec"] * mask)
            pe_o = omin.eval_step(o2, s2)
        assert int(carry.n_good) == fs.N_good and int(carry.n_bad) == fs.N_bad, K
        assert abs(float(carry.dt) - float(fs.dt)) <= 1e-15 and abs(float(carry.alpha) - float(fs.alpha)) <= 1e-15
        assert abs(float(carry.pe) - float(pe_o)) <= 1e-10 * abs(float(pe_o)) + 1e-300
        assert_close(carry.vel_pos, fs.vel["pos_c"], dtype, "vel", factor=1e3)


@pytest.mark.parametrize("collider", ["CellList", "naive"])
def test_minimize_converges_like_the_oracle(collider):
    # examples/jam_spheres.py recipe: bidisperse discs 0.5 / 0.7, phi = 0.4, spring k = 1, FIRE dt = 1e-2, f64:
    # an unjammed packing relaxes to pe / N <= 1e-16
    import jaxdem_b200 as jd
    dtype = np.float64
    rng = np.random.default_rng(0)
    N = 200
    rad = np.where(np.arange(N) < N // 2, 0.5, 0.7)
    L = (np.sum(np.pi * rad**2) / 0.4) ** 0.5
    pos = rng.uniform(0, L, (N, 2))
    ost = oracle.create_state(pos, rad=rad, mass=np.ones(N))
    mt = oracle.make_material_table([dict(young=1.0, poisson=0.5, density=1.0)], "harmonic")
    okw = dict(collider_kw=dict(state=ost)) if collider == "CellList" else {}
    osy = oracle.create_system(ost, dt=1e-2, collider_type=collider.lower(), domain_type="periodic",
                               domain_kw=dict(box_size=[L, L]), mat_table=mt, **okw)
    gst = jd.State.create(pos, rad=rad, mass=np.ones(N), dtype=torch.float64)
    gkw = dict(collider_kw=dict(state=gst)) if collider == "CellList" else {}
    gsy = jd.System.create(gst.shape, dt=1e-2, minimizer=jd.minimizers.fire, minimizer_kw=dict(dt=1e-2),
                           collider_type=collider, domain_type="periodic", domain_kw=dict(box_size=[L, L]),
                           force_model_type="spring",
                           mat_table=jd.MaterialTable.from_materials(
                               [jd.Material.create("elastic", young=1.0, poisson=0.5, density=1.0)]),
                           dtype=torch.float64, **gkw)
    steps_o, pe_o, hist = omin.minimize(ost, osy, omin.FireConfig(1e-2), max_steps=5000)
    gst, gsy, steps, pe = jd.System.minimize(gst, gsy, max_steps=5000)
    assert 10 < steps_o < 5000 and abs(int(steps) - steps_o) <= 2, (int(steps), steps_o)
    assert float(pe) <= 1e-16 and pe_o <= 1e-16
    # and a jammed one stops on the relative-change test instead, at the same energy
    ost2 = oracle.create_state(pos * 0.6, rad=rad, mass=np.ones(N))
    osy.domain.box_size = osy.domain.box_size * 0.6
    gst2 = jd.State.create(pos * 0.6, rad=rad, mass=np.ones(N), dtype=torch.float64)
    gsy.domain.box_size.mul_(0.6)
    gsy.collider.overflow.zero_()
    s_o, pe_o, _ = omin.minimize(ost2, osy, omin.FireConfig(1e-2), max_steps=400, pe_diff_tol=1e-9)
    _, _, s_g, pe_g = jd.System.minimize(gst2, gsy, max_steps=400, pe_diff_tol=1e-9)
    assert pe_o > 1e-6 and abs(float(pe_g) - pe_o) <= 1e-6 * pe_o and abs(int(s_g) - s_o) <= 3, (int(s_g), s_o, float(pe_g), pe_o)


def test_minimize_batched_ensemble_equals_singles():
    # config-4 shape in small: B independent 2D disc systems in one call; each system stops at its own
    # iteration and its carry is frozen from then on (batched while_loop semantics)
    import jaxdem_b200 as jd
    dtype = np.float64
    B, N = 4, 150
    rad = np.where(np.arange(N) < N // 2, 0.5, 0.7)
    L = (np.sum(np.pi * rad**2) / 0.45) ** 0.5
    mt = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1.0, poisson=0.5, density=1.0)])
    singles = []
    for b in range(B):
        pos = np.random.default_rng(b).uniform(0, L, (N, 2))
        st = jd.State.create(pos, rad=rad, mass=np.ones(N), dtype=torch.float64)
        sy = jd.System.create(st.shape, dt=1e-2, collider_type="CellList", collider_kw=dict(state=st),
                              domain_type="periodic", domain_kw=dict(box_size=[L, L]), mat_table=mt, dtype=torch.float64)
        singles.append((st, sy))
    stb = jd.State.stack([s.clone() for s, _ in singles])
    syb = jd.System.create(stb.shape, dt=1e-2, collider_type="CellList", collider_kw=dict(state=singles[0][0]),
                           domain_type="periodic", domain_kw=dict(box_size=[[L, L]] * B), mat_table=mt, dtype=torch.float64)
    stb, syb, steps_b, pe_b = jd.System.minimize(stb, syb, max_steps=3000, poll=50)
    counts = []
    for b, (st, sy) in enumerate(singles):
        st, sy, steps, pe = jd.System.minimize(st, sy, max_steps=3000, poll=13)
        counts.append(int(steps))
        assert int(steps_b[b]) == int(steps), (b, steps_b, steps)
        assert float(pe_b[b]) == float(pe)
        assert torch.equal(stb.pos_c[b], st.pos_c) and torch.equal(stb.force[b], st.force)
    assert len(set(counts)) > 1 and max(counts) < 3000, counts


def test_minimize_argument_errors():
    import jaxdem_b200 as jd
    inp = make_inputs(50, 2, seed=1, dtype=np.float64)
    st, sy = build_gpu(inp, dtype=np.float64)
    sy.minimizer = None
    with pytest.raises(ValueError, match="No minimizer configured"):
        jd.System.minimize(st, sy)
