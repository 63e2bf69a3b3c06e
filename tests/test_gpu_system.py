"""GPU tests of the step DRIVERS (row a1 / f4 of SURVEY §8): CUDA-graph replay, trajectory rollout, and the
reference's long-horizon physics checks run through the CUDA path (energy-drift order, analytic free top,
the README configuration for its full 1000 steps)."""

import numpy as np
import pytest
import torch

import oracle
from helpers import assert_close, build_gpu, build_oracle, compare_states, make_inputs

pytestmark = pytest.mark.gpu

FIELDS = ("pos_c", "vel", "force", "torque", "ang_vel", "_pos_p_rot")


def _same(a, b):
    for f in FIELDS:
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert torch.equal(a.q.w, b.q.w) and torch.equal(a.q.xyz, b.q.xyz)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("cfg", [("periodic", "spring", "", "auto"), ("periodic", "spring", "", "sorted"),
                                 ("reflect", "cundallstrack", "verletspiral", "auto"),
                                 ("free", "hertz", "spiral", "auto")], ids=lambda c: "-".join(c))
def test_compile_step_replay_is_bitwise_the_stream_path(dtype, cfg):
    # System.compile_step == jax.jit(System.step): a captured graph of n steps replayed twice must leave
    # the very same bits as 2n steps launched on the stream (incl. the radix fallback inside capture).
    import jaxdem_b200 as jd
    domain, law, rot, mode = cfg
    inp = make_inputs(3000, 3, seed=3, dtype=dtype, phi=0.55, poly=1.4, fixed_frac=0.02)
    kw = dict(dtype=dtype, domain=domain, law=law, rot=rot, dt=1e-3, grid_mode=mode,
              gravity=None if domain == "periodic" else [0.0, 0.0, -1.0])
    a_st, a_sy = build_gpu(inp, **kw)
    b_st, b_sy = build_gpu(inp, **kw)
    step = jd.System.compile_step(a_st, a_sy, n=3)
    # capture does not execute: both twins still hold the initial state
    _same(a_st, b_st)
    step()
    step()
    jd.System.step(b_st, b_sy, n=6)
    torch.cuda.synchronize()
    _same(a_st, b_st)
    assert int(a_sy.step_count) == int(b_sy.step_count) == 6
    assert float(a_sy.time) == float(b_sy.time)
    assert bool(a_sy.collider.overflow) == bool(b_sy.collider.overflow)


def test_compile_step_sees_forces_added_between_replays():
    # external buffers are read from device memory at replay time (never promised away inside a capture)
    import jaxdem_b200 as jd
    inp = make_inputs(500, 3, seed=4, dtype=np.float64, phi=0.3)
    a_st, a_sy = build_gpu(inp, dtype=np.float64, rot="")
    b_st, b_sy = build_gpu(inp, dtype=np.float64, rot="")
    step = jd.System.compile_step(a_st, a_sy, n=1)
    push = torch.zeros_like(a_st.force)
    push[..., 0] = 5.0
    for _ in range(3):
        jd.ForceManager.add_force(a_st, a_sy, push)
        step()
        jd.ForceManager.add_force(b_st, b_sy, push)
        jd.System.step(b_st, b_sy, n=1)
    _same(a_st, b_st)
    assert float(a_sy.force_manager.external_force.abs().max()) == 0.0


def test_trajectory_rollout_frames():
    # reference tests/test_public_api.py:69-78 + system.py:101-120: frame f is the state AFTER its strides[f]
    # steps; a leading 0 stride records the initial state; n/stride and strides agree.
    import jaxdem_b200 as jd
    inp = make_inputs(800, 3, seed=9, dtype=np.float64, phi=0.5)
    kw = dict(dtype=np.float64, law="hertz", rot="verletspiral", dt=1e-3)
    st, sy = build_gpu(inp, **kw)
    pos0 = st.pos.clone()
    st, sy, traj = jd.System.trajectory_rollout(st, sy, strides=[0, 1, 2, 3])
    assert traj.pos.shape == (4, 800, 3) and traj["vel"].shape == (4, 800, 3)
    assert torch.equal(traj.pos[0], pos0)
    assert traj.step_count.tolist() == [0, 1, 3, 6]
    # the same frames from plain stepping
    ref, rsy = build_gpu(inp, **kw)
    for f, k in enumerate([0, 1, 2, 3]):
        jd.System.step(ref, rsy, n=k)
        for name in ("pos_c", "vel", "force", "ang_vel"):
            assert torch.equal(traj[name][f], getattr(ref, name)), (f, name)
        assert torch.equal(traj.q_w[f], ref.q.w) and torch.equal(traj.q_xyz[f], ref.q.xyz)
        assert float(traj.time[f]) == float(rsy.time)
    _same(st, ref)
    # n + stride form, streamed to pinned host memory through the device ring (2 slots, 5 frames: slots reused)
    st2, sy2 = build_gpu(inp, **kw)
    st3, sy3 = build_gpu(inp, **kw)
    _, _, dev = jd.System.trajectory_rollout(st2, sy2, n=5, stride=2)
    _, _, host = jd.System.trajectory_rollout(st3, sy3, n=5, stride=2, to_host=True, ring=2)
    assert dev.pos.shape[0] == 5 and not host.pos.is_cuda and host.pos.is_pinned()
    for name in dev.keys():
        assert torch.equal(dev[name].cpu(), host[name]), name
    assert dev.step_count.tolist() == [2, 4, 6, 8, 10]
    # user save_fn: results stacked like lax.scan does
    st4, sy4 = build_gpu(inp, **kw)
    _, _, (e, c) = jd.System.trajectory_rollout(st4, sy4, n=3, stride=1,
                                                save_fn=lambda s, y: (s.vel.square().sum(), y.step_count.clone()))
    assert e.shape == (3,) and c.tolist() == [1, 2, 3]


def test_trajectory_rollout_batched_fields_subset():
    import jaxdem_b200 as jd
    B = 3
    sts = []
    for b in range(B):
        inp = make_inputs(300, 2, seed=20 + b, dtype=np.float32, phi=0.5)
        st, _ = build_gpu(inp, dtype=np.float32)
        sts.append((st, inp))
    stb = jd.State.stack([s.clone() for s, _ in sts])
    syb = jd.System.create(stb.shape, dt=1e-3, collider_type="CellList", collider_kw=dict(state=sts[0][0]),
                           domain_type="periodic", domain_kw=dict(box_size=np.stack([i["box"] for _, i in sts])),
                           rotation_integrator_type="", dtype=torch.float32)
    _, _, traj = jd.System.trajectory_rollout(stb, syb, n=4, stride=3, fields=("pos", "force"))
    assert set(traj.keys()) == {"pos", "force", "time", "step_count"}
    assert traj.pos.shape == (4, B, 300, 2) and traj.step_count.shape == (4, B)
    assert torch.equal(traj.pos[-1], stb.pos) and torch.equal(traj.force[-1], stb.force)
    with pytest.raises(KeyError):
        jd.System.trajectory_rollout(stb, syb, n=1, fields=("nope",))


def _gpu_spheres(dim, n, collider):
    """fixture recipe of reference tests/test_colliders_invariance.py:50-110 (numpy RNG instead of threefry)."""
    import jaxdem_b200 as jd
    spacing = 1.2
    ost = oracle.grid_state((n,) * dim, spacing, seed=0, dtype=np.float64, radius=0.5)
    box, anchor = np.full(dim, spacing * n), np.full(dim, -0.5)
    st = jd.State.create(ost.pos_c, vel=ost.vel, rad=ost.rad, mass=ost.mass, dtype=torch.float64)
    mt = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1.0e3, poisson=0.3, density=1.0)])

    def make(dt, lin, rot):
        s = st.clone()
        return s, jd.System.create(s.shape, dt=dt, linear_integrator_type=lin, rotation_integrator_type=rot,
                                   collider_type=collider, collider_kw=dict(state=s) if collider == "CellList" else {},
                                   domain_type="periodic", domain_kw=dict(box_size=box, anchor=anchor), mat_table=mt,
                                   dtype=torch.float64)
    return make


@pytest.mark.parametrize("lin,rot,thresh", [("verlet", "verletspiral", 1.8), ("euler", "spiral", 0.7)])
@pytest.mark.parametrize("collider,dim", [("naive", 2), ("CellList", 3)])
def test_energy_drift_order_through_cuda(lin, rot, thresh, collider, dim):
    # reference tests/test_energy_conservation.py:34-88: std of the total energy over a rollout scales like
    # dt^2 for velocity Verlet (slope > 1.8) and like dt for Euler (slope > 0.7); 5 log-spaced dts as there,
    # total time shortened to 1.0, frames through System.trajectory_rollout, energy through thermal.compute_energy
    import jaxdem_b200 as jd
    from jaxdem_b200 import thermal
    make = _gpu_spheres(dim, 6 if dim == 2 else 5, collider)
    dts = np.logspace(np.log10(2e-3), np.log10(1.25e-4), 5)
    stds = []
    for dt in dts:
        st, sy = make(float(dt), lin, rot)
        steps = int(1.0 / dt)
        stride = max(steps // 50, 1)
        es = []

        def energy(s, y):
            return thermal.compute_energy(s, y)

        st, sy, e = jd.System.trajectory_rollout(st, sy, n=50, stride=stride, save_fn=energy)
        stds.append(float(e.std(unbiased=False)))
    slope = np.polyfit(np.log(dts), np.log(stds), 1)[0]
    assert slope > thresh, (slope, stds)


@pytest.mark.parametrize("rot", ["spiral", "verletspiral"])
def test_aspherical_free_top_through_cuda(rot):
    # reference tests/test_rotation_integrators.py:24-97, same constants, 70 000 steps, f64: body-frame torque
    # tx about the symmetry axis; log10 relative error of the body-frame angular velocity < -4
    import jaxdem_b200 as jd
    rho, Rad, H = 7750.0, 0.05, 0.15
    Mass = rho * np.pi * Rad * Rad * H
    Ix = 0.5 * Mass * Rad * Rad
    Iy = Mass * H * H / 12.0 + 0.25 * Mass * Rad * Rad
    Iz = Iy
    tx, wx0, wy0, wz0 = Rad * 0.5, 0.3, -0.9, 0.6
    st = jd.State.create([[0.0, 0.0, 0.0]], ang_vel=[[wx0, wy0, wz0]], inertia=[[Ix, Iy, Iz]], mass=[Mass],
                         rad=[Rad], dtype=torch.float64)
    dt, steps = 1.0e-5, 70000
    sy = jd.System.create(st.shape, dt=dt, rotation_integrator_type=rot, collider_type="", dtype=torch.float64)
    tb = torch.tensor([[tx, 0.0, 0.0]], dtype=torch.float64, device=st.device)
    lin, rt, fm = sy.linear_integrator, sy.rotation_integrator, sy.force_manager
    for _ in range(steps):  # _step_once with the torque function evaluated where ForceManager.apply evaluates it
        lin.step_before_force(st, sy)
        rt.step_before_force(st, sy)
        sy.collider.compute_force(st, sy)
        fm.external_torque.copy_(jd.Quaternion.rotate(st.q, tb))
        fm.apply(st, sy)
        lin.step_after_force(st, sy)
        rt.step_after_force(st, sy)
    t = dt * steps
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
    w_body = jd.Quaternion.rotate_back(st.q, st.ang_vel)[0].cpu().numpy()
    err = np.log10(np.linalg.norm(w_body - ww) / np.linalg.norm(ww))
    assert err < -4.0, err


def test_readme_config_full_1000_steps_f64():
    # BASELINE config 1 exactly as the README runs it (10x10x10 grid, reflect box 20, naive collider, System.create
    # defaults, 1000 steps), f64.  The gas is chaotic: rounding-level differences grow ~10x per collision
    # generation (measured: 3e-4 in position after 1000 steps between two f64 implementations that agree to
    # 1e-12 per step), so the trajectory is compared with the numpy oracle for the first 200 steps (1 time unit,
    # 1-3 collisions per particle; bound 1e-9 of the field scale, element-wise 1e-8 relative on positions), and
    # the remaining 800 steps are checked through what stays meaningful: total energy is conserved by the elastic
    # walls + spring contacts to the integrator's accuracy, every sphere stays inside the box, and the clock.
    import jaxdem_b200 as jd
    from jaxdem_b200 import thermal
    ost = oracle.grid_state((10, 10, 10), 0.5, 0.1, seed=0, dtype=np.float64)
    pos0 = ost.pos_c.copy()
    osy = oracle.create_system(ost, domain_type="reflect", domain_kw=dict(box_size=[20.0] * 3))
    gst = jd.State.create(ost.pos_c, vel=ost.vel, rad=ost.rad, mass=ost.mass, dtype=torch.float64)
    gsy = jd.System.create(gst.shape, domain_type="reflect", domain_kw=dict(box_size=[20.0] * 3), dtype=torch.float64)
    e0 = float(thermal.compute_energy(gst, gsy))
    oracle.step(ost, osy, 200)
    jd.System.step(gst, gsy, n=200)
    compare_states(gst, ost, np.float64, factor=1.0e3)
    rel = np.abs(gst.pos_c.cpu().numpy() - ost.pos_c) / np.maximum(np.abs(ost.pos_c), 1e-3)
    assert rel.max() < 1e-8, rel.max()
    assert abs(float(thermal.compute_energy(gst, gsy)) - oracle.system.total_energy(ost, osy)) < 1e-9 * abs(e0)
    jd.System.step(gst, gsy, n=800)
    assert int(gsy.step_count) == 1000 and abs(float(gsy.time) - 5.0) < 1e-9
    e1 = float(thermal.compute_energy(gst, gsy))
    assert abs(e1 - e0) < 1e-2 * abs(e0), (e0, e1)  # measured 2.4e-3: dt = 0.005 against k = 1e4 contacts
    p = gst.pos_c.cpu().numpy()
    assert np.abs(p - pos0).max() > 1.0  # the gas really evolved
    # the drift follows domain.apply inside a step (system.py:66-74), so at the END of a step a sphere may stand
    # in a wall by at most one drift, |v| dt; the next apply mirrors it back
    slack = float(gst.vel.abs().max()) * 0.005 + 1e-9
    assert (p - 0.1 >= -slack).all() and (p + 0.1 <= 20.0 + slack).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_one_step_elementwise_relative_error(dtype):
    # next to the field-scale bound, an ELEMENT-WISE relative bound on the fields that are
    # not cancelling sums.  pos_c and vel after one step: |err| <= 4 tol |want| (+ 4 tol x 1e-3 x scale for the
    # components that are themselves ~0).
    import jaxdem_b200 as jd
    inp = make_inputs(20000, 3, seed=23, dtype=dtype, phi=0.5)
    kw = dict(dtype=dtype, law="cundallstrack", rot="verletspiral", dt=1e-3)
    ost, osy = build_oracle(inp, **kw)
    gst, gsy = build_gpu(inp, **kw)
    oracle.step(ost, osy, 1)
    jd.System.step(gst, gsy, n=1)
    t = 4.0 * (1e-5 if dtype == np.float32 else 1e-12)
    for f in ("pos_c", "vel", "ang_vel"):
        want = getattr(ost, f).astype(np.float64)
        got = getattr(gst, f).cpu().numpy().astype(np.float64)
        floor = 1e-3 * np.abs(want).max()
        if f == "ang_vel":  # the kick dt * torque / I of a cancelling contact-torque sum: its scale, not the element's
            floor += 1e-3 * np.abs(ost.torque).max() / float(ost.inertia.min())
        bad = np.abs(got - want) > t * (np.abs(want) + floor)
        assert not bad.any(), (f, int(bad.sum()), float(np.abs(got - want).max()))
