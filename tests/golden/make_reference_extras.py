"""Regenerate tests/golden/extras/*.npz:  python tests/golden/make_reference_extras.py   (build container only)

REFERENCE OUTPUTS for the rows of SURVEY §8 beyond the cell-list step (see make_reference_golden.py for how
the unmodified reference runs here on the numpy stand-in for JAX, and what that pins): the FIRE minimiser
loop (f1), the Verlet NeighborList collider (f2), MultiCellList (f3), trajectory_rollout frames (f4), the naive
collider with clumps and bonds, ForceManager with gravity and external forces on clumps, a reflecting box
with restitution, and a vmap-batched step.  One .npz per feature: inputs (``in_*``), outputs, ``meta``.
Consumed by tests/test_reference_golden.py (CPU: oracle vs these; GPU: CUDA path vs these)."""
import os
import sys
import time

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "extras")

from helpers import make_inputs  # noqa: E402
from make_reference_golden import build_reference  # noqa: E402

F64 = np.float64  # dtype of the case being generated (main() switches it for the *_f32 variants)


def _np(x):
    return np.array(np.asarray(x))


def _state_out(st, prefix=""):
    d = {prefix + f: _np(getattr(st, f)) for f in ("pos_c", "vel", "force", "torque", "ang_vel")}
    d[prefix + "q"] = np.concatenate([_np(st.q.w), _np(st.q.xyz)], axis=-1)
    return d


def _inputs(inp):
    return {f"in_{k}": np.asarray(v) for k, v in inp.items() if not isinstance(v, list)}


def case_fire(jd, jax, jnp):
    """minimizers/routines.py:151-383 + optimizers.py:150-340: K iterations with every stop test disabled, and a
    run to convergence of the examples/jam_spheres.py recipe (bidisperse discs, phi 0.4, spring k = 1)."""
    out = {}
    for tag, dim, clumps, law in (("a", 2, False, "spring"), ("b", 3, True, "hertz")):
        inp = make_inputs(90, dim, seed=3, dtype=F64, phi=0.7, poly=1.4, clumps=clumps, fixed_frac=0.05,
                          nmat=2 if law != "spring" else 1)
        kw = dict(domain="periodic", law=law, lin="verlet", rot="verletspiral", dt=1e-2, nmat=2 if law != "spring" else 1)
        st, sy = build_reference(jd, jnp, inp, **kw)
        K = 10
        st, sy, steps, pe = jd.minimize(st, sy, max_steps=K, pe_tol=0.0, pe_diff_tol=0.0, force_tol=-1.0)
        out.update({f"{tag}_{k}": v for k, v in _inputs(inp).items()})
        out.update(_state_out(st, f"{tag}_"))
        out[f"{tag}_steps"], out[f"{tag}_pe"] = _np(steps), _np(pe)
        out[f"{tag}_pos_p_rot"] = _np(st._pos_p_rot)
    # convergence: unjammed packing relaxes to pe / N <= 1e-16
    rng = np.random.default_rng(0)
    N = 60
    rad = np.where(np.arange(N) < N // 2, 0.5, 0.7)
    L = (np.sum(np.pi * rad**2) / 0.4) ** 0.5
    pos = rng.uniform(0, L, (N, 2))
    st = jd.State.create(jnp.asarray(pos), rad=jnp.asarray(rad), mass=jnp.ones(N))
    mt = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1.0, poisson=0.5, density=1.0)])
    sy = jd.System.create(st.shape, dt=1e-2, collider_type="celllist", collider_kw=dict(state=st),
                          domain_type="periodic", domain_kw=dict(box_size=jnp.asarray([L, L])),
                          force_model_type="spring", mat_table=mt, minimizer_kw=dict(dt=1e-2))
    st, sy, steps, pe = jd.minimize(st, sy, max_steps=5000)
    out.update(c_in_pos=pos, c_in_rad=rad, c_in_box=np.array([L, L]), c_steps=_np(steps), c_pe=_np(pe),
               c_pos_c=_np(st.pos_c))
    return out, dict(K=10)


def case_naive(jd, jax, jnp):
    """colliders/naive.py with clumps, two materials and bonds (valid_interaction_mask, colliders/__init__.py)."""
    out = {}
    for tag, dim, law in (("a", 3, "cundallstrack"), ("b", 2, "hertz")):
        inp = make_inputs(70, dim, seed=5, dtype=F64, phi=0.6, poly=1.3, clumps=True, nmat=2)
        n = 70
        bond = np.full((n, 1), -1, dtype=np.int64)
        bond[::3, 0] = (np.arange(n)[::3] + 1) % n
        st0, _ = build_reference(jd, jnp, inp, domain="periodic", law=law, lin="verlet", rot="verletspiral", dt=1e-3,
                                 nmat=2, collider="naive")
        kw = {k: jnp.asarray(inp[k]) for k in ("vel", "ang_vel", "rad", "mass", "clump_id", "pos_p", "q", "inertia",
                                               "mat_id")}
        st = jd.State.create(jnp.asarray(inp["pos"]), bond_id=jnp.asarray(bond), **kw)
        _, sy = build_reference(jd, jnp, inp, domain="periodic", law=law, lin="verlet", rot="verletspiral", dt=1e-3,
                                nmat=2, collider="naive")
        st, sy = sy.collider.compute_force(st, sy)
        e = sy.collider.compute_potential_energy(st, sy)[2]
        out.update({f"{tag}_{k}": v for k, v in _inputs(inp).items()})
        out[f"{tag}_in_bond"] = bond
        out[f"{tag}_bond_id"] = _np(st.bond_id)  # symmetrised and padded by State.create
        out[f"{tag}_force"], out[f"{tag}_torque"], out[f"{tag}_energy"] = _np(st.force), _np(st.torque), _np(e)
    return out, {}


def case_nlist(jd, jax, jnp):
    """colliders/neighbor_list.py:286-480,542-669: list at build, forces, then steps with a thin skin so the
    list is rebuilt on the way (n_build_times)."""
    inp = make_inputs(150, 3, seed=31, dtype=F64, phi=0.6, poly=1.5, nmat=2)
    kw = dict(domain="periodic", law="hertz", lin="verlet", rot="verletspiral", dt=2e-3, nmat=2)
    ckw = dict(cutoff=1.0, skin=0.04)
    st, sy = build_reference(jd, jnp, inp, collider="neighborlist", collider_kw=ckw, **kw)
    out = _inputs(inp)
    out["max_neighbors"] = np.asarray(int(sy.collider.max_neighbors))
    st, sy = sy.collider.compute_force(st, sy)
    out.update(nl0=_np(sy.collider.neighbor_list), builds0=_np(sy.collider.n_build_times), force0=_np(st.force),
               torque0=_np(st.torque), overflow0=_np(sy.collider.overflow))
    out["energy0"] = _np(sy.collider.compute_potential_energy(st, sy)[2])
    builds = []
    for _ in range(6):
        st, sy = jd.System.step(st, sy, n=5)
        builds.append(int(sy.collider.n_build_times))
    out["builds"] = np.asarray(builds)
    out["nl_after"] = _np(sy.collider.neighbor_list)
    out.update(_state_out(st, "after_"))
    return out, dict(steps=30, **ckw)


def case_multicell(jd, jax, jnp):
    """colliders/multi_cell_list.py:46-73,142-254 on a strongly polydisperse packing."""
    out = {}
    for tag, dim, law, domain in (("a", 3, "spring", "periodic"), ("b", 2, "hertz", "reflect")):
        inp = make_inputs(130, dim, seed=17, dtype=F64, phi=0.55, poly=3.0, nmat=2)
        st, sy = build_reference(jd, jnp, inp, domain=domain, law=law, lin="verlet", rot="verletspiral", dt=1e-3,
                                 nmat=2, collider="multicelllist")
        st, sy = sy.collider.compute_force(st, sy)
        e = sy.collider.compute_potential_energy(st, sy)[2]
        out.update({f"{tag}_{k}": v for k, v in _inputs(inp).items()})
        out[f"{tag}_force"], out[f"{tag}_torque"], out[f"{tag}_energy"] = _np(st.force), _np(st.torque), _np(e)
        out[f"{tag}_cell_size"] = _np(sy.collider.cell_size)
        st, sy = jd.System.step(st, sy, n=2)
        out.update(_state_out(st, f"{tag}_after_"))
    return out, {}


def case_force_manager(jd, jax, jnp):
    """forces/force_manager.py: gravity + external force on particles, at the clump COM, external torque, then
    apply (the clump segment sums) inside two steps."""
    inp = make_inputs(90, 3, seed=9, dtype=F64, phi=0.5, poly=1.3, clumps=True)
    g = np.array([0.0, -2.0, -9.81], dtype=F64)
    st, sy = build_reference(jd, jnp, inp, domain="periodic", law="spring", lin="verlet", rot="verletspiral",
                             dt=1e-3, nmat=1, gravity=jnp.asarray(g))
    rng = np.random.default_rng(4)
    fe, fc, te = (rng.normal(size=(90, 3)).astype(F64) for _ in range(3))
    sy = sy.force_manager.add_force(st, sy, jnp.asarray(fe))
    sy = sy.force_manager.add_force(st, sy, jnp.asarray(fc), is_com=True)
    sy = sy.force_manager.add_torque(st, sy, jnp.asarray(te))
    out = _inputs(inp)
    out.update(in_gravity=g, in_fe=fe, in_fc=fc, in_te=te)
    st, sy = jd.System.step(st, sy, n=1)
    out.update(_state_out(st, "s1_"))
    st, sy = jd.System.step(st, sy, n=1)  # external buffers were cleared by the first apply
    out.update(_state_out(st, "s2_"))
    return out, {}


def case_reflect(jd, jax, jnp):
    """domains/reflect.py with restitution 0.7: fast spheres and clumps crossing the walls over 12 steps."""
    out = {}
    for tag, clumps in (("a", False), ("b", True)):
        inp = make_inputs(80, 3, seed=21, dtype=F64, phi=0.35, poly=1.3, clumps=clumps)
        inp["vel"] = inp["vel"] * 40.0
        from helpers import MATS
        mats = [jd.Material.create("elasticfrict", **MATS[0])]
        mt = jd.MaterialTable.from_materials(mats, matcher=jd.MaterialMatchmaker.create("harmonic"))
        kw = {k: jnp.asarray(inp[k]) for k in ("vel", "ang_vel", "rad", "mass", "clump_id", "pos_p", "q", "inertia")
              if k in inp}
        st = jd.State.create(jnp.asarray(inp["pos"]), **kw)
        sy = jd.System.create(st.shape, dt=2e-3, collider_type="celllist", collider_kw=dict(state=st),
                              domain_type="reflect",
                              domain_kw=dict(box_size=jnp.asarray(inp["box"]), restitution_coefficient=0.7),
                              force_model_type="spring", mat_table=mt)
        st, sy = jd.System.step(st, sy, n=12)
        out.update({f"{tag}_{k}": v for k, v in _inputs(inp).items()})
        out.update(_state_out(st, f"{tag}_"))
    return out, dict(steps=12, dt=2e-3, restitution=0.7)


def case_rollout(jd, jax, jnp):
    """system.py:606-699 trajectory_rollout(n=3, stride=2): frames after 2, 4, 6 steps; final state = last frame."""
    inp = make_inputs(80, 2, seed=2, dtype=F64, phi=0.6, poly=1.2)
    st, sy = build_reference(jd, jnp, inp, domain="periodic", law="spring", lin="verlet", rot="verletspiral",
                             dt=1e-3, nmat=1)
    st, sy, (fst, fsy) = jd.System.trajectory_rollout(st, sy, n=3, stride=2)
    out = _inputs(inp)
    out.update(frames_pos_c=_np(fst.pos_c), frames_vel=_np(fst.vel), frames_time=_np(fsy.time),
               frames_step_count=_np(fsy.step_count), final_pos_c=_np(st.pos_c), final_time=_np(sy.time))
    return out, dict(n=3, stride=2)


def case_batched(jd, jax, jnp):
    """vmap(System.step) over B = 3 independent systems (system.py:701-748 under jax.vmap)."""
    B = 3
    inps = [make_inputs(70, 2, seed=40 + b, dtype=F64, phi=0.6, poly=1.3, box=9.0) for b in range(B)]
    pairs = [build_reference(jd, jnp, i, domain="periodic", law="spring", lin="verlet", rot="verletspiral", dt=1e-3,
                             nmat=1) for i in inps]
    stb = jd.State.stack([p[0] for p in pairs])
    syb = jax.tree.map(lambda *xs: jnp.stack(xs), *[p[1] for p in pairs])
    stb, syb = jax.vmap(lambda s, y: jd.System.step(s, y, n=2))(stb, syb)
    out = {}
    for b, i in enumerate(inps):
        out.update({f"b{b}_{k}": v for k, v in _inputs(i).items()})
    out.update(_state_out(stb))
    return out, dict(B=B, steps=2)


def case_cross(jd, jax, jnp):
    """colliders/cell_list.py:598-700 create_cross_neighbor_list: query points (some un-wrapped, some outside a
    reflecting box's grid) against the state's positions; a roomy and a too-small max_neighbors (overflow flag)."""
    out = {}
    for tag, dim, domain in (("a", 3, "periodic"), ("b", 2, "reflect")):
        inp = make_inputs(160, dim, seed=6, dtype=F64, phi=0.55, poly=1.3)
        st, sy = build_reference(jd, jnp, inp, domain=domain, law="spring", lin="verlet", rot="verletspiral",
                                 dt=1e-3, nmat=1)
        rng = np.random.default_rng(12)
        pos_a = (rng.uniform(0.0, 0.999, (90, dim)) * inp["box"]).astype(F64)
        if domain == "periodic":
            pos_a[::5] += inp["box"].astype(F64)
        out.update({f"{tag}_{k}": v for k, v in _inputs(inp).items()})
        out[f"{tag}_in_pos_a"] = pos_a
        for K in (40, 4):
            nl, ovf = sy.collider.create_cross_neighbor_list(jnp.asarray(pos_a), st.pos, sy, 1.2, K)
            out[f"{tag}_nl{K}"], out[f"{tag}_ovf{K}"] = _np(nl), _np(ovf)
    return out, dict(cutoff=1.2)


STATE_FIELDS = ("pos_c", "pos_p", "vel", "force", "ang_vel", "torque", "rad", "volume", "mass", "inertia", "clump_id",
                "bond_id", "mat_id", "species_id", "fixed", "_pos_p_rot")


def case_state_create(jd, jax, jnp):
    """state.py:371-867 State.create defaults: a bare position array in 2-D and 3-D (unit radii / masses, solid
    disc / sphere inertia, identity quaternions), masses from a material table's densities, and a clump system
    with sparse clump ids (relabelled to dense ids), ragged bonds (symmetrised, padded) and given quaternions
    (the ``_pos_p_rot`` cache)."""
    from helpers import MATS
    rng = np.random.default_rng(8)
    out = {}

    def put(tag, st):
        for f in STATE_FIELDS:
            out[f"{tag}_{f}"] = _np(getattr(st, f))
        out[f"{tag}_q"] = np.concatenate([_np(st.q.w), _np(st.q.xyz)], axis=-1)

    for tag, dim in (("a", 2), ("b", 3)):
        pos = rng.uniform(0, 5, (7, dim))
        out[f"{tag}_in_pos"] = pos
        put(tag, jd.State.create(jnp.asarray(pos)))
    pos, rad, mat_id = rng.uniform(0, 5, (9, 3)), rng.uniform(0.2, 0.6, 9), rng.integers(0, 3, 9)
    mt = jd.MaterialTable.from_materials([jd.Material.create("elasticfrict", **m) for m in MATS],
                                         matcher=jd.MaterialMatchmaker.create("harmonic"))
    out.update(c_in_pos=pos, c_in_rad=rad, c_in_mat_id=mat_id)
    put("c", jd.State.create(jnp.asarray(pos), rad=jnp.asarray(rad), mat_id=jnp.asarray(mat_id), mat_table=mt))
    n = 10
    pos, pos_p = rng.uniform(0, 5, (n, 3)), rng.normal(0, 0.3, (n, 3))
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    cid = np.array([40, 40, 7, 7, 7, 3, 99, 99, 12, 5])
    bonds = np.full((n, 2), -1, dtype=np.int64)
    bonds[0] = [3, 8]
    bonds[4, 0] = 9
    bonds[6, 0] = 0
    out.update(d_in_pos=pos, d_in_pos_p=pos_p, d_in_q=q, d_in_clump_id=cid, d_in_bond=bonds)
    put("d", jd.State.create(jnp.asarray(pos), pos_p=jnp.asarray(pos_p), q=jnp.asarray(q), clump_id=jnp.asarray(cid),
                             bond_id=jnp.asarray(bonds)))
    return out, {}


def case_collider_create(jd, jax, jnp):
    """colliders/cell_list.py:375-433 CellList.Create (cell size and stencil from the radius spread, with and
    without a box), multi_cell_list.py:324-370, neighbor_list.py:287-402 (skin, max_neighbors sizing)."""
    rng = np.random.default_rng(0)
    out = {}
    for i, (poly, box, dim) in enumerate(((1.0, None, 3), (3.0, None, 3), (1.2, [2.5, 2.5, 2.5], 3), (6.0, None, 2))):
        rad = rng.uniform(0.5 / poly, 0.5, 50)
        pos = rng.uniform(0, 5, (50, dim))
        st = jd.State.create(jnp.asarray(pos), rad=jnp.asarray(rad))
        kw = {} if box is None else dict(box_size=jnp.asarray(box))
        c = jd.Collider.create("celllist", state=st, **kw)
        m = jd.Collider.create("multicelllist", state=st, **kw)
        out.update({f"k{i}_in_pos": pos, f"k{i}_in_rad": rad, f"k{i}_in_box": np.asarray([] if box is None else box),
                    f"k{i}_cell_size": _np(c.cell_size), f"k{i}_neighbor_mask": _np(c.neighbor_mask),
                    f"k{i}_multi_cell_size": _np(m.cell_size), f"k{i}_multi_neighbor_mask": _np(m.neighbor_mask)})
        for j, nkw in enumerate((dict(cutoff=1.0), dict(cutoff=1.0, skin=0.3), dict(cutoff=1.5, skin_fraction=0.2),
                                 dict(cutoff=1.0, skin=0.1, max_neighbors=7),
                                 dict(cutoff=1.2, skin=0.1, number_density=0.6, safety_factor=1.5))):
            nl = jd.Collider.create("neighborlist", state=st, **nkw)
            out[f"k{i}_nl{j}"] = np.asarray([float(np.asarray(nl.cutoff)), float(np.asarray(nl.skin)),
                                             float(int(nl.max_neighbors)),
                                             float(np.asarray(nl.secondary_collider.cell_size))])
    return out, {}


def case_pair_laws(jd, jax, jnp):
    """forces/{spring,hertz,cundall_strack}.py: the per-pair plugin contract ``force(i, j, pos, state, system)`` /
    ``energy(...)`` on touching pairs (nearest periodic neighbours), random pairs and self pairs, two materials."""
    out = {}
    for dim in (2, 3):
        inp = make_inputs(120, dim, seed=5, dtype=F64, phi=0.7, nmat=2)
        rng = np.random.default_rng(1)
        d = inp["pos"][:, None, :] - inp["pos"][None, :, :]
        d -= inp["box"] * np.round(d / inp["box"])
        near = np.argsort((d**2).sum(-1), axis=1)[:, 1]
        i = np.concatenate([np.arange(120), rng.integers(0, 120, 80), np.arange(5)])
        j = np.concatenate([near, rng.integers(0, 120, 80), np.arange(5)])
        out.update({f"d{dim}_{k}": v for k, v in _inputs(inp).items()})
        out[f"d{dim}_i"], out[f"d{dim}_j"] = i, j
        for law in ("spring", "hertz", "cundallstrack"):
            st, sy = build_reference(jd, jnp, inp, domain="periodic", law=law, lin="verlet", rot="verletspiral",
                                     dt=1e-3, nmat=2, collider="naive")
            fm = sy.force_model
            f, t = jax.vmap(lambda a, b: fm.force(a, b, st.pos, st, sy))(jnp.asarray(i), jnp.asarray(j))
            e = jax.vmap(lambda a, b: fm.energy(a, b, st.pos, st, sy))(jnp.asarray(i), jnp.asarray(j))
            out[f"d{dim}_{law}_f"], out[f"d{dim}_{law}_t"], out[f"d{dim}_{law}_e"] = _np(f), _np(t), _np(e)
    return out, {}


def case_energy(jd, jax, jnp):
    """utils/thermal.py:83-177: translational / rotational kinetic energy, potential energy (collider + gravity) and
    total energy of a clump system under gravity — the quantities the energy-drift tests integrate."""
    from jaxdem.utils import thermal
    out = {}
    for tag, dim in (("a", 3), ("b", 2)):
        inp = make_inputs(90, dim, seed=13, dtype=F64, phi=0.55, poly=1.3, clumps=True)
        g = np.array([0.0, -9.81, 0.0][:dim] if dim == 3 else [0.0, -9.81], dtype=F64)
        st, sy = build_reference(jd, jnp, inp, domain="periodic", law="hertz", lin="verlet", rot="verletspiral",
                                 dt=1e-3, nmat=1, gravity=jnp.asarray(g))
        out.update({f"{tag}_{k}": v for k, v in _inputs(inp).items()})
        out[f"{tag}_in_gravity"] = g
        out[f"{tag}_ke_t"] = _np(thermal.compute_translational_kinetic_energy(st))
        out[f"{tag}_ke_r"] = _np(thermal.compute_rotational_kinetic_energy(st))
        out[f"{tag}_pe"] = _np(thermal.compute_potential_energy(st, sy))
        out[f"{tag}_e"] = _np(thermal.compute_energy(st, sy))
    return out, {}


def case_errors(jd, jax, jnp):
    """Error behaviour of the factories / builders on the path (factory.py:240-320, system.py:474-540,
    neighbor_list.py:330-340, system.py:684-690): exception type and message of seven invalid calls."""
    st = jd.State.create(jnp.asarray(np.random.default_rng(0).uniform(0, 3, (6, 2))))
    mt_el = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1.0, poisson=0.3, density=1.0)])
    sy = jd.System.create(st.shape)
    calls = dict(
        missing_material_fields=lambda: jd.System.create(st.shape, force_model_type="cundallstrack", mat_table=mt_el),
        skin_and_fraction=lambda: jd.Collider.create("neighborlist", state=st, cutoff=1.0, skin=0.1, skin_fraction=0.1),
        unknown_collider=lambda: jd.Collider.create("nosuchcollider"),
        rollout_without_n=lambda: jd.System.trajectory_rollout(st, sy),
        unknown_material=lambda: jd.Material.create("nosuchmaterial"),
        unknown_force_model=lambda: jd.ForceModel.create("nosuchlaw"),
        unknown_domain=lambda: jd.Domain.create("nosuchdomain", dim=2),
    )
    out = {}
    for name, f in calls.items():
        try:
            f()
            out[name] = np.array(["", ""])
        except Exception as e:  # noqa: BLE001 - the point is to record whatever the reference raises
            out[name] = np.array([type(e).__name__, str(e)])
    return out, {}


def case_materials(jd, jax, jnp):
    """materials/material_table.py:87-140 + material_matchmakers/{harmonic,linear}.py: per-material arrays and the
    effective pair tables of three elastic-friction materials under both matchmakers."""
    from helpers import MATS
    out = {}
    for matcher in ("harmonic", "linear"):
        mats = [jd.Material.create("elasticfrict", **m) for m in MATS]
        mt = jd.MaterialTable.from_materials(mats, matcher=jd.MaterialMatchmaker.create(matcher))
        for k in ("young", "poisson", "density", "mu", "e", "mu_r"):
            out[f"{matcher}_{k}"] = _np(getattr(mt, k))
            out[f"{matcher}_{k}_eff"] = _np(getattr(mt, k + "_eff"))
    return out, {}


CASES = dict(errors=case_errors, energy=case_energy, pair_laws=case_pair_laws, collider_create=case_collider_create, state_create=case_state_create, materials=case_materials, cross=case_cross, cross_f32=case_cross, fire=case_fire, naive=case_naive, nlist=case_nlist, multicell=case_multicell,
             force_manager=case_force_manager, reflect=case_reflect, rollout=case_rollout, batched=case_batched,
             nlist_f32=case_nlist, multicell_f32=case_multicell, force_manager_f32=case_force_manager,
             reflect_f32=case_reflect, batched_f32=case_batched)


def main():
    from _ref_import import import_reference
    jd = import_reference()
    import jax
    import jax.numpy as jnp
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in CASES.items():
        if only and name not in only:
            continue
        t = time.time()
        global F64
        F64 = np.float32 if name.endswith("_f32") else np.float64
        jax.config.update("jax_enable_x64", F64 == np.float64)  # float32 = JAX's default mode, no 64-bit arrays
        with np.errstate(all="ignore"):
            data, meta = fn(jd, jax, jnp)
        bad = {k: v.dtype for k, v in data.items() if not k.startswith(("in_", "a_in", "b_in", "c_in", "b0_", "b1_", "b2_"))
               and v.dtype.kind == "f" and v.dtype != np.dtype(F64)}
        assert not bad, bad
        meta = dict(meta, source="reference sources on tests/golden/jaxshim", dtype=np.dtype(F64).name)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **data, meta=np.array(repr(meta)))
        print(name, f"{time.time() - t:.1f} s", len(data), "arrays", flush=True)


if __name__ == "__main__":
    main()
