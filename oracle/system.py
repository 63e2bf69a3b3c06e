"""numpy restatement of System / _step_once.  Oracle only.

jaxdem/system.py:60-98 (step order), :230-270,474-477 (defaults).
"""

from __future__ import annotations

import numpy as np

from . import colliders, domains, force_manager, integrators
from .materials import make_material_table


class OSystem:
    pass


def create_system(state, *, dt=0.005, linear_integrator_type="verlet",
                  rotation_integrator_type="verletspiral", collider_type="naive", collider_kw=None,
                  domain_type="free", domain_kw=None, force_model_type="spring", mat_table=None,
                  gravity=None, interact_same_bond_id=False) -> OSystem:
    """Defaults of System.create (system.py:230-270,474-477)."""
    F = state.fdtype
    s = OSystem()
    s.dt = F.type(dt)
    s.time = F.type(0.0)
    s.step_count = 0
    s.linear_integrator = linear_integrator_type.lower()
    s.rotation_integrator = rotation_integrator_type.lower()
    s.collider_type = collider_type.lower()
    s.force_model = force_model_type.lower()
    s.interact_same_bond_id = bool(interact_same_bond_id)
    if mat_table is None:
        mat_table = make_material_table([dict(density=0.27, young=1.0e4, poisson=0.3)], "harmonic")
    s.mat_table = mat_table.astype(F)
    s.domain = domains.ODomain(domain_type, state.dim, dtype=F, **(domain_kw or {}))
    s.force_manager = force_manager.OForceManager(state.N, state.dim, gravity=gravity, dtype=F)
    if s.collider_type == "celllist":
        kw = dict(collider_kw or {})
        kw.pop("state", None)
        s.collider = colliders.celllist_create(state, **kw)
    elif s.collider_type == "multicelllist":
        kw = dict(collider_kw or {})
        kw.pop("state", None)
        s.collider = colliders.multicelllist_create(state, **kw)
    elif s.collider_type == "neighborlist":
        kw = dict(collider_kw or {})
        kw.pop("state", None)
        s.collider = colliders.neighborlist_create(state, **kw)
    else:
        s.collider = None
    return s


def step_once(state, system):
    """_step_once (system.py:60-82)."""
    system.time = system.time + system.dt
    system.step_count += 1
    domains.domain_apply(state, system)
    system.domain.inv_box_size = state.fdtype.type(1.0) / system.domain.box_size
    integrators.LINEAR[system.linear_integrator][0](state, system)
    integrators.ROTATION[system.rotation_integrator][0](state, system)
    colliders.compute_force(state, system)
    force_manager.apply(state, system)
    integrators.LINEAR[system.linear_integrator][1](state, system)
    integrators.ROTATION[system.rotation_integrator][1](state, system)


def step(state, system, n=1):
    """System.step (system.py:701-748), unbatched."""
    for _ in range(int(n)):
        step_once(state, system)
    return state, system


def kinetic_energy(state):
    """utils/thermal.py:24-78: translational + rotational KE, clump members
    weighted by 1/count (each member stores the clump totals)."""
    from . import linalg as la
    c = state.fdtype.type
    count = np.bincount(state.clump_id, minlength=state.N)[state.clump_id]
    ke_t = c(0.5) * state.mass * la.norm2(state.vel)
    if state.dim == 3:
        w = la.q_rotate_back(state.q_w, state.q_xyz, state.ang_vel)
    else:
        w = state.ang_vel
    ke_r = c(0.5) * la.dot(w, state.inertia * w)
    return np.sum((ke_t + ke_r) / count)


def total_energy(state, system):
    """utils/thermal.py:125-177."""
    pe = force_manager.compute_potential_energy(state, system) + colliders.compute_potential_energy(state, system)
    return kinetic_energy(state) + pe
