"""numpy restatement of the verlet / euler / verletspiral / spiral integrators.

Oracle only.  jaxdem/integrators/__init__.py:21-31, velocity_verlet.py:57-95,
direct_euler.py:62-66, velocity_verlet_spiral.py:83-180, spiral.py:22-141.
"""

from __future__ import annotations

import numpy as np

from . import linalg as la


def free_mask(state):
    """integrators/__init__.py:21-31."""
    return (~state.fixed)[:, None]


# ---- linear ---------------------------------------------------------------


def verlet_before(state, system):
    """velocity_verlet.py:57-61."""
    dt = system.dt
    half = state.fdtype.type(0.5)
    state.vel = state.vel + state.force * (dt * half / state.mass)[:, None] * free_mask(state)
    state.pos_c = state.pos_c + dt * state.vel


def verlet_after(state, system):
    """velocity_verlet.py:92-95."""
    dt = system.dt
    half = state.fdtype.type(0.5)
    state.vel = state.vel + state.force * (dt * half / state.mass)[:, None] * free_mask(state)


def euler_after(state, system):
    """direct_euler.py:62-66."""
    dt = system.dt
    state.vel = state.vel + state.force * (dt / state.mass)[:, None] * free_mask(state)
    state.pos_c = state.pos_c + dt * state.vel


# ---- rotation ---------------------------------------------------------------


def omega_dot(w, torque, inertia, inv_inertia):
    """spiral.py:22-40 (diagonal inertia)."""
    if w.shape[-1] == 3:
        wx, wy, wz = w[..., 0:1], w[..., 1:2], w[..., 2:3]
        ix, iy, iz = inertia[..., 0:1], inertia[..., 1:2], inertia[..., 2:3]
        tx, ty, tz = torque[..., 0:1], torque[..., 1:2], torque[..., 2:3]
        cx = wy * (iz * wz) - wz * (iy * wy)
        cy = wz * (ix * wx) - wx * (iz * wz)
        cz = wx * (iy * wy) - wy * (ix * wx)
        return np.concatenate([tx - cx, ty - cy, tz - cz], axis=-1) * inv_inertia
    return torque * inv_inertia


def _to_body(state):
    if state.dim == 3:
        return (
            la.q_rotate_back(state.q_w, state.q_xyz, state.ang_vel),
            la.q_rotate_back(state.q_w, state.q_xyz, state.torque),
        )
    return state.ang_vel, state.torque


def _ssprk3(state, h, w, tau):
    c = state.fdtype.type
    inv_I = c(1.0) / state.inertia
    k1 = h * omega_dot(w, tau, state.inertia, inv_I)
    k2 = h * omega_dot(w + k1, tau, state.inertia, inv_I)
    k3 = h * omega_dot(w + c(0.25) * (k1 + k2), tau, state.inertia, inv_I)
    return w + free_mask(state) * (k1 + k2 + c(4.0) * k3) / c(6.0)


def _rotvec3(state, rv):
    if state.dim == 2:
        return np.array([0.0, 0.0, 1.0], dtype=state.fdtype) * rv
    return rv


def verletspiral_before(state, system):
    """velocity_verlet_spiral.py:83-116."""
    c = state.fdtype.type
    dt_2 = system.dt / c(2.0)
    w, tau = _to_body(state)
    w = _ssprk3(state, dt_2, w, tau)
    dq_w, dq_xyz = la.q_from_small_rotvec(_rotvec3(state, c(2.0) * dt_2 * w))
    qw, qxyz = la.q_mul(state.q_w, state.q_xyz, dq_w, dq_xyz)
    qw, qxyz = la.q_unit(qw, qxyz)
    state.set_q(qw, qxyz)  # refreshes _pos_p_rot (state.py:264-273)
    state.ang_vel = la.q_rotate(state.q_w, state.q_xyz, w) if state.dim == 3 else w


def verletspiral_after(state, system):
    """velocity_verlet_spiral.py:156-180."""
    c = state.fdtype.type
    dt_2 = system.dt / c(2.0)
    w, tau = _to_body(state)
    w = _ssprk3(state, dt_2, w, tau)
    state.ang_vel = la.q_rotate(state.q_w, state.q_xyz, w) if state.dim == 3 else w


def spiral_after(state, system):
    """spiral.py:104-141."""
    c = state.fdtype.type
    dt = system.dt
    dt_2 = dt / c(2)
    inv_I = c(1.0) / state.inertia
    w, tau = _to_body(state)
    w_dot = omega_dot(w, tau, state.inertia, inv_I)
    rv1 = _rotvec3(state, c(2.0) * dt_2 * w)
    rv2 = _rotvec3(state, c(2.0) * dt_2 * dt_2 * w_dot)
    a_w, a_xyz = la.q_from_small_rotvec(rv1)
    b_w, b_xyz = la.q_from_small_rotvec(rv2)
    dq_w, dq_xyz = la.q_mul(a_w, a_xyz, b_w, b_xyz)
    qw, qxyz = la.q_mul(state.q_w, state.q_xyz, dq_w, dq_xyz)
    qw, qxyz = la.q_unit(qw, qxyz)
    state.set_q(qw, qxyz)
    k1 = dt * w_dot
    k2 = dt * omega_dot(w + k1, tau, state.inertia, inv_I)
    k3 = dt * omega_dot(w + c(0.25) * (k1 + k2), tau, state.inertia, inv_I)
    w = w + free_mask(state) * (k1 + k2 + c(4.0) * k3) / c(6.0)
    state.ang_vel = la.q_rotate(state.q_w, state.q_xyz, w) if state.dim == 3 else w


def _noop(state, system):
    pass


LINEAR = {
    "verlet": (verlet_before, verlet_after),
    "euler": (_noop, euler_after),
    "": (_noop, _noop),
}
ROTATION = {
    "verletspiral": (verletspiral_before, verletspiral_after),
    "spiral": (_noop, spiral_after),
    "": (_noop, _noop),
}
