"""numpy restatement of the free / periodic / reflect domains.  Oracle only.

jaxdem/domains/__init__.py:26-215, periodic.py:31-116, reflect.py:99-299,
_toc.py:11-94, free.py:42-64.
"""

from __future__ import annotations

import numpy as np

from . import linalg as la


class ODomain:
    def __init__(self, kind: str, dim: int, box_size=None, anchor=None,
                 restitution_coefficient=1.0, dtype=np.float64):
        kind = kind.lower()
        assert kind in ("free", "periodic", "reflect")
        F = np.dtype(dtype)
        self.kind = kind
        self.box_size = np.ones(dim, F) if box_size is None else np.broadcast_to(
            np.asarray(box_size, dtype=F), (dim,)).copy()
        self.anchor = np.zeros(dim, F) if anchor is None else np.broadcast_to(
            np.asarray(anchor, dtype=F), (dim,)).copy()
        self.inv_box_size = (F.type(1.0) / self.box_size).astype(F)
        self.restitution_coefficient = F.type(restitution_coefficient)

    @property
    def periodic(self) -> bool:
        return self.kind == "periodic"

    # -- displacement ------------------------------------------------------
    def displacement(self, ri, rj):
        """Division form (periodic.py:70-71), used by neighbour lists."""
        rij = ri - rj
        if self.periodic:
            return rij - self.box_size * np.round(rij / self.box_size)  # np.round = half-to-even
        return rij

    def _displacement(self, ri, rj):
        """Multiply form (periodic.py:75-79), used by the force laws."""
        rij = ri - rj
        if self.periodic:
            return rij - self.box_size * np.round(rij * self.inv_box_size)
        return rij

    def copy(self):
        d = ODomain(self.kind, self.box_size.shape[0], dtype=self.box_size.dtype)
        d.box_size, d.anchor = self.box_size.copy(), self.anchor.copy()
        d.inv_box_size = self.inv_box_size.copy()
        d.restitution_coefficient = self.restitution_coefficient
        return d


def verlet_collision_fraction(v_end, acc, delta, wall_sign, dt):
    """jaxdem/domains/_toc.py:68-94."""
    c = lambda x: v_end.dtype.type(x)
    v_0 = v_end - dt * acc
    v_mid = v_end - c(0.5) * dt * acc
    v_0_n = v_0 * wall_sign
    v_mid_n = v_mid * wall_sign
    acc_n = acc * wall_sign
    A = c(0.5) * acc_n * dt * dt
    B = v_0_n * dt
    dw = delta + v_mid_n * dt
    disc = np.maximum(c(0.0), B * B + c(4.0) * A * dw)
    safe = np.where(disc > 0.0, disc, c(1.0))
    sq = np.where(disc > 0.0, np.sqrt(safe), c(0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = np.where(
            B < 0,
            c(2.0) * dw / np.where(B - sq < -1e-10, B - sq, c(-1.0)),
            (-B - sq) / np.where(np.abs(c(2.0) * A) > 1e-10, c(2.0) * A, c(1.0)),
        )
    return np.clip(alpha, c(0.0), c(1.0))


def _segment(op, data, ids, n):
    """jax.ops.segment_{sum,max,min} with num_segments=n (empty segments are
    never gathered back by the callers, so their fill value is irrelevant)."""
    if op == "sum":
        out = np.zeros((n,) + data.shape[1:], data.dtype)
        np.add.at(out, ids, data)
    elif op == "max":
        out = np.full((n,) + data.shape[1:], -np.inf, data.dtype)
        np.maximum.at(out, ids, data)
    else:
        out = np.full((n,) + data.shape[1:], np.inf, data.dtype)
        np.minimum.at(out, ids, data)
    return out


def free_apply(state, system):
    """jaxdem/domains/free.py:59-63."""
    br = (la.norm(state.pos_p) + state.rad)[:, None]
    p_min = np.min(state.pos_c - br, axis=0)
    p_max = np.max(state.pos_c + br, axis=0)
    system.domain.box_size = p_max - p_min
    system.domain.anchor = p_min


def reflect_apply(state, system):
    """jaxdem/domains/reflect.py:165-299 (clump-aware impulse + Verlet TOC)."""
    dom = system.domain
    F = state.fdtype
    c = F.type
    N, dim = state.N, state.dim
    dt = system.dt
    e = dom.restitution_coefficient
    cid = state.clump_id
    pos_p_lab = state._pos_p_rot
    pos = state.pos_c + pos_p_lab
    rad = state.rad[:, None]
    lo = dom.anchor + rad
    hi = dom.anchor + dom.box_size - rad
    over_lo = np.maximum(c(0.0), lo - pos)
    over_hi = np.maximum(c(0.0), pos - hi)
    max_lo = _segment("max", over_lo, cid, N)[cid]
    max_hi = _segment("max", over_hi, cid, N)[cid]
    inv_mass = c(1.0) / state.mass
    inv_inertia = c(1.0) / state.inertia

    eye = np.eye(dim, dtype=F)
    # n_prime[p, a, :] = rotate_back(q_p, e_a)  (reflect.py:189-190)
    n_prime = np.stack(
        [la.q_rotate_back(state.q_w, state.q_xyz, np.broadcast_to(eye[a], (N, dim))) for a in range(dim)],
        axis=1,
    )
    r_p_cross_n = la.cross(state.pos_p[:, None, :], n_prime)  # (N, dim, A)
    denom_rot = np.einsum("ni,nji->nj", inv_inertia, r_p_cross_n * r_p_cross_n)
    denom = inv_mass[:, None] + denom_rot
    denom = np.where(denom == 0.0, c(1.0), denom)

    delta = np.maximum(max_lo, max_hi)
    is_lo = (over_lo > 0) & (over_lo == max_lo)
    is_hi = (over_hi > 0) & (over_hi == max_hi)
    wall_sign = is_lo.astype(F) - is_hi.astype(F)
    active = np.abs(wall_sign)

    v_contact_step = state.vel + la.cross_3X3D_1X2D(state.ang_vel, pos_p_lab)
    acc_clump = state.force * inv_mass[:, None]
    if dim == 3:
        R = n_prime
        R_T = np.swapaxes(n_prime, -1, -2)
        torque_body = np.einsum("nij,nj->ni", R_T, state.torque)
        ang_vel_body = np.einsum("nij,nj->ni", R_T, state.ang_vel)
        alpha_body = (torque_body - la.cross(ang_vel_body, state.inertia * ang_vel_body)) * inv_inertia
        alpha_rot = np.einsum("nij,nj->ni", R, alpha_body)
    else:
        alpha_rot = state.torque * inv_inertia
    acc_contact = acc_clump + la.cross_3X3D_1X2D(alpha_rot, pos_p_lab)

    alpha = verlet_collision_fraction(v_contact_step, acc_contact, delta, wall_sign, dt)
    alpha_min_dim = np.min(np.where(active > 0, alpha, c(1.0)), axis=-1, keepdims=True)
    alpha_clump = _segment("min", alpha_min_dim, cid, N)[cid]  # (N,1)

    dt_factor = (alpha_clump - c(1.0)) * dt
    v_contact = v_contact_step + dt_factor * acc_contact
    j_mag = -(c(1.0) + e) * v_contact / denom
    closing = (v_contact * wall_sign) < 0.0
    count_active = _segment("sum", active, cid, N)
    count_safe = np.maximum(count_active, c(1.0))
    weight = active / count_safe[cid]
    j_mag = j_mag * (weight * closing)

    dv = _segment("sum", j_mag * inv_mass[:, None], cid, N)[cid]
    dv = np.where(state.fixed[:, None], c(0.0), dv)
    state.vel = state.vel + dv

    if dim == 3:
        j_body = np.einsum("nij,nj->ni", R_T, j_mag)
    else:
        j_body = np.einsum("nji,nj->ni", n_prime, j_mag)
    moment = la.cross(state.pos_p, j_body)
    if dim == 2:
        d_omega_lab = moment[:, -1:] * inv_inertia
    else:
        d_omega_lab = np.einsum("nij,nj->ni", R, moment * inv_inertia)
    d_omega = _segment("sum", d_omega_lab, cid, N)[cid]
    d_omega = np.where(state.fixed[:, None], c(0.0), d_omega)
    state.ang_vel = state.ang_vel + d_omega

    dt_rem = (c(1.0) - alpha_clump) * dt
    dtheta = d_omega * dt_rem
    if dim == 2:
        dtheta = np.concatenate([np.zeros_like(dtheta), np.zeros_like(dtheta), dtheta], axis=-1)
    dq_w, dq_xyz = la.q_from_small_rotvec(dtheta)
    qw, qxyz = la.q_mul(dq_w, dq_xyz, state.q_w, state.q_xyz)  # LEFT multiply (reflect.py:293)
    qw, qxyz = la.q_unit(qw, qxyz)
    state.set_q(qw, qxyz)
    state.pos_c = state.pos_c + dv * dt_rem


def domain_apply(state, system):
    k = system.domain.kind
    if k == "reflect":
        reflect_apply(state, system)
    elif k == "free":
        free_apply(state, system)
    # periodic: inherited no-op (domains/__init__.py:156-189)
