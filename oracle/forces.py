"""numpy restatement of the spring / hertz / cundallstrack pair laws.  Oracle only.

Each law is ``force(i, j, pos, state, system) -> (F, T)`` = force/torque ON i
DUE TO j, and ``energy(i, j, ...) -> E``, exactly as ForceModel
(jaxdem/forces/__init__.py:55-150).  ``i`` and ``j`` are equal-length index
arrays (vectorised over pairs).
"""

from __future__ import annotations

import numpy as np

from . import linalg as la


def _c(x, ref):
    return ref.dtype.type(x)


def spring_force(i, j, pos, state, system):
    """jaxdem/forces/spring.py:97-108."""
    R = state.rad[i] + state.rad[j]
    rij = system.domain._displacement(pos[i], pos[j])
    d2 = np.sum(rij**2, axis=-1)
    k = system.mat_table.young_eff[state.mat_id[i], state.mat_id[j]]
    inv = np.where(d2 == 0.0, _c(0.0, d2), la.rsqrt(np.maximum(d2, _c(1e-16, d2))))
    r = d2 * inv
    delta = np.maximum(_c(0.0, r), R - r) * (i != j)
    mag = k * delta * inv
    A = state.torque.shape[-1]
    return mag[..., None] * rij, np.zeros(i.shape + (A,), state.torque.dtype)


def spring_energy(i, j, pos, state, system):
    """jaxdem/forces/spring.py:136-147."""
    R = state.rad[i] + state.rad[j]
    rij = system.domain._displacement(pos[i], pos[j])
    d2 = np.sum(rij**2, axis=-1)
    k = system.mat_table.young_eff[state.mat_id[i], state.mat_id[j]]
    inv = np.where(d2 == 0.0, _c(0.0, d2), la.rsqrt(np.maximum(d2, _c(1e-16, d2))))
    r = d2 * inv
    s = np.maximum(_c(0.0, r), R - r) * (i != j)
    return _c(0.5, r) * k * s**2


def _hertz_k(i, j, state, system):
    mi, mj = state.mat_id[i], state.mat_id[j]
    mt = system.mat_table
    E_i, E_j = mt.young[mi], mt.young[mj]
    nu_i, nu_j = mt.poisson[mi], mt.poisson[mj]
    R_i, R_j = state.rad[i], state.rad[j]
    one = _c(1.0, R_i)
    R_star = (R_i * R_j) / (R_i + R_j)
    E_star = one / ((one - nu_i**2) / E_i + (one - nu_j**2) / E_j)
    k = _c(4.0 / 3.0, R_i) * E_star * np.sqrt(R_star)
    return k, R_i, R_j


def hertz_force(i, j, pos, state, system):
    """jaxdem/forces/hertz.py:98-119."""
    k, R_i, R_j = _hertz_k(i, j, state, system)
    rij = system.domain._displacement(pos[i], pos[j])
    n, r = la.unit_and_norm(rij)
    delta = np.maximum(_c(0.0, r), R_i + R_j - r) * (i != j)
    mag = k * np.power(delta, _c(1.5, r))
    A = state.torque.shape[-1]
    return mag[..., None] * n, np.zeros(i.shape + (A,), state.torque.dtype)


def hertz_energy(i, j, pos, state, system):
    """jaxdem/forces/hertz.py:146-162."""
    k, R_i, R_j = _hertz_k(i, j, state, system)
    rij = system.domain._displacement(pos[i], pos[j])
    r = la.norm(rij)
    delta = R_i + R_j - r
    delta = delta * ((delta > 0) & (i != j))
    return _c(0.4, r) * k * np.power(delta, _c(2.5, r))


def cundallstrack_force(i, j, pos, state, system):
    """jaxdem/forces/cundall_strack.py:119-198."""
    mt = system.mat_table
    mi, mj = state.mat_id[i], state.mat_id[j]
    E_i, E_j = mt.young[mi], mt.young[mj]
    nu_i, nu_j = mt.poisson[mi], mt.poisson[mj]
    e_i, e_j = mt.e[mi], mt.e[mj]
    mu_i, mu_j = mt.mu[mi], mt.mu[mj]
    mur_i, mur_j = mt.mu_r[mi], mt.mu_r[mj]
    m_i, m_j = state.mass[i], state.mass[j]
    R_i, R_j = state.rad[i], state.rad[j]
    c = lambda x: _c(x, R_i)

    G_i = E_i / (c(2.0) * (c(1.0) + nu_i))
    G_j = E_j / (c(2.0) * (c(1.0) + nu_j))
    kn = (c(2.0) * E_i * R_i * E_j * R_j) / (E_i * R_i + E_j * R_j)
    kt = (c(2.0) * G_i * R_i * G_j * R_j) / (G_i * R_i + G_j * R_j)
    m_eff = (m_i * m_j) / (m_i + m_j)
    e_eff = np.minimum(e_i, e_j)
    mu_eff = np.minimum(mu_i, mu_j)

    e_safe = np.where(e_eff > 0.0, e_eff, c(1.0))
    ln_e = np.log(e_safe)
    beta = np.where(e_eff > 0.0, -ln_e / np.sqrt(c(np.pi) ** 2 + ln_e**2), c(1.0))
    gamma_n = c(2.0) * beta * np.sqrt(kn * m_eff)
    gamma_t = c(2.0) * beta * np.sqrt(kt * m_eff)

    rij = system.domain._displacement(pos[i], pos[j])
    n, r = la.unit_and_norm(rij)
    delta = R_i + R_j - r
    is_contact = ((delta > 0) & (i != j)).astype(r.dtype)
    delta = delta * is_contact

    r_ci = -R_i[..., None] * n
    r_cj = R_j[..., None] * n
    v_ci = state.vel[i] + la.cross_3X3D_1X2D(state.ang_vel[i], r_ci)
    v_cj = state.vel[j] + la.cross_3X3D_1X2D(state.ang_vel[j], r_cj)
    v_rel = v_ci - v_cj
    vn = la.dot(v_rel, n)
    vt_vec = v_rel - vn[..., None] * n
    t, vt = la.unit_and_norm(vt_vec)

    Fn = np.maximum(c(0.0), kn * delta - gamma_n * vn) * is_contact
    Ft = gamma_t * vt
    Ft = np.minimum(Ft, mu_eff * Fn) * is_contact
    F = Fn[..., None] * n - Ft[..., None] * t
    torque = la.cross(r_ci, F)

    mu_r_eff = np.minimum(mur_i, mur_j)
    R_eff = (R_i * R_j) / (R_i + R_j)
    omega_hat = la.unit(state.ang_vel[i] - state.ang_vel[j])
    torque = torque - (mu_r_eff * R_eff * Fn)[..., None] * omega_hat
    return F, torque


def cundallstrack_energy(i, j, pos, state, system):
    """jaxdem/forces/cundall_strack.py:222-235."""
    mt = system.mat_table
    mi, mj = state.mat_id[i], state.mat_id[j]
    E_i, E_j = mt.young[mi], mt.young[mj]
    R_i, R_j = state.rad[i], state.rad[j]
    kn = (_c(2.0, R_i) * E_i * R_i * E_j * R_j) / (E_i * R_i + E_j * R_j)
    rij = system.domain._displacement(pos[i], pos[j])
    r = la.norm(rij)
    delta = R_i + R_j - r
    delta = delta * ((delta > 0) & (i != j))
    return _c(0.5, r) * kn * np.power(delta, _c(2.0, r))


LAWS = {
    "spring": (spring_force, spring_energy),
    "hertz": (hertz_force, hertz_energy),
    "cundallstrack": (cundallstrack_force, cundallstrack_energy),
}
