"""Per-pair ``force`` / ``energy`` of the contact laws in torch: the ForceModel plugin contract
(jaxdem/forces/__init__.py:55-150 — ``force(i, j, pos, state, system) -> (force, torque)`` ON particle i DUE TO
particle j, ``energy(i, j, pos, state, system)``; i, j are indices or index tensors).

These are NOT the step path: every collider evaluates the laws inside its CUDA kernels (csrc/laws.cuh, the same
formulas).  They exist so that code written against the reference's contract — a user collider, an analysis
routine, a unit test of a law (tests/test_clump_pair_friction.py:167-217 calls ``force`` directly) — finds the
methods it expects, on the State's device, vectorised over (i, j)."""

from __future__ import annotations

import math

import torch


def _disp(pos, i, j, system):
    """Domain._displacement, multiply form (domains/periodic.py:75-79)."""
    rij = pos[..., i, :] - pos[..., j, :]
    dom = system.domain
    if dom.native_kind == "periodic":
        rij = rij - dom.box_size * torch.round(rij * dom.inv_box_size)
    return rij


def _different(i, j, like):
    return torch.as_tensor(i, device=like.device).ne(torch.as_tensor(j, device=like.device)).to(like.dtype)


def _unit_and_norm(v):
    """utils/linalg.py:162-181."""
    n2 = (v * v).sum(-1)
    inv = torch.where(n2 == 0, torch.zeros_like(n2), torch.rsqrt(n2.clamp_min(1e-16)))
    return v * inv[..., None], n2 * inv


def _pad3(v):
    """2D vectors / scalar angular quantities embedded in 3D (z = 0, omega = (0, 0, w))."""
    if v.shape[-1] == 3:
        return v
    if v.shape[-1] == 2:
        return torch.cat([v, torch.zeros_like(v[..., :1])], -1)
    return torch.cat([torch.zeros_like(v), torch.zeros_like(v), v], -1)


def spring_force(i, j, pos, state, system):
    """forces/spring.py:97-108."""
    R = state.rad[..., i] + state.rad[..., j]
    rij = _disp(pos, i, j, system)
    d2 = (rij * rij).sum(-1)
    k = system.mat_table.young_eff[state.mat_id[..., i], state.mat_id[..., j]]
    inv = torch.where(d2 == 0, torch.zeros_like(d2), torch.rsqrt(d2.clamp_min(1e-16)))
    delta = (R - d2 * inv).clamp_min(0) * _different(i, j, d2)
    f = (k * delta * inv)[..., None] * rij
    return f, torch.zeros((*f.shape[:-1], state.torque.shape[-1]), dtype=f.dtype, device=f.device)


def spring_energy(i, j, pos, state, system):
    """forces/spring.py:136-147."""
    R = state.rad[..., i] + state.rad[..., j]
    rij = _disp(pos, i, j, system)
    d2 = (rij * rij).sum(-1)
    k = system.mat_table.young_eff[state.mat_id[..., i], state.mat_id[..., j]]
    inv = torch.where(d2 == 0, torch.zeros_like(d2), torch.rsqrt(d2.clamp_min(1e-16)))
    s = (R - d2 * inv).clamp_min(0) * _different(i, j, d2)
    return 0.5 * k * s * s


def _hertz_k(i, j, state, system):
    mt = system.mat_table
    mi, mj = state.mat_id[..., i], state.mat_id[..., j]
    Ei, Ej, ni, nj = mt.young[mi], mt.young[mj], mt.poisson[mi], mt.poisson[mj]
    ri, rj = state.rad[..., i], state.rad[..., j]
    Rs = (ri * rj) / (ri + rj)
    Es = 1.0 / ((1.0 - ni * ni) / Ei + (1.0 - nj * nj) / Ej)
    return (4.0 / 3.0) * Es * torch.sqrt(Rs), ri + rj


def hertz_force(i, j, pos, state, system):
    """forces/hertz.py:98-119."""
    k, R = _hertz_k(i, j, state, system)
    n, r = _unit_and_norm(_disp(pos, i, j, system))
    delta = (R - r).clamp_min(0) * _different(i, j, r)
    f = (k * delta * torch.sqrt(delta))[..., None] * n
    return f, torch.zeros((*f.shape[:-1], state.torque.shape[-1]), dtype=f.dtype, device=f.device)


def hertz_energy(i, j, pos, state, system):
    """forces/hertz.py:146-162."""
    k, R = _hertz_k(i, j, state, system)
    rij = _disp(pos, i, j, system)
    n2 = (rij * rij).sum(-1)
    r = torch.where(n2 == 0, torch.zeros_like(n2), torch.sqrt(n2.clamp_min(1e-16)))
    delta = (R - r).clamp_min(0) * _different(i, j, r)
    return 0.4 * k * delta * delta * torch.sqrt(delta)


def cundallstrack_force(i, j, pos, state, system):
    """forces/cundall_strack.py:119-198: damped normal spring, Coulomb-capped tangential damping, rolling
    resistance; force and torque on i."""
    mt = system.mat_table
    mi, mj = state.mat_id[..., i], state.mat_id[..., j]
    Ei, Ej, nui, nuj = mt.young[mi], mt.young[mj], mt.poisson[mi], mt.poisson[mj]
    ri, rj = state.rad[..., i], state.rad[..., j]
    ma, mb = state.mass[..., i], state.mass[..., j]
    Gi, Gj = Ei / (2.0 * (1.0 + nui)), Ej / (2.0 * (1.0 + nuj))
    kn = (2.0 * Ei * ri * Ej * rj) / (Ei * ri + Ej * rj)
    kt = (2.0 * Gi * ri * Gj * rj) / (Gi * ri + Gj * rj)
    m_eff = (ma * mb) / (ma + mb)
    e_eff = torch.minimum(mt.e[mi], mt.e[mj])
    mu_eff = torch.minimum(mt.mu[mi], mt.mu[mj])
    ln_e = torch.log(torch.where(e_eff > 0, e_eff, torch.ones_like(e_eff)))
    beta = torch.where(e_eff > 0, -ln_e / torch.sqrt(math.pi * math.pi + ln_e * ln_e), torch.ones_like(e_eff))
    gamma_n, gamma_t = 2.0 * beta * torch.sqrt(kn * m_eff), 2.0 * beta * torch.sqrt(kt * m_eff)
    n, r = _unit_and_norm(_pad3(_disp(pos, i, j, system)))
    delta = ri + rj - r
    contact = (delta > 0).to(delta.dtype) * _different(i, j, delta)
    delta = delta * contact
    rci, rcj = -ri[..., None] * n, rj[..., None] * n
    wi, wj = _pad3(state.ang_vel[..., i, :]), _pad3(state.ang_vel[..., j, :])
    vi, vj = _pad3(state.vel[..., i, :]), _pad3(state.vel[..., j, :])
    vrel = (vi + torch.linalg.cross(wi, rci)) - (vj + torch.linalg.cross(wj, rcj))
    vn = (vrel * n).sum(-1)
    tt, vt = _unit_and_norm(vrel - vn[..., None] * n)
    Fn = (kn * delta - gamma_n * vn).clamp_min(0) * contact
    Ft = torch.minimum(gamma_t * vt, mu_eff * Fn) * contact
    f3 = Fn[..., None] * n - Ft[..., None] * tt
    tq = torch.linalg.cross(rci, f3)
    mur = torch.minimum(mt.mu_r[mi], mt.mu_r[mj])
    R_eff = (ri * rj) / (ri + rj)
    orel = wi - wj
    on2 = (orel * orel).sum(-1)
    ounit = orel * torch.rsqrt(torch.where(on2 == 0, torch.ones_like(on2), on2))[..., None]
    t3 = tq - (mur * R_eff * Fn)[..., None] * ounit
    dim = pos.shape[-1]
    return f3[..., :dim], (t3 if dim == 3 else t3[..., 2:3])


def cundallstrack_energy(i, j, pos, state, system):
    """forces/cundall_strack.py:222-235: elastic energy of the normal spring."""
    mt = system.mat_table
    Ei, Ej = mt.young[state.mat_id[..., i]], mt.young[state.mat_id[..., j]]
    ri, rj = state.rad[..., i], state.rad[..., j]
    kn = (2.0 * Ei * ri * Ej * rj) / (Ei * ri + Ej * rj)
    rij = _disp(pos, i, j, system)
    n2 = (rij * rij).sum(-1)
    r = torch.where(n2 == 0, torch.zeros_like(n2), torch.sqrt(n2.clamp_min(1e-16)))
    delta = (ri + rj - r).clamp_min(0) * _different(i, j, r)
    return 0.5 * kn * delta * delta
