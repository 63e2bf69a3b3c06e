"""Energy diagnostics mirroring jaxdem/utils/thermal.py:24-177 (host-driven, off the step path): torch
reductions over the State leaves plus the collider's potential-energy hook (a C-ABI call)."""

from __future__ import annotations

import torch

from .state import Quaternion, State


def _member_count(state: State) -> torch.Tensor:
    cid = state.clump_id.long()
    cnt = torch.zeros_like(cid).scatter_add_(-1, cid, torch.ones_like(cid))
    return torch.gather(cnt, -1, cid).to(state.dtype)


def compute_translational_kinetic_energy(state: State) -> torch.Tensor:
    """thermal.py:24-47,52-78: sum of 0.5 (m / count) |v|^2 (clump members each carry the clump's mass)."""
    w = state.mass / _member_count(state)
    return (0.5 * w * (state.vel * state.vel).sum(-1)).sum(-1)


def compute_rotational_kinetic_energy(state: State) -> torch.Tensor:
    """thermal.py:52-78: 0.5 w_body . (I w_body) / count, body-frame angular velocity in 3D."""
    wb = state.ang_vel if state.dim == 2 else Quaternion.rotate_back(state.q, state.ang_vel)
    return (0.5 * (wb * state.inertia * wb).sum(-1) / _member_count(state)).sum(-1)


def compute_potential_energy(state: State, system) -> torch.Tensor:
    """thermal.py:125-150: ForceManager (gravity, force_manager.py:427-479) + collider."""
    g = system.force_manager.gravity
    pe_g = -((g[..., None, :] * state.pos_c).sum(-1) * state.mass / _member_count(state)).sum(-1)
    _, _, pe_c = system.collider.compute_potential_energy(state, system)
    return pe_g + pe_c


def compute_energy(state: State, system) -> torch.Tensor:
    """thermal.py:153-177: total mechanical energy."""
    return (compute_potential_energy(state, system) + compute_translational_kinetic_energy(state)
            + compute_rotational_kinetic_energy(state))
