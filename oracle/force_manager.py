"""numpy restatement of ForceManager.apply / compute_potential_energy.  Oracle only.

jaxdem/forces/force_manager.py:338-479 (user force functions omitted: the hot
path under test uses the array buffers + gravity only).
"""

from __future__ import annotations

import numpy as np

from . import linalg as la


class OForceManager:
    def __init__(self, N, dim, gravity=None, dtype=np.float64):
        F = np.dtype(dtype)
        A = 1 if dim == 2 else 3
        self.gravity = np.zeros(dim, F) if gravity is None else np.asarray(gravity, dtype=F)
        self.external_force = np.zeros((N, dim), F)
        self.external_force_com = np.zeros((N, dim), F)
        self.external_torque = np.zeros((N, A), F)


def segment_sum(data, ids, n):
    out = np.zeros((n,) + data.shape[1:], data.dtype)
    np.add.at(out, ids, data)  # sequential in index order: deterministic
    return out


def apply(state, system):
    """force_manager.py:359-423."""
    fm = system.force_manager
    F_contact, T_contact = state.force, state.torque
    r_i = state._pos_p_rot
    F_part = fm.external_force
    T_part = fm.external_torque
    F_com = fm.external_force_com
    count = np.bincount(state.clump_id, minlength=state.N)[state.clump_id].astype(state.idtype)
    F_com = F_com + fm.gravity * (state.mass / count)[:, None]
    T_part = T_part + la.cross(r_i, F_part)
    F_total = F_contact + F_part + F_com
    T_total = T_contact + T_part
    F_clump = segment_sum(F_total, state.clump_id, state.N)
    T_clump = segment_sum(T_total, state.clump_id, state.N)
    state.force = F_clump[state.clump_id].astype(state.fdtype)
    state.torque = T_clump[state.clump_id].astype(state.fdtype)
    fm.external_force = np.zeros_like(fm.external_force)
    fm.external_force_com = np.zeros_like(fm.external_force_com)
    fm.external_torque = np.zeros_like(fm.external_torque)


def compute_potential_energy(state, system):
    """force_manager.py:427-479 (gravity part)."""
    fm = system.force_manager
    count = np.bincount(state.clump_id, minlength=state.N)[state.clump_id]
    return -np.sum(la.dot(fm.gravity, state.pos_c) * state.mass / count)
