"""numpy restatement of the State container (jaxdem/state.py:103-304,647-867).

Oracle only.  Unbatched: arrays are (N, dim) / (N,).  ``float``/``int`` pairs
are float32/int32 ("x64 off") or float64/int64 ("x64 on", what the reference
test-suite runs, tests/conftest.py:4).
"""

from __future__ import annotations

import numpy as np

from . import linalg as la


def int_dtype_for(fdtype) -> np.dtype:
    return np.dtype(np.int32) if np.dtype(fdtype) == np.float32 else np.dtype(np.int64)


class OState:
    """Fields as jaxdem/state.py:103-228.  Assigning ``q_w``/``q_xyz``/``pos_p``
    refreshes the cached ``_pos_p_rot`` like State.__setattr__
    (jaxdem/state.py:264-273)."""

    _FIELDS = (
        "pos_c", "pos_p", "vel", "force", "q_w", "q_xyz", "ang_vel", "torque",
        "rad", "_rad", "volume", "mass", "inertia", "clump_id", "bond_id",
        "mat_id", "species_id", "fixed", "_pos_p_rot",
    )

    def __init__(self, **kw):
        for k in self._FIELDS:
            object.__setattr__(self, k, kw.get(k))
        if self._pos_p_rot is None:
            self.refresh_cache()

    def refresh_cache(self):
        object.__setattr__(self, "_pos_p_rot", la.q_rotate(self.q_w, self.q_xyz, self.pos_p))

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name in ("q_w", "q_xyz", "pos_p"):
            self.refresh_cache()

    def set_q(self, w, xyz):
        object.__setattr__(self, "q_w", w)
        object.__setattr__(self, "q_xyz", xyz)
        self.refresh_cache()

    @property
    def N(self) -> int:
        return self.pos_c.shape[-2]

    @property
    def dim(self) -> int:
        return self.pos_c.shape[-1]

    @property
    def fdtype(self):
        return self.pos_c.dtype

    @property
    def idtype(self):
        return self.clump_id.dtype

    @property
    def pos(self) -> np.ndarray:
        """jaxdem/state.py:295-304."""
        return self.pos_c + self._pos_p_rot

    def copy(self) -> "OState":
        return OState(**{k: np.array(getattr(self, k), copy=True) for k in self._FIELDS})


def _hypersphere_volume(rad, dim):
    if dim == 2:
        return np.pi * rad**2
    return (4.0 / 3.0) * np.pi * rad**3


def symmetrize_bonds(conn: np.ndarray, N: int) -> np.ndarray:
    """jaxdem/state.py `_symmetrize` inside State.create (:728-757)."""
    w = max(conn.shape[-1], 1)
    rows_idx = np.repeat(np.arange(N), conn.shape[-1])
    cols = conn.reshape(-1)
    valid = (cols >= 0) & (cols < N)
    i_idx, j_idx = rows_idx[valid], cols[valid]
    if i_idx.size == 0:
        return np.full((N, w), -1, dtype=np.int64)
    pairs = np.unique(
        np.stack((np.concatenate((i_idx, j_idx)), np.concatenate((j_idx, i_idx))), axis=1), axis=0
    )
    counts = np.bincount(pairs[:, 0], minlength=N)
    max_deg = max(int(counts.max()), w)
    out = np.full((N, max_deg), -1, dtype=np.int64)
    starts = np.cumsum(counts) - counts
    slot = np.arange(pairs.shape[0]) - starts[pairs[:, 0]]
    out[pairs[:, 0], slot] = pairs[:, 1]
    return out


def create_state(
    pos, *, pos_p=None, vel=None, force=None, q=None, ang_vel=None, torque=None, rad=None,
    mass=None, inertia=None, clump_id=None, bond_id=None, mat_id=None, fixed=None,
    mat_table=None, dtype=np.float64,
) -> OState:
    """Defaults of State.create (jaxdem/state.py:647-867)."""
    F = np.dtype(dtype)
    I = int_dtype_for(F)
    pos_c = np.asarray(pos, dtype=F)
    N, dim = pos_c.shape
    A = 1 if dim == 2 else 3
    f = lambda x, shape: np.zeros(shape, F) if x is None else np.asarray(x, dtype=F)
    pos_p = f(pos_p, (N, dim))
    vel = f(vel, (N, dim))
    force = f(force, (N, dim))
    if q is None:
        q_w, q_xyz = np.ones((N, 1), F), np.zeros((N, 3), F)
    else:
        q = np.asarray(q, dtype=F)
        q_w, q_xyz = q[:, 0:1].copy(), q[:, 1:].copy()
    ang_vel = f(ang_vel, (N, A))
    torque = f(torque, (N, A))
    rad = np.ones(N, F) if rad is None else np.asarray(rad, dtype=F)
    volume = _hypersphere_volume(rad.astype(np.float64), dim).astype(F)
    clump = np.arange(N, dtype=I) if clump_id is None else np.asarray(clump_id, dtype=I)
    if bond_id is None:
        bond = np.full((N, 1), -1, dtype=I)
    else:
        try:
            conn = np.asarray(bond_id, dtype=np.int64)
            if conn.ndim == 1:
                conn = conn[:, None]
        except (ValueError, TypeError):
            rows = [list(np.asarray(r).ravel()) for r in bond_id]
            width = max(max((len(r) for r in rows), default=0), 1)
            conn = np.full((N, width), -1, dtype=np.int64)
            for i, r in enumerate(rows):
                conn[i, : len(r)] = r
        bond = symmetrize_bonds(conn, N).astype(I)
    mat_id = np.zeros(N, I) if mat_id is None else np.asarray(mat_id, dtype=I)
    fixed = np.zeros(N, bool) if fixed is None else np.asarray(fixed, dtype=bool)
    if mat_table is not None:
        mass = (mat_table.density[mat_id] * volume).astype(F)
    else:
        mass = np.ones(N, F) if mass is None else np.asarray(mass, dtype=F)
    coeff = F.type(0.5 if dim == 2 else 0.4)
    inertia = (
        (coeff * mass * rad**2)[:, None] * np.ones((N, A), F)
        if inertia is None
        else np.asarray(inertia, dtype=F)
    )
    if N > 0:
        clump = np.unique(clump, return_inverse=True)[1].astype(I)  # relabel to dense ids
    return OState(
        pos_c=pos_c, pos_p=pos_p, vel=vel, force=force, q_w=q_w, q_xyz=q_xyz, ang_vel=ang_vel,
        torque=torque, rad=rad, _rad=rad.copy(), volume=volume, mass=mass, inertia=inertia,
        clump_id=clump, bond_id=bond, mat_id=mat_id, species_id=np.zeros(N, I), fixed=fixed,
    )


def grid_state(n_per_axis, spacing, radius=1.0, *, mass=1.0, jitter=0.0, vel_range=None,
               radius_range=None, seed=0, dtype=np.float64) -> OState:
    """Lattice state in the spirit of jaxdem/utils/grid_state.py:20-140 (numpy
    PCG64 RNG, not threefry: values differ from the reference's, recipe is the same)."""
    rng = np.random.default_rng(seed)
    dim = len(n_per_axis)
    spacing = np.broadcast_to(np.asarray(spacing, dtype=np.float64), (dim,))
    axes = [np.arange(n) * s for n, s in zip(n_per_axis, spacing)]
    mesh = np.meshgrid(*axes, indexing="ij")
    pos = np.stack([m.ravel() for m in mesh], axis=1)
    N = pos.shape[0]
    if jitter:
        pos = pos + rng.uniform(-jitter, jitter, size=pos.shape)
    vr = (-1.0, 1.0) if vel_range is None else vel_range
    vel = rng.uniform(vr[0], vr[1], size=(N, dim))
    if radius_range is not None:
        rad = rng.uniform(radius_range[0], radius_range[1], size=N)
    else:
        rad = np.full(N, radius)
    return create_state(pos, vel=vel, rad=rad, mass=np.full(N, mass), dtype=dtype)
