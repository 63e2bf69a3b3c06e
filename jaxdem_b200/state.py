"""State container mirroring jaxdem.State (reference jaxdem/state.py:103-304,647-867).

SoA particle data as dense row-major CUDA tensors; these buffers ARE the leaves the
C ABI reads and updates in place (include/jaxdem_b200.h: jdb200_state).  Constructors
are host-side setup code (run once); none of it is on the step path.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

_DEFAULT_FLOAT = torch.float32  # "jax_enable_x64 off"


def set_default_dtype(dtype: torch.dtype) -> None:
    """float32/int32 (x64 off, default) or float64/int64 (x64 on; what the reference's
    test-suite uses, tests/conftest.py:4)."""
    global _DEFAULT_FLOAT
    assert dtype in (torch.float32, torch.float64)
    _DEFAULT_FLOAT = dtype


def default_float() -> torch.dtype:
    return _DEFAULT_FLOAT


def int_dtype_for(fdtype: torch.dtype) -> torch.dtype:
    return torch.int32 if fdtype == torch.float32 else torch.int64


def default_device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


@dataclass
class Quaternion:
    """(w (...,1), xyz (...,3)) as jaxdem.utils.Quaternion (utils/quaternion.py:23-70)."""

    w: torch.Tensor
    xyz: torch.Tensor

    @staticmethod
    def rotate(q: "Quaternion", v: torch.Tensor) -> torch.Tensor:
        """Host-side setup helper (quaternion.py:190-242); the step path uses the CUDA kernels."""
        dim = v.shape[-1]
        if dim == 2:
            qz = q.xyz[..., 2:3]
            c = q.w * q.w - qz * qz
            s = 2.0 * q.w * qz
            return torch.cat([c * v[..., 0:1] - s * v[..., 1:2], s * v[..., 0:1] + c * v[..., 1:2]], dim=-1)
        t = 2.0 * torch.linalg.cross(q.xyz, v)
        return v + q.w * t + torch.linalg.cross(q.xyz, t)

    @staticmethod
    def rotate_back(q: "Quaternion", v: torch.Tensor) -> torch.Tensor:
        return Quaternion.rotate(Quaternion(q.w, -q.xyz), v)


def _hypersphere_volume(rad: torch.Tensor, dim: int) -> torch.Tensor:
    return math.pi * rad**2 if dim == 2 else (4.0 / 3.0) * math.pi * rad**3


def _symmetrize(conn: np.ndarray, N: int) -> np.ndarray:
    """Symmetric bond table: if j is listed for i then i is listed for j (what state.py:728-757 guarantees);
    rows hold ascending partner ids, -1 padded, width = the widest row (at least the input width).
    Every directed edge (i -> j) and its mirror become one 64-bit key i * N + j; sorting the unique keys
    groups them by row with partners ascending."""
    width = max(conn.shape[-1], 1)
    src = np.broadcast_to(np.arange(N, dtype=np.int64)[:, None], conn.shape)
    keep = (conn >= 0) & (conn < N)
    a, b = src[keep], conn[keep].astype(np.int64)
    table = np.full((N, width), -1, dtype=np.int64)
    if a.size == 0:
        return table
    edge = np.unique(np.concatenate((a * N + b, b * N + a)))
    row, partner = edge // N, edge % N
    first = np.searchsorted(row, np.arange(N))            # where each row's block of keys begins
    col = np.arange(edge.size) - first[row]
    if col.max() >= width:
        table = np.full((N, int(col.max()) + 1), -1, dtype=np.int64)
    table[row, col] = partner
    return table


_FIELDS = ("pos_c", "pos_p", "vel", "force", "q", "ang_vel", "torque", "rad", "_rad", "volume", "mass",
           "inertia", "clump_id", "bond_id", "mat_id", "species_id", "fixed", "_pos_p_rot")


class State:
    """Fields as jaxdem/state.py:103-228 (facet fields omitted: meshes are out of scope).
    ``has_clumps`` is static metadata computed once at construction: False promises
    ``clump_id == arange(N)`` so the clump reductions are identities (jdb200_params.clumps)."""

    def __init__(self, **kw: Any) -> None:
        for k in _FIELDS:
            object.__setattr__(self, k, kw.get(k))
        object.__setattr__(self, "has_clumps", bool(kw.get("has_clumps", True)))
        if self._pos_p_rot is None:
            self.refresh_cache()

    def promise_bits(self) -> int:
        """jdb200_params.promises this State can vouch for: no bonds / nothing fixed / pos_p == 0.  The three
        reductions run once (one host synchronisation at first use) and are cached against the tensors' identity
        and torch version counters, so an in-place edit or a re-assigned leaf re-evaluates them."""
        from . import _lib
        key = tuple((t.data_ptr(), t._version) for t in (self.bond_id, self.fixed, self.pos_p))
        cached = self.__dict__.get("_promise_cache")
        if cached is None or cached[0] != key:
            bits = 0
            if self.bond_id.numel() == 0 or bool((self.bond_id < 0).all()):
                bits |= _lib.PROMISE_NO_BONDS
            if not bool(self.fixed.any()):
                bits |= _lib.PROMISE_NO_FIXED
            if not bool((self.pos_p != 0).any()) and not bool((self._pos_p_rot != 0).any()):
                bits |= _lib.PROMISE_NO_POS_P
            cached = (key, bits)
            object.__setattr__(self, "_promise_cache", cached)
        return cached[1]

    def refresh_cache(self) -> None:
        object.__setattr__(self, "_pos_p_rot", Quaternion.rotate(self.q, self.pos_p).contiguous())

    def __setattr__(self, name: str, value: Any) -> None:
        object.__setattr__(self, name, value)
        if name in ("q", "pos_p"):  # State.__setattr__ cache refresh (state.py:264-273)
            self.refresh_cache()
        if name == "clump_id":
            object.__setattr__(self, "has_clumps", True)

    # -- shape helpers (state.py:275-293) -----------------------------------
    @property
    def N(self) -> int:
        return self.pos_c.shape[-2]

    @property
    def dim(self) -> int:
        return self.pos_c.shape[-1]

    @property
    def shape(self) -> tuple:
        return tuple(self.pos_c.shape)

    @property
    def batch_size(self) -> int:
        return 1 if self.pos_c.ndim < 3 else self.pos_c.shape[-3]

    @property
    def dtype(self) -> torch.dtype:
        return self.pos_c.dtype

    @property
    def device(self) -> torch.device:
        return self.pos_c.device

    @property
    def pos(self) -> torch.Tensor:
        """pos = pos_c + R(q) pos_p (state.py:295-304)."""
        return self.pos_c + self._pos_p_rot

    def clone(self) -> "State":
        kw = {}
        for k in _FIELDS:
            v = getattr(self, k)
            kw[k] = Quaternion(v.w.clone(), v.xyz.clone()) if k == "q" else v.clone()
        return State(has_clumps=self.has_clumps, **kw)

    def to_numpy(self) -> dict:
        out = {}
        for k in _FIELDS:
            v = getattr(self, k)
            if k == "q":
                out["q_w"], out["q_xyz"] = v.w.cpu().numpy(), v.xyz.cpu().numpy()
            else:
                out[k] = v.cpu().numpy()
        return out

    # -- constructors ---------------------------------------------------------
    @staticmethod
    def create(pos, *, pos_p=None, vel=None, force=None, q=None, ang_vel=None, torque=None, rad=None,
               volume=None, mass=None, inertia=None, clump_id=None, bond_id=None, mat_id=None,
               species_id=None, fixed=None, mat_table=None, dtype: torch.dtype | None = None,
               device=None) -> "State":
        """Defaults of State.create (state.py:647-867).  Accepts (N, dim) or (B, N, dim)."""
        F = dtype or _DEFAULT_FLOAT
        I = int_dtype_for(F)
        dev = torch.device(device) if device is not None else default_device()
        tf = lambda x: torch.as_tensor(np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x), dtype=F)
        pos_c = tf(pos)
        lead = pos_c.shape[:-1]
        N, dim = pos_c.shape[-2], pos_c.shape[-1]
        A = 1 if dim == 2 else 3
        zeros = lambda shape: torch.zeros(shape, dtype=F)
        pos_p = zeros(pos_c.shape) if pos_p is None else tf(pos_p)
        vel = zeros(pos_c.shape) if vel is None else tf(vel)
        force = zeros(pos_c.shape) if force is None else tf(force)
        if q is None:
            qq = Quaternion(torch.ones((*lead, 1), dtype=F), zeros((*lead, 3)))
        elif isinstance(q, Quaternion):
            qq = Quaternion(tf(q.w), tf(q.xyz))
        else:
            qa = tf(q)
            qq = Quaternion(qa[..., 0:1].clone(), qa[..., 1:].clone())
        ang_vel = zeros((*lead, A)) if ang_vel is None else tf(ang_vel)
        torque = zeros((*lead, A)) if torque is None else tf(torque)
        rad = torch.ones(lead, dtype=F) if rad is None else tf(rad)
        volume = _hypersphere_volume(rad.double(), dim).to(F) if volume is None else tf(volume)
        has_clumps = clump_id is not None
        if clump_id is None:
            cid = torch.arange(N, dtype=torch.int64).expand(lead).clone()
        else:
            cid_np = np.asarray(clump_id.cpu() if isinstance(clump_id, torch.Tensor) else clump_id, dtype=np.int64)
            flat = cid_np.reshape(-1, N)
            flat = np.stack([np.unique(r, return_inverse=True)[1] for r in flat])  # dense relabel (state.py:825-833)
            cid = torch.as_tensor(flat.reshape(cid_np.shape))
            has_clumps = bool((flat != np.arange(N)[None, :]).any())
        if bond_id is None:
            bond = torch.full((*lead, 1), -1, dtype=torch.int64)
        else:
            try:
                conn = np.asarray(bond_id, dtype=np.int64)
                if conn.ndim == len(lead):
                    conn = conn[..., None]
            except (ValueError, TypeError):
                rows = [list(np.asarray(r).ravel()) for r in bond_id]
                width = max(max((len(r) for r in rows), default=0), 1)
                conn = np.full((N, width), -1, dtype=np.int64)
                for i, r in enumerate(rows):
                    conn[i, : len(r)] = r
            if conn.ndim == 2:
                bond_np = _symmetrize(conn, N)
            else:
                outs = [_symmetrize(c, N) for c in conn.reshape(-1, *conn.shape[-2:])]
                w = max(o.shape[-1] for o in outs)
                outs = [np.pad(o, ((0, 0), (0, w - o.shape[-1])), constant_values=-1) for o in outs]
                bond_np = np.stack(outs).reshape(*conn.shape[:-2], N, w)
            bond = torch.as_tensor(bond_np)
        mat = torch.zeros(lead, dtype=torch.int64) if mat_id is None else torch.as_tensor(np.asarray(mat_id), dtype=torch.int64)
        species = torch.zeros(lead, dtype=torch.int64) if species_id is None else torch.as_tensor(np.asarray(species_id), dtype=torch.int64)
        fixed_t = torch.zeros(lead, dtype=torch.bool) if fixed is None else torch.as_tensor(np.asarray(fixed), dtype=torch.bool)
        if mat_table is not None:
            mass_t = (mat_table.density.cpu().double()[mat] * volume.double()).to(F)
        else:
            mass_t = torch.ones(lead, dtype=F) if mass is None else tf(mass)
        coeff = 0.5 if dim == 2 else 0.4
        inertia_t = ((coeff * mass_t * rad**2)[..., None] * torch.ones((*lead, A), dtype=F)
                     if inertia is None else tf(inertia))
        d = lambda t: t.to(dev).contiguous()
        return State(
            pos_c=d(pos_c), pos_p=d(pos_p), vel=d(vel), force=d(force), q=Quaternion(d(qq.w), d(qq.xyz)),
            ang_vel=d(ang_vel), torque=d(torque), rad=d(rad), _rad=d(rad.clone()), volume=d(volume),
            mass=d(mass_t), inertia=d(inertia_t), clump_id=d(cid.to(I)), bond_id=d(bond.to(I)),
            mat_id=d(mat.to(I)), species_id=d(species.to(I)), fixed=d(fixed_t), has_clumps=has_clumps,
        )

    @staticmethod
    def stack(states: list["State"]) -> "State":
        """State.stack (state.py): batch along a new leading axis."""
        kw = {}
        for k in _FIELDS:
            if k == "q":
                kw[k] = Quaternion(torch.stack([s.q.w for s in states]), torch.stack([s.q.xyz for s in states]))
            else:
                kw[k] = torch.stack([getattr(s, k) for s in states]).contiguous()
        return State(has_clumps=any(s.has_clumps for s in states), **kw)
