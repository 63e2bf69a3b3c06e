"""numpy restatement of jaxdem/utils/linalg.py and jaxdem/utils/quaternion.py.

Oracle only (see oracle/__init__.py).  All functions preserve the input float
dtype (float32 stays float32) so f32 runs mimic the reference with x64 off.
"""

from __future__ import annotations

import numpy as np


def _c(x, ref):
    """Python scalar -> numpy scalar of ref's dtype (keeps f32 math in f32)."""
    return ref.dtype.type(x)


def cross(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """jaxdem/utils/linalg.py:13-66.  2D returns shape (..., 1)."""
    if a.shape[-1] == 2:
        return a[..., 0:1] * b[..., 1:2] - a[..., 1:2] * b[..., 0:1]
    ax, ay, az = a[..., 0], a[..., 1], a[..., 2]
    bx, by, bz = b[..., 0], b[..., 1], b[..., 2]
    return np.stack([ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx], axis=-1)


def dot(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """jaxdem/utils/linalg.py:69-89."""
    return np.sum(a * b, axis=-1)


def norm2(v: np.ndarray) -> np.ndarray:
    """jaxdem/utils/linalg.py:92-110."""
    return dot(v, v)


def rsqrt(x: np.ndarray) -> np.ndarray:
    one = np.asarray(1.0, dtype=x.dtype)
    return one / np.sqrt(x)


def norm(v: np.ndarray) -> np.ndarray:
    """jaxdem/utils/linalg.py:113-133 (zero-safe)."""
    n2 = norm2(v)
    safe = np.maximum(n2, _c(1e-16, n2))
    return np.where(n2 == 0.0, _c(0.0, n2), np.sqrt(safe))


def unit(v: np.ndarray) -> np.ndarray:
    """jaxdem/utils/linalg.py:136-159."""
    n2 = norm2(v)
    safe = np.where(n2 == 0.0, _c(1.0, n2), n2)
    return v * rsqrt(safe)[..., None]


def unit_and_norm(v: np.ndarray):
    """jaxdem/utils/linalg.py:162-181."""
    n2 = norm2(v)
    safe = np.maximum(n2, _c(1e-16, n2))
    inv = np.where(n2 == 0.0, _c(0.0, n2), rsqrt(safe))
    return v * inv[..., None], n2 * inv


def cross_3X3D_1X2D(w: np.ndarray, r: np.ndarray) -> np.ndarray:
    """jaxdem/utils/linalg.py:184-231.  w: (...,3) in 3D, (...,1) in 2D."""
    if r.shape[-1] == 2:
        return np.concatenate([-w * r[..., 1:2], w * r[..., 0:1]], axis=-1)
    return cross(w, r)


# --------------------------------------------------------------------------
# Quaternions: (w (...,1), xyz (...,3)) pairs, as in jaxdem/utils/quaternion.py
# --------------------------------------------------------------------------


def q_unit(w: np.ndarray, xyz: np.ndarray):
    """jaxdem/utils/quaternion.py:72-94."""
    n2 = w[..., 0] * w[..., 0] + dot(xyz, xyz)
    safe = np.where(n2 == 0.0, _c(1.0, n2), n2)
    inv = rsqrt(safe)
    return w * inv[..., None], xyz * inv[..., None]


def q_from_small_rotvec(rotvec: np.ndarray):
    """jaxdem/utils/quaternion.py:129-141 (2nd-order Taylor)."""
    n2 = dot(rotvec, rotvec)
    cos_half = _c(1.0, n2) - n2 / _c(8.0, n2)
    sinc = _c(0.5, n2) - n2 / _c(48.0, n2)
    return cos_half[..., None], rotvec * sinc[..., None]


def q_rotate(w: np.ndarray, xyz: np.ndarray, v: np.ndarray) -> np.ndarray:
    """jaxdem/utils/quaternion.py:190-242 (body -> lab)."""
    dim = v.shape[-1]
    if dim == 2:
        qz = xyz[..., 2:3]
        c = w * w - qz * qz
        s = _c(2.0, w) * w * qz
        vx, vy = v[..., 0:1], v[..., 1:2]
        return np.concatenate([c * vx - s * vy, s * vx + c * vy], axis=-1)
    t = _c(2.0, w) * cross(xyz, v)
    return v + w * t + cross(xyz, t)


def q_rotate_back(w: np.ndarray, xyz: np.ndarray, v: np.ndarray) -> np.ndarray:
    """jaxdem/utils/quaternion.py:244-268 (lab -> body)."""
    return q_rotate(w, -xyz, v)


def q_mul(w1, xyz1, w2, xyz2):
    """Hamilton product, jaxdem/utils/quaternion.py:345-367."""
    w = w1 * w2 - dot(xyz1, xyz2)[..., None]
    xyz = w1 * xyz2 + w2 * xyz1 + cross(xyz1, xyz2)
    return w, xyz
