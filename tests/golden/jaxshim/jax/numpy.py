"""jax.numpy on numpy: every function returns ``Arr``; JAX-only keyword arguments are translated."""
import sys
import types

import numpy as _np

from ._core import Arr, asarr, canon, canon_dtype, wrap

ndarray = Arr
pi, inf, nan, e, newaxis = _np.pi, _np.inf, _np.nan, _np.e, None

_TYPES = ("float16", "float32", "float64", "int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64",
          "bool_", "integer", "floating", "number", "inexact", "complex64", "complex128", "dtype",
          "issubdtype", "promote_types", "result_type", "can_cast", "shape", "ndim", "size", "isscalar",
          "signedinteger", "unsignedinteger", "generic")
for _t in _TYPES:
    globals()[_t] = getattr(_np, _t)


def _dt(dtype):
    return canon_dtype(dtype)


def iinfo(dtype):
    return _np.iinfo(canon_dtype(dtype))


def finfo(dtype):
    return _np.finfo(canon_dtype(dtype))


def asarray(x, dtype=None, **kw):
    return asarr(x, _dt(dtype))


def array(x, dtype=None, copy=True, **kw):
    return canon(_np.array(x, dtype=_dt(dtype), copy=True)).view(Arr)


def zeros(shape, dtype=float, **kw):
    return canon(_np.zeros(shape, dtype=_dt(dtype))).view(Arr)


def ones(shape, dtype=float, **kw):
    return canon(_np.ones(shape, dtype=_dt(dtype))).view(Arr)


def empty(shape, dtype=float, **kw):
    return canon(_np.zeros(shape, dtype=_dt(dtype))).view(Arr)


def full(shape, fill_value, dtype=None, **kw):
    return canon(_np.full(shape, fill_value, dtype=_dt(dtype))).view(Arr)


def arange(*a, dtype=None, **kw):
    return canon(_np.arange(*[_np.asarray(v).item() if isinstance(v, _np.ndarray) else v for v in a],
                            dtype=_dt(dtype))).view(Arr)


def eye(n, m=None, k=0, dtype=float, **kw):
    return canon(_np.eye(n, m, k, dtype=_dt(dtype))).view(Arr)


def zeros_like(x, dtype=None, **kw):
    return canon(_np.zeros_like(_np.asarray(x), dtype=_dt(dtype))).view(Arr)


def ones_like(x, dtype=None, **kw):
    return canon(_np.ones_like(_np.asarray(x), dtype=_dt(dtype))).view(Arr)


def full_like(x, v, dtype=None, **kw):
    return canon(_np.full_like(_np.asarray(x), v, dtype=_dt(dtype))).view(Arr)


def bincount(x, weights=None, minlength=0, *, length=None):
    x = _np.asarray(x)
    if length is not None:
        ok = (x >= 0) & (x < length)
        w = None if weights is None else _np.asarray(weights)[ok]
        return wrap(_np.bincount(x[ok], weights=w, minlength=length)[:length])
    return wrap(_np.bincount(x, weights=weights, minlength=minlength))


def searchsorted(a, v, side="left", sorter=None, *, method=None):
    return wrap(_np.searchsorted(_np.asarray(a), _np.asarray(v), side=side, sorter=sorter))


def unique(x, return_index=False, return_inverse=False, return_counts=False, axis=None, *, size=None,
           fill_value=None, **kw):
    res = _np.unique(_np.asarray(x), return_index=return_index, return_inverse=return_inverse,
                     return_counts=return_counts, axis=axis)
    if size is None:
        return wrap(res)
    tup = res if isinstance(res, tuple) else (res,)
    u = tup[0]
    n = u.shape[0]
    if n >= size:
        u2 = u[:size]
    else:
        fv = u[0] if fill_value is None else fill_value  # jnp pads with the minimum element by default
        u2 = _np.concatenate([u, _np.full((size - n,) + u.shape[1:], fv, dtype=u.dtype)])
    rest = []
    for r, name in zip(tup[1:], [k for k, f in (("index", return_index), ("inverse", return_inverse),
                                                ("counts", return_counts)) if f]):
        if name == "inverse":
            rest.append(r)
        else:
            pad = _np.zeros(max(size - n, 0), dtype=r.dtype)
            rest.append(_np.concatenate([r[:size], pad]))
    out = (u2,) + tuple(rest)
    return wrap(out if len(out) > 1 else out[0])


def where(cond, x=None, y=None, *, size=None, fill_value=None):
    if x is None:
        res = _np.where(_np.asarray(cond))
        if size is not None:
            fv = 0 if fill_value is None else fill_value
            res = tuple(_np.concatenate([r[:size], _np.full(max(size - len(r), 0), fv, dtype=r.dtype)]) for r in res)
        return wrap(res)
    return wrap(_np.where(_np.asarray(cond), _np.asarray(x), _np.asarray(y)))


def nonzero(x, *, size=None, fill_value=None):
    return where(_np.asarray(x) != 0, size=size, fill_value=fill_value)


def argsort(a, axis=-1, kind=None, stable=True, descending=False, **kw):
    a = _np.asarray(a)
    if descending:
        return wrap(_np.argsort(-a, axis=axis, kind="stable"))
    return wrap(_np.argsort(a, axis=axis, kind="stable"))


def sort(a, axis=-1, kind=None, stable=True, descending=False, **kw):
    r = _np.sort(_np.asarray(a), axis=axis, kind="stable")
    return wrap(_np.flip(r, axis=axis) if descending else r)


def take(a, indices, axis=None, mode=None, **kw):
    return wrap(_np.take(_np.asarray(a), _np.asarray(indices), axis=axis, mode="clip"))


def take_along_axis(a, idx, axis, mode=None, **kw):
    a, idx = _np.asarray(a), _np.asarray(idx)
    return wrap(_np.take_along_axis(a, _np.clip(idx, -a.shape[axis], a.shape[axis] - 1), axis=axis))


def clip(x, min=None, max=None, *, a_min=None, a_max=None):
    lo = min if min is not None else a_min
    hi = max if max is not None else a_max
    return wrap(_np.clip(_np.asarray(x), lo, hi))


def setdiff1d(a, b, assume_unique=False, *, size=None, fill_value=None):
    r = _np.setdiff1d(_np.asarray(a), _np.asarray(b), assume_unique=assume_unique)
    if size is not None:
        fv = (r[0] if len(r) else 0) if fill_value is None else fill_value
        r = _np.concatenate([r[:size], _np.full(max(size - len(r), 0), fv, dtype=r.dtype)])
    return wrap(r)


def _generic(name):
    f = getattr(_np, name)

    def g(*a, **k):
        a = [_np.asarray(v) if isinstance(v, Arr) else v for v in a]
        if "dtype" in k:
            k["dtype"] = _dt(k["dtype"])
        return wrap(f(*a, **k))

    g.__name__ = name
    return g


class _Sub(types.ModuleType):
    def __init__(self, name, src):
        super().__init__(name)
        self._src = src

    def __getattr__(self, name):
        f = getattr(self._src, name)
        if callable(f) and not isinstance(f, type):
            def g(*a, **k):
                return wrap(f(*[_np.asarray(v) if isinstance(v, Arr) else v for v in a], **k))
            return g
        return f


linalg = _Sub(__name__ + ".linalg", _np.linalg)
sys.modules[linalg.__name__] = linalg


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    f = getattr(_np, name)
    if callable(f) and not isinstance(f, type):
        g = _generic(name)
        globals()[name] = g
        return g
    return f
