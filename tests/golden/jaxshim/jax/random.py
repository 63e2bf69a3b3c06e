"""Deterministic numpy generators behind the jax.random call signatures (NOT JAX's threefry streams)."""
import numpy as np
from ._core import wrap


def PRNGKey(seed):
    return wrap(np.array([0, int(seed)], dtype=np.uint32))


key = PRNGKey


def _rng(k):
    return np.random.default_rng([int(v) for v in np.asarray(k).ravel()])


def split(k, num=2):
    r = _rng(k)
    return wrap(r.integers(0, 2**32, size=(num, 2), dtype=np.uint32))


def fold_in(k, data):
    return wrap(np.array([int(np.asarray(k).ravel()[-1]) ^ 0x9E3779B9, int(data)], dtype=np.uint32))


def uniform(k, shape=(), dtype=np.float64, minval=0.0, maxval=1.0):
    u = _rng(k).random(size=tuple(shape)).astype(dtype)
    return wrap(u * (np.asarray(maxval) - np.asarray(minval)) + np.asarray(minval))


def normal(k, shape=(), dtype=np.float64):
    return wrap(_rng(k).standard_normal(size=tuple(shape)).astype(dtype))


def permutation(k, x, axis=0, independent=False):
    if isinstance(x, (int, np.integer)):
        return wrap(_rng(k).permutation(int(x)))
    return wrap(_rng(k).permutation(np.asarray(x), axis=axis))


def choice(k, a, shape=(), replace=True, p=None, axis=0):
    return wrap(_rng(k).choice(np.asarray(a) if not isinstance(a, int) else a, size=tuple(shape), replace=replace, p=p))


def truncated_normal(k, lower, upper, shape=(), dtype=np.float64):
    r = _rng(k)
    out = r.standard_normal(size=tuple(shape))
    bad = (out < lower) | (out > upper)
    while bad.any():
        out[bad] = r.standard_normal(size=int(bad.sum()))
        bad = (out < lower) | (out > upper)
    return wrap(out.astype(dtype))
