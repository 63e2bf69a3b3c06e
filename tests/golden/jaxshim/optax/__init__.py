"""The three optax names the reference's FIRE minimiser touches (minimizers/optimizers.py:69,299;
routines.py:357), on the numpy stand-in for JAX.  Test infrastructure for the golden generator only."""
from typing import Any, NamedTuple

import numpy as _np

import jax
import jax.numpy as jnp


class GradientTransformation(NamedTuple):
    init: Any
    update: Any


class GradientTransformationExtraArgs(GradientTransformation):
    pass


def apply_updates(params, updates):
    """optax.apply_updates: p + u leaf by leaf, result cast to the parameter's dtype."""
    return jax.tree.map(lambda p, u: None if p is None else jnp.asarray(p + u).astype(jnp.asarray(p).dtype),
                        params, updates)


def safe_norm(x, min_norm, ord=None, axis=None, keepdims=False):
    """optax.safe_norm: max(||x||, min_norm) (the gradient-safe branch selection does not change the value)."""
    n = jnp.linalg.norm(x, ord=ord, axis=axis, keepdims=keepdims)
    return jnp.where(n <= min_norm, jnp.full_like(n, min_norm), n)


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)

    def missing(*a, **k):
        raise NotImplementedError(f"optax.{name} is outside the stand-in")
    return missing
