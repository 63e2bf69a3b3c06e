"""numpy-backed stand-in for ``jax`` (see _core.py).  Test infrastructure for the golden generator only."""

from . import _core
from ._core import Arr as Array, jit, named_call, vmap
from . import numpy, lax, tree, tree_util, ops, random, nn, typing, debug, scipy  # noqa: F401

__version__ = "0.8.1+numpy-stand-in"


class _Config:
    jax_enable_x64 = True

    def update(self, key, value):
        if key == "jax_enable_x64":
            _core.set_x64(value)
        setattr(self, key, value)


config = _Config()


def device_get(x):
    return x


def device_put(x, *a, **k):
    return x


def block_until_ready(x):
    return x


def checkpoint(f=None, **kw):
    return f if f is not None else (lambda g: g)


remat = checkpoint


def _no_autodiff(*a, **k):
    raise NotImplementedError("automatic differentiation is outside the numpy stand-in")


grad = value_and_grad = hessian = jacfwd = jacrev = _no_autodiff


class custom_vjp:
    def __init__(self, f, **kw):
        self.f = f

    def defvjp(self, fwd, bwd, **kw):
        pass

    def __call__(self, *a, **k):
        return self.f(*a, **k)


custom_jvp = custom_vjp


def pure_callback(cb, shape, *args, **kw):
    return _core.wrap(cb(*args))


class ShapeDtypeStruct:
    def __init__(self, shape, dtype, **kw):
        self.shape, self.dtype = shape, dtype


def devices(*a):
    return ["cpu:0"]


def default_backend():
    return "cpu"
