"""Import the reference package from /root/reference on the numpy stand-in for JAX (build container only).

Third-party packages the reference imports at package-import time but that its step path never calls
(flax / optax / orbax / distrax / vtk / h5py ...) are replaced by permissive stub modules."""

import importlib.abc
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
_STUB_ROOTS = ("flax", "orbax", "distrax", "vtk", "h5py", "tqdm", "chex", "tensorflow_probability")


class _Anything:
    """Class-like placeholder: subclassable, callable, attribute access returns more placeholders."""

    def __init__(self, *a, **k):
        pass

    def __init_subclass__(cls, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __class_getitem__(cls, item):
        return cls


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (_Anything,), {})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def import_reference():
    """-> the reference's ``jaxdem`` package.  With a real JAX installed (a maintainer's machine) it is used as is and
    the generators produce goldens from the reference on XLA; in the build image (no JAX) the numpy stand-in under
    ``jaxshim/`` takes its place."""
    if not os.path.isdir(REFERENCE):
        raise RuntimeError(f"{REFERENCE} is not present: the reference goldens are generated in the build container")
    sys.dont_write_bytecode = True  # nothing is written under the read-only reference tree
    try:
        import jax
        real = "numpy-stand-in" not in jax.__version__
    except ImportError:
        real = False
    if not real:
        shim = os.path.join(HERE, "jaxshim")
        if shim not in sys.path:
            sys.path.insert(0, shim)
        if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
            sys.meta_path.append(_StubFinder())
        import jax  # noqa: F401  (the stand-in)
    if REFERENCE not in sys.path:
        sys.path.insert(1, REFERENCE)
    import jaxdem
    return jaxdem
