"""numpy-backed stand-in for the small part of the JAX API that cdelv/JaxDEM's step path uses.

TEST INFRASTRUCTURE, used only by ``tests/golden/make_reference_golden.py`` in the build container: JAX is
not installed there (and there is no network), so this package lets the UNMODIFIED reference sources under
``/root/reference`` run, eagerly and slowly, on numpy: ``jit`` is the identity, ``vmap`` is a Python loop
over the mapped axis, ``lax.while_loop`` / ``scan`` / ``cond`` are Python control flow, arrays are an
``ndarray`` subclass with JAX's functional ``.at[...]`` updates and its out-of-range index rules (gathers
clamp, scatters drop).  It emulates ``jax_enable_x64=True`` (numpy's default widths).  Nothing in
``jaxdem_b200/`` or on the GPU box imports it."""

from __future__ import annotations

import dataclasses
import functools

import numpy as np

# ----------------------------------------------------------------------------- 32-bit mode
# jax_enable_x64=False (JAX's default): 64-bit dtypes do not exist, every array is truncated to its 32-bit
# counterpart.  The stand-in applies that after EVERY operation (``canon`` in ``wrap`` / ``__array_wrap__``), so
# no float64 / int64 array ever reaches the next operation.  Where numpy's promotion differs from JAX's
# (int32 op float32 -> float64 in numpy, float32 in JAX) the operation runs in float64 on operands that are
# exactly representable and is rounded once more to float32; for + - * / sqrt that equals the float32
# operation bit for bit (53 >= 2 * 24 + 2: the double rounding is innocuous).  Python scalars are weak in both
# (NEP 50).
X64 = True
_TRUNC = {np.dtype(np.float64): np.float32, np.dtype(np.int64): np.int32, np.dtype(np.uint64): np.uint32,
          np.dtype(np.complex128): np.complex64}


def set_x64(on):
    global X64
    X64 = bool(on)


def canon_dtype(dtype):
    """dtype request -> the dtype JAX would give in the current mode (None stays None)."""
    if dtype is None:
        return None
    if dtype is int:
        dtype = np.int64
    elif dtype is float:
        dtype = np.float64
    elif dtype is bool:
        dtype = np.bool_
    dt = np.dtype(dtype)
    if not X64 and dt in _TRUNC:
        return np.dtype(_TRUNC[dt])
    return dt


def canon(a):
    if not X64 and a.dtype in _TRUNC:
        return a.astype(_TRUNC[a.dtype])
    return a


# ----------------------------------------------------------------------------- arrays
def _is_int_index(k):
    return isinstance(k, (int, np.integer)) or (isinstance(k, np.ndarray) and k.dtype.kind in "iu")


def _clamp_index(arr, key):
    """JAX gather rule: out-of-range integer indices are clamped (after numpy's negative wrap)."""
    keys = key if isinstance(key, tuple) else (key,)
    out, ax = [], 0
    for k in keys:
        if k is None:
            out.append(k)
            continue
        if k is Ellipsis:
            ax = arr.ndim - (len([q for q in keys[keys.index(k) + 1:] if q is not None]))
            out.append(k)
            continue
        if _is_int_index(k):
            n = arr.shape[ax]
            kk = np.asarray(k)
            kk = np.where(kk < 0, kk + n, kk)
            out.append(np.clip(kk, 0, n - 1))
        else:
            out.append(k)
        ax += 1
    return tuple(out)


class _At:
    __slots__ = ("a", "k")

    def __init__(self, a, k=None):
        self.a, self.k = a, k

    def __getitem__(self, k):
        return _At(self.a, k)

    def _apply(self, fn, v, kw):
        kw.pop("indices_are_sorted", None), kw.pop("unique_indices", None), kw.pop("mode", None)
        out = np.array(self.a, copy=True).view(Arr)
        key = self.k
        v = np.asarray(v)
        try:
            fn(out, key, v)
        except IndexError:
            # JAX scatter rule: updates at out-of-range indices are dropped (negative indices wrap first).
            # Integer-array keys only (what the reference uses): the index arrays broadcast together, the
            # updates broadcast to that shape + the trailing axes, out-of-range entries are filtered out.
            keys = key if isinstance(key, tuple) else (key,)
            if not all(_is_int_index(k) for k in keys):
                raise
            idx = np.broadcast_arrays(*[np.asarray(k) for k in keys])
            ok = np.ones(idx[0].shape, dtype=bool)
            fixed = []
            for a, k in enumerate(idx):
                n = out.shape[a]
                k = np.where(k < 0, k + n, k)
                ok &= (k >= 0) & (k < n)
                fixed.append(k)
            vb = np.broadcast_to(v, idx[0].shape + out.shape[len(keys):])
            fn(out, tuple(k[ok] for k in fixed), vb[ok])
        return out

    def set(self, v, **kw):
        def f(o, k, v):
            np.ndarray.__setitem__(o, k, v)
        return self._apply(f, v, kw)

    def add(self, v, **kw):
        return self._apply(lambda o, k, v: np.add.at(o, k, v), v, kw)

    def multiply(self, v, **kw):
        return self._apply(lambda o, k, v: np.multiply.at(o, k, v), v, kw)

    def max(self, v, **kw):
        return self._apply(lambda o, k, v: np.maximum.at(o, k, v), v, kw)

    def min(self, v, **kw):
        return self._apply(lambda o, k, v: np.minimum.at(o, k, v), v, kw)

    def get(self, **kw):
        return self.a[self.k]


class Arr(np.ndarray):
    """ndarray with ``.at`` and JAX's gather rule; results of every operation stay ``Arr``."""

    __array_priority__ = 100.0

    @property
    def at(self):
        return _At(self)

    def __getitem__(self, key):
        try:
            r = np.ndarray.__getitem__(self, key)
        except IndexError:
            r = np.ndarray.__getitem__(self, _clamp_index(self, key))
        if isinstance(r, np.generic):
            r = np.asarray(r).view(Arr)
        return r

    def __iter__(self):  # by shape: the sequence protocol would never see an IndexError from the clamping gather
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (self[i] for i in range(self.shape[0]))

    def __setitem__(self, key, value):  # JAX arrays are immutable; the reference never does this
        raise TypeError("jax arrays are immutable; use .at[...].set()")

    def __array_wrap__(self, obj, context=None, return_scalar=False):
        return canon(np.asarray(obj)).view(Arr)

    def __hash__(self):
        raise TypeError("unhashable type: Arr")

    # JAX arrays are immutable: ``x += y`` rebinds the name to a new array and never touches aliases
    def _inplace(opname):
        def f(self, other):
            return getattr(self, opname)(other)
        return f

    for _i, _o in (("__iadd__", "__add__"), ("__isub__", "__sub__"), ("__imul__", "__mul__"),
                   ("__itruediv__", "__truediv__"), ("__ifloordiv__", "__floordiv__"), ("__imod__", "__mod__"),
                   ("__ipow__", "__pow__"), ("__iand__", "__and__"), ("__ior__", "__or__"), ("__ixor__", "__xor__"),
                   ("__ilshift__", "__lshift__"), ("__irshift__", "__rshift__"), ("__imatmul__", "__matmul__")):
        locals()[_i] = _inplace(_o)
    del _i, _o, _inplace

    def block_until_ready(self):
        return self

    def item(self, *a):
        return np.asarray(self).item(*a)

    # reductions of an ndarray subclass return numpy scalars; keep 0-d arrays like JAX
    def _red(name):
        base = getattr(np.ndarray, name)

        def f(self, *a, **k):
            if name == "astype":
                a = (canon_dtype(a[0]),) + a[1:] if a else a
                if "dtype" in k:
                    k["dtype"] = canon_dtype(k["dtype"])
            return wrap(base(np.asarray(self), *a, **k))
        f.__name__ = name
        return f

    for _n in ("sum", "prod", "max", "min", "mean", "any", "all", "argmax", "argmin", "std", "var", "dot",
               "cumsum", "astype", "reshape", "squeeze", "ravel", "transpose", "take", "clip", "round"):
        locals()[_n] = _red(_n)
    del _n, _red

    @property
    def T(self):
        return wrap(np.asarray(self).T)


def wrap(x):
    """numpy results -> Arr (recursively through tuples / lists)."""
    if isinstance(x, Arr):
        return x
    if isinstance(x, np.ndarray):
        return canon(x).view(Arr)
    if isinstance(x, (np.generic,)):
        return canon(np.asarray(x)).view(Arr)
    if isinstance(x, tuple) and not hasattr(x, "_fields"):
        return tuple(wrap(v) for v in x)
    if isinstance(x, list):
        return [wrap(v) for v in x]
    return x


def asarr(x, dtype=None):
    dtype = canon_dtype(dtype)
    if isinstance(x, Arr) and (dtype is None or x.dtype == dtype):
        return x
    return canon(np.array(x, dtype=dtype, copy=True)).view(Arr)  # never alias caller-owned numpy memory


# ----------------------------------------------------------------------------- pytrees
_REGISTRY = {}  # cls -> (data_fields, meta_fields)


def register_dataclass(cls=None, data_fields=None, meta_fields=None, drop_fields=()):
    if cls is None:
        return functools.partial(register_dataclass, data_fields=data_fields, meta_fields=meta_fields,
                                 drop_fields=drop_fields)
    if data_fields is None or meta_fields is None:
        data_fields, meta_fields = [], []
        for f in dataclasses.fields(cls):
            if not f.init:
                continue
            (meta_fields if f.metadata.get("static", False) else data_fields).append(f.name)
    _REGISTRY[cls] = (tuple(data_fields), tuple(meta_fields))
    return cls


_CUSTOM = {}  # cls -> (flatten, unflatten)


def register_pytree_node(cls, flatten, unflatten):
    _CUSTOM[cls] = (flatten, unflatten)


def register_pytree_node_class(cls):
    _CUSTOM[cls] = (lambda x: x.tree_flatten(), lambda aux, ch: cls.tree_unflatten(aux, ch))
    return cls


class _Leaf:
    def __repr__(self):
        return "*"


LEAF = _Leaf()


class TreeDef:
    def __init__(self, kind, aux, children):
        self.kind, self.aux, self.children = kind, aux, children

    @property
    def num_leaves(self):
        if self.kind == "leaf":
            return 1
        return sum(c.num_leaves for c in self.children)

    def __eq__(self, o):
        return isinstance(o, TreeDef) and self.kind == o.kind and _aux_eq(self.aux, o.aux) and \
            self.children == o.children

    def __repr__(self):
        return f"TreeDef({self.kind}, {self.children})"

    def unflatten(self, leaves):
        return tree_unflatten(self, leaves)


def _aux_eq(a, b):
    try:
        return bool(a == b)
    except Exception:
        return a is b


def _children(x):
    """(kind, aux, children) of one node, or None for a leaf."""
    t = type(x)
    if x is None:
        return ("none", None, [])
    if t in _REGISTRY:
        d, m = _REGISTRY[t]
        return ("dc", (t, d, m, tuple(getattr(x, k) for k in m)), [getattr(x, k) for k in d])
    if t in _CUSTOM:
        ch, aux = _CUSTOM[t][0](x)
        return ("custom", (t, aux), list(ch))
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return ("namedtuple", t, list(x))
    if t is tuple:
        return ("tuple", None, list(x))
    if t is list:
        return ("list", None, list(x))
    if t is dict:
        ks = sorted(x.keys())
        return ("dict", tuple(ks), [x[k] for k in ks])
    return None


def tree_flatten(tree, is_leaf=None):
    leaves = []

    def rec(x):
        if is_leaf is not None and is_leaf(x):
            leaves.append(x)
            return TreeDef("leaf", None, [])
        n = _children(x)
        if n is None:
            leaves.append(x)
            return TreeDef("leaf", None, [])
        kind, aux, ch = n
        return TreeDef(kind, aux, [rec(c) for c in ch])

    td = rec(tree)
    return leaves, td


def tree_unflatten(td, leaves):
    it = iter(leaves)

    def rec(d):
        if d.kind == "leaf":
            return next(it)
        ch = [rec(c) for c in d.children]
        if d.kind == "none":
            return None
        if d.kind == "dc":
            t, dn, mn, mv = d.aux
            return t(**dict(zip(dn, ch)), **dict(zip(mn, mv)))
        if d.kind == "custom":
            t, aux = d.aux
            return _CUSTOM[t][1](aux, ch)
        if d.kind == "namedtuple":
            return d.aux(*ch)
        if d.kind == "tuple":
            return tuple(ch)
        if d.kind == "list":
            return ch
        if d.kind == "dict":
            return dict(zip(d.aux, ch))
        raise TypeError(d.kind)

    return rec(td)


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_structure(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[1]


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, td = tree_flatten(tree, is_leaf)
    others = []
    for r in rest:
        others.append(_flatten_up_to(td, r))
    return tree_unflatten(td, [f(*xs) for xs in zip(leaves, *others)])


def _flatten_up_to(td, tree):
    """Leaves of ``tree`` at the positions of ``td``'s leaves (``tree`` may be deeper there)."""
    out = []

    def rec(d, x):
        if d.kind == "leaf":
            out.append(x)
            return
        n = _children(x)
        if n is None or len(n[2]) != len(d.children):
            raise ValueError(f"tree structure mismatch: {d} vs {type(x)}")
        for dc, c in zip(d.children, n[2]):
            rec(dc, c)

    rec(td, tree)
    return out


def _broadcast_prefix(prefix, tree, is_leaf=lambda x: x is None):
    """in_axes-style prefix tree -> one entry per leaf of ``tree``."""
    out = []

    def rec(p, x):
        if p is None or isinstance(p, (int, np.integer)):
            out.extend([p] * len(tree_leaves(x)))
            return
        pn, xn = _children(p), _children(x)
        if pn is None or xn is None or len(pn[2]) != len(xn[2]):
            raise ValueError(f"in_axes prefix mismatch: {p!r} vs {type(x)}")
        for pc, xc in zip(pn[2], xn[2]):
            rec(pc, xc)

    rec(prefix, tree)
    return out


# ----------------------------------------------------------------------------- transformations
def fresh(tree):
    """New pytree containers around the same (immutable) leaves.  JAX hands every traced function freshly
    unflattened arguments, so a body that assigns to a field of a dataclass argument (the reference's hooks do:
    ``state.pos_c = ...``) never changes the caller's object; the stand-in keeps that by rebuilding containers at
    the same boundaries (jit, vmap, scan, while_loop, fori_loop, cond)."""
    leaves, td = tree_flatten(tree)
    return tree_unflatten(td, leaves)


def jit(fun=None, **kw):
    if fun is None:
        return lambda f: jit(f, **kw)

    @functools.wraps(fun)
    def jitted(*args, **kwargs):
        return fun(*fresh(tuple(args)), **fresh(kwargs))

    return jitted


def named_call(fun=None, *, name=None):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0, **_kw):
    @functools.wraps(fun)
    def mapped(*args, **kwargs):
        axes_in = in_axes
        if isinstance(axes_in, (int, np.integer)) or axes_in is None:
            axes_in = (axes_in,) * len(args)
        axes_in = tuple(axes_in) if isinstance(axes_in, list) else axes_in
        leaves, td = tree_flatten(tuple(args))
        ax = _broadcast_prefix(tuple(axes_in), tuple(args))
        kw_leaves, kw_td = tree_flatten(kwargs)  # keyword arguments are mapped over axis 0
        n = None
        for l, a in list(zip(leaves, ax)) + [(l, 0) for l in kw_leaves]:
            if a is not None:
                n = np.shape(l)[a]
                break
        if n is None:
            raise ValueError("vmap needs at least one mapped argument")
        leaves = [asarr(l) if a is not None else l for l, a in zip(leaves, ax)]
        kw_leaves = [asarr(l) for l in kw_leaves]
        outs = []
        for i in range(n):
            li = [l if a is None else wrap(np.take(np.asarray(l), i, axis=a)) for l, a in zip(leaves, ax)]
            ki = [wrap(np.asarray(l)[i]) for l in kw_leaves]
            outs.append(fun(*tree_unflatten(td, li), **tree_unflatten(kw_td, ki)))
        if n == 0:
            raise ValueError("vmap over an empty axis is not supported by this stand-in")
        o_leaves0, o_td = tree_flatten(outs[0])
        if isinstance(out_axes, (int, np.integer)) or out_axes is None:
            oax = [out_axes] * len(o_leaves0)
        else:
            oax = _broadcast_prefix(out_axes, outs[0])
        cols = [tree_leaves(o) for o in outs]
        res = []
        for j, a in enumerate(oax):
            if a is None:
                res.append(cols[0][j])
            else:
                res.append(wrap(np.stack([np.asarray(c[j]) for c in cols], axis=a)))
        return tree_unflatten(o_td, res)

    return mapped
