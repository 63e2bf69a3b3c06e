"""jax.lax control flow and a few primitives, eagerly on numpy."""
import numpy as np
from ._core import asarr, fresh, wrap, tree_flatten, tree_leaves, tree_map, tree_unflatten


def iota(dtype, size):
    return asarr(np.arange(size), dtype=dtype)


def stop_gradient(x):
    return x


def rsqrt(x):
    x = np.asarray(x)
    return wrap(1.0 / np.sqrt(x))


def select_n(which, *cases):
    w = np.asarray(which)
    if w.dtype == np.bool_:
        w = w.astype(np.int64)
    return wrap(np.choose(w, [np.asarray(c) for c in cases]))


def select(pred, on_true, on_false):
    return wrap(np.where(np.asarray(pred), np.asarray(on_true), np.asarray(on_false)))


def cond(pred, true_fun, false_fun, *operands, **kw):
    if "operand" in kw:  # the older single-operand spelling
        operands = (kw["operand"],)
    operands = fresh(tuple(operands))
    return true_fun(*operands) if bool(np.asarray(pred)) else false_fun(*operands)


def switch(index, branches, *operands):
    i = int(np.clip(int(np.asarray(index)), 0, len(branches) - 1))
    return branches[i](*fresh(tuple(operands)))


def while_loop(cond_fun, body_fun, init_val):
    val = fresh(init_val)
    while bool(np.asarray(cond_fun(fresh(val)))):
        val = body_fun(fresh(val))
    return val


def fori_loop(lower, upper, body_fun, init_val, **kw):
    val = fresh(init_val)
    for i in range(int(np.asarray(lower)), int(np.asarray(upper))):
        val = body_fun(asarr(i), fresh(val))
    return val


def scan(f, init, xs=None, length=None, reverse=False, unroll=1, **kw):
    if xs is None:
        n = int(length)
    else:
        ls = tree_leaves(xs)
        n = int(np.shape(ls[0])[0]) if ls else int(length)
    carry, ys = init, []
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        x = None if xs is None else tree_map(lambda a: wrap(np.asarray(a)[i]), xs)
        carry, y = f(fresh(carry), x)
        ys.append(fresh(y))
    if reverse:
        ys = ys[::-1]
    if not ys:
        return carry, None
    leaves0, td = tree_flatten(ys[0])
    cols = [tree_leaves(y) for y in ys]
    stacked = [wrap(np.stack([np.asarray(c[j]) for c in cols], axis=0)) for j in range(len(leaves0))]
    return carry, tree_unflatten(td, stacked)


def map(f, xs, **kw):
    return scan(lambda c, x: (c, f(x)), None, xs)[1]


def sort(operand, dimension=-1, is_stable=True, num_keys=1):
    if isinstance(operand, (tuple, list)):
        ops = [np.asarray(o) for o in operand]
        if ops[0].ndim != 1:
            raise NotImplementedError("multi-operand sort of 1-D operands only")
        order = np.lexsort(tuple(ops[k] for k in range(num_keys - 1, -1, -1)))  # stable, first key primary
        return tuple(wrap(o[order]) for o in ops)
    return wrap(np.sort(np.asarray(operand), axis=dimension, kind="stable"))


def top_k(x, k):
    x = np.asarray(x)
    idx = np.argsort(-x, axis=-1, kind="stable")[..., :k]
    return wrap(np.take_along_axis(x, idx, axis=-1)), wrap(idx)


def associative_scan(fn, elems, reverse=False, axis=0):
    leaves, td = tree_flatten(elems)
    n = np.shape(leaves[0])[axis]
    take = lambda i: tree_unflatten(td, [wrap(np.take(np.asarray(l), i, axis=axis)) for l in leaves])
    idx = range(n - 1, -1, -1) if reverse else range(n)
    acc, outs = None, []
    for i in idx:
        acc = take(i) if acc is None else (fn(take(i), acc) if reverse else fn(acc, take(i)))
        outs.append(acc)
    if reverse:
        outs = outs[::-1]
    cols = [tree_leaves(o) for o in outs]
    return tree_unflatten(td, [wrap(np.stack([np.asarray(c[j]) for c in cols], axis=axis))
                               for j in range(len(leaves))])


def dynamic_slice(x, start, sizes):
    x = np.asarray(x)
    sl = tuple(slice(int(np.clip(int(s), 0, x.shape[a] - n)), int(np.clip(int(s), 0, x.shape[a] - n)) + n)
               for a, (s, n) in enumerate(zip(start, sizes)))
    return wrap(x[sl])
