import numpy as np
from ._core import wrap


def _seg(ufunc, init):
    def f(data, segment_ids, num_segments=None, indices_are_sorted=False, unique_indices=False, **kw):
        data, ids = np.asarray(data), np.asarray(segment_ids)
        n = int(ids.max()) + 1 if num_segments is None else int(num_segments)
        out = np.full((n,) + data.shape[1:], init(data.dtype), dtype=data.dtype)
        ok = (ids >= 0) & (ids < n)  # out-of-range segments are dropped
        ufunc.at(out, ids[ok], data[ok])
        return wrap(out)
    return f


def _lo(dt):
    return -np.inf if dt.kind == "f" else np.iinfo(dt).min


def _hi(dt):
    return np.inf if dt.kind == "f" else np.iinfo(dt).max


segment_sum = _seg(np.add, lambda dt: 0)
segment_max = _seg(np.maximum, _lo)
segment_min = _seg(np.minimum, _hi)
segment_prod = _seg(np.multiply, lambda dt: 1)
