import math
import numpy as np
from .._core import wrap


def gammaln(x):
    return wrap(np.vectorize(math.lgamma)(np.asarray(x, dtype=np.float64)))
