import numpy as np
from ._core import wrap


def softplus(x):
    return wrap(np.logaddexp(np.asarray(x), 0.0))


def sigmoid(x):
    return wrap(1.0 / (1.0 + np.exp(-np.asarray(x))))


def relu(x):
    return wrap(np.maximum(np.asarray(x), 0))


def one_hot(x, num_classes, dtype=np.float64, axis=-1):
    return wrap((np.asarray(x)[..., None] == np.arange(num_classes)).astype(dtype))
