"""numpy restatement of MaterialTable + matchmakers.  Oracle only.

jaxdem/materials/material_table.py:77-128,
jaxdem/material_matchmakers/harmonic.py:29-39, linear.py:28-32.
"""

from __future__ import annotations

import numpy as np

PROPS = ("density", "young", "poisson", "mu", "e", "mu_r")


def harmonic(a, b):
    is_zero = (a == 0.0) | (b == 0.0)
    s1 = np.where(is_zero, 1.0, a)
    s2 = np.where(is_zero, 1.0, b)
    return np.where(is_zero, 0.0, 2.0 * s1 * s2 / (s1 + s2))


def linear(a, b):
    return (a + b) / 2


class MaterialTable:
    """props: (M,) arrays; pair: (M, M) ``*_eff`` arrays."""

    def __init__(self, props: dict, pair: dict, matcher: str):
        self.props, self.pair, self.matcher = props, pair, matcher

    def __getattr__(self, item):
        if item in ("props", "pair", "matcher") or item.startswith("__"):
            raise AttributeError(item)
        if item in self.props:
            return self.props[item]
        if item in self.pair:
            return self.pair[item]
        raise AttributeError(item)

    def astype(self, dtype) -> "MaterialTable":
        return MaterialTable(
            {k: v.astype(dtype) for k, v in self.props.items()},
            {k: v.astype(dtype) for k, v in self.pair.items()},
            self.matcher,
        )


def make_material_table(mats, matcher: str = "harmonic", fill: float = 0.0, dtype=np.float64):
    """mats: list of dicts, e.g. ``{"young": 1e4, "poisson": 0.3, "density": 0.27}``.
    Missing properties take ``fill`` (material_table.py:111-116)."""
    keys = sorted({k for m in mats for k in m}) if mats else []
    # every table carries the full elastic-friction key set so all laws can run
    keys = sorted(set(keys) | set(PROPS))
    props = {k: np.asarray([m.get(k, fill) for m in mats], dtype=np.float64) for k in keys}
    fn = {"harmonic": harmonic, "linear": linear}[matcher]
    pair = {f"{k}_eff": fn(a[:, None], a[None, :]) for k, a in props.items()}
    return MaterialTable(props, pair, matcher).astype(dtype)
