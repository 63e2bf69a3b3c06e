"""Small host-side builders used by examples, tests and the benchmark (setup code,
run once; reference jaxdem/utils/grid_state.py:20-140).  numpy PCG64 RNG."""

from __future__ import annotations

import numpy as np

from .state import State


def grid_state(*, n_per_axis, spacing, radius=1.0, mass=1.0, jitter=0.0, vel_range=None,
               radius_range=None, seed=0, dtype=None, device=None) -> State:
    rng = np.random.default_rng(seed)
    dim = len(n_per_axis)
    spacing = np.broadcast_to(np.asarray(spacing, dtype=np.float64), (dim,))
    mesh = np.meshgrid(*[np.arange(n) * s for n, s in zip(n_per_axis, spacing)], indexing="ij")
    pos = np.stack([m.ravel() for m in mesh], axis=1)
    N = pos.shape[0]
    if jitter:
        pos = pos + rng.uniform(-jitter, jitter, size=pos.shape)
    vr = (-1.0, 1.0) if vel_range is None else vel_range
    vel = rng.uniform(vr[0], vr[1], size=(N, dim))
    rad = rng.uniform(*radius_range, size=N) if radius_range is not None else np.full(N, radius)
    return State.create(pos, vel=vel, rad=rad, mass=np.full(N, mass), dtype=dtype, device=device)
