"""The C/OpenMP restatement (oracle/c) against the numpy oracle: same partition bit for
bit, same forces and trajectories to rounding.  Both are CPU checkers; no GPU here."""

import numpy as np
import pytest

import oracle
from oracle import c_oracle
from oracle import colliders as ocol
from helpers import build_oracle, make_inputs


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,domain", [(3, "periodic"), (2, "periodic"), (3, "free")])
@pytest.mark.parametrize("law", ["spring", "hertz", "cundallstrack"])
def test_c_force_matches_numpy(dtype, dim, domain, law):
    inp = make_inputs(700, dim, seed=3, dtype=dtype, phi=0.6, poly=1.6, bonds=True, nmat=2)
    if domain == "free":
        inp["pos"] = inp["pos"] * 1.1 - 0.05 * inp["box"]  # some particles outside the grid
    kw = dict(dtype=dtype, domain=domain, law=law, nmat=2, lin="verlet", rot="")
    ost, osy = build_oracle(inp, **kw)
    cst, csy = build_oracle(inp, **kw)
    pos, perm, sh, nh, ovf = ocol._partition_for(ost, osy, osy.collider.cell_size)
    cs = c_oracle.CStep(cst, csy)
    cperm, csh, _ = cs.partition()
    assert np.array_equal(cperm, perm) and np.array_equal(csh, sh)
    ocol.celllist_compute_force(ost, osy)
    cs.compute_force()
    eps = np.finfo(dtype).eps
    for f in ("force", "torque"):
        a, b = getattr(cst, f), getattr(ost, f)
        assert np.abs(a - b).max() <= 64 * eps * max(1.0, np.abs(b).max()), f


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c_step_matches_numpy(dtype):
    inp = make_inputs(600, 3, seed=5, dtype=dtype, phi=0.55, fixed_frac=0.05)
    kw = dict(dtype=dtype, domain="periodic", law="spring", lin="verlet", rot="", dt=1e-3,
              gravity=[0.0, 0.0, -9.8])
    ost, osy = build_oracle(inp, **kw)
    cst, csy = build_oracle(inp, **kw)
    oracle.step(ost, osy, 5)
    c_oracle.CStep(cst, csy).step(5)
    eps = np.finfo(dtype).eps
    for f in ("pos_c", "vel", "force"):
        a, b = getattr(cst, f), getattr(ost, f)
        assert np.abs(a - b).max() <= 256 * eps * max(1.0, np.abs(b).max()), f


def test_c_threads_deterministic():
    """Thread count must not change a single bit (one thread owns one particle)."""
    inp = make_inputs(3000, 3, seed=9, dtype=np.float32, phi=0.5)
    kw = dict(dtype=np.float32, domain="periodic", law="spring", lin="verlet", rot="")
    outs = []
    for nt in (1, max(2, c_oracle.num_threads())):
        st, sy = build_oracle(inp, **kw)
        c_oracle.set_num_threads(nt)
        c_oracle.CStep(st, sy).step(3)
        outs.append((st.pos_c.copy(), st.force.copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
