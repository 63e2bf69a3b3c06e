"""ctypes wrapper of oracle/c/liboracle_c.so — the C/OpenMP restatement of the step path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as the full-size
checker (1M particles in seconds) and by bench.py's ``cpu_baseline`` / ``--impl
reference`` legs as the CPU baseline.  It operates IN PLACE on the numpy arrays of an
``OState`` / ``OSystem`` pair built by the numpy oracle.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "c", "liboracle_c.so")
LAW = {"spring": 0, "hertz": 1, "cundallstrack": 2}

_PTR_FIELDS_1 = ("pos_c", "pos_p_rot", "vel", "force", "torque", "ang_vel", "rad", "mass",
                 "clump_id", "mat_id", "bond_id", "fixed",
                 "dt", "box", "inv_box", "anchor", "cell_size", "gravity", "mask",
                 "young", "poisson", "e", "mu", "mu_r", "young_eff",
                 "hash", "sorted_hash", "perm", "tmp_key", "tmp_val", "pos")


class _Sys(C.Structure):
    _fields_ = ([("n", C.c_int64)]
                + [(k, C.c_int32) for k in ("dim", "periodic", "law", "M", "W", "nmat", "interact", "pad")]
                + [(k, C.c_void_p) for k in _PTR_FIELDS_1]
                + [("overflow", C.c_int32), ("pad2", C.c_int32)])


_lib = None


def build() -> None:
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(int(n)))


class CStep:
    """Binds the arrays of (state, system) once; every call works in place on them."""

    def __init__(self, state, system):
        F = np.dtype(state.fdtype)
        I = np.dtype(state.idtype)
        self.sfx = "_f32" if F == np.float32 else "_f64"
        self.state, self.system = state, system
        N, D = state.N, state.dim
        col = system.collider
        mt = system.mat_table
        self._keep = {}

        def arr(name, a, dt):
            a = np.array(a, dtype=dt, order="C", copy=True)  # never alias the caller's inputs
            self._keep[name] = a
            return a

        # State arrays are adopted (made contiguous) and written back after each call
        for k in ("pos_c", "vel", "force", "torque", "ang_vel", "rad", "mass"):
            arr(k, getattr(state, k), F)
        arr("pos_p_rot", state._pos_p_rot, F)
        arr("clump_id", state.clump_id, I)
        arr("mat_id", state.mat_id, I)
        arr("bond_id", state.bond_id, I)
        arr("fixed", state.fixed, np.uint8)
        arr("dt", np.array([system.dt]), F)
        arr("box", system.domain.box_size, F)
        arr("inv_box", system.domain.inv_box_size, F)
        arr("anchor", system.domain.anchor, F)
        arr("cell_size", np.array([col.cell_size]).reshape(1), F)
        arr("gravity", system.force_manager.gravity, F)
        arr("mask", col.neighbor_mask, I)
        for k in ("young", "poisson", "e", "mu", "mu_r", "young_eff"):
            v = getattr(mt, k, None)
            arr(k, np.zeros(1) if v is None else v, F)
        for k in ("hash", "sorted_hash", "perm", "tmp_key", "tmp_val"):
            arr(k, np.zeros(max(N, 1)), I)
        arr("pos", np.zeros((max(N, 1), D)), F)
        s = _Sys()
        s.n, s.dim, s.periodic = N, D, int(system.domain.periodic)
        s.law = LAW[system.force_model]
        s.M, s.W = col.neighbor_mask.shape[0], state.bond_id.shape[-1]
        s.nmat = int(np.asarray(self._keep["young_eff"]).shape[0]) if self._keep["young_eff"].ndim == 2 else 1
        s.interact = int(bool(system.interact_same_bond_id))
        for k in _PTR_FIELDS_1:
            setattr(s, k, self._keep[k].ctypes.data)
        self.s = s

    def _call(self, name, *extra):
        getattr(lib(), name + self.sfx)(C.byref(self.s), *extra)

    def _writeback(self):
        st, k = self.state, self._keep
        for f in ("pos_c", "vel", "force", "torque"):
            object.__setattr__(st, f, k[f])
        self.system.collider.overflow = bool(self.s.overflow)
        self.system.domain.inv_box_size = k["inv_box"]

    def partition(self):
        """-> (perm, sorted_hash, hash) of _get_spatial_partition."""
        self._call("orc_partition")
        return self._keep["perm"].copy(), self._keep["sorted_hash"].copy(), self._keep["hash"].copy()

    def compute_force(self):
        self._call("orc_partition")
        self._call("orc_celllist_force")
        self._writeback()

    def step(self, n=1):
        """n x _step_once for sphere systems, periodic / fixed box, verlet, no rotation."""
        st, sy = self.state, self.system
        assert sy.linear_integrator == "verlet" and sy.rotation_integrator == "" and sy.domain.kind == "periodic"
        assert (st.clump_id == np.arange(st.N)).all(), "orc_step_spheres: sphere systems only"
        self._call("orc_step_spheres", C.c_int64(int(n)))
        self._writeback()
