"""ctypes binding of libjaxdem_b200.so (C ABI declared in include/jaxdem_b200.h).

The product path has NO fallback: if the shared library is missing or a call
returns an error code, a RuntimeError is raised.  Nothing under ``oracle/`` is
ever imported from here.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjaxdem_b200.so")

JDB200_F32, JDB200_F64 = 0, 1
DOMAIN = {"free": 0, "periodic": 1, "reflect": 2}
LAW = {"spring": 0, "hertz": 1, "cundallstrack": 2}
LIN = {"": 0, "verlet": 1, "euler": 2}
ROT = {"": 0, "verletspiral": 1, "spiral": 2}
COLLIDER = {"": 0, "celllist": 1, "naive": 2, "neighborlist": 3, "multicelllist": 4}
GRID = {"auto": 0, "dense": 1, "sorted": 2}
# jdb200_params.promises (include/jaxdem_b200.h)
PROMISE_NO_EXT, PROMISE_NO_BONDS, PROMISE_NO_FIXED, PROMISE_NO_POS_P = 1, 2, 4, 8
ABI_VERSION = 4
FRAME_FIELDS = ("pos_c", "vel", "force", "ang_vel", "torque", "q_w", "q_xyz", "pos")  # bit f of jdb200_frame_pack's `fields`
ERRORS = {-1: "JDB200_EINVAL (bad params)", -2: "JDB200_ENULL (NULL pointer)",
          -3: "JDB200_EWORKSPACE (workspace too small)", -4: "JDB200_ECUDA (kernel launch failed)"}


class Params(C.Structure):
    _fields_ = [
        ("batch", C.c_int64), ("n", C.c_int64), ("max_cells", C.c_int64),
        ("key_window_lo", C.c_int64 * 2), ("key_window_len", C.c_int64 * 2),
        ("dim", C.c_int32), ("dtype", C.c_int32), ("domain", C.c_int32), ("law", C.c_int32),
        ("collider", C.c_int32), ("linear_integrator", C.c_int32), ("rotation_integrator", C.c_int32),
        ("stencil_m", C.c_int32), ("bond_width", C.c_int32), ("n_materials", C.c_int32),
        ("max_neighbors", C.c_int32), ("grid_mode", C.c_int32), ("clumps", C.c_int32),
        ("promises", C.c_int32),
    ]


STATE_FIELDS = ("pos_c", "pos_p", "vel", "force", "q_w", "q_xyz", "ang_vel", "torque", "inertia",
                "rad", "mass", "clump_id", "mat_id", "bond_id", "fixed", "pos_p_rot", "n_rows", "order_id")
SYSTEM_FIELDS = ("dt", "box_size", "inv_box_size", "anchor", "restitution", "cell_size",
                 "neighbor_mask", "collider_overflow", "interact_same_bond_id", "gravity",
                 "external_force", "external_force_com", "external_torque", "mat_young",
                 "mat_poisson", "mat_e", "mat_mu", "mat_mu_r", "mat_young_eff", "time", "step_count")


class StateView(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in STATE_FIELDS]


class SystemView(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in SYSTEM_FIELDS]


class NList(C.Structure):
    """jdb200_nlist: the NeighborList collider's own leaves."""
    _fields_ = [(k, C.c_void_p) for k in ("neighbor_list", "old_pos", "n_build_times", "cutoff", "skin")]


class FireStateView(C.Structure):
    """jdb200_fire_state"""
    _fields_ = [(k, C.c_void_p) for k in ("vel_pos", "vel_rot", "dt", "alpha", "n_good", "n_bad", "pe", "prev_pe",
                                          "steps", "active")]


class FireParams(C.Structure):
    """jdb200_fire_params"""
    _fields_ = [(k, C.c_double) for k in ("dt", "alpha_init", "f_inc", "f_dec", "f_alpha", "dt_max_scale",
                                          "dt_min_scale", "pe_tol", "pe_diff_tol", "force_tol")] + \
               [(k, C.c_int64) for k in ("n_min", "n_bad_max", "max_steps")]


class SlabDesc(C.Structure):
    _fields_ = [("n", C.c_int64), ("cap_mig", C.c_int64), ("cap_ghost", C.c_int64), ("dim", C.c_int32),
                ("dtype", C.c_int32), ("n_layers", C.c_int32), ("lo_layer", C.c_int32), ("up_layer", C.c_int32),
                ("search_range", C.c_int32), ("anchor", C.c_void_p), ("box_size", C.c_void_p),
                ("cell_size", C.c_void_p), ("dt", C.c_void_p)]


SLAB_ROW_FIELDS = ("pos_c", "vel", "force", "ang_vel", "torque", "inertia", "q_w", "q_xyz", "rad", "mass",
                   "mat_id", "fixed", "gid")


class SlabRows(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in SLAB_ROW_FIELDS]


_PP, _PS, _PY = C.POINTER(Params), C.POINTER(StateView), C.POINTER(SystemView)
_PD, _PR, _PN = C.POINTER(SlabDesc), C.POINTER(SlabRows), C.POINTER(NList)
_V, _SZ = C.c_void_p, C.c_size_t

# symbol -> (restype, argtypes); every symbol include/jaxdem_b200.h declares
SYMBOLS = {
    "jdb200_abi_version": (C.c_int, []),
    "jdb200_launch_count": (C.c_int64, []),
    "jdb200_workspace_bytes": (_SZ, [_PP]),
    "jdb200_celllist_partition": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, _V, _V, _V, _V]),
    "jdb200_celllist_compute_force": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ]),
    "jdb200_celllist_compute_potential_energy": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, _V]),
    "jdb200_celllist_create_neighbor_list": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, _V, _V, _V]),
    "jdb200_celllist_create_cross_neighbor_list": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, _V, C.c_int64, _V, _V, _V]),
    "jdb200_naive_compute_force": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ]),
    "jdb200_naive_compute_potential_energy": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, _V]),
    "jdb200_force_manager_apply": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ]),
    "jdb200_linear_step_before_force": (C.c_int, [_V, _PP, _PS, _PY]),
    "jdb200_linear_step_after_force": (C.c_int, [_V, _PP, _PS, _PY]),
    "jdb200_rotation_step_before_force": (C.c_int, [_V, _PP, _PS, _PY]),
    "jdb200_rotation_step_after_force": (C.c_int, [_V, _PP, _PS, _PY]),
    "jdb200_domain_apply": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ]),
    "jdb200_system_step": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, C.c_int64]),
    "jdb200_celllist_force_step_after": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ]),
    "jdb200_neighborlist_refresh": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, _PN]),
    "jdb200_neighborlist_compute_force": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, _PN]),
    "jdb200_neighborlist_compute_potential_energy": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, _PN, _V]),
    "jdb200_system_step_nl": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, C.c_int64, _PN]),
    "jdb200_minimize_fire": (C.c_int, [_V, _PP, _PS, _PY, _V, _SZ, _V, C.POINTER(FireStateView),
                                       C.POINTER(FireParams), C.c_int64, C.c_int32]),
    "jdb200_frame_pack": (C.c_int, [_V, _PP, _PS, _PY, C.c_int32, _V]),
    "jdb200_slab_message_bytes": (_SZ, [_PD]),
    "jdb200_slab_kept_bytes": (_SZ, [_PD]),
    "jdb200_slab_scratch_bytes": (_SZ, [_PD]),
    "jdb200_slab_holes_bytes": (_SZ, [_PD]),
    "jdb200_slab_pack": (C.c_int, [_V, _PD, _PR, _V, _V, _V, _V, _V, _V, _SZ]),
    "jdb200_slab_unpack": (C.c_int, [_V, _PD, _PR, C.POINTER(C.c_int64), _V, _V, _V, _V]),
    "jdb200_slab_pack_dev": (C.c_int, [_V, _PD, _PR, _V, C.POINTER(_V), C.POINTER(_V), _V, _V, _V, _V, _SZ]),
    "jdb200_slab_unpack_dev": (C.c_int, [_V, _PD, _PR, _V, C.POINTER(_V), C.POINTER(_V), _V, _V, _V, C.c_int64]),
    "jdb200_timing_enable": (C.c_int, [C.c_int]),
    "jdb200_timing_collect": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; loud failure if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C jaxdem_b200/csrc`.  jaxdem_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if a declared symbol is missing
            fn.restype, fn.argtypes = res, args
        if handle.jdb200_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH} has ABI version {handle.jdb200_abi_version()}, this package expects "
                               f"{ABI_VERSION}: rebuild it (make -C jaxdem_b200/csrc)")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {ERRORS.get(rc, rc)}")


def kernel_timing(enable: bool) -> None:
    """Diagnostic: bracket every kernel launch with CUDA events (bench.py roofline)."""
    lib().jdb200_timing_enable(1 if enable else 0)


def kernel_timing_collect(max_entries: int = 64) -> dict:
    """-> {kernel name: (total device ms, launches)} since the last collect."""
    names = C.create_string_buffer(64 * max_entries)
    ms = (C.c_double * max_entries)()
    cnt = (C.c_int64 * max_entries)()
    n = lib().jdb200_timing_collect(max_entries, names, ms, cnt)
    raw = names.raw
    return {raw[i * 64:(i + 1) * 64].split(b"\0", 1)[0].decode(): (ms[i], int(cnt[i])) for i in range(n)}
