"""Marshalling between the Python plugin objects and the C ABI structs."""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_WORKSPACES: dict = {}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda(state) -> None:
    if state.pos_c.device.type != "cuda":
        raise RuntimeError(
            "jaxdem_b200 runs on CUDA devices only (sm_100a kernels, no CPU fallback): "
            f"state is on {state.pos_c.device}."
        )


def params_for(state, system, *, max_neighbors: int = 0) -> L.Params:
    col = system.collider
    sec = getattr(col, "secondary_collider", None) or col  # NeighborList: the grid leaves are its cell list's
    batched = state.pos_c.ndim == 3
    p = L.Params()
    p.batch = state.pos_c.shape[0] if batched else 1
    p.n = state.N
    p.dim = state.dim
    p.dtype = L.JDB200_F32 if state.dtype == torch.float32 else L.JDB200_F64
    p.domain = L.DOMAIN[system.domain.native_kind]
    p.law = L.LAW[system.force_model.native_kind]
    p.collider = L.COLLIDER.get(getattr(col, "native_kind", None), 0)
    p.linear_integrator = L.LIN[system.linear_integrator.native_kind]
    p.rotation_integrator = L.ROT[system.rotation_integrator.native_kind]
    mask = getattr(sec, "neighbor_mask", None)
    p.stencil_m = 0 if mask is None else mask.shape[-2]
    p.bond_width = state.bond_id.shape[-1]
    p.n_materials = len(system.mat_table)
    p.max_neighbors = int(max_neighbors or getattr(col, "max_neighbors", 0) or 0)
    p.grid_mode = L.GRID[getattr(sec, "grid_mode", "auto")]
    p.max_cells = int(getattr(sec, "max_cells", 0) or 0)
    kw = getattr(col, "key_windows", None)  # ((lo, len), (lo, len)): rows of the dense table in use (slab.py)
    if kw:
        for w, (lo, ln) in enumerate(kw):
            p.key_window_lo[w], p.key_window_len[w] = int(lo), int(ln)
    p.clumps = 1 if state.has_clumps else 0
    p.promises = state.promise_bits()
    fm = getattr(system, "force_manager", None)
    # never baked into a CUDA graph: a replay cannot see forces added after the capture
    if (fm is not None and fm.buffers_clean() and state.pos_c.device.type == "cuda"
            and not torch.cuda.is_current_stream_capturing()):
        p.promises |= L.PROMISE_NO_EXT
    return p


def workspace(p: L.Params, device, owner=None) -> torch.Tensor:
    """Scratch for one entry-point call, owned by ``owner`` (the System): no state is shared between Systems or
    streams.  It grows on demand; an outgrown tensor is only dropped here, so whoever still references it (a
    captured CUDA graph keeps its own reference, System.compile_step) keeps it alive."""
    nbytes = L.lib().jdb200_workspace_bytes(C.byref(p))
    if nbytes == 0:
        raise RuntimeError("jdb200_workspace_bytes rejected the parameters (JDB200_EINVAL)")
    store = owner.__dict__ if owner is not None else _WORKSPACES
    key = "_workspace" if owner is not None else str(device)
    ws = store.get(key)
    if ws is None or ws.numel() < nbytes or ws.device != torch.device(device):
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        store[key] = ws
    return ws


def state_view(state) -> L.StateView:
    v = L.StateView()
    for k in ("pos_c", "pos_p", "vel", "force", "ang_vel", "torque", "inertia", "rad", "mass",
              "clump_id", "mat_id", "bond_id", "fixed"):
        t = getattr(state, k)
        if not t.is_contiguous():
            raise RuntimeError(f"State.{k} must be contiguous")
        setattr(v, k, t.data_ptr())
    v.q_w = state.q.w.data_ptr()
    v.q_xyz = state.q.xyz.data_ptr()
    v.pos_p_rot = state._pos_p_rot.data_ptr()
    nr = getattr(state, "n_rows", None)  # optional () int64 on the device: live row count, State.N is then the launch bound
    if nr is not None:
        if nr.dtype != torch.int64 or nr.numel() != 1 or nr.device != state.pos_c.device:
            raise RuntimeError("State.n_rows must be a one-element int64 tensor on the State's device")
        if state.pos_c.ndim == 3:
            raise RuntimeError("State.n_rows (ragged rows) is not available for batched States")
        v.n_rows = nr.data_ptr()
    oid = getattr(state, "order_id", None)  # optional (B,N) int64: in-cell order of the partition (jdb200_state.order_id)
    if oid is not None:
        if oid.dtype != torch.int64 or not oid.is_contiguous() or oid.device != state.pos_c.device:
            raise RuntimeError("State.order_id must be a contiguous int64 tensor on the State's device")
        v.order_id = oid.data_ptr()
    return v


def system_view(system) -> L.SystemView:
    v = L.SystemView()
    dom, fm, mt, col = system.domain, system.force_manager, system.mat_table, system.collider
    v.dt = system.dt.data_ptr()
    v.box_size = dom.box_size.data_ptr()
    v.inv_box_size = dom.inv_box_size.data_ptr()
    v.anchor = dom.anchor.data_ptr()
    rc = getattr(dom, "restitution_coefficient", None)
    v.restitution = None if rc is None else rc.data_ptr()
    sec = getattr(col, "secondary_collider", None) or col
    cs = getattr(sec, "cell_size", None)
    v.cell_size = None if cs is None else cs.data_ptr()
    nm = getattr(sec, "neighbor_mask", None)
    v.neighbor_mask = None if nm is None else nm.data_ptr()
    v.collider_overflow = col.overflow.data_ptr()
    v.interact_same_bond_id = system.interact_same_bond_id.data_ptr()
    v.gravity = fm.gravity.data_ptr()
    v.external_force = fm.external_force.data_ptr()
    v.external_force_com = fm.external_force_com.data_ptr()
    v.external_torque = fm.external_torque.data_ptr()
    for field, key in (("mat_young", "young"), ("mat_poisson", "poisson"), ("mat_e", "e"), ("mat_mu", "mu"),
                       ("mat_mu_r", "mu_r"), ("mat_young_eff", "young_eff")):
        t = getattr(mt, key, None)  # NULL for tables the materials do not define (System.create checked the law's needs)
        setattr(v, field, None if t is None else t.data_ptr())
    v.time = v.step_count = None  # the caller keeps the clock (jdb200_system_step sets them when asked)
    return v


def stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def call(name: str, state, system, *extra, needs_ws: bool = True, max_neighbors: int = 0, clock: bool = False):
    """Invoke one C-ABI entry point on the current CUDA stream (asynchronous).  ``clock``: hand
    System.time / step_count to the library (jdb200_system_step advances them on the device)."""
    require_cuda(state)
    lib = L.lib()
    p = params_for(state, system, max_neighbors=max_neighbors)
    sv, yv = state_view(state), system_view(system)
    if clock:
        yv.time, yv.step_count = system.time.data_ptr(), system.step_count.data_ptr()
    args = [stream_ptr(state.device), C.byref(p), C.byref(sv), C.byref(yv)]
    if needs_ws:
        ws = workspace(p, state.device, system)
        args += [C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel())]
    args += [(_ptr(e) if isinstance(e, torch.Tensor) or e is None else e) for e in extra]
    L.check(getattr(lib, name)(*args), name)
