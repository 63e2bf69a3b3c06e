"""Slab decomposition of ONE large periodic sphere system over the ranks of a
``torch.distributed`` process group (BASELINE configs 2-3 at N > 1 GPU).

The reference has no multi-device path: ``System.step`` (jaxdem/system.py:60-98) runs one
system on one device.  Here the step keeps the reference's hook order, and each rank runs
those hooks — the same CUDA entry points as the single-GPU path — on the particles it owns:

* the box is cut into slabs along the LAST axis (z in 3D, y in 2D), in units of cell
  layers of the collider's own grid (``_grid_params``, colliders/_partition.py:54-99).  The
  linear cell hash is x-fastest (``_partition.py:91-93``), so a rank's cells are one
  contiguous range of the global hash and every rank bins with the GLOBAL box: positions
  are never shifted, the minimum image (domains/periodic.py:75-79) closes the periodic
  wrap between the first and the last rank;
* after the drift (``step_before_force``) every rank classifies its particles by cell
  layer, sends the ones that left to the neighbour that now owns them (full record) and a
  copy of the ones within ``search_range`` layers of a face (ghost record: what the force
  laws read); a particle that just left stays behind as a ghost, so ONE neighbour exchange
  per step is enough;
* ``collider.compute_force`` runs on owned + ghost rows (forces on owned rows are complete,
  ghost rows are discarded), the remaining hooks on owned rows.

Classification, compaction and (un)packing are CUDA kernels (csrc/slab.cu, C ABI
``jdb200_slab_pack`` / ``jdb200_slab_unpack``): three launches before the exchange, one after.
Transport: ``torch.distributed`` point-to-point (NCCL over NVLink on GPUs; gloo in the CPU
tests), one message per direction.  Messages have fixed capacities and carry their counts, so
a step needs ONE host synchronisation (the new row counts, which are launch parameters).
Exceeding a capacity raises, in the spirit of ``Collider.overflow``.

The compute hooks and the exchange kernels are injected as an *engine* (``CudaEngine`` below
is the product and has no CPU path; the CPU tests plug numpy stand-ins — the oracle for the
hooks, a restatement of the pack / unpack kernels on the same message layout — to exercise
this file's protocol with gloo).
"""

from __future__ import annotations

import math
from typing import Any

import numpy as np
import torch
import torch.distributed as dist

from .state import Quaternion, State, int_dtype_for


def _A(dim: int) -> int:
    return 1 if dim == 2 else 3


class SlabLayout:
    """Which cell layers (last axis) each rank owns: rank r owns [bounds[r], bounds[r+1])."""

    def __init__(self, n_layers: int, world: int, search_range: int):
        if world > 1 and n_layers // world < 2 * search_range + 1:
            raise ValueError(f"{n_layers} cell layers over {world} ranks: slabs thinner than "
                             f"{2 * search_range + 1} layers cannot be served by one neighbour exchange")
        self.n_layers, self.world, self.R = n_layers, world, search_range
        self.bounds = [(r * n_layers) // world for r in range(world + 1)]
        owner = np.empty(n_layers, dtype=np.int64)
        for r in range(world):
            owner[self.bounds[r]:self.bounds[r + 1]] = r
        self.owner = owner


def cell_layer(z: torch.Tensor, anchor: float, box: float, n_layers: int) -> torch.Tensor:
    """Cell coordinate along one periodic axis, op for op what the hash kernel computes
    (colliders/cell_list.py:55-60: floor((((x - a) / B) mod 1) * g)), clamped to the grid."""
    u = (z - anchor) / box
    r = u - torch.trunc(u)
    r = torch.where((r != 0) & (r < 0), r + 1, r)
    c = torch.floor(r * float(n_layers)).to(torch.int64)
    return c.clamp_(0, n_layers - 1)


ROW_FLOAT_FIELDS = ("pos_c", "vel", "force", "ang_vel", "torque", "inertia", "q_w", "q_xyz", "rad", "mass")
GHOST_FLOAT_FIELDS = ("pos_c", "vel", "ang_vel", "rad", "mass")


def message_layout(dim: int, fbytes: int, cap_m: int, cap_g: int) -> dict:
    """Byte layout of one exchange message (csrc/slab.cu: SlabMsg): int64 header[8] — [0] full
    records, [1] ghost records, [2] strays, [7] device protocol: exchange number + 1 once the message
    is complete —, full records (floats in ROW_FLOAT_FIELDS order, then int64 gid / mat_id / fixed),
    ghost records (floats in GHOST_FLOAT_FIELDS order, then int64 gid / mat_id); sections 16-byte
    aligned, records field-major inside a section."""
    A = _A(dim)
    WF, WG = 3 * dim + 3 * A + 1 + 3 + 2, 2 * dim + A + 2
    al = lambda x: (x + 15) & ~15
    o = 64
    L = dict(WF=WF, WG=WG, cap_m=cap_m, cap_g=cap_g, mig_f=o)
    o = al(o + cap_m * WF * fbytes)
    L["mig_i"] = o
    o = al(o + cap_m * 3 * 8)
    L["gh_f"] = o
    o = al(o + cap_g * WG * fbytes)
    L["gh_i"] = o
    o = al(o + cap_g * 2 * 8)
    L["bytes"] = o
    return L


def message_views(buf: torch.Tensor, L: dict, F: torch.dtype) -> dict:
    """(records, words) views of one message.  Inside a section the records are FIELD-MAJOR (word c of
    record r at base[c * capacity + r], csrc/slab.cu): the views are transposes, index them in place."""
    fb = torch.empty((), dtype=F).element_size()
    cm, cg, WF, WG = L["cap_m"], L["cap_g"], L["WF"], L["WG"]
    return dict(
        header=buf[0:64].view(torch.int64),
        mig_f=buf[L["mig_f"]:L["mig_f"] + cm * WF * fb].view(F).view(WF, cm).t(),
        mig_i=buf[L["mig_i"]:L["mig_i"] + cm * 24].view(torch.int64).view(3, cm).t(),
        gh_f=buf[L["gh_f"]:L["gh_f"] + cg * WG * fb].view(F).view(WG, cg).t(),
        gh_i=buf[L["gh_i"]:L["gh_i"] + cg * 16].view(torch.int64).view(2, cg).t(),
    )


def key_windows(lo: int, up: int, R: int, n_layers: int, layer_cells: int, table_rows: int):
    """Rows of the dense cell table one rank can touch: the hashes of its owned + ghost rows and of
    their stencil cells lie in the layers [lo - 2R - 1, up + 2R + 1]; with the x-fastest hash that
    is one window of the table, or two where the layers wrap around the periodic box
    (``jdb200_params.key_window_*``).  None: the whole table."""
    pad = 2 * R + 1
    L0, L1 = lo - pad, up + pad + 1
    G, S = n_layers, layer_cells
    if L1 - L0 >= G:
        return None
    if L0 >= 0 and L1 <= G:
        return ((L0 * S, (L1 - L0) * S + 1), (0, 0))
    if L0 < 0:  # wraps below: layers [0, L1) and [G + L0, G)
        a_len, b_lo = L1 * S, (G + L0) * S
    else:       # wraps above: layers [0, L1 - G) and [L0, G)
        a_len, b_lo = (L1 - G) * S, L0 * S
    a_len = (a_len + 1 + 4095) // 4096 * 4096
    if a_len >= b_lo:
        return None
    return ((0, a_len), (b_lo, table_rows - b_lo))


class CudaEngine:
    """The product engine: the reference's hooks and the exchange kernels through
    libjaxdem_b200.so (no CPU path)."""

    def __init__(self, system):
        self.system = system
        self._scratch = None
        self._bound = None

    # The row buffers never move (the exchange works in place), so the C-ABI argument structs are
    # filled ONCE; per call only the row count changes.  This keeps the host side of a step at a few
    # microseconds per hook instead of rebuilding ~40 ctypes fields from Python objects every time.
    def bind(self, slab):
        import ctypes as C
        from . import _call, _lib as L
        full = slab.view(slab.cap)
        p = _call.params_for(full, self.system)
        ws = _call.workspace(p, slab.device)
        sy = self.system
        # collider + force manager + after-kick in ONE call when the configuration allows it
        # (with a rotation integrator the fused call also applies its step_after_force)
        self._fuse_after = sy.linear_integrator.native_kind == "verlet" and sy.domain.native_kind == "periodic"
        # ... and the before-force kick + drift inside the exchange's classify kernel
        self.fuse_before = self._fuse_after and slab.world > 1
        self._bound = dict(p=p, sv=_call.state_view(full), yv=_call.system_view(self.system), ws=ws, lib=L.lib(),
                           keep=full, dev=slab.device, C=C, check=L.check, stream=_call.stream_ptr)
        if slab.dev_state is not None:
            # device protocol: the hooks run over a launch bound and read the live row count themselves
            # (jdb200_state.n_rows): word 0 = owned rows, word 1 = owned + ghost rows
            for key, word in (("sv_own", 0), ("sv_local", 1)):
                v = _call.state_view(full)
                v.n_rows = slab.dev_state.data_ptr() + 8 * word
                self._bound[key] = v

    def _hook(self, name, n, ws=True, rows=None):
        """``rows``: None (n is the exact row count) or "own" / "local" (n is the launch bound, the live count is
        the slab's device-side word)."""
        b = self._bound
        C = b["C"]
        b["p"].n = int(n)
        sv = b["sv"] if rows is None else b["sv_" + rows]
        args = [b["stream"](b["dev"]), C.byref(b["p"]), C.byref(sv), C.byref(b["yv"])]
        if ws:
            args += [C.c_void_p(b["ws"].data_ptr()), C.c_size_t(b["ws"].numel())]
        b["check"](getattr(b["lib"], name)(*args), name)

    def before_force(self, state):
        sy = self.system
        if self._bound is None or sy.domain.native_kind != "periodic":
            sy.domain.apply(state, sy)
            torch.reciprocal(sy.domain.box_size, out=sy.domain.inv_box_size)
            sy.linear_integrator.step_before_force(state, sy)
            sy.rotation_integrator.step_before_force(state, sy)
            return
        n = state if isinstance(state, int) else state.N
        torch.reciprocal(sy.domain.box_size, out=sy.domain.inv_box_size)
        if getattr(self, "fuse_before", False):
            # the linear part is done by jdb200_slab_pack (SlabSystem.step passes integrate=True)
            if sy.rotation_integrator.native_kind:
                self._hook("jdb200_rotation_step_before_force", n, ws=False)
            return
        if sy.linear_integrator.native_kind:
            self._hook("jdb200_linear_step_before_force", n, ws=False)
        if sy.rotation_integrator.native_kind:
            self._hook("jdb200_rotation_step_before_force", n, ws=False)

    def compute_force(self, state):
        if self._bound is None:
            self.system.collider.compute_force(state, self.system)
            return
        self._hook("jdb200_celllist_compute_force", state if isinstance(state, int) else state.N)

    def force_and_after(self, n_local: int, n_own: int):
        """collider.compute_force on owned + ghost rows, then force manager + after-force hooks on
        the owned rows; one fused call for sphere / Verlet / no-rotation systems."""
        if self._bound is not None and self._fuse_after:
            self._hook("jdb200_celllist_force_step_after", n_local)
        else:
            self.compute_force(n_local)
            self.after_force(n_own)

    def after_force(self, state):
        sy = self.system
        if self._bound is None:
            sy.force_manager.apply(state, sy)
            sy.linear_integrator.step_after_force(state, sy)
            sy.rotation_integrator.step_after_force(state, sy)
            return
        n = state if isinstance(state, int) else state.N
        self._hook("jdb200_force_manager_apply", n)
        if sy.linear_integrator.native_kind:
            self._hook("jdb200_linear_step_after_force", n, ws=False)
        if sy.rotation_integrator.native_kind:
            self._hook("jdb200_rotation_step_after_force", n, ws=False)

    # -- exchange kernels (csrc/slab.cu) -----------------------------------------
    def _desc(self, slab, n):
        import ctypes as C
        from . import _lib as L
        sy = self.system
        d = L.SlabDesc()
        d.n, d.cap_mig, d.cap_ghost = int(n), slab.migrant_cap, slab.ghost_cap
        d.dim = slab.dim
        d.dtype = L.JDB200_F32 if slab.dtype == torch.float32 else L.JDB200_F64
        d.n_layers = slab.layout.n_layers
        d.lo_layer, d.up_layer = slab.layout.bounds[slab.rank], slab.layout.bounds[slab.rank + 1]
        d.search_range = slab.layout.R
        d.anchor, d.box_size = sy.domain.anchor.data_ptr(), sy.domain.box_size.data_ptr()
        d.cell_size = sy.collider.cell_size.data_ptr()
        d.dt = None
        return d

    @staticmethod
    def _rows(bufs):
        from . import _lib as L
        r = L.SlabRows()
        for k in L.SLAB_ROW_FIELDS:
            setattr(r, k, bufs[k].data_ptr())
        return r

    def _exchange_args(self, slab):
        """SlabDesc / SlabRows / scratch, built once per message capacity (pointers are stable)."""
        import ctypes as C
        from . import _lib as L
        key = (slab.migrant_cap, slab.ghost_cap, slab.kept.data_ptr())
        ex = getattr(self, "_ex", None)
        if ex is None or ex["key"] != key:
            lib = L.lib()
            d = self._desc(slab, slab.cap)
            need = lib.jdb200_slab_scratch_bytes(C.byref(d))
            if self._scratch is None or self._scratch.numel() < need:
                # zero once: the kernels keep their accumulators and the block ticket zeroed between exchanges
                self._scratch = torch.zeros(need, dtype=torch.uint8, device=slab.device)
            ex = self._ex = dict(key=key, d=d, rows=self._rows(slab.buf), lib=lib, C=C, check=L.check)
        return ex

    def pack(self, slab, integrate=False):
        from . import _call
        ex = self._exchange_args(slab)
        C = ex["C"]
        ex["d"].n = int(slab.n_own)
        # the before-force kick + drift of velocity Verlet rides along with the classification
        ex["d"].dt = self.system.dt.data_ptr() if integrate else None
        ex["check"](ex["lib"].jdb200_slab_pack(
            _call.stream_ptr(slab.device), C.byref(ex["d"]), C.byref(ex["rows"]), slab.send_ptr("lo"),
            slab.send_ptr("up"), slab.kept.data_ptr(), slab.holes.data_ptr(), slab.header_local.data_ptr(),
            self._scratch.data_ptr(), self._scratch.numel()), "jdb200_slab_pack")

    def pack_dev(self, slab, integrate=False):
        """Device protocol: classify / pack over the launch bound, messages stored into the neighbours'
        receive buffers of the current parity, "complete" flag behind them (jdb200_slab_pack_dev)."""
        from . import _call
        ex = self._exchange_args(slab)
        C = ex["C"]
        ex["d"].n = int(slab.bound)
        ex["d"].dt = self.system.dt.data_ptr() if integrate else None
        pr = slab.ports
        ex["check"](ex["lib"].jdb200_slab_pack_dev(
            _call.stream_ptr(slab.device), C.byref(ex["d"]), C.byref(ex["rows"]), slab.dev_state.data_ptr(),
            pr["send_lo"], pr["send_up"], slab.kept.data_ptr(), slab.holes.data_ptr(), slab.header_local.data_ptr(),
            self._scratch.data_ptr(), self._scratch.numel()), "jdb200_slab_pack_dev")

    def unpack_dev(self, slab):
        from . import _call
        ex = self._exchange_args(slab)
        C = ex["C"]
        ex["d"].n = int(slab.bound)
        pr = slab.ports
        ex["check"](ex["lib"].jdb200_slab_unpack_dev(
            _call.stream_ptr(slab.device), C.byref(ex["d"]), C.byref(ex["rows"]), slab.dev_state.data_ptr(),
            pr["recv_lo"], pr["recv_up"], slab.header_local.data_ptr(), slab.kept.data_ptr(), slab.holes.data_ptr(),
            C.c_int64(int(slab.timeout_s * 1e9))), "jdb200_slab_unpack_dev")

    def step_dev(self, slab):
        """One _step_once on the decomposed system, device protocol: no host synchronisation."""
        sy, B = self.system, slab.bound
        if not self._fuse_after:  # (the fused tail refreshes inv_box_size in its setup kernel; nothing before it reads it)
            torch.reciprocal(sy.domain.box_size, out=sy.domain.inv_box_size)
        if not self.fuse_before and sy.linear_integrator.native_kind:
            self._hook("jdb200_linear_step_before_force", B, ws=False, rows="own")
        if sy.rotation_integrator.native_kind:
            self._hook("jdb200_rotation_step_before_force", B, ws=False, rows="own")
        self.pack_dev(slab, integrate=self.fuse_before)
        self.unpack_dev(slab)
        if self._fuse_after:
            self._hook("jdb200_celllist_force_step_after", B, rows="local")
        else:
            self._hook("jdb200_celllist_compute_force", B, rows="local")
            self._hook("jdb200_force_manager_apply", B, rows="own")
            if sy.linear_integrator.native_kind:
                self._hook("jdb200_linear_step_after_force", B, ws=False, rows="own")
            if sy.rotation_integrator.native_kind:
                self._hook("jdb200_rotation_step_after_force", B, ws=False, rows="own")

    def unpack(self, slab, counts):
        from . import _call
        ex = self._exchange_args(slab)
        C = ex["C"]
        ex["d"].n = int(slab.n_own)
        arr = (C.c_int64 * 7)(*[int(c) for c in counts])
        ex["check"](ex["lib"].jdb200_slab_unpack(
            _call.stream_ptr(slab.device), C.byref(ex["d"]), C.byref(ex["rows"]), arr, slab.recv_lo.data_ptr(),
            slab.recv_up.data_ptr(), slab.kept.data_ptr(), slab.holes.data_ptr()), "jdb200_slab_unpack")


class SlabSystem:
    """Owned + ghost particles of one rank, in capacity-sized row buffers; ``view(n)`` exposes the
    first n rows as a ``State`` whose tensors alias the buffers (the C ABI works in place)."""

    def __init__(self, *, dim, dtype, device, capacity, box, anchor, n_layers, search_range, group=None,
                 ghost_capacity=None, migrant_capacity=None, transport="auto"):
        self.dim, self.dtype, self.device = dim, dtype, torch.device(device)
        self.group = group
        # "peer": the pack kernels store the messages straight into the neighbours' receive buffers
        # (symmetric memory over NVLink, one device-side barrier per step, no NCCL call);
        # "sendrecv": torch.distributed point-to-point (NCCL / gloo).  "auto": peer on CUDA when available.
        self.transport = transport
        self._symm = None
        self._parity = 0
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.layout = SlabLayout(int(n_layers), self.world, int(search_range))
        self.box_last, self.anchor_last = float(box[-1]), float(anchor[-1])
        self.cap = int(capacity)
        A, F, dev = _A(dim), dtype, self.device
        I = int_dtype_for(F)
        shapes = dict(pos_c=(dim,), vel=(dim,), force=(dim,), ang_vel=(A,), torque=(A,), inertia=(A,), q_w=(1,),
                      q_xyz=(3,), rad=(), mass=())
        self.widths = {k: int(np.prod(s)) if s else 1 for k, s in shapes.items()}

        def make_set():
            b = {k: torch.zeros((self.cap, *s), dtype=F, device=dev) for k, s in shapes.items()}
            for k in ("q_w", "rad", "mass", "inertia"):
                b[k].fill_(1)
            b["gid"] = torch.full((self.cap,), -1, dtype=torch.int64, device=dev)
            b["mat_id"] = torch.zeros(self.cap, dtype=I, device=dev)
            b["fixed"] = torch.zeros(self.cap, dtype=torch.bool, device=dev)
            return b

        self.buf = make_set()
        # static rows of a sphere system: pos_p = 0, clump_id = arange, no bonds
        self.static = dict(
            pos_p=torch.zeros((self.cap, dim), dtype=F, device=dev),
            _pos_p_rot=torch.zeros((self.cap, dim), dtype=F, device=dev),
            clump_id=torch.arange(self.cap, dtype=I, device=dev),
            bond_id=torch.full((self.cap, 1), -1, dtype=I, device=dev),
            species_id=torch.zeros(self.cap, dtype=I, device=dev),
        )
        self.n_own = 0
        self.n_ghost = 0
        self.engine: Any = None
        self.steps_done = 0
        lo, up = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        self.lo_rank, self.up_rank = lo, up
        self.header_local = torch.zeros(8, dtype=torch.int64, device=dev)
        # device protocol (peer transport only): row counts, exchange number and status bits live on the device
        # (csrc/slab.cu: kDev*), the host only looks at them every `check_every` steps, asynchronously
        self.dev_state = None
        self.bound = self.cap      # rows the kernels are launched over (>= the live owned + ghost rows)
        self.check_every = 32
        self.timeout_s = 5.0       # a neighbour's message that takes longer sets JDB200_SLAB_TIMEOUT
        self._pending = None
        self._graph = None         # (CUDAGraph of one step, the bound it was captured with), see compile_step
        self.graph_steps = False
        self._host_headers = torch.zeros((3, 8), dtype=torch.int64)
        if dev.type == "cuda":
            self._host_headers = self._host_headers.pin_memory()
        self._hh_np = self._host_headers.numpy()  # shares the (pinned) memory
        self.set_capacities(int(ghost_capacity or max(1024, self.cap // 4)),
                            int(migrant_capacity or max(256, self.cap // 16)))

    def set_capacities(self, ghost_cap: int, migrant_cap: int) -> None:
        """(Re)allocate the exchange messages: collective, every rank must use the same capacities."""
        self.ghost_cap, self.migrant_cap = int(ghost_cap), int(migrant_cap)
        self._dev_synced = False
        fb = torch.empty((), dtype=self.dtype).element_size()
        self.msg_layout = message_layout(self.dim, fb, self.migrant_cap, self.ghost_cap)
        self.kept_layout = message_layout(self.dim, fb, 0, 2 * self.migrant_cap)
        mk = lambda nbytes: torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        nb = self.msg_layout["bytes"]
        self._symm = None
        if self.world > 1 and self.device.type == "cuda" and self.transport in ("auto", "peer"):
            try:
                import torch.distributed._symmetric_memory as symm
                grp = self.group if self.group is not None else dist.group.WORLD
                # [parity][lower / upper] receive buffers; the parity alternates every exchange so a
                # neighbour never overwrites a message this rank may still be unpacking
                pool = symm.empty(4 * nb, dtype=torch.uint8, device=self.device)
                pool.zero_()
                hdl = symm.rendezvous(pool, grp.group_name)
                self._symm = dict(pool=pool, hdl=hdl, ptrs=list(hdl.buffer_ptrs), nb=nb)
            except Exception:
                if self.transport == "peer":
                    raise
                self._symm = None
        if self._symm is None:
            self.send_lo, self.send_up = mk(nb), mk(nb)
            self.recv_lo, self.recv_up = mk(nb), mk(nb)
            self.dev_state = None
        else:
            self._parity = 0
            self._select_parity()
            import ctypes as C
            sm = self._symm
            mine = sm["ptrs"][self.rank]
            arr = lambda f: (C.c_void_p * 2)(*[C.c_void_p(f(q)) for q in (0, 1)])
            # [parity] pointers: what goes down arrives in the lower neighbour's "from above" buffer and vice versa
            self.ports = dict(send_lo=arr(lambda q: sm["ptrs"][self.lo_rank] + (2 * q + 1) * nb),
                              send_up=arr(lambda q: sm["ptrs"][self.up_rank] + (2 * q) * nb),
                              recv_lo=arr(lambda q: mine + (2 * q) * nb), recv_up=arr(lambda q: mine + (2 * q + 1) * nb))
            self.dev_state = torch.zeros(16, dtype=torch.int64, device=self.device)
            self._dev_host = torch.zeros((2, 16), dtype=torch.int64).pin_memory()
            self._pending = None
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)  # every pool is zeroed (flag words 0) before anyone sends
        self.kept = mk(self.kept_layout["bytes"])
        # row indices of the leavers: [0, cap_m) downwards, [cap_m, 2 cap_m) upwards, [2 cap_m, 4 cap_m) merged
        self.holes = torch.zeros(4 * self.migrant_cap + 16, dtype=torch.int32, device=self.device)

    def _select_parity(self) -> None:
        sm, p = self._symm, self._parity
        nb = sm["nb"]
        self.recv_lo = sm["pool"][(2 * p) * nb:(2 * p + 1) * nb]
        self.recv_up = sm["pool"][(2 * p + 1) * nb:(2 * p + 2) * nb]

    def send_ptr(self, face: str) -> int:
        """Where the pack kernels write the message for the lower / upper neighbour: a local send
        buffer, or — peer transport — the neighbour's receive buffer itself (what goes down arrives
        from above, and vice versa)."""
        if self._symm is None:
            return (self.send_lo if face == "lo" else self.send_up).data_ptr()
        sm, p = self._symm, self._parity
        if face == "lo":
            return sm["ptrs"][self.lo_rank] + (2 * p + 1) * sm["nb"]
        return sm["ptrs"][self.up_rank] + (2 * p) * sm["nb"]

    def tune_capacities(self, slack: float = 2.0) -> None:
        """Shrink the messages to ``slack`` x the counts of one trial exchange (collective)."""
        if self.world == 1:
            return
        self.exchange()
        self.sync_counts()
        seen = torch.tensor(self.last_counts, dtype=torch.int64, device=self.device)
        dist.all_reduce(seen, op=dist.ReduceOp.MAX, group=self.group)
        mig, gh = int(seen[0]), int(seen[1])
        self.set_capacities(max(1024, int(slack * gh)), max(1024, int(4 * slack * mig), int(slack * gh) // 16))

    # ------------------------------------------------------------------ set-up
    def load_global(self, arrays: dict) -> None:
        """Take this rank's share of a global particle set (numpy arrays keyed like State.create:
        pos, vel, ang_vel, rad, mass, inertia, mat_id, fixed; every rank passes the same arrays)."""
        pos = np.asarray(arrays["pos"])
        n, dim = pos.shape
        F = self.dtype
        z = torch.as_tensor(pos[:, -1]).to(F)
        layer = cell_layer(z, self.anchor_last, self.box_last, self.layout.n_layers).numpy()
        mine = np.nonzero(self.layout.owner[layer] == self.rank)[0]
        m = len(mine)
        if m > self.cap:
            raise RuntimeError(f"rank {self.rank}: {m} particles exceed the capacity {self.cap}")
        A = _A(dim)
        rad = np.asarray(arrays.get("rad", np.ones(n)))
        mass = np.asarray(arrays.get("mass", np.ones(n)))
        coeff = 0.5 if dim == 2 else 0.4
        defaults = dict(
            pos_c=pos, vel=arrays.get("vel", np.zeros((n, dim))), force=np.zeros((n, dim)),
            ang_vel=arrays.get("ang_vel", np.zeros((n, A))), torque=np.zeros((n, A)),
            q_w=np.ones((n, 1)), q_xyz=np.zeros((n, 3)), rad=rad, mass=mass,
            inertia=arrays.get("inertia", (coeff * mass * rad**2)[:, None] * np.ones((1, A))),
        )
        for k, v in defaults.items():
            t = torch.as_tensor(np.asarray(v)[mine]).to(F)
            self.buf[k][:m] = t.reshape(self.buf[k][:m].shape).to(self.device)
        self.buf["gid"][:m] = torch.as_tensor(mine, dtype=torch.int64).to(self.device)
        if "mat_id" in arrays:
            self.buf["mat_id"][:m] = torch.as_tensor(np.asarray(arrays["mat_id"])[mine]).to(self.buf["mat_id"].dtype).to(self.device)
        if "fixed" in arrays:
            self.buf["fixed"][:m] = torch.as_tensor(np.asarray(arrays["fixed"])[mine], dtype=torch.bool).to(self.device)
        self.n_own, self.n_ghost = m, 0
        self._dev_synced = False

    def load_local(self, arrays: dict, gid) -> None:
        """Take rows this rank generated itself (they should lie in or next to its slab: rows that
        belong to a neighbour migrate in the first exchange) with their global ids."""
        pos = np.asarray(arrays["pos"])
        m, dim = pos.shape
        if m > self.cap:
            raise RuntimeError(f"rank {self.rank}: {m} particles exceed the capacity {self.cap}")
        A = _A(dim)
        rad = np.asarray(arrays.get("rad", np.ones(m)))
        mass = np.asarray(arrays.get("mass", np.ones(m)))
        coeff = 0.5 if dim == 2 else 0.4
        vals = dict(pos_c=pos, vel=arrays.get("vel", np.zeros((m, dim))), ang_vel=arrays.get("ang_vel", np.zeros((m, A))),
                    rad=rad, mass=mass, inertia=arrays.get("inertia", (coeff * mass * rad**2)[:, None] * np.ones((1, A))))
        for k, v in vals.items():
            t = torch.as_tensor(np.asarray(v)).to(self.dtype)
            self.buf[k][:m] = t.reshape(self.buf[k][:m].shape).to(self.device)
        for k in ("force", "torque", "q_xyz"):
            self.buf[k][:m] = 0
        self.buf["q_w"][:m] = 1
        self.buf["gid"][:m] = torch.as_tensor(np.asarray(gid), dtype=torch.int64).to(self.device)
        self.buf["mat_id"][:m] = 0
        self.buf["fixed"][:m] = False
        self.n_own, self.n_ghost = m, 0
        self._dev_synced = False

    def view(self, n: int) -> State:
        b, s = self.buf, self.static
        st = State(
            pos_c=b["pos_c"][:n], pos_p=s["pos_p"][:n], vel=b["vel"][:n], force=b["force"][:n],
            q=Quaternion(b["q_w"][:n], b["q_xyz"][:n]), ang_vel=b["ang_vel"][:n], torque=b["torque"][:n],
            rad=b["rad"][:n], _rad=b["rad"][:n], volume=b["rad"][:n], mass=b["mass"][:n], inertia=b["inertia"][:n],
            clump_id=s["clump_id"][:n], bond_id=s["bond_id"][:n], mat_id=b["mat_id"][:n],
            species_id=s["species_id"][:n], fixed=b["fixed"][:n], _pos_p_rot=s["_pos_p_rot"][:n], has_clumps=False)
        # particles that share a cell are ordered by GLOBAL id: every rank orders a cell exactly as the undecomposed
        # system does (stable sort of (hash, iota), colliders/_partition.py:91-93), so contact sums keep their order
        st.order_id = b["gid"][:n]
        return st

    # ------------------------------------------------------------------ device protocol
    @property
    def device_protocol(self) -> bool:
        return self.dev_state is not None and hasattr(self.engine, "pack_dev")

    def _push_counts(self) -> None:
        """Host row counts -> device words (start of the device protocol / after a host-side load)."""
        st = torch.zeros(16, dtype=torch.int64)
        st[0], st[1] = self.n_own, self.n_own + self.n_ghost
        st[2:4] = self.dev_state[2:4].cpu()
        self.dev_state.copy_(st.to(self.device))
        self._dev_synced = True

    _STATUS = {1: "a particle moved further than the halo in one step, or the box changed under the static slab layout",
               2: "migrants / ghosts exceed the message capacities", 4: "owned + ghost rows exceed the launch bound",
               8: "a neighbour's message did not arrive in time"}

    def _digest(self, words) -> None:
        """Look at one read-back of the device words: raise on status bits, refresh the host copies of the
        counts and grow the launch bound when the live rows come close to it."""
        st = int(words[3])
        if st:
            raise RuntimeError("slab exchange: " + "; ".join(m for b, m in self._STATUS.items() if st & b)
                               + f" (rank {self.rank}, status {st})")
        self.n_own, self.n_ghost = int(words[0]), int(words[1]) - int(words[0])
        self.last_counts = (int(words[11]), int(words[12]))
        live = int(words[1])
        target = min(self.cap, live + live // 32 + 4096)
        if live + live // 64 + 1024 > self.bound or target < self.bound - max(8192, self.bound // 16):
            self.bound = target

    def sync_counts(self) -> None:
        """Blocking read of the device-side row counts and status (gather, capacity tuning, tests)."""
        if not self.device_protocol:
            return
        self._pending = None
        self._digest(self.dev_state.cpu().tolist())

    def _poll(self) -> None:
        """Every `check_every` steps: digest the previous asynchronous read-back if it has landed and start
        the next one.  Never blocks the stream."""
        if self._pending is not None:
            slot, ev = self._pending
            if not ev.query():
                return
            self._pending = None
            self._digest(self._dev_host[slot].tolist())
        if self.steps_done % self.check_every == 0:
            slot = (self.steps_done // self.check_every) & 1
            self._dev_host[slot].copy_(self.dev_state, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._pending = (slot, ev)

    # ------------------------------------------------------------------ the exchange
    def exchange(self, integrate: bool = False) -> None:
        """Migration + halo exchange after the drift (one neighbour exchange, one host sync).
        ``integrate``: the engine's pack also applies the before-force kick + drift (fused step)."""
        if self.world == 1:
            self.n_ghost = 0
            return
        if self.device_protocol:
            if not getattr(self, "_dev_synced", False):
                self._push_counts()
            self.engine.pack_dev(self, integrate=integrate)
            self.engine.unpack_dev(self)
            return
        pack = (lambda: self.engine.pack(self, integrate=True)) if integrate else (lambda: self.engine.pack(self))
        if self._symm is not None:
            self._select_parity()
            pack()  # messages are stored straight into the neighbours' receive buffers
            self._symm["hdl"].barrier(channel=0)  # device-side: every rank's messages have landed
            self._parity ^= 1
        else:
            pack()  # leavers / halo rows -> messages, leavers' rows -> holes, counts -> headers
            # sends in (lower, upper) order, receives in (upper, lower) order: with two ranks both
            # messages travel between the same pair and are matched in posting order
            ops = [dist.P2POp(dist.isend, self.send_lo, self.lo_rank, group=self.group),
                   dist.P2POp(dist.isend, self.send_up, self.up_rank, group=self.group),
                   dist.P2POp(dist.irecv, self.recv_up, self.up_rank, group=self.group),
                   dist.P2POp(dist.irecv, self.recv_lo, self.lo_rank, group=self.group)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        # ---- the one host synchronisation of the step: the counts (they are launch parameters) ----
        hh = self._host_headers
        if self.device.type == "cuda":
            # the headers land in pinned memory; the host spins on a sentinel in the LAST one instead of
            # blocking in cudaStreamSynchronize (whose wake-up costs tens of microseconds of idle GPU)
            self._hh_np[2, 7] = -1
        hh[0].copy_(self.header_local, non_blocking=True)
        hh[1].copy_(self.recv_lo[0:64].view(torch.int64), non_blocking=True)
        hh[2].copy_(self.recv_up[0:64].view(torch.int64), non_blocking=True)
        if self.device.type == "cuda":
            flag, spins = self._hh_np, 0
            while flag[2, 7] == -1:  # header word 7 is never written by the kernels: it arrives as 0
                spins += 1
                if spins > 20_000_000:
                    torch.cuda.current_stream(self.device).synchronize()
                    if flag[2, 7] == -1:
                        raise RuntimeError("slab exchange: the headers never arrived")
        h = hh.tolist()
        k_lo, k_up, stray = h[0][1], h[0][2], h[0][3]
        a_lo, g_lo, a_up, g_up = h[1][0], h[1][1], h[2][0], h[2][1]
        if stray or h[1][2] or h[2][2]:
            raise RuntimeError("slab exchange: a particle moved further than the halo in one step, or the box "
                               "changed under the static slab layout")
        if max(k_lo, k_up, a_lo, a_up) > self.migrant_cap or max(g_lo, g_up) > self.ghost_cap:
            raise RuntimeError(f"slab exchange: {max(k_lo, k_up, a_lo, a_up)} migrants / {max(g_lo, g_up)} ghosts "
                               f"exceed the message capacities {self.migrant_cap} / {self.ghost_cap}")
        n_new = self.n_own - k_lo - k_up + a_lo + a_up
        n_gh = k_lo + k_up + g_lo + g_up
        if max(n_new, self.n_own) + n_gh > self.cap:
            raise RuntimeError(f"rank {self.rank}: {n_new} owned + {n_gh} ghost rows exceed the capacity {self.cap}")
        # owned rows repaired in place (arrivals into the holes, tail rows into what is left), ghosts behind
        self.engine.unpack(self, (self.n_own, a_lo, a_up, k_lo, k_up, g_lo, g_up))
        self.n_own, self.n_ghost = n_new, n_gh
        self.last_counts = (max(k_lo, k_up, a_lo, a_up), max(g_lo, g_up))

    # ------------------------------------------------------------------ stepping
    def step(self, n: int = 1) -> None:
        """n x _step_once (system.py:60-82) on the decomposed system."""
        eng = self.engine
        rows = (lambda k: k) if getattr(eng, "_bound", None) is not None else self.view
        fused = getattr(eng, "_bound", None) is not None
        if fused and self.world > 1 and self.device_protocol:
            if not getattr(self, "_dev_synced", False):
                self._push_counts()
            for _ in range(int(n)):
                g = self._graph
                if g is not None and g[1] == self.bound:
                    g[0].replay()
                else:
                    if g is not None:  # the launch bound moved: the captured grids are stale
                        self._graph = None
                        if self.graph_steps:
                            self.compile_step()
                            self._graph[0].replay()
                            self.steps_done += 1
                            self._poll()
                            continue
                    eng.step_dev(self)
                self.steps_done += 1
                self._poll()
            return
        for _ in range(int(n)):
            eng.before_force(rows(self.n_own))
            self.exchange(integrate=bool(getattr(eng, "fuse_before", False)))
            if fused:
                eng.force_and_after(self.n_own + self.n_ghost, self.n_own)
            else:
                eng.compute_force(rows(self.n_own + self.n_ghost))
                eng.after_force(rows(self.n_own))
            self.steps_done += 1

    def compile_step(self) -> None:
        """Capture ONE decomposed step — integrator hooks, the whole exchange, partition, pair kernel, epilogue — in
        a CUDA graph (the counterpart of ``jax.jit`` for this path; System.compile_step does the same on one GPU).
        Possible because the device protocol changes no argument from step to step: row counts, buffer parities
        and the exchange number are read on the device.  ``step`` replays it from then on (collectively: every
        rank replays the same number of steps) and re-captures when the launch bound is adjusted."""
        if not (self.device_protocol and self.world > 1 and getattr(self.engine, "_bound", None) is not None):
            raise RuntimeError("compile_step needs the device protocol (peer transport, CUDA engine, world > 1)")
        if not getattr(self, "_dev_synced", False):
            self._push_counts()
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            self.engine.step_dev(self)
        self._graph = (graph, self.bound)
        self.graph_steps = True

    def compute_force(self) -> None:
        """collider.compute_force on the decomposed system at the current positions."""
        self.exchange()
        if self.device_protocol and self.world > 1:
            if getattr(self.engine, "_bound", None) is not None:
                self.engine._hook("jdb200_celllist_compute_force", self.bound, rows="local")
                return
            self.sync_counts()
        self.engine.compute_force(self.view(self.n_own + self.n_ghost))

    # ------------------------------------------------------------------ results
    def gather(self, fields=("pos_c", "vel", "force", "torque", "ang_vel")) -> dict | None:
        """Owned rows of every rank, ordered by global particle id (all ranks get the result)."""
        self.sync_counts()
        n = self.n_own
        local = {k: self.buf[k][:n].detach().cpu() for k in (*fields, "gid")}
        if self.world == 1:
            parts = [local]
        else:
            parts = [None] * self.world
            dist.all_gather_object(parts, local, group=self.group)
        gid = torch.cat([p["gid"] for p in parts])
        order = torch.argsort(gid)
        out = {k: torch.cat([p[k] for p in parts])[order].numpy() for k in fields}
        out["gid"] = gid[order].numpy()
        return out


def create_slab_system(arrays: dict, *, box_size, anchor=None, dt=0.005, force_model_type="spring",
                       linear_integrator_type="verlet", rotation_integrator_type="verletspiral", mat_table=None,
                       gravity=None, dtype=torch.float32, device=None, group=None, capacity_factor=1.6,
                       cell_size=None, transport="auto", local_gid=None, n_total=None, rad_range=None):
    """Build the rank-local ``SlabSystem`` + the native ``System`` it drives (CUDA).  ``arrays`` is
    the GLOBAL particle set (same on every rank; numpy, keys as ``State.create``) — or, with
    ``local_gid`` (global ids of the rows), ``n_total`` and ``rad_range`` = (min, max) radius of
    the whole system, only the rows this rank generated for its own slab."""
    from . import System  # local import: jaxdem_b200.__init__ imports this module
    from .components import Collider
    from .state import default_device

    dev = torch.device(device) if device is not None else default_device()
    pos = np.asarray(arrays["pos"])
    n, dim = pos.shape
    if local_gid is not None:
        n = int(n_total)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    box = np.asarray(box_size, dtype=np.float64)
    anc = np.zeros(dim) if anchor is None else np.asarray(anchor, dtype=np.float64)
    # the collider is configured from the GLOBAL radii, identically on every rank
    rad = np.asarray(arrays.get("rad", np.ones(n)))
    rmin, rmax = (rad.min(), rad.max()) if rad_range is None else rad_range
    probe = State.create(np.zeros((2, dim)), rad=[rmin, rmax], dtype=dtype, device="cpu")
    col = Collider.create("CellList", state=probe, cell_size=cell_size)
    F = dtype
    cs = torch.as_tensor(col.cell_size, dtype=F)
    gd = torch.clamp(torch.floor(torch.as_tensor(box, dtype=F) / cs).to(torch.int64), min=1)
    col.max_cells = int(torch.prod(gd).item() + int(gd[:-1].prod().item()) * 2 + 1024)
    R = int(col.neighbor_mask.abs().max())
    if world > 1:  # only the layers this rank can touch are zeroed and scanned every step
        rank = dist.get_rank(group)
        lay = SlabLayout(int(gd[-1]), world, R)
        col.key_windows = key_windows(lay.bounds[rank], lay.bounds[rank + 1], R, int(gd[-1]),
                                      int(gd[:-1].prod().item()), col.max_cells + 1)
    cap = int(math.ceil(capacity_factor * n / world)) + 1024
    slab = SlabSystem(dim=dim, dtype=dtype, device=dev, capacity=cap, box=box, anchor=anc,
                      n_layers=int(gd[-1]), search_range=R, group=group, transport=transport)
    if local_gid is None:
        slab.load_global(arrays)
    else:
        slab.load_local(arrays, local_gid)
    system = System.create((cap, dim), dt=dt, linear_integrator_type=linear_integrator_type,
                           rotation_integrator_type=rotation_integrator_type, collider=col,
                           domain_type="periodic", domain_kw=dict(box_size=box, anchor=anc),
                           force_model_type=force_model_type, mat_table=mat_table,
                           force_manager_kw=dict(gravity=gravity), dtype=dtype, device=dev)
    slab.engine = CudaEngine(system)
    slab.system = system
    slab.tune_capacities()
    slab.engine.bind(slab)
    return slab
