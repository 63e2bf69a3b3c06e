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

Transport: ``torch.distributed`` point-to-point (NCCL over NVLink on GPUs; gloo in the CPU
tests).  Messages have fixed capacities and carry their counts, so a step needs ONE host
synchronisation (the new row counts, which are launch parameters).  Exceeding a capacity
raises, in the spirit of ``Collider.overflow``.

The compute hooks are injected as an *engine* (``CudaEngine`` below is the product; the CPU
tests plug the oracle in its place to exercise this file's host logic with gloo).
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


class CudaEngine:
    """The product engine: the reference's hooks through libjaxdem_b200.so (no CPU path)."""

    def __init__(self, system):
        self.system = system

    def before_force(self, state):
        sy = self.system
        sy.domain.apply(state, sy)
        torch.reciprocal(sy.domain.box_size, out=sy.domain.inv_box_size)
        sy.linear_integrator.step_before_force(state, sy)
        sy.rotation_integrator.step_before_force(state, sy)

    def compute_force(self, state):
        self.system.collider.compute_force(state, self.system)

    def after_force(self, state):
        sy = self.system
        sy.force_manager.apply(state, sy)
        sy.linear_integrator.step_after_force(state, sy)
        sy.rotation_integrator.step_after_force(state, sy)


class SlabSystem:
    """Owned + ghost particles of one rank, in capacity-sized buffers; ``view(n)`` exposes the
    first n rows as a ``State`` whose tensors alias the buffers (the C ABI works in place)."""

    FLOAT_FIELDS = ("pos_c", "vel", "force", "ang_vel", "torque", "q_w", "q_xyz", "rad", "mass", "inertia")
    GHOST_FLOAT_FIELDS = ("pos_c", "vel", "ang_vel", "rad", "mass")

    def __init__(self, *, dim, dtype, device, capacity, box, anchor, n_layers, search_range, group=None,
                 ghost_capacity=None, migrant_capacity=None):
        self.dim, self.dtype, self.device = dim, dtype, torch.device(device)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.layout = SlabLayout(int(n_layers), self.world, int(search_range))
        self.box_last, self.anchor_last = float(box[-1]), float(anchor[-1])
        self.cap = int(capacity)
        self.ghost_cap = int(ghost_capacity or max(1024, self.cap // 4))
        self.migrant_cap = int(migrant_capacity or max(256, self.cap // 16))
        A, F, dev = _A(dim), dtype, self.device
        I = int_dtype_for(F)
        shapes = dict(pos_c=(dim,), vel=(dim,), force=(dim,), ang_vel=(A,), torque=(A,), q_w=(1,), q_xyz=(3,),
                      rad=(), mass=(), inertia=(A,))
        self.widths = {k: int(np.prod(s)) if s else 1 for k, s in shapes.items()}
        self.buf = {k: torch.zeros((self.cap, *s), dtype=F, device=dev) for k, s in shapes.items()}
        self.buf["q_w"].fill_(1)
        self.buf["rad"].fill_(1)
        self.buf["mass"].fill_(1)
        self.buf["inertia"].fill_(1)
        self.buf["gid"] = torch.full((self.cap,), -1, dtype=torch.int64, device=dev)
        self.buf["mat_id"] = torch.zeros(self.cap, dtype=I, device=dev)
        self.buf["fixed"] = torch.zeros(self.cap, dtype=torch.bool, device=dev)
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
        self._nbrs = self._neighbours()

    # ------------------------------------------------------------------ set-up
    def _neighbours(self):
        """[(rank, faces)] with faces a subset of {"lo", "up"}; one entry per distinct neighbour."""
        if self.world == 1:
            return []
        lo, up = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        if lo == up:
            return [(lo, ("lo", "up"))]
        return [(lo, ("lo",)), (up, ("up",))]

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

    def view(self, n: int) -> State:
        b, s = self.buf, self.static
        return State(
            pos_c=b["pos_c"][:n], pos_p=s["pos_p"][:n], vel=b["vel"][:n], force=b["force"][:n],
            q=Quaternion(b["q_w"][:n], b["q_xyz"][:n]), ang_vel=b["ang_vel"][:n], torque=b["torque"][:n],
            rad=b["rad"][:n], _rad=b["rad"][:n], volume=b["rad"][:n], mass=b["mass"][:n], inertia=b["inertia"][:n],
            clump_id=s["clump_id"][:n], bond_id=s["bond_id"][:n], mat_id=b["mat_id"][:n],
            species_id=s["species_id"][:n], fixed=b["fixed"][:n], _pos_p_rot=s["_pos_p_rot"][:n], has_clumps=False)

    # ------------------------------------------------------------------ packing
    def _pack(self, idx: torch.Tensor, fields) -> torch.Tensor:
        return torch.cat([self.buf[k].index_select(0, idx).reshape(idx.numel(), self.widths[k]) for k in fields], dim=1)

    def _unpack(self, block: torch.Tensor, fields, dst: slice) -> None:
        c = 0
        for k in fields:
            w = self.widths[k]
            self.buf[k][dst] = block[:, c:c + w].reshape(self.buf[k][dst].shape)
            c += w

    def _width(self, fields) -> int:
        return sum(self.widths[k] for k in fields)

    # ------------------------------------------------------------------ the exchange
    def exchange(self) -> None:
        """Migration + halo exchange after the drift: owned rows are re-packed in place
        (holes left by leavers are filled by arrivals, then by rows from the tail), ghost rows
        are rebuilt behind them."""
        n, lay, R = self.n_own, self.layout, self.layout.R
        dev = self.device
        lo_l, up_l = lay.bounds[self.rank], lay.bounds[self.rank + 1]
        g = lay.n_layers
        layer = cell_layer(self.buf["pos_c"][:n, -1], self.anchor_last, self.box_last, g)
        owner = torch.as_tensor(lay.owner, device=dev)[layer]
        leave = owner != self.rank
        FF, GF = self.FLOAT_FIELDS, self.GHOST_FLOAT_FIELDS
        wF, wG = self._width(FF), self._width(GF)
        sends, recvs, ops = [], [], []
        kept_f, kept_i = [], []
        # distance (in layers, periodic) below the lower face / above the upper face
        below = (lo_l - layer) % g
        above = (layer - (up_l - 1)) % g
        stray = leave & (below > R) & (above > R)
        for nb, faces in self._nbrs:
            mig = leave & (owner == nb)
            halo = torch.zeros_like(leave)
            if "lo" in faces:
                halo |= ~leave & (layer < lo_l + R)
            if "up" in faces:
                halo |= ~leave & (layer >= up_l - R)
            mi, hi = torch.nonzero(mig).flatten(), torch.nonzero(halo).flatten()
            mf = torch.zeros((self.migrant_cap, wF), dtype=self.dtype, device=dev)
            mint = torch.zeros((self.migrant_cap, 3), dtype=torch.int64, device=dev)
            hf = torch.zeros((self.ghost_cap, wG), dtype=self.dtype, device=dev)
            hint = torch.zeros((self.ghost_cap, 2), dtype=torch.int64, device=dev)
            cm, ch = min(mi.numel(), self.migrant_cap), min(hi.numel(), self.ghost_cap)
            mf[:cm] = self._pack(mi[:cm], FF)
            mint[:cm] = torch.stack([self.buf["gid"][mi[:cm]], self.buf["mat_id"][mi[:cm]].long(),
                                     self.buf["fixed"][mi[:cm]].long()], dim=1)
            hf[:ch] = self._pack(hi[:ch], GF)
            hint[:ch] = torch.stack([self.buf["gid"][hi[:ch]], self.buf["mat_id"][hi[:ch]].long()], dim=1)
            cnt = torch.tensor([mi.numel(), hi.numel(), int(stray.any())], dtype=torch.int64, device=dev)
            # the leavers stay behind as ghosts (they sit within R layers of the face they crossed)
            kept_f.append(self._pack(mi[:cm], GF))
            kept_i.append(mint[:cm, :2])
            rbuf = dict(cnt=torch.zeros(3, dtype=torch.int64, device=dev), mf=torch.empty_like(mf),
                        mint=torch.empty_like(mint), hf=torch.empty_like(hf), hint=torch.empty_like(hint))
            sbuf = dict(cnt=cnt, mf=mf, mint=mint, hf=hf, hint=hint)
            sends.append(sbuf)
            recvs.append(rbuf)
            for k in ("cnt", "mf", "mint", "hf", "hint"):
                ops.append(dist.P2POp(dist.isend, sbuf[k], nb, group=self.group))
                ops.append(dist.P2POp(dist.irecv, rbuf[k], nb, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        # ---- the one host synchronisation of the step: counts ----
        counts = torch.stack([s["cnt"] for s in sends] + [r["cnt"] for r in recvs]).tolist() if sends else []
        k = len(sends)
        for (nb, _), c in zip(self._nbrs * 2, counts):
            if c[0] > self.migrant_cap or c[1] > self.ghost_cap:
                raise RuntimeError(f"slab exchange with rank {nb}: {c[0]} migrants / {c[1]} ghosts exceed the "
                                   f"capacities {self.migrant_cap} / {self.ghost_cap}")
            if c[2]:
                raise RuntimeError("a particle moved further than the halo in one step (time step too large "
                                   "for this slab decomposition)")
        # ---- owned rows: fill the holes ----
        leave_idx = torch.nonzero(leave).flatten()
        n_leave = leave_idx.numel()
        arr_f = [recvs[j]["mf"][:counts[k + j][0]] for j in range(k)]
        arr_i = [recvs[j]["mint"][:counts[k + j][0]] for j in range(k)]
        arr_f = torch.cat(arr_f) if arr_f else torch.zeros((0, wF), dtype=self.dtype, device=dev)
        arr_i = torch.cat(arr_i) if arr_i else torch.zeros((0, 3), dtype=torch.int64, device=dev)
        n_arr = arr_f.shape[0]
        new_n = n - n_leave + n_arr
        if new_n > self.cap:
            raise RuntimeError(f"rank {self.rank}: {new_n} owned particles exceed the capacity {self.cap}")
        fill = min(n_leave, n_arr)
        dst = torch.cat([leave_idx[:fill], torch.arange(n, n + n_arr - fill, device=dev)])
        self._write_rows(dst, arr_f, arr_i)
        if n_leave > n_arr:  # more leavers than arrivals: pull rows from the tail into the remaining holes
            holes = leave_idx[fill:]
            holes_low = holes[holes < new_n]
            tail = torch.arange(new_n, n, device=dev)
            tail = tail[~leave[new_n:n]]
            assert tail.numel() == holes_low.numel()
            for name in (*FF, "gid", "mat_id", "fixed"):
                self.buf[name][holes_low] = self.buf[name][tail]
        self.n_own = new_n
        # ---- ghost rows: leavers kept behind, then the neighbours' halos ----
        gf = kept_f + [recvs[j]["hf"][:counts[k + j][1]] for j in range(k)]
        gi = kept_i + [recvs[j]["hint"][:counts[k + j][1]] for j in range(k)]
        gf = torch.cat(gf) if gf else torch.zeros((0, wG), dtype=self.dtype, device=dev)
        gi = torch.cat(gi) if gi else torch.zeros((0, 2), dtype=torch.int64, device=dev)
        ng = gf.shape[0]
        if new_n + ng > self.cap:
            raise RuntimeError(f"rank {self.rank}: {new_n} owned + {ng} ghost rows exceed the capacity {self.cap}")
        sl = slice(new_n, new_n + ng)
        self._unpack(gf, GF, sl)
        self.buf["gid"][sl] = gi[:, 0]
        self.buf["mat_id"][sl] = gi[:, 1].to(self.buf["mat_id"].dtype)
        self.buf["fixed"][sl] = False
        self.n_ghost = ng

    def _write_rows(self, dst: torch.Tensor, f: torch.Tensor, i: torch.Tensor) -> None:
        c = 0
        for name in self.FLOAT_FIELDS:
            w = self.widths[name]
            self.buf[name][dst] = f[:, c:c + w].reshape(dst.numel(), *self.buf[name].shape[1:])
            c += w
        self.buf["gid"][dst] = i[:, 0]
        self.buf["mat_id"][dst] = i[:, 1].to(self.buf["mat_id"].dtype)
        self.buf["fixed"][dst] = i[:, 2].bool()

    # ------------------------------------------------------------------ stepping
    def step(self, n: int = 1) -> None:
        """n x _step_once (system.py:60-82) on the decomposed system."""
        eng = self.engine
        for _ in range(int(n)):
            eng.before_force(self.view(self.n_own))
            self.exchange()
            eng.compute_force(self.view(self.n_own + self.n_ghost))
            eng.after_force(self.view(self.n_own))
            self.steps_done += 1

    def compute_force(self) -> None:
        """collider.compute_force on the decomposed system at the current positions."""
        self.exchange()
        self.engine.compute_force(self.view(self.n_own + self.n_ghost))

    # ------------------------------------------------------------------ results
    def gather(self, fields=("pos_c", "vel", "force", "torque", "ang_vel")) -> dict | None:
        """Owned rows of every rank, ordered by global particle id (all ranks get the result)."""
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
                       cell_size=None):
    """Build the rank-local ``SlabSystem`` + the native ``System`` it drives (CUDA).  ``arrays`` is
    the GLOBAL particle set (same on every rank; numpy, keys as ``State.create``)."""
    from . import System  # local import: jaxdem_b200.__init__ imports this module
    from .components import Collider
    from .state import default_device

    dev = torch.device(device) if device is not None else default_device()
    pos = np.asarray(arrays["pos"])
    n, dim = pos.shape
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    box = np.asarray(box_size, dtype=np.float64)
    anc = np.zeros(dim) if anchor is None else np.asarray(anchor, dtype=np.float64)
    # the collider is configured from the GLOBAL radii, identically on every rank
    rad = np.asarray(arrays.get("rad", np.ones(n)))
    probe = State.create(np.zeros((2, dim)), rad=[rad.min(), rad.max()], dtype=dtype, device="cpu")
    col = Collider.create("CellList", state=probe, cell_size=cell_size)
    F = dtype
    cs = torch.as_tensor(col.cell_size, dtype=F)
    gd = torch.clamp(torch.floor(torch.as_tensor(box, dtype=F) / cs).to(torch.int64), min=1)
    col.max_cells = int(torch.prod(gd).item() + int(gd[:-1].prod().item()) * 2 + 1024)
    R = int(col.neighbor_mask.abs().max())
    cap = int(math.ceil(capacity_factor * n / world)) + 1024
    slab = SlabSystem(dim=dim, dtype=dtype, device=dev, capacity=cap, box=box, anchor=anc,
                      n_layers=int(gd[-1]), search_range=R, group=group)
    slab.load_global(arrays)
    system = System.create((cap, dim), dt=dt, linear_integrator_type=linear_integrator_type,
                           rotation_integrator_type=rotation_integrator_type, collider=col,
                           domain_type="periodic", domain_kw=dict(box_size=box, anchor=anc),
                           force_model_type=force_model_type, mat_table=mat_table,
                           force_manager_kw=dict(gravity=gravity), dtype=dtype, device=dev)
    slab.engine = CudaEngine(system)
    slab.system = system
    return slab
