"""Plugin components mirroring JaxDEM's Factory roots for the step path:
Domain, ForceModel, Collider, LinearIntegrator, RotationIntegrator, ForceManager.

Every hook has the reference's name and argument meaning, ``(state, system) ->
(state, system)``; the work is done by one C-ABI entry point of
libjaxdem_b200.so on the current CUDA stream.  Buffers are updated IN PLACE and
the same objects are returned (the reference returns fresh pytrees).
"""

from __future__ import annotations

import math
from typing import Any

import torch

from . import _call, pair_laws
from .factory import Factory
from .state import State, int_dtype_for


def _leaf(x, dtype, device, batch, shape):
    """System leaf as a dense tensor with an optional leading batch axis."""
    t = torch.as_tensor(x, dtype=dtype).to(device)
    full = (batch, *shape) if batch is not None else tuple(shape)
    return t.expand(full).contiguous().clone()


# ---------------------------------------------------------------------------
# Domain (reference jaxdem/domains/__init__.py:26-215)
# ---------------------------------------------------------------------------
class Domain(Factory):
    native_kind = "free"

    def __init__(self, box_size, inv_box_size, anchor, **kw: Any):
        self.box_size, self.inv_box_size, self.anchor = box_size, inv_box_size, anchor
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def periodic(self) -> bool:
        return False

    @classmethod
    def Create(cls, dim: int, box_size=None, anchor=None, *, dtype=torch.float32, device="cpu",
               batch=None, **kw: Any):
        box = torch.ones(dim) if box_size is None else torch.as_tensor(box_size, dtype=torch.float64)
        if box.shape[-1:] != (dim,):
            raise ValueError(f"box_size must have shape ({dim},), got shape {tuple(box.shape)}.")
        anc = torch.zeros(dim) if anchor is None else torch.as_tensor(anchor, dtype=torch.float64)
        if anc.shape[-1:] != (dim,):
            raise ValueError(f"anchor must have shape ({dim},), got shape {tuple(anc.shape)}.")
        box = _leaf(box, dtype, device, batch, (dim,))
        anc = _leaf(anc, dtype, device, batch, (dim,))
        extra = {k: _leaf(v, dtype, device, batch, ()) for k, v in kw.items()}
        return cls(box_size=box, inv_box_size=1.0 / box, anchor=anc, **extra)

    @staticmethod
    def displacement(ri, rj, system):
        return ri - rj

    @staticmethod
    def apply(state: State, system: "Any"):
        """Domain.apply -> jdb200_domain_apply (free: bbox; reflect: impulses; periodic: no-op)."""
        if system.domain.native_kind != "periodic":
            _call.call("jdb200_domain_apply", state, system)
        return state, system

    @staticmethod
    def shift(state: State, system: "Any"):
        return state, system


@Domain.register("free")
class FreeDomain(Domain):
    """reference jaxdem/domains/free.py:42-64."""
    native_kind = "free"


@Domain.register("periodic")
class PeriodicDomain(Domain):
    """reference jaxdem/domains/periodic.py:31-116."""
    native_kind = "periodic"

    @property
    def periodic(self) -> bool:
        return True

    @staticmethod
    def displacement(ri, rj, system):
        rij = ri - rj
        return rij - system.domain.box_size * torch.round(rij / system.domain.box_size)

    @staticmethod
    def shift(state, system):
        """Wrap pos_c into the primary box (periodic.py:84-116); writer-side, not on the step path."""
        box, anc = system.domain.box_size[..., None, :], system.domain.anchor[..., None, :]
        state.pos_c -= box * torch.floor((state.pos_c - anc) / box)
        return state, system


@Domain.register("reflect")
class ReflectDomain(Domain):
    """reference jaxdem/domains/reflect.py:99-299."""
    native_kind = "reflect"

    @classmethod
    def Create(cls, dim, box_size=None, anchor=None, restitution_coefficient=1.0, **kw):
        return super().Create(dim, box_size=box_size, anchor=anchor,
                              restitution_coefficient=restitution_coefficient, **kw)


# ---------------------------------------------------------------------------
# ForceModel (reference jaxdem/forces/__init__.py:55-150): the law is selected by
# name; its arithmetic lives in csrc/laws.cuh.
# ---------------------------------------------------------------------------
class ForceModel(Factory):
    """``force(i, j, pos, state, system) -> (force, torque)`` on i due to j and ``energy(i, j, pos, state,
    system)`` are the reference's per-pair contract (forces/__init__.py:55-150), here in torch
    (jaxdem_b200/pair_laws.py) for code that calls a law directly; the colliders evaluate the same formulas
    inside their CUDA kernels."""
    native_kind = "spring"
    required_material_properties: tuple = ()
    requires_history = False

    @staticmethod
    def force(i, j, pos, state, system):
        raise NotImplementedError

    @staticmethod
    def energy(i, j, pos, state, system):
        raise NotImplementedError

    def init_history(self, shape):
        return None

    @staticmethod
    def force_and_history(i, j, pos, state, system, history):
        f, t = system.force_model.force(i, j, pos, state, system)
        return f, t, history


@ForceModel.register("spring")
class SpringForce(ForceModel):
    """reference jaxdem/forces/spring.py:69-147."""
    native_kind = "spring"
    required_material_properties = ("young_eff",)
    force = staticmethod(pair_laws.spring_force)
    energy = staticmethod(pair_laws.spring_energy)


@ForceModel.register("hertz")
class HertzianForce(ForceModel):
    """reference jaxdem/forces/hertz.py:69-162."""
    native_kind = "hertz"
    required_material_properties = ("young", "poisson")
    force = staticmethod(pair_laws.hertz_force)
    energy = staticmethod(pair_laws.hertz_energy)


@ForceModel.register("cundallstrack")
class CundallStrackForce(ForceModel):
    """reference jaxdem/forces/cundall_strack.py:99-235."""
    native_kind = "cundallstrack"
    required_material_properties = ("young", "poisson", "e", "mu", "mu_r")
    force = staticmethod(pair_laws.cundallstrack_force)
    energy = staticmethod(pair_laws.cundallstrack_energy)


# ---------------------------------------------------------------------------
# Collider (reference jaxdem/colliders/__init__.py:22-222)
# ---------------------------------------------------------------------------
class Collider(Factory):
    native_kind = ""

    def __init__(self, overflow=None):
        self.overflow = overflow

    def _bind(self, dtype, device, batch):
        """Allocate device leaves once the owning System knows dtype/device/batch."""
        shape = (batch,) if batch is not None else ()
        self.overflow = torch.zeros(shape, dtype=torch.bool, device=device)
        return self

    @staticmethod
    def compute_force(state, system):
        """No-op collider "": zero force and torque (colliders/__init__.py:56-88)."""
        state.force.zero_()
        state.torque.zero_()
        return state, system

    @staticmethod
    def compute_potential_energy(state, system):
        shape = state.pos_c.shape[:-2]
        return state, system, torch.zeros(shape, dtype=state.dtype, device=state.device)


Collider.register("")(Collider)


@Collider.register("naive")
class NaiveSimulator(Collider):
    """O(N^2) collider, the reference's default (jaxdem/colliders/naive.py:73-235)."""
    native_kind = "naive"

    @staticmethod
    def compute_force(state, system):
        _call.call("jdb200_naive_compute_force", state, system)
        return state, system

    @staticmethod
    def compute_potential_energy(state, system):
        e = torch.empty(state.pos_c.shape[:-2], dtype=state.dtype, device=state.device)
        _call.call("jdb200_naive_compute_potential_energy", state, system, e)
        return state, system, e


@Collider.register("CellList")
class DynamicCellList(Collider):
    """Cell-list collider (reference jaxdem/colliders/cell_list.py:264-595).

    Extra static knobs of this implementation: ``max_cells`` (capacity of the dense
    cell table per system, default 4 N + 1024) and ``grid_mode`` ("auto" | "dense" |
    "sorted", see include/jaxdem_b200.h)."""
    native_kind = "celllist"

    def __init__(self, neighbor_mask, cell_size, max_cells=None, grid_mode="auto", overflow=None):
        super().__init__(overflow)
        self.neighbor_mask, self.cell_size = neighbor_mask, cell_size
        self.max_cells, self.grid_mode = max_cells, grid_mode

    @classmethod
    def Create(cls, state: State, cell_size=None, search_range=None, box_size=None, max_cells=None,
               grid_mode="auto"):
        """DynamicCellList.Create (cell_list.py:374-432); host-side, once."""
        F = state.dtype
        rad = state._rad.detach().cpu().to(F)
        min_rad, max_rad = rad.min(), rad.max()
        alpha = max_rad / min_rad
        if cell_size is None:
            cell_size = 2.0 * max_rad if bool(alpha < 2.5) else 0.5 * max_rad
        cell_size = torch.as_tensor(cell_size, dtype=F)
        if box_size is not None:
            box = torch.as_tensor(box_size, dtype=F)
            for _ in range(2):
                sr = max(1, int(torch.ceil(2 * max_rad / cell_size))) if search_range is None else int(search_range)
                gd = torch.clamp(torch.floor(box / cell_size).to(torch.int64), min=2 * sr + 1)
                cell_size = torch.min(box / gd.to(F))
        if search_range is None:
            search_range = max(1, int(torch.ceil(2 * max_rad / cell_size)))
        r = torch.arange(-int(search_range), int(search_range) + 1)
        mesh = torch.meshgrid(*([r] * state.dim), indexing="ij")
        mask = torch.stack([m.reshape(-1) for m in mesh], dim=1).to(int_dtype_for(F))
        if max_cells is None:
            max_cells = 4 * state.N + 1024
        return cls(neighbor_mask=mask, cell_size=cell_size, max_cells=int(max_cells), grid_mode=grid_mode)

    def _bind(self, dtype, device, batch):
        super()._bind(dtype, device, batch)
        I = int_dtype_for(dtype)
        self.cell_size = _leaf(self.cell_size, dtype, device, batch, ())
        m = torch.as_tensor(self.neighbor_mask)
        self.neighbor_mask = _leaf(m, I, device, batch, tuple(m.shape[-2:]))
        return self

    @staticmethod
    def compute_force(state, system):
        """-> jdb200_celllist_compute_force (cell_list.py:434-464)."""
        _call.call("jdb200_celllist_compute_force", state, system)
        return state, system

    @staticmethod
    def compute_potential_energy(state, system):
        """-> jdb200_celllist_compute_potential_energy (cell_list.py:466-496)."""
        e = torch.empty(state.pos_c.shape[:-2], dtype=state.dtype, device=state.device)
        _call.call("jdb200_celllist_compute_potential_energy", state, system, e)
        return state, system, e

    @staticmethod
    def create_neighbor_list(state, system, cutoff, max_neighbors: int):
        """-> jdb200_celllist_create_neighbor_list (cell_list.py:498-595).
        Returns (state, system, (.., N, K) int list padded with -1, overflow flag)."""
        lead = state.pos_c.shape[:-2]
        I = int_dtype_for(state.dtype)
        nl = torch.empty((*lead, state.N, max_neighbors), dtype=I, device=state.device)
        ovf = torch.zeros(lead, dtype=torch.bool, device=state.device)
        cut = torch.as_tensor(cutoff, dtype=state.dtype).to(state.device).expand(lead).contiguous()
        if max_neighbors > 0 and state.N > 0:
            _call.call("jdb200_celllist_create_neighbor_list", state, system, cut, nl, ovf,
                       max_neighbors=max_neighbors)
        return state, system, nl, ovf

    @staticmethod
    def create_cross_neighbor_list(pos_a, pos_b, system, cutoff, max_neighbors: int):
        """-> jdb200_celllist_create_cross_neighbor_list (cell_list.py:600-715): for every query
        point of ``pos_a`` the points of ``pos_b`` within ``cutoff``.  Returns ((.., N_A, K) int
        list of indices into pos_b, padded with -1; overflow flag)."""
        import ctypes as C
        pos_a, pos_b = pos_a.contiguous(), pos_b.contiguous()
        lead = pos_b.shape[:-2]
        n_a, n_b, F = pos_a.shape[-2], pos_b.shape[-2], pos_b.dtype
        I = int_dtype_for(F)
        dev = pos_b.device
        nl = torch.full((*lead, n_a, max_neighbors), -1, dtype=I, device=dev)
        ovf = torch.zeros(lead, dtype=torch.bool, device=dev)
        if n_a == 0 or n_b == 0 or max_neighbors == 0:
            return nl, ovf
        z = lambda *s, dt=F: torch.zeros((*lead, *s), dtype=dt, device=dev)
        from .state import Quaternion
        A = 1 if pos_b.shape[-1] == 2 else 3
        zi = z(n_b, dt=I)
        db = State(pos_c=pos_b, pos_p=z(n_b, pos_b.shape[-1]), vel=pos_b, force=pos_b, q=Quaternion(z(n_b, 1), z(n_b, 3)),
                   ang_vel=z(n_b, A), torque=z(n_b, A), rad=z(n_b), _rad=z(n_b), volume=z(n_b), mass=z(n_b),
                   inertia=z(n_b, A), clump_id=zi, bond_id=torch.full((*lead, n_b, 1), -1, dtype=I, device=dev),
                   mat_id=zi, species_id=zi, fixed=z(n_b, dt=torch.bool), _pos_p_rot=z(n_b, pos_b.shape[-1]),
                   has_clumps=False)
        cut = torch.as_tensor(cutoff, dtype=F).to(dev).expand(lead).contiguous()
        _call.call("jdb200_celllist_create_cross_neighbor_list", db, system, pos_a, C.c_int64(n_a), cut, nl, ovf,
                   max_neighbors=max_neighbors)
        return nl, ovf

    @staticmethod
    def partition(state, system):
        """Cell permutation, sorted hashes, de-duplicated neighbour-cell hashes and the
        strategy used (-> jdb200_celllist_partition; _get_spatial_partition,
        cell_list.py:35-96)."""
        lead = state.pos_c.shape[:-2]
        I = int_dtype_for(state.dtype)
        M = system.collider.neighbor_mask.shape[-2]
        perm = torch.empty((*lead, state.N), dtype=I, device=state.device)
        sh = torch.empty_like(perm)
        nh = torch.empty((*lead, state.N, M), dtype=I, device=state.device)
        dense = torch.zeros(lead, dtype=torch.bool, device=state.device)
        _call.call("jdb200_celllist_partition", state, system, perm, sh, nh, dense)
        return perm, sh, nh, dense


    @staticmethod
    def table_strategy(state, system):
        """How force / energy / neighbour-list calls would address the cells of this State right now:
        0 = sorted keys + binary search (radix sort), 1 = dense cell table, 2 = hashed cell table (grid larger than
        the table; csrc/common.cuh GridInfo.hashed).  Per system of a batch."""
        lead = state.pos_c.shape[:-2]
        out = torch.zeros(lead, dtype=torch.uint8, device=state.device)
        _call.call("jdb200_celllist_partition", state, system, None, None, None, out)
        return out


Collider.register("b200celllist")(DynamicCellList)


@Collider.register("MultiCellList")
class DynamicMultiCellList(DynamicCellList):
    """Loose-grid collider (reference jaxdem/colliders/multi_cell_list.py:283-702): the cell list's partition plus
    an expandable AABB per occupied cell (segmented min / max of the members' boxes over the cell's sorted run);
    a stencil cell whose AABB does not reach the query box is skipped before its run is walked.  The prune only
    drops cells without contacts, so forces, energies and neighbour lists equal DynamicCellList's.  Same C-ABI entry
    points as the cell list with ``collider = JDB200_COLLIDER_MULTICELLLIST``."""
    native_kind = "multicelllist"

    @classmethod
    def Create(cls, state: State, cell_size=None, search_range=None, box_size=None, max_hashes=None, max_cells=None,
               grid_mode="auto"):
        """DynamicMultiCellList.Create (multi_cell_list.py:330-392): the cell defaults to 2 r_max whatever the
        polydispersity; ``max_hashes`` is the reference's deprecated no-op."""
        del max_hashes
        if cell_size is None:
            cell_size = 2.0 * state._rad.detach().cpu().to(state.dtype).max()
        return super().Create(state, cell_size=cell_size, search_range=search_range, box_size=box_size,
                              max_cells=max_cells, grid_mode=grid_mode)


Collider.register("b200multicelllist")(DynamicMultiCellList)


@Collider.register("NeighborList")
class NeighborList(Collider):
    """Verlet neighbour-list collider (reference jaxdem/colliders/neighbor_list.py:133-776): a cached
    ``(N, max_neighbors)`` list built by the secondary collider with radius ``cutoff + skin`` and rebuilt when a
    particle has moved further than ``skin / 2`` since the last build.  The decision is taken on the device inside
    the same call that walks the list (csrc/nlist.cu): no host round trip, legal in CUDA-graph capture, and under a
    batch axis each system rebuilds only when IT must."""
    native_kind = "neighborlist"

    def __init__(self, secondary_collider, neighbor_list, old_pos, n_build_times, cutoff, skin, max_neighbors,
                 overflow=None):
        super().__init__(overflow)
        self.secondary_collider = secondary_collider
        self.neighbor_list, self.old_pos, self.n_build_times = neighbor_list, old_pos, n_build_times
        self.cutoff, self.skin, self.max_neighbors = cutoff, skin, int(max_neighbors)

    @classmethod
    def Create(cls, state: State, cutoff, skin=None, skin_fraction=None, max_neighbors=None, number_density=1.0,
               safety_factor=1.2, secondary_collider_type="CellList", secondary_collider_kw=None):
        """NeighborList.Create (neighbor_list.py:286-402); host-side, once.  ``skin`` is the absolute buffer
        distance, ``skin_fraction`` the same as a fraction of ``cutoff`` (default 0.05)."""
        import warnings
        from .factory import _normalize_key
        if skin is not None and skin_fraction is not None:
            raise ValueError("Pass either `skin` (absolute distance) or `skin_fraction` (fraction of the cutoff), "
                             "not both.")
        cutoff = float(cutoff)
        skin_val = (0.05 if skin_fraction is None else float(skin_fraction)) * cutoff if skin is None else float(skin)
        reach = cutoff + skin_val
        dim, N = state.dim, state.N
        rad = state._rad.detach().double().cpu()
        pos = state.pos.detach().double().cpu().reshape(-1, N, dim)[0]  # batched: sized from the first system
        # buffer size: density estimate vs typical packing, capped by the densest packing and by N (:346-392)
        extent = torch.clamp(pos.amax(0) - pos.amin(0) + 2.0 * rad.max(), min=1.0)
        density = max(float(number_density), float(pos.shape[0] / extent.prod()))
        shell = lambda r: ((reach + 0.9 * r) / (0.9 * r)) ** dim
        hard_cap = int(math.ceil((0.91 if dim == 2 else 0.74) * shell(float(rad.min()))))
        typical = int(math.ceil(shell(float(rad.mean()))))
        asked = max_neighbors
        if max_neighbors is None:
            ball = math.pi * reach ** dim * (1.0 if dim == 2 else 4.0 / 3.0)
            max_neighbors = max(int(math.ceil(safety_factor * ball * density)), typical)
        K = max(min(int(max_neighbors), hard_cap, N), 0)
        if asked is not None and K < int(asked):
            warnings.warn(f"NeighborList max_neighbors={asked} clamped to {K} (bounded by N={N} and the physical "
                          f"packing limit of {hard_cap} neighbors within the search radius).", stacklevel=2)
        key = _normalize_key(secondary_collider_type)
        if key not in ("celllist", "b200celllist"):
            raise NotImplementedError("jaxdem_b200's NeighborList rebuilds through the cell-list collider only "
                                      f"(secondary_collider_type={secondary_collider_type!r})")
        kw = dict(secondary_collider_kw or {})
        kw["state"] = state
        kw.setdefault("cell_size", reach)
        sec = Collider.create(secondary_collider_type, **kw)
        return cls(secondary_collider=sec, neighbor_list=None, old_pos=state.pos.detach().clone(),
                   n_build_times=None, cutoff=cutoff, skin=skin_val, max_neighbors=K)

    def _bind(self, dtype, device, batch):
        super()._bind(dtype, device, batch)
        self.secondary_collider._bind(dtype, device, batch)
        I = int_dtype_for(dtype)
        self.old_pos = self.old_pos.to(device=device, dtype=dtype).contiguous()
        lead = self.old_pos.shape[:-1]
        self.neighbor_list = torch.full((*lead, self.max_neighbors), -1, dtype=I, device=device)
        self.n_build_times = torch.zeros((batch,) if batch is not None else (), dtype=I, device=device)
        self.cutoff = _leaf(self.cutoff, dtype, device, batch, ())
        self.skin = _leaf(self.skin, dtype, device, batch, ())
        return self

    def _nlist_view(self):
        import ctypes as C
        from . import _lib
        v = _lib.NList()
        v.neighbor_list = self.neighbor_list.data_ptr() if self.neighbor_list.numel() else None
        v.old_pos, v.n_build_times = self.old_pos.data_ptr(), self.n_build_times.data_ptr()
        v.cutoff, v.skin = self.cutoff.data_ptr(), self.skin.data_ptr()
        self._nl_keep = v  # keep the struct alive across the asynchronous call
        return C.byref(v)

    @staticmethod
    def compute_force(state, system):
        """-> jdb200_neighborlist_compute_force (neighbor_list.py:542-632)."""
        _call.call("jdb200_neighborlist_compute_force", state, system, system.collider._nlist_view())
        return state, system

    @staticmethod
    def compute_potential_energy(state, system):
        """-> jdb200_neighborlist_compute_potential_energy (neighbor_list.py:634-727)."""
        e = torch.empty(state.pos_c.shape[:-2], dtype=state.dtype, device=state.device)
        _call.call("jdb200_neighborlist_compute_potential_energy", state, system, system.collider._nlist_view(), e)
        return state, system, e

    @staticmethod
    def create_neighbor_list(state, system, cutoff=None, max_neighbors=None):
        """NeighborList.create_neighbor_list (neighbor_list.py:404-441): refresh the cached list if it is stale and
        return it; ``cutoff`` and ``max_neighbors`` are ignored like in the reference."""
        col = system.collider
        _call.call("jdb200_neighborlist_refresh", state, system, col._nlist_view())
        return state, system, col.neighbor_list, col.overflow

    @staticmethod
    def create_cross_neighbor_list(pos_a, pos_b, system, cutoff, max_neighbors: int):
        """Delegates to the secondary collider (neighbor_list.py:729-776)."""
        import copy
        inner = copy.copy(system)
        inner.collider = system.collider.secondary_collider
        return inner.collider.create_cross_neighbor_list(pos_a, pos_b, inner, cutoff, max_neighbors)


Collider.register("b200neighborlist")(NeighborList)


# ---------------------------------------------------------------------------
# Integrators (reference jaxdem/integrators/__init__.py:21-149)
# ---------------------------------------------------------------------------
class Integrator(Factory):
    native_kind = ""

    @staticmethod
    def step_before_force(state, system):
        return state, system

    @staticmethod
    def step_after_force(state, system):
        return state, system

    @staticmethod
    def initialize(state, system):
        return state, system


class LinearIntegrator(Integrator):
    _registry: dict = {}

    @staticmethod
    def step_before_force(state, system):
        if system.linear_integrator.native_kind:
            _call.call("jdb200_linear_step_before_force", state, system, needs_ws=False)
        return state, system

    @staticmethod
    def step_after_force(state, system):
        if system.linear_integrator.native_kind:
            _call.call("jdb200_linear_step_after_force", state, system, needs_ws=False)
        return state, system


class RotationIntegrator(Integrator):
    _registry: dict = {}

    @staticmethod
    def step_before_force(state, system):
        if system.rotation_integrator.native_kind:
            _call.call("jdb200_rotation_step_before_force", state, system, needs_ws=False)
        return state, system

    @staticmethod
    def step_after_force(state, system):
        if system.rotation_integrator.native_kind:
            _call.call("jdb200_rotation_step_after_force", state, system, needs_ws=False)
        return state, system


LinearIntegrator.register("")(LinearIntegrator)
RotationIntegrator.register("")(RotationIntegrator)


@LinearIntegrator.register("verlet")
class VelocityVerlet(LinearIntegrator):
    """reference jaxdem/integrators/velocity_verlet.py:57-95."""
    native_kind = "verlet"


@LinearIntegrator.register("euler")
class DirectEuler(LinearIntegrator):
    """reference jaxdem/integrators/direct_euler.py:62-66."""
    native_kind = "euler"


@RotationIntegrator.register("verletspiral")
class VelocityVerletSpiral(RotationIntegrator):
    """reference jaxdem/integrators/velocity_verlet_spiral.py:83-180."""
    native_kind = "verletspiral"


@RotationIntegrator.register("spiral")
class Spiral(RotationIntegrator):
    """reference jaxdem/integrators/spiral.py:104-141."""
    native_kind = "spiral"


for _k, _c in (("b200verlet", VelocityVerlet), ("b200euler", DirectEuler)):
    LinearIntegrator.register(_k)(_c)
for _k, _c in (("b200verletspiral", VelocityVerletSpiral), ("b200spiral", Spiral)):
    RotationIntegrator.register(_k)(_c)


# ---------------------------------------------------------------------------
# ForceManager (reference jaxdem/forces/force_manager.py:31-479; not a Factory)
# ---------------------------------------------------------------------------
class ForceManager:
    def __init__(self, gravity, external_force, external_force_com, external_torque):
        self.gravity = gravity
        self.external_force = external_force
        self.external_force_com = external_force_com
        self.external_torque = external_torque

    @staticmethod
    def create(state_shape, *, gravity=None, dtype=torch.float32, device="cpu") -> "ForceManager":
        dim = state_shape[-1]
        A = 1 if dim == 2 else 3
        batch = state_shape[0] if len(state_shape) == 3 else None
        g = torch.zeros(dim) if gravity is None else gravity
        z = lambda *s: torch.zeros(s, dtype=dtype, device=device)
        fm = ForceManager(_leaf(g, dtype, device, batch, (dim,)), z(*state_shape), z(*state_shape),
                          z(*state_shape[:-1], A))
        fm.mark_clean()
        return fm

    # -- host-side knowledge "the three external buffers are all zero" (jdb200_params.promises) ----------
    def _versions(self):
        return tuple((t.data_ptr(), t._version) for t in (self.external_force, self.external_force_com,
                                                           self.external_torque))

    def buffers_clean(self) -> bool:
        """True when nothing was added since the buffers were created as zeros or last cleared by ``apply``
        (any torch-side in-place edit bumps the tensors' version counters)."""
        return getattr(self, "_clean_at", None) == self._versions()

    def mark_clean(self) -> None:
        """Called after an entry point that runs ForceManager.apply (which zeroes the buffers on the device)."""
        self._clean_at = self._versions()

    @staticmethod
    def _member_count(state):
        """count = bincount(clump_id)[clump_id] (force_manager.py:227-228)."""
        cid = state.clump_id.long()
        cnt = torch.zeros_like(cid).scatter_add_(-1, cid, torch.ones_like(cid))
        return torch.gather(cnt, -1, cid).to(state.dtype)

    @staticmethod
    def add_force(state, system, force, *, is_com=False):
        """ForceManager.add_force (force_manager.py:196-232): a COM force is shared out over the members of the
        clump, because ``apply`` segment-sums the buffer over each clump without dividing."""
        fm = system.force_manager
        force = torch.as_tensor(force, dtype=state.dtype, device=state.device)
        if is_com:
            if state.has_clumps:
                force = force / ForceManager._member_count(state)[..., None]
            fm.external_force_com.add_(force)
        else:
            fm.external_force.add_(force)
        return system

    @staticmethod
    def add_torque(state, system, torque):
        """ForceManager.add_torque (force_manager.py:270-303): divided by the member count like a COM force."""
        torque = torch.as_tensor(torque, dtype=state.dtype, device=state.device)
        if state.has_clumps:
            torque = torque / ForceManager._member_count(state)[..., None]
        system.force_manager.external_torque.add_(torque)
        return system

    @staticmethod
    def apply(state, system):
        """-> jdb200_force_manager_apply (force_manager.py:338-425)."""
        _call.call("jdb200_force_manager_apply", state, system)
        system.force_manager.mark_clean()
        return state, system
