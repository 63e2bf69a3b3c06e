"""System mirroring jaxdem.System (reference jaxdem/system.py:123-934): holds the
plugin components and drives the step loop.

``System.step(state, system, n=k)`` keeps the reference's hook order
(_step_once, system.py:60-82).  When every component is one of this package's
native plugins and the user hooks are identity, the k steps run inside ONE C-ABI
call (``jdb200_system_step``): kernels are enqueued back to back on the current
CUDA stream with no host round trip.  Otherwise the hooks are called one by one,
exactly like the reference.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Callable

import torch

from . import _call, _lib
from .components import (Collider, Domain, ForceManager, ForceModel, LinearIntegrator, RotationIntegrator,
                         _leaf)
from .materials import Material, MaterialMatchmaker, MaterialTable
from .minimizers import make_minimizer
from .state import State


def _identity(state, system):
    return state, system


class System:
    def __init__(self, **kw: Any):
        for k, v in kw.items():
            setattr(self, k, v)

    # -- construction (system.py:230-270,425-604) ---------------------------------
    @staticmethod
    def create(state_shape, *, dt=0.005, time=0.0, linear_integrator_type="verlet",
               rotation_integrator_type="verletspiral", collider_type="naive", domain_type="free",
               force_model_type="spring", force_manager_kw=None, mat_table=None,
               linear_integrator_kw=None, rotation_integrator_kw=None, collider_kw=None, domain_kw=None,
               force_model_kw=None, collider=None, domain=None, force_manager=None,
               interact_same_bond_id=False, user_pre_step_actions: Callable = _identity,
               user_post_step_actions: Callable = _identity, minimizer=None, minimizer_kw=None, dtype=None,
               device=None) -> "System":
        from .state import default_device, default_float
        F = dtype or default_float()
        dev = torch.device(device) if device is not None else default_device()
        state_shape = tuple(state_shape)
        dim = state_shape[-1]
        batch = state_shape[0] if len(state_shape) == 3 else None
        if mat_table is None:
            mat_table = MaterialTable.from_materials(
                [Material.create("elastic", density=0.27, young=1.0e4, poisson=0.3)],
                matcher=MaterialMatchmaker.create("harmonic"))
        fmodel = ForceModel.create(force_model_type, **(force_model_kw or {}))
        missing = [k for k in fmodel.required_material_properties
                   if k not in mat_table.props and k not in mat_table.pair]
        if missing:
            raise KeyError(f"MaterialTable lacks fields {missing}, required by the selected force model.")
        if fmodel.requires_history:
            raise ValueError("history-tracking force models need the NeighborList collider")
        mt = mat_table.to(device=dev, dtype=F)
        if batch is not None:  # every System leaf carries the batch axis (vmap semantics)
            bc = lambda t: t.expand(batch, *t.shape).contiguous()
            mt = MaterialTable({k: bc(v) for k, v in mt.props.items()}, {k: bc(v) for k, v in mt.pair.items()},
                               mt.matcher)
        if domain is None:
            domain = Domain.create(domain_type, dim=dim, dtype=F, device=dev, batch=batch, **(domain_kw or {}))
        if collider is None:
            collider = Collider.create(collider_type, **(collider_kw or {}))
        collider._bind(F, dev, batch)
        if force_manager is None:
            force_manager = ForceManager.create(state_shape, dtype=F, device=dev, **(force_manager_kw or {}))
        return System(
            linear_integrator=LinearIntegrator.create(linear_integrator_type, **(linear_integrator_kw or {})),
            rotation_integrator=RotationIntegrator.create(rotation_integrator_type,
                                                          **(rotation_integrator_kw or {})),
            collider=collider, domain=domain, force_manager=force_manager, force_model=fmodel, mat_table=mt,
            dt=_leaf(dt, F, dev, batch, ()), time=_leaf(time, F, dev, batch, ()),
            step_count=torch.zeros((batch,) if batch is not None else (), dtype=torch.int64, device=dev),
            dim=dim, interact_same_bond_id=_leaf(bool(interact_same_bond_id), torch.bool, dev, batch, ()),
            user_pre_step_actions=user_pre_step_actions, user_post_step_actions=user_post_step_actions,
            minimizer=make_minimizer(minimizer, minimizer_kw, dt),
        )

    # -- stepping -------------------------------------------------------------------
    def _is_native(self) -> bool:
        """True when the single-call driver may replace the hook-by-hook loop: identity user hooks and every
        hook of every component is this package's own implementation (a registered plugin that overrides a
        hook in Python must run that override, reference plugin contract factory.py:240-320)."""
        def own(obj, root, hooks):
            impls = [root] + list(root._registry.values())
            return all(any(getattr(type(obj), h, None) is getattr(k, h, None) for k in impls
                           if k.__module__.startswith(__package__)) for h in hooks)

        from .components import ForceManager as _FM
        return (self.user_pre_step_actions is _identity and self.user_post_step_actions is _identity
                and own(self.collider, Collider, ("compute_force",))
                and getattr(self.collider, "native_kind", None) in _lib.COLLIDER
                and own(self.linear_integrator, LinearIntegrator, ("step_before_force", "step_after_force"))
                and own(self.rotation_integrator, RotationIntegrator, ("step_before_force", "step_after_force"))
                and own(self.domain, Domain, ("apply",))
                and type(self.force_manager).apply is _FM.apply
                and type(self.force_model).__module__.startswith(__package__))

    @staticmethod
    def _step_once(state: State, system: "System"):
        """_step_once (system.py:60-82), hook by hook."""
        system.time += system.dt
        system.step_count += 1
        state, system = system.user_pre_step_actions(state, system)
        state, system = system.domain.apply(state, system)
        torch.reciprocal(system.domain.box_size, out=system.domain.inv_box_size)
        state, system = system.linear_integrator.step_before_force(state, system)
        state, system = system.rotation_integrator.step_before_force(state, system)
        state, system = system.collider.compute_force(state, system)
        state, system = system.force_manager.apply(state, system)
        state, system = system.linear_integrator.step_after_force(state, system)
        state, system = system.rotation_integrator.step_after_force(state, system)
        state, system = system.user_post_step_actions(state, system)
        return state, system

    @staticmethod
    def step(state: State, system: "System", *, n: int = 1, fused: bool | None = None):
        """System.step (system.py:701-748).  ``fused=None`` picks the single-call driver
        whenever it is legal; ``fused=False`` forces the hook-by-hook path."""
        n = int(n)
        if fused is None:
            fused = system._is_native()
        if fused:
            # time and step_count advance on the device, once per step, inside the same call
            if system.collider.native_kind == "neighborlist":
                _call.call("jdb200_system_step_nl", state, system, C.c_int64(n), system.collider._nlist_view(),
                           clock=True)
            else:
                _call.call("jdb200_system_step", state, system, C.c_int64(n), clock=True)
            if n > 0:
                system.force_manager.mark_clean()  # every step ends with the external buffers cleared
        else:
            for _ in range(n):
                state, system = System._step_once(state, system)
        return state, system

    @staticmethod
    def minimize(state: State, system: "System", **kw):
        """System.minimize (system.py:811-842) -> jaxdem_b200.minimizers.minimize."""
        from .minimizers import minimize
        return minimize(state, system, **kw)

    @staticmethod
    def compile_step(state: State, system: "System", *, n: int = 1) -> "CompiledStep":
        """The counterpart of ``jax.jit(System.step)``: capture ``n`` steps into a CUDA graph
        (every C-ABI entry point is stream-ordered and capture-legal, include/jaxdem_b200.h)
        and return a callable that replays it on the current stream."""
        return CompiledStep(state, system, int(n))

    @staticmethod
    def trajectory_rollout(state: State, system: "System", *, n: int | None = None, stride: int = 1, strides=None,
                           save_fn: Callable | None = None, fields=("pos_c", "vel", "force", "ang_vel", "q_w",
                                                                     "q_xyz", "pos"),
                           to_host: bool = False, ring: int = 4):
        """System.trajectory_rollout (system.py:606-699): one saved frame per entry of ``strides`` (or ``n``
        frames ``stride`` steps apart), each saved AFTER its steps; a leading 0 stride records the initial state.
        Returns ``(state, system, trajectory)``.

        ``save_fn=None`` (default) is the reference's "save the state" in this package's form: the selected
        State ``fields`` of a frame are packed by ONE kernel (jdb200_frame_pack) and the trajectory is a
        :class:`Frames` object whose attributes (``traj.pos``, ``traj.vel`` ... , ``traj.time``,
        ``traj.step_count``) carry a new leading frame axis.  ``to_host=False`` keeps it on the device;
        ``to_host=True`` streams it out while the simulation continues: frames are packed into a small device
        ring, a second stream copies each slot to pinned host memory, events keep the two streams honest
        (``ring`` slots in flight).  A user ``save_fn(state, system)`` is called per frame instead and its
        results are stacked (dict / tuple / tensor), exactly like the reference's scan."""
        if strides is not None:
            strides = [int(k) for k in (strides.tolist() if hasattr(strides, "tolist") else strides)]
        else:
            if n is None:
                raise ValueError("`n` must be provided when `strides` is None.")
            strides = [int(stride)] * int(n)
        if save_fn is not None:
            frames = []
            for k in strides:
                state, system = System.step(state, system, n=k)
                frames.append(save_fn(state, system))
            return state, system, _stack_frames(frames)
        rec = FrameRing(state, system, len(strides), fields, to_host=to_host, ring=ring)
        for f, k in enumerate(strides):
            state, system = System.step(state, system, n=k)
            rec.save(f, state, system)
        return state, system, rec.finish()


def _stack_frames(frames):
    if not frames:
        return frames
    f0 = frames[0]
    if isinstance(f0, dict):
        return {k: torch.stack([fr[k] for fr in frames]) for k in f0}
    if isinstance(f0, (tuple, list)):
        return type(f0)(_stack_frames([fr[i] for fr in frames]) for i in range(len(f0)))
    if isinstance(f0, torch.Tensor):
        return torch.stack(frames)
    return frames


class Frames:
    """Stacked frames of a rollout: one tensor per saved field, leading axis = frame."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getitem__(self, k):
        return self.__dict__[k]

    def keys(self):
        return self.__dict__.keys()


class FrameRing:
    """Trajectory output path (reference system.py:101-120 stacks a copy of the whole pytree per frame).
    Device side: jdb200_frame_pack writes the selected leaves of a frame as ONE contiguous record.
    ``to_host``: the record goes to a slot of a device ring; a copy stream moves it to its row of one pinned
    host array (cudaMemcpyAsync D2H) while the step kernels of the next frame already run; per-slot events
    order "packed -> copied -> slot free again".  Otherwise the record is a row of a device-resident array."""

    def __init__(self, state: State, system: "System", n_frames: int, fields, *, to_host: bool, ring: int = 4):
        _call.require_cuda(state)
        self.fields = tuple(fields)
        unknown = [f for f in self.fields if f not in _lib.FRAME_FIELDS]
        if unknown:
            raise KeyError(f"unknown frame fields {unknown}; known: {_lib.FRAME_FIELDS}")
        self.mask = sum(1 << _lib.FRAME_FIELDS.index(f) for f in set(self.fields))
        lead = tuple(state.pos_c.shape[:-2])
        N, D, A = state.N, state.dim, (1 if state.dim == 2 else 3)
        widths = {"pos_c": (N, D), "vel": (N, D), "force": (N, D), "ang_vel": (N, A), "torque": (N, A),
                  "q_w": (N, 1), "q_xyz": (N, 3), "pos": (N, D)}
        self.layout, off = [], 0
        for f in _lib.FRAME_FIELDS:  # the kernel's order
            if (self.mask >> _lib.FRAME_FIELDS.index(f)) & 1:
                ln = widths[f][0] * widths[f][1]
                self.layout.append((f, off, widths[f]))
                off += ln
        self.frame_len, self.lead, self.n_frames, self.to_host = off, lead, n_frames, to_host
        B = 1
        for x in lead:
            B *= x
        dev, F = state.device, state.dtype
        self.times = torch.empty((n_frames, *lead), dtype=F, device=dev)
        self.counts = torch.empty((n_frames, *lead), dtype=torch.int64, device=dev)
        if to_host:
            self.ring = max(1, min(int(ring), max(n_frames, 1)))
            self.slots = torch.empty((self.ring, B, self.frame_len), dtype=F, device=dev)
            self.host = torch.empty((n_frames, B, self.frame_len), dtype=F).pin_memory()
            self.copy_stream = torch.cuda.Stream(device=dev)
            self.packed = [torch.cuda.Event() for _ in range(self.ring)]
            self.copied = [None] * self.ring
        else:
            self.dev_frames = torch.empty((n_frames, B, self.frame_len), dtype=F, device=dev)

    def save(self, f: int, state: State, system: "System") -> None:
        import ctypes as C
        self.times[f].copy_(system.time)
        self.counts[f].copy_(system.step_count)
        if not self.to_host:
            _call.call("jdb200_frame_pack", state, system, C.c_int32(self.mask), self.dev_frames[f], needs_ws=False)
            return
        slot = f % self.ring
        cur = torch.cuda.current_stream(state.device)
        if self.copied[slot] is not None:
            cur.wait_event(self.copied[slot])  # the slot's previous frame has left the device
        _call.call("jdb200_frame_pack", state, system, C.c_int32(self.mask), self.slots[slot], needs_ws=False)
        self.packed[slot].record(cur)
        self.copy_stream.wait_event(self.packed[slot])
        with torch.cuda.stream(self.copy_stream):
            self.host[f].copy_(self.slots[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.copied[slot] = ev

    def finish(self) -> Frames:
        if self.to_host:
            self.copy_stream.synchronize()
            base = self.host
            times, counts = self.times.cpu(), self.counts.cpu()
        else:
            base, times, counts = self.dev_frames, self.times, self.counts
        out = {}
        for name, off, (n, w) in self.layout:
            t = base[:, :, off:off + n * w].reshape(self.n_frames, *self.lead, n, w)
            out[name] = t
        return Frames(time=times, step_count=counts, **out)


class CompiledStep:
    """``n`` fused steps captured once in a CUDA graph; ``__call__`` replays them in place on
    the same State / System buffers (pointers are baked into the graph)."""

    def __init__(self, state: State, system: System, n: int):
        if not system._is_native():
            raise RuntimeError("compile_step needs native components and identity user hooks")
        self.state, self.system, self.n = state, system, n
        _call.require_cuda(state)
        # allocate before capture; the graph bakes this pointer in, so keep the tensor alive with it
        self._ws = _call.workspace(_call.params_for(state, system), state.device, system)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=state.device)
        side.wait_stream(torch.cuda.current_stream(state.device))
        with torch.cuda.graph(self.graph, stream=side):
            System.step(state, system, n=n, fused=True)

    def __call__(self):
        self.graph.replay()
        return self.state, self.system
