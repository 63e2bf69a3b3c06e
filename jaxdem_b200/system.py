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
               user_post_step_actions: Callable = _identity, dtype=None, device=None) -> "System":
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
        )

    # -- stepping -------------------------------------------------------------------
    def _is_native(self) -> bool:
        return (self.user_pre_step_actions is _identity and self.user_post_step_actions is _identity
                and type(self.collider).compute_force in (Collider.compute_force,)
                + tuple(c.compute_force for c in Collider._registry.values())
                and getattr(self.collider, "native_kind", None) in _lib.COLLIDER)

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
            _call.call("jdb200_system_step", state, system, C.c_int64(n), clock=True)
        else:
            for _ in range(n):
                state, system = System._step_once(state, system)
        return state, system

    @staticmethod
    def compile_step(state: State, system: "System", *, n: int = 1) -> "CompiledStep":
        """The counterpart of ``jax.jit(System.step)``: capture ``n`` steps into a CUDA graph
        (every C-ABI entry point is stream-ordered and capture-legal, include/jaxdem_b200.h)
        and return a callable that replays it on the current stream."""
        return CompiledStep(state, system, int(n))

    @staticmethod
    def trajectory_rollout(state: State, system: "System", *, n: int, stride: int = 1, strides=None,
                           save_fn: Callable | None = None):
        """System.trajectory_rollout (system.py:606-699): n frames, each saved AFTER
        ``stride`` steps; returns (state, system, stacked frames).  The default save_fn
        keeps pos_c, vel, force, ang_vel, q (device tensors, stacked on a new leading axis)."""
        if save_fn is None:
            save_fn = lambda st, sy: {"pos_c": st.pos_c.clone(), "vel": st.vel.clone(), "force": st.force.clone(),
                                      "ang_vel": st.ang_vel.clone(), "q_w": st.q.w.clone(),
                                      "q_xyz": st.q.xyz.clone(), "time": sy.time.clone()}
        frames = []
        for f in range(n):
            k = int(strides[f]) if strides is not None else stride
            state, system = System.step(state, system, n=k)
            frames.append(save_fn(state, system))
        if frames and isinstance(frames[0], dict):
            traj = {k: torch.stack([fr[k] for fr in frames]) for k in frames[0]}
        else:
            traj = frames
        return state, system, traj


class CompiledStep:
    """``n`` fused steps captured once in a CUDA graph; ``__call__`` replays them in place on
    the same State / System buffers (pointers are baked into the graph)."""

    def __init__(self, state: State, system: System, n: int):
        if not system._is_native():
            raise RuntimeError("compile_step needs native components and identity user hooks")
        self.state, self.system, self.n = state, system, n
        _call.require_cuda(state)
        _call.workspace(_call.params_for(state, system), state.device)  # allocate before capture
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=state.device)
        side.wait_stream(torch.cuda.current_stream(state.device))
        with torch.cuda.graph(self.graph, stream=side):
            System.step(state, system, n=n, fused=True)

    def __call__(self):
        self.graph.replay()
        return self.state, self.system
