"""Energy minimisation mirroring jaxdem.minimizers (reference jaxdem/minimizers/routines.py:151-383 ``minimize``,
jaxdem/minimizers/optimizers.py:127-340 ``fire``).

The reference expresses the optimiser as an optax gradient transformation driven by a ``lax.while_loop``; here the
whole loop body — FIRE update, parameter -> State map, ONE force / energy evaluation, termination test — runs as
CUDA kernels enqueued ``poll`` iterations at a time by ``jdb200_minimize_fire`` (csrc/minimize.cu).  The loop carry
(FIREState, pe, prev_pe, step count, the per-system ``active`` flag) stays in device memory; the host only reads the
``active`` flags between chunks to know when to stop launching.
"""

from __future__ import annotations

import ctypes as C
import inspect
from typing import Any

import torch

from . import _call, _lib


class FireMinimizer:
    """What ``fire(...)`` returns: the FIRE hyper-parameters (optimizers.py:127-137, 324-340)."""
    type_name = "fire"

    def __init__(self, **kw: Any):
        self.kw = dict(kw)
        for k, v in kw.items():
            setattr(self, k, v)

    def __eq__(self, other):
        return isinstance(other, FireMinimizer) and self.kw == other.kw

    def __hash__(self):
        return hash(("fire", tuple(sorted((k, str(v)) for k, v in self.kw.items()))))


def fire(dt: float, alpha_init: float = 0.1, f_inc: float = 1.1, f_dec: float = 0.5, f_alpha: float = 0.99,
         N_min: int = 5, N_bad_max: int = 10, dt_max_scale: float = 10.0, dt_min_scale: float = 1e-3) -> FireMinimizer:
    """Fast Inertial Relaxation Engine (Bitzek et al. 2006); same arguments and defaults as the reference's ``fire``."""
    return FireMinimizer(dt=float(dt), alpha_init=float(alpha_init), f_inc=float(f_inc), f_dec=float(f_dec),
                         f_alpha=float(f_alpha), N_min=int(N_min), N_bad_max=int(N_bad_max),
                         dt_max_scale=float(dt_max_scale), dt_min_scale=float(dt_min_scale))


def make_minimizer(minimizer, minimizer_kw, dt):
    """System.create's minimiser plumbing (system.py:517-541): default FIRE; ``dt`` defaults to the system's."""
    minimizer = fire if minimizer is None else minimizer
    kw = {} if minimizer_kw is None else dict(minimizer_kw)
    if "dt" in inspect.signature(minimizer).parameters:
        kw.setdefault("dt", float(dt))
    return minimizer(**kw)


class FireState:
    """The while_loop carry of ``minimize`` (FIREState + pe / prev_pe / steps / active), device resident."""

    def __init__(self, state):
        lead, dev, F = tuple(state.pos_c.shape[:-2]), state.device, state.dtype
        A = 1 if state.dim == 2 else 3
        z = lambda *s, dt=F: torch.zeros((*lead, *s), dtype=dt, device=dev)
        self.vel_pos, self.vel_rot = z(state.N, state.dim), z(state.N, A)
        self.dt, self.alpha, self.pe, self.prev_pe = z(), z(), z(), z()
        self.n_good, self.n_bad, self.steps = z(dt=torch.int64), z(dt=torch.int64), z(dt=torch.int64)
        self.active = z(dt=torch.int32)

    def view(self) -> _lib.FireStateView:
        v = _lib.FireStateView()
        for k in ("vel_pos", "vel_rot", "dt", "alpha", "n_good", "n_bad", "pe", "prev_pe", "steps", "active"):
            setattr(v, k, getattr(self, k).data_ptr())
        return v


def minimize(state, system, max_steps: int = 10000, pe_tol: float = 1e-16, pe_diff_tol: float = 1e-16,
             force_tol: float = 0.0, *, poll: int = 64, return_carry: bool = False):
    """``minimize`` (routines.py:151-383): returns ``(state, system, steps, final_pe)`` with ``final_pe`` per particle,
    ``steps`` an int64 tensor (one entry per system of a batch).  Terminates per system when ``steps == max_steps``,
    ``|pe / N| <= pe_tol``, the relative pe change ``< pe_diff_tol`` or ``max |grad| <= force_tol``.

    ``poll``: iterations enqueued per host visit (the host reads one int32 per system between chunks)."""
    mz = getattr(system, "minimizer", None)
    if mz is None:
        raise ValueError("No minimizer configured in System. Please configure `minimizer` in System.create.")
    if not isinstance(mz, FireMinimizer):
        raise NotImplementedError("jaxdem_b200.minimize runs the FIRE optimiser (jaxdem_b200.minimizers.fire)")
    kind = getattr(system.collider, "native_kind", "")
    if kind not in ("celllist", "multicelllist", "naive", "neighborlist"):
        raise NotImplementedError(f"minimize needs a native collider, got {type(system.collider).__name__}")
    _call.require_cuda(state)
    fp = _lib.FireParams(dt=mz.dt, alpha_init=mz.alpha_init, f_inc=mz.f_inc, f_dec=mz.f_dec, f_alpha=mz.f_alpha,
                         dt_max_scale=mz.dt_max_scale, dt_min_scale=mz.dt_min_scale, pe_tol=float(pe_tol),
                         pe_diff_tol=float(pe_diff_tol), force_tol=float(force_tol), n_min=mz.N_min,
                         n_bad_max=mz.N_bad_max, max_steps=int(max_steps))
    carry = FireState(state)
    fv = carry.view()
    nl = system.collider._nlist_view() if kind == "neighborlist" else None
    done, init = 0, 1
    while True:
        chunk = max(0, min(int(poll), int(max_steps) - done))
        _call.call("jdb200_minimize_fire", state, system, nl, C.byref(fv), C.byref(fp), C.c_int64(chunk),
                   C.c_int32(init))
        init = 0
        done += chunk
        system.force_manager.mark_clean()
        if done >= int(max_steps) or not bool(carry.active.any()):
            break
    final_pe = carry.pe / state.N
    if return_carry:
        return state, system, carry.steps, final_pe, carry
    return state, system, carry.steps, final_pe
