"""numpy restatement of the energy-minimisation loop and the FIRE optimiser.  Oracle only.

jaxdem/minimizers/routines.py:22-62 (delta parameters), :151-383 (``minimize``: one force/energy evaluation per
iteration, termination tests) and jaxdem/minimizers/optimizers.py:127-340 (``fire``).

Third-party arithmetic: the reference drives the loop through optax (``optax>=0.2.6`` in pyproject.toml:26, not
vendored, not installed here).  Two optax functions are on the path and are restated from their published
definitions: ``optax.apply_updates(params, updates) = params + updates`` and
``optax.safe_norm(x, min_norm, axis=-1, keepdims=True)`` = ``where(norm <= min_norm, min_norm,
norm(where(norm <= min_norm, 1, x)))`` — i.e. the row norm, floored at ``min_norm``.  The reference holds no test
for ``minimizers/``; this restatement is pinned on outputs of the reference's own ``minimize`` / ``fire``
run on the numpy stand-in for JAX (tests/golden/extras/fire.npz, tests/test_reference_golden.py): iterates after K
iterations, energy, and the stop iteration of a run to convergence, float64.
"""

from __future__ import annotations

import numpy as np

from . import colliders, force_manager
from . import linalg as la


class FireConfig:
    """Arguments of ``fire`` (optimizers.py:127-137)."""

    def __init__(self, dt, alpha_init=0.1, f_inc=1.1, f_dec=0.5, f_alpha=0.99, N_min=5, N_bad_max=10,
                 dt_max_scale=10.0, dt_min_scale=1e-3):
        self.dt, self.alpha_init, self.f_inc, self.f_dec, self.f_alpha = dt, alpha_init, f_inc, f_dec, f_alpha
        self.N_min, self.N_bad_max, self.dt_max_scale, self.dt_min_scale = N_min, N_bad_max, dt_max_scale, dt_min_scale


class FireState:
    """FIREState (optimizers.py:107-125): vel is the {'pos_c', 'rotvec'} pytree."""

    def __init__(self, cfg: FireConfig, state):
        F = state.fdtype
        A = 1 if state.dim == 2 else 3
        self.vel = {"pos_c": np.zeros_like(state.pos_c), "rotvec": np.zeros((state.N, A), F)}
        self.dt = F.type(cfg.dt)
        self.alpha = F.type(cfg.alpha_init)
        self.N_good = 0
        self.N_bad = 0


def _safe_norm(x, min_norm):
    """optax.safe_norm(x, min_norm, axis=-1, keepdims=True)."""
    n = np.sqrt(np.sum(x * x, axis=-1, keepdims=True))
    masked = np.sqrt(np.sum(np.where(n <= min_norm, np.ones_like(x), x) ** 2, axis=-1, keepdims=True))
    return np.where(n <= min_norm, x.dtype.type(min_norm), masked)


def fire_update(cfg: FireConfig, fs: FireState, grads):
    """``update`` of ``fire`` (optimizers.py:227-322).  grads: {'pos_c', 'rotvec'}; returns the updates to add to
    the parameters and advances ``fs`` in place."""
    F_t = fs.dt.dtype.type
    keys = ("pos_c", "rotvec")  # jax.tree.leaves order of the dict
    Fo = {k: -grads[k] for k in keys}
    v_old = {k: fs.vel[k] + Fo[k] * fs.dt / F_t(2.0) for k in keys}
    power = sum(np.sum(Fo[k] * v_old[k]) for k in keys)
    dt_inc = min(fs.dt * F_t(cfg.f_inc), F_t(cfg.dt * cfg.dt_max_scale))
    dt_dec = max(fs.dt * F_t(cfg.f_dec), F_t(cfg.dt * cfg.dt_min_scale))
    if power > 0.0:  # downhill
        N_good = fs.N_good + 1
        new_dt = dt_inc if N_good > cfg.N_min else fs.dt
        new_alpha = fs.alpha * F_t(cfg.f_alpha) if N_good > cfg.N_min else fs.alpha
        N_bad, dt_rev, vscale = 0, F_t(0.0), F_t(1.0)
    else:  # uphill
        N_bad = fs.N_bad + 1
        exceeded = N_bad > cfg.N_bad_max
        new_dt = F_t(cfg.dt) if exceeded else dt_dec
        N_bad = 0 if exceeded else N_bad
        new_alpha, N_good, dt_rev, vscale = F_t(cfg.alpha_init), 0, -new_dt, F_t(0.0)
    upd = {}
    for k in keys:
        v_half = v_old[k] * vscale + Fo[k] * new_dt / F_t(2.0)
        vn = _safe_norm(v_half, 1e-16)
        fn = _safe_norm(Fo[k], 1e-16)
        mix = np.where(fn > 1e-16, vn / fn * new_alpha, F_t(0.0))
        v_half = (v_half * (F_t(1.0) - new_alpha) + Fo[k] * mix) * vscale
        upd[k] = v_old[k] * dt_rev / F_t(2.0) + v_half * new_dt / F_t(2.0)
        fs.vel[k] = v_half.astype(v_old[k].dtype)
    fs.dt, fs.alpha, fs.N_good, fs.N_bad = F_t(new_dt), F_t(new_alpha), N_good, N_bad
    return upd


def q_from_rotvec(rotvec):
    """Quaternion.from_rotvec (utils/quaternion.py:99-129)."""
    n2 = la.dot(rotvec, rotvec)
    theta = np.sqrt(np.maximum(n2, n2.dtype.type(1e-16)))
    half = n2.dtype.type(0.5) * theta
    return np.cos(half)[..., None], rotvec * (np.sin(half) / theta)[..., None]


def apply_delta_params(state, pos_c, rotvec):
    """_delta_params_to_state (routines.py:38-62): q <- unit(from_rotvec(rotvec) @ q), pos_c <- pos_c."""
    if state.dim == 2:
        rotvec = np.concatenate([np.zeros_like(pos_c), rotvec], axis=-1)
    dw, dxyz = q_from_rotvec(rotvec.astype(state.fdtype))
    w, xyz = la.q_mul(dw, dxyz, state.q_w, state.q_xyz)
    w, xyz = la.q_unit(w, xyz)
    state.pos_c = pos_c.astype(state.fdtype)
    state.set_q(w.astype(state.fdtype), xyz.astype(state.fdtype))


def eval_step(state, system):
    """_objective_energy (routines.py:65-100): collider.compute_force -> force_manager.apply -> potential energy
    (utils/thermal.py:125-150: force manager + collider)."""
    colliders.compute_force(state, system)
    force_manager.apply(state, system)
    return force_manager.compute_potential_energy(state, system) + colliders.compute_potential_energy(state, system)


def minimize(state, system, cfg: FireConfig, max_steps=10000, pe_tol=1e-16, pe_diff_tol=1e-16, force_tol=0.0):
    """``minimize`` (routines.py:151-383), target_fn None.  Returns (steps, final_pe / N); state is updated in
    place.  ``history`` (list of per-iteration pe) is returned third for the tests."""
    F_t = state.fdtype.type
    N = state.N
    fs = FireState(cfg, state)
    pe = eval_step(state, system)
    prev_pe = F_t(np.inf)
    steps = 0
    tiny = np.finfo(state.fdtype).tiny
    history = [pe]
    while True:
        # cond_fun (:284-310); the gradient tested is the unmasked one of the last evaluation
        conv_pe = abs(pe / N) <= pe_tol
        denom = max(abs(pe), abs(prev_pe), tiny)
        conv_rel = (abs(pe - prev_pe) / denom < pe_diff_tol) if np.isfinite(prev_pe) else False
        max_grad = max(np.max(np.abs(state.force), initial=0.0), np.max(np.abs(state.torque), initial=0.0))
        if not (steps < max_steps) or conv_pe or conv_rel or (max_grad <= force_tol):
            break
        # body_fun (:312-372)
        mask = (~state.fixed)[:, None]
        grads = {"pos_c": -state.force * mask, "rotvec": -state.torque * mask}
        upd = fire_update(cfg, fs, grads)
        A = state.torque.shape[-1]
        new_pos = np.where(mask, state.pos_c + upd["pos_c"] * mask, state.pos_c)
        new_rot = np.where(mask, np.zeros((N, A), state.fdtype) + upd["rotvec"] * mask, np.zeros((N, A), state.fdtype))
        apply_delta_params(state, new_pos, new_rot)
        new_pe = eval_step(state, system)
        prev_pe, pe = pe, new_pe
        steps += 1
        history.append(pe)
    return steps, pe / N, history
