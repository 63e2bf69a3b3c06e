// jaxdem_b200 — energy minimisation loop with the FIRE optimiser, device resident
// (reference jaxdem/minimizers/routines.py:151-383 `minimize`, jaxdem/minimizers/optimizers.py:127-340 `fire`).
//
// One iteration of the reference's while_loop body = FIRE update of the {pos_c, rotvec} parameters from the
// gradient {-force, -torque} (masked by ~fixed), parameters -> State (pos_c, q <- unit(from_rotvec(d) @ q)),
// ONE force / energy evaluation (collider.compute_force -> force_manager.apply -> potential energy), and the
// termination test.  Here all of it is enqueued on the stream for `n_iter` iterations at a time: the two global
// reductions (power, max |grad|) are two-level block reductions with fixed order, the FIRE scalars and the loop
// carry (pe, prev_pe, steps, active) live in device memory per system, and every update kernel starts with
// `if (!active[b]) return` — the batched while_loop's "finished elements keep their carry".  The host only polls
// `active` between chunks.
#include "ctx.cuh"
#include "launch.cuh"

namespace jdb {

template <typename F> int celllist_force(cudaStream_t, Ctx<F>&, int, bool, bool);
template <typename F> int celllist_energy(cudaStream_t, Ctx<F>&, F*, bool);
template <typename F> int naive_force(cudaStream_t, Ctx<F>&);
template <typename F> int naive_energy(cudaStream_t, Ctx<F>&, F*);
template <typename F> int neighborlist_force(cudaStream_t, Ctx<F>&);
template <typename F> int neighborlist_energy(cudaStream_t, Ctx<F>&, F*);
template <typename F> int force_manager_apply_pe(cudaStream_t, Ctx<F>&, F*);
template <typename F> int force_manager_fire_tail(cudaStream_t, Ctx<F>&, const F*, const F*, const F*);

template <typename F>
struct Fire {  // device view of jdb200_fire_state + the scalars of jdb200_fire_params in F
  F *vel_pos, *vel_rot, *dt, *alpha, *pe, *prev_pe;
  long long *n_good, *n_bad, *steps;
  int* active;
  F dt0, alpha_init, f_inc, f_dec, f_alpha, dt_max, dt_min, pe_tol, pe_diff_tol, force_tol;
  long long n_min, n_bad_max, max_steps;
};

template <typename F>
__device__ __forceinline__ F blk_sum(F v, F* sm) {
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int s = kReduceBlock / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  const F r = sm[0];
  __syncthreads();
  return r;
}
template <typename F>
__device__ __forceinline__ F blk_max(F v, F* sm) {
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int s = kReduceBlock / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] = RT<F>::fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
    __syncthreads();
  }
  const F r = sm[0];
  __syncthreads();
  return r;
}

// power = sum F . v_old over both parameter leaves, v_old = vel + F dt / 2 (optimizers.py:233-238)
template <typename F, int D>
__global__ void __launch_bounds__(kReduceBlock) k_fire_power(Ctx<F> c, Fire<F> fs) {
  pdl_prologue();
  constexpr int A = D == 3 ? 3 : 1;
  __shared__ F sm[kReduceBlock];
  const int b = blockIdx.y;
  if (!fs.active[b]) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  F pl = F(0), pr = F(0);
  if (i < c.n) {
    const size_t g = (size_t)b * c.n + i;
    const F m = c.fixed[g] ? F(0) : F(1);
    const F dt = fs.dt[b];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const F f = c.force[g * D + d] * m;
      const F vo = fs.vel_pos[g * D + d] + f * dt / F(2);
      pl += f * vo;
    }
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const F f = c.torque[g * A + a] * m;
      const F vo = fs.vel_rot[g * A + a] + f * dt / F(2);
      pr += f * vo;
    }
  }
  const F sl = blk_sum(pl, sm), sr = blk_sum(pr, sm);
  if (threadIdx.x == 0) {
    F* o = c.min_part + ((size_t)b * c.reduce_blocks + blockIdx.x) * 4;
    o[0] = sl;
    o[1] = sr;
  }
}

// the scalar half of `update` (optimizers.py:239-277): downhill / uphill bookkeeping
template <typename F>
__global__ void __launch_bounds__(kReduceBlock) k_fire_scalars(Ctx<F> c, Fire<F> fs) {
  pdl_prologue();
  __shared__ F sm[kReduceBlock];
  const int b = blockIdx.x;
  if (!fs.active[b]) return;
  F al = F(0), ar = F(0);
  for (int i = threadIdx.x; i < c.reduce_blocks; i += kReduceBlock) {
    const F* o = c.min_part + ((size_t)b * c.reduce_blocks + i) * 4;
    al += o[0];
    ar += o[1];
  }
  const F sl = blk_sum(al, sm), sr = blk_sum(ar, sm);
  if (threadIdx.x != 0) return;
  const F power = sl + sr;  // pos_c leaf + rotvec leaf
  const F dt = fs.dt[b], alpha = fs.alpha[b];
  const F dt_inc = RT<F>::fmin(dt * fs.f_inc, fs.dt_max);
  const F dt_dec = RT<F>::fmax(dt * fs.f_dec, fs.dt_min);
  F new_dt, new_alpha, dt_rev, vscale;
  long long n_good = fs.n_good[b], n_bad = fs.n_bad[b];
  if (power > F(0)) {
    n_good += 1;
    new_dt = n_good > fs.n_min ? dt_inc : dt;
    new_alpha = n_good > fs.n_min ? alpha * fs.f_alpha : alpha;
    n_bad = 0;
    dt_rev = F(0);
    vscale = F(1);
  } else {
    n_bad += 1;
    const bool exceeded = n_bad > fs.n_bad_max;
    new_dt = exceeded ? fs.dt0 : dt_dec;
    n_bad = exceeded ? 0 : n_bad;
    new_alpha = fs.alpha_init;
    n_good = 0;
    dt_rev = -new_dt;
    vscale = F(0);
  }
  F* t = c.min_scal + (size_t)b * 8;
  t[0] = dt; t[1] = new_dt; t[2] = new_alpha; t[3] = dt_rev; t[4] = vscale;
  fs.dt[b] = new_dt;
  fs.alpha[b] = new_alpha;
  fs.n_good[b] = n_good;
  fs.n_bad[b] = n_bad;
}

// optax.safe_norm(x, 1e-16, axis=-1): the row norm, floored
template <typename F>
__device__ __forceinline__ F fire_safe_norm(const F* x, int n) {
  F s = F(0);
  for (int k = 0; k < n; ++k) s += x[k] * x[k];
  const F nr = RT<F>::sqrt(s);
  return nr <= F(1e-16) ? F(1e-16) : nr;
}

template <typename F>
__device__ __forceinline__ void fire_leaf(const F* f, F* vel, F* upd, int n, const F* t) {
  const F dt_old = t[0], new_dt = t[1], new_alpha = t[2], dt_rev = t[3], vscale = t[4];
  F vo[3], vh[3];
  for (int k = 0; k < n; ++k) {
    vo[k] = vel[k] + f[k] * dt_old / F(2);
    vh[k] = vo[k] * vscale + f[k] * new_dt / F(2);
  }
  const F vn = fire_safe_norm(vh, n), fn = fire_safe_norm(f, n);
  const F mix = fn > F(1e-16) ? vn / fn * new_alpha : F(0);
  for (int k = 0; k < n; ++k) {
    vh[k] = (vh[k] * (F(1) - new_alpha) + f[k] * mix) * vscale;
    upd[k] = vo[k] * dt_rev / F(2) + vh[k] * new_dt / F(2);
    vel[k] = vh[k];
  }
}

// the per-particle half of `update` (optimizers.py:279-313) + apply_updates + _delta_params_to_state
// (routines.py:38-62, 336-343): pos_c += upd (free particles), q <- unit(from_rotvec(upd_rot) @ q) (all particles;
// the delta is zero for fixed ones), _pos_p_rot refreshed (State.q setter, state.py:264-273).
template <typename F, int D>
__global__ void __launch_bounds__(256) k_fire_update(Ctx<F> c, Fire<F> fs) {
  pdl_prologue();
  using T = RT<F>;
  constexpr int A = D == 3 ? 3 : 1;
  const int b = blockIdx.y;
  if (!fs.active[b]) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const size_t g = (size_t)b * c.n + i;
  const F* t = c.min_scal + (size_t)b * 8;
  const bool fixed = c.fixed[g] != 0;
  const F m = fixed ? F(0) : F(1);
  F fl[3], fr[3], ul[3], ur[3] = {0, 0, 0};
  for (int d = 0; d < D; ++d) fl[d] = c.force[g * D + d] * m;
  for (int a = 0; a < A; ++a) fr[a] = c.torque[g * A + a] * m;
  fire_leaf<F>(fl, fs.vel_pos + g * D, ul, D, t);
  fire_leaf<F>(fr, fs.vel_rot + g * A, ur, A, t);
  if (!fixed)
    for (int d = 0; d < D; ++d) c.pos_c[g * D + d] = c.pos_c[g * D + d] + ul[d] * m;
  // rotation vector of this iteration, anchored at the current orientation
  F rv[3];
  if (D == 3) { rv[0] = fixed ? F(0) : ur[0] * m; rv[1] = fixed ? F(0) : ur[1] * m; rv[2] = fixed ? F(0) : ur[2] * m; }
  else { rv[0] = F(0); rv[1] = F(0); rv[2] = fixed ? F(0) : ur[0] * m; }
  const F n2 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
  const F theta = T::sqrt(T::fmax(n2, F(1e-16)));
  const F half = F(0.5) * theta;
  const F cw = cos(half), sf = sin(half) / theta;
  const V3<F> dq = {rv[0] * sf, rv[1] * sf, rv[2] * sf};
  const F qw = c.q_w[g];
  const V3<F> qv = {c.q_xyz[g * 3 + 0], c.q_xyz[g * 3 + 1], c.q_xyz[g * 3 + 2]};
  const V3<F> cr = cross3(dq, qv);
  F w = cw * qw - dot3(dq, qv);
  V3<F> v = {cw * qv.x + qw * dq.x + cr.x, cw * qv.y + qw * dq.y + cr.y, cw * qv.z + qw * dq.z + cr.z};
  const F q2 = w * w + dot3(v, v);
  const F inv = T::rsqrt(q2 == F(0) ? F(1) : q2);
  w *= inv; v.x *= inv; v.y *= inv; v.z *= inv;
  c.q_w[g] = w;
  c.q_xyz[g * 3 + 0] = v.x; c.q_xyz[g * 3 + 1] = v.y; c.q_xyz[g * 3 + 2] = v.z;
  if (!(c.promises & JDB200_PROMISE_NO_POS_P)) {
    if (D == 3) {
      const V3<F> pp = {c.pos_p[g * 3 + 0], c.pos_p[g * 3 + 1], c.pos_p[g * 3 + 2]};
      const V3<F> r = q_rotate3(w, v, pp);
      c.pos_p_rot[g * 3 + 0] = r.x; c.pos_p_rot[g * 3 + 1] = r.y; c.pos_p_rot[g * 3 + 2] = r.z;
    } else {  // Quaternion.rotate, 2D branch (utils/quaternion.py:222-231)
      const F cc = w * w - v.z * v.z, ss = F(2) * w * v.z;
      const F px = c.pos_p[g * 2 + 0], py = c.pos_p[g * 2 + 1];
      c.pos_p_rot[g * 2 + 0] = cc * px - ss * py;
      c.pos_p_rot[g * 2 + 1] = ss * px + cc * py;
    }
  }
}

// max |grad| of cond_fun (routines.py:299-305): over force and torque of the evaluated state, unmasked
template <typename F>
__global__ void __launch_bounds__(kReduceBlock) k_fire_maxgrad(Ctx<F> c) {
  pdl_prologue();
  __shared__ F sm[kReduceBlock];
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  F m = F(0);
  if (i < c.n) {
    const size_t g = (size_t)b * c.n + i;
    for (int d = 0; d < c.dim; ++d) m = RT<F>::fmax(m, RT<F>::abs(c.force[g * c.dim + d]));
    for (int a = 0; a < c.A; ++a) m = RT<F>::fmax(m, RT<F>::abs(c.torque[g * c.A + a]));
  }
  const F r = blk_max(m, sm);
  if (threadIdx.x == 0) c.min_part[((size_t)b * c.reduce_blocks + blockIdx.x) * 4 + 2] = r;
}

// loop carry + cond_fun (routines.py:284-310, 364-372)
template <typename F>
__global__ void __launch_bounds__(kReduceBlock) k_fire_cond(Ctx<F> c, Fire<F> fs, int init, int fm_in_parts) {
  pdl_prologue();
  __shared__ F sm[kReduceBlock];
  const int b = blockIdx.x;
  if (!init && !fs.active[b]) return;
  F m = F(0), ge = F(0);
  for (int i = threadIdx.x; i < c.reduce_blocks; i += kReduceBlock) {
    m = RT<F>::fmax(m, c.min_part[((size_t)b * c.reduce_blocks + i) * 4 + 2]);
    if (fm_in_parts) ge += c.min_part[((size_t)b * c.reduce_blocks + i) * 4 + 3];  // k_fm_fire_tail (the order of k_fm_energy_final)
  }
  m = blk_max(m, sm);
  if (fm_in_parts) ge = blk_sum(ge, sm);
  if (threadIdx.x != 0) return;
  const F pe_fm = fm_in_parts ? -ge : c.min_pe[c.batch + b];
  const F pe_new = pe_fm + c.min_pe[b];  // force manager + collider (thermal.py:148-150)
  F pe, prev;
  long long steps;
  if (init) {
    pe = pe_new;
    prev = F(INFINITY);
    steps = 0;
  } else {
    prev = fs.pe[b];
    pe = pe_new;
    steps = fs.steps[b] + 1;
  }
  fs.pe[b] = pe;
  fs.prev_pe[b] = prev;
  fs.steps[b] = steps;
  const F pe_n = pe / (F)c.n;
  const bool running = steps < fs.max_steps;
  const bool conv_pe = RT<F>::abs(pe_n) <= fs.pe_tol;
  const F tiny = sizeof(F) == 4 ? F(1.17549435e-38) : F(2.2250738585072014e-308);
  const F denom = RT<F>::fmax(RT<F>::fmax(RT<F>::abs(pe), RT<F>::abs(prev)), tiny);
  const bool conv_rel = RT<F>::abs(pe - prev) / denom < fs.pe_diff_tol;  // inf / inf = NaN on the first test: false
  const bool conv_force = m <= fs.force_tol;
  fs.active[b] = running && !(conv_pe || conv_rel || conv_force);
}

template <typename F>
__global__ void k_fire_init(Ctx<F> c, Fire<F> fs) {
  pdl_prologue();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= c.batch) return;
  fs.dt[b] = fs.dt0;
  fs.alpha[b] = fs.alpha_init;
  fs.n_good[b] = 0;
  fs.n_bad[b] = 0;
}

// eval_step (routines.py:221-237, 65-100): collider.compute_force -> force_manager.apply -> potential energy
template <typename F>
static int fire_eval(cudaStream_t s, Ctx<F>& c, int collider) {
  int rc = 0;
  if (collider == JDB200_COLLIDER_CELLLIST || collider == JDB200_COLLIDER_MULTICELLLIST) {
    // default stencil on a dense table: the row kernel walks every pair once for force AND energy (each particle's
    // share lands in sforce.w); whatever it does not serve keeps the separate energy walk
    c.want_energy = (c.max_cells > 0 && c.M == (c.dim == 3 ? 27 : 9) && !c.prune && c.reduce_blocks == cdiv(c.n, kReduceBlock)) ? 1 : 0;
    if ((rc = celllist_force<F>(s, c, 0, false, true))) return rc;
    if ((rc = celllist_energy<F>(s, c, c.min_pe, true))) return rc;  // same positions: the partition is reused
    c.want_energy = 0;
  } else if (collider == JDB200_COLLIDER_NAIVE) {
    if ((rc = naive_force<F>(s, c))) return rc;
    if ((rc = naive_energy<F>(s, c, c.min_pe))) return rc;
  } else if (collider == JDB200_COLLIDER_NEIGHBORLIST) {
    if ((rc = neighborlist_force<F>(s, c))) return rc;
    if ((rc = neighborlist_energy<F>(s, c, c.min_pe))) return rc;
  } else {
    return JDB200_EINVAL;
  }
  return 0;
}

template <typename F>
int minimize_fire(cudaStream_t s, Ctx<F>& c, int collider, const jdb200_fire_state* st, const jdb200_fire_params* fp,
                  long long n_iter, int init) {
  if (c.n == 0) return 0;
  Fire<F> fs;
  fs.vel_pos = (F*)st->vel_pos; fs.vel_rot = (F*)st->vel_rot; fs.dt = (F*)st->dt; fs.alpha = (F*)st->alpha;
  fs.pe = (F*)st->pe; fs.prev_pe = (F*)st->prev_pe;
  fs.n_good = (long long*)st->n_good; fs.n_bad = (long long*)st->n_bad; fs.steps = (long long*)st->steps;
  fs.active = (int*)st->active;
  fs.dt0 = (F)fp->dt; fs.alpha_init = (F)fp->alpha_init; fs.f_inc = (F)fp->f_inc; fs.f_dec = (F)fp->f_dec;
  fs.f_alpha = (F)fp->f_alpha; fs.dt_max = (F)(fp->dt * fp->dt_max_scale); fs.dt_min = (F)(fp->dt * fp->dt_min_scale);
  fs.pe_tol = (F)fp->pe_tol; fs.pe_diff_tol = (F)fp->pe_diff_tol; fs.force_tol = (F)fp->force_tol;
  fs.n_min = fp->n_min; fs.n_bad_max = fp->n_bad_max; fs.max_steps = fp->max_steps;
  const int B = c.batch;
  const dim3 gr(c.reduce_blocks, B), gp(cdiv(c.n, 256), B), gb(cdiv(B, 64));
  F* pe_fm = c.min_pe + B;  // min_pe = [collider energies (B) | force-manager energies (B)]
  int rc = 0;
  // sphere systems: force manager, its energy, max |grad| and the NEXT iteration's FIRE power in one pass
  const bool tail = !c.clumps;
  auto evaluate = [&](int is_init) -> int {
    if ((rc = fire_eval<F>(s, c, collider))) return rc;
    if (tail) {
      if ((rc = force_manager_fire_tail<F>(s, c, fs.vel_pos, fs.vel_rot, fs.dt))) return rc;
    } else {
      if ((rc = force_manager_apply_pe<F>(s, c, pe_fm))) return rc;
      JDB_LAUNCH(k_fire_maxgrad<F>, gr, kReduceBlock, s, c);
    }
    JDB_LAUNCH(k_fire_cond<F>, dim3(B), kReduceBlock, s, c, fs, is_init, tail ? 1 : 0);
    return 0;
  };
  if (init) {
    if (cudaMemsetAsync(fs.vel_pos, 0, sizeof(F) * B * c.n * c.dim, s) != cudaSuccess) return JDB200_ECUDA;
    if (cudaMemsetAsync(fs.vel_rot, 0, sizeof(F) * B * c.n * c.A, s) != cudaSuccess) return JDB200_ECUDA;
    JDB_LAUNCH(k_fire_init<F>, gb, 64, s, c, fs);
    if ((rc = evaluate(1))) return rc;
  }
  for (long long it = 0; it < n_iter; ++it) {
    if (!tail || (it == 0 && !init)) {  // (a call that continues a run: the partials of the previous call are gone)
      if (c.dim == 3) JDB_LAUNCH((k_fire_power<F, 3>), gr, kReduceBlock, s, c, fs);
      else JDB_LAUNCH((k_fire_power<F, 2>), gr, kReduceBlock, s, c, fs);
    }
    JDB_LAUNCH(k_fire_scalars<F>, dim3(B), kReduceBlock, s, c, fs);
    if (c.dim == 3) JDB_LAUNCH((k_fire_update<F, 3>), gp, 256, s, c, fs);
    else JDB_LAUNCH((k_fire_update<F, 2>), gp, 256, s, c, fs);
    if ((rc = evaluate(0))) return rc;
  }
  return 0;
}

template int minimize_fire<float>(cudaStream_t, Ctx<float>&, int, const jdb200_fire_state*, const jdb200_fire_params*, long long, int);
template int minimize_fire<double>(cudaStream_t, Ctx<double>&, int, const jdb200_fire_state*, const jdb200_fire_params*, long long, int);

}  // namespace jdb
