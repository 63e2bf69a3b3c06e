// jaxdem_b200 — rotation integrators (SPIRAL family), per-particle device code shared by the
// streaming kernel k_rotation (step.cu) and the fused step epilogue (pair.cu, EPI 2).
#pragma once
#include "ctx.cuh"

namespace jdb {

// ---------------------------------------------------------------------------
// K9  rotation integrators (SPIRAL)
// ---------------------------------------------------------------------------
template <typename F>
struct Q4 {
  F w, x, y, z;
};

template <typename F>
__device__ __forceinline__ V3<F> xcross(const V3<F>& a, const V3<F>& b) {  // unfused cross
  using T = RT<F>;
  return {T::sub(T::mul(a.y, b.z), T::mul(a.z, b.y)), T::sub(T::mul(a.z, b.x), T::mul(a.x, b.z)),
          T::sub(T::mul(a.x, b.y), T::mul(a.y, b.x))};
}
template <typename F>
__device__ __forceinline__ F xdot(const V3<F>& a, const V3<F>& b) {
  using T = RT<F>;
  return T::add(T::add(T::mul(a.x, b.x), T::mul(a.y, b.y)), T::mul(a.z, b.z));
}
// Quaternion.rotate 3D (quaternion.py:236-238): t = 2 q x v; v + w t + q x t
template <typename F>
__device__ __forceinline__ V3<F> xrotate(const Q4<F>& q, const V3<F>& v) {
  using T = RT<F>;
  const V3<F> qv = {q.x, q.y, q.z};
  V3<F> t = xcross(qv, v);
  t = {T::mul(F(2), t.x), T::mul(F(2), t.y), T::mul(F(2), t.z)};
  const V3<F> u = xcross(qv, t);
  return {T::add(T::add(v.x, T::mul(q.w, t.x)), u.x), T::add(T::add(v.y, T::mul(q.w, t.y)), u.y),
          T::add(T::add(v.z, T::mul(q.w, t.z)), u.z)};
}
template <typename F>
__device__ __forceinline__ V3<F> xrotate_back(const Q4<F>& q, const V3<F>& v) {
  return xrotate(Q4<F>{q.w, -q.x, -q.y, -q.z}, v);
}
// 2D rotate (quaternion.py:224-234)
template <typename F>
__device__ __forceinline__ void xrotate2(const Q4<F>& q, F vx, F vy, F& rx, F& ry) {
  using T = RT<F>;
  const F cth = T::sub(T::mul(q.w, q.w), T::mul(q.z, q.z));
  const F sth = T::mul(T::mul(F(2), q.w), q.z);
  rx = T::sub(T::mul(cth, vx), T::mul(sth, vy));
  ry = T::add(T::mul(sth, vx), T::mul(cth, vy));
}
// Hamilton product (quaternion.py:345-367)
template <typename F>
__device__ __forceinline__ Q4<F> xqmul(const Q4<F>& a, const Q4<F>& b) {
  using T = RT<F>;
  const V3<F> av = {a.x, a.y, a.z}, bv = {b.x, b.y, b.z};
  const V3<F> cr = xcross(av, bv);
  Q4<F> r;
  r.w = T::sub(T::mul(a.w, b.w), xdot(av, bv));
  r.x = T::add(T::add(T::mul(a.w, b.x), T::mul(b.w, a.x)), cr.x);
  r.y = T::add(T::add(T::mul(a.w, b.y), T::mul(b.w, a.y)), cr.y);
  r.z = T::add(T::add(T::mul(a.w, b.z), T::mul(b.w, a.z)), cr.z);
  return r;
}
// Quaternion.unit (quaternion.py:72-94)
template <typename F>
__device__ __forceinline__ Q4<F> xqunit(const Q4<F>& q) {
  using T = RT<F>;
  const V3<F> v = {q.x, q.y, q.z};
  const F n2 = T::add(T::mul(q.w, q.w), xdot(v, v));
  const F inv = T::div(F(1), T::sqrt(n2 == F(0) ? F(1) : n2));
  return {T::mul(q.w, inv), T::mul(q.x, inv), T::mul(q.y, inv), T::mul(q.z, inv)};
}
// Quaternion.from_small_rotvec (quaternion.py:129-141)
template <typename F>
__device__ __forceinline__ Q4<F> xsmall(const V3<F>& rv) {
  using T = RT<F>;
  const F n2 = xdot(rv, rv);
  const F ch = T::sub(F(1), T::div(n2, F(8)));
  const F sc = T::sub(F(0.5), T::div(n2, F(48)));
  return {ch, T::mul(rv.x, sc), T::mul(rv.y, sc), T::mul(rv.z, sc)};
}
// omega_dot (spiral.py:22-40), 3D diagonal inertia
template <typename F>
__device__ __forceinline__ V3<F> xomega_dot(const V3<F>& w, const V3<F>& tq, const V3<F>& I,
                                            const V3<F>& invI) {
  using T = RT<F>;
  const F cx = T::sub(T::mul(w.y, T::mul(I.z, w.z)), T::mul(w.z, T::mul(I.y, w.y)));
  const F cy = T::sub(T::mul(w.z, T::mul(I.x, w.x)), T::mul(w.x, T::mul(I.z, w.z)));
  const F cz = T::sub(T::mul(w.x, T::mul(I.y, w.y)), T::mul(w.y, T::mul(I.x, w.x)));
  return {T::mul(T::sub(tq.x, cx), invI.x), T::mul(T::sub(tq.y, cy), invI.y),
          T::mul(T::sub(tq.z, cz), invI.z)};
}
template <typename F>
__device__ __forceinline__ V3<F> xaxpy(F a, const V3<F>& x, const V3<F>& y) {  // a*x + y, unfused
  using T = RT<F>;
  return {T::add(T::mul(a, x.x), y.x), T::add(T::mul(a, x.y), y.y), T::add(T::mul(a, x.z), y.z)};
}
template <typename F>
__device__ __forceinline__ V3<F> xscale(F a, const V3<F>& x) {
  using T = RT<F>;
  return {T::mul(a, x.x), T::mul(a, x.y), T::mul(a, x.z)};
}
template <typename F>
__device__ __forceinline__ V3<F> xadd(const V3<F>& x, const V3<F>& y) {
  using T = RT<F>;
  return {T::add(x.x, y.x), T::add(x.y, y.y), T::add(x.z, y.z)};
}

// SSPRK3 update of the body-frame angular velocity with step h
// (velocity_verlet_spiral.py:93-99 / spiral.py:129-134): w + free*(k1+k2+4k3)/6.
// `k1_out` returns omega_dot(w) for Spiral's quaternion update.
template <typename F>
__device__ __forceinline__ V3<F> ssprk3(F h, const V3<F>& w, const V3<F>& tq, const V3<F>& I,
                                        const V3<F>& invI, F free, V3<F>* wdot_out) {
  using T = RT<F>;
  const V3<F> wd = xomega_dot(w, tq, I, invI);
  if (wdot_out) *wdot_out = wd;
  const V3<F> k1 = xscale(h, wd);
  const V3<F> k2 = xscale(h, xomega_dot(xadd(w, k1), tq, I, invI));
  const V3<F> k3 = xscale(h, xomega_dot(xadd(w, xscale(F(0.25), xadd(k1, k2))), tq, I, invI));
  const V3<F> s = xadd(xadd(k1, k2), xscale(F(4), k3));
  return {T::add(w.x, T::div(T::mul(free, s.x), F(6))), T::add(w.y, T::div(T::mul(free, s.y), F(6))),
          T::add(w.z, T::div(T::mul(free, s.z), F(6)))};
}

template <typename F>
__device__ __forceinline__ Q4<F> load_q(const Ctx<F>& c, size_t gi) {
  return {c.q_w[gi], c.q_xyz[gi * 3 + 0], c.q_xyz[gi * 3 + 1], c.q_xyz[gi * 3 + 2]};
}
template <typename F>
__device__ __forceinline__ void store_q_and_cache(const Ctx<F>& c, size_t gi, const Q4<F>& q) {
  c.q_w[gi] = q.w;
  c.q_xyz[gi * 3 + 0] = q.x;
  c.q_xyz[gi * 3 + 1] = q.y;
  c.q_xyz[gi * 3 + 2] = q.z;
  // State.__setattr__("q") refreshes _pos_p_rot = rotate(q, pos_p) (state.py:264-273)
  if (c.dim == 3) {
    const V3<F> p = {c.pos_p[gi * 3], c.pos_p[gi * 3 + 1], c.pos_p[gi * 3 + 2]};
    const V3<F> r = xrotate(q, p);
    c.pos_p_rot[gi * 3] = r.x;
    c.pos_p_rot[gi * 3 + 1] = r.y;
    c.pos_p_rot[gi * 3 + 2] = r.z;
  } else {
    F rx, ry;
    xrotate2(q, c.pos_p[gi * 2], c.pos_p[gi * 2 + 1], rx, ry);
    c.pos_p_rot[gi * 2] = rx;
    c.pos_p_rot[gi * 2 + 1] = ry;
  }
}

// One particle of a rotation hook.
// MODE 0: VelocityVerletSpiral.step_before_force (velocity_verlet_spiral.py:83-116)
// MODE 1: VelocityVerletSpiral.step_after_force  (:156-180)
// MODE 2: Spiral.step_after_force                (spiral.py:104-141)
// tov: the torque to use instead of State.torque (the fused epilogue holds it in registers), or NULL.
template <typename F, int MODE>
__device__ __forceinline__ void rotation_update(const Ctx<F>& c, int b, size_t gi, const F* tov) {
  using T = RT<F>;
  const F dt = c.dt[b];
  const F dt_2 = T::div(dt, F(2));
  const F free = c.fixed[gi] ? F(0) : F(1);
  const F h = MODE == 2 ? dt : dt_2;
  Q4<F> q = load_q(c, gi);
  if (c.dim == 3) {
    const V3<F> I = {c.inertia[gi * 3], c.inertia[gi * 3 + 1], c.inertia[gi * 3 + 2]};
    const V3<F> invI = {T::div(F(1), I.x), T::div(F(1), I.y), T::div(F(1), I.z)};
    const V3<F> wl = {c.ang_vel[gi * 3], c.ang_vel[gi * 3 + 1], c.ang_vel[gi * 3 + 2]};
    const V3<F> tl = tov ? V3<F>{tov[0], tov[1], tov[2]} : V3<F>{c.torque[gi * 3], c.torque[gi * 3 + 1], c.torque[gi * 3 + 2]};
    const V3<F> w0 = xrotate_back(q, wl);
    const V3<F> tq = xrotate_back(q, tl);
    V3<F> wdot;
    const V3<F> w = ssprk3(h, w0, tq, I, invI, free, &wdot);
    if (MODE == 0) {
      const Q4<F> dq = xsmall(xscale(T::mul(F(2), dt_2), w));
      q = xqunit(xqmul(q, dq));
      store_q_and_cache(c, gi, q);
    } else if (MODE == 2) {
      const Q4<F> a = xsmall(xscale(T::mul(F(2), dt_2), w0));
      const Q4<F> bq = xsmall(xscale(T::mul(T::mul(F(2), dt_2), dt_2), wdot));
      q = xqunit(xqmul(q, xqmul(a, bq)));
      store_q_and_cache(c, gi, q);
    }
    const V3<F> wn = xrotate(q, w);
    c.ang_vel[gi * 3] = wn.x;
    c.ang_vel[gi * 3 + 1] = wn.y;
    c.ang_vel[gi * 3 + 2] = wn.z;
  } else {
    const F I = c.inertia[gi], invI = T::div(F(1), I);
    const F w0 = c.ang_vel[gi], tq = tov ? tov[0] : c.torque[gi];
    const F wd = T::mul(tq, invI);  // omega_dot 2D: torque * inv_inertia (independent of w)
    const F k = T::mul(h, wd);
    const F s = T::add(T::add(k, k), T::mul(F(4), k));
    const F w = T::add(w0, T::div(T::mul(free, s), F(6)));
    if (MODE == 0) {
      const Q4<F> dq = xsmall(V3<F>{T::mul(F(0), T::mul(T::mul(F(2), dt_2), w)),
                                    T::mul(F(0), T::mul(T::mul(F(2), dt_2), w)),
                                    T::mul(F(1), T::mul(T::mul(F(2), dt_2), w))});
      q = xqunit(xqmul(q, dq));
      store_q_and_cache(c, gi, q);
    } else if (MODE == 2) {
      const F r1 = T::mul(T::mul(F(2), dt_2), w0);
      const F r2 = T::mul(T::mul(T::mul(F(2), dt_2), dt_2), wd);
      const Q4<F> a = xsmall(V3<F>{T::mul(F(0), r1), T::mul(F(0), r1), r1});
      const Q4<F> bq = xsmall(V3<F>{T::mul(F(0), r2), T::mul(F(0), r2), r2});
      q = xqunit(xqmul(q, xqmul(a, bq)));
      store_q_and_cache(c, gi, q);
    }
    c.ang_vel[gi] = w;
  }
}

}  // namespace jdb
