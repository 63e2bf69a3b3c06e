// jaxdem_b200 — pair force laws (spring / hertz / cundallstrack), device side.
//
// Each law returns the force and torque ON body a DUE TO body b, exactly the
// ForceModel.force contract (jaxdem/forces/__init__.py:55-150).  2D bodies are
// embedded in 3D (z = 0, omega = (0, 0, w)); every extra term is an exact zero,
// so the 2D results equal the reference's dimension-specific formulas.
#pragma once
#include "ctx.cuh"

namespace jdb {

template <typename F>
struct Body {
  F x, y, z, r;     // position (pos_c + pos_p_rot), radius
  F vx, vy, vz, m;  // velocity, mass            (cundallstrack only)
  F wx, wy, wz;     // angular velocity           (cundallstrack only)
  int mat;
};

template <typename F>
struct LawCtx {
  F box[3], inv_box[3];
  const F *young, *poisson, *e, *mu, *mu_r, *young_eff;
  F k0;  // young_eff[0, 0], preloaded: the spring stiffness of single-material systems
  // single-material systems: the terms of hertz / cundallstrack that depend on the materials only, evaluated
  // once per thread instead of once per contact (same expressions, same roundings)
  F E0, G0, Es0, beta0, mu0, mur0;
  int nmat;
  bool periodic;
};

// cundall_strack.py:131-141: damping ratio from the restitution coefficient
template <typename F>
__device__ __forceinline__ F cs_beta(F e_eff) {
  using T = RT<F>;
  const F e_safe = e_eff > F(0) ? e_eff : F(1);
  const F ln_e = T::log(e_safe);
  const F pi = F(3.14159265358979323846);
  return e_eff > F(0) ? -ln_e / T::sqrt(pi * pi + ln_e * ln_e) : F(1);
}

template <typename F>
__device__ __forceinline__ LawCtx<F> make_law_ctx(const Ctx<F>& c, int b) {
  LawCtx<F> lc;
  for (int d = 0; d < 3; ++d) {
    lc.box[d] = d < c.dim ? c.box[b * c.dim + d] : F(1);
    lc.inv_box[d] = d < c.dim ? c.inv_box[b * c.dim + d] : F(1);
  }
  const size_t mo = (size_t)b * c.nmat;
  lc.young = c.young ? c.young + mo : nullptr;
  lc.poisson = c.poisson ? c.poisson + mo : nullptr;
  lc.e = c.e ? c.e + mo : nullptr;
  lc.mu = c.mu ? c.mu + mo : nullptr;
  lc.mu_r = c.mu_r ? c.mu_r + mo : nullptr;
  lc.young_eff = c.young_eff ? c.young_eff + mo * c.nmat : nullptr;
  lc.nmat = c.nmat;
  lc.k0 = (c.law == JDB200_LAW_SPRING && c.nmat == 1 && c.young_eff) ? lc.young_eff[0] : F(0);
  lc.periodic = c.periodic;
  lc.E0 = lc.G0 = lc.Es0 = lc.beta0 = lc.mu0 = lc.mur0 = F(0);
  if (c.nmat == 1 && c.law == JDB200_LAW_HERTZ) {
    const F E = lc.young[0], nu = lc.poisson[0];
    lc.Es0 = F(1) / ((F(1) - nu * nu) / E + (F(1) - nu * nu) / E);
  } else if (c.nmat == 1 && c.law == JDB200_LAW_CUNDALLSTRACK) {
    lc.E0 = lc.young[0];
    lc.G0 = lc.E0 / (F(2) * (F(1) + lc.poisson[0]));
    lc.beta0 = cs_beta(RT<F>::fmin(lc.e[0], lc.e[0]));
    lc.mu0 = RT<F>::fmin(lc.mu[0], lc.mu[0]);
    lc.mur0 = RT<F>::fmin(lc.mu_r[0], lc.mu_r[0]);
  }
  return lc;
}

// Domain._displacement: multiply form (jaxdem/domains/periodic.py:75-79).  Unfused
// on purpose: rij - B*round(rij*invB) cancels catastrophically across the
// boundary, and an FMA there would differ from the reference expression.
template <typename F>
__device__ __forceinline__ void displacement_mul(const LawCtx<F>& lc, const Body<F>& a,
                                                 const Body<F>& b, F* r) {
  using T = RT<F>;
  r[0] = T::sub(a.x, b.x);
  r[1] = T::sub(a.y, b.y);
  r[2] = T::sub(a.z, b.z);
  if (lc.periodic) {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      r[d] = T::sub(r[d], T::mul(lc.box[d], T::rint(T::mul(r[d], lc.inv_box[d]))));
  }
}
// Domain.displacement: division form (periodic.py:70-71), used by neighbour lists.
template <typename F>
__device__ __forceinline__ void displacement_div(const LawCtx<F>& lc, const Body<F>& a,
                                                 const Body<F>& b, F* r) {
  using T = RT<F>;
  r[0] = T::sub(a.x, b.x);
  r[1] = T::sub(a.y, b.y);
  r[2] = T::sub(a.z, b.z);
  if (lc.periodic) {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      r[d] = T::sub(r[d], T::mul(lc.box[d], T::rint(T::div(r[d], lc.box[d]))));
  }
}

// unit_and_norm (jaxdem/utils/linalg.py:162-181)
template <typename F>
__device__ __forceinline__ void unit_and_norm(const F* v, F* u, F& nrm) {
  using T = RT<F>;
  const F n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const F inv = n2 == F(0) ? F(0) : T::rsqrt(T::fmax(n2, F(1e-16)));
  nrm = n2 * inv;
  u[0] = v[0] * inv;
  u[1] = v[1] * inv;
  u[2] = v[2] * inv;
}
// norm (linalg.py:113-133)
template <typename F>
__device__ __forceinline__ F safe_norm(const F* v) {
  using T = RT<F>;
  const F n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  return n2 == F(0) ? F(0) : T::sqrt(T::fmax(n2, F(1e-16)));
}

template <typename F>
__device__ __forceinline__ F pow15(F x) { return x * RT<F>::sqrt(x); }   // x^1.5, x >= 0
template <typename F>
__device__ __forceinline__ F pow25(F x) { return x * x * RT<F>::sqrt(x); }  // x^2.5, x >= 0

// ---- force on a due to b --------------------------------------------------------
template <typename F, int LAW>
__device__ __forceinline__ void pair_force_rij(const LawCtx<F>& lc, const Body<F>& a, const Body<F>& b,
                                               const F* rij, F* f, F* t) {
  using T = RT<F>;
  if (LAW == JDB200_LAW_SPRING) {
    // jaxdem/forces/spring.py:97-108
    const F R = a.r + b.r;
    const F d2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
    const F k = lc.nmat == 1 ? lc.k0 : lc.young_eff[a.mat * lc.nmat + b.mat];
    const F inv = d2 == F(0) ? F(0) : T::rsqrt(T::fmax(d2, F(1e-16)));
    const F r = d2 * inv;
    const F delta = T::fmax(F(0), R - r);
    const F mag = k * delta * inv;
    f[0] = mag * rij[0];
    f[1] = mag * rij[1];
    f[2] = mag * rij[2];
    t[0] = t[1] = t[2] = F(0);
  } else if (LAW == JDB200_LAW_HERTZ) {
    // jaxdem/forces/hertz.py:98-119
    const F Rs = (a.r * b.r) / (a.r + b.r);
    F Es = lc.Es0;
    if (lc.nmat != 1) {
      const F Ei = lc.young[a.mat], Ej = lc.young[b.mat];
      const F ni = lc.poisson[a.mat], nj = lc.poisson[b.mat];
      Es = F(1) / ((F(1) - ni * ni) / Ei + (F(1) - nj * nj) / Ej);
    }
    const F k = F(4.0 / 3.0) * Es * T::sqrt(Rs);
    F n[3], r;
    unit_and_norm(rij, n, r);
    const F delta = T::fmax(F(0), a.r + b.r - r);
    const F mag = k * pow15(delta);
    f[0] = mag * n[0];
    f[1] = mag * n[1];
    f[2] = mag * n[2];
    t[0] = t[1] = t[2] = F(0);
  } else {
    // jaxdem/forces/cundall_strack.py:119-198
    F Ei = lc.E0, Ej = lc.E0, Gi = lc.G0, Gj = lc.G0, beta = lc.beta0, mu_eff = lc.mu0, mu_r_eff = lc.mur0;
    if (lc.nmat != 1) {
      Ei = lc.young[a.mat];
      Ej = lc.young[b.mat];
      const F nui = lc.poisson[a.mat], nuj = lc.poisson[b.mat];
      Gi = Ei / (F(2) * (F(1) + nui));
      Gj = Ej / (F(2) * (F(1) + nuj));
      beta = cs_beta(T::fmin(lc.e[a.mat], lc.e[b.mat]));
      mu_eff = T::fmin(lc.mu[a.mat], lc.mu[b.mat]);
      mu_r_eff = T::fmin(lc.mu_r[a.mat], lc.mu_r[b.mat]);
    }
    const F kn = (F(2) * Ei * a.r * Ej * b.r) / (Ei * a.r + Ej * b.r);
    const F kt = (F(2) * Gi * a.r * Gj * b.r) / (Gi * a.r + Gj * b.r);
    const F m_eff = (a.m * b.m) / (a.m + b.m);
    const F gamma_n = F(2) * beta * T::sqrt(kn * m_eff);
    const F gamma_t = F(2) * beta * T::sqrt(kt * m_eff);
    F n[3], r;
    unit_and_norm(rij, n, r);
    F delta = a.r + b.r - r;
    const F is_contact = delta > F(0) ? F(1) : F(0);
    delta *= is_contact;
    const V3<F> rci = {-a.r * n[0], -a.r * n[1], -a.r * n[2]};
    const V3<F> rcj = {b.r * n[0], b.r * n[1], b.r * n[2]};
    const V3<F> wi = {a.wx, a.wy, a.wz}, wj = {b.wx, b.wy, b.wz};
    const V3<F> ci = cross3(wi, rci), cj = cross3(wj, rcj);
    const F vrel[3] = {(a.vx + ci.x) - (b.vx + cj.x), (a.vy + ci.y) - (b.vy + cj.y),
                       (a.vz + ci.z) - (b.vz + cj.z)};
    const F vn = vrel[0] * n[0] + vrel[1] * n[1] + vrel[2] * n[2];
    const F vt_vec[3] = {vrel[0] - vn * n[0], vrel[1] - vn * n[1], vrel[2] - vn * n[2]};
    F tt[3], vt;
    unit_and_norm(vt_vec, tt, vt);
    const F Fn = T::fmax(F(0), kn * delta - gamma_n * vn) * is_contact;
    const F Ft = T::fmin(gamma_t * vt, mu_eff * Fn) * is_contact;
    f[0] = Fn * n[0] - Ft * tt[0];
    f[1] = Fn * n[1] - Ft * tt[1];
    f[2] = Fn * n[2] - Ft * tt[2];
    const V3<F> fv = {f[0], f[1], f[2]};
    const V3<F> tq = cross3(rci, fv);
    const F R_eff = (a.r * b.r) / (a.r + b.r);
    const F orel[3] = {a.wx - b.wx, a.wy - b.wy, a.wz - b.wz};
    const F on2 = orel[0] * orel[0] + orel[1] * orel[1] + orel[2] * orel[2];
    const F oinv = T::rsqrt(on2 == F(0) ? F(1) : on2);  // linalg.unit (linalg.py:136-159)
    const F roll = mu_r_eff * R_eff * Fn;
    t[0] = tq.x - roll * orel[0] * oinv;
    t[1] = tq.y - roll * orel[1] * oinv;
    t[2] = tq.z - roll * orel[2] * oinv;
  }
}

template <typename F, int LAW>
__device__ __forceinline__ void pair_force(const LawCtx<F>& lc, const Body<F>& a, const Body<F>& b,
                                           F* f, F* t) {
  F rij[3];
  displacement_mul(lc, a, b, rij);
  pair_force_rij<F, LAW>(lc, a, b, rij, f, t);
}

// ---- pair energy E_ab (the 0.5 is applied by the caller, _partition.py:39-51) ----
template <typename F, int LAW>
__device__ __forceinline__ F pair_energy_rij(const LawCtx<F>& lc, const Body<F>& a, const Body<F>& b,
                                             const F* rij) {
  using T = RT<F>;
  if (LAW == JDB200_LAW_SPRING) {
    // spring.py:136-147
    const F d2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
    const F k = lc.nmat == 1 ? lc.k0 : lc.young_eff[a.mat * lc.nmat + b.mat];  // (k0: preloaded, as in the force law)
    const F inv = d2 == F(0) ? F(0) : T::rsqrt(T::fmax(d2, F(1e-16)));
    const F s = T::fmax(F(0), a.r + b.r - d2 * inv);
    return F(0.5) * k * s * s;
  } else if (LAW == JDB200_LAW_HERTZ) {
    // hertz.py:146-162
    const F Ei = lc.young[a.mat], Ej = lc.young[b.mat];
    const F ni = lc.poisson[a.mat], nj = lc.poisson[b.mat];
    const F Rs = (a.r * b.r) / (a.r + b.r);
    const F Es = F(1) / ((F(1) - ni * ni) / Ei + (F(1) - nj * nj) / Ej);
    const F k = F(4.0 / 3.0) * Es * T::sqrt(Rs);
    F delta = a.r + b.r - safe_norm(rij);
    delta = delta > F(0) ? delta : F(0);
    return F(0.4) * k * pow25(delta);
  } else {
    // cundall_strack.py:222-235
    const F Ei = lc.young[a.mat], Ej = lc.young[b.mat];
    const F kn = (F(2) * Ei * a.r * Ej * b.r) / (Ei * a.r + Ej * b.r);
    F delta = a.r + b.r - safe_norm(rij);
    delta = delta > F(0) ? delta : F(0);
    return F(0.5) * kn * delta * delta;
  }
}

template <typename F, int LAW>
__device__ __forceinline__ F pair_energy(const LawCtx<F>& lc, const Body<F>& a, const Body<F>& b) {
  F rij[3];
  displacement_mul(lc, a, b, rij);
  return pair_energy_rij<F, LAW>(lc, a, b, rij);
}

}  // namespace jdb
