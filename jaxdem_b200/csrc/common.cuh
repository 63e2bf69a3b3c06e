// jaxdem_b200 — shared device helpers (sm_100a).  See include/jaxdem_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/jaxdem_b200.h"

namespace jdb {

// ---------------------------------------------------------------------------
// Real-number traits.  The *_rn intrinsics are never contracted into FMAs, so
// the integer results derived from them (cell coordinates, hashes) are
// bit-identical to an IEEE evaluation of the reference expressions.
// ---------------------------------------------------------------------------
template <typename F>
struct RT;

template <>
struct RT<float> {
  using I = int32_t;
  using U = uint32_t;
  static constexpr int kDtype = JDB200_F32;
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float floor(float a) { return floorf(a); }
  static __device__ __forceinline__ float ceil(float a) { return ceilf(a); }
  static __device__ __forceinline__ float fmod(float a, float b) { return fmodf(a, b); }
  static __device__ __forceinline__ float trunc(float a) { return truncf(a); }
  static __device__ __forceinline__ float rint(float a) { return rintf(a); }  // half-to-even
  static __device__ __forceinline__ float sqrt(float a) { return sqrtf(a); }
  static __device__ __forceinline__ float rsqrt(float a) { return rsqrtf(a); }
  static __device__ __forceinline__ float log(float a) { return logf(a); }
  static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
  static __device__ __forceinline__ float fmax(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float fmin(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ I to_int(float a) { return __float2int_rz(a); }  // saturating, NaN->0
  static __device__ __forceinline__ float from_int(I a) { return __int2float_rn(a); }
  static __device__ __forceinline__ float imax_as_real() { return 2147483648.0f; }  // float(iinfo(int32).max)
};

template <>
struct RT<double> {
  using I = int64_t;
  using U = uint64_t;
  static constexpr int kDtype = JDB200_F64;
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double floor(double a) { return ::floor(a); }
  static __device__ __forceinline__ double ceil(double a) { return ::ceil(a); }
  static __device__ __forceinline__ double fmod(double a, double b) { return ::fmod(a, b); }
  static __device__ __forceinline__ double trunc(double a) { return ::trunc(a); }
  static __device__ __forceinline__ double rint(double a) { return ::rint(a); }
  static __device__ __forceinline__ double sqrt(double a) { return ::sqrt(a); }
  static __device__ __forceinline__ double rsqrt(double a) { return 1.0 / ::sqrt(a); }
  static __device__ __forceinline__ double log(double a) { return ::log(a); }
  static __device__ __forceinline__ double abs(double a) { return ::fabs(a); }
  static __device__ __forceinline__ double fmax(double a, double b) { return ::fmax(a, b); }
  static __device__ __forceinline__ double fmin(double a, double b) { return ::fmin(a, b); }
  static __device__ __forceinline__ I to_int(double a) { return __double2ll_rz(a); }
  static __device__ __forceinline__ double from_int(I a) { return __ll2double_rn(a); }
  static __device__ __forceinline__ double imax_as_real() { return 9223372036854775808.0; }
};

// 4-wide vector of F (16 B for float, 32 B for double), the sorted-shadow record.
template <typename F>
struct alignas(4 * sizeof(F)) Vec4 {
  F x, y, z, w;
};

template <typename F>
struct V3 {
  F x, y, z;
};

template <typename F>
__device__ __forceinline__ V3<F> cross3(const V3<F>& a, const V3<F>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename F>
__device__ __forceinline__ F dot3(const V3<F>& a, const V3<F>& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z;
}

// Quaternion.rotate (jaxdem/utils/quaternion.py:190-242), 3D.
template <typename F>
__device__ __forceinline__ V3<F> q_rotate3(F w, const V3<F>& q, const V3<F>& v) {
  V3<F> t = cross3(q, v);
  t.x *= F(2);
  t.y *= F(2);
  t.z *= F(2);
  V3<F> u = cross3(q, t);
  return {v.x + w * t.x + u.x, v.y + w * t.y + u.y, v.z + w * t.z + u.z};
}
template <typename F>
__device__ __forceinline__ V3<F> q_rotate_back3(F w, const V3<F>& q, const V3<F>& v) {
  V3<F> qc = {-q.x, -q.y, -q.z};
  return q_rotate3(w, qc, v);
}

// ---------------------------------------------------------------------------
// Grid description of one system, derived ON DEVICE from box/cell_size each
// call (_grid_params, jaxdem/colliders/_partition.py:54-99).
// ---------------------------------------------------------------------------
template <typename I>
struct GridInfo {
  I gd[3];
  I stride[3];
  long long bound;   // dense table rows in use: every particle hash must lie in [0, bound)
  int hash_overflow; // reference's hash_overflow flag
  int need_dedup;    // periodic and stencil rows can collide after the wrap
  int dense;         // 1: dense cell table valid for this call, 0: sorted/binary search
  int dense_fail;    // set by kernels when the dense strategy cannot represent the data
  int nl_overflow;   // neighbour-list overflow accumulator
  int canonical;     // neighbor_mask == meshgrid(-R..R)^D ("ij", last axis fastest): x-run walk is legal
  int range;         // R of the canonical stencil
  int any_bond;      // some particle has a bond_id >= 0 (set by the hash kernel)
  int any_ppr;       // some particle has a non-zero _pos_p_rot (set by the hash kernel)
  int any_ext;       // some external force / torque buffer entry is non-zero (fused hash kernel)
  int any_fixed;     // some particle is fixed (fused hash kernel)
  int edge;          // some cell coordinate lies outside [0, g) (periodic: rounded up to g), so its hash aliases another cell: keys cannot be decoded
  int hashed;        // 1: the grid has more cells than the table has rows — the table is addressed by a HASH of the cell
                     //    key (rows = 2^(32 - hshift)); a row may hold several cells, kept sorted by (key, index), and
                     //    every look-up filters by the exact key.  Same pair sets as the sorted strategy, no sort.
  int hshift;
};

// row of the cell table that holds cell `key`
template <typename I>
__device__ __forceinline__ long long table_row(const GridInfo<I>& g, long long key) {
  return g.hashed ? (long long)(((unsigned)key * 2654435761u) >> g.hshift) : key;
}

template <typename F, typename I>
__device__ __forceinline__ void grid_dims(const F* __restrict__ box, F cell_size, int dim,
                                          bool periodic, I* gd, I* stride, int* overflow) {
  using T = RT<F>;
  F total = F(1);
  for (int d = 0; d < dim; ++d) {
    F q = T::div(box[d], cell_size);
    I g = T::to_int(periodic ? T::floor(q) : T::ceil(q));
    g = g < I(1) ? I(1) : g;
    gd[d] = g;
    total = T::mul(total, T::from_int(g));
  }
  typename T::U s = 1;
  for (int d = 0; d < dim; ++d) {
    stride[d] = (I)s;
    s = s * (typename T::U)gd[d];  // wrapping cumprod, like the int cumprod of the reference
  }
  for (int d = dim; d < 3; ++d) {
    gd[d] = 1;
    stride[d] = 0;
  }
  *overflow = total > T::imax_as_real();
}

// Cell coordinate of one axis (jaxdem/colliders/cell_list.py:55-60).
template <typename F, typename I>
__device__ __forceinline__ I cell_coord(F x, F anchor, F box, F cell_size, I g, bool periodic) {
  using T = RT<F>;
  if (periodic) {
    F u = T::div(T::sub(x, anchor), box);
    // jnp.remainder(u, 1) = fmod(u, 1) + sign fix-up; fmod(u, 1) == u - trunc(u) EXACTLY for
    // every finite u (trunc(u) is exact and the difference is representable), NaN for inf/NaN
    F r = T::sub(u, T::trunc(u));
    if (r != F(0) && r < F(0)) r = T::add(r, F(1));
    return T::to_int(T::floor(T::mul(r, T::from_int(g))));
  }
  return T::to_int(T::floor(T::div(T::sub(x, anchor), cell_size)));
}

// Neighbour-cell hash for stencil row `off` (cell_list.py:66-80).  Returns -1
// for out-of-grid cells of non-periodic domains.
template <typename F, typename I>
__device__ __forceinline__ I neighbor_hash(const I* c, const I* off, const I* gd, const I* stride,
                                           int dim, bool periodic) {
  using T = RT<F>;
  typename T::U h = 0;
  bool oob = false;
  for (int d = 0; d < dim; ++d) {
    I nc = c[d] + off[d];
    if (periodic) {
      nc -= gd[d] * T::to_int(T::floor(T::div(T::from_int(nc), T::from_int(gd[d]))));
    } else {
      oob |= (nc < 0) | (nc >= gd[d]);
    }
    h += (typename T::U)nc * (typename T::U)stride[d];
  }
  return oob ? I(-1) : (I)h;
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

struct Bump {  // workspace bump allocator (256-B aligned), shared by sizing and use
  char* base;
  size_t off;
  explicit Bump(void* b) : base((char*)b), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? (T*)(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

}  // namespace jdb
