// jaxdem_b200 — pair traversal kernels (K4 force, K5 energy, K6 neighbour list)
// over the cell partition, plus the O(N^2) naive collider.
//
// Replaces _traverse_pairs / compute_force / compute_potential_energy /
// create_neighbor_list (jaxdem/colliders/cell_list.py:187-261,434-595) and
// NaiveSimulator (jaxdem/colliders/naive.py:73-235).  One thread owns one
// particle and walks its stencil rows in neighbor_mask order and each cell's run
// in sorted order: a fixed summation order, no atomics => deterministic.
#include "laws.cuh"
#include "launch.cuh"
#include "rot.cuh"

namespace jdb {

template <typename I>
__device__ __forceinline__ bool use_dense(const GridInfo<I>& g) {
  return g.dense && !g.dense_fail;
}

template <typename F>
__device__ __forceinline__ Body<F> load_sorted(const Ctx<F>& c, size_t off, int k, bool cs) {
  Body<F> a;
  const Vec4<F> p = c.spos[off + k];
  a.x = p.x; a.y = p.y; a.z = p.z; a.r = p.w;
  a.mat = (c.nmat > 1) ? c.smat[off + k] : 0;
  if (cs) {
    const Vec4<F> v = c.svel[off + k];
    const Vec4<F> w = c.sang[off + k];
    a.vx = v.x; a.vy = v.y; a.vz = v.z; a.m = v.w;
    a.wx = w.x; a.wy = w.y; a.wz = w.z;
  }
  return a;
}

template <typename F>
__device__ __forceinline__ Body<F> load_orig(const Ctx<F>& c, size_t gidx, bool cs) {
  using T = RT<F>;
  Body<F> a;
  const F* pc = c.pos_c + gidx * c.dim;
  const F* pr = c.pos_p_rot + gidx * c.dim;
  a.x = T::add(pc[0], pr[0]);
  a.y = T::add(pc[1], pr[1]);
  a.z = c.dim == 3 ? T::add(pc[2], pr[2]) : F(0);
  a.r = c.rad[gidx];
  a.mat = (int)c.mat_id[gidx];
  if (cs) {
    const F* v = c.vel + gidx * c.dim;
    a.vx = v[0]; a.vy = v[1]; a.vz = c.dim == 3 ? v[2] : F(0);
    a.m = c.mass[gidx];
    const F* w = c.ang_vel + gidx * c.A;
    if (c.dim == 3) { a.wx = w[0]; a.wy = w[1]; a.wz = w[2]; }
    else { a.wx = F(0); a.wy = F(0); a.wz = w[0]; }
  }
  return a;
}

// valid_interaction_mask as called from the cell list (colliders/__init__.py:225-243,
// cell_list.py:240-246): candidate's clump / bond row against the owner's clump / index.
template <typename F>
__device__ __forceinline__ bool pair_valid(const Ctx<F>& c, size_t off, int owner_idx, int owner_clump,
                                           int cand_sclump, int cand_slot, bool interact) {
  if ((cand_sclump & 0x7fffffff) == owner_clump) return false;
  if ((cand_sclump & 0x80000000) && !interact) {
    const int j = c.perm[off + cand_slot];
    const typename RT<F>::I* row = c.bond_id + (off + j) * c.W;
    for (int w = 0; w < c.W; ++w)
      if (row[w] == owner_idx) return false;
  }
  return true;
}

// Walk the stencil of sorted slot k: calls vis.cell(m, start, end) for every stencil
// row whose target cell is looked up (rows removed by the periodic de-dup are skipped;
// the reference turns them into hash -1, which no periodic cell carries).
// MultiCellList (multi_cell_list.py:186-204, 475-481): does the query box overlap the expandable AABB of the cell
// whose sorted run starts at slot s?  dr = Domain.displacement(query centre, cell centre) (division form).
template <typename F>
__device__ __forceinline__ bool aabb_overlap(const Ctx<F>& c, int b, size_t off, int s, const F* qc, const F* qh) {
  using T = RT<F>;
  const Vec4<F> cc = c.aabb_c[off + s], hh = c.aabb_h[off + s];
  const F cv[3] = {cc.x, cc.y, cc.z}, hv[3] = {hh.x, hh.y, hh.z};
  for (int d = 0; d < c.dim; ++d) {
    F dr = T::sub(qc[d], cv[d]);
    if (c.periodic) {
      const F bx = c.box[b * c.dim + d];
      dr = T::sub(dr, T::mul(bx, T::rint(T::div(dr, bx))));
    }
    if (!(T::abs(dr) <= T::add(qh[d], hv[d]))) return false;
  }
  return true;
}

template <typename F, typename Vis>
__device__ __forceinline__ void walk_stencil_at(const Ctx<F>& c, int b, const F* pp, const F* cell_size_override,
                                                Vis& vis, F qrad = F(-1)) {
  using I = typename RT<F>::I;
  using T = RT<F>;
  // loose-grid pruning: query box = the particle's own box [pos - rad, pos + rad] (force / energy), or the point
  // +- cutoff (neighbour-list builds)
  F qc[3] = {0, 0, 0}, qh[3] = {0, 0, 0};
  bool prune = false;
  if (c.prune) {
    if (c.prune_cut) {
      prune = true;
      for (int d = 0; d < c.dim; ++d) { qc[d] = pp[d]; qh[d] = c.prune_cut[b]; }
    } else if (qrad >= F(0)) {
      prune = true;
      for (int d = 0; d < c.dim; ++d) {
        const F lo = T::sub(pp[d], qrad), hi = T::add(pp[d], qrad);
        qc[d] = T::mul(F(0.5), T::add(lo, hi));
        qh[d] = T::mul(F(0.5), T::sub(hi, lo));
      }
    }
  }
  const GridInfo<I> g = c.gi[b];
  const size_t off = (size_t)b * c.n;
  const bool dense = use_dense(g);
  const F cs = cell_size_override ? cell_size_override[b] : c.cell_size[b];
  I cc[3] = {0, 0, 0};
  for (int d = 0; d < c.dim; ++d)
    cc[d] = cell_coord<F, I>(pp[d], c.anchor[b * c.dim + d], c.box[b * c.dim + d], cs, g.gd[d],
                             c.periodic);
  const I* mask = c.mask + (size_t)b * c.M * c.dim;
  const I* skey = c.skey + off;
  const int* cstart = c.cell_start + (size_t)b * c.cell_stride;
  const int n = (int)(c.n_dev ? *c.n_dev : c.n);
  for (int m = 0; m < c.M; ++m) {
    const I h = neighbor_hash<F, I>(cc, mask + m * c.dim, g.gd, g.stride, c.dim, c.periodic);
    if (c.periodic && g.need_dedup) {
      bool dup = false;
      for (int m2 = 0; m2 < m && !dup; ++m2)
        dup = neighbor_hash<F, I>(cc, mask + m2 * c.dim, g.gd, g.stride, c.dim, c.periodic) == h;
      if (dup) continue;
    }
    int s, e;
    if (dense && g.hashed) {  // the row of the hashed table, narrowed to the run of cell h (rows are sorted by key)
      const int hk = (int)h;
      const long long row = table_row(g, (long long)hk);
      const int* tk = c.tmp_key + off;
      s = cstart[row];
      const int re = cstart[row + 1];
      while (s < re && tk[s] < hk) ++s;
      e = s;
      while (e < re && tk[e] == hk) ++e;
    } else if (dense) {
      if (h < 0 || (long long)h >= g.bound) continue;
      s = cstart[h];
      e = cstart[h + 1];
    } else {
      int lo = 0, hi = n;  // lower_bound: searchsorted(side="left") (cell_list.py:227-229)
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (skey[mid] < h) lo = mid + 1; else hi = mid;
      }
      s = lo;
      hi = n;              // upper_bound: end of the run with key == h
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (skey[mid] <= h) lo = mid + 1; else hi = mid;
      }
      e = lo;
    }
    if (e > s && (!prune || aabb_overlap(c, b, off, s, qc, qh))) vis.cell(m, s, e);
  }
}

template <typename F, typename Vis>
__device__ __forceinline__ void walk_stencil(const Ctx<F>& c, int b, int k, const F* cell_size_override,
                                             Vis& vis) {
  const Vec4<F> p = c.spos[(size_t)b * c.n + k];
  const F pp[3] = {p.x, p.y, p.z};
  walk_stencil_at<F>(c, b, pp, cell_size_override, vis, p.w);
}

// Hashed table + the default 3^D stencil, every particle inside the grid, no de-dup, no pruning: the cells in mask
// order (last axis fastest) with their keys taken from the strides — a short loop instead of the general walk above
// (mask loads, wrapped hashes, de-dup and prune tests per cell).  Used by the force kernels (the dilute systems whose
// grid outgrows the dense table spend their time there); energy and neighbour-list builds keep the general walk.
template <typename I>
__device__ __forceinline__ bool hashed_walk_ok(const GridInfo<I>& g) {
  return g.dense && !g.dense_fail && g.hashed && g.canonical && g.range == 1 && !g.need_dedup && !g.edge;
}
template <typename F, typename Vis>
__device__ __forceinline__ void walk_hashed(const Ctx<F>& c, int b, const GridInfo<typename RT<F>::I>& g, const F* pp,
                                            Vis& vis) {
  using I = typename RT<F>::I;
  const size_t off = (size_t)b * c.n;
  const F cs = c.cell_size[b];
  I cc[3] = {0, 0, 0};
  for (int d = 0; d < c.dim; ++d)
    cc[d] = cell_coord<F, I>(pp[d], c.anchor[b * c.dim + d], c.box[b * c.dim + d], cs, g.gd[d], c.periodic);
  const int* cst = c.cell_start + (size_t)b * c.cell_stride;
  const int* tk = c.tmp_key + off;
  const int D = c.dim;
  const long long g0 = (long long)g.gd[0], g1 = (long long)g.gd[1], g2 = D == 3 ? (long long)g.gd[2] : 1;
  const long long s0 = (long long)g.stride[0], s1 = (long long)g.stride[1], s2 = D == 3 ? (long long)g.stride[2] : 0;
  const int M = D == 3 ? 27 : 9;
#pragma unroll 1
  for (int m = 0; m < M; ++m) {
    // offsets of stencil row m: digits of m in base 3, last axis fastest
    const int o0 = (D == 3 ? m / 9 : m / 3) - 1, o1 = (D == 3 ? (m / 3) % 3 : m % 3) - 1, o2 = D == 3 ? m % 3 - 1 : 0;
    long long n0 = (long long)cc[0] + o0, n1 = (long long)cc[1] + o1, n2 = (long long)cc[2] + o2;
    if (c.periodic) {
      n0 = n0 < 0 ? n0 + g0 : (n0 >= g0 ? n0 - g0 : n0);
      n1 = n1 < 0 ? n1 + g1 : (n1 >= g1 ? n1 - g1 : n1);
      n2 = n2 < 0 ? n2 + g2 : (n2 >= g2 ? n2 - g2 : n2);
    } else if (n0 < 0 || n0 >= g0 || n1 < 0 || n1 >= g1 || n2 < 0 || n2 >= g2) {
      continue;  // outside a non-periodic grid: no particle lives there (no edge particles in this branch)
    }
    const int hk = (int)(n0 * s0 + n1 * s1 + n2 * s2);
    const long long row = table_row(g, (long long)hk);
    int s = cst[row];
    const int re = cst[row + 1];
    while (s < re && tk[s] < hk) ++s;
    int e = s;
    while (e < re && tk[e] == hk) ++e;
    if (e > s) vis.cell(m, s, e);
  }
}

// ---------------------------------------------------------------------------
// Fast stencil walk (dense table, canonical cubic stencil, no periodic de-dup needed):
// the x-fastest linear hash makes the cells (cx-R .. cx+R, ny, nz) ONE contiguous run of
// the sorted arrays (two where the run wraps around the periodic box), so a (2R+1)^D
// stencil costs (2R+1)^(D-1) range look-ups.  The set of cells visited is exactly the
// reference's stencil (same wrapped / out-of-grid rules, cell_list.py:66-80); only the
// ORDER differs from neighbor_mask order (x innermost), which changes floating-point
// summation order but nothing else.  Row-exact consumers (the neighbour list) keep
// walk_stencil.
// ---------------------------------------------------------------------------
template <typename I>
__device__ __forceinline__ bool fast_walk_ok(const GridInfo<I>& g) {
  return g.dense && !g.dense_fail && !g.hashed && g.canonical && !g.need_dedup;  // hashed rows: adjacent cells are not adjacent rows
}

// the flat kernel (k_pair_flat) additionally wants the default stencil range R = 1
template <typename I>
__device__ __forceinline__ bool flat_walk_ok(const GridInfo<I>& g) {
  return fast_walk_ok(g) && g.range == 1 && !g.edge;
}

// the row kernel (k_pair_rows) packs a run start and its length into one word: starts below 2^27
template <typename F>
__device__ __forceinline__ bool rows_ok(const Ctx<F>& c, const GridInfo<typename RT<F>::I>& g) {
  return flat_walk_ok(g) && c.n < (1ll << 27) && !c.prune;  // MultiCellList: the generic walk prunes per cell
}

// wrap n into [0, g) for |n| < 2g: identical to n - g*floor(n/g) (cell_list.py:70-72)
__device__ __forceinline__ int wrap1(int n, int g) { return n < 0 ? n + g : (n >= g ? n - g : n); }

// Calls vis.run(s1, e1, s2, e2) once per stencil row: candidates are the sorted slots
// [s1, e1) followed by [s2, e2) (the second segment is empty unless the x-run wraps).
template <typename F, int D, bool PERIODIC, typename Vis>
__device__ __forceinline__ void walk_runs(const Ctx<F>& c, int b, const GridInfo<typename RT<F>::I>& g,
                                          const F* pp, Vis& vis) {
  using I = typename RT<F>::I;
  const F cs = c.cell_size[b];
  int cc[3] = {0, 0, 0}, gd[3] = {1, 1, 1};
#pragma unroll
  for (int d = 0; d < D; ++d) {
    cc[d] = (int)cell_coord<F, I>(pp[d], c.anchor[b * D + d], c.box[b * D + d], cs, g.gd[d], PERIODIC);
    gd[d] = (int)g.gd[d];
  }
  const int R = g.range, len = 2 * R + 1;
  const int sy = (int)g.stride[1], sz = D == 3 ? (int)g.stride[2] : 0;
  const int* __restrict__ cstart = c.cell_start + (size_t)b * c.cell_stride;
  // x segment(s): [x1, x1 + n1) and, when it wraps, [0, n2)
  int x1, n1, n2 = 0;
  if (PERIODIC) {
    x1 = wrap1(cc[0] - R, gd[0]);
    n1 = min(len, gd[0] - x1);
    n2 = len - n1;
  } else {
    x1 = max(cc[0] - R, 0);
    n1 = min(cc[0] + R, gd[0] - 1) - x1 + 1;  // <= 0: the whole x-range is out of the grid
  }
  const int zlo = D == 3 ? -R : 0, zhi = D == 3 ? R : 0;
  for (int dz = zlo; dz <= zhi; ++dz) {
    int nz = cc[2] + dz;
    if (D == 3) {
      if (PERIODIC) nz = wrap1(nz, gd[2]);
      else if (nz < 0 || nz >= gd[2]) continue;
    }
    for (int dy = -R; dy <= R; ++dy) {
      int ny = cc[1] + dy;
      if (PERIODIC) ny = wrap1(ny, gd[1]);
      else if (ny < 0 || ny >= gd[1]) continue;
      const int hb = ny * sy + nz * sz;
      int s1 = 0, e1 = 0, s2 = 0, e2 = 0;
      if (n1 > 0) {
        s1 = cstart[hb + x1];
        e1 = cstart[hb + x1 + n1];
      }
      if (PERIODIC && n2 > 0) {
        s2 = cstart[hb];
        e2 = cstart[hb + n2];
      }
      vis.run(s1, e1, s2, e2);
    }
  }
}

// ---------------------------------------------------------------------------
// K4  pair force.  One thread owns one particle (sorted slot k) and accumulates its
// contacts in a fixed order: deterministic, no atomics.
// ---------------------------------------------------------------------------
template <typename F, int LAW, int D, bool PERIODIC, bool SIMPLE>
struct ForceVis {
  using T = RT<F>;
  const Ctx<F>& c;
  const LawCtx<F>& lc;
  size_t off;
  const Vec4<F>* __restrict__ sp;  // c.spos + off
  Body<F> a;
  int k, idx, clump;
  bool interact;
  F hb[3];  // |rij| below this => the minimum-image term is exactly zero
  F f[3], t[3];
  F e = F(0);  // c.want_energy: this particle's share 0.5 * sum_j E_ij (non-overlapping pairs carry exactly 0)

  __device__ __forceinline__ void one(int kj, const Vec4<F>& q) {
    constexpr bool CS = LAW == JDB200_LAW_CUNDALLSTRACK;
    if (!SIMPLE) {
      const int sc = c.sclump[off + kj];
      if (!pair_valid(c, off, idx, clump, sc, kj, interact)) return;
    }
    F rij[3] = {T::sub(a.x, q.x), T::sub(a.y, q.y), D == 3 ? T::sub(a.z, q.z) : F(0)};
    if (PERIODIC) {
      // Domain._displacement (periodic.py:75-79); rint(rij * inv_box) is exactly 0 below hb
      bool far = false;
#pragma unroll
      for (int d = 0; d < D; ++d) far |= !(T::abs(rij[d]) < hb[d]);
      if (far) {
#pragma unroll
        for (int d = 0; d < D; ++d)
          rij[d] = T::sub(rij[d], T::mul(lc.box[d], T::rint(T::mul(rij[d], lc.inv_box[d]))));
      }
    }
    const F d2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
    const F rs = a.r + q.w;
    // no overlap => every law returns exactly zero force and torque (the margin keeps
    // pairs within rounding of touching on the full path); SIMPLE: clump_id == arange(N),
    // only the particle itself is excluded
    if (!(d2 < rs * rs * F(1.00001)) || (SIMPLE && kj == k)) return;
    Body<F> bj;
    bj.x = q.x; bj.y = q.y; bj.z = q.z; bj.r = q.w;
    bj.mat = (c.nmat > 1) ? c.smat[off + kj] : 0;
    if (CS) {
      const Vec4<F> v = c.svel[off + kj];
      const Vec4<F> w = c.sang[off + kj];
      bj.vx = v.x; bj.vy = v.y; bj.vz = v.z; bj.m = v.w;
      bj.wx = w.x; bj.wy = w.y; bj.wz = w.z;
    }
    F ff[3], tt[3];
    pair_force_rij<F, LAW>(lc, a, bj, rij, ff, tt);
    f[0] += ff[0]; f[1] += ff[1]; f[2] += ff[2];
    if (CS) { t[0] += tt[0]; t[1] += tt[1]; t[2] += tt[2]; }
    if (c.want_energy) e += F(0.5) * pair_energy_rij<F, LAW>(lc, a, bj, rij);
  }
  // two-segment iterator: [s1, e1) then [s2, e2)
  __device__ __forceinline__ void run(int s1, int e1, int s2, int e2) {
    int kj = s1, e = e1;
    while (true) {
      if (kj == e) {
        if (s2 == e2) break;
        kj = s2; e = e2; s2 = e2;
        continue;
      }
      one(kj, sp[kj]);
      ++kj;
    }
  }
  __device__ __forceinline__ void cell(int, int s, int e) {  // general walk
    for (int kj = s; kj < e; ++kj) one(kj, sp[kj]);
  }
};

// epilogue of DynamicCellList.compute_force (cell_list.py:461-462):
// torque = sum T + cross(_pos_p_rot, sum F).  with_torque = false skips the torque store
// (fused driver, steps whose torque nobody can observe).
template <typename F, int D>
__device__ __forceinline__ void store_force_torque(const Ctx<F>& c, size_t gidx, const F* f, const F* t,
                                                   bool any_ppr, bool with_torque) {
  F* fo = c.force + gidx * D;
#pragma unroll
  for (int d = 0; d < D; ++d) fo[d] = f[d];
  if (!with_torque) return;
  constexpr int A = D == 3 ? 3 : 1;
  F* to = c.torque + gidx * A;
  F pr[3] = {0, 0, 0};
  if (any_ppr) {
#pragma unroll
    for (int d = 0; d < D; ++d) pr[d] = c.pos_p_rot[gidx * D + d];
  }
  if (D == 3) {
    to[0] = t[0] + (pr[1] * f[2] - pr[2] * f[1]);
    to[1] = t[1] + (pr[2] * f[0] - pr[0] * f[2]);
    to[2] = t[2] + (pr[0] * f[1] - pr[1] * f[0]);
  } else {
    to[0] = t[2] + (pr[0] * f[1] - pr[1] * f[0]);
  }
}

// Fused sphere driver (abi.cu system_step): what follows the collider inside _step_once for a
// sphere system (clump_id == arange(N)) with velocity Verlet and no rotation integrator, done
// by the thread that owns the particle while the contact force is still in registers:
//   DynamicCellList.compute_force epilogue  torque = sum T + cross(_pos_p_rot, sum F)   (cell_list.py:461-462)
//   ForceManager.apply with count == 1      (force_manager.py:359-423)
//   VelocityVerlet.step_after_force         (velocity_verlet.py:92-95)
// Same expressions, same order per particle as k_fm_spheres + k_linear.  The kicked velocity
// and the mass (vm) come from the sorted shadow record written by k_hash<MODE 3>; external
// buffers and fixed flags are gathered only when that kernel saw a non-zero entry (they are
// zero otherwise, and stay zero).
template <typename F, int D>
__device__ __forceinline__ void fused_sphere_compute(const Ctx<F>& c, int b, const GridInfo<typename RT<F>::I>& g,
                                                     const Vec4<F>& vm, size_t gi, const F* f, const F* t,
                                                     bool with_torque, F* o_force, F* o_vel, F* o_torque) {
  using T = RT<F>;
  constexpr int A = D == 3 ? 3 : 1;
  const F v[3] = {vm.x, vm.y, vm.z}, mass = vm.w;
  const F dt = c.dt[b];
  F fp[3] = {0, 0, 0}, fc[3] = {0, 0, 0}, r[3] = {0, 0, 0}, et[3] = {0, 0, 0}, grav[3] = {0, 0, 0};
  F free = F(1);
  if (g.any_fixed) free = c.fixed[gi] ? F(0) : F(1);
#pragma unroll
  for (int d = 0; d < D; ++d) grav[d] = c.gravity[b * D + d];
  if (g.any_ppr) {
#pragma unroll
    for (int d = 0; d < D; ++d) r[d] = c.pos_p_rot[gi * D + d];
  }
  if (g.any_ext) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      fp[d] = c.ext_force[gi * D + d];
      fc[d] = c.ext_force_com[gi * D + d];
    }
#pragma unroll
    for (int a = 0; a < A; ++a) et[a] = c.ext_torque[gi * A + a];
  }
  const F sc = T::div(T::mul(dt, F(0.5)), mass);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const F fcom = T::add(fc[d], T::mul(grav[d], T::div(mass, F(1))));
    const F ft = T::add(T::add(f[d], fp[d]), fcom);
    o_force[d] = ft;
    o_vel[d] = T::add(v[d], T::mul(T::mul(ft, sc), free));
  }
  if (g.any_ext) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      c.ext_force[gi * D + d] = F(0);
      c.ext_force_com[gi * D + d] = F(0);
    }
#pragma unroll
    for (int a = 0; a < A; ++a) c.ext_torque[gi * A + a] = F(0);
  }
  if (!with_torque) return;
  if (D == 3) {
    // collider: t + cross(r, f) (contracted like store_force_torque); manager: unfused cross(r, fp)
    const F tc[3] = {t[0] + (r[1] * f[2] - r[2] * f[1]), t[1] + (r[2] * f[0] - r[0] * f[2]),
                     t[2] + (r[0] * f[1] - r[1] * f[0])};
    const F cr[3] = {T::sub(T::mul(r[1], fp[2]), T::mul(r[2], fp[1])), T::sub(T::mul(r[2], fp[0]), T::mul(r[0], fp[2])),
                     T::sub(T::mul(r[0], fp[1]), T::mul(r[1], fp[0]))};
#pragma unroll
    for (int a = 0; a < 3; ++a) o_torque[a] = T::add(tc[a], T::add(et[a], cr[a]));
  } else {
    const F tc = t[2] + (r[0] * f[1] - r[1] * f[0]);
    const F cr = T::sub(T::mul(r[0], fp[1]), T::mul(r[1], fp[0]));
    o_torque[0] = T::add(tc, T::add(et[0], cr));
  }
}

template <typename F, int D>
__device__ __forceinline__ void fused_sphere_epilogue(const Ctx<F>& c, int b,
                                                      const GridInfo<typename RT<F>::I>& g, const Vec4<F>& vm,
                                                      int idx, const F* f, const F* t, bool with_torque) {
  constexpr int A = D == 3 ? 3 : 1;
  const size_t gi = (size_t)b * c.n + idx;
  F of[3], ov[3], ot[3];
  fused_sphere_compute<F, D>(c, b, g, vm, gi, f, t, with_torque, of, ov, ot);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    c.force[gi * D + d] = of[d];
    c.vel[gi * D + d] = ov[d];
  }
  if (with_torque) {
#pragma unroll
    for (int a = 0; a < A; ++a) c.torque[gi * A + a] = ot[a];
  }
  // fused == 2: the rotation integrator's step_after_force on the same particle, with the torque still in
  // registers (velocity_verlet_spiral.py:156-180 / spiral.py:104-141; callers pass with_torque = true)
  if (c.fused == 2) {
    if (c.rot == JDB200_ROT_VERLETSPIRAL) rotation_update<F, 1>(c, b, gi, ot);
    else rotation_update<F, 2>(c, b, gi, ot);
  }
}

template <typename F, int LAW, int D, bool PERIODIC, bool SIMPLE, int EPI>
__device__ __forceinline__ void pair_force_body(const Ctx<F>& c, int b, int k,
                                                const GridInfo<typename RT<F>::I>& g, bool fast,
                                                int with_torque) {
  const size_t off = (size_t)b * c.n;
  const LawCtx<F> lc = make_law_ctx(c, b);
  ForceVis<F, LAW, D, PERIODIC, SIMPLE> vis{c, lc, off, c.spos + off};
  vis.a = load_sorted(c, off, k, LAW == JDB200_LAW_CUNDALLSTRACK);
  vis.k = k;
  vis.idx = c.perm[off + k];
  vis.clump = SIMPLE ? 0 : (c.sclump[off + k] & 0x7fffffff);
  vis.interact = c.interact && c.interact[b];
#pragma unroll
  for (int d = 0; d < 3; ++d) vis.hb[d] = lc.box[d] * F(0.499999);
  vis.f[0] = vis.f[1] = vis.f[2] = F(0);
  vis.t[0] = vis.t[1] = vis.t[2] = F(0);
  if (fast) {
    const F pp[3] = {vis.a.x, vis.a.y, vis.a.z};
    walk_runs<F, D, PERIODIC>(c, b, g, pp, vis);
  } else if (hashed_walk_ok(g) && !c.prune) {
    const F pp[3] = {vis.a.x, vis.a.y, vis.a.z};
    walk_hashed<F>(c, b, g, pp, vis);
  } else {
    walk_stencil<F>(c, b, k, nullptr, vis);
  }
  if (EPI == 1) fused_sphere_epilogue<F, D>(c, b, g, c.uvm[off + vis.idx], vis.idx, vis.f, vis.t, with_torque != 0);
  else store_force_torque<F, D>(c, off + vis.idx, vis.f, vis.t, g.any_ppr != 0, with_torque != 0);
}


// ---------------------------------------------------------------------------
// K4 main path: the row kernel.  Serves the systems whose partition is a dense table with
// the default 3^D stencil (flat_walk_ok).  One thread owns one particle (sorted slot k).
//
// What bounds this path on B200 is (i) instruction issue and (ii) the SM's single L1TEX
// pipe, which spends ~2 cycles per 128-byte line a warp-wide access touches: a fully
// scattered 4-byte store costs a warp as much as 32 coalesced ones.  Hence:
//   A  the 3^(D-1) stencil rows become slot ranges (x-runs of up to three cells);
//   B  the first kU = 4 slots of every run are loaded and tested WITHOUT a loop — twelve
//      independent 16-byte loads in flight per z-plane, ~12 instructions per candidate, hits
//      recorded as bits of a 36-bit mask; the rare longer runs and the cells that wrap around
//      the periodic box in x go through a small per-thread list in shared memory;
//   C  the force law runs only on the hits (exact reference arithmetic, min-image included),
//      in a fixed order: deterministic, no atomics, no barriers;
//   D  the sums are stored in SORTED order (one coalesced 16-byte store per thread);
//      k_after, one thread per ORIGINAL particle, fetches them through the inverse
//      permutation with one 16-byte gather and finishes the hook (collider epilogue, or
//      the fused force manager + step_after_force) with coalesced State stores.
// ---------------------------------------------------------------------------
#ifndef JDB_ROWS_MINB
#define JDB_ROWS_MINB 10  // resident CTAs per SM the f32 row kernel is compiled for (spring / hertz: 48 registers, no spills)
#endif
template <int D>
struct RowsCfg {
  static constexpr int kThreads = 128;
  static constexpr int kRows = D == 3 ? 9 : 3;
  static constexpr int kU = 4;       // slots per stencil row tested without a loop
  static constexpr int kExtra = 12;  // per-thread list: run tails beyond kU and wrapped cells
  static constexpr int kOffExtra = kRows * kThreads * 4;           // u32 [kRows][kThreads] run starts at 0
  static constexpr int kBytes = kOffExtra + kExtra * kThreads * 4;  // u32 [kExtra][kThreads]
};

// keeps the compiler from folding a base pointer back into per-access 64-bit index arithmetic
template <typename P>
__device__ __forceinline__ P* opaque_ptr(P* p) {
  asm volatile("" : "+l"(p));
  return p;
}
__device__ __forceinline__ Vec4<float> ldg_vec4(const Vec4<float>* p) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return Vec4<float>{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Vec4<double> ldg_vec4(const Vec4<double>* p) {
  const double2 u = __ldg(reinterpret_cast<const double2*>(p));
  const double2 v = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return Vec4<double>{u.x, u.y, v.x, v.y};
}
__device__ __forceinline__ unsigned lds_u32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// n / d for 0 <= n < 2^31, d >= 1, with rcp = 1.0f / d: float estimate + exact fix-up
__device__ __forceinline__ int div_fix(int n, int d, float rcp, int& rem) {
  int q = __float2int_rz(__int2float_rz(n) * rcp);
  int r = n - q * d;
  while (r < 0) { --q; r += d; }
  while (r >= d) { ++q; r -= d; }
  rem = r;
  return q;
}

// candidate test of phase B: a hit is anything the exact law could turn into a non-zero force
// (|rij|^2 < (Ri + Rj)^2 with a margin for rounding) or whose minimum image is not the plain
// difference (pairs across the periodic seam: rare, decided exactly in phase C)
template <typename F, int D, bool PERIODIC>
__device__ __forceinline__ unsigned rows_hit(const Body<F>& a, const Vec4<F>& q, F hbmin2) {
  const F rx = a.x - q.x, ry = a.y - q.y, rz = D == 3 ? a.z - q.z : F(0);
  const F d2 = rx * rx + ry * ry + rz * rz;
  const F rs = a.r + q.w;
  // branch-free: both comparisons are evaluated and combined as integers
  unsigned hit = (unsigned)(d2 < rs * rs * F(1.00001));
  if (PERIODIC) hit |= (unsigned)!(d2 < hbmin2);
  return hit;
}

// packed f32x2 arithmetic (sm_100: FADD2 / FMUL2 work on 64-bit register pairs)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// The kU candidate tests of one stencil row -> kU hit bits.  f32: the record (x, y | z, rad) is two register
// pairs; (ax, ay) - (x, y) and (az, -ar) - (z, rad) = (dz, -(ar + rad)) are two FADD2, their squares two FMUL2:
// 4 packed + 2 adds + 1 multiply + 2 compares per candidate.
template <typename F, int D, bool PERIODIC, int kU>
struct RowTest {
  Body<F> a;
  F hbmin2;
  __device__ __forceinline__ RowTest(const Body<F>& a_, F hb) : a(a_), hbmin2(hb) {}
  __device__ __forceinline__ unsigned run(const Vec4<F>* p) const {
    Vec4<F> q[kU];
#pragma unroll
    for (int j = 0; j < kU; ++j) q[j] = ldg_vec4(p + j);  // past the run: harmless, masked by the caller
    unsigned bits = 0;
#pragma unroll
    for (int j = 0; j < kU; ++j) bits |= rows_hit<F, D, PERIODIC>(a, q[j], hbmin2) << j;
    return bits;
  }
};
template <int D, bool PERIODIC, int kU>
struct RowTest<float, D, PERIODIC, kU> {
  unsigned long long axy, azr;
  float hbmin2;
  __device__ __forceinline__ RowTest(const Body<float>& a, float hb)
      : axy(pack2(a.x, a.y)), azr(pack2(a.z, -a.r)), hbmin2(hb) {}
  __device__ __forceinline__ unsigned run(const Vec4<float>* p) const {
    ulonglong2 q[kU];
#pragma unroll
    for (int j = 0; j < kU; ++j) q[j] = __ldg(reinterpret_cast<const ulonglong2*>(p + j));
    unsigned bits = 0;
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const unsigned long long u = sub2(axy, q[j].x), v = sub2(azr, q[j].y);
      float dx2, dy2, dz2, rs2;
      unpack2(mul2(u, u), dx2, dy2);
      unpack2(mul2(v, v), dz2, rs2);
      const float d2 = dx2 + dy2 + dz2;
      unsigned hit = (unsigned)(d2 < rs2 * 1.00001f);
      if (PERIODIC) hit |= (unsigned)!(d2 < hbmin2);
      bits |= hit << j;
    }
    return bits;
  }
};

template <typename F, int LAW, int D, bool PERIODIC, bool SIMPLE>
__device__ __forceinline__ void pair_rows_body(const Ctx<F>& c, int b, int k,
                                               const GridInfo<typename RT<F>::I>& g, unsigned sbase, F* f,
                                               F* t, F* en) {
  using T = RT<F>;
  using Cfg = RowsCfg<D>;
  constexpr bool CS = LAW == JDB200_LAW_CUNDALLSTRACK;
  constexpr int kU = Cfg::kU;
  constexpr int kR = Cfg::kRows;
  constexpr unsigned kT = Cfg::kThreads;
  constexpr unsigned kSlotMask = (1u << 27) - 1u;  // run word = start | min(len, 31) << 27 (rows_ok: n < 2^27)
  const unsigned tid = threadIdx.x, lane = tid & 31u;
  const unsigned rs0 = sbase + tid * 4;                    // run words, this thread's column
  const unsigned el0 = sbase + Cfg::kOffExtra + tid * 4;   // extra list
  const size_t off = (size_t)b * c.n;
  const Vec4<F>* sp = opaque_ptr(c.spos + off);
  const int* cst = opaque_ptr(c.cell_start + (size_t)b * c.cell_stride);
  const Body<F> a = load_sorted(c, off, k, CS);
  const int key = c.tmp_key[off + k];  // hash of slot k (the in-cell fix-up keeps cells in place)
  // |rij|^2 below this => every |rij_d| < box_d / 2 => the minimum-image term is exactly zero
  F hbmin2 = F(0);
  if (PERIODIC) {
    F m = c.box[b * D];
#pragma unroll
    for (int d = 1; d < D; ++d) m = T::fmin(m, c.box[b * D + d]);
    m *= F(0.499999);
    hbmin2 = m * m;
  }

  // ---- A: cell coordinates decoded from the hash (flat_walk_ok(): every coordinate lies in
  // [0, g), so hash <-> coordinates is one-to-one); wrapped / out-of-grid rules of cell_list.py:66-80 ----
  const int gx = (int)g.gd[0], gy = (int)g.gd[1], gz = (int)g.gd[2];
  const int sy = gx, sz = gx * gy;  // strides of the x-fastest hash (_partition.py:91-93)
  int cx, cy, cz = 0;
  {
    const int tq = div_fix(key, gx, __frcp_rn(__int2float_rn(gx)), cx);
    if (D == 3) cz = div_fix(tq, gy, __frcp_rn(__int2float_rn(gy)), cy);
    else cy = tq;
  }
  // x segment: the cells of [cx - 1, cx + 1] inside the grid; periodic: the cell outside wraps
  // around and is handled on its own (x2; the grid has >= 3 cells per axis here)
  const int x1 = max(cx - 1, 0);
  const int n1 = min(cx + 1, gx - 1) - x1 + 1;
  int x2 = -1;
  if (PERIODIC) x2 = cx == 0 ? gx - 1 : (cx == gx - 1 ? 0 : -1);
  // hash of cell (0, ny, nz), negative outside a non-periodic grid
  auto axis_base = [&](int n, int gn, int stride) -> int {
    if (PERIODIC) return wrap1(n, gn) * stride;
    return (n >= 0 && n < gn) ? n * stride : (int)0x80000000;  // stays negative after adding the other axis
  };
  // all kRows run ranges at once (2 kRows independent loads), parked as run words in this thread's column
  {
    const int ybs[3] = {axis_base(cy - 1, gy, sy), cy * sy, axis_base(cy + 1, gy, sy)};
    int zbs[3] = {0, 0, 0};
    if (D == 3) { zbs[0] = axis_base(cz - 1, gz, sz); zbs[1] = cz * sz; zbs[2] = axis_base(cz + 1, gz, sz); }
    int s[kR], e[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int yb = ybs[r % 3], zb = zbs[r / 3];
      const bool ok = PERIODIC || (yb | zb) >= 0;
      const int h0 = yb + zb + x1;
      s[r] = ok ? cst[h0] : 0;
      e[r] = ok ? cst[h0 + n1] : 0;
    }
#pragma unroll
    for (int r = 0; r < kR; ++r)
      sts_u32(rs0 + (unsigned)r * (kT * 4), (unsigned)s[r] | ((unsigned)min(e[r] - s[r], 31) << 27));
  }

  // ---- B: candidate tests, one stencil row per trip ----
  unsigned mlo = 0u, mhi = 0u;     // bit r * kU + j of (mhi : mlo): slot (run start of row r) + j is a hit
  int ne = 0;                      // entries of the extra list
  bool ovf = false;                // extra list full (or a run of >= 31): the generic walk redoes this particle
  int s_self = 0;
  auto push_extra = [&](int kj) {
    if (ne < Cfg::kExtra) {
      sts_u32(el0 + (unsigned)ne * (kT * 4), (unsigned)kj);
      ++ne;
    } else {
      ovf = true;
    }
  };
  const RowTest<F, D, PERIODIC, kU> rt(a, hbmin2);
#pragma unroll 1
  for (int r = 0; r < kR; ++r) {
    const unsigned w = lds_u32(rs0 + (unsigned)r * (kT * 4));
    const int s0 = (int)(w & kSlotMask), len = (int)(w >> 27);
    const unsigned bits = rt.run(sp + s0) & ((1u << min(len, kU)) - 1u);
    if (r < 32 / kU) mlo |= bits << (r * kU);
    else mhi |= bits << (r * kU - 32);
    if (r == kR / 2) s_self = s0;
    if (len > kU) {  // run tail (more than kU particles in three cells: dense random packings)
      if (len == 31) ovf = true;
      for (int kj = s0 + kU; kj < s0 + len; ++kj)
        if (rows_hit<F, D, PERIODIC>(a, ldg_vec4(sp + kj), hbmin2)) push_extra(kj);
    }
  }
  // the cells that wrap around the periodic box in x (owners in the first / last cell of an x-row, at most a
  // few lanes of a warp): the warp serves them one owner at a time, lane r taking stencil row r
  if (PERIODIC) {
    unsigned todo = __ballot_sync(0xffffffffu, x2 >= 0);
    while (todo) {
      const int L = __ffs((int)todo) - 1;
      todo &= todo - 1u;
      const int cyL = __shfl_sync(0xffffffffu, cy, L), czL = __shfl_sync(0xffffffffu, cz, L);
      const int x2L = __shfl_sync(0xffffffffu, x2, L);
      Body<F> o;
      o.x = __shfl_sync(0xffffffffu, a.x, L); o.y = __shfl_sync(0xffffffffu, a.y, L);
      o.z = __shfl_sync(0xffffffffu, a.z, L); o.r = __shfl_sync(0xffffffffu, a.r, L);
      int neL = __shfl_sync(0xffffffffu, ne, L);
      bool ovfL = false;
      int s2 = 0, e2 = 0;
      if ((int)lane < kR) {
        const int h2 = wrap1(cyL + (int)lane % 3 - 1, gy) * sy + (D == 3 ? wrap1(czL + (int)lane / 3 - 1, gz) * sz : 0) + x2L;
        s2 = cst[h2];
        e2 = cst[h2 + 1];
      }
      const unsigned elL = sbase + Cfg::kOffExtra + (tid - lane + (unsigned)L) * 4;  // owner's extra list
      for (int it = 0; __any_sync(0xffffffffu, s2 + it < e2); ++it) {
        const int kj = s2 + it;
        bool hit = false;
        if (kj < e2) {
          const Vec4<F> qq = ldg_vec4(sp + kj);
          // Domain._displacement (periodic.py:75-79)
          F rr[3] = {o.x - qq.x, o.y - qq.y, D == 3 ? o.z - qq.z : F(0)};
          F d2 = F(0);
#pragma unroll
          for (int d = 0; d < D; ++d) {
            rr[d] = T::sub(rr[d], T::mul(c.box[b * D + d], T::rint(T::mul(rr[d], c.inv_box[b * D + d]))));
            d2 += rr[d] * rr[d];
          }
          const F rs = o.r + qq.w;
          hit = d2 < rs * rs * F(1.00001);
        }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        const int pos = neL + __popc(hm & ((1u << lane) - 1u));
        if (hit) {
          if (pos < Cfg::kExtra) sts_u32(elL + (unsigned)pos * (kT * 4), (unsigned)kj);
          else ovfL = true;
        }
        neL += __popc(hm);
      }
      ovfL = __any_sync(0xffffffffu, ovfL);
      __syncwarp();
      if ((int)lane == L) {
        ne = min(neL, Cfg::kExtra);
        ovf |= ovfL;
      }
    }
  }
  const int idx = SIMPLE ? 0 : c.perm[off + k];
  const int clump = SIMPLE ? 0 : (c.sclump[off + k] & 0x7fffffff);
  const bool interact = c.interact && c.interact[b];
  const LawCtx<F> lc = make_law_ctx(c, b);
  if (SIMPLE) {  // clump_id == arange(N): only the particle itself is excluded; drop its bit
    const int js = k - s_self;
    if (js < kU) mlo &= ~(1u << (Cfg::kRows / 2 * kU + js));
  }

  // ---- C: the force law on the hits ----
  f[0] = f[1] = f[2] = F(0);
  t[0] = t[1] = t[2] = F(0);
  F e_acc = F(0);
  auto contact = [&](int kcur) {
    const Vec4<F> qc = ldg_vec4(sp + kcur);
    if (SIMPLE) {
      if (kcur == k) return;
    } else {
      const int sc = c.sclump[off + kcur];
      if (!pair_valid(c, off, idx, clump, sc, kcur, interact)) return;
    }
    F rij[3] = {T::sub(a.x, qc.x), T::sub(a.y, qc.y), D == 3 ? T::sub(a.z, qc.z) : F(0)};
    if (PERIODIC) {
      const F d2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
      if (!(d2 < hbmin2)) {
#pragma unroll
        for (int d = 0; d < D; ++d)
          rij[d] = T::sub(rij[d], T::mul(lc.box[d], T::rint(T::mul(rij[d], lc.inv_box[d]))));
      }
    }
    Body<F> bj;
    bj.x = qc.x; bj.y = qc.y; bj.z = qc.z; bj.r = qc.w;
    bj.mat = (c.nmat > 1) ? c.smat[off + kcur] : 0;
    if (CS) {
      const Vec4<F> v = c.svel[off + kcur];
      const Vec4<F> w = c.sang[off + kcur];
      bj.vx = v.x; bj.vy = v.y; bj.vz = v.z; bj.m = v.w;
      bj.wx = w.x; bj.wy = w.y; bj.wz = w.z;
    }
    F ff[3], tt[3];
    pair_force_rij<F, LAW>(lc, a, bj, rij, ff, tt);
    f[0] += ff[0]; f[1] += ff[1]; f[2] += ff[2];
    if (CS) { t[0] += tt[0]; t[1] += tt[1]; t[2] += tt[2]; }
    // minimiser loop: the pair energy from the same displacement (0.5 per side, _partition.py:39-51)
    if (c.want_energy) e_acc += F(0.5) * pair_energy_rij<F, LAW>(lc, a, bj, rij);
  };
  if (!ovf) {
    int e = 0;
    while (true) {  // one instance of the law: mask bits in row order, then the extra list
      int kcur;
      if (mlo | mhi) {
        const unsigned m = mlo ? mlo : mhi;
        const unsigned bit = 31u - (unsigned)__clz((int)(m & (0u - m)));
        const unsigned row = (bit / kU) + (mlo ? 0u : 32u / kU);
        if (mlo) mlo &= mlo - 1u;
        else mhi &= mhi - 1u;
        kcur = (int)((lds_u32(rs0 + row * (kT * 4)) & kSlotMask) + bit % kU);
      } else if (e < ne) {
        kcur = (int)lds_u32(el0 + (unsigned)e * (kT * 4));
        ++e;
      } else {
        break;
      }
      contact(kcur);
    }
  } else {
    // generic x-run walk (law inside the candidate loop): any occupancy
    ForceVis<F, LAW, D, PERIODIC, SIMPLE> vis{c, lc, off, c.spos + off};
    vis.a = a;
    vis.k = k;
    vis.idx = c.perm[off + k];
    vis.clump = clump;
    vis.interact = interact;
#pragma unroll
    for (int d = 0; d < 3; ++d) vis.hb[d] = lc.box[d] * F(0.499999);
    vis.f[0] = vis.f[1] = vis.f[2] = F(0);
    vis.t[0] = vis.t[1] = vis.t[2] = F(0);
    const F pp[3] = {a.x, a.y, a.z};
    walk_runs<F, D, PERIODIC>(c, b, g, pp, vis);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      f[d] = vis.f[d];
      t[d] = vis.t[d];
    }
    e_acc = vis.e;
  }
  *en = e_acc;
}

// The systems the row kernel cannot serve (sorted fallback, periodic de-dup, particles outside the grid)
// take the generic walk INSIDE the same launch (no idle fallback kernel behind every step); kept out of
// line so that its registers do not weigh on the main path.
template <typename F, int LAW, int D, bool PERIODIC>
__device__ __noinline__ void pair_generic(const Ctx<F>& c, int b, int k, int with_torque) {
  const GridInfo<typename RT<F>::I> g = c.gi[b];
  const bool fast = fast_walk_ok(g) && !c.prune, simple = !c.clumps && !g.any_bond;
  if (c.fused) {
    if (simple) pair_force_body<F, LAW, D, PERIODIC, true, 1>(c, b, k, g, fast, with_torque);
    else pair_force_body<F, LAW, D, PERIODIC, false, 1>(c, b, k, g, fast, with_torque);
  } else {
    if (simple) pair_force_body<F, LAW, D, PERIODIC, true, 0>(c, b, k, g, fast, with_torque);
    else pair_force_body<F, LAW, D, PERIODIC, false, 0>(c, b, k, g, fast, with_torque);
  }
}

template <typename F, int LAW, int D, bool PERIODIC>
__global__ void __launch_bounds__(RowsCfg<D>::kThreads, sizeof(F) == 4 ? (LAW == JDB200_LAW_CUNDALLSTRACK ? 8 : JDB_ROWS_MINB) : 4)
    k_pair_rows(const __grid_constant__ Ctx<F> c, int with_torque) {  // grid constant: pair_generic takes its address
  pdl_prologue();
  using I = typename RT<F>::I;
  __shared__ __align__(16) unsigned char smem[RowsCfg<D>::kBytes];
  const int b = blockIdx.y;
  const GridInfo<I> g = c.gi[b];
  const bool mine = rows_ok(c, g);
  if (blockIdx.x == 0 && threadIdx.x == 0 && c.overflow) {
    if (mine || c.grid_mode != JDB200_GRID_DENSE) c.overflow[b] = (uint8_t)g.hash_overflow;
    else c.overflow[b] = 1;  // JDB200_GRID_DENSE and the dense table cannot hold this system
  }
  const long long n_live = c.n_dev ? *c.n_dev : c.n;  // ragged rows (batch == 1: offsets do not depend on n)
  if ((long long)blockIdx.x * blockDim.x >= n_live) return;  // a block of the launch bound past the live rows
  const long long k0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k0 < n_live;
  if (!mine) {
    if (live && c.grid_mode != JDB200_GRID_DENSE) pair_generic<F, LAW, D, PERIODIC>(c, b, (int)k0, with_torque);
    return;
  }
  // the body uses warp-wide votes: lanes past the end of the array redo the last particle and store nothing
  const long long k = live ? k0 : n_live - 1;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
  F f[3], t[3], en;
  if (!c.clumps && !g.any_bond) pair_rows_body<F, LAW, D, PERIODIC, true>(c, b, (int)k, g, sbase, f, t, &en);
  else pair_rows_body<F, LAW, D, PERIODIC, false>(c, b, (int)k, g, sbase, f, t, &en);
  if (!live) return;
  const size_t off = (size_t)b * c.n;
  c.sforce[off + k] = Vec4<F>{f[0], f[1], f[2], en};  // .w: the particle's energy share (c.want_energy), else 0
  if (LAW == JDB200_LAW_CUNDALLSTRACK) c.storque[off + k] = Vec4<F>{t[0], t[1], t[2], F(0)};
}

// D of the row kernel: one thread per ORIGINAL particle i.  EPI 0: the epilogue of
// DynamicCellList.compute_force (cell_list.py:461-462); EPI 1: the fused sphere driver
// (fused_sphere_epilogue: collider epilogue + ForceManager.apply + step_after_force).
template <typename F, int D, int EPI, bool WITH_E>
__global__ void __launch_bounds__(256) k_after(Ctx<F> c, int with_torque) {
  pdl_prologue();
  JDB_LIVE_ROWS(c);
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const GridInfo<I> g = c.gi[b];
  const bool served = rows_ok(c, g);
  // not served by the row kernel: the generic kernel stores by itself — unless JDB200_GRID_DENSE kept it from
  // being launched; then Collider.overflow is up and the hook still completes, with zero contact forces
  if (!served && c.grid_mode != JDB200_GRID_DENSE) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < c.n;
  if (!live && !WITH_E) return;
  const size_t off = (size_t)b * c.n;
  Vec4<F> fs = Vec4<F>{0, 0, 0, 0}, ts = Vec4<F>{0, 0, 0, 0};
  if (live) {
    if (served) {
      const int slot = c.inv[off + i];
      fs = ldg_vec4(c.sforce + off + slot);
      if (c.law == JDB200_LAW_CUNDALLSTRACK) ts = ldg_vec4(c.storque + off + slot);
    }
    const F f[3] = {fs.x, fs.y, fs.z}, t[3] = {ts.x, ts.y, ts.z};
    if (EPI == 1) fused_sphere_epilogue<F, D>(c, b, g, c.uvm[off + i], (int)i, f, t, with_torque != 0);
    else store_force_torque<F, D>(c, off + i, f, t, g.any_ppr != 0, with_torque != 0);
  }
  // minimiser loop: the gathered record also carries the particle's energy share — block sums in the partial
  // layout of k_pair_energy (same block size), which then skips the systems served here
  if (WITH_E && served) {  // (its own instantiation: the shared array and the barriers stay off the step path)
    __shared__ F sm_e[256];
    sm_e[threadIdx.x] = live ? fs.w : F(0);
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if ((int)threadIdx.x < st) sm_e[threadIdx.x] += sm_e[threadIdx.x + st];
      __syncthreads();
    }
    if (threadIdx.x == 0) c.partial[(size_t)b * c.reduce_blocks + blockIdx.x] = sm_e[0];
  }
}

// FAST = true: systems whose partition allows the x-run walk; FAST = false: the rest
// (sorted fallback, custom stencils, periodic de-dup).  Both kernels are launched; each
// exits at once for the systems the other one owns.
template <typename F, int LAW, int D, bool PERIODIC, bool FAST, int EPI>
__global__ void __launch_bounds__(128) k_pair_force(Ctx<F> c, int with_torque) {
  pdl_prologue();
  JDB_LIVE_ROWS(c);
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const GridInfo<I> g = c.gi[b];
  // FAST = false also serves the systems the flat kernel (below) cannot: stencil range > 1
  // FAST: wider canonical stencils (range != 1; launched instead of the flat kernel when M != 3^D)
  const bool mine = c.prune ? !FAST
                            : (FAST ? (fast_walk_ok(g) && g.range != 1) : !(flat_walk_ok(g) || (fast_walk_ok(g) && g.range != 1)));
  // Collider.overflow (cell_list.py:463).  JDB200_GRID_DENSE launches only the FAST kernel: a
  // system it cannot serve (table too small, custom stencil, periodic de-dup) raises the flag.
  if (blockIdx.x == 0 && threadIdx.x == 0 && c.overflow) {
    if (mine) c.overflow[b] = (uint8_t)g.hash_overflow;
    else if (FAST && c.grid_mode == JDB200_GRID_DENSE && !fast_walk_ok(g)) c.overflow[b] = 1;
  }
  if (!mine) return;
  const bool simple = !c.clumps && !g.any_bond;
  // the FAST kernel has one thread per particle; the fallback kernel is launched with a
  // small grid (an idle launch must cost nothing) and strides over the particles
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < c.n;
       k += (long long)gridDim.x * blockDim.x) {
    if (simple) pair_force_body<F, LAW, D, PERIODIC, true, EPI>(c, b, (int)k, g, FAST && !c.prune, with_torque);
    else pair_force_body<F, LAW, D, PERIODIC, false, EPI>(c, b, (int)k, g, FAST && !c.prune, with_torque);
  }
}

// ---------------------------------------------------------------------------
// K5  pair energy: per-particle 0.5*sum E, block partials, fixed-order final sum
// ---------------------------------------------------------------------------
template <typename F, int LAW>
struct EnergyVis {
  const Ctx<F>& c;
  const LawCtx<F>& lc;
  size_t off;
  Body<F> a;
  int k, idx, clump;
  bool interact, simple;
  F e;
  __device__ __forceinline__ void cell(int, int s, int en) {
    for (int kj = s; kj < en; ++kj) {
      if (simple) {
        if (kj == k) continue;
      } else {
        const int sc = c.sclump[off + kj];
        if (!pair_valid(c, off, idx, clump, sc, kj, interact)) continue;
      }
      const Body<F> bj = load_sorted(c, off, kj, false);
      e += F(0.5) * pair_energy<F, LAW>(lc, a, bj);
    }
  }
  __device__ __forceinline__ void run(int s1, int e1, int s2, int e2) {
    cell(0, s1, e1);
    cell(0, s2, e2);
  }
};

template <typename F>
__device__ __forceinline__ F block_sum_256(F v) {  // fixed tree order => deterministic
  __shared__ F sm[kReduceBlock];
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int s = kReduceBlock / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  return sm[0];
}

// One instantiation per (law, dim, periodic) — like the force kernels — so that each kernel carries ONE x-run walk
// next to the general one (the single four-way kernel was very large, see DESIGN.md "open observation").
template <typename F, int LAW, int D, bool PERIODIC>
__global__ void __launch_bounds__(kReduceBlock) k_pair_energy(Ctx<F> c) {
  pdl_prologue();
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  if (c.want_energy && rows_ok(c, c.gi[b])) return;  // block sums of the row kernel's per-particle shares: k_after
  const size_t off = (size_t)b * c.n;
  // one chunk of kReduceBlock particles per trip: gridDim.x == reduce_blocks in the stand-alone hook; the minimiser
  // loop launches ONE block per system (this kernel is then only the fallback for systems the row kernel skipped)
  for (int blk = blockIdx.x; blk < c.reduce_blocks; blk += gridDim.x) {
    const int k = blk * blockDim.x + threadIdx.x;
    F e = F(0);
    if (k < c.n) {
      const GridInfo<I> g = c.gi[b];
      const LawCtx<F> lc = make_law_ctx(c, b);
      EnergyVis<F, LAW> vis{c, lc, off};
      vis.a = load_sorted(c, off, k, false);
      vis.k = k;
      vis.idx = c.perm[off + k];
      vis.simple = !c.clumps && !g.any_bond;
      vis.clump = vis.simple ? 0 : (c.sclump[off + k] & 0x7fffffff);
      vis.interact = c.interact && c.interact[b];
      vis.e = F(0);
      if (fast_walk_ok(g) && !c.prune) {
        const F pp[3] = {vis.a.x, vis.a.y, vis.a.z};
        walk_runs<F, D, PERIODIC>(c, b, g, pp, vis);
      } else {
        walk_stencil<F>(c, b, k, nullptr, vis);
      }
      e = vis.e;
    }
    const F tot = block_sum_256(e);
    if (threadIdx.x == 0) c.partial[(size_t)b * c.reduce_blocks + blk] = tot;
    __syncthreads();  // block_sum_256's shared array is reused by the next trip
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && c.overflow) c.overflow[b] = (uint8_t)c.gi[b].hash_overflow;
}

template <typename F>
__global__ void __launch_bounds__(kReduceBlock) k_final_sum(const F* __restrict__ partial, int nblocks,
                                                             F* __restrict__ out) {
  pdl_prologue();
  const int b = blockIdx.x;
  F acc = F(0);
  for (int i = threadIdx.x; i < nblocks; i += kReduceBlock) acc += partial[(size_t)b * nblocks + i];
  const F tot = block_sum_256(acc);
  if (threadIdx.x == 0) out[b] = tot;
}

// ---------------------------------------------------------------------------
// K6  neighbour list (cell_list.py:498-595).  Row order = stencil order x sorted-run
// order; the PAIR_UNROLL=4 chunked counting of _make_stencil_body (:99-174) and the
// prefix-sum packing of _pack_stencil_lists (_partition.py:102-150) are reproduced so
// rows, padding and the overflow flag are bit-identical, also when buffers overflow.
// ---------------------------------------------------------------------------
template <typename F>
struct NlVis {
  using I = typename RT<F>::I;
  const Ctx<F>& c;
  const LawCtx<F>& lc;
  size_t off;
  Body<F> a;
  int k, idx, clump;
  bool interact, simple;
  F cutoff_sq;
  I* row;
  long long row_off;  // running sum of raw per-cell counts
  bool stencil_overflow;
  __device__ __forceinline__ void cell(int, int s, int e) {
    const int cap = c.K;
    int cnt = 0;
    for (int k0 = s; k0 < e && cnt < cap + 1; k0 += 4) {  // cond_fun: in_cell * has_space
      for (int u = 0; u < 4; ++u) {
        const int kj = k0 + u;
        if (kj >= e) break;
        if (simple) {
          if (kj == k) continue;
        } else {
          const int sc = c.sclump[off + kj];
          if (!pair_valid(c, off, idx, clump, sc, kj, interact)) continue;
        }
        const Body<F> bj = load_sorted(c, off, kj, false);
        F r[3];
        displacement_div(lc, a, bj, r);
        const F d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        if (!(d2 <= cutoff_sq)) continue;
        const long long dest = row_off + cnt;
        if (cnt < cap && dest < cap) row[dest] = (I)c.perm[off + kj];
        ++cnt;
      }
      stencil_overflow |= cnt > cap;
    }
    row_off += cnt;
  }
};

template <typename F>
__global__ void __launch_bounds__(128) k_neighbor_list(Ctx<F> c, const F* __restrict__ cell_size_nl,
                                                        const F* __restrict__ cutoff,
                                                        typename RT<F>::I* __restrict__ nl) {
  pdl_prologue();
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n || (c.gate && !c.gate[b])) return;
  const size_t off = (size_t)b * c.n;
  const LawCtx<F> lc = make_law_ctx(c, b);
  NlVis<F> vis{c, lc, off};
  vis.a = load_sorted(c, off, k, false);
  vis.k = k;
  vis.idx = c.perm[off + k];
  vis.simple = !c.clumps && !c.gi[b].any_bond;
  vis.clump = vis.simple ? 0 : (c.sclump[off + k] & 0x7fffffff);
  vis.interact = c.interact && c.interact[b];
  vis.cutoff_sq = cutoff[b] * cutoff[b];
  vis.row = nl + (off + vis.idx) * c.K;
  for (int q = 0; q < c.K; ++q) vis.row[q] = I(-1);
  vis.row_off = 0;
  vis.stencil_overflow = false;
  walk_stencil<F>(c, b, k, cell_size_nl, vis);
  if (vis.stencil_overflow || vis.row_off > c.K) c.gi[b].nl_overflow = 1;
}

template <typename F>
__global__ void k_nl_cell_size(Ctx<F> c, const F* __restrict__ cutoff, F* __restrict__ out) {
  pdl_prologue();
  // cell_list.py:537-538: cell_size = max(cell_size, cutoff / max(max(neighbor_mask), 1))
  using I = typename RT<F>::I;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= c.batch) return;
  I mx = 1;
  const I* mask = c.mask + (size_t)b * c.M * c.dim;
  for (int i = 0; i < c.M * c.dim; ++i) mx = mask[i] > mx ? mask[i] : mx;
  out[b] = RT<F>::fmax(c.cell_size[b], RT<F>::div(cutoff[b], RT<F>::from_int(mx)));
}

template <typename F>
__global__ void k_nl_flag(Ctx<F> c, uint8_t* __restrict__ overflow) {
  pdl_prologue();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= c.batch || (c.gate && !c.gate[b])) return;
  overflow[b] = (uint8_t)(c.gi[b].nl_overflow || c.gi[b].hash_overflow);
}

// create_cross_neighbor_list (cell_list.py:600-715): queries pos_a (N_A, D) against the partition
// of the database points (the State's positions); no clump / bond mask; same chunked counting and
// prefix-sum packing as K6, entries are original database indices.
template <typename F>
struct CrossVis {
  using I = typename RT<F>::I;
  const Ctx<F>& c;
  const LawCtx<F>& lc;
  size_t off;
  Body<F> a;
  F cutoff_sq;
  I* row;
  long long row_off;
  bool stencil_overflow;
  __device__ __forceinline__ void cell(int, int s, int e) {
    const int cap = c.K;
    int cnt = 0;
    for (int k0 = s; k0 < e && cnt < cap + 1; k0 += 4) {  // cond_fun: in_cell * has_space
      for (int u = 0; u < 4; ++u) {
        const int kj = k0 + u;
        if (kj >= e) break;
        const Body<F> bj = load_sorted(c, off, kj, false);
        F r[3];
        displacement_div(lc, a, bj, r);
        const F d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        if (!(d2 <= cutoff_sq)) continue;
        const long long dest = row_off + cnt;
        if (cnt < cap && dest < cap) row[dest] = (I)c.perm[off + kj];
        ++cnt;
      }
      stencil_overflow |= cnt > cap;
    }
    row_off += cnt;
  }
};

template <typename F>
__global__ void __launch_bounds__(128) k_cross_neighbor_list(Ctx<F> c, const F* __restrict__ pos_a, long long n_a,
                                                              const F* __restrict__ cell_size_nl,
                                                              const F* __restrict__ cutoff,
                                                              typename RT<F>::I* __restrict__ nl) {
  pdl_prologue();
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_a) return;
  const LawCtx<F> lc = make_law_ctx(c, b);
  CrossVis<F> vis{c, lc, (size_t)b * c.n};
  const F* pa = pos_a + ((size_t)b * n_a + i) * c.dim;
  F pp[3] = {pa[0], pa[1], c.dim == 3 ? pa[2] : F(0)};
  vis.a.x = pp[0]; vis.a.y = pp[1]; vis.a.z = pp[2];
  vis.cutoff_sq = cutoff[b] * cutoff[b];
  vis.row = nl + ((size_t)b * n_a + i) * c.K;
  for (int q = 0; q < c.K; ++q) vis.row[q] = I(-1);
  vis.row_off = 0;
  vis.stencil_overflow = false;
  walk_stencil_at<F>(c, b, pp, cell_size_nl, vis);
  if (vis.stencil_overflow || vis.row_off > c.K) c.gi[b].nl_overflow = 1;
}

// ---------------------------------------------------------------------------
// naive O(N^2) collider (naive.py:187-235, 73-113): thread per particle, all j in
// index order straight from the State arrays.
// ---------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ bool naive_valid(const Ctx<F>& c, size_t off, int i, int j, bool interact) {
  // naive.py:218-224: mask(clump[i], clump[:], bond_id[i], iota): OWNER's bond row vs candidate index
  if (c.clump_id[off + i] == c.clump_id[off + j]) return false;
  if (!interact) {
    const typename RT<F>::I* row = c.bond_id + (off + i) * c.W;
    for (int w = 0; w < c.W; ++w)
      if (row[w] == j) return false;
  }
  return true;
}

template <typename F, int LAW>
__global__ void __launch_bounds__(128) k_naive_force(Ctx<F> c) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const size_t off = (size_t)b * c.n;
  const LawCtx<F> lc = make_law_ctx(c, b);
  const bool interact = c.interact && c.interact[b];
  const bool cs = LAW == JDB200_LAW_CUNDALLSTRACK;
  const Body<F> a = load_orig(c, off + i, cs);
  F f[3] = {0, 0, 0}, t[3] = {0, 0, 0};
  for (int j = 0; j < c.n; ++j) {
    if (!naive_valid(c, off, i, j, interact)) continue;
    const Body<F> bj = load_orig(c, off + j, cs);
    F ff[3], tt[3];
    pair_force<F, LAW>(lc, a, bj, ff, tt);
    f[0] += ff[0]; f[1] += ff[1]; f[2] += ff[2];
    t[0] += tt[0]; t[1] += tt[1]; t[2] += tt[2];
  }
  if (c.dim == 3) store_force_torque<F, 3>(c, off + i, f, t, true, true);
  else store_force_torque<F, 2>(c, off + i, f, t, true, true);
}

template <typename F, int LAW>
__global__ void __launch_bounds__(kReduceBlock) k_naive_energy(Ctx<F> c) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t off = (size_t)b * c.n;
  F e = F(0);
  if (i < c.n) {
    const LawCtx<F> lc = make_law_ctx(c, b);
    const bool interact = c.interact && c.interact[b];
    const Body<F> a = load_orig(c, off + i, false);
    for (int j = 0; j < c.n; ++j) {
      if (!naive_valid(c, off, i, j, interact)) continue;
      e += F(0.5) * pair_energy<F, LAW>(lc, a, load_orig(c, off + j, false));
    }
  }
  const F tot = block_sum_256(e);
  if (threadIdx.x == 0) c.partial[(size_t)b * c.reduce_blocks + blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------
// MultiCellList: expandable AABB of every occupied cell (_loose_cell_aabbs, multi_cell_list.py:46-73): segmented
// min / max of the members' boxes [pos - rad, pos + rad] (with_rad) or of the bare positions (neighbour-list
// builds, :466-468) over the cell's sorted run; centre and half extent are stored at the run's first slot.
// ---------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256) k_cell_aabb(Ctx<F> c, int with_rad) {
  pdl_prologue();
  using T = RT<F>;
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  if (c.gate && !c.gate[b]) return;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n) return;
  const size_t off = (size_t)b * c.n;
  const I key = c.skey[off + k];
  if (k > 0 && c.skey[off + k - 1] == key) return;  // not the first slot of a run
  F lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  for (long long j = k; j < c.n && c.skey[off + j] == key; ++j) {
    const Vec4<F> p = c.spos[off + j];
    const F r = with_rad ? p.w : F(0);
    const F x[3] = {p.x, p.y, p.z};
    for (int d = 0; d < 3; ++d) {
      const F a = T::sub(x[d], r), z = T::add(x[d], r);
      lo[d] = j == k ? a : T::fmin(lo[d], a);
      hi[d] = j == k ? z : T::fmax(hi[d], z);
    }
  }
  c.aabb_c[off + k] = Vec4<F>{T::mul(F(0.5), T::add(lo[0], hi[0])), T::mul(F(0.5), T::add(lo[1], hi[1])),
                              T::mul(F(0.5), T::add(lo[2], hi[2])), F(0)};
  c.aabb_h[off + k] = Vec4<F>{T::mul(F(0.5), T::sub(hi[0], lo[0])), T::mul(F(0.5), T::sub(hi[1], lo[1])),
                              T::mul(F(0.5), T::sub(hi[2], lo[2])), F(0)};
}
template <typename F>
static int cell_aabbs(cudaStream_t s, Ctx<F>& c, int with_rad) {
  JDB_LAUNCH(k_cell_aabb<F>, dim3(cdiv(c.n, 256), c.batch), 256, s, c, with_rad);
  return 0;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <typename F>
int build_partition(cudaStream_t s, Ctx<F>& c, const F* cell_size_override, int hash_mode, bool ext);

#define JDB_LAW_SWITCH(law, CALL)                                     \
  switch (law) {                                                      \
    case JDB200_LAW_SPRING: { constexpr int L = JDB200_LAW_SPRING; CALL; } break;           \
    case JDB200_LAW_HERTZ: { constexpr int L = JDB200_LAW_HERTZ; CALL; } break;             \
    default: { constexpr int L = JDB200_LAW_CUNDALLSTRACK; CALL; } break;                   \
  }

template <typename F, int D, int EPI>
int launch_pair_force_epi(cudaStream_t s, Ctx<F>& c, bool with_torque) {
  const dim3 grid(cdiv(c.n, 128), c.batch);
  const int wt = with_torque ? 1 : 0;
  if (c.prune) {  // MultiCellList: one thread per particle, generic stencil walk with the per-cell AABB test
    if (c.periodic) {
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, true, false, EPI>), grid, 128, s, c, wt));
    } else {
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, false, false, EPI>), grid, 128, s, c, wt));
    }
    return 0;
  }
  if (c.max_cells > 0) {  // a dense table exists
    if (c.M == (D == 3 ? 27 : 9)) {  // default stencil: the row kernel owns the systems it can serve
      constexpr int kT = RowsCfg<D>::kThreads;
      const dim3 gf(cdiv(c.n, kT), c.batch);
      if (c.periodic) {
        JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_rows<F, L, D, true>), gf, kT, s, c, wt));
      } else {
        JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_rows<F, L, D, false>), gf, kT, s, c, wt));
      }
      // (a four-particles-per-thread variant with 128-bit stores was measured slower in every configuration:
      // 24.7 vs 22.6 us fused, 81 vs 69 us on the 4 M-sphere clump workload)
      if (c.want_energy) JDB_LAUNCH((k_after<F, D, EPI, true>), dim3(cdiv(c.n, 256), c.batch), 256, s, c, wt);
      else JDB_LAUNCH((k_after<F, D, EPI, false>), dim3(cdiv(c.n, 256), c.batch), 256, s, c, wt);
      return 0;  // the systems it cannot serve took the generic walk inside the same launch
    } else if (c.periodic) {         // wider canonical stencils: x-run kernel
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, true, true, EPI>), grid, 128, s, c, wt));
    } else {
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, false, true, EPI>), grid, 128, s, c, wt));
    }
  }
  if (c.grid_mode != JDB200_GRID_DENSE || c.max_cells == 0) {  // everything else
    const dim3 gs(std::min(cdiv(c.n, 128), std::max(1, 1184 / c.batch)), c.batch);  // <= 8 CTAs per SM: an idle launch is cheap
    if (c.periodic) {
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, true, false, EPI>), gs, 128, s, c, wt));
    } else {
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, false, false, EPI>), gs, 128, s, c, wt));
    }
  }
  return 0;
}

template <typename F, int D>
int launch_pair_force(cudaStream_t s, Ctx<F>& c, bool with_torque) {
  return c.fused ? launch_pair_force_epi<F, D, 1>(s, c, with_torque)
                 : launch_pair_force_epi<F, D, 0>(s, c, with_torque);
}

// The four (dtype, dim) instantiations of the force launchers are the bulk of this file's compile
// time: each is built by its own translation unit (pair_f32_3d.cu ... include this file with
// JDB_PAIR_SLICE_F / JDB_PAIR_SLICE_D set), the rest by pair.cu itself.
#ifdef JDB_PAIR_SLICE_F
template int launch_pair_force<JDB_PAIR_SLICE_F, JDB_PAIR_SLICE_D>(cudaStream_t, Ctx<JDB_PAIR_SLICE_F>&, bool);
#else
extern template int launch_pair_force<float, 2>(cudaStream_t, Ctx<float>&, bool);
extern template int launch_pair_force<float, 3>(cudaStream_t, Ctx<float>&, bool);
extern template int launch_pair_force<double, 2>(cudaStream_t, Ctx<double>&, bool);
extern template int launch_pair_force<double, 3>(cudaStream_t, Ctx<double>&, bool);
#endif

// hash_mode / ext: fusion of the linear integrator into the hash kernel (celllist.cu k_hash);
// with_torque = false skips the torque store (fused driver only).
template <typename F>
int celllist_force(cudaStream_t s, Ctx<F>& c, int hash_mode, bool ext, bool with_torque) {
  if (c.n == 0) return 0;
  int rc = build_partition<F>(s, c, nullptr, hash_mode, ext);
  if (rc) return rc;
  if (c.prune && (rc = cell_aabbs<F>(s, c, 1))) return rc;
  return c.dim == 3 ? launch_pair_force<F, 3>(s, c, with_torque) : launch_pair_force<F, 2>(s, c, with_torque);
}

// reuse: the partition in the workspace is the one a force call of the same State just built (minimiser loop)
template <typename F>
int celllist_energy(cudaStream_t s, Ctx<F>& c, F* energy, bool reuse) {
  if (c.n == 0) return cudaMemsetAsync(energy, 0, sizeof(F) * c.batch, s) == cudaSuccess ? 0 : JDB200_ECUDA;
  int rc = reuse ? 0 : build_partition<F>(s, c, nullptr, 0, false);
  if (rc) return rc;
  if (c.prune && !reuse && (rc = cell_aabbs<F>(s, c, 1))) return rc;
  const dim3 grid(c.want_energy ? 1 : c.reduce_blocks, c.batch);
  if (c.dim == 3) {
    if (c.periodic) { JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_energy<F, L, 3, true>), grid, kReduceBlock, s, c)); }
    else { JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_energy<F, L, 3, false>), grid, kReduceBlock, s, c)); }
  } else {
    if (c.periodic) { JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_energy<F, L, 2, true>), grid, kReduceBlock, s, c)); }
    else { JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_energy<F, L, 2, false>), grid, kReduceBlock, s, c)); }
  }
  JDB_LAUNCH(k_final_sum<F>, dim3(c.batch), kReduceBlock, s, c.partial, c.reduce_blocks, energy);
  return 0;
}

template <typename F>
int celllist_neighbor_list(cudaStream_t s, Ctx<F>& c, const F* cutoff, typename RT<F>::I* nl,
                           uint8_t* overflow) {
  if (c.n == 0 || c.K == 0)
    return cudaMemsetAsync(overflow, 0, c.batch, s) == cudaSuccess ? 0 : JDB200_ECUDA;
  F* cs_nl = c.partial;  // [B] scratch for the inflated cell size
  JDB_LAUNCH(k_nl_cell_size<F>, dim3(cdiv(c.batch, 64)), 64, s, c, cutoff, cs_nl);
  int rc = build_partition<F>(s, c, cs_nl, 0, false);
  if (rc) return rc;
  if (c.prune) {  // multi_cell_list.py:466-481: AABBs of the bare positions, query box = point +- cutoff
    if ((rc = cell_aabbs<F>(s, c, 0))) return rc;
    c.prune_cut = cutoff;
  }
  JDB_LAUNCH(k_neighbor_list<F>, dim3(cdiv(c.n, 128), c.batch), 128, s, c, cs_nl, cutoff, nl);
  JDB_LAUNCH(k_nl_flag<F>, dim3(cdiv(c.batch, 64)), 64, s, c, overflow);
  return 0;
}

template <typename F>
int celllist_cross_neighbor_list(cudaStream_t s, Ctx<F>& c, const F* pos_a, long long n_a, const F* cutoff,
                                 typename RT<F>::I* nl, uint8_t* overflow) {
  if (cudaMemsetAsync(overflow, 0, c.batch, s) != cudaSuccess) return JDB200_ECUDA;
  if (n_a == 0 || c.K == 0) return 0;
  if (c.n == 0)  // empty database: every row is padding
    return cudaMemsetAsync(nl, 0xff, sizeof(typename RT<F>::I) * c.batch * n_a * c.K, s) == cudaSuccess ? 0 : JDB200_ECUDA;
  F* cs_nl = c.partial;  // [B] scratch for the inflated cell size
  JDB_LAUNCH(k_nl_cell_size<F>, dim3(cdiv(c.batch, 64)), 64, s, c, cutoff, cs_nl);
  int rc = build_partition<F>(s, c, cs_nl, 0, false);
  if (rc) return rc;
  if (c.prune) {  // multi_cell_list.py:640-643, 676-682
    if ((rc = cell_aabbs<F>(s, c, 0))) return rc;
    c.prune_cut = cutoff;
  }
  JDB_LAUNCH(k_cross_neighbor_list<F>, dim3(cdiv(n_a, 128), c.batch), 128, s, c, pos_a, n_a, cs_nl, cutoff, nl);
  JDB_LAUNCH(k_nl_flag<F>, dim3(cdiv(c.batch, 64)), 64, s, c, overflow);
  return 0;
}

template <typename F>
int naive_force(cudaStream_t s, Ctx<F>& c) {
  if (c.n == 0) return 0;
  const dim3 grid(cdiv(c.n, 128), c.batch);
  JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_naive_force<F, L>), grid, 128, s, c));
  return 0;
}

template <typename F>
int naive_energy(cudaStream_t s, Ctx<F>& c, F* energy) {
  if (c.n == 0) return cudaMemsetAsync(energy, 0, sizeof(F) * c.batch, s) == cudaSuccess ? 0 : JDB200_ECUDA;
  const dim3 grid(c.reduce_blocks, c.batch);
  JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_naive_energy<F, L>), grid, kReduceBlock, s, c));
  JDB_LAUNCH(k_final_sum<F>, dim3(c.batch), kReduceBlock, s, c.partial, c.reduce_blocks, energy);
  return 0;
}

#define JDB_INST(F)                                                                     \
  template int celllist_force<F>(cudaStream_t, Ctx<F>&, int, bool, bool);                          \
  template int celllist_energy<F>(cudaStream_t, Ctx<F>&, F*, bool);                     \
  template int celllist_neighbor_list<F>(cudaStream_t, Ctx<F>&, const F*, RT<F>::I*, uint8_t*); \
  template int celllist_cross_neighbor_list<F>(cudaStream_t, Ctx<F>&, const F*, long long, const F*, RT<F>::I*, uint8_t*); \
  template int naive_force<F>(cudaStream_t, Ctx<F>&);                                   \
  template int naive_energy<F>(cudaStream_t, Ctx<F>&, F*);
#ifndef JDB_PAIR_SLICE_F
JDB_INST(float)
JDB_INST(double)
#endif

}  // namespace jdb
