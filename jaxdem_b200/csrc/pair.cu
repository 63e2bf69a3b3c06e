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
template <typename F, typename Vis>
__device__ __forceinline__ void walk_stencil(const Ctx<F>& c, int b, int k, const F* cell_size_override,
                                             Vis& vis) {
  using I = typename RT<F>::I;
  const GridInfo<I> g = c.gi[b];
  const size_t off = (size_t)b * c.n;
  const bool dense = use_dense(g);
  const Vec4<F> p = c.spos[off + k];
  const F pp[3] = {p.x, p.y, p.z};
  const F cs = cell_size_override ? cell_size_override[b] : c.cell_size[b];
  I cc[3] = {0, 0, 0};
  for (int d = 0; d < c.dim; ++d)
    cc[d] = cell_coord<F, I>(pp[d], c.anchor[b * c.dim + d], c.box[b * c.dim + d], cs, g.gd[d],
                             c.periodic);
  const I* mask = c.mask + (size_t)b * c.M * c.dim;
  const I* skey = c.skey + off;
  const int* cstart = c.cell_start + (size_t)b * c.cell_stride;
  const int n = (int)c.n;
  for (int m = 0; m < c.M; ++m) {
    const I h = neighbor_hash<F, I>(cc, mask + m * c.dim, g.gd, g.stride, c.dim, c.periodic);
    if (c.periodic && g.need_dedup) {
      bool dup = false;
      for (int m2 = 0; m2 < m && !dup; ++m2)
        dup = neighbor_hash<F, I>(cc, mask + m2 * c.dim, g.gd, g.stride, c.dim, c.periodic) == h;
      if (dup) continue;
    }
    int s, e;
    if (dense) {
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
  return g.dense && !g.dense_fail && g.canonical && !g.need_dedup;
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

template <typename F, int LAW, int D, bool PERIODIC, bool SIMPLE>
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
  } else {
    walk_stencil<F>(c, b, k, nullptr, vis);
  }
  store_force_torque<F, D>(c, off + vis.idx, vis.f, vis.t, g.any_ppr != 0, with_torque != 0);
}

// FAST = true: systems whose partition allows the x-run walk; FAST = false: the rest
// (sorted fallback, custom stencils, periodic de-dup).  Both kernels are launched; each
// exits at once for the systems the other one owns.
template <typename F, int LAW, int D, bool PERIODIC, bool FAST>
__global__ void __launch_bounds__(128) k_pair_force(Ctx<F> c, int with_torque) {
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const GridInfo<I> g = c.gi[b];
  const bool mine = fast_walk_ok(g) == FAST;
  // Collider.overflow (cell_list.py:463).  JDB200_GRID_DENSE launches only the FAST kernel: a
  // system it cannot serve (table too small, custom stencil, periodic de-dup) raises the flag.
  if (blockIdx.x == 0 && threadIdx.x == 0 && c.overflow &&
      (mine || (FAST && c.grid_mode == JDB200_GRID_DENSE)))
    c.overflow[b] = (uint8_t)(g.hash_overflow || !mine);
  if (!mine) return;
  const bool simple = !c.clumps && !g.any_bond;
  // the FAST kernel has one thread per particle; the fallback kernel is launched with a
  // small grid (an idle launch must cost nothing) and strides over the particles
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < c.n;
       k += (long long)gridDim.x * blockDim.x) {
    if (simple) pair_force_body<F, LAW, D, PERIODIC, true>(c, b, (int)k, g, FAST, with_torque);
    else pair_force_body<F, LAW, D, PERIODIC, false>(c, b, (int)k, g, FAST, with_torque);
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

template <typename F, int LAW>
__global__ void __launch_bounds__(kReduceBlock) k_pair_energy(Ctx<F> c) {
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t off = (size_t)b * c.n;
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
    if (fast_walk_ok(g)) {
      const F pp[3] = {vis.a.x, vis.a.y, vis.a.z};
      if (c.dim == 3) {
        if (c.periodic) walk_runs<F, 3, true>(c, b, g, pp, vis);
        else walk_runs<F, 3, false>(c, b, g, pp, vis);
      } else {
        if (c.periodic) walk_runs<F, 2, true>(c, b, g, pp, vis);
        else walk_runs<F, 2, false>(c, b, g, pp, vis);
      }
    } else {
      walk_stencil<F>(c, b, k, nullptr, vis);
    }
    e = vis.e;
  }
  const F tot = block_sum_256(e);
  if (threadIdx.x == 0) c.partial[(size_t)b * c.reduce_blocks + blockIdx.x] = tot;
  if (k == 0 && c.overflow) c.overflow[b] = (uint8_t)c.gi[b].hash_overflow;
}

template <typename F>
__global__ void __launch_bounds__(kReduceBlock) k_final_sum(const F* __restrict__ partial, int nblocks,
                                                             F* __restrict__ out) {
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
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n) return;
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
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= c.batch) return;
  overflow[b] = (uint8_t)(c.gi[b].nl_overflow || c.gi[b].hash_overflow);
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

template <typename F, int D>
int launch_pair_force(cudaStream_t s, Ctx<F>& c, bool with_torque) {
  const dim3 grid(cdiv(c.n, 128), c.batch);
  const int wt = with_torque ? 1 : 0;
  if (c.max_cells > 0) {  // a dense table exists: the x-run kernel owns the systems it can serve
    if (c.periodic) {
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, true, true>), grid, 128, s, c, wt));
    } else {
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, false, true>), grid, 128, s, c, wt));
    }
  }
  if (c.grid_mode != JDB200_GRID_DENSE || c.max_cells == 0) {  // everything else
    const dim3 gs(std::min(cdiv(c.n, 128), std::max(1, 4736 / c.batch)), c.batch);  // <= 32 CTAs per SM
    if (c.periodic) {
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, true, false>), gs, 128, s, c, wt));
    } else {
      JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L, D, false, false>), gs, 128, s, c, wt));
    }
  }
  return 0;
}

// hash_mode / ext: fusion of the linear integrator into the hash kernel (celllist.cu k_hash);
// with_torque = false skips the torque store (fused driver only).
template <typename F>
int celllist_force(cudaStream_t s, Ctx<F>& c, int hash_mode, bool ext, bool with_torque) {
  if (c.n == 0) return 0;
  int rc = build_partition<F>(s, c, nullptr, hash_mode, ext);
  if (rc) return rc;
  return c.dim == 3 ? launch_pair_force<F, 3>(s, c, with_torque) : launch_pair_force<F, 2>(s, c, with_torque);
}

template <typename F>
int celllist_energy(cudaStream_t s, Ctx<F>& c, F* energy) {
  if (c.n == 0) return cudaMemsetAsync(energy, 0, sizeof(F) * c.batch, s) == cudaSuccess ? 0 : JDB200_ECUDA;
  int rc = build_partition<F>(s, c, nullptr, 0, false);
  if (rc) return rc;
  const dim3 grid(c.reduce_blocks, c.batch);
  JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_energy<F, L>), grid, kReduceBlock, s, c));
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
  JDB_LAUNCH(k_neighbor_list<F>, dim3(cdiv(c.n, 128), c.batch), 128, s, c, cs_nl, cutoff, nl);
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
  template int celllist_energy<F>(cudaStream_t, Ctx<F>&, F*);                           \
  template int celllist_neighbor_list<F>(cudaStream_t, Ctx<F>&, const F*, RT<F>::I*, uint8_t*); \
  template int naive_force<F>(cudaStream_t, Ctx<F>&);                                   \
  template int naive_energy<F>(cudaStream_t, Ctx<F>&, F*);
JDB_INST(float)
JDB_INST(double)

}  // namespace jdb
