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
  const int* cstart = c.cell_start + (size_t)b * (c.max_cells + 1);
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
// K4  pair force
// ---------------------------------------------------------------------------
template <typename F, int LAW>
struct ForceVis {
  const Ctx<F>& c;
  const LawCtx<F>& lc;
  size_t off;
  Body<F> a;
  int idx, clump;
  bool interact;
  F f[3], t[3];
  __device__ __forceinline__ void cell(int, int s, int e) {
    for (int kj = s; kj < e; ++kj) {
      const int sc = c.sclump[off + kj];
      if (!pair_valid(c, off, idx, clump, sc, kj, interact)) continue;
      const Body<F> bj = load_sorted(c, off, kj, LAW == JDB200_LAW_CUNDALLSTRACK);
      F ff[3], tt[3];
      pair_force<F, LAW>(lc, a, bj, ff, tt);
      f[0] += ff[0]; f[1] += ff[1]; f[2] += ff[2];
      t[0] += tt[0]; t[1] += tt[1]; t[2] += tt[2];
    }
  }
};

template <typename F>
__device__ __forceinline__ void store_force_torque(const Ctx<F>& c, size_t gidx, const F* f, const F* t) {
  // collider epilogue (cell_list.py:461-462): torque = sum T + cross(_pos_p_rot, sum F)
  F* fo = c.force + gidx * c.dim;
  F* to = c.torque + gidx * c.A;
  const F* pr = c.pos_p_rot + gidx * c.dim;
  if (c.dim == 3) {
    fo[0] = f[0]; fo[1] = f[1]; fo[2] = f[2];
    to[0] = t[0] + (pr[1] * f[2] - pr[2] * f[1]);
    to[1] = t[1] + (pr[2] * f[0] - pr[0] * f[2]);
    to[2] = t[2] + (pr[0] * f[1] - pr[1] * f[0]);
  } else {
    fo[0] = f[0]; fo[1] = f[1];
    to[0] = t[2] + (pr[0] * f[1] - pr[1] * f[0]);
  }
}

template <typename F, int LAW>
__global__ void __launch_bounds__(128) k_pair_force(Ctx<F> c) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n) return;
  const size_t off = (size_t)b * c.n;
  const LawCtx<F> lc = make_law_ctx(c, b);
  ForceVis<F, LAW> vis{c, lc, off};
  vis.a = load_sorted(c, off, k, LAW == JDB200_LAW_CUNDALLSTRACK);
  vis.idx = c.perm[off + k];
  vis.clump = c.sclump[off + k] & 0x7fffffff;
  vis.interact = c.interact && c.interact[b];
  vis.f[0] = vis.f[1] = vis.f[2] = F(0);
  vis.t[0] = vis.t[1] = vis.t[2] = F(0);
  walk_stencil<F>(c, b, k, nullptr, vis);
  store_force_torque(c, off + vis.idx, vis.f, vis.t);
  if (k == 0 && c.overflow) c.overflow[b] = (uint8_t)c.gi[b].hash_overflow;  // cell_list.py:463
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
  int idx, clump;
  bool interact;
  F e;
  __device__ __forceinline__ void cell(int, int s, int en) {
    for (int kj = s; kj < en; ++kj) {
      const int sc = c.sclump[off + kj];
      if (!pair_valid(c, off, idx, clump, sc, kj, interact)) continue;
      const Body<F> bj = load_sorted(c, off, kj, false);
      e += F(0.5) * pair_energy<F, LAW>(lc, a, bj);
    }
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
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t off = (size_t)b * c.n;
  F e = F(0);
  if (k < c.n) {
    const LawCtx<F> lc = make_law_ctx(c, b);
    EnergyVis<F, LAW> vis{c, lc, off};
    vis.a = load_sorted(c, off, k, false);
    vis.idx = c.perm[off + k];
    vis.clump = c.sclump[off + k] & 0x7fffffff;
    vis.interact = c.interact && c.interact[b];
    vis.e = F(0);
    walk_stencil<F>(c, b, k, nullptr, vis);
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
  int idx, clump;
  bool interact;
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
        const int sc = c.sclump[off + kj];
        if (!pair_valid(c, off, idx, clump, sc, kj, interact)) continue;
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
  vis.idx = c.perm[off + k];
  vis.clump = c.sclump[off + k] & 0x7fffffff;
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
  store_force_torque(c, off + i, f, t);
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
int build_partition(cudaStream_t s, Ctx<F>& c, const F* cell_size_override);

#define JDB_LAW_SWITCH(law, CALL)                                     \
  switch (law) {                                                      \
    case JDB200_LAW_SPRING: { constexpr int L = JDB200_LAW_SPRING; CALL; } break;           \
    case JDB200_LAW_HERTZ: { constexpr int L = JDB200_LAW_HERTZ; CALL; } break;             \
    default: { constexpr int L = JDB200_LAW_CUNDALLSTRACK; CALL; } break;                   \
  }

template <typename F>
int celllist_force(cudaStream_t s, Ctx<F>& c, bool rebuild) {
  if (c.n == 0) return 0;
  if (rebuild) {
    int rc = build_partition<F>(s, c, nullptr);
    if (rc) return rc;
  }
  const dim3 grid(cdiv(c.n, 128), c.batch);
  JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_pair_force<F, L>), grid, 128, s, c));
  return 0;
}

template <typename F>
int celllist_energy(cudaStream_t s, Ctx<F>& c, F* energy) {
  if (c.n == 0) return cudaMemsetAsync(energy, 0, sizeof(F) * c.batch, s) == cudaSuccess ? 0 : JDB200_ECUDA;
  int rc = build_partition<F>(s, c, nullptr);
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
  int rc = build_partition<F>(s, c, cs_nl);
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
  template int celllist_force<F>(cudaStream_t, Ctx<F>&, bool);                          \
  template int celllist_energy<F>(cudaStream_t, Ctx<F>&, F*);                           \
  template int celllist_neighbor_list<F>(cudaStream_t, Ctx<F>&, const F*, RT<F>::I*, uint8_t*); \
  template int naive_force<F>(cudaStream_t, Ctx<F>&);                                   \
  template int naive_energy<F>(cudaStream_t, Ctx<F>&, F*);
JDB_INST(float)
JDB_INST(double)

}  // namespace jdb
