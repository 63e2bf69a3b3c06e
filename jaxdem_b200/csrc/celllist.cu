// jaxdem_b200 — cell-list partition (K1-K3): grid setup, cell hash (optionally fused with
// the velocity-Verlet kicks of the fused step driver), dense counting sort with a
// single-pass scan, cooperative stable LSD radix sort as the fallback, sorted shadow arrays.
//
// Replaces _get_spatial_partition (jaxdem/colliders/cell_list.py:35-87) and
// _grid_params (jaxdem/colliders/_partition.py:54-99).  Outputs are bit-identical
// to a stable sort of (hash, iota) on the reference's linear x-fastest hash.
#include "ctx.cuh"
#include "launch.cuh"
#include "scan.cuh"

namespace jdb {

// ---------------------------------------------------------------------------
// K0  setup: GridInfo per system, reset scan descriptors, zero the rows of the dense
// count table in use.  grid = (blocks, B); every block recomputes the (cheap) grid
// dims, block 0 also classifies the stencil and publishes GridInfo.
// ---------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256) k_setup(Ctx<F> c, const F* __restrict__ cell_size_override) {
  pdl_prologue();
  JDB_LIVE_ROWS(c);
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  if (c.gate && !c.gate[b]) {  // NeighborList: no rebuild for this system in this call
    if (b == 0 && blockIdx.x == 0 && threadIdx.x == 0) c.coop_bar[0] = 0u;
    return;
  }
  __shared__ I s_gd[3], s_stride[3];
  __shared__ int s_ovf, s_dense, s_hashed, s_hshift;
  __shared__ long long s_bound;
  if (threadIdx.x == 0) {
    const F cs = cell_size_override ? cell_size_override[b] : c.cell_size[b];
    grid_dims<F, I>(c.box + (size_t)b * c.dim, cs, c.dim, c.periodic, s_gd, s_stride, &s_ovf);
    // rows of the dense table in use: an upper bound on every particle hash + 1
    double bound = 0;
    if (c.periodic) {  // coords in [0, g] (g itself reachable through rounding, see cell_coord)
      for (int d = 0; d < c.dim; ++d) bound += (double)s_gd[d] * (double)s_stride[d];
      bound += 1;
    } else {
      bound = 1;
      for (int d = 0; d < c.dim; ++d) bound *= (double)s_gd[d];
    }
    bool dense = c.max_cells > 0 && !s_ovf && bound <= (double)c.max_cells;
    s_hashed = 0;
    s_hshift = 0;
    // a grid with more cells than the table has rows (dilute systems, the reference's own benchmark box): the
    // table is addressed by a hash of the cell key instead of falling back to the radix sort.  Not when the
    // caller wants the globally sorted permutation (partition export, MultiCellList AABB runs) or key windows.
    if (!dense && c.max_cells >= 4096 && !s_ovf && bound < 2147483647.0 && c.grid_mode == JDB200_GRID_AUTO &&
        !c.want_skey && !c.prune && c.win_len[0] == 0) {
      int bits = 12;
      while (bits < 30 && (2ll << bits) <= c.max_cells) ++bits;  // rows = 2^bits <= max_cells
      s_hashed = 1;
      s_hshift = 32 - bits;
      dense = true;
      bound = (double)(1ll << bits);
    }
    s_dense = dense;
    s_bound = dense ? (long long)bound : 0;
  }
  __syncthreads();
  if (s_dense) {  // zero count rows [0, bound] (the table stride is padded to 64 ints), or the windows in use
    for (int w = 0; w < 2; ++w) {
      long long lo = 0, len = s_bound + 1;
      if (c.win_len[0] > 0) {
        lo = c.win_lo[w] & ~3ll;
        len = c.win_len[w] > 0 ? min(c.win_lo[w] + c.win_len[w], s_bound + 1) - lo : 0;
      } else if (w == 1) {
        len = 0;
      }
      int4* cnt = reinterpret_cast<int4*>(c.cell_count + (size_t)b * c.cell_stride + lo);
      const long long rows4 = (len + 3) / 4;
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows4;
           i += (long long)gridDim.x * blockDim.x)
        cnt[i] = make_int4(0, 0, 0, 0);
    }
    unsigned long long* ts = c.tile_state + (size_t)b * c.scan_tiles;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.scan_tiles; i += gridDim.x * blockDim.x)
      ts[i] = 0ull;
  }
  if (blockIdx.x != 0) return;
  // fused driver: the _step_once refresh of inv_box_size (system.py:69-74) rides along
  if (c.fused && (int)threadIdx.x < c.dim)
    c.inv_box[b * c.dim + threadIdx.x] = RT<F>::div(F(1), c.box[b * c.dim + threadIdx.x]);
  // ... and so does the clock of _step_once (system.py:62-63): time += dt, step_count += 1
  if (c.tick && threadIdx.x == 0) {
    if (c.time) c.time[b] = RT<F>::add(c.time[b], c.dt[b]);
    if (c.step_count) c.step_count[b] += 1;
  }
  // ---- stencil classification (block 0 of each system) ----
  const I* mask = c.mask + (size_t)b * c.M * c.dim;
  // canonical cube: M == (2R+1)^D and row m == digits of m in base (2R+1), last axis fastest
  int w = 1;
  while (true) {
    long long pw = 1;
    for (int d = 0; d < c.dim; ++d) pw *= w;
    if (pw >= c.M) break;
    w += 2;
  }
  long long pw = 1;
  for (int d = 0; d < c.dim; ++d) pw *= w;
  const int R = (w - 1) / 2;
  int ok = (pw == c.M) && c.M > 0;
  if (ok) {
    for (int m = threadIdx.x; m < c.M; m += blockDim.x) {
      int rem = m;
      for (int d = c.dim - 1; d >= 0; --d) {
        if (mask[m * c.dim + d] != (I)(rem % w - R)) ok = 0;
        rem /= w;
      }
    }
  }
  const int canonical = __syncthreads_and(ok);
  // can two stencil rows collide after the periodic wrap?  (rows equal, or span >= g)
  int dd = 0;
  if (c.periodic) {
    for (int d = 0; d < c.dim; ++d) {
      I lo = 0, hi = 0;
      for (int m = 0; m < c.M; ++m) {
        I v = mask[m * c.dim + d];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
      }
      if (hi - lo >= s_gd[d]) dd = 1;
    }
    if (!canonical)
      for (int m = threadIdx.x; m < c.M && !dd; m += blockDim.x)
        for (int m2 = 0; m2 < m; ++m2) {
          bool same = true;
          for (int d = 0; d < c.dim; ++d) same &= mask[m * c.dim + d] == mask[m2 * c.dim + d];
          if (same) dd = 1;
        }
  }
  const int dedup = __syncthreads_or(dd);
  if (threadIdx.x == 0) {
    GridInfo<I>& g = c.gi[b];
    for (int d = 0; d < 3; ++d) {
      g.gd[d] = s_gd[d];
      g.stride[d] = s_stride[d];
    }
    g.bound = s_bound;
    g.hashed = s_hashed;
    g.hshift = s_hshift;
    g.hash_overflow = s_ovf;
    g.need_dedup = dedup;
    g.dense = s_dense;
    g.dense_fail = 0;
    g.nl_overflow = 0;
    g.canonical = canonical;
    g.range = R;
    g.any_bond = 0;
    g.any_ppr = 0;
    g.any_ext = 0;
    g.any_fixed = 0;
    g.edge = 0;
    c.tile_counter[b] = 0;
    c.radix_skip[b] = 0;
    if (b == 0) c.coop_bar[0] = 0u;  // arrival counter of the sort fallback's grid barrier
  }
}

template <typename I>
__device__ __forceinline__ bool use_dense(const GridInfo<I>& g) {
  return g.dense && !g.dense_fail;
}

// ---------------------------------------------------------------------------
// K1  cell hash (+ dense histogram with arrival rank), optionally fused with the
// linear velocity-Verlet updates of the fused step driver (sphere systems):
//   MODE 0  hash only                                  (collider hooks)
//   MODE 3  fused driver: step_before_force, then hash; the kicked velocity goes to the
//           (vx, vy, vz, mass) shadow record instead of State.vel (the pair kernel's epilogue
//           writes the final velocity), and the external buffers / fixed flags are scanned
//           so that epilogue knows whether it has to gather them at all
//   MODE 4  slab driver: the drift was done before the neighbour exchange; hash only, but
//           with the shadow record and the scans of MODE 3, so that the pair kernel's fused
//           epilogue (force manager + step_after_force) can follow
// References: velocity_verlet.py:57-61,92-95; force_manager.py:359-423;
// cell_list.py:51-62; state.py:295-304.
// ---------------------------------------------------------------------------
template <typename F, int D, int MODE>
__global__ void __launch_bounds__(256) k_hash(Ctx<F> c, const F* __restrict__ cell_size_override) {
  pdl_prologue();
  JDB_LIVE_ROWS(c);
  using T = RT<F>;
  using I = typename RT<F>::I;
  using U = typename RT<F>::U;
  constexpr int A = D == 3 ? 3 : 1;
  const int b = blockIdx.y;
  if (c.gate && !c.gate[b]) return;  // NeighborList: this system keeps its list in this call
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < c.n;
  const size_t gidx = (size_t)b * c.n + (live ? i : 0);
  bool bond = false, ppr_nz = false, ext_nz = false, fixed_any = false, edge = false;
  if (live) {
    // ---- all loads first (stores below may alias as far as the compiler knows) ----
    const GridInfo<I> g = c.gi[b];
    F pc[3] = {0, 0, 0}, pr[3] = {0, 0, 0}, f[3] = {0, 0, 0}, v[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < D; ++d) {
      pc[d] = __ldcs(&c.pos_c[gidx * D + d]);  // State leaves are streamed once: evict-first, so the
      // shadow records written below stay in L2 for k_scatter / k_finalize / the pair kernel
      if (!(c.promises & JDB200_PROMISE_NO_POS_P)) pr[d] = __ldcs(&c.pos_p_rot[gidx * D + d]);
      if (MODE != 0) {
        f[d] = __ldcs(&c.force[gidx * D + d]);
        v[d] = __ldcs(&c.vel[gidx * D + d]);
      }
    }
    const F rad = __ldcs(&c.rad[gidx]);
    F dt = F(0), mass = F(1);
    bool fixed = false;
    if (MODE != 0) {
      dt = c.dt[b];
      mass = __ldcs(&c.mass[gidx]);
      if (!(c.promises & JDB200_PROMISE_NO_FIXED)) fixed = __ldcs(&c.fixed[gidx]) != 0;
    }
    // streams the caller vouches for (jdb200_params.promises) are not read at all
    if (!(c.promises & JDB200_PROMISE_NO_BONDS))
      for (int w = 0; w < c.W; ++w) bond |= __ldcs(&c.bond_id[gidx * c.W + w]) >= 0;
    if (MODE == 3 || MODE == 4) {
      fixed_any = fixed;
      if (!(c.promises & JDB200_PROMISE_NO_EXT)) {
#pragma unroll
        for (int d = 0; d < D; ++d)
          ext_nz |= (__ldcs(&c.ext_force[gidx * D + d]) != F(0)) | (__ldcs(&c.ext_force_com[gidx * D + d]) != F(0));
#pragma unroll
        for (int a = 0; a < A; ++a) ext_nz |= __ldcs(&c.ext_torque[gidx * A + a]) != F(0);
      }
    }
    const F cs = cell_size_override ? cell_size_override[b] : c.cell_size[b];
    F anchor[3] = {0, 0, 0}, box[3] = {1, 1, 1};
#pragma unroll
    for (int d = 0; d < D; ++d) {
      anchor[d] = c.anchor[b * D + d];
      box[d] = c.box[b * D + d];
    }
    // ---- arithmetic ----
    if (MODE != 0 && MODE != 4) {
      const F sc = T::div(T::mul(dt, F(0.5)), mass);  // dt * 0.5 / mass
      const F free = fixed ? F(0) : F(1);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        v[d] = T::add(v[d], T::mul(T::mul(f[d], sc), free));  // step_before_force: kick ...
        pc[d] = T::add(pc[d], T::mul(dt, v[d]));              // ... and drift
      }
    }
    F p[3] = {0, 0, 0};
    U h = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      ppr_nz |= pr[d] != F(0);
      p[d] = T::add(pc[d], pr[d]);  // State.pos = pos_c + _pos_p_rot
      const I cd = cell_coord<F, I>(p[d], anchor[d], box[d], cs, g.gd[d], c.periodic);
      edge |= cd < 0 || cd >= g.gd[d];  // hash aliases another cell (or leaves the table)
      h += (U)cd * (U)g.stride[d];
    }
    const I key = (I)h;
    // ---- stores ----
    if (MODE != 0 && MODE != 4) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (MODE != 3) c.vel[gidx * D + d] = v[d];
        c.pos_c[gidx * D + d] = pc[d];
      }
    }
    if (MODE == 3 || MODE == 4) c.uvm[gidx] = Vec4<F>{v[0], v[1], v[2], mass};
    c.key[gidx] = key;
    c.urec[gidx] = Vec4<F>{p[0], p[1], p[2], rad};
    if (g.dense) {
      // hashed table: any key that fits the 32-bit records (keys of particles far outside a non-periodic grid may not)
      bool in_table = g.hashed ? (long long)key == (long long)(int)key : (key >= 0 && (long long)key < g.bound);
      if (in_table && c.win_len[0] > 0) {
        const long long k = (long long)key;
        in_table = (k >= c.win_lo[0] && k < c.win_lo[0] + c.win_len[0]) ||
                   (k >= c.win_lo[1] && k < c.win_lo[1] + c.win_len[1]);
      }
      if (in_table) {
        c.rank[gidx] = atomicAdd(c.cell_count + (size_t)b * c.cell_stride + table_row(g, (long long)key), 1);
      } else {
        c.gi[b].dense_fail = 1;  // hash outside the dense table: the sorted fallback takes over
      }
    }
  }
  // flags consumed by the pair kernels (one store per warp at most)
  if (__any_sync(0xffffffffu, bond) && (threadIdx.x & 31) == 0) c.gi[b].any_bond = 1;
  if (__any_sync(0xffffffffu, ppr_nz) && (threadIdx.x & 31) == 0) c.gi[b].any_ppr = 1;
  if (__any_sync(0xffffffffu, edge) && (threadIdx.x & 31) == 0) c.gi[b].edge = 1;
  if (MODE == 3 || MODE == 4) {
    if (__any_sync(0xffffffffu, ext_nz) && (threadIdx.x & 31) == 0) c.gi[b].any_ext = 1;
    if (__any_sync(0xffffffffu, fixed_any) && (threadIdx.x & 31) == 0) c.gi[b].any_fixed = 1;
  }
}

// ---------------------------------------------------------------------------
// K1 (vectorised): the same hook arithmetic, FOUR particles per thread, every State stream moved
// as 128-bit accesses — the (N, 3) row-major leaves as 3 float4 per 4 particles (SURVEY §7
// "layout").  f32, n % 4 == 0 (so each system's rows start 16-byte aligned); otherwise k_hash.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float4 ldcs4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

template <int D, int MODE>
__global__ void __launch_bounds__(128) k_hash4(Ctx<float> c, const float* __restrict__ cell_size_override) {
  pdl_prologue();
  JDB_LIVE_ROWS(c);
  using F = float;
  using T = RT<F>;
  using I = int32_t;
  using U = uint32_t;
  constexpr int A = D == 3 ? 3 : 1;
  constexpr int P = 4;  // particles per thread
  const int b = blockIdx.y;
  if (c.gate && !c.gate[b]) return;  // NeighborList: this system keeps its list in this call
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * P;
  const bool live = i0 < c.n;
  const size_t g0 = (size_t)b * c.n + (live ? i0 : 0);
  bool bond = false, ppr_nz = false, ext_nz = false, fixed_any = false, edge = false;
  if (live) {
    const GridInfo<I> g = c.gi[b];
    // ---- loads: P * D floats of every (N, D) leaf = D float4 ----
    F pc[P * D], pr[P * D], f[P * D], v[P * D], rad[P], mass[P];
    bool fixed[P];
    auto load_rows = [&](const F* base, F* out) {
#pragma unroll
      for (int q = 0; q < D; ++q) {
        const float4 t = ldcs4(base + g0 * D + 4 * q);
        out[4 * q] = t.x; out[4 * q + 1] = t.y; out[4 * q + 2] = t.z; out[4 * q + 3] = t.w;
      }
    };
    load_rows(c.pos_c, pc);
#pragma unroll
    for (int e = 0; e < P * D; ++e) pr[e] = f[e] = v[e] = F(0);
    if (!(c.promises & JDB200_PROMISE_NO_POS_P)) load_rows(c.pos_p_rot, pr);
    if (MODE != 0) {
      load_rows(c.force, f);
      load_rows(c.vel, v);
    }
    {
      const float4 t = ldcs4(c.rad + g0);
      rad[0] = t.x; rad[1] = t.y; rad[2] = t.z; rad[3] = t.w;
    }
    F dt = F(0);
#pragma unroll
    for (int p = 0; p < P; ++p) { mass[p] = F(1); fixed[p] = false; }
    if (MODE != 0) {
      dt = c.dt[b];
      const float4 t = ldcs4(c.mass + g0);
      mass[0] = t.x; mass[1] = t.y; mass[2] = t.z; mass[3] = t.w;
      if (!(c.promises & JDB200_PROMISE_NO_FIXED)) {
        const unsigned fx = __ldcs(reinterpret_cast<const unsigned*>(c.fixed + g0));
#pragma unroll
        for (int p = 0; p < P; ++p) fixed[p] = ((fx >> (8 * p)) & 0xffu) != 0u;
      }
    }
    if (!(c.promises & JDB200_PROMISE_NO_BONDS))
      for (int w = 0; w < P * c.W; ++w) bond |= __ldcs(&c.bond_id[g0 * c.W + w]) >= 0;
    if (MODE == 3 || MODE == 4) {
#pragma unroll
      for (int p = 0; p < P; ++p) fixed_any |= fixed[p];
      if (!(c.promises & JDB200_PROMISE_NO_EXT)) {
#pragma unroll
        for (int q = 0; q < D; ++q) {
          const float4 s1 = ldcs4(c.ext_force + g0 * D + 4 * q), s2 = ldcs4(c.ext_force_com + g0 * D + 4 * q);
          ext_nz |= s1.x != 0.f || s1.y != 0.f || s1.z != 0.f || s1.w != 0.f || s2.x != 0.f || s2.y != 0.f ||
                    s2.z != 0.f || s2.w != 0.f;
        }
#pragma unroll
        for (int q = 0; q < A; ++q) {
          const float4 s3 = ldcs4(c.ext_torque + g0 * A + 4 * q);
          ext_nz |= s3.x != 0.f || s3.y != 0.f || s3.z != 0.f || s3.w != 0.f;
        }
      }
    }
    const F cs = cell_size_override ? cell_size_override[b] : c.cell_size[b];
    F anchor[3] = {0, 0, 0}, box[3] = {1, 1, 1};
#pragma unroll
    for (int d = 0; d < D; ++d) {
      anchor[d] = c.anchor[b * D + d];
      box[d] = c.box[b * D + d];
    }
    // ---- arithmetic (identical, op for op, to k_hash) ----
    I key[P];
    F pos[P * D];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      if (MODE != 0 && MODE != 4) {
        const F sc = T::div(T::mul(dt, F(0.5)), mass[p]);
        const F free = fixed[p] ? F(0) : F(1);
#pragma unroll
        for (int d = 0; d < D; ++d) {
          v[p * D + d] = T::add(v[p * D + d], T::mul(T::mul(f[p * D + d], sc), free));
          pc[p * D + d] = T::add(pc[p * D + d], T::mul(dt, v[p * D + d]));
        }
      }
      U h = 0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        ppr_nz |= pr[p * D + d] != F(0);
        pos[p * D + d] = T::add(pc[p * D + d], pr[p * D + d]);
        const I cd = cell_coord<F, I>(pos[p * D + d], anchor[d], box[d], cs, g.gd[d], c.periodic);
        edge |= cd < 0 || cd >= g.gd[d];
        h += (U)cd * (U)g.stride[d];
      }
      key[p] = (I)h;
    }
    // ---- stores ----
    auto store_rows = [&](F* base, const F* in) {
#pragma unroll
      for (int q = 0; q < D; ++q)
        *reinterpret_cast<float4*>(base + g0 * D + 4 * q) = make_float4(in[4 * q], in[4 * q + 1], in[4 * q + 2], in[4 * q + 3]);
    };
    if (MODE != 0 && MODE != 4) {
      if (MODE != 3) store_rows(c.vel, v);
      store_rows(c.pos_c, pc);
    }
    int rk[P] = {0, 0, 0, 0};
    if (g.dense) {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        bool in_table = g.hashed || (key[p] >= 0 && (long long)key[p] < g.bound);  // (f32: keys are int32)
        if (in_table && c.win_len[0] > 0) {
          const long long k = (long long)key[p];
          in_table = (k >= c.win_lo[0] && k < c.win_lo[0] + c.win_len[0]) ||
                     (k >= c.win_lo[1] && k < c.win_lo[1] + c.win_len[1]);
        }
        if (in_table) rk[p] = atomicAdd(c.cell_count + (size_t)b * c.cell_stride + table_row(g, (long long)key[p]), 1);
        else c.gi[b].dense_fail = 1;
      }
      *reinterpret_cast<int4*>(c.rank + g0) = make_int4(rk[0], rk[1], rk[2], rk[3]);
    }
    *reinterpret_cast<int4*>(c.key + g0) = make_int4(key[0], key[1], key[2], key[3]);
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const F z = D == 3 ? pos[p * D + D - 1] : F(0), vz = D == 3 ? v[p * D + D - 1] : F(0);
      c.urec[g0 + p] = Vec4<F>{pos[p * D], pos[p * D + 1], z, rad[p]};
      if (MODE == 3 || MODE == 4) c.uvm[g0 + p] = Vec4<F>{v[p * D], v[p * D + 1], vz, mass[p]};
    }
  }
  if (__any_sync(0xffffffffu, bond) && (threadIdx.x & 31) == 0) c.gi[b].any_bond = 1;
  if (__any_sync(0xffffffffu, ppr_nz) && (threadIdx.x & 31) == 0) c.gi[b].any_ppr = 1;
  if (__any_sync(0xffffffffu, edge) && (threadIdx.x & 31) == 0) c.gi[b].edge = 1;
  if (MODE == 3 || MODE == 4) {
    if (__any_sync(0xffffffffu, ext_nz) && (threadIdx.x & 31) == 0) c.gi[b].any_ext = 1;
    if (__any_sync(0xffffffffu, fixed_any) && (threadIdx.x & 31) == 0) c.gi[b].any_fixed = 1;
  }
}

// ---------------------------------------------------------------------------
// K2a  dense: exclusive scan of the per-cell counts (single pass, decoupled
// look-back).  Also rejects cells above JDB200_DENSE_MAX_OCC.
// ---------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(512) k_scan(Ctx<F> c) {
  pdl_prologue();
  JDB_LIVE_ROWS(c);
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  if (c.gate && !c.gate[b]) return;  // NeighborList: this system keeps its list in this call
  GridInfo<I>& g = c.gi[b];
  if (!g.dense) return;
  const size_t co = (size_t)b * c.cell_stride;
  bool too_many;
  if (c.win_len[0] > 0) {  // only the windows of the table in use (slab decomposition): scanned as one array
    const long long end0 = min(c.win_lo[0] + c.win_len[0], (long long)g.bound + 1);
    const long long len_a = c.win_len[1] > 0 ? c.win_len[0] : max(end0 - c.win_lo[0], 0ll);
    const long long len_b = c.win_len[1] > 0 ? max(min(c.win_lo[1] + c.win_len[1], (long long)g.bound + 1) - c.win_lo[1], 0ll) : 0;
    too_many = scan_tile(c.cell_count + co, c.cell_start + co, len_a + len_b,
                         c.tile_state + (size_t)b * c.scan_tiles, &c.tile_counter[b], JDB200_DENSE_MAX_OCC,
                         len_a, c.win_lo[0], c.win_lo[1]);
  } else {
    too_many = scan_tile(c.cell_count + co, c.cell_start + co, g.bound + 1,
                         c.tile_state + (size_t)b * c.scan_tiles, &c.tile_counter[b],
                         JDB200_DENSE_MAX_OCC);
  }
  if (too_many) g.dense_fail = 1;
}

// K2b  dense: place particle i at cell_start[h] + arrival rank (order inside a cell is fixed
// up by k_finalize).  The particle's (x, y, z, rad) record travels with it: ONE 256-bit store
// per particle (f32: the record, the index and the key share a 32-byte sector) — a random
// access costs the L1TEX pipe the same 32 wavefronts per warp whatever its width, so the
// record is moved here, once, instead of being gathered again by k_finalize.
__device__ __forceinline__ void st256(void* p, float a0, float a1, float a2, float a3, float a4, float a5, float a6,
                                      float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3),
               "f"(a4), "f"(a5), "f"(a6), "f"(a7)
               : "memory");
}
__device__ __forceinline__ void st256(void* p, double a0, double a1, double a2, double a3) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a0), "d"(a1), "d"(a2), "d"(a3) : "memory");
}
__device__ __forceinline__ void ld256(const void* p, float* a) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7])
               : "l"(p));
}
__device__ __forceinline__ void ld256(const void* p, double* a) {
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a[0]), "=d"(a[1]), "=d"(a[2]), "=d"(a[3]) : "l"(p));
}

template <typename F>
__global__ void __launch_bounds__(256) k_scatter(Ctx<F> c) {
  pdl_prologue();
  JDB_LIVE_ROWS(c);
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  if (c.gate && !c.gate[b]) return;  // NeighborList: this system keeps its list in this call
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const GridInfo<I>& g = c.gi[b];
  if (!use_dense(g)) return;
  const size_t gidx = (size_t)b * c.n + i;
  const int key = (int)c.key[gidx];
  const Vec4<F> p = c.urec[gidx];
  const size_t slot = (size_t)b * c.n + c.cell_start[(size_t)b * c.cell_stride + table_row(g, (long long)key)] + c.rank[gidx];
  if (sizeof(F) == 4) {
    st256(c.arec + 32 * slot, (float)p.x, (float)p.y, (float)p.z, (float)p.w, __int_as_float((int)i),
          __int_as_float(key), 0.f, 0.f);
  } else {
    st256(c.arec + 32 * slot, (double)p.x, (double)p.y, (double)p.z, (double)p.w);
    c.slot_rec[slot] = make_int2((int)i, key);
  }
}

// ---------------------------------------------------------------------------
// K2c  sorted fallback: stable LSD radix sort, 8 bits per pass, on (hash ^ signbit), in ONE
// cooperative launch (count / scan / scatter phases separated by grid syncs).  Systems whose
// dense table worked are skipped; when every system is dense the kernel exits at once.
// ---------------------------------------------------------------------------
template <typename I>
__device__ __forceinline__ unsigned radix_digit(I key, int pass) {
  using U = typename std::make_unsigned<I>::type;
  U u = (U)key ^ (U(1) << (sizeof(I) * 8 - 1));
  return (unsigned)((u >> (pass * 8)) & 0xff);
}

template <typename F>
__device__ __forceinline__ void radix_count_tile(const Ctx<F>& c, int b, int tile,
                                                 const typename RT<F>::I* __restrict__ kin, int pass,
                                                 int* hist /*smem[256]*/) {
  using I = typename RT<F>::I;
  hist[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)tile * kRadixTile;
  const I* k = kin + (size_t)b * c.n;
#pragma unroll
  for (int r = 0; r < kRadixTile / 256; ++r) {
    long long idx = base + r * 256 + threadIdx.x;
    if (idx < c.n) atomicAdd(&hist[radix_digit<I>(k[idx], pass)], 1);
  }
  __syncthreads();
  c.radix_counts[((size_t)b * 256 + threadIdx.x) * c.radix_blocks + tile] = hist[threadIdx.x];
  __syncthreads();
}

// one block: exclusive scan over counts[digit][tile] in (digit, tile) order; thread d owns digit d
template <typename F>
__device__ __forceinline__ void radix_scan_block(const Ctx<F>& c, int b, int* s_warp /*smem[8]*/) {
  int* cnt = c.radix_counts + ((size_t)b * 256 + threadIdx.x) * c.radix_blocks;
  int sum = 0;
  for (int t = 0; t < c.radix_blocks; ++t) sum += cnt[t];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int woff = 0;
  for (int w2 = 0; w2 < warp; ++w2) woff += s_warp[w2];
  int run = woff + incl - sum;
  for (int t = 0; t < c.radix_blocks; ++t) {
    int v = cnt[t];
    cnt[t] = run;
    run += v;
  }
  // a pass whose digit is the same for every key is a plain copy
  if (threadIdx.x == 0) c.radix_skip[b] = 0;
  __syncthreads();
  if (sum == (int)c.n && c.n > 0) c.radix_skip[b] = 1;
  __syncthreads();
}

template <typename F>
__device__ __forceinline__ void radix_scatter_tile(const Ctx<F>& c, int b, int tile,
                                                   const typename RT<F>::I* __restrict__ kin,
                                                   const int* __restrict__ vin,
                                                   typename RT<F>::I* __restrict__ kout,
                                                   int* __restrict__ vout, int pass,
                                                   int (*whist)[256] /*smem[8][256]*/) {
  using I = typename RT<F>::I;
  const size_t off = (size_t)b * c.n;
  const long long base = (long long)tile * kRadixTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int R = kRadixTile / 256;  // rounds per warp
  if (c.radix_skip[b]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      long long idx = base + r * 256 + threadIdx.x;
      if (idx < c.n) {
        kout[off + idx] = kin[off + idx];
        vout[off + idx] = vin ? vin[off + idx] : (int)idx;
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&whist[0][0])[i] = 0;
  __syncthreads();
  I keys[R];
  int vals[R];
  int ranks[R];
  unsigned digs[R];
  // warp w owns the contiguous keys [base + w*R*32, base + (w+1)*R*32): stable order
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long idx = base + (long long)warp * (R * 32) + r * 32 + lane;
    const bool valid = idx < c.n;
    keys[r] = valid ? kin[off + idx] : I(0);
    vals[r] = valid ? (vin ? vin[off + idx] : (int)idx) : 0;
    const unsigned d = valid ? radix_digit<I>(keys[r], pass) : (256u + lane);
    digs[r] = d;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int before = __popc(peers & ((1u << lane) - 1u));
    int basecnt = 0;
    if (valid) basecnt = whist[warp][d];
    __syncwarp();
    if (valid && before == 0) whist[warp][d] = basecnt + __popc(peers);
    __syncwarp();
    ranks[r] = basecnt + before;
  }
  __syncthreads();
  {  // per digit: global offset of this tile + exclusive prefix over the warps
    const int d = threadIdx.x;
    int acc = c.radix_counts[((size_t)b * 256 + d) * c.radix_blocks + tile];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      int t = whist[w][d];
      whist[w][d] = acc;
      acc += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (digs[r] < 256u) {
      const int pos = whist[warp][digs[r]] + ranks[r];
      kout[off + pos] = keys[r];
      vout[off + pos] = vals[r];
    }
  }
  __syncthreads();
}

// LSD passes: key -> key_b -> key_c -> key_b ...; values: (iota) -> perm_c -> rank -> perm_c ...
// (rank[] is free in the sorted path).  The final permutation lands in perm_c or rank,
// see sorted_perm_buffer().
// Grid-wide barrier on a monotonic arrival counter (zeroed by k_setup): instance `phase` (1, 2, ...) is
// passed once phase * gridDim.x blocks have arrived.  The launch keeps every block resident (grid <= one
// wave, see build_partition), so the spin cannot starve a block that has not started.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& phase) {
  __syncthreads();
  ++phase;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned target = phase * gridDim.x;
    while (*((volatile unsigned*)counter) < target) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
}

template <typename F>
__global__ void __launch_bounds__(256) k_radix_sort(Ctx<F> c) {
  pdl_prologue();
  JDB_LIVE_ROWS(c);
  using I = typename RT<F>::I;
  __shared__ int whist[8][256];
  __shared__ int s_warp[8];
  constexpr int passes = (int)sizeof(I);
  unsigned phase = 0;
  struct { unsigned* ctr; unsigned* ph; __device__ void sync() { grid_barrier(ctr, *ph); } } grid{c.coop_bar, &phase};
  {  // idle launch (every system dense, or gated out): one strided look at the B descriptors, then exit — the
     // serial loop below would cost one dependent global load per system (1.7 ms at B = 4096)
    int need = 0;
    for (int b = threadIdx.x; b < c.batch; b += blockDim.x)
      need |= !((c.gate && !c.gate[b]) || use_dense(c.gi[b]));
    if (!__syncthreads_or(need)) return;  // block-uniform and grid-uniform: every block sees the same descriptors
  }
  for (int b = 0; b < c.batch; ++b) {
    if ((c.gate && !c.gate[b]) || use_dense(c.gi[b])) continue;  // grid-uniform
    const I* kin = c.key;
    const int* vin = nullptr;
    for (int p = 0; p < passes; ++p) {
      I* kout = (p & 1) ? c.key_c : c.key_b;
      int* vout = (p & 1) ? c.rank : c.perm_c;
      for (int tile = blockIdx.x; tile < c.radix_blocks; tile += gridDim.x)
        radix_count_tile<F>(c, b, tile, kin, p, &whist[0][0]);
      grid.sync();
      if (blockIdx.x == 0) radix_scan_block<F>(c, b, s_warp);
      grid.sync();
      for (int tile = blockIdx.x; tile < c.radix_blocks; tile += gridDim.x)
        radix_scatter_tile<F>(c, b, tile, kin, vin, kout, vout, p, whist);
      grid.sync();
      kin = kout;
      vin = vout;
    }
  }
}

template <typename F>
inline const int* sorted_perm_buffer(const Ctx<F>& c) {
  constexpr int passes = (int)sizeof(typename RT<F>::I);
  return ((passes - 1) & 1) ? c.rank : c.perm_c;
}

// ---------------------------------------------------------------------------
// K3  finalize: deterministic order inside dense cells (rank = number of members
// with a smaller original index) + gather the sorted shadow arrays.
// ---------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256) k_finalize(Ctx<F> c, const int* __restrict__ sorted_perm) {
  pdl_prologue();
  JDB_LIVE_ROWS(c);
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  if (c.gate && !c.gate[b]) return;  // NeighborList: this system keeps its list in this call
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n) return;
  const GridInfo<I>& g = c.gi[b];
  const size_t off = (size_t)b * c.n;
  int i;
  long long dest;
  bool dense = use_dense(g);
  Vec4<F> mine = Vec4<F>{0, 0, 0, 0};
  if (dense) {
    const int* cs = c.cell_start + (size_t)b * c.cell_stride;
    int key;
    if (sizeof(F) == 4) {
      float a[8];
      ld256(c.arec + 32 * (off + k), a);
      mine = Vec4<F>{(F)a[0], (F)a[1], (F)a[2], (F)a[3]};
      i = __float_as_int(a[4]);
      key = __float_as_int(a[5]);
    } else {
      double a[4];
      ld256(c.arec + 32 * (off + k), a);
      mine = Vec4<F>{(F)a[0], (F)a[1], (F)a[2], (F)a[3]};
      const int2 me = c.slot_rec[off + k];
      i = me.x;
      key = me.y;
    }
    if (!g.hashed) c.tmp_key[off + k] = key;  // cells stay in place: the key of arrival slot k is the key of final slot k
    const long long row = table_row(g, (long long)key);
    const int s = cs[row], e = cs[row + 1];
    int r = 0;
    if (g.hashed) {  // a row may hold several cells: order by (key, original index)
      if (e - s > 1) {
        if (sizeof(F) == 4) {
          const int* rec = reinterpret_cast<const int*>(c.arec + 32 * off) + 4;  // (index, key) of every arrival slot
          for (int kk = s; kk < e; ++kk) {
            const int ik = rec[8 * (size_t)kk], kk_key = rec[8 * (size_t)kk + 1];
            r += kk_key < key || (kk_key == key && ik < i);
          }
        } else {
          const int2* __restrict__ tmp = c.slot_rec + off;
          for (int kk = s; kk < e; ++kk) r += tmp[kk].y < key || (tmp[kk].y == key && tmp[kk].x < i);
        }
      }
      c.tmp_key[off + s + r] = key;
    } else if (e - s > 1 && c.order_id) {  // rank = members of the cell with a smaller order id (jdb200_state.order_id)
      const long long* oid = c.order_id + off;
      const long long mine_id = oid[i];
      if (sizeof(F) == 4) {
        const int* ids = reinterpret_cast<const int*>(c.arec + 32 * off) + 4;
        for (int kk = s; kk < e; ++kk) r += oid[ids[8 * (size_t)kk]] < mine_id;
      } else {
        const int2* __restrict__ tmp = c.slot_rec + off;
        for (int kk = s; kk < e; ++kk) r += oid[tmp[kk].x] < mine_id;
      }
    } else if (e - s > 1) {  // rank = members of the cell with a smaller original index
      if (sizeof(F) == 4) {
        const int* ids = reinterpret_cast<const int*>(c.arec + 32 * off) + 4;
        for (int kk = s; kk < e; ++kk) r += ids[8 * (size_t)kk] < i;
      } else {
        const int2* __restrict__ tmp = c.slot_rec + off;
        for (int kk = s; kk < e; ++kk) r += tmp[kk].x < i;
      }
    }
    dest = s + r;
    c.inv[off + i] = (int)dest;  // original index -> sorted slot (k_after un-permutes the row kernel's sums)
  } else if (sorted_perm) {
    i = sorted_perm[off + k];
    dest = k;
  } else {
    // JDB200_GRID_DENSE was forced but the data does not fit the dense table: keep
    // every access in bounds (identity order) and raise the overflow flag.
    i = (int)k;
    dest = k;
    if (k == 0) c.gi[b].hash_overflow = 1;
  }
  const size_t gi = off + i, gd = off + dest;
  c.perm[gd] = i;
  if (!dense || c.want_skey) c.skey[gd] = c.key[gi];
  c.spos[gd] = dense ? mine : c.urec[gi];
  if (c.clumps || g.any_bond) {
    bool has_bond = false;
    if (g.any_bond)
      for (int w = 0; w < c.W; ++w) has_bond |= c.bond_id[gi * c.W + w] >= 0;
    c.sclump[gd] = (int)c.clump_id[gi] | (has_bond ? 0x80000000 : 0);
  }
  if (c.nmat > 1) c.smat[gd] = (int)c.mat_id[gi];
  // (the fused flows keep the kicked velocity in the original-order record uvm[i]: k_after reads it there)
  if (c.law == JDB200_LAW_CUNDALLSTRACK) {
    const F* v = c.vel + gi * c.dim;
    if (c.fused) c.svel[gd] = c.uvm[gi];
    else c.svel[gd] = Vec4<F>{v[0], v[1], c.dim == 3 ? v[2] : F(0), c.mass[gi]};
    const F* w = c.ang_vel + gi * c.A;
    c.sang[gd] = c.dim == 3 ? Vec4<F>{w[0], w[1], w[2], F(0)} : Vec4<F>{F(0), F(0), w[0], F(0)};
  }
}

// ---------------------------------------------------------------------------
// host: enqueue the partition
// ---------------------------------------------------------------------------
static int coop_grid_limit(const void* fn, int block) {
  // co-resident blocks for a cooperative launch on the current device (queried per call:
  // cheap, and keeps the library free of global mutable state)
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, 0) != cudaSuccess) return 0;
  return sms * per_sm;
}

template <int D>
static bool launch_hash4(cudaStream_t s, Ctx<double>&, const double*, int) { return false; }
template <int D>
static bool launch_hash4(cudaStream_t s, Ctx<float>& c, const float* cso, int mode) {
  if (c.n % 4 != 0 || c.n_dev) return false;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!(al(c.pos_c) && al(c.pos_p_rot) && al(c.force) && al(c.vel) && al(c.rad) && al(c.mass) && al(c.ext_force) &&
        al(c.ext_force_com) && al(c.ext_torque) && (reinterpret_cast<uintptr_t>(c.fixed) & 3) == 0))
    return false;
  const dim3 grid(cdiv(c.n, 4 * 128), c.batch);
  auto go = [&]() -> int {
    if (mode == 0) JDB_LAUNCH((k_hash4<D, 0>), grid, 128, s, c, cso);
    else if (mode == 3) JDB_LAUNCH((k_hash4<D, 3>), grid, 128, s, c, cso);
    else JDB_LAUNCH((k_hash4<D, 4>), grid, 128, s, c, cso);
    return 0;
  };
  return go() == 0;
}

template <typename F, int D>
static int launch_hash(cudaStream_t s, Ctx<F>& c, const F* cso, int mode, bool ext) {
  // The four-particles-per-thread kernel wins where the hash runs on its own (mode 0: 321 vs 913 us on the 4 M-sphere
  // clump workload, whose members are neighbours in index AND in space) and loses where the integrator rides along
  // (modes 3 / 4: 33 vs 29 us at 1 M spheres, 102-128 registers at 22 % occupancy): measured, profiles/README.md.
  if (mode == 0 && launch_hash4<D>(s, c, cso, mode)) return 0;
  const dim3 grid(cdiv(c.n, 256), c.batch);
  if (mode == 0) JDB_LAUNCH((k_hash<F, D, 0>), grid, 256, s, c, cso);
  else if (mode == 3) JDB_LAUNCH((k_hash<F, D, 3>), grid, 256, s, c, cso);
  else if (mode == 4) JDB_LAUNCH((k_hash<F, D, 4>), grid, 256, s, c, cso);
  else return JDB200_EINVAL;
  return 0;
}

// hash_mode / ext: see k_hash (0 for the plain collider hooks)
template <typename F>
int build_partition(cudaStream_t s, Ctx<F>& c, const F* cell_size_override, int hash_mode, bool ext) {
  using I = typename RT<F>::I;
  if (c.n == 0) return 0;
  const int B = c.batch;
  const int pb = cdiv(c.n, 256);
  int setup_blocks = c.max_cells > 0 ? (int)std::min<long long>(cdiv(c.max_cells + 1, 256 * 16), 592) : 1;
  setup_blocks = std::max(setup_blocks, 1);
  JDB_LAUNCH(k_setup<F>, dim3(setup_blocks, B), 256, s, c, cell_size_override);
  int rc = c.dim == 3 ? launch_hash<F, 3>(s, c, cell_size_override, hash_mode, ext)
                      : launch_hash<F, 2>(s, c, cell_size_override, hash_mode, ext);
  if (rc) return rc;
  const int* sorted_perm = nullptr;
  if (c.max_cells > 0) {
    JDB_LAUNCH(k_scan<F>, dim3(c.scan_tiles, B), 512, s, c);
    JDB_LAUNCH(k_scatter<F>, dim3(pb, B), 256, s, c);
  }
  if (c.max_cells == 0 || c.grid_mode == JDB200_GRID_AUTO) {
    // ONE ordinary launch (programmatic stream serialization like every other kernel of the library; legal in
    // graph capture) with a software grid barrier between the count / scan / scatter phases; at most one
    // resident wave of blocks.  When every system is dense the blocks exit at once.
    const int limit = coop_grid_limit((const void*)k_radix_sort<F>, 256);
    if (limit <= 0) return JDB200_ECUDA;
    const int blocks = std::max(1, std::min(limit / 2, c.radix_blocks));
    JDB_LAUNCH(k_radix_sort<F>, dim3(blocks), 256, s, c);
    sorted_perm = sorted_perm_buffer(c);
  }
  JDB_LAUNCH(k_finalize<F>, dim3(pb, B), 256, s, c, sorted_perm);
  return 0;
}

template int build_partition<float>(cudaStream_t, Ctx<float>&, const float*, int, bool);
template int build_partition<double>(cudaStream_t, Ctx<double>&, const double*, int, bool);

}  // namespace jdb
