// jaxdem_b200 — cell-list partition (K1-K3): grid setup, cell hash, dense counting
// sort / stable LSD radix sort, sorted shadow arrays.
//
// Replaces _get_spatial_partition (jaxdem/colliders/cell_list.py:35-87) and
// _grid_params (jaxdem/colliders/_partition.py:54-99).  Outputs are bit-identical
// to a stable sort of (hash, iota) on the reference's linear x-fastest hash.
#include "ctx.cuh"
#include "launch.cuh"
#include "scan.cuh"

namespace jdb {

// ---------------------------------------------------------------------------
// K0  setup: GridInfo per system, reset scan descriptors, zero the dense table.
// grid = (blocks, B); every block recomputes the (cheap) grid dims, block 0
// publishes them.
// ---------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256) k_setup(Ctx<F> c, const F* __restrict__ cell_size_override) {
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  __shared__ I s_gd[3], s_stride[3];
  __shared__ int s_ovf, s_dedup;
  __shared__ long long s_bound;
  __shared__ int s_dense;
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
    const bool dense = c.max_cells > 0 && !s_ovf && bound <= (double)c.max_cells;
    s_dense = dense;
    s_bound = dense ? (long long)bound : 0;
    s_dedup = 0;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    // can two stencil rows collide after the periodic wrap?  (rows equal, or span >= g)
    if (c.periodic) {
      const I* mask = c.mask + (size_t)b * c.M * c.dim;
      int dd = 0;
      for (int d = 0; d < c.dim; ++d) {
        I lo = 0, hi = 0;
        for (int m = 0; m < c.M; ++m) {
          I v = mask[m * c.dim + d];
          lo = v < lo ? v : lo;
          hi = v > hi ? v : hi;
        }
        if (hi - lo >= s_gd[d]) dd = 1;
      }
      for (int m = threadIdx.x; m < c.M && !dd; m += blockDim.x)
        for (int m2 = 0; m2 < m; ++m2) {
          bool same = true;
          for (int d = 0; d < c.dim; ++d) same &= mask[m * c.dim + d] == mask[m2 * c.dim + d];
          if (same) dd = 1;
        }
      if (dd) atomicOr(&s_dedup, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      GridInfo<I>& g = c.gi[b];
      for (int d = 0; d < 3; ++d) {
        g.gd[d] = s_gd[d];
        g.stride[d] = s_stride[d];
      }
      g.bound = s_bound;
      g.hash_overflow = s_ovf;
      g.need_dedup = s_dedup;
      g.dense = s_dense;
      g.dense_fail = 0;
      g.nl_overflow = 0;
      c.tile_counter[b] = 0;
      c.radix_skip[b] = 0;
    }
  }
  if (s_dense) {
    int* cs = c.cell_start + (size_t)b * (c.max_cells + 1);
    const long long rows = s_bound + 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows;
         i += (long long)gridDim.x * blockDim.x)
      cs[i] = 0;
    unsigned long long* ts = c.tile_state + (size_t)b * c.scan_tiles;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.scan_tiles; i += gridDim.x * blockDim.x)
      ts[i] = 0ull;
  }
}

template <typename F>
__device__ __forceinline__ void load_pos(const Ctx<F>& c, size_t gidx, F* p) {
  using T = RT<F>;
  const F* pc = c.pos_c + gidx * c.dim;
  const F* pr = c.pos_p_rot + gidx * c.dim;
  p[0] = T::add(pc[0], pr[0]);  // State.pos = pos_c + _pos_p_rot (state.py:295-304)
  p[1] = T::add(pc[1], pr[1]);
  p[2] = c.dim == 3 ? T::add(pc[2], pr[2]) : F(0);
}

// ---------------------------------------------------------------------------
// K1  cell hash (+ dense histogram with arrival rank).
// ---------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256) k_hash(Ctx<F> c, const F* __restrict__ cell_size_override) {
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const GridInfo<I> g = c.gi[b];
  const size_t gidx = (size_t)b * c.n + i;
  F p[3];
  load_pos(c, gidx, p);
  const F cs = cell_size_override ? cell_size_override[b] : c.cell_size[b];
  typename RT<F>::U h = 0;
  for (int d = 0; d < c.dim; ++d) {
    I cd = cell_coord<F, I>(p[d], c.anchor[b * c.dim + d], c.box[b * c.dim + d], cs, g.gd[d], c.periodic);
    h += (typename RT<F>::U)cd * (typename RT<F>::U)g.stride[d];
  }
  const I key = (I)h;
  c.key[gidx] = key;
  if (g.dense) {
    if (key >= 0 && (long long)key < g.bound) {
      c.rank[gidx] = atomicAdd(c.cell_start + (size_t)b * (c.max_cells + 1) + key, 1);
    } else {
      c.gi[b].dense_fail = 1;  // hash outside the dense table: fall back to the sorted path
    }
  }
}

// ---------------------------------------------------------------------------
// K2a  dense: exclusive scan of the per-cell counts (single pass, decoupled
// look-back), in place.  Also rejects cells above JDB200_DENSE_MAX_OCC.
// ---------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(512) k_scan(Ctx<F> c) {
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  GridInfo<I>& g = c.gi[b];
  if (!g.dense) return;
  const bool too_many =
      scan_tile(c.cell_start + (size_t)b * (c.max_cells + 1), g.bound + 1,
                c.tile_state + (size_t)b * c.scan_tiles, &c.tile_counter[b], JDB200_DENSE_MAX_OCC);
  if (too_many) g.dense_fail = 1;
}

template <typename I>
__device__ __forceinline__ bool use_dense(const GridInfo<I>& g) {
  return g.dense && !g.dense_fail;
}

// K2b  dense: place particle i at cell_start[h] + arrival rank (order inside a
// cell is fixed up by k_finalize).
template <typename F>
__global__ void __launch_bounds__(256) k_scatter(Ctx<F> c) {
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const GridInfo<I>& g = c.gi[b];
  if (!use_dense(g)) return;
  const size_t gidx = (size_t)b * c.n + i;
  const I key = c.key[gidx];
  const int* cs = c.cell_start + (size_t)b * (c.max_cells + 1);
  c.perm_b[(size_t)b * c.n + cs[key] + c.rank[gidx]] = (int)i;
}

// ---------------------------------------------------------------------------
// K2c  sorted fallback: stable LSD radix sort, 8 bits per pass, on
// (hash ^ signbit).  Three kernels per pass: count, scan, scatter.
// ---------------------------------------------------------------------------
template <typename I>
__device__ __forceinline__ unsigned radix_digit(I key, int pass) {
  using U = typename std::make_unsigned<I>::type;
  U u = (U)key ^ (U(1) << (sizeof(I) * 8 - 1));
  return (unsigned)((u >> (pass * 8)) & 0xff);
}

template <typename F>
__global__ void __launch_bounds__(256) k_radix_count(Ctx<F> c, const typename RT<F>::I* __restrict__ kin,
                                                      int pass) {
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  if (use_dense(c.gi[b])) return;
  __shared__ int hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kRadixTile;
  const I* k = kin + (size_t)b * c.n;
#pragma unroll
  for (int r = 0; r < kRadixTile / 256; ++r) {
    long long idx = base + r * 256 + threadIdx.x;
    if (idx < c.n) atomicAdd(&hist[radix_digit<I>(k[idx], pass)], 1);
  }
  __syncthreads();
  c.radix_counts[((size_t)b * 256 + threadIdx.x) * c.radix_blocks + blockIdx.x] = hist[threadIdx.x];
}

template <typename F>
__global__ void __launch_bounds__(1024) k_radix_scan(Ctx<F> c) {
  const int b = blockIdx.x;
  if (use_dense(c.gi[b])) return;
  int* cnt = c.radix_counts + (size_t)b * 256 * c.radix_blocks;
  const int total = 256 * c.radix_blocks;
  const int per = (total + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(lo + per, total);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += cnt[i];
  __shared__ int s_warp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;
  }
  __syncthreads();
  int run = s_warp[warp] + incl - sum;
  for (int i = lo; i < hi; ++i) {
    int v = cnt[i];
    cnt[i] = run;
    run += v;
  }
  __syncthreads();
  // a pass whose digit is the same for every key is a plain copy
  if (threadIdx.x == 0) c.radix_skip[b] = 0;
  __syncthreads();
  if (threadIdx.x < 256) {
    const int d = threadIdx.x;
    const int start = cnt[(size_t)d * c.radix_blocks];
    const int end = d == 255 ? (int)c.n : cnt[(size_t)(d + 1) * c.radix_blocks];
    if (end - start == (int)c.n && c.n > 0) c.radix_skip[b] = 1;
  }
}

template <typename F>
__global__ void __launch_bounds__(256) k_radix_scatter(Ctx<F> c, const typename RT<F>::I* __restrict__ kin,
                                                        const int* __restrict__ vin,
                                                        typename RT<F>::I* __restrict__ kout,
                                                        int* __restrict__ vout, int pass) {
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  if (use_dense(c.gi[b])) return;
  const size_t off = (size_t)b * c.n;
  const long long base = (long long)blockIdx.x * kRadixTile;
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
  __shared__ int whist[8][256];
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
  {  // per digit: global offset of this block + exclusive prefix over the warps
    const int d = threadIdx.x;
    int acc = c.radix_counts[((size_t)b * 256 + d) * c.radix_blocks + blockIdx.x];
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
}

// ---------------------------------------------------------------------------
// K3  finalize: deterministic order inside dense cells (rank = number of members
// with a smaller original index) + gather the sorted shadow arrays.
// ---------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256) k_finalize(Ctx<F> c, const int* __restrict__ sorted_perm) {
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n) return;
  const GridInfo<I>& g = c.gi[b];
  const size_t off = (size_t)b * c.n;
  int i;
  long long dest;
  if (use_dense(g)) {
    const int* cs = c.cell_start + (size_t)b * (c.max_cells + 1);
    const int* tmp = c.perm_b + off;
    i = tmp[k];
    const I key = c.key[off + i];
    const int s = cs[key], e = cs[key + 1];
    int r = 0;
    for (int kk = s; kk < e; ++kk) r += tmp[kk] < i;
    dest = s + r;
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
  c.skey[gd] = c.key[gi];
  F p[3];
  load_pos(c, gi, p);
  c.spos[gd] = Vec4<F>{p[0], p[1], p[2], c.rad[gi]};
  bool has_bond = false;
  for (int w = 0; w < c.W; ++w) has_bond |= c.bond_id[gi * c.W + w] >= 0;
  c.sclump[gd] = (int)c.clump_id[gi] | (has_bond ? 0x80000000 : 0);
  if (c.nmat > 1 || c.law != JDB200_LAW_SPRING) c.smat[gd] = (int)c.mat_id[gi];
  if (c.law == JDB200_LAW_CUNDALLSTRACK) {
    const F* v = c.vel + gi * c.dim;
    c.svel[gd] = Vec4<F>{v[0], v[1], c.dim == 3 ? v[2] : F(0), c.mass[gi]};
    const F* w = c.ang_vel + gi * c.A;
    c.sang[gd] = c.dim == 3 ? Vec4<F>{w[0], w[1], w[2], F(0)} : Vec4<F>{F(0), F(0), w[0], F(0)};
  }
}

// ---------------------------------------------------------------------------
// host: enqueue the partition
// ---------------------------------------------------------------------------
template <typename F>
int build_partition(cudaStream_t s, Ctx<F>& c, const F* cell_size_override) {
  using I = typename RT<F>::I;
  if (c.n == 0) return 0;
  const int B = c.batch;
  const int pb = cdiv(c.n, 256);
  int setup_blocks = c.max_cells > 0 ? (int)std::min<long long>(cdiv(c.max_cells + 1, 256 * 8), 1184) : 1;
  setup_blocks = std::max(setup_blocks, 1);
  JDB_LAUNCH(k_setup<F>, dim3(setup_blocks, B), 256, s, c, cell_size_override);
  JDB_LAUNCH(k_hash<F>, dim3(pb, B), 256, s, c, cell_size_override);
  const int* sorted_perm = nullptr;
  if (c.max_cells > 0) {
    JDB_LAUNCH(k_scan<F>, dim3(c.scan_tiles, B), 512, s, c);
    JDB_LAUNCH(k_scatter<F>, dim3(pb, B), 256, s, c);
  }
  if (c.max_cells == 0 || c.grid_mode == JDB200_GRID_AUTO) {
    // LSD passes: key -> key_b -> key_c -> key_b ...; perm: (iota) -> perm_c -> perm -> perm_c ...
    const int passes = (int)sizeof(I);
    const I* kin = c.key;
    const int* vin = nullptr;
    for (int p = 0; p < passes; ++p) {
      I* kout = (p & 1) ? c.key_c : c.key_b;
      int* vout = (p & 1) ? c.rank : c.perm_c;  // rank[] is free in the sorted path
      JDB_LAUNCH(k_radix_count<F>, dim3(c.radix_blocks, B), 256, s, c, kin, p);
      JDB_LAUNCH(k_radix_scan<F>, dim3(B), 1024, s, c);
      JDB_LAUNCH(k_radix_scatter<F>, dim3(c.radix_blocks, B), 256, s, c, kin, vin, kout, vout, p);
      kin = kout;
      vin = vout;
    }
    sorted_perm = vin;
  }
  JDB_LAUNCH(k_finalize<F>, dim3(pb, B), 256, s, c, sorted_perm);
  return 0;
}

template int build_partition<float>(cudaStream_t, Ctx<float>&, const float*);
template int build_partition<double>(cudaStream_t, Ctx<double>&, const double*);

}  // namespace jdb
