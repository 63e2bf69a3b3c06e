// jaxdem_b200 — slab decomposition: classify / pack / unpack kernels of the per-step
// neighbour exchange (jaxdem_b200/slab.py).  No reference equivalent: the
// reference runs one system on one device (jaxdem/system.py:60-98).
//
// After the drift every owned particle is classified by the cell layer of its last
// coordinate — the SAME arithmetic as the collider's hash (cell_coord, cell_list.py:55-60) —
// into: stays (and lies in the halo of the lower / upper face), leaves to the lower / upper
// neighbour, or strays (moved further than the halo: reported, never lost).  Leavers are
// written as full records into the message of their direction and as ghost records into `kept`
// (they stay behind as ghosts), and their row indices are listed as holes; halo particles are
// written as ghost records.  All lists keep the index order (block counts -> one-block scan ->
// in-block ballot ranks): bitwise repeatable.  After the exchange the owned rows are repaired IN
// PLACE — arrivals fill the lowest holes (or are appended), rows from the tail fill the holes
// that are left — so a step moves O(leavers + halo) records, not the whole slab.
#include "ctx.cuh"
#include "launch.cuh"

namespace jdb {

constexpr int kSlabBlock = 256;

template <typename F>
struct SlabRows {
  F *pos_c, *vel, *force, *ang_vel, *torque, *q_w, *q_xyz, *rad, *mass, *inertia;
  typename RT<F>::I* mat_id;
  uint8_t* fixed;
  long long* gid;
};

template <typename F>
inline SlabRows<F> slab_rows(const jdb200_slab_rows* r) {
  SlabRows<F> o;
  o.pos_c = (F*)r->pos_c; o.vel = (F*)r->vel; o.force = (F*)r->force; o.ang_vel = (F*)r->ang_vel;
  o.torque = (F*)r->torque; o.q_w = (F*)r->q_w; o.q_xyz = (F*)r->q_xyz; o.rad = (F*)r->rad;
  o.mass = (F*)r->mass; o.inertia = (F*)r->inertia;
  o.mat_id = (typename RT<F>::I*)r->mat_id; o.fixed = (uint8_t*)r->fixed; o.gid = (long long*)r->gid;
  return o;
}

// message layout (jaxdem_b200/slab.py: message_layout): int64 header[8], full records
// (floats, then int64 triples gid / mat / fixed), ghost records (floats, then int64 pairs
// gid / mat); every section 16-byte aligned
template <typename F, int D>
struct SlabMsg {
  static constexpr int A = D == 3 ? 3 : 1;
  static constexpr int WF = 3 * D + 3 * A + 1 + 3 + 2;  // pos_c vel force | ang_vel torque inertia | q_w | q_xyz | rad mass
  static constexpr int WG = 2 * D + A + 2;              // pos_c vel | ang_vel | rad mass
  long long* header;
  F* mig_f;
  long long* mig_i;
  F* gh_f;
  long long* gh_i;
  __host__ __device__ static size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
  __host__ __device__ static size_t bytes(long long cap_m, long long cap_g) {
    size_t o = 64;
    o = align16(o + (size_t)cap_m * WF * sizeof(F));
    o = align16(o + (size_t)cap_m * 3 * 8);
    o = align16(o + (size_t)cap_g * WG * sizeof(F));
    o = align16(o + (size_t)cap_g * 2 * 8);
    return o;
  }
  __host__ __device__ SlabMsg(void* base, long long cap_m, long long cap_g) {
    char* b = (char*)base;
    size_t o = 0;
    header = (long long*)b;
    o = 64;
    mig_f = (F*)(b + o);
    o = align16(o + (size_t)cap_m * WF * sizeof(F));
    mig_i = (long long*)(b + o);
    o = align16(o + (size_t)cap_m * 3 * 8);
    gh_f = (F*)(b + o);
    o = align16(o + (size_t)cap_g * WG * sizeof(F));
    gh_i = (long long*)(b + o);
  }
};

template <typename F, int D>
__device__ __forceinline__ void write_full(const SlabRows<F>& s, long long i, F* f, long long* iv) {
  constexpr int A = D == 3 ? 3 : 1;
  int c = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) f[c++] = s.pos_c[i * D + d];
#pragma unroll
  for (int d = 0; d < D; ++d) f[c++] = s.vel[i * D + d];
#pragma unroll
  for (int d = 0; d < D; ++d) f[c++] = s.force[i * D + d];
#pragma unroll
  for (int a = 0; a < A; ++a) f[c++] = s.ang_vel[i * A + a];
#pragma unroll
  for (int a = 0; a < A; ++a) f[c++] = s.torque[i * A + a];
#pragma unroll
  for (int a = 0; a < A; ++a) f[c++] = s.inertia[i * A + a];
  f[c++] = s.q_w[i];
#pragma unroll
  for (int a = 0; a < 3; ++a) f[c++] = s.q_xyz[i * 3 + a];
  f[c++] = s.rad[i];
  f[c++] = s.mass[i];
  iv[0] = s.gid[i];
  iv[1] = (long long)s.mat_id[i];
  iv[2] = (long long)s.fixed[i];
}
template <typename F, int D>
__device__ __forceinline__ void read_full(const SlabRows<F>& s, long long i, const F* f, const long long* iv) {
  constexpr int A = D == 3 ? 3 : 1;
  int c = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) s.pos_c[i * D + d] = f[c++];
#pragma unroll
  for (int d = 0; d < D; ++d) s.vel[i * D + d] = f[c++];
#pragma unroll
  for (int d = 0; d < D; ++d) s.force[i * D + d] = f[c++];
#pragma unroll
  for (int a = 0; a < A; ++a) s.ang_vel[i * A + a] = f[c++];
#pragma unroll
  for (int a = 0; a < A; ++a) s.torque[i * A + a] = f[c++];
#pragma unroll
  for (int a = 0; a < A; ++a) s.inertia[i * A + a] = f[c++];
  s.q_w[i] = f[c++];
#pragma unroll
  for (int a = 0; a < 3; ++a) s.q_xyz[i * 3 + a] = f[c++];
  s.rad[i] = f[c++];
  s.mass[i] = f[c++];
  s.gid[i] = iv[0];
  s.mat_id[i] = (typename RT<F>::I)iv[1];
  s.fixed[i] = (uint8_t)iv[2];
}
template <typename F, int D>
__device__ __forceinline__ void write_ghost(const SlabRows<F>& s, long long i, F* f, long long* iv) {
  constexpr int A = D == 3 ? 3 : 1;
  int c = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) f[c++] = s.pos_c[i * D + d];
#pragma unroll
  for (int d = 0; d < D; ++d) f[c++] = s.vel[i * D + d];
#pragma unroll
  for (int a = 0; a < A; ++a) f[c++] = s.ang_vel[i * A + a];
  f[c++] = s.rad[i];
  f[c++] = s.mass[i];
  iv[0] = s.gid[i];
  iv[1] = (long long)s.mat_id[i];
}
// a ghost row: what the force laws read; the rest of the row is set to inert values
template <typename F, int D>
__device__ __forceinline__ void read_ghost(const SlabRows<F>& s, long long i, const F* f, const long long* iv) {
  constexpr int A = D == 3 ? 3 : 1;
  int c = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) s.pos_c[i * D + d] = f[c++];
#pragma unroll
  for (int d = 0; d < D; ++d) s.vel[i * D + d] = f[c++];
#pragma unroll
  for (int a = 0; a < A; ++a) s.ang_vel[i * A + a] = f[c++];
  s.rad[i] = f[c++];
  s.mass[i] = f[c++];
  s.gid[i] = iv[0];
  s.mat_id[i] = (typename RT<F>::I)iv[1];
  s.fixed[i] = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) s.force[i * D + d] = F(0);
#pragma unroll
  for (int a = 0; a < A; ++a) {
    s.torque[i * A + a] = F(0);
    s.inertia[i * A + a] = F(1);
  }
  s.q_w[i] = F(1);
#pragma unroll
  for (int a = 0; a < 3; ++a) s.q_xyz[i * 3 + a] = F(0);
}

struct SlabGeom {
  long long n;
  int n_layers, lo, up, R;
  long long cap_m, cap_g;
};

// category bits: 1 halo-lo, 2 halo-up, 4 leave-lo, 8 leave-up, 16 stray
// INTEGRATE: VelocityVerlet.step_before_force (velocity_verlet.py:57-61; the arithmetic of
// k_linear) rides along, so the drift and the classification read the positions once.
template <typename F, int D, bool INTEGRATE>
__global__ void __launch_bounds__(kSlabBlock) k_slab_classify(SlabGeom gm, SlabRows<F> rows, const F* __restrict__ dtp,
                                                               const F* __restrict__ anchor, const F* __restrict__ box,
                                                               const F* __restrict__ cell_size,
                                                               uint8_t* __restrict__ cat, int* __restrict__ bc) {
  pdl_prologue();
  using I = typename RT<F>::I;
  using T = RT<F>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c8 = 0;
  bool live = i < gm.n;
  if (live) {
    F* pos_c = rows.pos_c;
    if (INTEGRATE) {
      const F dt = dtp[0];
      const F sc = T::div(T::mul(dt, F(0.5)), rows.mass[i]);  // dt * 0.5 / mass
      const F free = rows.fixed[i] ? F(0) : F(1);
      F v[D], p[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        v[d] = T::add(rows.vel[i * D + d], T::mul(T::mul(rows.force[i * D + d], sc), free));
        p[d] = T::add(pos_c[i * D + d], T::mul(dt, v[d]));
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        rows.vel[i * D + d] = v[d];
        pos_c[i * D + d] = p[d];
      }
    }
    const F B = box[D - 1];
    // grid size of the last axis, as _grid_params computes it (periodic: floor(B / cs), >= 1)
    I g = RT<F>::to_int(RT<F>::floor(RT<F>::div(B, cell_size[0])));
    g = g < I(1) ? I(1) : g;
    I c = cell_coord<F, I>(pos_c[i * D + (D - 1)], anchor[D - 1], B, cell_size[0], g, true);
    c = c < 0 ? I(0) : (c >= g ? g - 1 : c);
    const int G = gm.n_layers, ci = (int)c;
    if ((int)g != G) {
      c8 = 16;  // the box changed under the static slab layout
    } else if (ci >= gm.lo && ci < gm.up) {
      if (ci < gm.lo + gm.R) c8 |= 1;
      if (ci >= gm.up - gm.R) c8 |= 2;
    } else {
      const int below = ((gm.lo - ci) % G + G) % G;        // layers below my lower face (periodic)
      const int above = ((ci - (gm.up - 1)) % G + G) % G;  // layers above my upper face
      if (below >= 1 && below <= gm.R) c8 = 4;
      else if (above >= 1 && above <= gm.R) c8 = 8;
      else c8 = 16;
    }
    cat[i] = (uint8_t)c8;
  }
  const bool stay = live && !(c8 & 12);  // strays stay (and are reported): no particle is ever dropped
  const int n0 = __syncthreads_count(stay);
  const int n1 = __syncthreads_count(live && (c8 & 4));
  const int n2 = __syncthreads_count(live && (c8 & 8));
  const int n3 = __syncthreads_count(stay && (c8 & 1));
  const int n4 = __syncthreads_count(stay && (c8 & 2));
  const int n5 = __syncthreads_count(live && (c8 & 16));
  if (threadIdx.x == 0) {
    int* o = bc + (size_t)blockIdx.x * 8;
    o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5;
  }
}

// exclusive scan of the block counts (one block: thread t owns a contiguous chunk of blocks,
// the 1024 chunk sums are scanned with shuffles), totals into the message headers and the
// local header
__global__ void __launch_bounds__(1024) k_slab_scan(int nblocks, int* __restrict__ bc, long long* __restrict__ hdr_lo,
                                                    long long* __restrict__ hdr_up, long long* __restrict__ hdr_local) {
  pdl_prologue();
  __shared__ int s_warp[32][6];
  __shared__ int s_tot[6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (nblocks + 1023) / 1024;
  const int b0 = min((int)threadIdx.x * per, nblocks), b1 = min(b0 + per, nblocks);
  int sum[6] = {0, 0, 0, 0, 0, 0};
  for (int b = b0; b < b1; ++b) {
#pragma unroll
    for (int j = 0; j < 6; ++j) sum[j] += bc[(size_t)b * 8 + j];
  }
  int excl[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    int incl = sum[j];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp][j] = incl;
    excl[j] = incl - sum[j];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int v = s_warp[lane][j];
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      s_warp[lane][j] = incl - v;
      if (lane == 31) s_tot[j] = incl;
    }
  }
  __syncthreads();
  int run[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) run[j] = excl[j] + s_warp[warp][j];
  for (int b = b0; b < b1; ++b) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int v = bc[(size_t)b * 8 + j];
      bc[(size_t)b * 8 + j] = run[j];
      run[j] += v;
    }
  }
  if (threadIdx.x == 0) {
    // header: [0] full records, [1] ghost records, [2] strays seen by the sender
    hdr_lo[0] = s_tot[1]; hdr_lo[1] = s_tot[3]; hdr_lo[2] = s_tot[5];
    hdr_up[0] = s_tot[2]; hdr_up[1] = s_tot[4]; hdr_up[2] = s_tot[5];
    hdr_local[0] = s_tot[0]; hdr_local[1] = s_tot[1]; hdr_local[2] = s_tot[2]; hdr_local[3] = s_tot[5];
  }
}

constexpr int kPackLists = 4;  // leave-lo, leave-up, halo-lo, halo-up (block-count columns 1..4)

template <typename F, int D>
__global__ void __launch_bounds__(kSlabBlock) k_slab_pack(SlabGeom gm, SlabRows<F> src, const uint8_t* __restrict__ cat,
                                                           const int* __restrict__ bc, void* msg_lo, void* msg_up,
                                                           void* kept, int* __restrict__ holes) {
  pdl_prologue();
  using M = SlabMsg<F, D>;
  __shared__ int s_w[kSlabBlock / 32][kPackLists];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < gm.n;
  const int c8 = live ? cat[i] : 0;
  const bool stay = live && !(c8 & 12);
  const bool fl[kPackLists] = {live && (c8 & 4) != 0, live && (c8 & 8) != 0, stay && (c8 & 1) != 0,
                               stay && (c8 & 2) != 0};
  if (!__syncthreads_or(fl[0] | fl[1] | fl[2] | fl[3])) return;  // interior block: nothing to pack
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int rank[kPackLists];
#pragma unroll
  for (int j = 0; j < kPackLists; ++j) {
    const unsigned m = __ballot_sync(0xffffffffu, fl[j]);
    rank[j] = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_w[warp][j] = __popc(m);
  }
  __syncthreads();
  const int* base = bc + (size_t)blockIdx.x * 8 + 1;
#pragma unroll
  for (int j = 0; j < kPackLists; ++j) {
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_w[w][j];
    rank[j] += woff + base[j];
  }
  if (!live) return;
  const M lo(msg_lo, gm.cap_m, gm.cap_g), up(msg_up, gm.cap_m, gm.cap_g);
  // kept: ghost records of the leavers, lower direction in rows [0, cap_m), upper in [cap_m, 2 cap_m);
  // holes: their row indices, same split
  const M kp(kept, 0, 2 * gm.cap_m);
  if (fl[0] && rank[0] < gm.cap_m) {
    write_full<F, D>(src, i, lo.mig_f + (size_t)rank[0] * M::WF, lo.mig_i + (size_t)rank[0] * 3);
    write_ghost<F, D>(src, i, kp.gh_f + (size_t)rank[0] * M::WG, kp.gh_i + (size_t)rank[0] * 2);
    holes[rank[0]] = (int)i;
  }
  if (fl[1] && rank[1] < gm.cap_m) {
    write_full<F, D>(src, i, up.mig_f + (size_t)rank[1] * M::WF, up.mig_i + (size_t)rank[1] * 3);
    write_ghost<F, D>(src, i, kp.gh_f + (size_t)(gm.cap_m + rank[1]) * M::WG, kp.gh_i + (size_t)(gm.cap_m + rank[1]) * 2);
    holes[gm.cap_m + rank[1]] = (int)i;
  }
  if (fl[2] && rank[2] < gm.cap_g)
    write_ghost<F, D>(src, i, lo.gh_f + (size_t)rank[2] * M::WG, lo.gh_i + (size_t)rank[2] * 2);
  if (fl[3] && rank[3] < gm.cap_g)
    write_ghost<F, D>(src, i, up.gh_f + (size_t)rank[3] * M::WG, up.gh_i + (size_t)rank[3] * 2);
}

// counts of one exchange, known on the host after it: owned rows before, arrivals from the
// lower / upper neighbour, leavers to the lower / upper neighbour, halo of the lower / upper
struct SlabCounts {
  long long n_old, a_lo, a_up, k_lo, k_up, g_lo, g_up;
};

// holes[0, k_lo) and holes[cap_m, cap_m + k_up) are sorted; merged (sorted) list -> holes[2 cap_m ...)
__global__ void __launch_bounds__(kSlabBlock) k_slab_merge_holes(SlabGeom gm, SlabCounts cn, int* __restrict__ holes) {
  pdl_prologue();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long l = cn.k_lo + cn.k_up;
  if (t >= l) return;
  const bool first = t < cn.k_lo;
  const int* mine = first ? holes : holes + gm.cap_m;
  const int* other = first ? holes + gm.cap_m : holes;
  const long long r = first ? t : t - cn.k_lo, no = first ? cn.k_up : cn.k_lo;
  const int v = mine[r];
  long long lo = 0, hi = no;  // number of entries of the other list below v (indices are distinct)
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (other[mid] < v) lo = mid + 1; else hi = mid;
  }
  holes[2 * gm.cap_m + r + lo] = v;
}

template <typename F, int D>
__device__ __forceinline__ void copy_row(const SlabRows<F>& s, long long from, long long to) {
  F f[SlabMsg<F, D>::WF];
  long long iv[3];
  write_full<F, D>(s, from, f, iv);
  read_full<F, D>(s, to, f, iv);
}

// more leavers than arrivals: the holes the arrivals do not fill are R = sorted_holes[a, l); the new
// row count is n' = n_old - (l - a).  Rows of the tail [n', n_old) that are not holes move, in index
// order, into the holes below n' (a prefix of R).  One block: the tail is l - a rows long.
template <typename F, int D>
__global__ void __launch_bounds__(1024) k_slab_tail_move(SlabGeom gm, SlabRows<F> rows, SlabCounts cn,
                                                         const int* __restrict__ holes) {
  pdl_prologue();
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const long long l = cn.k_lo + cn.k_up, a = cn.a_lo + cn.a_up;
  const long long nr = l - a, n_new = cn.n_old - nr;
  const int* R = holes + 2 * gm.cap_m + a;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long t0 = 0; t0 < nr; t0 += 1024) {
    const long long row = n_new + t0 + threadIdx.x;
    bool mover = false;
    if (row < cn.n_old) {
      long long lo = 0, hi = nr;  // is `row` one of the holes?
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (R[mid] < row) lo = mid + 1; else hi = mid;
      }
      mover = !(lo < nr && R[lo] == row);
    }
    const unsigned m = __ballot_sync(0xffffffffu, mover);
    int rank = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) woff += s_warp[w];
      tot += s_warp[w];
    }
    rank += woff + s_base;
    if (mover) copy_row<F, D>(rows, row, R[rank]);
    __syncthreads();
    if (threadIdx.x == 0) s_base += tot;
    __syncthreads();
  }
}

// arrivals from the lower, then the upper neighbour: into the lowest holes, the rest appended;
// then the ghost rows behind the owned rows: leavers kept behind (lower, upper), halo of the
// lower, halo of the upper neighbour
template <typename F, int D>
__global__ void __launch_bounds__(kSlabBlock) k_slab_unpack(SlabGeom gm, SlabRows<F> dst, SlabCounts cn,
                                                             const void* from_lo, const void* from_up, const void* kept,
                                                             const int* __restrict__ holes) {
  pdl_prologue();
  using M = SlabMsg<F, D>;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const M lo((void*)from_lo, gm.cap_m, gm.cap_g), up((void*)from_up, gm.cap_m, gm.cap_g), kp((void*)kept, 0, 2 * gm.cap_m);
  const long long l = cn.k_lo + cn.k_up, a = cn.a_lo + cn.a_up;
  if (t < a) {
    const long long row = t < l ? (long long)holes[2 * gm.cap_m + t] : cn.n_old + (t - l);
    if (t < cn.a_lo) read_full<F, D>(dst, row, lo.mig_f + (size_t)t * M::WF, lo.mig_i + (size_t)t * 3);
    else read_full<F, D>(dst, row, up.mig_f + (size_t)(t - cn.a_lo) * M::WF, up.mig_i + (size_t)(t - cn.a_lo) * 3);
    return;
  }
  t -= a;
  const long long row = cn.n_old - l + a + t;
  if (t < cn.k_lo) { read_ghost<F, D>(dst, row, kp.gh_f + (size_t)t * M::WG, kp.gh_i + (size_t)t * 2); return; }
  t -= cn.k_lo;
  if (t < cn.k_up) {
    read_ghost<F, D>(dst, row, kp.gh_f + (size_t)(gm.cap_m + t) * M::WG, kp.gh_i + (size_t)(gm.cap_m + t) * 2);
    return;
  }
  t -= cn.k_up;
  if (t < cn.g_lo) { read_ghost<F, D>(dst, row, lo.gh_f + (size_t)t * M::WG, lo.gh_i + (size_t)t * 2); return; }
  t -= cn.g_lo;
  if (t < cn.g_up) read_ghost<F, D>(dst, row, up.gh_f + (size_t)t * M::WG, up.gh_i + (size_t)t * 2);
}

static inline int slab_check(const jdb200_slab_desc* d) {
  if (!d) return JDB200_ENULL;
  if (d->dim != 2 && d->dim != 3) return JDB200_EINVAL;
  if (d->dtype != JDB200_F32 && d->dtype != JDB200_F64) return JDB200_EINVAL;
  if (d->n < 0 || d->cap_mig < 0 || d->cap_ghost < 0 || d->n_layers < 1 || d->search_range < 1) return JDB200_EINVAL;
  if (d->lo_layer < 0 || d->up_layer > d->n_layers || d->lo_layer >= d->up_layer) return JDB200_EINVAL;
  return 0;
}
static inline SlabGeom slab_geom(const jdb200_slab_desc* d) {
  return SlabGeom{d->n, d->n_layers, d->lo_layer, d->up_layer, d->search_range, d->cap_mig, d->cap_ghost};
}

template <typename F, int D>
int slab_pack(cudaStream_t s, const jdb200_slab_desc* d, const jdb200_slab_rows* rows, void* msg_lo, void* msg_up,
              void* kept, void* holes, void* header_local, void* scratch) {
  using M = SlabMsg<F, D>;
  const SlabGeom gm = slab_geom(d);
  const int nb = std::max(1, cdiv(d->n, kSlabBlock));
  uint8_t* cat = (uint8_t*)scratch;
  int* bc = (int*)((char*)scratch + (((size_t)nb * kSlabBlock + 255) & ~size_t(255)));
  if (d->dt)
    JDB_LAUNCH((k_slab_classify<F, D, true>), dim3(nb), kSlabBlock, s, gm, slab_rows<F>(rows), (const F*)d->dt,
               (const F*)d->anchor, (const F*)d->box_size, (const F*)d->cell_size, cat, bc);
  else
    JDB_LAUNCH((k_slab_classify<F, D, false>), dim3(nb), kSlabBlock, s, gm, slab_rows<F>(rows), (const F*)nullptr,
               (const F*)d->anchor, (const F*)d->box_size, (const F*)d->cell_size, cat, bc);
  const M lo(msg_lo, gm.cap_m, gm.cap_g), up(msg_up, gm.cap_m, gm.cap_g);
  JDB_LAUNCH(k_slab_scan, dim3(1), 1024, s, nb, bc, lo.header, up.header, (long long*)header_local);
  JDB_LAUNCH((k_slab_pack<F, D>), dim3(nb), kSlabBlock, s, gm, slab_rows<F>(rows), cat, bc, msg_lo, msg_up, kept,
             (int*)holes);
  return 0;
}

template <typename F, int D>
int slab_unpack(cudaStream_t s, const jdb200_slab_desc* d, const jdb200_slab_rows* rows, const int64_t* counts,
                const void* from_lo, const void* from_up, const void* kept, void* holes) {
  const SlabCounts cn{counts[0], counts[1], counts[2], counts[3], counts[4], counts[5], counts[6]};
  const SlabGeom gm = slab_geom(d);
  const long long l = cn.k_lo + cn.k_up, a = cn.a_lo + cn.a_up;
  if (l > 0) JDB_LAUNCH(k_slab_merge_holes, dim3(cdiv(l, kSlabBlock)), kSlabBlock, s, gm, cn, (int*)holes);
  if (l > a) JDB_LAUNCH((k_slab_tail_move<F, D>), dim3(1), 1024, s, gm, slab_rows<F>(rows), cn, (const int*)holes);
  const long long tot = a + cn.k_lo + cn.k_up + cn.g_lo + cn.g_up;
  if (tot > 0)
    JDB_LAUNCH((k_slab_unpack<F, D>), dim3(cdiv(tot, kSlabBlock)), kSlabBlock, s, gm, slab_rows<F>(rows), cn, from_lo,
               from_up, kept, (const int*)holes);
  return 0;
}

}  // namespace jdb

using namespace jdb;

#define SLAB_DISPATCH(CALL)                                   \
  if (d->dtype == JDB200_F32) {                               \
    using F = float;                                          \
    if (d->dim == 3) { constexpr int D = 3; return CALL; }    \
    constexpr int D = 2;                                      \
    return CALL;                                              \
  } else {                                                    \
    using F = double;                                         \
    if (d->dim == 3) { constexpr int D = 3; return CALL; }    \
    constexpr int D = 2;                                      \
    return CALL;                                              \
  }

extern "C" {

JDB200_API size_t jdb200_slab_message_bytes(const jdb200_slab_desc* d) {
  if (slab_check(d)) return 0;
  SLAB_DISPATCH((SlabMsg<F, D>::bytes(d->cap_mig, d->cap_ghost)))
}

JDB200_API size_t jdb200_slab_kept_bytes(const jdb200_slab_desc* d) {
  if (slab_check(d)) return 0;
  SLAB_DISPATCH((SlabMsg<F, D>::bytes(0, 2 * d->cap_mig)))
}

JDB200_API size_t jdb200_slab_scratch_bytes(const jdb200_slab_desc* d) {
  if (slab_check(d)) return 0;
  const size_t nb = (size_t)std::max(1, cdiv(d->n, kSlabBlock));
  return ((nb * kSlabBlock + 255) & ~size_t(255)) + nb * 8 * sizeof(int) + 256;
}

JDB200_API size_t jdb200_slab_holes_bytes(const jdb200_slab_desc* d) {
  if (slab_check(d)) return 0;
  return (size_t)d->cap_mig * 4 * sizeof(int) + 64;
}

JDB200_API int jdb200_slab_pack(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* rows, void* msg_lo,
                                void* msg_up, void* kept, void* holes, void* header_local, void* scratch,
                                size_t scratch_bytes) {
  int rc = slab_check(d);
  if (rc) return rc;
  if (!rows || !msg_lo || !msg_up || !kept || !holes || !header_local || !scratch || !d->anchor || !d->box_size ||
      !d->cell_size)
    return JDB200_ENULL;
  if (scratch_bytes < jdb200_slab_scratch_bytes(d)) return JDB200_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  SLAB_DISPATCH((slab_pack<F, D>(s, d, rows, msg_lo, msg_up, kept, holes, header_local, scratch)))
}

JDB200_API int jdb200_slab_unpack(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* rows,
                                  const int64_t* counts, const void* from_lo, const void* from_up, const void* kept,
                                  void* holes) {
  int rc = slab_check(d);
  if (rc) return rc;
  if (!rows || !counts || !from_lo || !from_up || !kept || !holes) return JDB200_ENULL;
  cudaStream_t s = (cudaStream_t)stream;
  SLAB_DISPATCH((slab_unpack<F, D>(s, d, rows, counts, from_lo, from_up, kept, holes)))
}

}  // extern "C"
