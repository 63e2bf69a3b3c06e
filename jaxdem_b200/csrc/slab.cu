// jaxdem_b200 — slab decomposition: classify / pack / unpack kernels of the per-step
// neighbour exchange (jaxdem_b200/slab.py).  No reference equivalent: the
// reference runs one system on one device (jaxdem/system.py:60-98).
//
// After the drift every owned particle is classified by the cell layer of its last
// coordinate — the SAME arithmetic as the collider's hash (cell_coord, cell_list.py:55-60) —
// into: stays (and lies in the halo of the lower / upper face), leaves to the lower / upper
// neighbour, or strays (moved further than the halo: reported, never lost).  Stayers are
// compacted, in index order, into the alternate row buffers; leavers are written as full
// records into the message of their direction and as ghost records into `kept` (they stay
// behind as ghosts); halo particles are written as ghost records.  All lists keep the index
// order (block counts -> one-block scan -> in-block ballot ranks): bitwise repeatable.
#include "ctx.cuh"
#include "launch.cuh"

namespace jdb {

constexpr int kSlabBlock = 256;
constexpr int kSlabLists = 5;  // stay, leave-lo, leave-up, halo-lo, halo-up

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
template <typename F, int D>
__global__ void __launch_bounds__(kSlabBlock) k_slab_classify(SlabGeom gm, const F* __restrict__ pos_c,
                                                               const F* __restrict__ anchor, const F* __restrict__ box,
                                                               const F* __restrict__ cell_size,
                                                               uint8_t* __restrict__ cat, int* __restrict__ bc) {
  pdl_prologue();
  using I = typename RT<F>::I;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c8 = 0;
  bool live = i < gm.n;
  if (live) {
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

// exclusive scan of the block counts (one block; the lists are short: n / 256 entries), totals
// into the message headers and the local header
__global__ void __launch_bounds__(1024) k_slab_scan(int nblocks, int* __restrict__ bc, long long* __restrict__ hdr_lo,
                                                    long long* __restrict__ hdr_up, long long* __restrict__ hdr_local) {
  pdl_prologue();
  __shared__ long long s_run[6];
  __shared__ int s_warp[32][6];
  if (threadIdx.x < 6) s_run[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 1024) {
    const int b = base + threadIdx.x;
    int v[6], incl[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      v[j] = b < nblocks ? bc[(size_t)b * 8 + j] : 0;
      incl[j] = v[j];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl[j], o);
        if (lane >= o) incl[j] += t;
      }
      if (lane == 31) s_warp[warp][j] = incl[j];
    }
    __syncthreads();
    long long tot[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      long long woff = 0, all = 0;
      for (int w = 0; w < 32; ++w) {
        if (w < warp) woff += s_warp[w][j];
        all += s_warp[w][j];
      }
      tot[j] = all;
      if (b < nblocks) bc[(size_t)b * 8 + j] = (int)(s_run[j] + woff + incl[j] - v[j]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < 6; ++j) s_run[j] += tot[j];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // header: [0] full records, [1] ghost records, [2] strays seen by the sender
    hdr_lo[0] = s_run[1]; hdr_lo[1] = s_run[3]; hdr_lo[2] = s_run[5];
    hdr_up[0] = s_run[2]; hdr_up[1] = s_run[4]; hdr_up[2] = s_run[5];
    hdr_local[0] = s_run[0]; hdr_local[1] = s_run[1]; hdr_local[2] = s_run[2]; hdr_local[3] = s_run[5];
  }
}

template <typename F, int D>
__global__ void __launch_bounds__(kSlabBlock) k_slab_pack(SlabGeom gm, SlabRows<F> src, SlabRows<F> dst,
                                                           const uint8_t* __restrict__ cat, const int* __restrict__ bc,
                                                           void* msg_lo, void* msg_up, void* kept) {
  pdl_prologue();
  using M = SlabMsg<F, D>;
  __shared__ int s_w[kSlabBlock / 32][kSlabLists];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < gm.n;
  const int c8 = live ? cat[i] : 0;
  const bool stay = live && !(c8 & 12);
  const bool fl[kSlabLists] = {stay, live && (c8 & 4) != 0, live && (c8 & 8) != 0, stay && (c8 & 1) != 0,
                               stay && (c8 & 2) != 0};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int rank[kSlabLists];
#pragma unroll
  for (int j = 0; j < kSlabLists; ++j) {
    const unsigned m = __ballot_sync(0xffffffffu, fl[j]);
    rank[j] = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_w[warp][j] = __popc(m);
  }
  __syncthreads();
  const int* base = bc + (size_t)blockIdx.x * 8;
#pragma unroll
  for (int j = 0; j < kSlabLists; ++j) {
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_w[w][j];
    rank[j] += woff + base[j];
  }
  if (!live) return;
  if (stay) {
    F f[M::WF];
    long long iv[3];
    write_full<F, D>(src, i, f, iv);
    read_full<F, D>(dst, rank[0], f, iv);
  }
  const M lo(msg_lo, gm.cap_m, gm.cap_g), up(msg_up, gm.cap_m, gm.cap_g);
  // kept: ghost records of the leavers, lower direction in rows [0, cap_m), upper in [cap_m, 2 cap_m)
  const M kp(kept, 0, 2 * gm.cap_m);
  if (fl[1] && rank[1] < gm.cap_m) {
    write_full<F, D>(src, i, lo.mig_f + (size_t)rank[1] * M::WF, lo.mig_i + (size_t)rank[1] * 3);
    write_ghost<F, D>(src, i, kp.gh_f + (size_t)rank[1] * M::WG, kp.gh_i + (size_t)rank[1] * 2);
  }
  if (fl[2] && rank[2] < gm.cap_m) {
    write_full<F, D>(src, i, up.mig_f + (size_t)rank[2] * M::WF, up.mig_i + (size_t)rank[2] * 3);
    write_ghost<F, D>(src, i, kp.gh_f + (size_t)(gm.cap_m + rank[2]) * M::WG, kp.gh_i + (size_t)(gm.cap_m + rank[2]) * 2);
  }
  if (fl[3] && rank[3] < gm.cap_g)
    write_ghost<F, D>(src, i, lo.gh_f + (size_t)rank[3] * M::WG, lo.gh_i + (size_t)rank[3] * 2);
  if (fl[4] && rank[4] < gm.cap_g)
    write_ghost<F, D>(src, i, up.gh_f + (size_t)rank[4] * M::WG, up.gh_i + (size_t)rank[4] * 2);
}

// rows behind the stayers: arrivals from the lower, then the upper neighbour (owned), then the
// ghosts: leavers kept behind (lower, upper), halo of the lower, halo of the upper neighbour
struct SlabCounts {
  long long n_stay, a_lo, a_up, k_lo, k_up, g_lo, g_up;
};

template <typename F, int D>
__global__ void __launch_bounds__(kSlabBlock) k_slab_unpack(SlabGeom gm, SlabRows<F> dst, SlabCounts cn,
                                                             const void* from_lo, const void* from_up, const void* kept) {
  pdl_prologue();
  using M = SlabMsg<F, D>;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const M lo((void*)from_lo, gm.cap_m, gm.cap_g), up((void*)from_up, gm.cap_m, gm.cap_g), kp((void*)kept, 0, 2 * gm.cap_m);
  long long row = cn.n_stay + t;
  if (t < cn.a_lo) { read_full<F, D>(dst, row, lo.mig_f + (size_t)t * M::WF, lo.mig_i + (size_t)t * 3); return; }
  t -= cn.a_lo;
  if (t < cn.a_up) { read_full<F, D>(dst, row, up.mig_f + (size_t)t * M::WF, up.mig_i + (size_t)t * 3); return; }
  t -= cn.a_up;
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
int slab_pack(cudaStream_t s, const jdb200_slab_desc* d, const jdb200_slab_rows* src, const jdb200_slab_rows* dst,
              void* msg_lo, void* msg_up, void* kept, void* header_local, void* scratch) {
  using M = SlabMsg<F, D>;
  const SlabGeom gm = slab_geom(d);
  const int nb = std::max(1, cdiv(d->n, kSlabBlock));
  uint8_t* cat = (uint8_t*)scratch;
  int* bc = (int*)((char*)scratch + (((size_t)nb * kSlabBlock + 255) & ~size_t(255)));
  JDB_LAUNCH((k_slab_classify<F, D>), dim3(nb), kSlabBlock, s, gm, (const F*)src->pos_c, (const F*)d->anchor,
             (const F*)d->box_size, (const F*)d->cell_size, cat, bc);
  const M lo(msg_lo, gm.cap_m, gm.cap_g), up(msg_up, gm.cap_m, gm.cap_g);
  JDB_LAUNCH(k_slab_scan, dim3(1), 1024, s, nb, bc, lo.header, up.header, (long long*)header_local);
  JDB_LAUNCH((k_slab_pack<F, D>), dim3(nb), kSlabBlock, s, gm, slab_rows<F>(src), slab_rows<F>(dst), cat, bc, msg_lo,
             msg_up, kept);
  return 0;
}

template <typename F, int D>
int slab_unpack(cudaStream_t s, const jdb200_slab_desc* d, const jdb200_slab_rows* dst, const int64_t* counts,
                const void* from_lo, const void* from_up, const void* kept) {
  const SlabCounts cn{counts[0], counts[1], counts[2], counts[3], counts[4], counts[5], counts[6]};
  const long long tot = cn.a_lo + cn.a_up + cn.k_lo + cn.k_up + cn.g_lo + cn.g_up;
  if (tot == 0) return 0;
  JDB_LAUNCH((k_slab_unpack<F, D>), dim3(cdiv(tot, kSlabBlock)), kSlabBlock, s, slab_geom(d), slab_rows<F>(dst), cn,
             from_lo, from_up, kept);
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

JDB200_API int jdb200_slab_pack(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* src,
                                const jdb200_slab_rows* dst, void* msg_lo, void* msg_up, void* kept,
                                void* header_local, void* scratch, size_t scratch_bytes) {
  int rc = slab_check(d);
  if (rc) return rc;
  if (!src || !dst || !msg_lo || !msg_up || !kept || !header_local || !scratch || !d->anchor || !d->box_size ||
      !d->cell_size)
    return JDB200_ENULL;
  if (scratch_bytes < jdb200_slab_scratch_bytes(d)) return JDB200_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  SLAB_DISPATCH((slab_pack<F, D>(s, d, src, dst, msg_lo, msg_up, kept, header_local, scratch)))
}

JDB200_API int jdb200_slab_unpack(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* dst,
                                  const int64_t* counts, const void* from_lo, const void* from_up, const void* kept) {
  int rc = slab_check(d);
  if (rc) return rc;
  if (!dst || !counts || !from_lo || !from_up || !kept) return JDB200_ENULL;
  cudaStream_t s = (cudaStream_t)stream;
  SLAB_DISPATCH((slab_unpack<F, D>(s, d, dst, counts, from_lo, from_up, kept)))
}

}  // extern "C"
