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

// records are FIELD-MAJOR inside a section: word c of record r lives at base[c * cap + r], so the
// lanes of a warp (consecutive ranks r) store consecutive addresses — over NVLink that is one
// 128-byte packet per warp store instead of 32 packets of 4 bytes
template <typename F, int D>
__device__ __forceinline__ void write_full(const SlabRows<F>& s, long long i, F* f, long long* iv, size_t cap, size_t r) {
  constexpr int A = D == 3 ? 3 : 1;
  int c = 0;
  f += r;
  iv += r;
#pragma unroll
  for (int d = 0; d < D; ++d) f[cap * c++] = s.pos_c[i * D + d];
#pragma unroll
  for (int d = 0; d < D; ++d) f[cap * c++] = s.vel[i * D + d];
#pragma unroll
  for (int d = 0; d < D; ++d) f[cap * c++] = s.force[i * D + d];
#pragma unroll
  for (int a = 0; a < A; ++a) f[cap * c++] = s.ang_vel[i * A + a];
#pragma unroll
  for (int a = 0; a < A; ++a) f[cap * c++] = s.torque[i * A + a];
#pragma unroll
  for (int a = 0; a < A; ++a) f[cap * c++] = s.inertia[i * A + a];
  f[cap * c++] = s.q_w[i];
#pragma unroll
  for (int a = 0; a < 3; ++a) f[cap * c++] = s.q_xyz[i * 3 + a];
  f[cap * c++] = s.rad[i];
  f[cap * c++] = s.mass[i];
  iv[0] = s.gid[i];
  iv[cap] = (long long)s.mat_id[i];
  iv[2 * cap] = (long long)s.fixed[i];
}
template <typename F, int D>
__device__ __forceinline__ void read_full(const SlabRows<F>& s, long long i, const F* f, const long long* iv, size_t cap,
                                          size_t r) {
  constexpr int A = D == 3 ? 3 : 1;
  int c = 0;
  f += r;
  iv += r;
#pragma unroll
  for (int d = 0; d < D; ++d) s.pos_c[i * D + d] = f[cap * c++];
#pragma unroll
  for (int d = 0; d < D; ++d) s.vel[i * D + d] = f[cap * c++];
#pragma unroll
  for (int d = 0; d < D; ++d) s.force[i * D + d] = f[cap * c++];
#pragma unroll
  for (int a = 0; a < A; ++a) s.ang_vel[i * A + a] = f[cap * c++];
#pragma unroll
  for (int a = 0; a < A; ++a) s.torque[i * A + a] = f[cap * c++];
#pragma unroll
  for (int a = 0; a < A; ++a) s.inertia[i * A + a] = f[cap * c++];
  s.q_w[i] = f[cap * c++];
#pragma unroll
  for (int a = 0; a < 3; ++a) s.q_xyz[i * 3 + a] = f[cap * c++];
  s.rad[i] = f[cap * c++];
  s.mass[i] = f[cap * c++];
  s.gid[i] = iv[0];
  s.mat_id[i] = (typename RT<F>::I)iv[cap];
  s.fixed[i] = (uint8_t)iv[2 * cap];
}
template <typename F, int D>
__device__ __forceinline__ void write_ghost(const SlabRows<F>& s, long long i, F* f, long long* iv, size_t cap, size_t r) {
  constexpr int A = D == 3 ? 3 : 1;
  int c = 0;
  f += r;
  iv += r;
#pragma unroll
  for (int d = 0; d < D; ++d) f[cap * c++] = s.pos_c[i * D + d];
#pragma unroll
  for (int d = 0; d < D; ++d) f[cap * c++] = s.vel[i * D + d];
#pragma unroll
  for (int a = 0; a < A; ++a) f[cap * c++] = s.ang_vel[i * A + a];
  f[cap * c++] = s.rad[i];
  f[cap * c++] = s.mass[i];
  iv[0] = s.gid[i];
  iv[cap] = (long long)s.mat_id[i];
}
// a ghost row: what the force laws read; the rest of the row is set to inert values
template <typename F, int D>
__device__ __forceinline__ void read_ghost(const SlabRows<F>& s, long long i, const F* f, const long long* iv, size_t cap,
                                           size_t r) {
  constexpr int A = D == 3 ? 3 : 1;
  int c = 0;
  f += r;
  iv += r;
#pragma unroll
  for (int d = 0; d < D; ++d) s.pos_c[i * D + d] = f[cap * c++];
#pragma unroll
  for (int d = 0; d < D; ++d) s.vel[i * D + d] = f[cap * c++];
#pragma unroll
  for (int a = 0; a < A; ++a) s.ang_vel[i * A + a] = f[cap * c++];
  s.rad[i] = f[cap * c++];
  s.mass[i] = f[cap * c++];
  s.gid[i] = iv[0];
  s.mat_id[i] = (typename RT<F>::I)iv[cap];
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

// ---- device-side exchange state (jdb200_slab_pack_dev / _unpack_dev): int64 words -------------
// The row counts of a slab change every step and are known on the device only.  With this block
// the whole exchange runs without the host: kernels are launched over a BOUND (desc.n) and read the
// live counts here; the two neighbours signal "message complete" through a flag word in the
// message header (peer-memory store with release semantics), not through a host-visible barrier.
enum : int {
  kDevOwn = 0,     // owned rows
  kDevLocal = 1,   // owned + ghost rows (jdb200_state.n_rows of the force evaluation points here)
  kDevSeq = 2,     // exchanges completed; parity of the message buffers = seq & 1
  kDevStatus = 3,  // sticky bits, JDB200_SLAB_* (include/jaxdem_b200.h)
  kDevCounts = 4,  // SlabCounts of the current exchange (7 words)
  kDevMaxMig = 11, // largest migrant / ghost counts seen (capacity tuning)
  kDevMaxGhost = 12,
  kDevTicket = 13, // block ticket of the pack kernel
  kDevWords = 16
};
constexpr int kHdrFlag = 7;  // header word 7: seq + 1 of the exchange whose message is complete

struct SlabGeom {
  long long n;  // owned rows — or, with n_dev, the launch bound
  int n_layers, lo, up, R;
  long long cap_m, cap_g;
  long long* dev;  // device-side exchange state or NULL (counts come from the host)
};
__device__ __forceinline__ long long slab_rows_live(const SlabGeom& gm) {
  if (!gm.dev) return gm.n;
  const long long n = gm.dev[kDevOwn];
  return n < gm.n ? n : gm.n;
}

// category bits: 1 halo-lo, 2 halo-up, 4 leave-lo, 8 leave-up, 16 stray
// INTEGRATE: VelocityVerlet.step_before_force (velocity_verlet.py:57-61; the arithmetic of
// k_linear) rides along, so the drift and the classification read the positions once.
template <typename F, int D, bool INTEGRATE>
__global__ void __launch_bounds__(kSlabBlock) k_slab_classify(SlabGeom gm, SlabRows<F> rows, const F* __restrict__ dtp,
                                                               const F* __restrict__ anchor, const F* __restrict__ box,
                                                               const F* __restrict__ cell_size,
                                                               uint8_t* __restrict__ cat, int* __restrict__ bc,
                                                               int* __restrict__ sup) {
  pdl_prologue();
  using I = typename RT<F>::I;
  using T = RT<F>;
  const long long n = slab_rows_live(gm);
  if ((long long)blockIdx.x * blockDim.x >= n && blockIdx.x > 0) return;  // past the live rows (block 0 always counts)
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c8 = 0;
  bool live = i < n;
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
    // sums over groups of 32 blocks (integer atomics: exact, order-free), so the scan below is one
    // short parallel pass; zeroed again by k_slab_pack
    int* sp = sup + (size_t)(blockIdx.x >> 5) * 8;
    if (n0) atomicAdd(sp + 0, n0);
    if (n1) atomicAdd(sp + 1, n1);
    if (n2) atomicAdd(sp + 2, n2);
    if (n3) atomicAdd(sp + 3, n3);
    if (n4) atomicAdd(sp + 4, n4);
    if (n5) atomicAdd(sp + 5, n5);
  }
}

// exclusive scan of the block counts -> per-block offsets of the six lists, totals -> tot[6] (stay,
// leave-lo, leave-up, halo-lo, halo-up, stray).  One warp per group of 32 blocks: its base is the sum of
// the earlier groups' accumulators (k_slab_classify), its own 32 counts are scanned with shuffles.
__global__ void __launch_bounds__(256) k_slab_scan(SlabGeom gm, int* __restrict__ bc, const int* __restrict__ sup,
                                                   long long* __restrict__ tot) {
  pdl_prologue();
  const long long n = slab_rows_live(gm);
  const int nblocks = (int)max((n + kSlabBlock - 1) / kSlabBlock, 1LL);
  const int nsup = (nblocks + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int sg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sg >= nsup) return;
  int pre[6] = {0, 0, 0, 0, 0, 0};
  for (int s2 = lane; s2 < sg; s2 += 32) {
    const int4 a = *reinterpret_cast<const int4*>(sup + (size_t)s2 * 8);
    const int2 h = *reinterpret_cast<const int2*>(sup + (size_t)s2 * 8 + 4);
    pre[0] += a.x; pre[1] += a.y; pre[2] += a.z; pre[3] += a.w; pre[4] += h.x; pre[5] += h.y;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre[j] += __shfl_xor_sync(0xffffffffu, pre[j], o);
  const int b = sg * 32 + lane;
  int v[6] = {0, 0, 0, 0, 0, 0};
  if (b < nblocks) {
    const int4 a = *reinterpret_cast<const int4*>(bc + (size_t)b * 8);
    const int2 h = *reinterpret_cast<const int2*>(bc + (size_t)b * 8 + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = h.x; v[5] = h.y;
  }
  int incl[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    incl[j] = v[j];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl[j], o);
      if (lane >= o) incl[j] += t;
    }
  }
  if (b < nblocks) {
    *reinterpret_cast<int4*>(bc + (size_t)b * 8) =
        make_int4(pre[0] + incl[0] - v[0], pre[1] + incl[1] - v[1], pre[2] + incl[2] - v[2], pre[3] + incl[3] - v[3]);
    *reinterpret_cast<int2*>(bc + (size_t)b * 8 + 4) = make_int2(pre[4] + incl[4] - v[4], pre[5] + incl[5] - v[5]);
  }
  if (sg == nsup - 1 && lane == 31) {
#pragma unroll
    for (int j = 0; j < 6; ++j) tot[j] = pre[j] + incl[j];
  }
}

__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

constexpr int kPackLists = 4;  // leave-lo, leave-up, halo-lo, halo-up (block-count columns 1..4)

// the two messages of one exchange: [parity] — the host protocol passes the same pointer twice
struct SlabPorts {
  void* lo[2];
  void* up[2];
};

// Row indices of the four lists, in index order: one WARP per 256-row chunk of k_slab_classify (lane l
// owns rows 8 l .. 8 l + 7 of the chunk: one 8-byte load of the category bytes), offsets from the scanned
// block counts.  The leavers' lists are `holes` itself ([0, cap_m) downwards, [cap_m, 2 cap_m) upwards),
// the halo lists go to halo[0, cap_g) and halo[cap_g, 2 cap_g).  No particle row is touched here, so the
// pass is short whatever the particle order is (the rows of a face are scattered over all chunks when
// the particle index carries no spatial order).
__global__ void __launch_bounds__(256) k_slab_compact(SlabGeom gm, const uint8_t* __restrict__ cat,
                                                      const int* __restrict__ bc, int* __restrict__ holes,
                                                      int* __restrict__ halo, int* __restrict__ sup, int nsup) {
  pdl_prologue();
  if (blockIdx.x == 0)  // the group accumulators of k_slab_classify: zero for the next exchange
    for (int t = threadIdx.x; t < nsup * 8; t += blockDim.x) sup[t] = 0;
  const long long n = slab_rows_live(gm);
  const int lane = threadIdx.x & 31;
  const long long chunk = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long base = chunk * kSlabBlock + lane * 8;
  if (chunk * kSlabBlock >= n) return;
  unsigned long long w = 0;
  if (base + 8 <= n) {
    w = *reinterpret_cast<const unsigned long long*>(cat + base);
  } else {
    for (int k = 0; k < 8; ++k)
      if (base + k < n) w |= (unsigned long long)cat[base + k] << (8 * k);
  }
  // per-byte masks: bit k of m[j] <=> row base + k is in list j
  unsigned m[kPackLists] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const unsigned c8 = (unsigned)(w >> (8 * k)) & 0xffu;
    const bool stay = !(c8 & 12u);
    m[0] |= ((c8 >> 2) & 1u) << k;
    m[1] |= ((c8 >> 3) & 1u) << k;
    m[2] |= (stay ? (c8 & 1u) : 0u) << k;
    m[3] |= (stay ? ((c8 >> 1) & 1u) : 0u) << k;
  }
  if (!__any_sync(0xffffffffu, (m[0] | m[1] | m[2] | m[3]) != 0)) return;
  const int4 off = *reinterpret_cast<const int4*>(bc + (size_t)chunk * 8);  // columns 0..3
  const int off4 = bc[(size_t)chunk * 8 + 4];
  const int start[kPackLists] = {off.y, off.z, off.w, off4};
  int* const dst[kPackLists] = {holes, holes + gm.cap_m, halo, halo + gm.cap_g};
  const long long cap[kPackLists] = {gm.cap_m, gm.cap_m, gm.cap_g, gm.cap_g};
#pragma unroll
  for (int j = 0; j < kPackLists; ++j) {
    if (!__any_sync(0xffffffffu, m[j] != 0)) continue;
    const int cnt = __popc(m[j]);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    long long r = start[j] + incl - cnt;
    unsigned mm = m[j];
    while (mm) {
      const int k = __ffs(mm) - 1;
      mm &= mm - 1;
      if (r < cap[j]) dst[j][r] = (int)(base + k);
      ++r;
    }
  }
}

// One thread per list entry: gather the row, store the record — consecutive threads store consecutive
// records of a field-major section, i.e. coalesced 128-byte stores into the neighbour's memory.  The
// block that finishes last (ticket) then writes the headers — [0] full records, [1] ghost records, [2]
// strays seen by the sender — and, device protocol, the "message complete" flag behind a system-scope
// fence (every storing thread fenced its own records before it took the ticket).
template <typename F, int D>
__global__ void __launch_bounds__(kSlabBlock) k_slab_pack(SlabGeom gm, SlabRows<F> src, const long long* __restrict__ tot,
                                                           SlabPorts out, void* kept, const int* __restrict__ holes,
                                                           const int* __restrict__ halo, long long* __restrict__ hdr_local,
                                                           unsigned* __restrict__ ticket) {
  pdl_prologue();
  using M = SlabMsg<F, D>;
  __shared__ bool s_last;
  const long long seq = gm.dev ? gm.dev[kDevSeq] : 0;
  const int par = (int)(seq & 1);
  const M lo(out.lo[par], gm.cap_m, gm.cap_g), up(out.up[par], gm.cap_m, gm.cap_g), kp(kept, 0, 2 * gm.cap_m);
  const size_t cm = (size_t)gm.cap_m, cg = (size_t)gm.cap_g;
  const long long c1 = min(tot[1], gm.cap_m), c2 = min(tot[2], gm.cap_m), c3 = min(tot[3], gm.cap_g),
                  c4 = min(tot[4], gm.cap_g);
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool wrote = true;
  if (t < c1) {
    const long long i = holes[t];
    write_full<F, D>(src, i, lo.mig_f, lo.mig_i, cm, t);
    write_ghost<F, D>(src, i, kp.gh_f, kp.gh_i, 2 * cm, t);
  } else if ((t -= c1) < c2) {
    const long long i = holes[cm + t];
    write_full<F, D>(src, i, up.mig_f, up.mig_i, cm, t);
    write_ghost<F, D>(src, i, kp.gh_f, kp.gh_i, 2 * cm, cm + t);
  } else if ((t -= c2) < c3) {
    write_ghost<F, D>(src, halo[t], lo.gh_f, lo.gh_i, cg, t);
  } else if ((t -= c3) < c4) {
    write_ghost<F, D>(src, halo[cg + t], up.gh_f, up.gh_i, cg, t);
  } else {
    wrote = false;
  }
  if (wrote) __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  *ticket = 0;
  lo.header[0] = tot[1]; lo.header[1] = tot[3]; lo.header[2] = tot[5];
  up.header[0] = tot[2]; up.header[1] = tot[4]; up.header[2] = tot[5];
  hdr_local[0] = tot[0]; hdr_local[1] = tot[1]; hdr_local[2] = tot[2]; hdr_local[3] = tot[5];
  if (gm.dev) {
    if (tot[1] > gm.cap_m || tot[2] > gm.cap_m || tot[3] > gm.cap_g || tot[4] > gm.cap_g)
      gm.dev[kDevStatus] |= JDB200_SLAB_MESSAGE_FULL;  // the receiver sees the same through the counts
    __threadfence_system();  // orders the counts above and (cumulatively) every block's fenced records
    *reinterpret_cast<volatile long long*>(lo.header + kHdrFlag) = seq + 1;
    *reinterpret_cast<volatile long long*>(up.header + kHdrFlag) = seq + 1;
  }
}

// counts of one exchange: owned rows before, arrivals from the lower / upper neighbour, leavers to
// the lower / upper neighbour, halo of the lower / upper neighbour.  Host protocol: kernel
// arguments; device protocol: dev[kDevCounts ...], written by k_slab_repair
struct SlabCounts {
  long long n_old, a_lo, a_up, k_lo, k_up, g_lo, g_up;
};
__device__ __forceinline__ SlabCounts slab_counts(const SlabGeom& gm, const SlabCounts& host) {
  if (!gm.dev) return host;
  const long long* c = gm.dev + kDevCounts;
  return SlabCounts{c[0], c[1], c[2], c[3], c[4], c[5], c[6]};
}

template <typename F, int D>
__device__ __forceinline__ void copy_row(const SlabRows<F>& s, long long from, long long to) {
  F f[SlabMsg<F, D>::WF];
  long long iv[3];
  write_full<F, D>(s, from, f, iv, 1, 0);
  read_full<F, D>(s, to, f, iv, 1, 0);
}

// One block repairs the owned rows after an exchange:
//  (device protocol) wait for both neighbours' "message complete" flags, read the counts from the
//  headers, check them against the capacities, publish the new row counts;
//  merge the two sorted hole lists holes[0, k_lo) and holes[cap_m, cap_m + k_up) -> holes[2 cap_m ...);
//  more leavers than arrivals: the holes the arrivals do not fill are R = sorted_holes[a, l), the new
//  row count is n' = n_old - (l - a); rows of the tail [n', n_old) that are not holes move, in index
//  order, into the holes below n' (a prefix of R).
template <typename F, int D>
__global__ void __launch_bounds__(1024) k_slab_repair(SlabGeom gm, SlabRows<F> rows, SlabCounts host, SlabPorts in,
                                                      const long long* __restrict__ hdr_local, int* __restrict__ holes,
                                                      long long timeout_ns) {
  pdl_prologue();
  __shared__ int s_warp[32];
  __shared__ int s_base;
  __shared__ SlabCounts s_cn;
  if (threadIdx.x == 0) {
    s_base = 0;
    if (!gm.dev) {
      s_cn = host;
    } else {
      long long* dev = gm.dev;
      const long long seq = dev[kDevSeq];
      const int par = (int)(seq & 1);
      const long long* hlo = (const long long*)in.lo[par];
      const long long* hup = (const long long*)in.up[par];
      long long status = 0;
      const unsigned long long t0 = global_timer_ns();
      unsigned spins = 0;
      while (ld_acquire_sys(hlo + kHdrFlag) != seq + 1 || ld_acquire_sys(hup + kHdrFlag) != seq + 1) {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > (unsigned long long)timeout_ns) {
          status |= JDB200_SLAB_TIMEOUT;
          break;
        }
      }
      SlabCounts cn;
      cn.n_old = dev[kDevOwn];
      cn.k_lo = hdr_local[1]; cn.k_up = hdr_local[2];
      cn.a_lo = hlo[0]; cn.g_lo = hlo[1];
      cn.a_up = hup[0]; cn.g_up = hup[1];
      if (hdr_local[3] || hlo[2] || hup[2]) status |= JDB200_SLAB_STRAY;
      const long long mig = max(max(cn.k_lo, cn.k_up), max(cn.a_lo, cn.a_up)), gh = max(cn.g_lo, cn.g_up);
      if (mig > gm.cap_m || gh > gm.cap_g) status |= JDB200_SLAB_MESSAGE_FULL;
      long long n_new = cn.n_old - cn.k_lo - cn.k_up + cn.a_lo + cn.a_up;
      long long n_gh = cn.k_lo + cn.k_up + cn.g_lo + cn.g_up;
      if (max(n_new, cn.n_old) + n_gh > gm.n) status |= JDB200_SLAB_ROWS_FULL;
      if (status & (JDB200_SLAB_TIMEOUT | JDB200_SLAB_MESSAGE_FULL | JDB200_SLAB_ROWS_FULL)) {
        // nothing can be placed safely: leave the rows as they are (the host raises at its next look)
        cn.a_lo = cn.a_up = cn.k_lo = cn.k_up = cn.g_lo = cn.g_up = 0;
        n_new = cn.n_old;
        n_gh = 0;
      }
      dev[kDevMaxMig] = max(dev[kDevMaxMig], mig);
      dev[kDevMaxGhost] = max(dev[kDevMaxGhost], gh);
      dev[kDevStatus] |= status;
      long long* c = dev + kDevCounts;
      c[0] = cn.n_old; c[1] = cn.a_lo; c[2] = cn.a_up; c[3] = cn.k_lo; c[4] = cn.k_up; c[5] = cn.g_lo; c[6] = cn.g_up;
      dev[kDevOwn] = n_new;
      dev[kDevLocal] = n_new + n_gh;
      dev[kDevSeq] = seq + 1;
      s_cn = cn;
    }
  }
  __syncthreads();
  const SlabCounts cn = s_cn;
  const long long l = cn.k_lo + cn.k_up, a = cn.a_lo + cn.a_up;
  // ---- merge the hole lists (indices are distinct) ----
  for (long long t = threadIdx.x; t < l; t += blockDim.x) {
    const bool first = t < cn.k_lo;
    const int* mine = first ? holes : holes + gm.cap_m;
    const int* other = first ? holes + gm.cap_m : holes;
    const long long r = first ? t : t - cn.k_lo, no = first ? cn.k_up : cn.k_lo;
    const int v = mine[r];
    long long lo = 0, hi = no;  // number of entries of the other list below v
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (other[mid] < v) lo = mid + 1; else hi = mid;
    }
    holes[2 * gm.cap_m + r + lo] = v;
  }
  if (l <= a) return;
  __syncthreads();
  // ---- tail move ----
  const long long nr = l - a, n_new = cn.n_old - nr;
  const int* R = holes + 2 * gm.cap_m + a;
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
__global__ void __launch_bounds__(kSlabBlock) k_slab_unpack(SlabGeom gm, SlabRows<F> dst, SlabCounts host, SlabPorts in,
                                                             const void* kept, const int* __restrict__ holes) {
  pdl_prologue();
  using M = SlabMsg<F, D>;
  const SlabCounts cn = slab_counts(gm, host);
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long l = cn.k_lo + cn.k_up, a = cn.a_lo + cn.a_up;
  if (t >= a + l + cn.g_lo + cn.g_up) return;
  const int par = gm.dev ? (int)((gm.dev[kDevSeq] - 1) & 1) : 0;  // k_slab_repair has advanced seq already
  const M lo(in.lo[par], gm.cap_m, gm.cap_g), up(in.up[par], gm.cap_m, gm.cap_g), kp((void*)kept, 0, 2 * gm.cap_m);
  const size_t cm = (size_t)gm.cap_m, cg = (size_t)gm.cap_g;
  if (t < a) {
    const long long row = t < l ? (long long)holes[2 * gm.cap_m + t] : cn.n_old + (t - l);
    if (t < cn.a_lo) read_full<F, D>(dst, row, lo.mig_f, lo.mig_i, cm, t);
    else read_full<F, D>(dst, row, up.mig_f, up.mig_i, cm, t - cn.a_lo);
    return;
  }
  t -= a;
  const long long row = cn.n_old - l + a + t;
  if (t < cn.k_lo) { read_ghost<F, D>(dst, row, kp.gh_f, kp.gh_i, 2 * cm, t); return; }
  t -= cn.k_lo;
  if (t < cn.k_up) { read_ghost<F, D>(dst, row, kp.gh_f, kp.gh_i, 2 * cm, cm + t); return; }
  t -= cn.k_up;
  if (t < cn.g_lo) { read_ghost<F, D>(dst, row, lo.gh_f, lo.gh_i, cg, t); return; }
  t -= cn.g_lo;
  if (t < cn.g_up) read_ghost<F, D>(dst, row, up.gh_f, up.gh_i, cg, t);
}

static inline int slab_check(const jdb200_slab_desc* d) {
  if (!d) return JDB200_ENULL;
  if (d->dim != 2 && d->dim != 3) return JDB200_EINVAL;
  if (d->dtype != JDB200_F32 && d->dtype != JDB200_F64) return JDB200_EINVAL;
  if (d->n < 0 || d->cap_mig < 0 || d->cap_ghost < 0 || d->n_layers < 1 || d->search_range < 1) return JDB200_EINVAL;
  if (d->lo_layer < 0 || d->up_layer > d->n_layers || d->lo_layer >= d->up_layer) return JDB200_EINVAL;
  return 0;
}
static inline SlabGeom slab_geom(const jdb200_slab_desc* d, void* dev) {
  return SlabGeom{d->n, d->n_layers, d->lo_layer, d->up_layer, d->search_range, d->cap_mig, d->cap_ghost, (long long*)dev};
}

// scratch layout (jdb200_slab_scratch_bytes): category bytes | block counts | totals | group accumulators |
// ticket | halo row lists.  The accumulators and the ticket must be zero before the first call and are
// left zero by every call.
struct SlabScratch {
  uint8_t* cat;
  int* bc;
  long long* tot;
  int* sup;
  unsigned* ticket;
  int* halo;
  int nb, nsup;
  size_t bytes;
  SlabScratch(void* base, long long n, long long cap_g) {
    nb = std::max(1, cdiv(n, kSlabBlock));
    nsup = cdiv(nb, 32);
    size_t o = 0;
    cat = (uint8_t*)base;
    o = ((size_t)nb * kSlabBlock + 255) & ~size_t(255);
    bc = (int*)((char*)base + o);
    o += (size_t)nb * 8 * sizeof(int);
    tot = (long long*)((char*)base + o);
    o += 64;
    sup = (int*)((char*)base + o);
    o += (size_t)nsup * 8 * sizeof(int);
    ticket = (unsigned*)((char*)base + o);
    o += 64;
    halo = (int*)((char*)base + o);
    o += (size_t)2 * cap_g * sizeof(int);
    bytes = o + 256;
  }
};

template <typename F, int D>
int slab_pack(cudaStream_t s, const jdb200_slab_desc* d, const jdb200_slab_rows* rows, SlabPorts out, void* kept,
              void* holes, void* header_local, void* scratch, void* dev) {
  const SlabGeom gm = slab_geom(d, dev);
  const SlabScratch sc(scratch, d->n, d->cap_ghost);
  if (d->dt)
    JDB_LAUNCH((k_slab_classify<F, D, true>), dim3(sc.nb), kSlabBlock, s, gm, slab_rows<F>(rows), (const F*)d->dt,
               (const F*)d->anchor, (const F*)d->box_size, (const F*)d->cell_size, sc.cat, sc.bc, sc.sup);
  else
    JDB_LAUNCH((k_slab_classify<F, D, false>), dim3(sc.nb), kSlabBlock, s, gm, slab_rows<F>(rows), (const F*)nullptr,
               (const F*)d->anchor, (const F*)d->box_size, (const F*)d->cell_size, sc.cat, sc.bc, sc.sup);
  JDB_LAUNCH(k_slab_scan, dim3(cdiv(sc.nsup, 8)), 256, s, gm, sc.bc, (const int*)sc.sup, sc.tot);
  JDB_LAUNCH(k_slab_compact, dim3(cdiv(sc.nb, 8)), 256, s, gm, (const uint8_t*)sc.cat, (const int*)sc.bc, (int*)holes,
             sc.halo, sc.sup, sc.nsup);
  const long long entries = 2 * d->cap_mig + 2 * d->cap_ghost;
  JDB_LAUNCH((k_slab_pack<F, D>), dim3(std::max(1, cdiv(entries, kSlabBlock))), kSlabBlock, s, gm, slab_rows<F>(rows),
             (const long long*)sc.tot, out, kept, (const int*)holes, (const int*)sc.halo, (long long*)header_local,
             sc.ticket);
  return 0;
}

template <typename F, int D>
int slab_unpack(cudaStream_t s, const jdb200_slab_desc* d, const jdb200_slab_rows* rows, const int64_t* counts,
                SlabPorts in, const void* header_local, const void* kept, void* holes, void* dev, long long timeout_ns) {
  SlabCounts cn{0, 0, 0, 0, 0, 0, 0};
  if (counts) cn = SlabCounts{counts[0], counts[1], counts[2], counts[3], counts[4], counts[5], counts[6]};
  const SlabGeom gm = slab_geom(d, dev);
  const long long l = cn.k_lo + cn.k_up;
  // host protocol: exact grids; device protocol: grids over the capacities, early exit on the live counts
  const long long tot = dev ? 4 * d->cap_mig + 2 * d->cap_ghost : cn.a_lo + cn.a_up + l + cn.g_lo + cn.g_up;
  if (dev || l > 0)
    JDB_LAUNCH((k_slab_repair<F, D>), dim3(1), 1024, s, gm, slab_rows<F>(rows), cn, in, (const long long*)header_local,
               (int*)holes, timeout_ns);
  if (tot > 0)
    JDB_LAUNCH((k_slab_unpack<F, D>), dim3(cdiv(tot, kSlabBlock)), kSlabBlock, s, gm, slab_rows<F>(rows), cn, in, kept,
               (const int*)holes);
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
  return SlabScratch(nullptr, d->n, d->cap_ghost).bytes;
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
  const SlabPorts out{{msg_lo, msg_lo}, {msg_up, msg_up}};
  SLAB_DISPATCH((slab_pack<F, D>(s, d, rows, out, kept, holes, header_local, scratch, nullptr)))
}

JDB200_API int jdb200_slab_unpack(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* rows,
                                  const int64_t* counts, const void* from_lo, const void* from_up, const void* kept,
                                  void* holes) {
  int rc = slab_check(d);
  if (rc) return rc;
  if (!rows || !counts || !from_lo || !from_up || !kept || !holes) return JDB200_ENULL;
  cudaStream_t s = (cudaStream_t)stream;
  const SlabPorts in{{(void*)from_lo, (void*)from_lo}, {(void*)from_up, (void*)from_up}};
  SLAB_DISPATCH((slab_unpack<F, D>(s, d, rows, counts, in, nullptr, kept, holes, nullptr, 0)))
}

JDB200_API int jdb200_slab_pack_dev(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* rows,
                                    void* dev_state, void* const msg_lo[2], void* const msg_up[2], void* kept,
                                    void* holes, void* header_local, void* scratch, size_t scratch_bytes) {
  int rc = slab_check(d);
  if (rc) return rc;
  if (!rows || !dev_state || !msg_lo || !msg_up || !msg_lo[0] || !msg_lo[1] || !msg_up[0] || !msg_up[1] || !kept ||
      !holes || !header_local || !scratch || !d->anchor || !d->box_size || !d->cell_size)
    return JDB200_ENULL;
  if (scratch_bytes < jdb200_slab_scratch_bytes(d)) return JDB200_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const SlabPorts out{{msg_lo[0], msg_lo[1]}, {msg_up[0], msg_up[1]}};
  SLAB_DISPATCH((slab_pack<F, D>(s, d, rows, out, kept, holes, header_local, scratch, dev_state)))
}

JDB200_API int jdb200_slab_unpack_dev(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* rows,
                                      void* dev_state, const void* const from_lo[2], const void* const from_up[2],
                                      const void* header_local, const void* kept, void* holes, int64_t timeout_ns) {
  int rc = slab_check(d);
  if (rc) return rc;
  if (!rows || !dev_state || !from_lo || !from_up || !from_lo[0] || !from_lo[1] || !from_up[0] || !from_up[1] ||
      !header_local || !kept || !holes)
    return JDB200_ENULL;
  if (timeout_ns <= 0) return JDB200_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const SlabPorts in{{(void*)from_lo[0], (void*)from_lo[1]}, {(void*)from_up[0], (void*)from_up[1]}};
  SLAB_DISPATCH((slab_unpack<F, D>(s, d, rows, nullptr, in, header_local, kept, holes, dev_state, (long long)timeout_ns)))
}

}  // extern "C"
