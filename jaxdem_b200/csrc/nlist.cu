// jaxdem_b200 — Verlet NeighborList collider (reference jaxdem/colliders/neighbor_list.py:57-131,286-480,
// 542-669): a cached (N, K) list built by the cell-list collider with radius cutoff + skin, rebuilt when some
// particle has moved further than skin / 2 since the last build (or never built), and a list-driven force /
// energy pass.
//
// The rebuild decision never leaves the device: k_nl_check reduces max |pos - old_pos|^2 per block,
// k_nl_decide turns it into a per-system gate, and every kernel of the rebuild (partition build of celllist.cu,
// k_neighbor_list) starts with `if (!gate[b]) return`.  A step that keeps its list therefore costs the check,
// a handful of empty launches and the list walk — no sort, no stencil.  Under a batch axis each system
// decides for itself (what lax.cond-turned-select computes per element in the reference).
#include "ctx.cuh"
#include "launch.cuh"
#include "laws.cuh"

namespace jdb {

template <typename F> int build_partition(cudaStream_t, Ctx<F>&, const F*, int, bool);
template <typename F> int celllist_neighbor_list(cudaStream_t, Ctx<F>&, const F*, typename RT<F>::I*, uint8_t*);

template <typename F>
__device__ __forceinline__ F nl_block_max(F v) {
  __shared__ F sm[kReduceBlock];
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int s = kReduceBlock / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] = RT<F>::fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
    __syncthreads();
  }
  return sm[0];
}
template <typename F>
__device__ __forceinline__ F nl_block_sum(F v) {  // fixed tree order => deterministic
  __shared__ F sm[kReduceBlock];
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int s = kReduceBlock / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  return sm[0];
}

// _check_and_rebuild, first half (neighbor_list.py:80-86): disp = state.pos - old_pos (NOT a periodic
// displacement), max over particles of |disp|^2.  Also leaves urec[i] = (pos, rad), the record the list walk gathers.
template <typename F, int D>
__global__ void __launch_bounds__(kReduceBlock) k_nl_check(Ctx<F> c) {
  pdl_prologue();
  using T = RT<F>;
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  F d2 = F(0);
  if (i < c.n) {
    const size_t g = (size_t)b * c.n + i;
    F p[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < D; ++d) {
      p[d] = c.pos_c[g * D + d];
      if (!(c.promises & JDB200_PROMISE_NO_POS_P)) p[d] = T::add(p[d], c.pos_p_rot[g * D + d]);
      const F dd = p[d] - c.nl_old_pos[g * D + d];
      d2 += dd * dd;
    }
    Vec4<F> r;
    r.x = p[0]; r.y = p[1]; r.z = p[2]; r.w = c.rad[g];
    c.urec[g] = r;
  }
  const F m = nl_block_max(d2);
  if (threadIdx.x == 0) c.nl_part[(size_t)b * c.reduce_blocks + blockIdx.x] = m;
}

// second half (neighbor_list.py:87-90): should_rebuild = max_disp_sq > skin^2 / 4  OR  n_build_times == 0.
// Publishes the gate, the list radius cutoff + skin (:463) and bumps n_build_times (:479).
template <typename F>
__global__ void __launch_bounds__(kReduceBlock) k_nl_decide(Ctx<F> c) {
  pdl_prologue();
  using I = typename RT<F>::I;
  const int b = blockIdx.x;
  F m = F(0);
  for (int i = threadIdx.x; i < c.reduce_blocks; i += kReduceBlock)
    m = RT<F>::fmax(m, c.nl_part[(size_t)b * c.reduce_blocks + i]);
  m = nl_block_max(m);
  if (threadIdx.x == 0) {
    const F skin = c.nl_skin[b];
    const I built = c.nl_builds[b];
    const int go = (m > skin * skin / F(4)) || built == I(0);
    c.nl_gate[b] = go;
    c.nl_cut[b] = c.nl_cutoff[b] + skin;
    if (go) c.nl_builds[b] = built + I(1);
  }
}

// rebuild branch: old_pos <- state.pos (neighbor_list.py:477)
template <typename F, int D>
__global__ void __launch_bounds__(256) k_nl_commit(Ctx<F> c, int clear_overflow) {
  pdl_prologue();
  const int b = blockIdx.y;
  if (!c.nl_gate[b]) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const size_t g = (size_t)b * c.n + i;
  const Vec4<F> r = c.urec[g];
  c.nl_old_pos[g * D + 0] = r.x;
  c.nl_old_pos[g * D + 1] = r.y;
  if (D == 3) c.nl_old_pos[g * D + 2] = r.z;
  if (clear_overflow && i == 0 && c.overflow) c.overflow[b] = 0;  // K == 0: create_neighbor_list returns (empty, False)
}

template <typename F>
__device__ __forceinline__ Body<F> nl_body(const Ctx<F>& c, size_t g, bool cs) {
  Body<F> a;
  const Vec4<F> r = c.urec[g];
  a.x = r.x; a.y = r.y; a.z = c.dim == 3 ? r.z : F(0); a.r = r.w;
  a.mat = c.nmat > 1 ? (int)c.mat_id[g] : 0;
  if (cs) {
    const F* v = c.vel + g * c.dim;
    a.vx = v[0]; a.vy = v[1]; a.vz = c.dim == 3 ? v[2] : F(0);
    a.m = c.mass[g];
    const F* w = c.ang_vel + g * c.A;
    if (c.dim == 3) { a.wx = w[0]; a.wy = w[1]; a.wz = w[2]; }
    else { a.wx = F(0); a.wy = F(0); a.wz = w[0]; }
  }
  return a;
}

// valid_interaction_mask as NeighborList calls it (neighbor_list.py:592-598): clump of i vs clump of j, and the
// OWNER's bond row against the candidate's index.
template <typename F>
__device__ __forceinline__ bool nl_valid(const Ctx<F>& c, size_t off, long long i, long long j, bool plain, bool interact) {
  if (plain) return true;  // sphere system without bonds: clump ids differ for i != j, and a list never holds i itself
  if (c.clump_id[off + i] == c.clump_id[off + j]) return false;
  if (!interact && !(c.promises & JDB200_PROMISE_NO_BONDS)) {
    const typename RT<F>::I* row = c.bond_id + (off + i) * c.W;
    for (int w = 0; w < c.W; ++w)
      if (row[w] == (typename RT<F>::I)j) return false;
  }
  return true;
}

// NeighborList.compute_force, list walk (neighbor_list.py:581-614): one thread owns particle i (original order),
// walks its row in list order — rows are packed by construction (_pack_stencil_lists, _partition.py:102-150), so
// the first -1 ends the row — and accumulates force and torque; epilogue torque += cross(_pos_p_rot_i, sum F).
template <typename F, int LAW, int D>
__global__ void __launch_bounds__(128) k_nl_force(Ctx<F> c) {
  pdl_prologue();
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const size_t off = (size_t)b * c.n;
  const LawCtx<F> lc = make_law_ctx(c, b);
  constexpr bool cs = LAW == JDB200_LAW_CUNDALLSTRACK;
  const bool interact = c.interact && c.interact[b];
  const bool plain = !c.clumps && (c.promises & JDB200_PROMISE_NO_BONDS);
  const Body<F> a = nl_body(c, off + i, cs);
  const I* row = c.nl_list + (off + i) * c.K;
  F f[3] = {0, 0, 0}, t[3] = {0, 0, 0};
  for (int q = 0; q < c.K; ++q) {
    const I j = row[q];
    if (j < I(0)) break;
    if (!nl_valid(c, off, i, (long long)j, plain, interact)) continue;
    const Body<F> bj = nl_body(c, off + (size_t)j, cs);
    F ff[3], tt[3];
    pair_force<F, LAW>(lc, a, bj, ff, tt);
    f[0] += ff[0]; f[1] += ff[1]; f[2] += ff[2];
    t[0] += tt[0]; t[1] += tt[1]; t[2] += tt[2];
  }
  const size_t g = off + i;
  F pr[3] = {0, 0, 0};
  if (!(c.promises & JDB200_PROMISE_NO_POS_P)) {
#pragma unroll
    for (int d = 0; d < D; ++d) pr[d] = c.pos_p_rot[g * D + d];
  }
#pragma unroll
  for (int d = 0; d < D; ++d) c.force[g * D + d] = f[d];
  if (D == 3) {
    c.torque[g * 3 + 0] = t[0] + (pr[1] * f[2] - pr[2] * f[1]);
    c.torque[g * 3 + 1] = t[1] + (pr[2] * f[0] - pr[0] * f[2]);
    c.torque[g * 3 + 2] = t[2] + (pr[0] * f[1] - pr[1] * f[0]);
  } else {
    c.torque[g] = t[2] + (pr[0] * f[1] - pr[1] * f[0]);
  }
}

// NeighborList.compute_potential_energy (neighbor_list.py:700-727): 0.5 * sum_j E_ij per particle, summed.
template <typename F, int LAW>
__global__ void __launch_bounds__(kReduceBlock) k_nl_energy(Ctx<F> c) {
  pdl_prologue();
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t off = (size_t)b * c.n;
  F e = F(0);
  if (i < c.n) {
    const LawCtx<F> lc = make_law_ctx(c, b);
    const bool interact = c.interact && c.interact[b];
    const bool plain = !c.clumps && (c.promises & JDB200_PROMISE_NO_BONDS);
    const Body<F> a = nl_body(c, off + i, false);
    const I* row = c.nl_list + (off + i) * c.K;
    F acc = F(0);
    for (int q = 0; q < c.K; ++q) {
      const I j = row[q];
      if (j < I(0)) break;
      if (!nl_valid(c, off, i, (long long)j, plain, interact)) continue;
      acc += pair_energy<F, LAW>(lc, a, nl_body(c, off + (size_t)j, false));
    }
    e = F(0.5) * acc;
  }
  const F tot = nl_block_sum(e);
  if (threadIdx.x == 0) c.partial[(size_t)b * c.reduce_blocks + blockIdx.x] = tot;
}

template <typename F>
__global__ void __launch_bounds__(kReduceBlock) k_nl_final_sum(const F* __restrict__ partial, int nblocks,
                                                                F* __restrict__ out) {
  pdl_prologue();
  const int b = blockIdx.x;
  F acc = F(0);
  for (int i = threadIdx.x; i < nblocks; i += kReduceBlock) acc += partial[(size_t)b * nblocks + i];
  const F tot = nl_block_sum(acc);
  if (threadIdx.x == 0) out[b] = tot;
}

#define JDB_LAW_SWITCH(law, CALL)                                     \
  switch (law) {                                                      \
    case JDB200_LAW_SPRING: { constexpr int L = JDB200_LAW_SPRING; CALL; } break;           \
    case JDB200_LAW_HERTZ: { constexpr int L = JDB200_LAW_HERTZ; CALL; } break;             \
    default: { constexpr int L = JDB200_LAW_CUNDALLSTRACK; CALL; } break;                   \
  }

// _check_and_rebuild (neighbor_list.py:57-131) without history (no history-tracking law on this path).
template <typename F>
int neighborlist_refresh(cudaStream_t s, Ctx<F>& c) {
  if (c.n == 0) return 0;
  const dim3 gr(c.reduce_blocks, c.batch);
  if (c.dim == 3) JDB_LAUNCH((k_nl_check<F, 3>), gr, kReduceBlock, s, c);
  else JDB_LAUNCH((k_nl_check<F, 2>), gr, kReduceBlock, s, c);
  JDB_LAUNCH(k_nl_decide<F>, dim3(c.batch), kReduceBlock, s, c);
  c.gate = c.nl_gate;
  int rc = 0;
  if (c.K > 0) rc = celllist_neighbor_list<F>(s, c, c.nl_cut, c.nl_list, c.overflow);  // every kernel of it honours c.gate
  if (rc) return rc;
  const dim3 gc(cdiv(c.n, 256), c.batch);
  if (c.dim == 3) JDB_LAUNCH((k_nl_commit<F, 3>), gc, 256, s, c, c.K == 0 ? 1 : 0);
  else JDB_LAUNCH((k_nl_commit<F, 2>), gc, 256, s, c, c.K == 0 ? 1 : 0);
  c.gate = nullptr;
  return 0;
}

template <typename F>
int neighborlist_force(cudaStream_t s, Ctx<F>& c) {
  if (c.n == 0) return 0;
  int rc = neighborlist_refresh<F>(s, c);
  if (rc) return rc;
  const dim3 grid(cdiv(c.n, 128), c.batch);
  if (c.dim == 3) {
    JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_nl_force<F, L, 3>), grid, 128, s, c));
  } else {
    JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_nl_force<F, L, 2>), grid, 128, s, c));
  }
  return 0;
}

template <typename F>
int neighborlist_energy(cudaStream_t s, Ctx<F>& c, F* energy) {
  if (c.n == 0) return cudaMemsetAsync(energy, 0, sizeof(F) * c.batch, s) == cudaSuccess ? 0 : JDB200_ECUDA;
  int rc = neighborlist_refresh<F>(s, c);
  if (rc) return rc;
  const dim3 grid(c.reduce_blocks, c.batch);
  JDB_LAW_SWITCH(c.law, JDB_LAUNCH((k_nl_energy<F, L>), grid, kReduceBlock, s, c));
  JDB_LAUNCH(k_nl_final_sum<F>, dim3(c.batch), kReduceBlock, s, c.partial, c.reduce_blocks, energy);
  return 0;
}

#define JDB_INST(F)                                                 \
  template int neighborlist_refresh<F>(cudaStream_t, Ctx<F>&);      \
  template int neighborlist_force<F>(cudaStream_t, Ctx<F>&);        \
  template int neighborlist_energy<F>(cudaStream_t, Ctx<F>&, F*);
JDB_INST(float)
JDB_INST(double)

}  // namespace jdb
