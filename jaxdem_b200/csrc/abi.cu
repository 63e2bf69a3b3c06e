// jaxdem_b200 — extern "C" entry points (include/jaxdem_b200.h) and the fused
// n-step driver replacing _step_once / System.step (jaxdem/system.py:60-98,701-748).
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "ctx.cuh"
#include "launch.cuh"

namespace jdb {

std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_timing{0};

// ---- diagnostic per-kernel timing (bench.py roofline; never on in production) ----
struct TimingRec {
  const char* name;
  cudaEvent_t e0, e1;
};
static std::mutex g_timing_mu;
static std::vector<TimingRec> g_timing_recs;

void timing_begin(const char* name, cudaStream_t s) {
  TimingRec r{name, nullptr, nullptr};
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, s);
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timing_recs.push_back(r);
}
void timing_end(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  cudaEventRecord(g_timing_recs.back().e1, s);
}

template <typename F> int build_partition(cudaStream_t, Ctx<F>&, const F*, int, bool);
template <typename F> int celllist_force(cudaStream_t, Ctx<F>&, int, bool, bool);
template <typename F> int celllist_energy(cudaStream_t, Ctx<F>&, F*, bool);
template <typename F> int celllist_neighbor_list(cudaStream_t, Ctx<F>&, const F*, typename RT<F>::I*, uint8_t*);
template <typename F> int celllist_cross_neighbor_list(cudaStream_t, Ctx<F>&, const F*, long long, const F*, typename RT<F>::I*, uint8_t*);
template <typename F> int naive_force(cudaStream_t, Ctx<F>&);
template <typename F> int naive_energy(cudaStream_t, Ctx<F>&, F*);
template <typename F> int force_manager_apply(cudaStream_t, Ctx<F>&);
template <typename F> int domain_apply(cudaStream_t, Ctx<F>&);
template <typename F> int refresh_inv_box(cudaStream_t, Ctx<F>&);
template <typename F> int linear_before(cudaStream_t, Ctx<F>&);
template <typename F> int linear_after(cudaStream_t, Ctx<F>&);
template <typename F> int rotation_before(cudaStream_t, Ctx<F>&);
template <typename F> int rotation_after(cudaStream_t, Ctx<F>&);
template <typename F> int frame_pack(cudaStream_t, Ctx<F>&, int, void*);
template <typename F> int neighborlist_refresh(cudaStream_t, Ctx<F>&);
template <typename F> int minimize_fire(cudaStream_t, Ctx<F>&, int, const jdb200_fire_state*, const jdb200_fire_params*, long long, int);
template <typename F> int neighborlist_force(cudaStream_t, Ctx<F>&);
template <typename F> int neighborlist_energy(cudaStream_t, Ctx<F>&, F*);

// ---- partition outputs for parity checks -------------------------------------
template <typename F>
__global__ void __launch_bounds__(256) k_export_partition(Ctx<F> c, typename RT<F>::I* perm,
                                                           typename RT<F>::I* sorted_hash,
                                                           typename RT<F>::I* nbr_hash, uint8_t* used_dense) {
  pdl_prologue();
  using I = typename RT<F>::I;
  const int b = blockIdx.y;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n) return;
  const size_t off = (size_t)b * c.n;
  const GridInfo<I> g = c.gi[b];
  if (perm) perm[off + k] = (I)c.perm[off + k];
  if (sorted_hash) sorted_hash[off + k] = c.skey[off + k];
  if (k == 0 && used_dense) used_dense[b] = (uint8_t)((g.dense && !g.dense_fail) ? (g.hashed ? 2 : 1) : 0);
  if (nbr_hash) {
    // row k here is ORIGINAL particle k (the reference builds the table from unsorted coords)
    const F* pc = c.pos_c + (off + k) * c.dim;
    const F* pr = c.pos_p_rot + (off + k) * c.dim;
    I cc[3] = {0, 0, 0};
    for (int d = 0; d < c.dim; ++d)
      cc[d] = cell_coord<F, I>(RT<F>::add(pc[d], pr[d]), c.anchor[b * c.dim + d], c.box[b * c.dim + d],
                               c.cell_size[b], g.gd[d], c.periodic);
    const I* mask = c.mask + (size_t)b * c.M * c.dim;
    I* row = nbr_hash + (off + k) * c.M;
    for (int m = 0; m < c.M; ++m) {
      I h = neighbor_hash<F, I>(cc, mask + m * c.dim, g.gd, g.stride, c.dim, c.periodic);
      if (c.periodic)  // _dedup_stencil_hashes: later duplicates -> -1
        for (int m2 = 0; m2 < m; ++m2)
          if (neighbor_hash<F, I>(cc, mask + m2 * c.dim, g.gd, g.stride, c.dim, c.periodic) == h) {
            h = I(-1);
            break;
          }
      row[m] = h;
    }
  }
}

template <typename F>
int partition_entry(cudaStream_t s, Ctx<F>& c, void* perm, void* sorted_hash, void* nbr_hash,
                    void* used_dense) {
  using I = typename RT<F>::I;
  c.want_skey = (perm || sorted_hash) ? 1 : 0;  // the sorted outputs need the globally sorted order (never the hashed table)
  int rc = build_partition<F>(s, c, nullptr, 0, false);
  if (rc || c.n == 0) return rc;
  JDB_LAUNCH(k_export_partition<F>, dim3(cdiv(c.n, 256), c.batch), 256, s, c, (I*)perm, (I*)sorted_hash,
             (I*)nbr_hash, (uint8_t*)used_dense);
  return 0;
}

template <typename F>
int collider_force(cudaStream_t s, Ctx<F>& c, int collider) {
  if (collider == JDB200_COLLIDER_CELLLIST || collider == JDB200_COLLIDER_MULTICELLLIST)
    return celllist_force<F>(s, c, 0, false, true);
  if (collider == JDB200_COLLIDER_NAIVE) return naive_force<F>(s, c);
  if (collider == JDB200_COLLIDER_NEIGHBORLIST) return neighborlist_force<F>(s, c);
  // "" no-op collider zeroes force and torque (colliders/__init__.py:56-88)
  if (c.n == 0) return 0;
  if (cudaMemsetAsync(c.force, 0, sizeof(F) * c.batch * c.n * c.dim, s) != cudaSuccess) return JDB200_ECUDA;
  if (cudaMemsetAsync(c.torque, 0, sizeof(F) * c.batch * c.n * c.A, s) != cudaSuccess) return JDB200_ECUDA;
  return 0;
}

template <typename F>
__global__ void k_advance_clock(Ctx<F> c, long long n_steps) {
  pdl_prologue();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= c.batch) return;
  if (c.time) {
    F t = c.time[b];
    const F dt = c.dt[b];
    for (long long i = 0; i < n_steps; ++i) t = RT<F>::add(t, dt);  // one rounding per step, as the reference
    c.time[b] = t;
  }
  if (c.step_count) c.step_count[b] += n_steps;
}
template <typename F>
int advance_clock(cudaStream_t s, Ctx<F>& c, long long n_steps) {
  JDB_LAUNCH(k_advance_clock<F>, dim3(cdiv(c.batch, 64)), 64, s, c, n_steps);
  return 0;
}

// _step_once (system.py:60-82), identity user hooks.  time/step_count are host-side
// bookkeeping of the caller.
template <typename F>
int system_step(cudaStream_t s, Ctx<F>& c, int collider, long long n_steps) {
  int rc = 0;
  // Fused flow for sphere systems (clump_id == arange), periodic box (domain.apply is a
  // no-op), linear velocity Verlet, no rotation integrator.  Per step: ONE partition build
  // whose setup kernel refreshes inv_box_size and whose hash kernel applies
  // step_before_force (kick + drift), and ONE pair kernel whose epilogue applies the collider
  // epilogue, ForceManager.apply and step_after_force to the particle it owns.  Same
  // arithmetic, same order per particle as the hook-by-hook sequence; the torque store is
  // skipped on steps whose torque nothing can observe.
  // With a rotation integrator (fused == 2) its step_before_force stays a streaming launch in front of the
  // partition and its step_after_force joins the epilogue (the torque is then stored on every step: the next
  // step_before_force reads it).
  const bool fused = collider == JDB200_COLLIDER_CELLLIST && !c.clumps && c.lin == JDB200_LIN_VERLET &&
                     c.domain == JDB200_DOMAIN_PERIODIC && c.n > 0;
  if (fused) {
    c.fused = c.rot == JDB200_ROT_NONE ? 1 : 2;
    c.tick = 1;  // the setup kernel of every step advances System.time / step_count
    for (long long it = 0; it < n_steps && !rc; ++it) {
      if (c.fused == 2 && (rc = rotation_before<F>(s, c))) break;
      rc = celllist_force<F>(s, c, 3, false, c.fused == 2 || it == n_steps - 1);
    }
    return rc;
  }
  if (c.time || c.step_count) {  // hook-by-hook flow: one tiny launch per call
    rc = advance_clock<F>(s, c, n_steps);
    if (rc) return rc;
  }
  if (c.domain != JDB200_DOMAIN_FREE && n_steps > 0) rc = refresh_inv_box<F>(s, c);  // box is constant
  for (long long it = 0; it < n_steps && !rc; ++it) {
    if ((rc = domain_apply<F>(s, c))) break;  // free: also refreshes inv_box_size
    if ((rc = linear_before<F>(s, c))) break;
    if ((rc = rotation_before<F>(s, c))) break;
    if ((rc = collider_force<F>(s, c, collider))) break;
    if ((rc = force_manager_apply<F>(s, c))) break;
    if ((rc = linear_after<F>(s, c))) break;
    if ((rc = rotation_after<F>(s, c))) break;
  }
  return rc;
}

}  // namespace jdb

using namespace jdb;

#define JDB_ENTER(NEED_WS)                                               \
  int rc_ = check_params(p);                                             \
  if (rc_) return rc_;                                                   \
  if (!st || !sys) return JDB200_ENULL;                                  \
  if (NEED_WS) {                                                         \
    if (!ws) return JDB200_ENULL;                                        \
    if (ws_bytes < jdb200_workspace_bytes(p)) return JDB200_EWORKSPACE;  \
  }                                                                      \
  cudaStream_t s = (cudaStream_t)stream;

#define JDB_DISPATCH(EXPR)                         \
  if (p->dtype == JDB200_F32) {                    \
    using F = float;                               \
    Ctx<F> c;                                      \
    make_ctx<F>(c, p, st, sys, ws);                \
    return EXPR;                                   \
  } else {                                         \
    using F = double;                              \
    Ctx<F> c;                                      \
    make_ctx<F>(c, p, st, sys, ws);                \
    return EXPR;                                   \
  }

template <typename F>
static int bind_nlist(Ctx<F>& c, const jdb200_nlist* nl) {
  using I = typename RT<F>::I;
  if (!nl || !nl->old_pos || !nl->n_build_times || !nl->cutoff || !nl->skin) return JDB200_ENULL;
  if (!nl->neighbor_list && c.K > 0 && c.n > 0) return JDB200_ENULL;
  c.nl_list = (I*)nl->neighbor_list;
  c.nl_old_pos = (F*)nl->old_pos;
  c.nl_builds = (I*)nl->n_build_times;
  c.nl_cutoff = (const F*)nl->cutoff;
  c.nl_skin = (const F*)nl->skin;
  return 0;
}

extern "C" {

JDB200_API int jdb200_abi_version(void) { return JDB200_ABI_VERSION; }

JDB200_API int64_t jdb200_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

JDB200_API int jdb200_timing_enable(int on) {
  g_timing.store(on ? 1 : 0, std::memory_order_relaxed);
  return 0;
}

JDB200_API int jdb200_timing_collect(int max_entries, char* names, double* total_ms, int64_t* launches) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  std::map<std::string, std::pair<double, int64_t>> acc;
  std::vector<std::string> order;
  for (auto& r : g_timing_recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess) cudaEventElapsedTime(&ms, r.e0, r.e1);
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    std::string nm(r.name);
    size_t a = nm.find_first_not_of("( ");
    size_t b = nm.find_first_of("<) ", a);
    nm = nm.substr(a, b == std::string::npos ? b : b - a);
    if (!acc.count(nm)) order.push_back(nm);
    acc[nm].first += ms;
    acc[nm].second += 1;
  }
  g_timing_recs.clear();
  int n = 0;
  for (auto& nm : order) {
    if (n >= max_entries) break;
    std::strncpy(names + (size_t)n * 64, nm.c_str(), 63);
    names[(size_t)n * 64 + 63] = 0;
    total_ms[n] = acc[nm].first;
    launches[n] = acc[nm].second;
    ++n;
  }
  return n;
}

JDB200_API size_t jdb200_workspace_bytes(const jdb200_params* p) {
  if (check_params(p)) return 0;
  if (p->dtype == JDB200_F32) {
    Ctx<float> c;
    make_ctx<float>(c, p, nullptr, nullptr, nullptr);
    return carve(c, nullptr);
  }
  Ctx<double> c;
  make_ctx<double>(c, p, nullptr, nullptr, nullptr);
  return carve(c, nullptr);
}

JDB200_API int jdb200_celllist_partition(void* stream, const jdb200_params* p, const jdb200_state* st,
                              const jdb200_system* sys, void* ws, size_t ws_bytes, void* perm,
                              void* sorted_hash, void* nbr_hash, void* used_dense) {
  JDB_ENTER(true)
  JDB_DISPATCH(partition_entry<F>(s, c, perm, sorted_hash, nbr_hash, used_dense))
}

JDB200_API int jdb200_celllist_compute_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                  const jdb200_system* sys, void* ws, size_t ws_bytes) {
  JDB_ENTER(true)
  JDB_DISPATCH(celllist_force<F>(s, c, 0, false, true))
}

JDB200_API int jdb200_celllist_compute_potential_energy(void* stream, const jdb200_params* p,
                                             const jdb200_state* st, const jdb200_system* sys,
                                             void* ws, size_t ws_bytes, void* energy) {
  JDB_ENTER(true)
  if (!energy) return JDB200_ENULL;
  JDB_DISPATCH(celllist_energy<F>(s, c, (F*)energy, false))
}

JDB200_API int jdb200_celllist_create_neighbor_list(void* stream, const jdb200_params* p,
                                         const jdb200_state* st, const jdb200_system* sys,
                                         void* ws, size_t ws_bytes, const void* cutoff,
                                         void* neighbor_list, void* overflow) {
  JDB_ENTER(true)
  if (!cutoff || !overflow || (!neighbor_list && p->max_neighbors > 0 && p->n > 0)) return JDB200_ENULL;
  JDB_DISPATCH(celllist_neighbor_list<F>(s, c, (const F*)cutoff, (RT<F>::I*)neighbor_list,
                                         (uint8_t*)overflow))
}

JDB200_API int jdb200_celllist_create_cross_neighbor_list(void* stream, const jdb200_params* p,
                                                          const jdb200_state* st, const jdb200_system* sys, void* ws,
                                                          size_t ws_bytes, const void* pos_a, int64_t n_a,
                                                          const void* cutoff, void* neighbor_list, void* overflow) {
  JDB_ENTER(true)
  if (n_a < 0) return JDB200_EINVAL;
  if (!cutoff || !overflow || ((!neighbor_list || !pos_a) && p->max_neighbors > 0 && n_a > 0)) return JDB200_ENULL;
  JDB_DISPATCH(celllist_cross_neighbor_list<F>(s, c, (const F*)pos_a, (long long)n_a, (const F*)cutoff,
                                               (RT<F>::I*)neighbor_list, (uint8_t*)overflow))
}

JDB200_API int jdb200_naive_compute_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                               const jdb200_system* sys, void* ws, size_t ws_bytes) {
  JDB_ENTER(true)
  JDB_DISPATCH(naive_force<F>(s, c))
}

JDB200_API int jdb200_naive_compute_potential_energy(void* stream, const jdb200_params* p,
                                          const jdb200_state* st, const jdb200_system* sys,
                                          void* ws, size_t ws_bytes, void* energy) {
  JDB_ENTER(true)
  if (!energy) return JDB200_ENULL;
  JDB_DISPATCH(naive_energy<F>(s, c, (F*)energy))
}

JDB200_API int jdb200_force_manager_apply(void* stream, const jdb200_params* p, const jdb200_state* st,
                               const jdb200_system* sys, void* ws, size_t ws_bytes) {
  JDB_ENTER(true)
  JDB_DISPATCH(force_manager_apply<F>(s, c))
}

JDB200_API int jdb200_linear_step_before_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                    const jdb200_system* sys) {
  void* ws = nullptr;
  size_t ws_bytes = 0;
  JDB_ENTER(false)
  (void)ws_bytes;
  JDB_DISPATCH(linear_before<F>(s, c))
}
JDB200_API int jdb200_linear_step_after_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                   const jdb200_system* sys) {
  void* ws = nullptr;
  size_t ws_bytes = 0;
  JDB_ENTER(false)
  (void)ws_bytes;
  JDB_DISPATCH(linear_after<F>(s, c))
}
JDB200_API int jdb200_rotation_step_before_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                      const jdb200_system* sys) {
  void* ws = nullptr;
  size_t ws_bytes = 0;
  JDB_ENTER(false)
  (void)ws_bytes;
  JDB_DISPATCH(rotation_before<F>(s, c))
}
JDB200_API int jdb200_rotation_step_after_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                     const jdb200_system* sys) {
  void* ws = nullptr;
  size_t ws_bytes = 0;
  JDB_ENTER(false)
  (void)ws_bytes;
  JDB_DISPATCH(rotation_after<F>(s, c))
}

JDB200_API int jdb200_domain_apply(void* stream, const jdb200_params* p, const jdb200_state* st,
                        const jdb200_system* sys, void* ws, size_t ws_bytes) {
  JDB_ENTER(true)
  JDB_DISPATCH(domain_apply<F>(s, c))
}

JDB200_API int jdb200_celllist_force_step_after(void* stream, const jdb200_params* p, const jdb200_state* st,
                                                const jdb200_system* sys, void* ws, size_t ws_bytes) {
  JDB_ENTER(true)
  if (p->clumps || p->linear_integrator != JDB200_LIN_VERLET) return JDB200_EINVAL;
  const int fused = p->rotation_integrator == JDB200_ROT_NONE ? 1 : 2;  // 2: + rotation step_after_force
  if (p->dtype == JDB200_F32) {
    using F = float;
    Ctx<F> c;
    make_ctx<F>(c, p, st, sys, ws);
    c.fused = fused;
    return celllist_force<F>(s, c, 4, false, true);
  } else {
    using F = double;
    Ctx<F> c;
    make_ctx<F>(c, p, st, sys, ws);
    c.fused = fused;
    return celllist_force<F>(s, c, 4, false, true);
  }
}

JDB200_API int jdb200_frame_pack(void* stream, const jdb200_params* p, const jdb200_state* st,
                                 const jdb200_system* sys, int32_t fields, void* out) {
  void* ws = nullptr;
  size_t ws_bytes = 0;
  JDB_ENTER(false)
  (void)ws_bytes;
  if (!out || fields < 0 || fields > 255) return fields < 0 || fields > 255 ? JDB200_EINVAL : JDB200_ENULL;
  JDB_DISPATCH(frame_pack<F>(s, c, fields, out))
}

#define JDB_DISPATCH_NL(EXPR)                      \
  if (p->dtype == JDB200_F32) {                    \
    using F = float;                               \
    Ctx<F> c;                                      \
    make_ctx<F>(c, p, st, sys, ws);                \
    if (int e_ = bind_nlist<F>(c, nl)) return e_;  \
    return EXPR;                                   \
  } else {                                         \
    using F = double;                              \
    Ctx<F> c;                                      \
    make_ctx<F>(c, p, st, sys, ws);                \
    if (int e_ = bind_nlist<F>(c, nl)) return e_;  \
    return EXPR;                                   \
  }

JDB200_API int jdb200_neighborlist_refresh(void* stream, const jdb200_params* p, const jdb200_state* st,
                                           const jdb200_system* sys, void* ws, size_t ws_bytes,
                                           const jdb200_nlist* nl) {
  JDB_ENTER(true)
  JDB_DISPATCH_NL(neighborlist_refresh<F>(s, c))
}

JDB200_API int jdb200_neighborlist_compute_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                                 const jdb200_system* sys, void* ws, size_t ws_bytes,
                                                 const jdb200_nlist* nl) {
  JDB_ENTER(true)
  JDB_DISPATCH_NL(neighborlist_force<F>(s, c))
}

JDB200_API int jdb200_neighborlist_compute_potential_energy(void* stream, const jdb200_params* p,
                                                            const jdb200_state* st, const jdb200_system* sys,
                                                            void* ws, size_t ws_bytes, const jdb200_nlist* nl,
                                                            void* energy) {
  JDB_ENTER(true)
  if (!energy) return JDB200_ENULL;
  JDB_DISPATCH_NL(neighborlist_energy<F>(s, c, (F*)energy))
}

JDB200_API int jdb200_system_step_nl(void* stream, const jdb200_params* p, const jdb200_state* st,
                                     const jdb200_system* sys, void* ws, size_t ws_bytes, int64_t n_steps,
                                     const jdb200_nlist* nl) {
  JDB_ENTER(true)
  if (n_steps < 0 || p->collider != JDB200_COLLIDER_NEIGHBORLIST) return JDB200_EINVAL;
  JDB_DISPATCH_NL(system_step<F>(s, c, p->collider, (long long)n_steps))
}

JDB200_API int jdb200_minimize_fire(void* stream, const jdb200_params* p, const jdb200_state* st,
                                    const jdb200_system* sys, void* ws, size_t ws_bytes, const jdb200_nlist* nl,
                                    const jdb200_fire_state* fs, const jdb200_fire_params* fp, int64_t n_iter,
                                    int32_t init) {
  JDB_ENTER(true)
  if (!fs || !fp) return JDB200_ENULL;
  if (n_iter < 0 || fp->max_steps < 0) return JDB200_EINVAL;
  if (!fs->vel_pos || !fs->vel_rot || !fs->dt || !fs->alpha || !fs->n_good || !fs->n_bad || !fs->pe || !fs->prev_pe ||
      !fs->steps || !fs->active)
    return JDB200_ENULL;
  if (p->collider == JDB200_COLLIDER_NEIGHBORLIST) {
    JDB_DISPATCH_NL(minimize_fire<F>(s, c, p->collider, fs, fp, (long long)n_iter, init))
  }
  if (p->collider != JDB200_COLLIDER_CELLLIST && p->collider != JDB200_COLLIDER_NAIVE &&
      p->collider != JDB200_COLLIDER_MULTICELLLIST)
    return JDB200_EINVAL;
  JDB_DISPATCH(minimize_fire<F>(s, c, p->collider, fs, fp, (long long)n_iter, init))
}

JDB200_API int jdb200_system_step(void* stream, const jdb200_params* p, const jdb200_state* st,
                       const jdb200_system* sys, void* ws, size_t ws_bytes, int64_t n_steps) {
  JDB_ENTER(true)
  if (n_steps < 0 || p->collider == JDB200_COLLIDER_NEIGHBORLIST) return JDB200_EINVAL;  // -> jdb200_system_step_nl
  JDB_DISPATCH(system_step<F>(s, c, p->collider, (long long)n_steps))
}

}  // extern "C"
