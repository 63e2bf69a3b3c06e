/* jaxdem_b200 — C ABI of the B200-native DEM step engine (libjaxdem_b200.so).
 *
 * Drop-in boundary for the per-timestep hot path of cdelv/JaxDEM
 * (System.step / _step_once, jaxdem/system.py:60-98).  Each entry point below
 * replaces one plugin hook of the reference; the hook it replaces is cited as
 * reference file:line.  INTEGRATION.md shows the jax.ffi binding a JaxDEM
 * maintainer would add on top of these symbols.
 *
 * Rules every entry point obeys:
 *  - stream-ordered and asynchronous: work is enqueued on `stream`, the call
 *    never synchronises with the host, never allocates, never throws, and
 *    keeps no state of the data path in globals => legal inside jit / lax.scan /
 *    lax.fori_loop and CUDA-graph capture, re-entrant across threads.  (The only
 *    globals are diagnostic: a relaxed launch counter and, while
 *    jdb200_timing_enable(1) is on, a mutex-guarded list of timing events;
 *    neither influences any result);
 *  - all buffers are caller-owned DEVICE pointers, dense row-major, exactly
 *    the State / System pytree leaves of the reference (jaxdem/state.py:103-228),
 *    with an optional leading batch axis of size `params.batch` (vmap);
 *  - everything the reference treats as a traced leaf (dt, box, anchor,
 *    cell_size, material tables, ...) is read from device memory; only shapes,
 *    dtypes and type names are host-side (`jdb200_params`);
 *  - scratch comes from a caller-provided workspace sized by
 *    jdb200_workspace_bytes();
 *  - return value: 0 on success, negative JDB200_E* for ARGUMENT errors only.
 *    Data-dependent failures (hash overflow, neighbour-list overflow) are
 *    written to device flags, mirroring Collider.overflow
 *    (jaxdem/colliders/__init__.py:51-54);
 *  - State buffers are updated IN PLACE (bind with operand/result aliasing
 *    under XLA FFI).
 *
 * dtype pairs: JDB200_F32 = <float, int32_t>, JDB200_F64 = <double, int64_t>
 * (jax_enable_x64 off / on).  `fixed` and boolean flags are uint8.
 */
#ifndef JAXDEM_B200_H
#define JAXDEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JDB200_ABI_VERSION 4

#if defined(__GNUC__)
#define JDB200_API __attribute__((visibility("default")))
#else
#define JDB200_API
#endif

/* error codes */
#define JDB200_OK 0
#define JDB200_EINVAL (-1)    /* bad params (dim, dtype, law, ...) */
#define JDB200_ENULL (-2)     /* a required pointer is NULL */
#define JDB200_EWORKSPACE (-3) /* workspace too small */
#define JDB200_ECUDA (-4)     /* kernel launch failed (cudaGetLastError) */

enum { JDB200_F32 = 0, JDB200_F64 = 1 };
enum { JDB200_DOMAIN_FREE = 0, JDB200_DOMAIN_PERIODIC = 1, JDB200_DOMAIN_REFLECT = 2 };
enum { JDB200_LAW_SPRING = 0, JDB200_LAW_HERTZ = 1, JDB200_LAW_CUNDALLSTRACK = 2 };
enum { JDB200_LIN_NONE = 0, JDB200_LIN_VERLET = 1, JDB200_LIN_EULER = 2 };
enum { JDB200_ROT_NONE = 0, JDB200_ROT_VERLETSPIRAL = 1, JDB200_ROT_SPIRAL = 2 };
enum { JDB200_COLLIDER_NONE = 0, JDB200_COLLIDER_CELLLIST = 1, JDB200_COLLIDER_NAIVE = 2,
       JDB200_COLLIDER_NEIGHBORLIST = 3 /* Verlet list on top of the cell list; needs a jdb200_nlist */,
       JDB200_COLLIDER_MULTICELLLIST = 4 /* DynamicMultiCellList (jaxdem/colliders/multi_cell_list.py:46-73,142-254):
                                            the jdb200_celllist_* entry points with this value prune every stencil
                                            cell by its expandable AABB (segmented min / max over the cell's sorted
                                            run) before walking it; same results as the cell list */ };
/* cell-table strategy.  AUTO picks, per system and per call, ON THE DEVICE:
 * DENSE (counting sort into a dense cell table) when every cell hash lies in
 * [0, max_cells) and no cell holds more than JDB200_DENSE_MAX_OCC particles,
 * else — force / energy / neighbour-list calls on a grid with fewer than 2^31 cells — the same
 * counting sort into a HASHED table (row = hash of the cell key, rows sorted by (key, index),
 * look-ups filter by the exact key), else SORTED (stable LSD radix sort + binary search; any
 * hash values).  All produce identical pair sets and neighbour lists; perm / sorted hashes
 * come from DENSE or SORTED and are bit-identical. */
enum { JDB200_GRID_AUTO = 0, JDB200_GRID_DENSE = 1, JDB200_GRID_SORTED = 2 };
#define JDB200_DENSE_MAX_OCC 64

/* Trace-time-static values only (shapes, dtypes, type names). */
typedef struct jdb200_params {
  int64_t batch;          /* leading vmap axis, >= 1 */
  int64_t n;              /* particles per system */
  int64_t max_cells;      /* dense cell-table capacity per system (0 => SORTED) */
  int64_t key_window_lo[2];  /* optional: the caller promises that every cell hash — of the particles AND of */
  int64_t key_window_len[2]; /* their stencil cells — lies in [lo[w], lo[w] + len[w]) for w = 0 or 1 (slab
                                decomposition: the layers of one rank, two windows where they wrap around the
                                periodic box); only those rows of the dense table are zeroed and scanned.
                                len[0] == 0: the whole table.  With two windows len[0] is a multiple of 4096
                                and the windows are disjoint and ascending.  A hash outside the windows sends
                                the system to the sorted fallback (AUTO) or raises Collider.overflow (DENSE). */
  int32_t dim;            /* 2 or 3 */
  int32_t dtype;          /* JDB200_F32 / JDB200_F64 */
  int32_t domain;         /* JDB200_DOMAIN_* (periodic <=> Domain.periodic) */
  int32_t law;            /* JDB200_LAW_* (ForceModel type_name) */
  int32_t collider;       /* JDB200_COLLIDER_* */
  int32_t linear_integrator;   /* JDB200_LIN_* */
  int32_t rotation_integrator; /* JDB200_ROT_* */
  int32_t stencil_m;      /* rows of neighbor_mask */
  int32_t bond_width;     /* W of bond_id (N, W) */
  int32_t n_materials;    /* M of the material tables */
  int32_t max_neighbors;  /* K of create_neighbor_list */
  int32_t grid_mode;      /* JDB200_GRID_* */
  int32_t clumps;         /* 0: caller promises clump_id == arange(N) (spheres only; the
                             clump reductions are identities), 1: general clump ids */
  int32_t promises;       /* JDB200_PROMISE_* bits: facts about the VALUES of some leaves that the caller
                             knows on the host (it created the arrays) and that let the kernels skip whole
                             streams.  0 is always legal; a broken promise gives wrong results, not a crash. */
} jdb200_params;

/* jdb200_params.promises */
#define JDB200_PROMISE_NO_EXT 1    /* external_force / external_force_com / external_torque are all zero on entry
                                      (ForceManager: nothing was added since the last apply cleared them) */
#define JDB200_PROMISE_NO_BONDS 2  /* every bond_id entry is -1 */
#define JDB200_PROMISE_NO_FIXED 4  /* no particle is fixed */
#define JDB200_PROMISE_NO_POS_P 8  /* pos_p == 0, hence pos_p_rot == 0 and pos == pos_c (sphere systems) */

/* State leaves (jaxdem/state.py:103-228).  F = float/double, I = int32/int64.
 * A = 3 in 3D, 1 in 2D.  NULL is allowed for leaves an entry point does not use. */
typedef struct jdb200_state {
  void* pos_c;     /* (B,N,D) F */
  void* pos_p;     /* (B,N,D) F */
  void* vel;       /* (B,N,D) F */
  void* force;     /* (B,N,D) F */
  void* q_w;       /* (B,N,1) F */
  void* q_xyz;     /* (B,N,3) F */
  void* ang_vel;   /* (B,N,A) F */
  void* torque;    /* (B,N,A) F */
  void* inertia;   /* (B,N,A) F */
  void* rad;       /* (B,N) F */
  void* mass;      /* (B,N) F */
  void* clump_id;  /* (B,N) I */
  void* mat_id;    /* (B,N) I */
  void* bond_id;   /* (B,N,W) I, -1 padded */
  void* fixed;     /* (B,N) uint8 */
  void* pos_p_rot; /* (B,N,D) F  cache R(q)·pos_p (State._pos_p_rot) */
  const void* n_rows; /* optional () int64 ON THE DEVICE, or NULL: the number of LIVE rows, <= params.n.  Non-NULL
                         (batch == 1, dense cell table): params.n is only the launch bound; every kernel reads the
                         live count itself and leaves rows [n_rows, n) alone.  Used by the slab decomposition, whose
                         owned + ghost row count changes every step and is known on the device only. */
  const void* order_id; /* optional (B,N) int64 on the device, or NULL: particles that share a cell are ordered by this
                           id instead of their row index (dense cell table only).  The reference's perm is the stable
                           sort of (hash, iota) (colliders/_partition.py:91-93); with order_id = the GLOBAL particle id
                           a rank of the slab decomposition orders each cell exactly as the undecomposed system does,
                           so every particle sums its contacts in the same order (bit-identical forces). */
} jdb200_state;

/* System leaves (jaxdem/system.py:123-228 and the components it holds). */
typedef struct jdb200_system {
  void* dt;                    /* (B,) F */
  void* box_size;              /* (B,D) F   Domain.box_size */
  void* inv_box_size;          /* (B,D) F   Domain.inv_box_size */
  void* anchor;                /* (B,D) F   Domain.anchor */
  void* restitution;           /* (B,) F    ReflectDomain.restitution_coefficient */
  void* cell_size;             /* (B,) F    DynamicCellList.cell_size */
  void* neighbor_mask;         /* (B,M,D) I DynamicCellList.neighbor_mask */
  void* collider_overflow;     /* (B,) uint8 out: Collider.overflow */
  void* interact_same_bond_id; /* (B,) uint8 */
  void* gravity;               /* (B,D) F   ForceManager.gravity */
  void* external_force;        /* (B,N,D) F ForceManager.external_force (cleared by apply) */
  void* external_force_com;    /* (B,N,D) F */
  void* external_torque;       /* (B,N,A) F */
  void* mat_young;             /* (B,Mt) F  MaterialTable.young */
  void* mat_poisson;           /* (B,Mt) F */
  void* mat_e;                 /* (B,Mt) F */
  void* mat_mu;                /* (B,Mt) F */
  void* mat_mu_r;              /* (B,Mt) F */
  void* mat_young_eff;         /* (B,Mt,Mt) F MaterialTable.young_eff */
  void* time;                  /* (B,) F      System.time       } optional (NULL: the caller keeps them): advanced by */
  void* step_count;            /* (B,) int64  System.step_count } jdb200_system_step, once per step (system.py:62-63)  */
} jdb200_system;

JDB200_API int jdb200_abi_version(void);

/* Bytes of scratch any entry point below needs for these params. */
JDB200_API size_t jdb200_workspace_bytes(const jdb200_params* p);

/* ---- collider: DynamicCellList ("CellList") ------------------------------ */

/* _get_spatial_partition (jaxdem/colliders/cell_list.py:35-87) +
 * _grid_params (_partition.py:54-99).  Optional outputs (NULL to skip), for
 * parity checks and for callers that want the partition itself:
 *   perm        (B,N) I   stable-sort permutation
 *   sorted_hash (B,N) I   sorted cell hashes
 *   nbr_hash    (B,N,M) I neighbour-cell hashes after the periodic de-dup
 *                         (_dedup_stencil_hashes, cell_list.py:90-96)
 *   used_dense  (B,) uint8  which cell-table strategy ran: 0 sorted keys + binary search, 1 dense
 *                         table, 2 hashed table (a grid with more cells than max_cells: the table is
 *                         addressed by a hash of the cell key, rows sorted by (key, index), look-ups
 *                         filter by the exact key — same pair sets, no sort; never chosen when perm or
 *                         sorted_hash are requested, those need the global order) */
JDB200_API int jdb200_celllist_partition(void* stream, const jdb200_params* p, const jdb200_state* st,
                              const jdb200_system* sys, void* ws, size_t ws_bytes, void* perm,
                              void* sorted_hash, void* nbr_hash, void* used_dense);

/* DynamicCellList.compute_force (cell_list.py:434-464): writes st->force,
 * st->torque (= sum T + cross(pos_p_rot, sum F)) and sys->collider_overflow. */
JDB200_API int jdb200_celllist_compute_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                  const jdb200_system* sys, void* ws, size_t ws_bytes);

/* DynamicCellList.compute_potential_energy (cell_list.py:466-496):
 * energy (B,) F. */
JDB200_API int jdb200_celllist_compute_potential_energy(void* stream, const jdb200_params* p,
                                             const jdb200_state* st, const jdb200_system* sys,
                                             void* ws, size_t ws_bytes, void* energy);

/* DynamicCellList.create_neighbor_list (cell_list.py:498-595): cutoff (B,) F
 * on device; neighbor_list (B,N,K) I padded with -1; overflow (B,) uint8. */
JDB200_API int jdb200_celllist_create_neighbor_list(void* stream, const jdb200_params* p,
                                         const jdb200_state* st, const jdb200_system* sys,
                                         void* ws, size_t ws_bytes, const void* cutoff,
                                         void* neighbor_list, void* overflow);

/* DynamicCellList.create_cross_neighbor_list (colliders/cell_list.py:600-715): for every
 * query point pos_a (B, N_A, D) the database points — the State's positions pos_c + pos_p_rot,
 * p->n of them — within `cutoff` (B,), through the partition of the DATABASE with the cell size
 * inflated to cover the cutoff; no clump / bond mask; rows (B, N_A, K) in stencil x sorted-run
 * order, original database indices, -1 padded; overflow (B,) uint8.  Only pos_c, pos_p_rot and
 * rad of `st` are read (bond_width may be 0). */
JDB200_API int jdb200_celllist_create_cross_neighbor_list(void* stream, const jdb200_params* p,
                                                          const jdb200_state* st, const jdb200_system* sys, void* ws,
                                                          size_t ws_bytes, const void* pos_a, int64_t n_a,
                                                          const void* cutoff, void* neighbor_list, void* overflow);

/* NaiveSimulator.compute_force / compute_potential_energy
 * (jaxdem/colliders/naive.py:187-235, 73-113): O(N^2), the reference's default
 * collider (README config). */
JDB200_API int jdb200_naive_compute_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                               const jdb200_system* sys, void* ws, size_t ws_bytes);
JDB200_API int jdb200_naive_compute_potential_energy(void* stream, const jdb200_params* p,
                                          const jdb200_state* st, const jdb200_system* sys,
                                          void* ws, size_t ws_bytes, void* energy);

/* ---- collider: NeighborList ("NeighborList", jaxdem/colliders/neighbor_list.py:133-776) ------------ */
/* The collider's own leaves.  `sys` carries the SECONDARY collider's leaves (cell_size, neighbor_mask of the
 * DynamicCellList that rebuilds the list, neighbor_list.py:436-440) and sys->collider_overflow is
 * NeighborList.overflow.  p->max_neighbors = K (static), p->collider = JDB200_COLLIDER_NEIGHBORLIST. */
typedef struct jdb200_nlist {
  void* neighbor_list;  /* (B,N,K) I  NeighborList.neighbor_list, -1 padded, rows packed */
  void* old_pos;        /* (B,N,D) F  NeighborList.old_pos: positions at the last build */
  void* n_build_times;  /* (B,) I     NeighborList.n_build_times */
  const void* cutoff;   /* (B,) F     NeighborList.cutoff */
  const void* skin;     /* (B,) F     NeighborList.skin (absolute) */
} jdb200_nlist;

/* _check_and_rebuild (neighbor_list.py:57-131): max |pos - old_pos|^2 > skin^2/4 or n_build_times == 0 =>
 * rebuild through DynamicCellList.create_neighbor_list with radius cutoff + skin (:443-480), old_pos <- pos,
 * n_build_times += 1, overflow <- the builder's flag.  The decision is taken and acted upon ON THE DEVICE, per
 * system of the batch; no host synchronisation.  This is NeighborList.create_neighbor_list (:404-441). */
JDB200_API int jdb200_neighborlist_refresh(void* stream, const jdb200_params* p, const jdb200_state* st,
                                           const jdb200_system* sys, void* ws, size_t ws_bytes,
                                           const jdb200_nlist* nl);
/* NeighborList.compute_force (neighbor_list.py:542-632): refresh, then the list-driven force pass. */
JDB200_API int jdb200_neighborlist_compute_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                                 const jdb200_system* sys, void* ws, size_t ws_bytes,
                                                 const jdb200_nlist* nl);
/* NeighborList.compute_potential_energy (neighbor_list.py:634-727): energy (B,) F. */
JDB200_API int jdb200_neighborlist_compute_potential_energy(void* stream, const jdb200_params* p,
                                                            const jdb200_state* st, const jdb200_system* sys,
                                                            void* ws, size_t ws_bytes, const jdb200_nlist* nl,
                                                            void* energy);
/* n x _step_once with the NeighborList collider (same contract as jdb200_system_step). */
JDB200_API int jdb200_system_step_nl(void* stream, const jdb200_params* p, const jdb200_state* st,
                                     const jdb200_system* sys, void* ws, size_t ws_bytes, int64_t n_steps,
                                     const jdb200_nlist* nl);

/* ---- minimiser inner loop: `minimize` with the FIRE optimiser ------------------------------------------
 * jaxdem/minimizers/routines.py:151-383 (loop, carry, termination tests) and jaxdem/minimizers/optimizers.py:127-340
 * (`fire`: FIREState and its update).  Every array below is a leaf of the reference's while_loop carry. */
typedef struct jdb200_fire_state {
  void* vel_pos;  /* (B,N,D) F  FIREState.vel["pos_c"]  */
  void* vel_rot;  /* (B,N,A) F  FIREState.vel["rotvec"] */
  void* dt;       /* (B,) F     FIREState.dt    */
  void* alpha;    /* (B,) F     FIREState.alpha */
  void* n_good;   /* (B,) int64 FIREState.N_good */
  void* n_bad;    /* (B,) int64 FIREState.N_bad  */
  void* pe;       /* (B,) F     carry: potential energy of the current state (total, not per particle) */
  void* prev_pe;  /* (B,) F     carry: the one before (inf before the first iteration) */
  void* steps;    /* (B,) int64 carry: iterations taken */
  void* active;   /* (B,) int32 cond_fun of the while loop: 1 while the system keeps iterating */
} jdb200_fire_state;
typedef struct jdb200_fire_params { /* arguments of fire(...) and minimize(...); host scalars (static under jit) */
  double dt, alpha_init, f_inc, f_dec, f_alpha, dt_max_scale, dt_min_scale;
  double pe_tol, pe_diff_tol, force_tol;
  int64_t n_min, n_bad_max, max_steps;
} jdb200_fire_params;
/* init != 0: opt_state = init(params), the initial evaluation (routines.py:239-258), steps = 0, active = cond.
 * Then `n_iter` iterations of body_fun, each a no-op for the systems whose `active` is 0; pe / steps / active are
 * updated on the device after every iteration, so the caller may poll `active` between calls (or not at all and
 * run max_steps iterations).  p->collider selects the force / energy evaluation (cell list, naive, or neighbour
 * list: then `nl` must be given, else NULL).  State is updated in place: pos_c, q, _pos_p_rot, force, torque. */
JDB200_API int jdb200_minimize_fire(void* stream, const jdb200_params* p, const jdb200_state* st,
                                    const jdb200_system* sys, void* ws, size_t ws_bytes, const jdb200_nlist* nl,
                                    const jdb200_fire_state* fs, const jdb200_fire_params* fp, int64_t n_iter,
                                    int32_t init);

/* ---- ForceManager.apply (jaxdem/forces/force_manager.py:338-425) --------- */
JDB200_API int jdb200_force_manager_apply(void* stream, const jdb200_params* p, const jdb200_state* st,
                               const jdb200_system* sys, void* ws, size_t ws_bytes);

/* ---- integrators ---------------------------------------------------------- */
/* p->linear_integrator selects VelocityVerlet (velocity_verlet.py:57-61,92-95)
 * or DirectEuler (direct_euler.py:62-66). */
JDB200_API int jdb200_linear_step_before_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                    const jdb200_system* sys);
JDB200_API int jdb200_linear_step_after_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                   const jdb200_system* sys);
/* p->rotation_integrator selects VelocityVerletSpiral
 * (velocity_verlet_spiral.py:83-116,156-180) or Spiral (spiral.py:104-141);
 * both refresh st->pos_p_rot when q changes (State.__setattr__, state.py:264-273). */
JDB200_API int jdb200_rotation_step_before_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                      const jdb200_system* sys);
JDB200_API int jdb200_rotation_step_after_force(void* stream, const jdb200_params* p, const jdb200_state* st,
                                     const jdb200_system* sys);

/* ---- domains --------------------------------------------------------------- */
/* Domain.apply for p->domain: periodic = no-op (domains/__init__.py:156-189),
 * reflect = ReflectDomain.apply (domains/reflect.py:99-299, _toc.py:11-94),
 * free = FreeDomain.apply (domains/free.py:42-64; rewrites box_size/anchor). */
JDB200_API int jdb200_domain_apply(void* stream, const jdb200_params* p, const jdb200_state* st,
                        const jdb200_system* sys, void* ws, size_t ws_bytes);

/* ---- fused driver: n x _step_once (jaxdem/system.py:60-98) ---------------- */
/* Runs domain.apply -> inv_box refresh -> linear/rotation before -> collider ->
 * force manager -> linear/rotation after, `n_steps` times, all on `stream`, with
 * no host round trip.  Identity user pre/post hooks only. */
JDB200_API int jdb200_system_step(void* stream, const jdb200_params* p, const jdb200_state* st,
                       const jdb200_system* sys, void* ws, size_t ws_bytes, int64_t n_steps);

/* ---- trajectory output: the default save_fn of System.trajectory_rollout ------------------- */
/* Replaces the per-frame copy of the State pytree (jaxdem/system.py:55-57 `_save_state_system`, stacked by the
 * scan of `_trajectory_rollout`, :101-120).  One launch packs the selected leaves of a frame into a contiguous
 * record per system, (B, frame_len) F, in the order pos_c (N,D) | vel (N,D) | force (N,D) | ang_vel (N,A) |
 * torque (N,A) | q_w (N) | q_xyz (N,3) | pos = pos_c + pos_p_rot (N,D); `fields` bit f selects the f-th of them
 * (JDB200_FRAME_*).  `out` may be a slot of a device ring buffer that the caller drains to pinned host memory on a
 * second stream (jaxdem_b200/system.py FrameRing), or a row of a device-resident trajectory. */
#define JDB200_FRAME_POS_C 1
#define JDB200_FRAME_VEL 2
#define JDB200_FRAME_FORCE 4
#define JDB200_FRAME_ANG_VEL 8
#define JDB200_FRAME_TORQUE 16
#define JDB200_FRAME_Q_W 32
#define JDB200_FRAME_Q_XYZ 64
#define JDB200_FRAME_POS 128
JDB200_API int jdb200_frame_pack(void* stream, const jdb200_params* p, const jdb200_state* st,
                                 const jdb200_system* sys, int32_t fields, void* out);

/* collider.compute_force -> force_manager.apply -> linear_integrator.step_after_force ->
 * rotation_integrator.step_after_force in one call for sphere systems (clumps == 0) with
 * velocity Verlet (any rotation integrator): the tail of _step_once (system.py:75-80) behind a
 * point where the caller has to intervene between the drift and the force evaluation (the slab
 * exchange below).  Same arithmetic per particle as the hooks in sequence (the pair kernel's
 * epilogue applies the manager, the kick and the rotation update to the particle it owns).
 * JDB200_EINVAL for other configurations. */
JDB200_API int jdb200_celllist_force_step_after(void* stream, const jdb200_params* p, const jdb200_state* st,
                                                const jdb200_system* sys, void* ws, size_t ws_bytes);

/* ---- slab decomposition: the per-step neighbour exchange ----------------------------- */
/* No reference equivalent (the reference runs one system on one device,
 * jaxdem/system.py:60-98); SURVEY.md 8(e).  One periodic sphere system is cut into slabs of
 * cell layers along the LAST axis, one slab per rank (jaxdem_b200/slab.py).  After the drift
 * (step_before_force) `jdb200_slab_pack` classifies the n owned rows by the cell
 * layer of their last coordinate — the arithmetic of the collider's hash,
 * colliders/cell_list.py:55-60, on the device copies of anchor / box_size / cell_size —
 * and writes the rows that left as full records into the message of their direction (and as
 * ghost records into `kept`: they stay behind as ghosts; their row indices go to `holes`) and
 * the rows within `search_range` layers of a face as ghost records.  Inside a message section the
 * records are FIELD-MAJOR (word c of record r at base[c * capacity + r]): a warp's stores are
 * contiguous, which is what peer-memory stores over NVLink need.  The caller exchanges the
 * two messages with its neighbours (NCCL), reads the counts from the 64-byte headers (int64:
 * [0] full records, [1] ghost records, [2] rows that moved further than the halo;
 * `header_local`: [0] rows that stay, [1] left downwards, [2] left upwards, [3] strays) and
 * calls `jdb200_slab_unpack`, which repairs the owned rows IN PLACE — arrivals (from the lower,
 * then the upper neighbour) fill the lowest holes or are appended, rows from the tail fill the
 * holes that are left — and writes the ghost rows behind them (kept lower, kept upper, halo of
 * the lower, halo of the upper neighbour).  counts = {n_old, a_lo, a_up, k_lo, k_up, g_lo,
 * g_up}; afterwards n_own = n_old - k_lo - k_up + a_lo + a_up.  Same rules as every entry
 * point: stream-ordered, no host synchronisation, no allocation; all lists keep index order. */
typedef struct jdb200_slab_desc {
  int64_t n;            /* owned rows */
  int64_t cap_mig;      /* full-record capacity of one message */
  int64_t cap_ghost;    /* ghost-record capacity of one message */
  int32_t dim, dtype;   /* as jdb200_params */
  int32_t n_layers;     /* cell layers along the last axis (static slab layout) */
  int32_t lo_layer, up_layer; /* this rank owns layers [lo_layer, up_layer) */
  int32_t search_range; /* halo width in layers (R of the collider's stencil) */
  const void* anchor;   /* (D,) F  Domain.anchor       (device) */
  const void* box_size; /* (D,) F  Domain.box_size     (device) */
  const void* cell_size;/* ()   F  DynamicCellList.cell_size (device) */
  const void* dt;       /* ()   F  System.dt (device), or NULL.  Non-NULL: jdb200_slab_pack first applies
                           VelocityVerlet.step_before_force (velocity_verlet.py:57-61) to the owned rows */
} jdb200_slab_desc;

/* Per-particle rows of a slab (State leaves of a sphere system + the global particle id). */
typedef struct jdb200_slab_rows {
  void *pos_c, *vel, *force;        /* (cap,D) F */
  void *ang_vel, *torque, *inertia; /* (cap,A) F */
  void *q_w, *q_xyz;                /* (cap,1), (cap,3) F */
  void *rad, *mass;                 /* (cap,) F */
  void* mat_id;                     /* (cap,) I */
  void* fixed;                      /* (cap,) uint8 */
  void* gid;                        /* (cap,) int64 global particle id */
} jdb200_slab_rows;

JDB200_API size_t jdb200_slab_message_bytes(const jdb200_slab_desc* d);
JDB200_API size_t jdb200_slab_kept_bytes(const jdb200_slab_desc* d);
JDB200_API size_t jdb200_slab_scratch_bytes(const jdb200_slab_desc* d);
JDB200_API size_t jdb200_slab_holes_bytes(const jdb200_slab_desc* d);
JDB200_API int jdb200_slab_pack(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* rows, void* msg_lo,
                                void* msg_up, void* kept, void* holes, void* header_local, void* scratch,
                                size_t scratch_bytes);
JDB200_API int jdb200_slab_unpack(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* rows,
                                  const int64_t* counts, const void* from_lo, const void* from_up, const void* kept,
                                  void* holes);

/* The same exchange WITHOUT the host (device protocol; peer-memory transport only).  The row
 * counts live in `dev_state` (16 int64 words on the device: [0] owned rows, [1] owned + ghost rows —
 * point jdb200_state.n_rows of the hooks at these words —, [2] exchanges completed, [3] sticky
 * JDB200_SLAB_* status bits, [4..10] the counts of the last exchange as in jdb200_slab_unpack,
 * [11] / [12] largest migrant / ghost counts seen, [13] internal ticket); desc.n is only the launch
 * BOUND (rows the kernels may touch, <= the row capacity).  msg_lo / msg_up and from_lo / from_up
 * hold one pointer per PARITY (exchange number & 1): the messages this rank writes — normally
 * straight into the neighbours' receive buffers over NVLink — and the buffers the neighbours
 * write into.  jdb200_slab_pack_dev stores the records (every storing thread fences them at system
 * scope), then — from the block that finishes last — the counts and, behind another system-scope
 * fence, header word 7 = exchange number + 1.  `scratch` must be zero before the first call (the
 * calls leave their accumulators zeroed).  jdb200_slab_unpack_dev starts with one block that spins (acquire loads, at
 * most timeout_ns nanoseconds) on word 7 of both of its receive buffers, derives the counts, checks
 * them against the capacities and the bound, publishes the new row counts and repairs the rows as
 * jdb200_slab_unpack does.  Two parities are enough: a neighbour can only write the message of
 * exchange s + 2 after it has seen this rank's flag of exchange s + 1, which this rank stores after
 * it has unpacked exchange s.  Errors never stop the stream: they set status bits (the rows are then
 * left untouched) which the caller reads whenever it likes.  Capturable in a CUDA graph: no
 * argument changes from step to step. */
#define JDB200_SLAB_STRAY 1        /* a particle moved further than the halo in one step / the box changed */
#define JDB200_SLAB_MESSAGE_FULL 2 /* more migrants / ghosts than the message capacities */
#define JDB200_SLAB_ROWS_FULL 4    /* owned + ghost rows exceed the bound */
#define JDB200_SLAB_TIMEOUT 8      /* a neighbour's message did not arrive in time */
#define JDB200_SLAB_DEV_WORDS 16
JDB200_API int jdb200_slab_pack_dev(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* rows,
                                    void* dev_state, void* const msg_lo[2], void* const msg_up[2], void* kept,
                                    void* holes, void* header_local, void* scratch, size_t scratch_bytes);
JDB200_API int jdb200_slab_unpack_dev(void* stream, const jdb200_slab_desc* d, const jdb200_slab_rows* rows,
                                      void* dev_state, const void* const from_lo[2], const void* const from_up[2],
                                      const void* header_local, const void* kept, void* holes, int64_t timeout_ns);

/* Number of kernels the library has launched since load (diagnostic counter for
 * bench.py's `gpu_launches`; relaxed atomic, not part of the data path). */
JDB200_API int64_t jdb200_launch_count(void);

/* Diagnostic per-kernel device timing for bench.py's roofline (no reference equivalent;
 * the reference's benchmarks/run_benchmarks.py:27-51 times whole hooks with timeit).
 * While enabled, every kernel launch is bracketed by CUDA events on its stream (this
 * mode is NOT graph-capture safe and is off by default).  jdb200_timing_collect waits
 * for the recorded events, sums device milliseconds and launch counts per kernel name
 * (names: max_entries x 64 chars), clears the records and returns the number of names. */
JDB200_API int jdb200_timing_enable(int on);
JDB200_API int jdb200_timing_collect(int max_entries, char* names, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* JAXDEM_B200_H */
