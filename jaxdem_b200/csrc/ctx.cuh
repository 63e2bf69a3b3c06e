// jaxdem_b200 — kernel argument bundle built from the C-ABI structs.
#pragma once
#include "common.cuh"

namespace jdb {

template <typename F>
struct Ctx {
  using I = typename RT<F>::I;
  // static
  long long n;
  long long max_cells;
  long long cell_stride;  // ints per system in the dense cell tables (max_cells + 1 rounded up to 64)
  int batch, dim, A, periodic, domain, law, M, W, nmat, K, grid_mode, clumps, lin, rot;
  long long win_lo[2], win_len[2];  // dense cell-table windows in use (jdb200_params.key_window_*; len 0: whole table)
  int promises;  // jdb200_params.promises (JDB200_PROMISE_*)
  int want_energy;  // minimiser loop: the row kernel also accumulates each particle's share of the pair energy (sforce.w)
  int fused;  // fused sphere step driver (abi.cu system_step): hash kernel integrates, pair kernel finishes the step
  // state (in place)
  F *pos_c, *pos_p, *vel, *force, *q_w, *q_xyz, *ang_vel, *torque, *inertia, *rad, *mass, *pos_p_rot;
  I *clump_id, *mat_id, *bond_id;
  uint8_t* fixed;
  // system
  F *dt, *box, *inv_box, *anchor, *restitution, *cell_size, *gravity;
  F *ext_force, *ext_force_com, *ext_torque;
  F *young, *poisson, *e, *mu, *mu_r, *young_eff;
  I* mask;
  uint8_t *overflow, *interact;
  F* time;               // System.time / step_count, advanced by the n-step driver when given
  long long* step_count;
  int tick;              // host flag: this partition build opens a step of jdb200_system_step
  // workspace
  GridInfo<I>* gi;            // [B]
  I *key, *key_b, *key_c;      // [B*N] unsorted hashes + radix ping-pong
  I* skey;                     // [B*N] sorted hashes
  int *perm, *perm_b, *perm_c; // [B*N] final perm + scratch
  int* rank;                   // [B*N] arrival rank inside the cell (dense)
  int* cell_count;             // [B*(max_cells+1)] per-cell counts (dense; zeroed by every build)
  int* cell_start;             // [B*(max_cells+1)] exclusive starts (dense)
  int* tmp_key;                // [B*N] dense key of the particle in arrival slot k
  int2* slot_rec;              // [B*N] (particle index, key) of arrival slot k: ONE 8-byte random store per particle
  // shadow records in ORIGINAL order, written by k_hash: urec[i] = (x, y, z, rad) with pos = pos_c + pos_p_rot
  // (read back, coalesced, by k_scatter) and — fused flows — uvm[i] = (vx, vy, vz, mass) after the before-force
  // kick (read back, coalesced, by k_after).  Two separate arrays: each reader streams exactly what it needs.
  Vec4<F>* urec;
  Vec4<F>* uvm;
  // [B*N] 32-byte records in ARRIVAL order (slot = cell_start[key] + arrival rank), written by k_scatter with ONE
  // 256-bit store per particle: f32 (x, y, z, rad, index, key, 0, 0); f64 (x, y, z, rad) with (index, key) in slot_rec
  char* arec;
  int* inv;                    // [B*N] sorted slot of original particle i (k_finalize, dense path; aliases perm_b)
  Vec4<F>* sforce;             // [B*N] sorted-order contact force sums of the row kernel (aliases segf)
  Vec4<F>* storque;            // [B*N] sorted-order contact torque sums (aliases segf2)
  unsigned* coop_bar;          // [4] software state of the cooperative sort fallback
  int want_skey;               // host flag: dense builds also fill skey (partition export)
  unsigned long long* tile_state;  // [B*scan_tiles] decoupled look-back descriptors
  int* tile_counter;           // [B]
  int* radix_counts;           // [B*256*radix_blocks]
  int* radix_skip;             // [B]
  Vec4<F>* spos;               // [B*N] sorted (x, y, z, rad)
  Vec4<F>* svel;               // [B*N] sorted (vx, vy, vz, mass)       (cundallstrack)
  Vec4<F>* sang;               // [B*N] sorted (wx, wy, wz, 0)          (cundallstrack)
  int* sclump;                 // [B*N] sorted clump id | has_bond << 31
  int* smat;                   // [B*N] sorted material id
  F* partial;                  // [B*reduce_blocks] reduction partials
  int* seg;                    // [B*N] clump arrival ranks (force manager / reflect)
  F *segf, *segf2;             // [B*N*8] per-sphere scratch of the clump reductions
  int* cell_start_clump;       // [B*(N+1)] clump CSR starts
  unsigned long long* tile_state_clump;  // [B*tiles(N+1)]
  int* tile_counter_clump;     // [B]
  int scan_tiles, radix_blocks, reduce_blocks;
  // ---- Verlet NeighborList collider (nlist.cu; NULL / unused for every other entry point) ----
  const int* gate;             // [B] or NULL: the partition / list-build kernels of system b run only when gate[b] != 0
  int* nl_gate;                // [B] rebuild decision of this call (workspace)
  F* nl_part;                  // [B*reduce_blocks] per-block max squared displacement (workspace)
  F* nl_cut;                   // [B] cutoff + skin (workspace)
  I* nl_list;                  // (B,N,K) NeighborList.neighbor_list
  F* nl_old_pos;               // (B,N,D) NeighborList.old_pos
  I* nl_builds;                // (B,)    NeighborList.n_build_times
  const F *nl_cutoff, *nl_skin;  // (B,)
  // ---- ragged rows (slab decomposition without host synchronisation) ----
  const long long* n_dev;      // jdb200_state.n_rows or NULL: the LIVE row count, read on the device by every kernel
                               // (n above is then the launch bound: grids cover it, rows >= *n_dev are not touched)
  const long long* order_id;   // jdb200_state.order_id or NULL: in-cell order by this id instead of the row index
  // ---- MultiCellList (loose-grid AABB pruning, pair.cu) ----
  int prune;                   // 1: every stencil cell is tested against its expandable AABB before its run is walked
  const F* prune_cut;          // [B] or NULL: neighbour-list builds prune with a query box of +-cutoff around the point
  Vec4<F>* aabb_c;             // [B*N] cell AABB centre, stored at the FIRST sorted slot of the cell's run
  Vec4<F>* aabb_h;             // [B*N] cell AABB half extent, same indexing
  // ---- minimiser (minimize.cu) ----
  F* min_part;                 // [B*reduce_blocks*4] reduction partials (power lin / rot, max|grad|)
  F* min_scal;                 // [B*8] per-iteration scalars (old dt, new dt, new alpha, reverse dt, velocity scale)
  F* min_pe;                   // [B*2] potential energy of the evaluation (collider, force manager)
};

constexpr int kScanTile = 4096;    // cells per look-back tile (512 threads x 8)
constexpr int kRadixTile = 2048;   // keys per radix block (256 threads x 8)
constexpr int kReduceBlock = 256;

template <typename F>
inline size_t carve(Ctx<F>& c, void* ws) {
  using I = typename RT<F>::I;
  Bump b(ws);
  const size_t B = (size_t)c.batch, N = (size_t)c.n, BN = B * N;
  c.cell_stride = (c.max_cells + 1 + 63) / 64 * 64;
  c.scan_tiles = cdiv(c.max_cells + 1, kScanTile);
  c.radix_blocks = cdiv(c.n, kRadixTile);
  c.reduce_blocks = cdiv(c.n, kReduceBlock);
  c.gi = b.take<GridInfo<I>>(B);
  c.key = b.take<I>(BN);
  c.key_b = b.take<I>(BN);
  c.key_c = b.take<I>(BN);
  c.skey = b.take<I>(BN);
  c.perm = b.take<int>(BN);
  c.perm_b = b.take<int>(BN);
  c.perm_c = b.take<int>(BN);
  c.rank = b.take<int>(BN);
  c.cell_count = b.take<int>(B * (size_t)c.cell_stride);
  c.cell_start = b.take<int>(B * (size_t)c.cell_stride);
  c.tmp_key = b.take<int>(BN);
  c.slot_rec = b.take<int2>(BN);
  c.urec = b.take<Vec4<F>>(BN);
  c.uvm = b.take<Vec4<F>>(BN);
  c.arec = b.take<char>(32 * BN);
  c.coop_bar = b.take<unsigned>(4);
  c.tile_state = b.take<unsigned long long>(B * (size_t)c.scan_tiles);
  c.tile_counter = b.take<int>(B);
  c.radix_counts = b.take<int>(B * 256 * (size_t)c.radix_blocks);
  c.radix_skip = b.take<int>(B);
  c.spos = b.take<Vec4<F>>(BN + 8);  // the row kernel loads up to kU records past the end of a run
  c.svel = b.take<Vec4<F>>(BN);
  c.sang = b.take<Vec4<F>>(BN);
  c.sclump = b.take<int>(BN);
  c.smat = b.take<int>(BN);
  c.partial = b.take<F>(B * (size_t)c.reduce_blocks);
  c.seg = b.take<int>(BN);
  c.segf = b.take<F>(BN * 8 + 64);
  c.segf2 = b.take<F>(BN * 8 + 64);
  c.cell_start_clump = b.take<int>(B * (N + 1));
  c.tile_state_clump = b.take<unsigned long long>(B * (size_t)cdiv(c.n + 1, kScanTile));
  c.tile_counter_clump = b.take<int>(B);
  c.aabb_c = b.take<Vec4<F>>(BN);
  c.aabb_h = b.take<Vec4<F>>(BN);
  c.nl_gate = b.take<int>(B);
  c.nl_part = b.take<F>(B * (size_t)c.reduce_blocks);
  c.nl_cut = b.take<F>(B);
  c.min_part = b.take<F>(B * (size_t)c.reduce_blocks * 4);
  c.min_scal = b.take<F>(B * 8);
  c.min_pe = b.take<F>(B * 2);
  c.inv = c.perm_b;
  c.sforce = reinterpret_cast<Vec4<F>*>(c.segf);
  c.storque = reinterpret_cast<Vec4<F>*>(c.segf2);
  return b.off + 256;
}

template <typename F>
inline int make_ctx(Ctx<F>& c, const jdb200_params* p, const jdb200_state* st,
                    const jdb200_system* sys, void* ws) {
  using I = typename RT<F>::I;
  c = Ctx<F>{};
  c.n = p->n;
  c.max_cells = p->grid_mode == JDB200_GRID_SORTED ? 0 : p->max_cells;
  c.batch = (int)p->batch;
  c.dim = p->dim;
  c.A = p->dim == 2 ? 1 : 3;
  c.domain = p->domain;
  c.periodic = p->domain == JDB200_DOMAIN_PERIODIC;
  c.law = p->law;
  c.M = p->stencil_m;
  c.W = p->bond_width;
  c.nmat = p->n_materials;
  c.K = p->max_neighbors;
  c.grid_mode = p->grid_mode;
  c.clumps = p->clumps;
  c.promises = p->promises;
  c.lin = p->linear_integrator;
  for (int w = 0; w < 2; ++w) {
    c.win_lo[w] = p->key_window_lo[w];
    c.win_len[w] = p->key_window_len[w];
  }
  c.rot = p->rotation_integrator;
  c.prune = p->collider == JDB200_COLLIDER_MULTICELLLIST;
  if (c.prune) c.want_skey = 1;  // run boundaries of the sorted hashes, dense or sorted
  if (st) {
    c.pos_c = (F*)st->pos_c; c.pos_p = (F*)st->pos_p; c.vel = (F*)st->vel; c.force = (F*)st->force;
    c.q_w = (F*)st->q_w; c.q_xyz = (F*)st->q_xyz; c.ang_vel = (F*)st->ang_vel;
    c.torque = (F*)st->torque; c.inertia = (F*)st->inertia; c.rad = (F*)st->rad;
    c.mass = (F*)st->mass; c.pos_p_rot = (F*)st->pos_p_rot;
    c.clump_id = (I*)st->clump_id; c.mat_id = (I*)st->mat_id; c.bond_id = (I*)st->bond_id;
    c.fixed = (uint8_t*)st->fixed;
    c.n_dev = (const long long*)st->n_rows;
    c.order_id = (const long long*)st->order_id;
  }
  if (sys) {
    c.dt = (F*)sys->dt; c.box = (F*)sys->box_size; c.inv_box = (F*)sys->inv_box_size;
    c.anchor = (F*)sys->anchor; c.restitution = (F*)sys->restitution;
    c.cell_size = (F*)sys->cell_size; c.gravity = (F*)sys->gravity;
    c.ext_force = (F*)sys->external_force; c.ext_force_com = (F*)sys->external_force_com;
    c.ext_torque = (F*)sys->external_torque;
    c.young = (F*)sys->mat_young; c.poisson = (F*)sys->mat_poisson; c.e = (F*)sys->mat_e;
    c.mu = (F*)sys->mat_mu; c.mu_r = (F*)sys->mat_mu_r; c.young_eff = (F*)sys->mat_young_eff;
    c.mask = (I*)sys->neighbor_mask;
    c.overflow = (uint8_t*)sys->collider_overflow; c.interact = (uint8_t*)sys->interact_same_bond_id;
    c.time = (F*)sys->time; c.step_count = (long long*)sys->step_count;
  }
  carve(c, ws);
  return 0;
}

inline int check_params(const jdb200_params* p) {
  if (!p) return JDB200_ENULL;
  if (p->dim != 2 && p->dim != 3) return JDB200_EINVAL;
  if (p->dtype != JDB200_F32 && p->dtype != JDB200_F64) return JDB200_EINVAL;
  if (p->batch < 1 || p->n < 0 || p->n > 0x7fffffffLL) return JDB200_EINVAL;
  if (p->domain < 0 || p->domain > 2 || p->law < 0 || p->law > 2) return JDB200_EINVAL;
  if (p->grid_mode < 0 || p->grid_mode > 2 || p->max_cells < 0 || p->max_cells > 0x7ffffff0LL) return JDB200_EINVAL;
  if (p->bond_width < 0 || p->n_materials < 0 || p->stencil_m < 0 || p->max_neighbors < 0)
    return JDB200_EINVAL;
  for (int w = 0; w < 2; ++w)
    if (p->key_window_lo[w] < 0 || p->key_window_len[w] < 0) return JDB200_EINVAL;
  if (p->key_window_len[1] > 0 && (p->key_window_len[0] % 4096 != 0 || p->key_window_lo[1] < p->key_window_lo[0] + p->key_window_len[0]))
    return JDB200_EINVAL;
  return 0;
}

}  // namespace jdb
