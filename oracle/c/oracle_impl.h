/* CPU oracle, C restatement of the JaxDEM step path.  TEST INFRASTRUCTURE ONLY.
 *
 * Included twice by oracle.c with (REAL, INT, SFX) = (float, int32_t, _f32) and
 * (double, int64_t, _f64): the reference computes in <float32,int32> by default and in
 * <float64,int64> under jax_enable_x64.  Compiled with
 * -ffp-contract=off and without -ffast-math so every expression rounds where the
 * reference's jnp expression rounds (XLA may still contract / reorder; that is what
 * the rel 1e-5 / 1e-12 tolerance absorbs).
 *
 * Each function cites the reference file:line it restates (paths relative to the
 * reference checkout).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef struct FN(orc_sys) {
  int64_t n;
  int32_t dim, periodic, law, M, W, nmat, interact, pad;
  /* State leaves (jaxdem/state.py:103-228) */
  REAL *pos_c, *pos_p_rot, *vel, *force, *torque, *ang_vel, *rad, *mass;
  INT *clump_id, *mat_id, *bond_id;
  uint8_t* fixed;
  /* System leaves */
  REAL *dt, *box, *inv_box, *anchor, *cell_size, *gravity;
  INT* mask; /* (M, dim) neighbor_mask */
  REAL *young, *poisson, *e, *mu, *mu_r, *young_eff;
  /* partition outputs / scratch, all (n,) */
  INT *hash, *sorted_hash, *perm, *tmp_key, *tmp_val;
  REAL* pos; /* (n, dim) scratch: pos_c + pos_p_rot */
  int32_t overflow, pad2;
} FN(orc_sys);

/* float -> int with the saturating semantics XLA and CUDA share (NaN -> 0). */
static inline INT FN(to_int)(REAL x) {
  if (x != x) return 0;
  if (x >= (REAL)INT_MAX_V) return INT_MAX_V;
  if (x <= (REAL)INT_MIN_V) return INT_MIN_V;
  return (INT)x;
}

/* _grid_params (jaxdem/colliders/_partition.py:54-99) */
static void FN(grid_params)(const FN(orc_sys) * s, INT* gd, INT* stride, int* overflow) {
  REAL total = (REAL)1;
  for (int d = 0; d < s->dim; ++d) {
    REAL q = s->box[d] / s->cell_size[0];
    INT g = FN(to_int)(s->periodic ? FLOOR(q) : CEIL(q));
    if (g < 1) g = 1;
    gd[d] = g;
    total = total * (REAL)g;
  }
  UINT acc = 1;
  for (int d = 0; d < s->dim; ++d) { /* wrapping integer cumprod */
    stride[d] = (INT)acc;
    acc = acc * (UINT)gd[d];
  }
  *overflow = total > (REAL)INT_MAX_V;
}

/* cell coordinate of one axis (jaxdem/colliders/cell_list.py:55-60) */
static inline INT FN(cell_coord)(const FN(orc_sys) * s, REAL x, int d, INT g) {
  if (s->periodic) {
    REAL u = (x - s->anchor[d]) / s->box[d];
    REAL r = FMOD(u, (REAL)1); /* jnp.remainder(u, 1) = fmod + sign fix-up */
    if (r != (REAL)0 && r < (REAL)0) r = r + (REAL)1;
    return FN(to_int)(FLOOR(r * (REAL)g));
  }
  return FN(to_int)(FLOOR((x - s->anchor[d]) / s->cell_size[0]));
}

/* stable LSD radix sort of (key, val) by key (signed), 8 bits per pass:
 * jax.lax.sort([hash, iota], num_keys=1) is stable (cell_list.py:64). */
static void FN(stable_sort)(int64_t n, INT* key, INT* val, INT* tk, INT* tv) {
  INT *ka = key, *va = val, *kb = tk, *vb = tv;
  for (int pass = 0; pass < (int)sizeof(INT); ++pass) {
    int64_t cnt[257];
    memset(cnt, 0, sizeof(cnt));
    const int sh = pass * 8;
    const UINT flip = (UINT)1 << (sizeof(INT) * 8 - 1);
    for (int64_t i = 0; i < n; ++i) cnt[((((UINT)ka[i]) ^ flip) >> sh & 0xff) + 1]++;
    int skip = 0;
    for (int b = 0; b < 256; ++b)
      if (cnt[b + 1] == n) skip = 1;
    if (skip) continue; /* all keys share this digit: the pass is the identity */
    for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
    for (int64_t i = 0; i < n; ++i) {
      int64_t p = cnt[(((UINT)ka[i]) ^ flip) >> sh & 0xff]++;
      kb[p] = ka[i];
      vb[p] = va[i];
    }
    INT* t = ka; ka = kb; kb = t;
    t = va; va = vb; vb = t;
  }
  if (ka != key) {
    memcpy(key, ka, sizeof(INT) * n);
    memcpy(val, va, sizeof(INT) * n);
  }
}

/* _get_spatial_partition (cell_list.py:35-87): hash (unsorted), sorted_hash, perm. */
void FN(orc_partition)(FN(orc_sys) * s) {
  INT gd[3] = {1, 1, 1}, stride[3] = {0, 0, 0};
  int ovf;
  FN(grid_params)(s, gd, stride, &ovf);
  s->overflow = ovf;
  const int D = s->dim;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < s->n; ++i) {
    UINT h = 0;
    for (int d = 0; d < D; ++d) {
      REAL x = s->pos_c[i * D + d] + s->pos_p_rot[i * D + d]; /* State.pos (state.py:295-304) */
      s->pos[i * D + d] = x;
      h += (UINT)FN(cell_coord)(s, x, d, gd[d]) * (UINT)stride[d];
    }
    s->hash[i] = (INT)h;
    s->sorted_hash[i] = (INT)h;
    s->perm[i] = (INT)i;
  }
  FN(stable_sort)(s->n, s->sorted_hash, s->perm, s->tmp_key, s->tmp_val);
}

/* Domain._displacement, multiply form (jaxdem/domains/periodic.py:75-79) */
static inline void FN(disp_mul)(const FN(orc_sys) * s, const REAL* a, const REAL* b, REAL* r) {
  for (int d = 0; d < 3; ++d) r[d] = (REAL)0;
  for (int d = 0; d < s->dim; ++d) {
    REAL x = a[d] - b[d];
    if (s->periodic) x = x - s->box[d] * RINT(x * s->inv_box[d]);
    r[d] = x;
  }
}

/* linalg.unit_and_norm (jaxdem/utils/linalg.py:162-181) */
static inline REAL FN(unit_and_norm)(const REAL* v, REAL* u) {
  REAL n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  REAL m = n2 > (REAL)1e-16 ? n2 : (REAL)1e-16;
  REAL inv = n2 == (REAL)0 ? (REAL)0 : (REAL)1 / SQRT(m);
  u[0] = v[0] * inv; u[1] = v[1] * inv; u[2] = v[2] * inv;
  return n2 * inv;
}

static inline void FN(cross)(const REAL* a, const REAL* b, REAL* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* force / torque ON i DUE TO j (ForceModel.force, jaxdem/forces/__init__.py:55-150) */
static void FN(pair_force)(const FN(orc_sys) * s, int64_t i, int64_t j, REAL* f, REAL* t) {
  const int D = s->dim, A = D == 3 ? 3 : 1;
  REAL rij[3];
  FN(disp_mul)(s, s->pos + i * D, s->pos + j * D, rij);
  const REAL Ri = s->rad[i], Rj = s->rad[j];
  const int mi = (int)s->mat_id[i], mj = (int)s->mat_id[j];
  const REAL neq = i != j ? (REAL)1 : (REAL)0;
  t[0] = t[1] = t[2] = (REAL)0;
  if (s->law == 0) { /* jaxdem/forces/spring.py:97-108 */
    REAL d2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
    REAL k = s->young_eff[mi * s->nmat + mj];
    REAL m = d2 > (REAL)1e-16 ? d2 : (REAL)1e-16;
    REAL inv = d2 == (REAL)0 ? (REAL)0 : (REAL)1 / SQRT(m);
    REAL r = d2 * inv;
    REAL delta = (Ri + Rj - r > (REAL)0 ? Ri + Rj - r : (REAL)0) * neq;
    REAL mag = k * delta * inv;
    f[0] = mag * rij[0]; f[1] = mag * rij[1]; f[2] = mag * rij[2];
  } else if (s->law == 1) { /* jaxdem/forces/hertz.py:98-119 */
    REAL Ei = s->young[mi], Ej = s->young[mj], ni = s->poisson[mi], nj = s->poisson[mj];
    REAL Rs = (Ri * Rj) / (Ri + Rj);
    REAL Es = (REAL)1 / (((REAL)1 - ni * ni) / Ei + ((REAL)1 - nj * nj) / Ej);
    REAL k = (REAL)(4.0 / 3.0) * Es * SQRT(Rs);
    REAL n[3];
    REAL r = FN(unit_and_norm)(rij, n);
    REAL delta = (Ri + Rj - r > (REAL)0 ? Ri + Rj - r : (REAL)0) * neq;
    REAL mag = k * POW(delta, (REAL)1.5);
    f[0] = mag * n[0]; f[1] = mag * n[1]; f[2] = mag * n[2];
  } else { /* jaxdem/forces/cundall_strack.py:119-198 */
    REAL Ei = s->young[mi], Ej = s->young[mj], nui = s->poisson[mi], nuj = s->poisson[mj];
    REAL Gi = Ei / ((REAL)2 * ((REAL)1 + nui)), Gj = Ej / ((REAL)2 * ((REAL)1 + nuj));
    REAL kn = ((REAL)2 * Ei * Ri * Ej * Rj) / (Ei * Ri + Ej * Rj);
    REAL kt = ((REAL)2 * Gi * Ri * Gj * Rj) / (Gi * Ri + Gj * Rj);
    REAL m_i = s->mass[i], m_j = s->mass[j];
    REAL m_eff = (m_i * m_j) / (m_i + m_j);
    REAL e_eff = s->e[mi] < s->e[mj] ? s->e[mi] : s->e[mj];
    REAL mu_eff = s->mu[mi] < s->mu[mj] ? s->mu[mi] : s->mu[mj];
    REAL e_safe = e_eff > (REAL)0 ? e_eff : (REAL)1;
    REAL ln_e = LOG(e_safe);
    REAL pi = (REAL)3.14159265358979323846;
    REAL beta = e_eff > (REAL)0 ? -ln_e / SQRT(pi * pi + ln_e * ln_e) : (REAL)1;
    REAL gamma_n = (REAL)2 * beta * SQRT(kn * m_eff);
    REAL gamma_t = (REAL)2 * beta * SQRT(kt * m_eff);
    REAL n[3];
    REAL r = FN(unit_and_norm)(rij, n);
    REAL delta = Ri + Rj - r;
    REAL is_contact = (delta > (REAL)0 && i != j) ? (REAL)1 : (REAL)0;
    delta = delta * is_contact;
    REAL rci[3] = {-Ri * n[0], -Ri * n[1], -Ri * n[2]};
    REAL rcj[3] = {Rj * n[0], Rj * n[1], Rj * n[2]};
    REAL wi[3] = {0, 0, 0}, wj[3] = {0, 0, 0}, vi[3] = {0, 0, 0}, vj[3] = {0, 0, 0};
    for (int d = 0; d < D; ++d) { vi[d] = s->vel[i * D + d]; vj[d] = s->vel[j * D + d]; }
    if (D == 3) for (int a = 0; a < 3; ++a) { wi[a] = s->ang_vel[i * 3 + a]; wj[a] = s->ang_vel[j * 3 + a]; }
    else { wi[2] = s->ang_vel[i]; wj[2] = s->ang_vel[j]; }
    REAL ci[3], cj[3];
    FN(cross)(wi, rci, ci);
    FN(cross)(wj, rcj, cj);
    REAL vrel[3], vt_vec[3], tt[3];
    for (int d = 0; d < 3; ++d) vrel[d] = (vi[d] + ci[d]) - (vj[d] + cj[d]);
    REAL vn = vrel[0] * n[0] + vrel[1] * n[1] + vrel[2] * n[2];
    for (int d = 0; d < 3; ++d) vt_vec[d] = vrel[d] - vn * n[d];
    REAL vt = FN(unit_and_norm)(vt_vec, tt);
    REAL Fn = kn * delta - gamma_n * vn;
    Fn = (Fn > (REAL)0 ? Fn : (REAL)0) * is_contact;
    REAL Ft = gamma_t * vt;
    Ft = (Ft < mu_eff * Fn ? Ft : mu_eff * Fn) * is_contact;
    for (int d = 0; d < 3; ++d) f[d] = Fn * n[d] - Ft * tt[d];
    REAL tq[3];
    FN(cross)(rci, f, tq);
    REAL mur = s->mu_r[mi] < s->mu_r[mj] ? s->mu_r[mi] : s->mu_r[mj];
    REAL R_eff = (Ri * Rj) / (Ri + Rj);
    REAL orel[3] = {wi[0] - wj[0], wi[1] - wj[1], wi[2] - wj[2]};
    REAL on2 = orel[0] * orel[0] + orel[1] * orel[1] + orel[2] * orel[2];
    REAL oinv = (REAL)1 / SQRT(on2 == (REAL)0 ? (REAL)1 : on2); /* linalg.unit (linalg.py:136-159) */
    REAL roll = mur * R_eff * Fn;
    for (int d = 0; d < 3; ++d) t[d] = tq[d] - roll * orel[d] * oinv;
  }
  (void)A;
}

/* valid_interaction_mask as called by the cell list (colliders/__init__.py:225-243,
 * cell_list.py:240-246): candidate's clump / bond row vs the owner's clump / index. */
static inline int FN(valid)(const FN(orc_sys) * s, int64_t i, int64_t j) {
  if (s->clump_id[j] == s->clump_id[i]) return 0;
  if (!s->interact)
    for (int w = 0; w < s->W; ++w)
      if (s->bond_id[j * s->W + w] == (INT)i) return 0;
  return 1;
}

/* DynamicCellList.compute_force (cell_list.py:434-464) over _traverse_pairs (:187-261):
 * particle i (original order), stencil rows in neighbor_mask order (unsorted coords,
 * periodic wrap + first-occurrence de-dup, :66-96), run from lower_bound while the
 * sorted hash matches, j = perm[k]; epilogue torque += cross(_pos_p_rot, F) (:461-462).
 * Requires orc_partition to have run on the same positions. */
void FN(orc_celllist_force)(FN(orc_sys) * s) {
  INT gd[3] = {1, 1, 1}, stride[3] = {0, 0, 0};
  int ovf;
  FN(grid_params)(s, gd, stride, &ovf);
  const int D = s->dim, A = D == 3 ? 3 : 1, M = s->M;
  const int64_t n = s->n;
#pragma omp parallel for schedule(dynamic, 512)
  for (int64_t i = 0; i < n; ++i) {
    INT cc[3] = {0, 0, 0};
    for (int d = 0; d < D; ++d) cc[d] = FN(cell_coord)(s, s->pos[i * D + d], d, gd[d]);
    INT nh[125];
    INT* nhp = M <= 125 ? nh : (INT*)malloc(sizeof(INT) * M);
    for (int m = 0; m < M; ++m) {
      UINT h = 0;
      int oob = 0;
      for (int d = 0; d < D; ++d) {
        INT nc = cc[d] + s->mask[m * D + d];
        if (s->periodic) nc -= gd[d] * FN(to_int)(FLOOR((REAL)nc / (REAL)gd[d]));
        else oob |= (nc < 0) | (nc >= gd[d]);
        h += (UINT)nc * (UINT)stride[d];
      }
      nhp[m] = oob ? (INT)-1 : (INT)h;
    }
    if (s->periodic) /* _dedup_stencil_hashes: later duplicates -> -1 */
      for (int m = M - 1; m > 0; --m)
        for (int m2 = 0; m2 < m; ++m2)
          if (nhp[m2] == nhp[m]) { nhp[m] = (INT)-1; break; }
    REAL F[3] = {0, 0, 0}, T[3] = {0, 0, 0};
    for (int m = 0; m < M; ++m) {
      const INT target = nhp[m];
      int64_t lo = 0, hi = n; /* searchsorted(side="left") */
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (s->sorted_hash[mid] < target) lo = mid + 1; else hi = mid;
      }
      for (int64_t k = lo; k < n && s->sorted_hash[k] == target; ++k) {
        const int64_t j = (int64_t)s->perm[k];
        if (!FN(valid)(s, i, j)) continue;
        REAL f[3], t[3];
        FN(pair_force)(s, i, j, f, t);
        for (int d = 0; d < 3; ++d) { F[d] += f[d]; T[d] += t[d]; }
      }
    }
    if (nhp != nh) free(nhp);
    const REAL* pr = s->pos_p_rot + i * D;
    if (D == 3) {
      for (int d = 0; d < 3; ++d) s->force[i * 3 + d] = F[d];
      s->torque[i * 3 + 0] = T[0] + (pr[1] * F[2] - pr[2] * F[1]);
      s->torque[i * 3 + 1] = T[1] + (pr[2] * F[0] - pr[0] * F[2]);
      s->torque[i * 3 + 2] = T[2] + (pr[0] * F[1] - pr[1] * F[0]);
    } else {
      s->force[i * 2] = F[0]; s->force[i * 2 + 1] = F[1];
      s->torque[i] = T[2] + (pr[0] * F[1] - pr[1] * F[0]);
    }
  }
  (void)A;
}

/* VelocityVerlet.step_before_force (jaxdem/integrators/velocity_verlet.py:57-61) */
void FN(orc_verlet_before)(FN(orc_sys) * s) {
  const int D = s->dim;
  const REAL dt = s->dt[0];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < s->n; ++i) {
    const REAL sc = dt * (REAL)0.5 / s->mass[i];
    const REAL fr = s->fixed[i] ? (REAL)0 : (REAL)1;
    for (int d = 0; d < D; ++d) {
      REAL v = s->vel[i * D + d] + s->force[i * D + d] * sc * fr;
      s->vel[i * D + d] = v;
      s->pos_c[i * D + d] = s->pos_c[i * D + d] + dt * v;
    }
  }
}

/* VelocityVerlet.step_after_force (velocity_verlet.py:92-95) */
void FN(orc_verlet_after)(FN(orc_sys) * s) {
  const int D = s->dim;
  const REAL dt = s->dt[0];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < s->n; ++i) {
    const REAL sc = dt * (REAL)0.5 / s->mass[i];
    const REAL fr = s->fixed[i] ? (REAL)0 : (REAL)1;
    for (int d = 0; d < D; ++d) s->vel[i * D + d] = s->vel[i * D + d] + s->force[i * D + d] * sc * fr;
  }
}

/* ForceManager.apply for sphere systems with empty external buffers
 * (jaxdem/forces/force_manager.py:359-423 with clump_id == arange(N), count == 1):
 * force += gravity * mass; torque unchanged. */
void FN(orc_force_manager_spheres)(FN(orc_sys) * s) {
  const int D = s->dim;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < s->n; ++i)
    for (int d = 0; d < D; ++d)
      s->force[i * D + d] = (s->force[i * D + d] + (REAL)0) + ((REAL)0 + s->gravity[d] * (s->mass[i] / (REAL)1));
}

/* n x _step_once (jaxdem/system.py:60-82) for sphere systems in a periodic (or fixed-box
 * non-periodic, domain.apply == no-op) domain, linear velocity Verlet, no rotation
 * integrator: BASELINE config 2. */
void FN(orc_step_spheres)(FN(orc_sys) * s, int64_t n_steps) {
  for (int64_t it = 0; it < n_steps; ++it) {
    for (int d = 0; d < s->dim; ++d) s->inv_box[d] = (REAL)1 / s->box[d]; /* system.py:69-74 */
    FN(orc_verlet_before)(s);
    FN(orc_partition)(s);
    FN(orc_celllist_force)(s);
    FN(orc_force_manager_spheres)(s);
    FN(orc_verlet_after)(s);
  }
}

#undef FN
#undef CAT
#undef CAT_
