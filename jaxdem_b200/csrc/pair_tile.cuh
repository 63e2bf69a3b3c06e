// jaxdem_b200 — K4 main path: shared-memory tile kernel (included by pair.cu).
//
// Same job as k_pair_flat (one thread per sorted slot, candidates of the 3^D stencil, contact
// list, force law, fused epilogue) for the CTAs whose 128 slots lie in the INTERIOR of the grid
// in y and z.  There the x-fastest linear hash (_partition.py:91-93) makes every stencil row a
// LINEAR shift of the CTA's own key range: row (dy, dz) of all 128 particles is the slot window
//     [ cell_start[h_lo - 1 + off], cell_start[h_hi + 2 + off] ),   off = dy*gx + dz*gx*gy,
// one contiguous piece of the sorted (x, y, z, rad) array.  Warp 0 issues one TMA bulk copy
// (cp.async.bulk global -> shared, mbarrier completion) per window and one per window of the
// cell-start table, so the candidate loop touches shared memory only: no global-load latency,
// no 64-bit address arithmetic, a third of the instructions.  Cells that wrap around the
// periodic box in x (needed by the particles of the first / last cell of an x-row) are staged
// next to the windows by the other warps.
//
// A CTA that does not qualify (touches the y/z faces of the grid, spans too many cells, or
// overflows the staging buffers) appends itself to a work list served by k_pair_flat right
// after.  Which kernel serves a particle is a function of the partition only, candidates are
// visited and contacts summed in a fixed order => bitwise repeatable, no atomics on floats.
#pragma once
// (included from inside namespace jdb)

template <typename F, int D>
struct TileCfg {
  static constexpr int kThreads = kPairBlock;
  static constexpr int kRows = D == 3 ? 9 : 3;
  static constexpr int kSpan = kThreads + kThreads / 4;  // cells between the CTA's first and last key
  static constexpr int kW = kSpan + 8;              // cell-start words staged per window (span + 3, + 3 alignment slack, /4)
  static constexpr int kRec = (int)sizeof(Vec4<F>);
  static constexpr int kExCap = kThreads / 2;       // staged records reserved for the wrapped cells
  static constexpr int kCap = kRows * (kThreads + 12) + kExCap;  // staged records (windows + wrapped cells), < 4096
  static constexpr int kMaxRows = 8;                // x-rows covered by the CTA's key range
  static constexpr int kEx = kMaxRows * 2;          // wrapped-cell runs: one per (x-row, side)
  static constexpr int kClRows = 8;                 // contact-list rows; a full list is evaluated and reused
  static constexpr int kOffPos = 0;
  static constexpr int kOffCst = kCap * kRec;                       // int [kRows][kW]
  static constexpr int kOffCl = kOffCst + kRows * kW * 4;           // u16 [kClRows][kThreads]
  static constexpr int kOffEx = kOffCl + kClRows * kThreads * 2;    // int [2][kEx]: first record, count; int [kExCap]: global slot per wrapped record
  static constexpr int kOffWin = kOffEx + (2 * kEx + kExCap) * 4;   // int [2][kRows]: delta (record = slot + delta), shift
  static constexpr int kOffMisc = kOffWin + 2 * kRows * 4 + 4;      // mbarrier (8 B aligned), counters
  static constexpr int kBytes = ((kOffMisc + 7) / 8) * 8 + 24;
};

__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_arrive(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ int lds_i32(unsigned a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_i32(unsigned a, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ Vec4<float> lds_vec4(unsigned a, float) {
  Vec4<float> v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ Vec4<double> lds_vec4(unsigned a, double) {
  Vec4<double> v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(a + 16));
  return v;
}
template <typename F>
__device__ __forceinline__ void sts_vec4(unsigned a, const Vec4<F>& v);
template <>
__device__ __forceinline__ void sts_vec4<float>(unsigned a, const Vec4<float>& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <>
__device__ __forceinline__ void sts_vec4<double>(unsigned a, const Vec4<double>& v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a + 16), "d"(v.z), "d"(v.w) : "memory");
}

__device__ __forceinline__ unsigned sign_word(float s) { return __float_as_uint(s); }   // top bit = sign
__device__ __forceinline__ unsigned sign_word(double s) { return (unsigned)__double2hiint(s); }
__device__ __forceinline__ float real_from_bits(unsigned long long b, float) { return __uint_as_float((unsigned)b); }
__device__ __forceinline__ double real_from_bits(unsigned long long b, double) { return __longlong_as_double((long long)b); }

// |rij|^2 through Domain._displacement (periodic.py:75-79), for candidates across the boundary
template <typename F, int D>
__device__ __forceinline__ void tile_image(const LawCtx<F>& lc, F* rij) {
  using T = RT<F>;
#pragma unroll
  for (int d = 0; d < D; ++d) rij[d] = T::sub(rij[d], T::mul(lc.box[d], T::rint(T::mul(rij[d], lc.inv_box[d]))));
}

// One contact.  ent (u16): staged record of the candidate (bits 0..11) and where it came from:
// window r (bits 12..15; 15 = a wrapped cell).  Candidates arrive through a CONSERVATIVE
// test, so the exact one (|rij| < Ri + Rj, after the minimum image) is made here; a pair that
// fails it contributes exactly zero with every law.
template <typename F, int LAW, int D, bool PERIODIC, bool SIMPLE>
__device__ __forceinline__ void tile_contact(const Ctx<F>& c, const LawCtx<F>& lc, size_t off, const Body<F>& a,
                                             int idx, int clump, bool interact, F hbmin2, unsigned sbase,
                                             unsigned self_rec, unsigned ent, F* f, F* t) {
  using T = RT<F>;
  using Cfg = TileCfg<F, D>;
  constexpr bool CS = LAW == JDB200_LAW_CUNDALLSTRACK;
  const unsigned rec = ent & 0xfffu;
  if (SIMPLE && rec == self_rec) return;  // clump_id == arange(N): only the particle itself is excluded
  const Vec4<F> qc = lds_vec4(sbase + rec * (unsigned)Cfg::kRec, F(0));
  F rij[3] = {T::sub(a.x, qc.x), T::sub(a.y, qc.y), D == 3 ? T::sub(a.z, qc.z) : F(0)};
  F d2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
  if (PERIODIC && !(d2 < hbmin2)) {
    tile_image<F, D>(lc, rij);
    d2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
  }
  const F rs = a.r + qc.w;
  if (!(d2 < rs * rs * F(1.00001))) return;
  int kcur = 0;
  if (!SIMPLE || CS || c.nmat > 1) {  // global slot of the candidate
    if ((ent >> 12) == 15u) kcur = lds_i32(sbase + Cfg::kOffEx + (2 * Cfg::kEx + rec - (Cfg::kCap - Cfg::kExCap)) * 4);
    else kcur = (int)rec - lds_i32(sbase + Cfg::kOffWin + (ent >> 12) * 4);
  }
  if (!SIMPLE) {
    const int sc = c.sclump[off + kcur];
    if (!pair_valid(c, off, idx, clump, sc, kcur, interact)) return;
  }
  Body<F> bj;
  bj.x = qc.x; bj.y = qc.y; bj.z = qc.z; bj.r = qc.w;
  bj.mat = (c.nmat > 1) ? c.smat[off + kcur] : 0;
  if (CS) {
    const Vec4<F> v = c.svel[off + kcur];
    const Vec4<F> w = c.sang[off + kcur];
    bj.vx = v.x; bj.vy = v.y; bj.vz = v.z; bj.m = v.w;
    bj.wx = w.x; bj.wy = w.y; bj.wz = w.z;
  }
  F ff[3], tt[3];
  pair_force_rij<F, LAW>(lc, a, bj, rij, ff, tt);
  f[0] += ff[0]; f[1] += ff[1]; f[2] += ff[2];
  if (CS) { t[0] += tt[0]; t[1] += tt[1]; t[2] += tt[2]; }
}
template <typename F, int LAW, int D, bool PERIODIC, bool SIMPLE, int EPI>
__device__ __forceinline__ void pair_tile_body(const Ctx<F>& c, int b, int k, int h, int h_lo, int rho_lo,
                                               const GridInfo<typename RT<F>::I>& g, int with_torque,
                                               unsigned sbase, const LawCtx<F>& lc, const StepUniform<F, D>& su,
                                               int idx, const Vec4<F>& vm) {
  using T = RT<F>;
  using Cfg = TileCfg<F, D>;
  constexpr bool CS = LAW == JDB200_LAW_CUNDALLSTRACK;
  constexpr unsigned kT = Cfg::kThreads;
  constexpr unsigned kRec = Cfg::kRec;
  constexpr int kRows = Cfg::kRows;
  constexpr int kCenter = D == 3 ? 4 : 1;  // window of the particle's own row
  const unsigned tid = threadIdx.x;
  const size_t off = (size_t)b * c.n;
  const int clump = SIMPLE ? 0 : (c.sclump[off + k] & 0x7fffffff);
  const bool interact = c.interact && c.interact[b];

  const unsigned self_rec = (unsigned)(k + lds_i32(sbase + Cfg::kOffWin + kCenter * 4));
  const Vec4<F> av = lds_vec4(sbase + self_rec * kRec, F(0));
  Body<F> a;
  a.x = av.x; a.y = av.y; a.z = av.z; a.r = av.w;
  a.mat = (c.nmat > 1) ? c.smat[off + k] : 0;
  if (CS) {
    const Vec4<F> v = c.svel[off + k];
    const Vec4<F> w = c.sang[off + k];
    a.vx = v.x; a.vy = v.y; a.vz = v.z; a.m = v.w;
    a.wx = w.x; a.wy = w.y; a.wz = w.z;
  }
  // |rij|^2 below hbmin2 => every |rij_d| < box_d / 2 => the minimum-image term is exactly zero
  F hbmin2 = F(0);
  if (PERIODIC) {
    F m = lc.box[0];
#pragma unroll
    for (int d = 1; d < D; ++d) m = T::fmin(m, lc.box[d]);
    m *= F(0.499999);
    hbmin2 = m * m;
  }
  // Conservative candidate test, one subtraction deep: a pair can only touch if |rij|^2 < tq
  // = (Ri + Rmax)^2 (1 + 1e-5); periodic: or if the raw |rij|^2 >= hbmin2 (other image).
  //   periodic:      hit <=> wq - | |rij|^2 - cq | < 0,  cq, wq = centre / half width of [tq, hbmin2)
  //   non-periodic:  hit <=> |rij|^2 - tq < 0
  const F rsm = a.r + real_from_bits(g.rmax_bits, F(0));
  const F tq = rsm * rsm * F(1.00001);
  const F cq = F(0.5) * (hbmin2 + tq), wq = F(0.5) * (hbmin2 - tq) * F(0.999999);

  F f[3] = {0, 0, 0}, t[3] = {0, 0, 0};
  const unsigned cl0 = sbase + Cfg::kOffCl + tid * 2;
  const unsigned clend = cl0 + Cfg::kClRows * kT * 2;
  unsigned cla = cl0;

  // ---- A + B: row by row, the x-run of (up to) three cells from the staged cell-start window.
  // The first four candidates of a run (runs are ~3 long) are tested without a loop and leave a
  // hit bit each; the rest go one at a time onto the contact list ----
  const int gx = (int)g.gd[0];
  int cx;
  const int rho = div_fix(h, gx, __frcp_rn(__int2float_rn(gx)), cx);
  const int x1 = max(cx - 1, 0);
  const int n1 = min(cx + 1, gx - 1) - x1 + 1;
  const unsigned hrel4 = (unsigned)(h - cx + x1 - (h_lo - 1)) * 4;
  unsigned m0 = 0, m1 = 0;  // hit bit of slot u of row r: bit 31 - (4 r + u) of m0 (rows 0..7) / m1 (row 8)
#pragma unroll 1
  for (int r = 0; r < kRows; ++r) {
    const int delta = lds_i32(sbase + Cfg::kOffWin + r * 4);
    const unsigned wa = sbase + Cfg::kOffCst + (unsigned)lds_i32(sbase + Cfg::kOffWin + (kRows + r) * 4) + hrel4;
    const int s = lds_i32(wa), e = lds_i32(wa + n1 * 4);
    const unsigned ja = (unsigned)(s + delta);  // first staged record of the run
    const int len = e - s;
    const unsigned pa = sbase + ja * kRec;
    unsigned nib = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const Vec4<F> q = lds_vec4(pa + u * kRec, F(0));  // past the run: some other staged record, masked below
      const F rx = a.x - q.x, ry = a.y - q.y, rz = D == 3 ? a.z - q.z : F(0);
      const F d2 = rx * rx + ry * ry + rz * rz;
      const F sg = PERIODIC ? wq - T::abs(d2 - cq) : d2 - tq;
      nib = __funnelshift_l(sign_word(sg), nib, 1);
    }
    nib &= 0xf0u >> min(len, 4);  // slots u < len
    if (r < 8) m0 = (m0 << 4) | (nib & 0xfu);
    else m1 = nib << 28;
    for (int j = 4; j < len; ++j) {
      const Vec4<F> q = lds_vec4(pa + j * kRec, F(0));
      const F rx = a.x - q.x, ry = a.y - q.y, rz = D == 3 ? a.z - q.z : F(0);
      const F d2 = rx * rx + ry * ry + rz * rz;
      const F sg = PERIODIC ? wq - T::abs(d2 - cq) : d2 - tq;
      if (sg < F(0)) {
        const unsigned ent = (ja + j) | ((unsigned)r << 12);
        if (cla != clend) {
          sts_u16(cla, ent);
          cla += kT * 2;
        } else {
          tile_contact<F, LAW, D, PERIODIC, SIMPLE>(c, lc, off, a, idx, clump, interact, hbmin2, sbase, self_rec,
                                                        ent, f, t);
        }
      }
    }
  }
  if (kRows < 8) m0 <<= 32 - 4 * kRows;
  // ---- B': particles of the first / last cell of an x-row: the cells wrapped around the box ----
  if (PERIODIC && (cx == 0 || cx == gx - 1)) {
    const unsigned e = (unsigned)((rho - rho_lo) * 2 + (cx == 0 ? 0 : 1));
    const int cnt = lds_i32(sbase + Cfg::kOffEx + (Cfg::kEx + e) * 4);
    unsigned ja = (unsigned)lds_i32(sbase + Cfg::kOffEx + e * 4);
    for (int j = 0; j < cnt; ++j, ++ja) {
      const Vec4<F> q = lds_vec4(sbase + ja * kRec, F(0));
      F rij[3] = {T::sub(a.x, q.x), T::sub(a.y, q.y), D == 3 ? T::sub(a.z, q.z) : F(0)};
      tile_image<F, D>(lc, rij);
      const F d2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
      if (d2 < tq) {
        const unsigned ent = ja | 0xf000u;
        if (cla != clend) {
          sts_u16(cla, ent);
          cla += kT * 2;
        } else {
          tile_contact<F, LAW, D, PERIODIC, SIMPLE>(c, lc, off, a, idx, clump, interact, hbmin2, sbase, self_rec,
                                                        ent, f, t);
        }
      }
    }
  }
  // ---- C: force law on the hits: bits in (row, slot) order, then the list ----
  unsigned ca = cl0;
  while (true) {
    unsigned ent;
    if (m0 | m1) {
      const bool lo = m0 != 0;
      const unsigned m = lo ? m0 : m1;
      const int i = __clz(m);
      const unsigned rest = m & ~(0x80000000u >> i);
      if (lo) m0 = rest; else m1 = rest;
      const int r = (i >> 2) + (lo ? 0 : 8);
      const unsigned wa = sbase + Cfg::kOffCst + (unsigned)lds_i32(sbase + Cfg::kOffWin + (kRows + r) * 4) + hrel4;
      ent = ((unsigned)(lds_i32(wa) + lds_i32(sbase + Cfg::kOffWin + r * 4)) + (unsigned)(i & 3)) | ((unsigned)r << 12);
    } else if (ca != cla) {
      ent = lds_u16(ca);
      ca += kT * 2;
    } else {
      break;
    }
    tile_contact<F, LAW, D, PERIODIC, SIMPLE>(c, lc, off, a, idx, clump, interact, hbmin2, sbase, self_rec, ent, f, t);
  }
  if (EPI == 1) fused_sphere_epilogue<F, D>(c, b, g, vm, idx, f, t, with_torque != 0, su);
  else store_force_torque<F, D>(c, off + idx, f, t, g.any_ppr != 0, with_torque != 0);
}

template <typename F, int LAW, int D, bool PERIODIC, int EPI>
__global__ void __launch_bounds__(TileCfg<F, D>::kThreads, (sizeof(F) == 4 ? 896 : 512) / TileCfg<F, D>::kThreads) k_pair_tile(Ctx<F> c, int with_torque) {
  using I = typename RT<F>::I;
  using Cfg = TileCfg<F, D>;
  constexpr int kRows = Cfg::kRows;
  constexpr unsigned kRec = Cfg::kRec;
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned mbar = sbase + ((Cfg::kOffMisc + 7) / 8) * 8;
  const unsigned s_exused = mbar + 8, s_fail = mbar + 12, s_issued = mbar + 16;
  const unsigned tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sts_i32(s_exused, 0);
    sts_i32(s_fail, 0);
    sts_i32(s_issued, 0);
  }
  pdl_prologue();
  const int b = blockIdx.y;
  // the tile's keys do not depend on the grid description: request them first
  const int n = (int)c.n;
  const int k0 = blockIdx.x * Cfg::kThreads;
  if (k0 >= n) return;
  const int kend = min(k0 + Cfg::kThreads, n);
  const size_t off = (size_t)b * c.n;
  const int k = k0 + (int)tid;
  const bool live = k < kend;
  const int* tk = c.tmp_key + off;
  const int h = tk[live ? k : kend - 1];  // hash of slot k (k_scatter; the in-cell fix-up keeps cells in place)
  const int h_lo = tk[k0], h_hi = tk[kend - 1];
  const GridInfo<I> g = c.gi[b];
  const bool mine = flat_walk_ok(g);
  if (blockIdx.x == 0 && tid == 0 && c.overflow) {
    if (mine) c.overflow[b] = (uint8_t)g.hash_overflow;
    else if (c.grid_mode == JDB200_GRID_DENSE) c.overflow[b] = 1;  // nobody else serves this system
  }
  if (!mine) return;
  // ---- does the tile qualify?  (CTA-uniform) ----
  const int gx = (int)g.gd[0], gy = (int)g.gd[1], gz = (int)g.gd[2];
  const int span = h_hi - h_lo + 1;
  int rem;
  const float rgx = __frcp_rn(__int2float_rn(gx));
  const int rho_lo = div_fix(h_lo, gx, rgx, rem);
  const int rho_hi = div_fix(h_hi, gx, rgx, rem);
  bool ok = span <= Cfg::kSpan && rho_hi - rho_lo < Cfg::kMaxRows;
  int off_min;
  if (D == 3) {
    int cy_lo, cy_hi;
    const float rgy = __frcp_rn(__int2float_rn(gy));
    const int cz_lo = div_fix(rho_lo, gy, rgy, cy_lo);
    const int cz_hi = div_fix(rho_hi, gy, rgy, cy_hi);
    ok = ok && cz_lo == cz_hi && cz_lo >= 1 && cz_lo <= gz - 2 && cy_lo >= 1 && cy_hi <= gy - 2;
    off_min = -gx - gx * gy;
  } else {
    ok = ok && rho_lo >= 1 && rho_hi <= gy - 2;
    off_min = -gx;
  }
  // the staged cell-start windows must lie inside the table (rows [0, bound])
  ok = ok && h_lo - 1 + off_min >= 0 && (long long)h_hi - off_min + 8 <= g.bound;
  if (!ok) {
    if (tid == 0) c.wl[(size_t)b * c.pair_blocks + atomicAdd(&c.wl_count[b], 1)] = blockIdx.x;
    return;
  }
  __syncthreads();  // mbarrier and counters initialised
  const int* cst = c.cell_start + (size_t)b * c.cell_stride;
  const Vec4<F>* sp = c.spos + off;
  if (tid < 32) {
    // ---- warp 0: one bulk copy per window of records and per window of cell starts ----
    const int r = (int)tid;
    int ws = 0, cnt = 0, a_r = 0, len = 0, shift = 0;
    if (r < kRows) {
      const int off_r = D == 3 ? ((r % 3) - 1) * gx + ((r / 3) - 1) * gx * gy : (r - 1) * gx;
      const int c_lo = h_lo - 1 + off_r;
      a_r = c_lo & ~3;
      shift = c_lo - a_r;
      len = (span + 3 + shift + 3) & ~3;
      ws = cst[c_lo];
      cnt = cst[c_lo + span + 2] - ws;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)tid >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, kRows - 1);
    const int base = incl - cnt;
    if (total > Cfg::kCap - Cfg::kExCap) {
      if (tid == 0) sts_i32(s_fail, 1);
    } else {
      int bytes = r < kRows ? cnt * (int)kRec + len * 4 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
      if (tid == 0) {
        mbar_expect_tx_arrive(mbar, (unsigned)bytes);
        sts_i32(s_issued, 1);
      }
      __syncwarp();
      if (r < kRows) {
        if (cnt > 0) bulk_g2s(sbase + Cfg::kOffPos + (unsigned)base * kRec, sp + ws, (unsigned)cnt * kRec, mbar);
        bulk_g2s(sbase + Cfg::kOffCst + (unsigned)(r * Cfg::kW) * 4, cst + a_r, (unsigned)len * 4, mbar);
        sts_i32(sbase + Cfg::kOffWin + r * 4, base - ws);
        sts_i32(sbase + Cfg::kOffWin + (kRows + r) * 4, (r * Cfg::kW + shift) * 4);
      }
    }
  } else if (PERIODIC) {
    // ---- other warps: the cells wrapped around the box in x, for every x-row the tile covers.
    // One warp per (x-row, side): lane r stages the wrapped cell of stencil row r; the cells of
    // one (x-row, side) are laid out back to back in row order, so the particles that need them
    // walk ONE run ----
    const int ncombo = (rho_hi - rho_lo + 1) * 2;
    const int lane = (int)tid & 31;
    for (int e = (int)(tid >> 5) - 1; e < ncombo; e += Cfg::kThreads / 32 - 1) {
      const int side = e & 1, ri = e >> 1, r = lane;
      const int row0 = (rho_lo + ri) * gx;  // key of the row's first cell
      // side 0: particles of cell (0, row) need cell (gx - 1, row + off); side 1: the mirror case
      const int owner = side == 0 ? row0 : row0 + gx - 1;
      int s = 0, cnt = 0;
      if (r < kRows && owner >= h_lo && owner <= h_hi) {
        const int off_r = D == 3 ? ((r % 3) - 1) * gx + ((r / 3) - 1) * gx * gy : (r - 1) * gx;
        const int cw = (side == 0 ? row0 + gx - 1 : row0) + off_r;
        s = cst[cw];
        cnt = cst[cw + 1] - s;
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total = __shfl_sync(0xffffffffu, incl, kRows - 1);
      int base = 0;
      if (lane == 0 && total > 0) base = atomicAdd((int*)(smem + (s_exused - sbase)), total);
      base = __shfl_sync(0xffffffffu, base, 0);
      int rec0 = 0, tot = total;
      if (base + total > Cfg::kExCap) {
        if (lane == 0) sts_i32(s_fail, 1);
        tot = 0;
      } else if (total > 0) {
        rec0 = Cfg::kCap - base - total;
        const int rec = rec0 + incl - cnt;
        for (int j = 0; j < cnt; ++j) {
          sts_vec4<F>(sbase + Cfg::kOffPos + (unsigned)(rec + j) * kRec, sp[s + j]);
          sts_i32(sbase + Cfg::kOffEx + (2 * Cfg::kEx + rec + j - (Cfg::kCap - Cfg::kExCap)) * 4, s + j);
        }
      }
      if (lane == 0) {
        sts_i32(sbase + Cfg::kOffEx + e * 4, rec0);
        sts_i32(sbase + Cfg::kOffEx + (Cfg::kEx + e) * 4, tot);
      }
    }
  }
  // per-system and per-particle values, requested before the wait so that they arrive during it
  const LawCtx<F> lc = make_law_ctx(c, b);
  StepUniform<F, D> su;
  if (EPI == 1) su = load_step_uniform<F, D>(c, b);
  const int kk = live ? k : kend - 1;
  const int idx = c.perm[off + kk];
  Vec4<F> vm = Vec4<F>{0, 0, 0, 0};
  if (EPI == 1) vm = c.svel[off + kk];
  __syncthreads();
  const bool issued = lds_i32(s_issued) != 0;
  if (issued) mbar_wait(mbar, 0);  // also on the way out: a CTA must not retire with copies in flight
  if (lds_i32(s_fail)) {
    if (tid == 0) c.wl[(size_t)b * c.pair_blocks + atomicAdd(&c.wl_count[b], 1)] = blockIdx.x;
    return;
  }
  if (!live) return;
  if (!c.clumps && !g.any_bond)
    pair_tile_body<F, LAW, D, PERIODIC, true, EPI>(c, b, k, h, h_lo, rho_lo, g, with_torque, sbase, lc, su, idx, vm);
  else
    pair_tile_body<F, LAW, D, PERIODIC, false, EPI>(c, b, k, h, h_lo, rho_lo, g, with_torque, sbase, lc, su, idx, vm);
}

// launch with the shared-memory carve-out the 7-CTAs-per-SM residency needs
#define JDB_LAUNCH_TILE(kernel, grid, block, smem_bytes, stream, ...)                                          \
  do {                                                                                                        \
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (smem_bytes));                  \
    JDB_LAUNCH_SMEM(kernel, grid, block, smem_bytes, stream, __VA_ARGS__);                                    \
  } while (0)

