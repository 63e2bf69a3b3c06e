// jaxdem_b200 — single-pass device scan shared by the cell table and clump CSR builds.
#pragma once
#include "ctx.cuh"

namespace jdb {

// Single-pass exclusive scan of ints (decoupled look-back, 512 threads x 8 per tile):
// out[0..rows) = exclusive prefix of in[0..rows) (in == out allowed).  Tile ids are handed
// out dynamically so a tile's predecessors are always resident => the look-back cannot
// deadlock.  Warp 0 inspects 32 predecessor descriptors per round.  Descriptor = status
// (0 invalid, 1 aggregate, 2 inclusive prefix) << 32 | value.  Returns true if any element
// exceeded `limit`.
// The scanned array may be the concatenation of two windows of in / out: virtual elements
// [0, len_a) live at [a_lo, a_lo + len_a), the rest at [b_lo, ...); len_a is a multiple of
// kScanTile, so a tile never straddles the seam.  One window: len_a = rows, a_lo = 0.
__device__ __forceinline__ bool scan_tile(const int* __restrict__ in, int* __restrict__ out, long long rows,
                                          unsigned long long* __restrict__ ts, int* tile_counter,
                                          int limit, long long len_a = -1, long long a_lo = 0, long long b_lo = 0) {
  const int ntiles = (int)((rows + kScanTile - 1) / kScanTile);
  __shared__ int s_tile;
  __shared__ int s_warp[16];
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1);
  __syncthreads();
  const int tile = s_tile;
  if (tile >= ntiles) return false;
  const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * 8;
  {  // window of this tile
    const long long t0 = (long long)tile * kScanTile;
    const long long shift = (len_a < 0 || t0 < len_a) ? a_lo : b_lo - len_a;
    in += shift;
    out += shift;
  }
  int v[8];
  int sum = 0;
  bool too_many = false;
  const bool vec = base + 8 <= rows && ((((size_t)in) | ((size_t)out)) & 15) == 0;  // 16-byte loads / stores
  if (vec) {
    const int4 a = *reinterpret_cast<const int4*>(in + base);
    const int4 b = *reinterpret_cast<const int4*>(in + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (base + k < rows) ? in[base + k] : 0;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    too_many |= v[k] > limit;
    sum += v[k];
  }
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
    int w = lane < 16 ? s_warp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < 16) s_warp[lane] = wi - w;  // exclusive warp offsets
    const int aggregate = __shfl_sync(0xffffffffu, wi, 15);
    int prefix = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(&ts[0], (2ull << 32) | (unsigned)aggregate);
    } else {
      if (lane == 0) atomicExch(&ts[tile], (1ull << 32) | (unsigned)aggregate);  // publish aggregate
      int t0 = tile - 1;  // lane l looks at tile t0 - l
      while (true) {
        const int t = t0 - lane;
        unsigned long long d = 2ull << 32;  // tiles before 0: inclusive prefix 0
        if (t >= 0) {
          do {
            d = *((volatile unsigned long long*)&ts[t]);
          } while ((d >> 32) == 0);
        }
        const unsigned st = (unsigned)(d >> 32);
        const unsigned full = __ballot_sync(0xffffffffu, st == 2);
        const int first = __ffs(full) - 1;  // nearest predecessor holding an inclusive prefix
        int val = (first < 0 || lane <= first) ? (int)(unsigned)d : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        prefix += val;
        if (first >= 0) break;
        t0 -= 32;
      }
      if (lane == 0) atomicExch(&ts[tile], (2ull << 32) | (unsigned)(prefix + aggregate));
    }
    if (lane == 0) s_prefix = prefix;
  }
  __syncthreads();
  int run = s_prefix + s_warp[warp] + (incl - sum);
  if (vec) {
    int4 a, b;
    a.x = run; run += v[0]; a.y = run; run += v[1]; a.z = run; run += v[2]; a.w = run; run += v[3];
    b.x = run; run += v[4]; b.y = run; run += v[5]; b.z = run; run += v[6]; b.w = run;
    *reinterpret_cast<int4*>(out + base) = a;
    *reinterpret_cast<int4*>(out + base + 4) = b;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (base + k < rows) out[base + k] = run;
      run += v[k];
    }
  }
  return too_many;
}

}  // namespace jdb
