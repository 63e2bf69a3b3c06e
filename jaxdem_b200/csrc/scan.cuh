// jaxdem_b200 — single-pass device scan shared by the cell table and clump CSR builds.
#pragma once
#include "ctx.cuh"

namespace jdb {

// Single-pass exclusive scan of ints (decoupled look-back, 512 threads x 8 per tile),
// in place over data[0..rows).  Tile ids are handed out dynamically so a tile's
// predecessors are always resident => the look-back spin cannot deadlock.
// Returns true if any element exceeded `limit`.
__device__ __forceinline__ bool scan_tile(int* __restrict__ data, long long rows,
                                          unsigned long long* __restrict__ ts, int* tile_counter,
                                          int limit) {
  const int ntiles = (int)((rows + kScanTile - 1) / kScanTile);
  __shared__ int s_tile;
  __shared__ int s_warp[16];
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1);
  __syncthreads();
  const int tile = s_tile;
  if (tile >= ntiles) return false;
  const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * 8;
  int v[8];
  int sum = 0;
  bool too_many = false;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = (base + k < rows) ? data[base + k] : 0;
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
    if (lane == 15) {
      const int aggregate = wi;
      int prefix = 0;
      if (tile == 0) {
        atomicExch(&ts[0], (2ull << 32) | (unsigned)aggregate);
      } else {
        atomicExch(&ts[tile], (1ull << 32) | (unsigned)aggregate);  // publish aggregate
        int t = tile - 1;
        while (true) {  // look back for the exclusive prefix of this tile
          const unsigned long long s = *((volatile unsigned long long*)&ts[t]);
          const unsigned st = (unsigned)(s >> 32);
          if (st == 0) continue;
          prefix += (int)(unsigned)s;
          if (st == 2) break;
          --t;
        }
        atomicExch(&ts[tile], (2ull << 32) | (unsigned)(prefix + aggregate));
      }
      s_prefix = prefix;
    }
  }
  __syncthreads();
  int run = s_prefix + s_warp[warp] + (incl - sum);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (base + k < rows) data[base + k] = run;
    run += v[k];
  }
  return too_many;
}

}  // namespace jdb
