// jaxdem_b200 — kernel launch macro: counts launches (jdb200_launch_count), turns launch
// failures into JDB200_ECUDA without synchronising, and (diagnostic mode only, see
// jdb200_timing_enable) brackets every launch with CUDA events on its own stream so
// bench.py can report per-kernel device time for the roofline.
#pragma once
#include <algorithm>
#include <atomic>
#include <type_traits>

#include "common.cuh"

namespace jdb {
extern std::atomic<unsigned long long> g_launches;
extern std::atomic<int> g_timing;
// records (name, start, stop) for the launch in flight; defined in abi.cu
void timing_begin(const char* name, cudaStream_t s);
void timing_end(cudaStream_t s);
}  // namespace jdb

#define JDB_LAUNCH(kernel, grid, block, stream, ...)                       \
  do {                                                                     \
    const bool jdb_t_ = jdb::g_timing.load(std::memory_order_relaxed) != 0; \
    if (jdb_t_) jdb::timing_begin(#kernel, (stream));                      \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                 \
    if (jdb_t_) jdb::timing_end((stream));                                 \
    jdb::g_launches.fetch_add(1, std::memory_order_relaxed);               \
    if (cudaPeekAtLastError() != cudaSuccess) return JDB200_ECUDA;         \
  } while (0)
