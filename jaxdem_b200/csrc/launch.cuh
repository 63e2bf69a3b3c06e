// jaxdem_b200 — kernel launch macro: counts launches (jdb200_launch_count) and turns
// launch failures into JDB200_ECUDA without synchronising.
#pragma once
#include <algorithm>
#include <atomic>
#include <type_traits>

#include "common.cuh"

namespace jdb {
extern std::atomic<unsigned long long> g_launches;
}

#define JDB_LAUNCH(kernel, grid, block, stream, ...)                       \
  do {                                                                     \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                 \
    jdb::g_launches.fetch_add(1, std::memory_order_relaxed);               \
    if (cudaPeekAtLastError() != cudaSuccess) return JDB200_ECUDA;         \
  } while (0)
