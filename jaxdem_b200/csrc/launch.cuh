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

namespace jdb {
// Programmatic dependent launch (sm_90+): every kernel of the library starts with
// pdl_prologue() — griddepcontrol.wait (all memory of the preceding kernel in the stream is
// visible after it) followed by griddepcontrol.launch_dependents — and is launched with
// programmatic stream serialization, so the NEXT kernel's blocks are scheduled while the
// last wave of this one drains instead of after a full launch round trip.  Nothing is read or
// written before the wait, so the stream semantics are unchanged.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace jdb

// ragged rows (Ctx.n_dev): a kernel that takes Ctx by value replaces its launch bound n by the live row count
#define JDB_LIVE_ROWS(c)                    \
  do {                                      \
    if ((c).n_dev) (c).n = *(c).n_dev;      \
  } while (0)

#define JDB_LAUNCH(kernel, grid, block, stream, ...)                       \
  do {                                                                     \
    const bool jdb_t_ = jdb::g_timing.load(std::memory_order_relaxed) != 0; \
    if (jdb_t_) jdb::timing_begin(#kernel, (stream));                      \
    jdb::launch_pdl(kernel, dim3(grid), dim3(block), (stream), __VA_ARGS__); \
    if (jdb_t_) jdb::timing_end((stream));                                 \
    jdb::g_launches.fetch_add(1, std::memory_order_relaxed);               \
    if (cudaPeekAtLastError() != cudaSuccess) return JDB200_ECUDA;         \
  } while (0)
