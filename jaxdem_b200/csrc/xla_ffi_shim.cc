// jaxdem_b200 — XLA FFI binding of the C ABI (include/jaxdem_b200.h) for jax.ffi.
//
// Compiled only where jaxlib's headers exist (`jax.ffi.include_dir()`); this build image has
// no JAX, so here the file compiles to nothing (the guard below) and HAS NOT BEEN EXECUTED —
// see INTEGRATION.md.  What runs today is the same C ABI through ctypes (jaxdem_b200/_lib.py).
//
//   g++ -O2 -std=c++17 -fPIC -shared xla_ffi_shim.cc -I../../include \
//       -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include -L.. -ljaxdem_b200 -o ../libjaxdem_b200_ffi.so
//
// Convention: every handler receives the State leaves and the System leaves as operands in
// the fixed order of jdb200_state / jdb200_system (a leaf the hook does not use may be passed
// as a zero-size buffer -> NULL), static configuration as attributes, and returns the leaves
// the hook updates as results bound with input_output_aliases (the C ABI works in place), plus
// one uint8 workspace result of jdb200_workspace_bytes(&params) bytes.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define JDB200_HAVE_XLA_FFI 1
#endif
#endif

#ifdef JDB200_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <cstdint>

#include "jaxdem_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

constexpr int kStateLeaves = 16;   // order of jdb200_state
constexpr int kSystemLeaves = 21;  // order of jdb200_system (time, step_count last; may be empty)

inline void* ptr_or_null(const ffi::AnyBuffer& b) { return b.element_count() == 0 ? nullptr : b.untyped_data(); }

// operands [0, 16) -> jdb200_state, [16, 37) -> jdb200_system
ffi::Error fill(ffi::RemainingArgs args, jdb200_state* st, jdb200_system* sy, jdb200_params* p) {
  if (args.size() < kStateLeaves + kSystemLeaves)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "jaxdem_b200: expected 37 operands (State + System leaves)");
  void** s = reinterpret_cast<void**>(st);
  for (int i = 0; i < kStateLeaves; ++i) {
    auto b = args.get<ffi::AnyBuffer>(i);
    if (b.has_error()) return b.error();
    s[i] = ptr_or_null(b.value());
  }
  void** y = reinterpret_cast<void**>(sy);
  for (int i = 0; i < kSystemLeaves; ++i) {
    auto b = args.get<ffi::AnyBuffer>(kStateLeaves + i);
    if (b.has_error()) return b.error();
    y[i] = ptr_or_null(b.value());
  }
  // shapes come from pos_c: (N, D) or, under vmap, (B, N, D)
  auto pos = args.get<ffi::AnyBuffer>(0).value();
  const auto dims = pos.dimensions();
  p->batch = dims.size() == 3 ? dims[0] : 1;
  p->n = dims[dims.size() - 2];
  p->dim = static_cast<int32_t>(dims[dims.size() - 1]);
  p->dtype = pos.element_type() == ffi::F32 ? JDB200_F32 : JDB200_F64;
  auto bond = args.get<ffi::AnyBuffer>(13).value();  // bond_id (.., N, W)
  p->bond_width = bond.element_count() ? static_cast<int32_t>(bond.dimensions().back()) : 0;
  auto mask = args.get<ffi::AnyBuffer>(kStateLeaves + 6).value();  // neighbor_mask (.., M, D)
  p->stencil_m = mask.element_count() ? static_cast<int32_t>(mask.dimensions()[mask.dimensions().size() - 2]) : 0;
  auto young = args.get<ffi::AnyBuffer>(kStateLeaves + 13).value();  // mat_young (.., Mt)
  p->n_materials = young.element_count() ? static_cast<int32_t>(young.dimensions().back()) : 0;
  return ffi::Error::Success();
}

struct Static {  // trace-time-static attributes shared by all handlers
  int64_t domain, law, collider, lin, rot, max_cells, grid_mode, clumps, max_neighbors;
};

void apply(const Static& a, jdb200_params* p) {
  p->domain = static_cast<int32_t>(a.domain);
  p->law = static_cast<int32_t>(a.law);
  p->collider = static_cast<int32_t>(a.collider);
  p->linear_integrator = static_cast<int32_t>(a.lin);
  p->rotation_integrator = static_cast<int32_t>(a.rot);
  p->max_cells = a.max_cells;
  p->grid_mode = static_cast<int32_t>(a.grid_mode);
  p->clumps = static_cast<int32_t>(a.clumps);
  p->max_neighbors = static_cast<int32_t>(a.max_neighbors);
}

ffi::Error status(int rc, const char* what) {
  if (rc == JDB200_OK) return ffi::Error::Success();
  return ffi::Error(rc == JDB200_ECUDA ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument, what);
}

using Hook4 = int (*)(void*, const jdb200_params*, const jdb200_state*, const jdb200_system*);
using Hook6 = int (*)(void*, const jdb200_params*, const jdb200_state*, const jdb200_system*, void*, size_t);

// last result = workspace; the other results alias operands (nothing to do with them here)
template <Hook6 FN>
ffi::Error Call6(cudaStream_t stream, int64_t domain, int64_t law, int64_t collider, int64_t lin, int64_t rot,
                 int64_t max_cells, int64_t grid_mode, int64_t clumps, ffi::RemainingArgs args,
                 ffi::RemainingRets rets) {
  jdb200_params p{};
  jdb200_state st{};
  jdb200_system sy{};
  if (auto e = fill(args, &st, &sy, &p); e.failure()) return e;
  apply(Static{domain, law, collider, lin, rot, max_cells, grid_mode, clumps, 0}, &p);
  auto ws = rets.get<ffi::AnyBuffer>(rets.size() - 1);
  if (ws.has_error()) return ws.error();
  return status(FN(stream, &p, &st, &sy, ws.value()->untyped_data(), ws.value()->size_bytes()), "jaxdem_b200 hook failed");
}

template <Hook4 FN>
ffi::Error Call4(cudaStream_t stream, int64_t domain, int64_t law, int64_t collider, int64_t lin, int64_t rot,
                 int64_t max_cells, int64_t grid_mode, int64_t clumps, ffi::RemainingArgs args,
                 ffi::RemainingRets rets) {
  jdb200_params p{};
  jdb200_state st{};
  jdb200_system sy{};
  if (auto e = fill(args, &st, &sy, &p); e.failure()) return e;
  apply(Static{domain, law, collider, lin, rot, max_cells, grid_mode, clumps, 0}, &p);
  (void)rets;
  return status(FN(stream, &p, &st, &sy), "jaxdem_b200 hook failed");
}

ffi::Error SystemStep(cudaStream_t stream, int64_t domain, int64_t law, int64_t collider, int64_t lin, int64_t rot,
                      int64_t max_cells, int64_t grid_mode, int64_t clumps, int64_t n_steps,
                      ffi::RemainingArgs args, ffi::RemainingRets rets) {
  jdb200_params p{};
  jdb200_state st{};
  jdb200_system sy{};
  if (auto e = fill(args, &st, &sy, &p); e.failure()) return e;
  apply(Static{domain, law, collider, lin, rot, max_cells, grid_mode, clumps, 0}, &p);
  auto ws = rets.get<ffi::AnyBuffer>(rets.size() - 1);
  if (ws.has_error()) return ws.error();
  return status(jdb200_system_step(stream, &p, &st, &sy, ws.value()->untyped_data(), ws.value()->size_bytes(), n_steps),
                "jdb200_system_step failed");
}

#define JDB200_COMMON_ATTRS()                                                                        \
  .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                           \
      .Attr<int64_t>("domain").Attr<int64_t>("law").Attr<int64_t>("collider")                         \
      .Attr<int64_t>("linear_integrator").Attr<int64_t>("rotation_integrator")                        \
      .Attr<int64_t>("max_cells").Attr<int64_t>("grid_mode").Attr<int64_t>("clumps")

}  // namespace

// stream-ordered, synchronisation-free => legal in XLA command buffers (CUDA graphs)
#define JDB200_HANDLER6(NAME, FN)                                                                     \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(NAME, Call6<FN>,                                                      \
                                ffi::Ffi::Bind() JDB200_COMMON_ATTRS().RemainingArgs().RemainingRets(), \
                                {ffi::Traits::kCmdBufferCompatible})
#define JDB200_HANDLER4(NAME, FN)                                                                     \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(NAME, Call4<FN>,                                                      \
                                ffi::Ffi::Bind() JDB200_COMMON_ATTRS().RemainingArgs().RemainingRets(), \
                                {ffi::Traits::kCmdBufferCompatible})

JDB200_HANDLER6(JdbCellListForce, jdb200_celllist_compute_force);          // Collider.compute_force
JDB200_HANDLER6(JdbNaiveForce, jdb200_naive_compute_force);
JDB200_HANDLER6(JdbForceManagerApply, jdb200_force_manager_apply);         // ForceManager.apply
JDB200_HANDLER6(JdbDomainApply, jdb200_domain_apply);                      // Domain.apply
JDB200_HANDLER6(JdbForceStepAfter, jdb200_celllist_force_step_after);      // collider + manager + after-kick
JDB200_HANDLER4(JdbLinearBefore, jdb200_linear_step_before_force);         // LinearIntegrator.step_before_force
JDB200_HANDLER4(JdbLinearAfter, jdb200_linear_step_after_force);
JDB200_HANDLER4(JdbRotationBefore, jdb200_rotation_step_before_force);     // RotationIntegrator.step_before_force
JDB200_HANDLER4(JdbRotationAfter, jdb200_rotation_step_after_force);

XLA_FFI_DEFINE_HANDLER_SYMBOL(JdbSystemStep, SystemStep,
                              ffi::Ffi::Bind() JDB200_COMMON_ATTRS().Attr<int64_t>("n_steps").RemainingArgs().RemainingRets(),
                              {ffi::Traits::kCmdBufferCompatible});

#endif  // JDB200_HAVE_XLA_FFI
