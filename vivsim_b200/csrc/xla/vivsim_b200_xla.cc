// XLA FFI custom-call handlers over the C ABI (include/vivsim_b200.h) -- NOT BUILT IN THIS IMAGE.
//
// north_star asks for the kernels to be reachable from jax as XLA FFI custom calls.  jax / jaxlib and the
// xla/ffi/api headers are absent here (no wheel, no network), so this translation unit is neither compiled by
// vivsim_b200/_build.py nor tested; it documents exactly what a maintainer with jaxlib would build:
//
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I../../../include \
//       vivsim_b200_xla.cc -L../.. -lvivsim_b200 -o libvivsim_b200_xla.so
//
// and register from Python:
//   jax.ffi.register_ffi_target("vsb_streaming", jax.ffi.pycapsule(lib.VsbStreaming), platform="CUDA")
//
// Conventions: XLA owns every buffer and passes the stream; handlers never allocate or synchronise; static
// configuration (omega, collision / forcing kind, loc, window) arrives as attributes; errors come back as
// ffi::Error with the text of vsb_last_error().
#include <cstdint>

#include "vivsim_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using Stream = ffi::PlatformStream<cudaStream_t>;

static ffi::Error Fail(int rc) {
  return rc == VSB_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, vsb_last_error());
}
static VsbGrid GridOf(ffi::Span<const int64_t> d) {   // (Q, NX, NY[, NZ])
  return VsbGrid{(int)d.size() - 1, (int)d[1], (int)d[2], d.size() == 4 ? (int)d[3] : 1};
}
static int64_t CellsOf(ffi::Span<const int64_t> d) {
  int64_t n = 1;
  for (size_t i = 1; i < d.size(); ++i) n *= d[i];
  return n;
}
static int DimOfQ(int64_t q) { return q == 9 ? 2 : 3; }

// lbm.streaming / lbm3d.streaming
static ffi::Error Streaming(cudaStream_t s, F32 f, RF32 out) {
  VsbGrid g = GridOf(f.dimensions());
  return Fail(vsb_streaming(&g, f.typed_data(), out->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbStreaming, Streaming, ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>());

// get_macroscopic: f -> (rho, u)
static ffi::Error Macroscopic(cudaStream_t s, F32 f, RF32 rho, RF32 u) {
  auto d = f.dimensions();
  return Fail(vsb_macroscopic(DimOfQ(d[0]), CellsOf(d), f.typed_data(), rho->typed_data(), u->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbMacroscopic, Macroscopic, ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>().Ret<F32>());

// get_equilibrium: (rho, u) -> feq
static ffi::Error Equilibrium(cudaStream_t s, F32 rho, F32 u, RF32 feq) {
  auto d = u.dimensions();
  return Fail(vsb_equilibrium((int)d[0], CellsOf(d), rho.typed_data(), u.typed_data(), feq->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbEquilibrium, Equilibrium, ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>());

// collision_bgk / _kbc / _reg (kind attribute); collision_mrt passes its Q x Q operator as a host-resident attribute
static ffi::Error Collision(cudaStream_t s, F32 f, F32 feq, RF32 out, int32_t kind, float omega,
                            ffi::Span<const float> op) {
  auto d = f.dimensions();
  return Fail(vsb_collision(DimOfQ(d[0]), CellsOf(d), kind, omega, op.size() ? op.begin() : nullptr, f.typed_data(),
                            feq.typed_data(), out->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbCollision, Collision,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Attr<int32_t>("kind")
                                  .Attr<float>("omega").Attr<ffi::Span<const float>>("op"));

// forcing_edm / forcing_guo_bgk / forcing_guo_mrt
static ffi::Error Forcing(cudaStream_t s, F32 f, F32 g, F32 u, RF32 out, int32_t kind, float omega,
                          ffi::Span<const float> fop) {
  auto d = f.dimensions();
  return Fail(vsb_forcing(DimOfQ(d[0]), CellsOf(d), kind, omega, fop.size() ? fop.begin() : nullptr, f.typed_data(),
                          g.typed_data(), u.typed_data(), out->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbForcing, Forcing,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>()
                                  .Attr<int32_t>("kind").Attr<float>("omega").Attr<ffi::Span<const float>>("fop"));

// boundary_{nee,nebb,equilibrium} with scalar wall values (input/output aliased: in place on XLA's donated buffer).
// Array-valued wall data would be extra Arg<F32> operands filling VsbWallValue::ptr.
static ffi::Error PostOpScalar(cudaStream_t s, F32 f_in, RF32 f, int32_t kind, int32_t wrap, int32_t loc, float rho,
                               ffi::Span<const float> u, ffi::Span<const float> g) {
  VsbGrid grid = GridOf(f_in.dimensions());
  VsbPostOp op{};
  op.kind = kind; op.wrap = wrap; op.loc = loc;
  op.rho.value = rho;
  for (size_t i = 0; i < u.size() && i < 3; ++i) op.u[i].value = u[i];
  for (size_t i = 0; i < g.size() && i < 3; ++i) op.g[i].value = g[i];
  if (f->typed_data() != f_in.typed_data())   // not aliased: copy first (cudaMemcpyAsync on XLA's stream)
    cudaMemcpyAsync(f->typed_data(), f_in.typed_data(), f_in.size_bytes(), cudaMemcpyDeviceToDevice, s);
  return Fail(vsb_post_op(&grid, &op, nullptr, f->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbPostOpScalar, PostOpScalar,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>().Attr<int32_t>("kind").Attr<int32_t>("wrap")
                                  .Attr<int32_t>("loc").Attr<float>("rho").Attr<ffi::Span<const float>>("u")
                                  .Attr<ffi::Span<const float>>("g"));

// The fused step on the post-collision state (periodic or with a force window; face operations would be passed as a
// serialized VsbPostOp array attribute).
static ffi::Error Step(cudaStream_t s, F32 s_in, F32 g_win, RF32 s_out, int32_t collision, int32_t forcing, float omega,
                       ffi::Span<const int32_t> win_origin, ffi::Span<const int32_t> win_size) {
  VsbStepArgs a{};
  a.grid = GridOf(s_in.dimensions());
  a.collision = collision; a.forcing = forcing; a.omega = omega;
  a.do_stream = 1; a.do_collide = 1;
  a.f_in = s_in.typed_data(); a.f_out = s_out->typed_data();
  a.g_win = g_win.element_count() ? g_win.typed_data() : nullptr;
  for (int i = 0; i < a.grid.dim && i < (int)win_size.size(); ++i) { a.win_origin[i] = win_origin[i]; a.win_size[i] = win_size[i]; }
  return Fail(vsb_step(&a, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbStep, Step,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Attr<int32_t>("collision")
                                  .Attr<int32_t>("forcing").Attr<float>("omega")
                                  .Attr<ffi::Span<const int32_t>>("win_origin").Attr<ffi::Span<const int32_t>>("win_size"));
