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
// Handlers in this file: streaming, moments, equilibrium, collisions, forcing, face operations with scalar AND
// array-valued wall data, bounce-back / specular reflection, obstacle mask, characteristic boundary, the IB functions
// (delta kernels, stencil, interpolate, spread, multi_direct_forcing on a window), the device Newmark update, the
// post.py diagnostics, the multigrid transfers and the fused step (21 of the ABI's entry points; the rest are
// orchestration calls -- host-ODE drivers, halo exchange, sharded chain -- that a jax program reaches through the
// Stepper / SlabStepper classes, not through jit-ed code).  The host-ODE path a custom call may use is
// vsb_enqueue_host_ode: it never blocks the calling thread.
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

// ------------------------------------------------------------------------------------------------------------------
// Array-valued wall data (reference lbm/boundary/_helpers.py:53-77: every wall quantity may be a scalar or a
// face-shaped array).  The face arrays arrive as extra operands; an operand with zero elements means "use the scalar".
static ffi::Error PostOpArray(cudaStream_t s, F32 f_in, F32 rho_w, F32 ux_w, F32 uy_w, F32 uz_w, F32 gx_w, F32 gy_w,
                              F32 gz_w, RF32 f, int32_t kind, int32_t wrap, int32_t loc, float rho,
                              ffi::Span<const float> u, ffi::Span<const float> g) {
  VsbGrid grid = GridOf(f_in.dimensions());
  VsbPostOp op{};
  op.kind = kind; op.wrap = wrap; op.loc = loc;
  auto wall = [](F32& a, float v) { return VsbWallValue{a.element_count() ? a.typed_data() : nullptr, v}; };
  op.rho = wall(rho_w, rho);
  F32* uw[3] = {&ux_w, &uy_w, &uz_w};
  F32* gw[3] = {&gx_w, &gy_w, &gz_w};
  for (int i = 0; i < 3; ++i) {
    op.u[i] = wall(*uw[i], i < (int)u.size() ? u[i] : 0.f);
    op.g[i] = wall(*gw[i], i < (int)g.size() ? g[i] : 0.f);
  }
  if (f->typed_data() != f_in.typed_data())
    cudaMemcpyAsync(f->typed_data(), f_in.typed_data(), f_in.size_bytes(), cudaMemcpyDeviceToDevice, s);
  return Fail(vsb_post_op(&grid, &op, nullptr, f->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbPostOpArray, PostOpArray,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Ret<F32>().Attr<int32_t>("kind").Attr<int32_t>("wrap")
                                  .Attr<int32_t>("loc").Attr<float>("rho").Attr<ffi::Span<const float>>("u")
                                  .Attr<ffi::Span<const float>>("g"));

// boundary_bounce_back / boundary_specular_reflection (need the pre-streaming populations) and obstacle_bounce_back
static ffi::Error PostOpReflect(cudaStream_t s, F32 f_pre, F32 f_in, RF32 f, int32_t kind, int32_t loc,
                                ffi::Span<const float> u) {
  VsbGrid grid = GridOf(f_in.dimensions());
  VsbPostOp op{};
  op.kind = kind; op.loc = loc;
  op.rho.value = 1.f;
  for (size_t i = 0; i < u.size() && i < 3; ++i) op.u[i].value = u[i];
  if (f->typed_data() != f_in.typed_data())
    cudaMemcpyAsync(f->typed_data(), f_in.typed_data(), f_in.size_bytes(), cudaMemcpyDeviceToDevice, s);
  return Fail(vsb_post_op(&grid, &op, f_pre.typed_data(), f->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbPostOpReflect, PostOpReflect,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Attr<int32_t>("kind")
                                  .Attr<int32_t>("loc").Attr<ffi::Span<const float>>("u"));

static ffi::Error ObstacleMask(cudaStream_t s, F32 f_in, ffi::Buffer<ffi::U8> mask, RF32 f) {
  VsbGrid grid = GridOf(f_in.dimensions());
  VsbPostOp op{};
  op.kind = VSB_POST_MASK;
  op.mask = mask.typed_data();
  if (f->typed_data() != f_in.typed_data())
    cudaMemcpyAsync(f->typed_data(), f_in.typed_data(), f_in.size_bytes(), cudaMemcpyDeviceToDevice, s);
  return Fail(vsb_post_op(&grid, &op, nullptr, f->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbObstacleMask, ObstacleMask,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<ffi::Buffer<ffi::U8>>().Ret<F32>());

// boundary_characteristic: (rho, u) -> face-shaped (rho_out, u_out)          lbm/boundary/cbc.py:14-53
static ffi::Error Characteristic(cudaStream_t s, F32 rho, F32 u, RF32 rho_out, RF32 u_out, int32_t loc) {
  auto d = u.dimensions();                                // (dim, NX, NY[, NZ])
  VsbGrid grid{(int)d[0], (int)d[1], (int)d[2], d.size() == 4 ? (int)d[3] : 1};
  return Fail(vsb_boundary_characteristic(&grid, loc, rho.typed_data(), u.typed_data(), rho_out->typed_data(),
                                          u_out->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbCharacteristic, Characteristic,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Attr<int32_t>("loc"));

// ---- immersed boundary: the per-function handlers (ib/kernels.py, ib/stencil.py) ...
using I32 = ffi::Buffer<ffi::S32>;
using RI32 = ffi::ResultBuffer<ffi::S32>;

static ffi::Error IbDelta(cudaStream_t s, F32 r, RF32 out, int32_t kind) {
  return Fail(vsb_ib_delta(kind, (int64_t)r.element_count(), r.typed_data(), out->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbIbDelta, IbDelta, ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>().Attr<int32_t>("kind"));

// get_ib_stencil: coords (M, dim) -> weights, indices (M, (2 radius)^dim)
static ffi::Error IbStencil(cudaStream_t s, F32 coords, RF32 weights, RI32 indices, int32_t kind, int32_t radius,
                            int32_t ny, int32_t nz) {
  auto d = coords.dimensions();
  return Fail(vsb_ib_stencil((int)d[1], kind, radius, d[0], coords.typed_data(), ny, nz, weights->typed_data(),
                             indices->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbIbStencil, IbStencil,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Attr<int32_t>("kind").Attr<int32_t>("radius").Attr<int32_t>("ny").Attr<int32_t>("nz"));

// interpolate: grid (C, *S), weights / indices (M, NS) -> (M, C)
static ffi::Error IbInterpolate(cudaStream_t s, F32 grid, F32 w, I32 idx, RF32 out) {
  auto gd = grid.dimensions();
  auto wd = w.dimensions();
  return Fail(vsb_ib_interpolate((int)gd[0], CellsOf(gd), grid.typed_data(), wd[0], (int)wd[1], w.typed_data(),
                                 idx.typed_data(), out->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbIbInterpolate, IbInterpolate,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<I32>().Ret<F32>());

// spread: grid_in + scatter(values) -> grid_out (aliased with grid_in when XLA donates it)
static ffi::Error IbSpread(cudaStream_t s, F32 values, F32 grid_in, F32 w, I32 idx, RF32 grid) {
  auto gd = grid_in.dimensions();
  auto wd = w.dimensions();
  if (grid->typed_data() != grid_in.typed_data())
    cudaMemcpyAsync(grid->typed_data(), grid_in.typed_data(), grid_in.size_bytes(), cudaMemcpyDeviceToDevice, s);
  return Fail(vsb_ib_spread((int)gd[0], CellsOf(gd), grid->typed_data(), wd[0], (int)wd[1], values.typed_data(),
                            w.typed_data(), idx.typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbIbSpread, IbSpread,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<I32>().Ret<F32>());

// ... and multi_direct_forcing (ib/mdf.py:10-64) with the stencil formed on the fly from the marker coordinates.
// Operands: the post-collision state the step reads, marker coordinates (M, dim), optional target velocities and
// per-marker ds (zero-sized operand = absent), the device body state (23 words) or a zero-sized operand for a fixed
// body.  Results: force window g (wnx, wny[, wnz], 2 | 4), marker force (M, dim), marker velocity (M, dim) and a
// zero-initialised scratch the caller threads through as XLA temporaries: the (n_iter - 1) work fields and the next
// step's buffers (XLA has no notion of "cleared one step ahead", so this handler clears what it needs itself).
static ffi::Error Mdf(cudaStream_t s, F32 state, F32 markers, F32 u_target, F32 ds, F32 body, RF32 g_win, RF32 marker_force,
                      RF32 marker_u, RF32 scratch, int32_t kernel, int32_t n_iter, float ds_value,
                      ffi::Span<const int32_t> win_origin, ffi::Span<const int32_t> win_size) {
  VsbStepArgs a{};
  a.grid = GridOf(state.dimensions());
  a.do_stream = 1; a.do_collide = 1;
  a.f_in = state.typed_data();
  a.f_out = scratch->typed_data();            // never written by vsb_ib_mdf; only has to differ from f_in
  VsbMdfArgs m{};
  auto md = markers.dimensions();
  m.dim = (int)md[1]; m.delta_kind = kernel; m.n_iter = n_iter; m.n_markers = md[0];
  int64_t wcells = 1;
  for (int i = 0; i < m.dim; ++i) { m.win_origin0[i] = win_origin[i]; m.win_size[i] = win_size[i]; wcells *= win_size[i]; }
  const int64_t field = wcells * (m.dim == 2 ? 2 : 4);
  m.markers0 = markers.typed_data();
  m.u_target = u_target.element_count() ? u_target.typed_data() : nullptr;
  m.ds_ptr = ds.element_count() ? ds.typed_data() : nullptr;
  m.ds_value = ds_value;
  m.g_win = g_win->typed_data();
  m.scratch = scratch->typed_data();                         // (n_iter - 1) work fields ...
  m.g_win_next = scratch->typed_data() + (int64_t)(n_iter - 1) * field;       // ... and a dummy "next step" set
  m.scratch_next = m.g_win_next + field;
  m.marker_u = marker_u->typed_data(); m.marker_force = marker_force->typed_data();
  m.body = body.element_count() ? reinterpret_cast<VsbBodyState*>(const_cast<float*>(body.typed_data())) : nullptr;
  m.chain_mode = 3;                                          // one launch per iteration: no barrier word to carry
  cudaMemsetAsync(g_win->typed_data(), 0, field * sizeof(float), s);
  cudaMemsetAsync(scratch->typed_data(), 0, (int64_t)(n_iter - 1) * field * sizeof(float), s);
  return Fail(vsb_ib_mdf(&a, &m, nullptr, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbMdf, Mdf,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Attr<int32_t>("kernel").Attr<int32_t>("n_iter")
                                  .Attr<float>("ds").Attr<ffi::Span<const int32_t>>("win_origin")
                                  .Attr<ffi::Span<const int32_t>>("win_size"));

// dyn.newmark on the device (dyn.py:5-51): in place on the 23-word body state
static ffi::Error BodyNewmark(cudaStream_t s, F32 body_in, RF32 body, int32_t n_dof, int32_t follow, int32_t parity, float m,
                              float k, float c, float added_mass, ffi::Span<const float> origin0,
                              ffi::Span<const int32_t> grid_size, ffi::Span<const int32_t> win_size) {
  VsbBodyParams bp{};
  bp.n_dof = n_dof; bp.follow = follow; bp.m = m; bp.k = k; bp.c = c; bp.added_mass = added_mass;
  for (size_t i = 0; i < 3; ++i) {
    bp.origin0[i] = i < origin0.size() ? origin0[i] : 0.f;
    bp.grid_size[i] = i < grid_size.size() ? grid_size[i] : 1;
    bp.win_size[i] = i < win_size.size() ? win_size[i] : 1;
  }
  if (body->typed_data() != body_in.typed_data())
    cudaMemcpyAsync(body->typed_data(), body_in.typed_data(), body_in.size_bytes(), cudaMemcpyDeviceToDevice, s);
  return Fail(vsb_body_newmark(reinterpret_cast<VsbBodyState*>(body->typed_data()), &bp, parity, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbBodyNewmark, BodyNewmark,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>().Attr<int32_t>("n_dof").Attr<int32_t>("follow")
                                  .Attr<int32_t>("parity").Attr<float>("m").Attr<float>("k").Attr<float>("c")
                                  .Attr<float>("added_mass").Attr<ffi::Span<const float>>("origin0")
                                  .Attr<ffi::Span<const int32_t>>("grid_size").Attr<ffi::Span<const int32_t>>("win_size"));

// post.py diagnostics and multigrid.py transfers
static ffi::Error PostField(cudaStream_t s, F32 in, RF32 out, int32_t kind, float param, ffi::Span<const int32_t> shape) {
  VsbGrid grid{(int)shape.size(), shape[0], shape[1], shape.size() == 3 ? shape[2] : 1};
  return Fail(vsb_post_field(&grid, kind, in.typed_data(), param, out->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbPostField, PostField,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Ret<F32>().Attr<int32_t>("kind").Attr<float>("param")
                                  .Attr<ffi::Span<const int32_t>>("shape"));

static ffi::Error MgTransfer(cudaStream_t s, F32 src, F32 dst_in, RF32 dst, int32_t dir, int32_t to_fine) {
  auto sd = src.dimensions();
  auto dd = dst_in.dimensions();
  if (dst->typed_data() != dst_in.typed_data())
    cudaMemcpyAsync(dst->typed_data(), dst_in.typed_data(), dst_in.size_bytes(), cudaMemcpyDeviceToDevice, s);
  return Fail(to_fine ? vsb_mg_coarse_to_fine((int)sd[1], (int)sd[2], src.typed_data(), (int)dd[1], (int)dd[2], dst->typed_data(), dir, s)
                      : vsb_mg_fine_to_coarse((int)sd[1], (int)sd[2], src.typed_data(), (int)dd[1], (int)dd[2], dst->typed_data(), dir, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(VsbMgTransfer, MgTransfer,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>().Attr<int32_t>("dir").Attr<int32_t>("to_fine"));
