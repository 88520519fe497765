// Post-streaming operations on domain faces: NEE / NEBB / equilibrium (+ velocity, pressure,
// force-corrected wrappers), bounce-back / specular reflection, obstacle mask, characteristic BC.
// One thread per face cell; all in place on the post-streaming populations.  SURVEY.md 8a rows a10-a16.
#include "vsb_bc.cuh"
#include "vsb_internal.h"

namespace vsb {

constexpr int kBlock = 128;

template <int DIM, int LOC>
__global__ void k_face_bc(float* __restrict__ f, int n0, int n1, int n2, int wall_layer, int kind, int wrap, WallVals w) {
  using L = Lat<DIM>;
  using G = FaceGeom<DIM, LOC>;
  constexpr int Q = L::Q, D = L::D;
  const int n[3] = {n0, n1, n2};
  const long long nface = (long long)n[G::TA] * n[G::TB];
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nface) return;
  int idx[3];
  idx[G::TA] = (int)(k / n[G::TB]);
  idx[G::TB] = (int)(k % n[G::TB]);
  idx[G::AX] = wall_layer;
  const long long ncell = (long long)n0 * n1 * n2;
  const long long cw = ((long long)idx[0] * n1 + idx[1]) * n2 + idx[2];
  idx[G::AX] += G::SIGN;
  const long long cn = ((long long)idx[0] * n1 + idx[1]) * n2 + idx[2];

  float fw[Q], fn[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) fw[q] = f[q * ncell + cw];
  if (bc_needs_neighbor(kind, wrap)) {
#pragma unroll
    for (int q = 0; q < Q; ++q) fn[q] = f[q * ncell + cn];
  } else {
#pragma unroll
    for (int q = 0; q < Q; ++q) fn[q] = 0.f;
  }
  float uw[D], gw[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { uw[d] = wv(w.u[d], k); gw[d] = wv(w.g[d], k); }
  apply_face_bc<DIM, LOC>(fw, fn, kind, wrap, wv(w.rho, k), uw, gw);
#pragma unroll
  for (int q = 0; q < Q; ++q) f[q * ncell + cw] = fw[q];
}

// bounce-back / specular reflection: needs the PRE-streaming populations on the wall.
template <int DIM, int LOC>
__global__ void k_face_reflect(const float* __restrict__ f_pre, float* __restrict__ f, int n0, int n1, int n2,
                               int wall_layer, int specular, WallVals w) {
  using L = Lat<DIM>;
  using G = FaceGeom<DIM, LOC>;
  constexpr int Q = L::Q, D = L::D;
  const int n[3] = {n0, n1, n2};
  const long long nface = (long long)n[G::TA] * n[G::TB];
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nface) return;
  int idx[3];
  idx[G::TA] = (int)(k / n[G::TB]);
  idx[G::TB] = (int)(k % n[G::TB]);
  idx[G::AX] = wall_layer;
  const long long ncell = (long long)n0 * n1 * n2;
  const long long cw = ((long long)idx[0] * n1 + idx[1]) * n2 + idx[2];
  float fw[Q], pre[Q], uw[D];
#pragma unroll
  for (int q = 0; q < Q; ++q) { fw[q] = f[q * ncell + cw]; pre[q] = f_pre[q * ncell + cw]; }
#pragma unroll
  for (int d = 0; d < D; ++d) uw[d] = wv(w.u[d], k);
  apply_face_reflect<DIM, LOC>(fw, pre, specular, uw);
#pragma unroll
  for (int q = 0; q < Q; ++q) f[q * ncell + cw] = fw[q];
}

// obstacle_bounce_back: masked cells f_q <- f_opp(q)          (lbm/boundary/bb.py:110, lbm3d/boundary/bb.py:59)
template <int DIM>
__global__ void k_mask_bb(float* __restrict__ f, const uint8_t* __restrict__ mask, long long ncell) {
  using L = Lat<DIM>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell || !mask[i]) return;
  float fl[L::Q];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) fl[q] = f[q * ncell + i];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) f[q * ncell + i] = fl[L::opp(q)];
}

// boundary_characteristic                                    (lbm/boundary/cbc.py:14-53, lbm3d/boundary/cbc.py:16-51)
__global__ void k_characteristic(const float* __restrict__ rho, const float* __restrict__ u, float* __restrict__ rho_out,
                                 float* __restrict__ u_out, int n0, int n1, int n2, int a0, int dim, int ax, int sign) {
  const int n[3] = {n0, n1, n2};
  const int ta = (ax == 0) ? 1 : 0, tb = (ax == 2) ? 1 : 2;
  const long long nface = (long long)n[ta] * n[tb];
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nface) return;
  const long long ncell = (long long)n0 * n1 * n2;
  int idx[3];
  idx[ta] = (int)(k / n[tb]);
  idx[tb] = (int)(k % n[tb]);
  long long c[3];
  for (int j = 0; j < 3; ++j) {
    idx[ax] = (sign > 0) ? j : n[ax] - 1 - j;
    c[j] = ((long long)idx[0] * n1 + idx[1]) * n2 + idx[2];
  }
  const int nd = ax - a0;
  const float s = (float)sign;
  const float cs = 0.57735026918962576f;
  const float r1 = rho[c[0]], r2 = rho[c[1]], r3 = rho[c[2]];
  const float* un = u + (long long)nd * ncell;
  const float q1 = un[c[0]], q2 = un[c[1]], q3 = un[c[2]];
  const float coef = -0.5f * s;
  const float drho = coef * (3.f * r1 - 4.f * r2 + r3);
  const float dun = coef * (3.f * q1 - 4.f * q2 + q3);
  const float l_out = (q1 - s * cs) * (dun - s * cs / r1 * drho);
  rho_out[k] = r1 - 0.5f * r1 / cs * l_out;
  for (int d = 0; d < dim; ++d) u_out[d * nface + k] = (d == nd) ? q1 - 0.5f * l_out : u[(long long)d * ncell + c[0]];
}

template <int DIM, int LOC>
static int launch_face(const VsbPostOp& op, const float* f_pre, float* f, int n0, int n1, int n2, cudaStream_t s,
                       int r_begin, int r_end) {
  using G = FaceGeom<DIM, LOC>;
  const int n[3] = {n0, n1, n2};
  // the slowest real axis may be restricted to rows [r_begin, r_end) (ghost layers of a slab decomposition)
  const int lo = (G::AX == Lat<DIM>::A0) ? r_begin : 0;
  const int hi = (G::AX == Lat<DIM>::A0 && r_end > 0) ? r_end : n[G::AX];
  VSB_REQUIRE(hi - lo >= 2, "boundary op: the grid needs at least 2 layers along the face normal");
  const int wall_layer = (G::SIGN > 0) ? lo : hi - 1;
  const long long nface = (long long)n[G::TA] * n[G::TB];
  WallVals w;
  w.rho = op.rho;
  for (int d = 0; d < 3; ++d) { w.u[d] = op.u[d]; w.g[d] = op.g[d]; }
  const unsigned nb = blocks_for(nface, kBlock);
  if (op.kind == VSB_BC_BOUNCE_BACK || op.kind == VSB_BC_SPECULAR) {
    VSB_REQUIRE(f_pre != nullptr, "bounce-back / specular reflection need the pre-streaming populations");
    k_face_reflect<DIM, LOC><<<nb, kBlock, 0, s>>>(f_pre, f, n0, n1, n2, wall_layer, op.kind == VSB_BC_SPECULAR ? 1 : 0, w);
  } else {
    k_face_bc<DIM, LOC><<<nb, kBlock, 0, s>>>(f, n0, n1, n2, wall_layer, op.kind, op.wrap, w);
  }
  VSB_LAUNCH_CHECK("boundary op");
  return VSB_OK;
}

int launch_post_op(int dim, int n0, int n1, int n2, const VsbPostOp& op, const float* f_pre, float* f, cudaStream_t s,
                   int r_begin, int r_end) {
  if (op.kind == VSB_POST_MASK) {
    VSB_REQUIRE(op.mask != nullptr, "mask op without a mask");
    const long long ncell = (long long)n0 * n1 * n2;
    if (dim == 2) k_mask_bb<2><<<blocks_for(ncell, 256), 256, 0, s>>>(f, op.mask, ncell);
    else k_mask_bb<3><<<blocks_for(ncell, 256), 256, 0, s>>>(f, op.mask, ncell);
    VSB_LAUNCH_CHECK("obstacle_bounce_back");
    return VSB_OK;
  }
  VSB_REQUIRE(op.kind >= VSB_BC_NEE && op.kind <= VSB_BC_SPECULAR, "unknown post op kind %d", op.kind);
  VSB_REQUIRE(op.wrap >= VSB_WRAP_NONE && op.wrap <= VSB_WRAP_FORCE_CORRECTED, "unknown wrapper %d", op.wrap);
  VSB_REQUIRE(op.loc >= 0 && op.loc < 2 * dim, "loc %d is not a face of a %d-D grid", op.loc, dim);
  VSB_REQUIRE(op.rho.ptr != nullptr || op.rho.value != 0.f || op.wrap == VSB_WRAP_VELOCITY ||
                  op.kind == VSB_BC_BOUNCE_BACK || op.kind == VSB_BC_SPECULAR,
              "rho_wall = 0 (did you leave VsbPostOp.rho unset? the reference default is 1)");
  if (dim == 2) {
    switch (op.loc) {
      case 0: return launch_face<2, 0>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
      case 1: return launch_face<2, 1>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
      case 2: return launch_face<2, 2>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
      default: return launch_face<2, 3>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
    }
  }
  switch (op.loc) {
    case 0: return launch_face<3, 0>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
    case 1: return launch_face<3, 1>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
    case 2: return launch_face<3, 2>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
    case 3: return launch_face<3, 3>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
    case 4: return launch_face<3, 4>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
    default: return launch_face<3, 5>(op, f_pre, f, n0, n1, n2, s, r_begin, r_end);
  }
}

}  // namespace vsb

using namespace vsb;

extern "C" {

int vsb_post_op(const VsbGrid* grid, const VsbPostOp* op, const float* f_pre, float* f, vsb_stream_t stream) {
  VSB_REQUIRE(grid && op && f, "vsb_post_op: null argument");
  VSB_REQUIRE(grid->dim == 2 || grid->dim == 3, "dim must be 2 or 3, got %d", grid->dim);
  int n0, n1, n2;
  grid_axes(*grid, n0, n1, n2);
  VSB_REQUIRE(n0 > 0 && n1 > 0 && n2 > 0, "vsb_post_op: bad grid");
  return launch_post_op(grid->dim, n0, n1, n2, *op, f_pre, f, (cudaStream_t)stream, 0, 0);
}

int vsb_boundary_characteristic(const VsbGrid* grid, int loc, const float* rho, const float* u, float* rho_out,
                                float* u_out, vsb_stream_t stream) {
  VSB_REQUIRE(grid && rho && u && rho_out && u_out, "vsb_boundary_characteristic: null argument");
  VSB_REQUIRE(grid->dim == 2 || grid->dim == 3, "dim must be 2 or 3, got %d", grid->dim);
  VSB_REQUIRE(loc >= 0 && loc < 2 * grid->dim, "loc %d is not a face of a %d-D grid", loc, grid->dim);
  int n0, n1, n2;
  grid_axes(*grid, n0, n1, n2);
  const int a0 = 3 - grid->dim, ax = loc / 2 + a0, sign = (loc % 2 == 0) ? 1 : -1;
  const int n[3] = {n0, n1, n2};
  VSB_REQUIRE(n[ax] >= 3, "boundary_characteristic needs 3 layers along the normal");
  const long long nface = (long long)n0 * n1 * n2 / n[ax];
  k_characteristic<<<blocks_for(nface, kBlock), kBlock, 0, (cudaStream_t)stream>>>(rho, u, rho_out, u_out, n0, n1, n2, a0,
                                                                                 grid->dim, ax, sign);
  VSB_LAUNCH_CHECK("vsb_boundary_characteristic");
  return VSB_OK;
}

}  // extern "C"
