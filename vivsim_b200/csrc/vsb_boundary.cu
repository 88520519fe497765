// Post-streaming operations on domain faces: NEE / NEBB / equilibrium (+ velocity, pressure,
// force-corrected wrappers), bounce-back / specular reflection, obstacle mask, characteristic BC.
// One thread per face cell; all in place on the post-streaming populations.  SURVEY.md 8a rows a10-a16.
#include "vsb_common.cuh"
#include "vsb_internal.h"

namespace vsb {

constexpr int kBlock = 128;

struct WallVals {
  VsbWallValue rho, u[3], g[3];
};

__device__ __forceinline__ float wv(const VsbWallValue& v, long long k) { return v.ptr ? v.ptr[k] : v.value; }

template <int DIM>
__host__ __device__ constexpr int find_dir(int c0, int c1, int c2) {
  using L = Lat<DIM>;
  for (int q = 0; q < L::Q; ++q)
    if (L::c(q, 0) == c0 && L::c(q, 1) == c1 && L::c(q, 2) == c2) return q;
  return -1;
}

// Geometry of one face in array-axis terms.  LOC as VSB_LOC_*.
template <int DIM, int LOC> struct FaceGeom {
  using L = Lat<DIM>;
  static constexpr int AX = LOC / 2 + L::A0;          // array axis normal to the face
  static constexpr int SIGN = (LOC % 2 == 0) ? 1 : -1; // inward normal direction along AX
  static constexpr int ND = AX - L::A0;               // velocity component normal to the face
  static constexpr int TA = (AX == 0) ? 1 : 0;        // remaining array axes, ascending
  static constexpr int TB = (AX == 2) ? 1 : 2;
  __host__ __device__ static constexpr int cn(int q) { return L::c(q, AX) * SIGN; }  // >0: enters the fluid
};

// sum_{zero} f + 2 sum_{out} f      (lbm/boundary/_helpers.py:135-145, lbm3d/boundary/_helpers.py:35-43)
template <int DIM, int LOC>
__device__ __forceinline__ float rho_numerator(const float (&fw)[Lat<DIM>::Q]) {
  using G = FaceGeom<DIM, LOC>;
  float zero = 0.f, out = 0.f;
#pragma unroll
  for (int q = 0; q < Lat<DIM>::Q; ++q) {
    if (G::cn(q) == 0) zero += fw[q];
    if (G::cn(q) < 0) out += fw[q];
  }
  return zero + 2.0f * out;
}

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
  const bool need_nb = (kind == VSB_BC_NEE) || (wrap == VSB_WRAP_PRESSURE);
  if (need_nb) {
#pragma unroll
    for (int q = 0; q < Q; ++q) fn[q] = f[q * ncell + cn];
  }

  float rho_w = wv(w.rho, k), uw[D];
#pragma unroll
  for (int d = 0; d < D; ++d) uw[d] = wv(w.u[d], k);

  if (wrap == VSB_WRAP_VELOCITY) {
    // rho_w = numerator / (1 - u_n)                (lbm/boundary/_helpers.py:80-95, lbm3d/.../_helpers.py:58-63)
    rho_w = rho_numerator<DIM, LOC>(fw) / (1.0f - (float)G::SIGN * uw[G::ND]);
  } else if (wrap == VSB_WRAP_PRESSURE) {
    // u_n from rho_w; tangential velocity from the adjacent fluid layer
    // (lbm/boundary/_helpers.py:98-132 ; lbm3d/boundary/_helpers.py:66-78)
    const float un = (float)G::SIGN * (1.0f - rho_numerator<DIM, LOC>(fw) / rho_w);
    float rho_nb, u_nb[D];
    moments<DIM>(fn, rho_nb, u_nb);
#pragma unroll
    for (int d = 0; d < D; ++d) uw[d] = u_nb[d];
    uw[G::ND] = un;
  } else if (wrap == VSB_WRAP_FORCE_CORRECTED) {
    // u_w -= g_w / (2 rho_w)                        (lbm/boundary/_helpers.py:156-177)
#pragma unroll
    for (int d = 0; d < D; ++d) uw[d] -= wv(w.g[d], k) * 0.5f / rho_w;
  }

  if (kind == VSB_BC_EQUILIBRIUM) {            // lbm/boundary/eq.py:45-56
    float fe[Q];
    equilibrium<DIM>(rho_w, uw, fe);
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q * ncell + cw] = fe[q];
  } else if (kind == VSB_BC_NEE) {             // lbm/boundary/nee.py:43-60
    float fe[Q], fen[Q], rho_nb, u_nb[D];
    equilibrium<DIM>(rho_w, uw, fe);
    moments<DIM>(fn, rho_nb, u_nb);
    equilibrium<DIM>(rho_nb, u_nb, fen);
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q * ncell + cw] = fe[q] + (fn[q] - fen[q]);
  } else if (kind == VSB_BC_NEBB) {
    if constexpr (DIM == 2) {
      // Zou/He with transverse correction          (lbm/boundary/nebb.py:41-58)
      constexpr int TAX = (G::AX == 1) ? 2 : 1;  // tangential array axis
      constexpr int TD = TAX - L::A0;
      constexpr int cN[3] = {0, G::AX == 1 ? G::SIGN : 0, G::AX == 2 ? G::SIGN : 0};
      constexpr int cT[3] = {0, TAX == 1 ? 1 : 0, TAX == 2 ? 1 : 0};
      constexpr int in0 = find_dir<2>(0, cN[1], cN[2]);
      constexpr int in1 = find_dir<2>(0, cN[1] + G::SIGN * cT[1], cN[2] + G::SIGN * cT[2]);
      constexpr int in2 = find_dir<2>(0, cN[1] - G::SIGN * cT[1], cN[2] - G::SIGN * cT[2]);
      constexpr int t0 = find_dir<2>(0, cT[1], cT[2]), t1 = find_dir<2>(0, -cT[1], -cT[2]);
      const float un = (float)G::SIGN * uw[G::ND], ut = (float)G::SIGN * uw[TD];
      const float shear = 0.5f * (fw[t0] - fw[t1]) * (float)G::SIGN;
      const float normal = (1.0f / 6.0f) * un * rho_w;
      const float tang = 0.5f * ut * rho_w;
      f[in0 * ncell + cw] = fw[L::opp(in0)] + (2.0f / 3.0f) * un * rho_w;
      f[in1 * ncell + cw] = fw[L::opp(in1)] - shear + normal + tang;
      f[in2 * ncell + cw] = fw[L::opp(in2)] + shear + normal - tang;
    } else {
      // f_in = f_opp(in) + feq_in - feq_opp(in)    (lbm3d/boundary/nebb.py:16-32)
      float fe[Q];
      equilibrium<DIM>(rho_w, uw, fe);
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (G::cn(q) > 0) f[q * ncell + cw] = fw[L::opp(q)] + fe[q] - fe[L::opp(q)];
    }
  }
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
  float uw[D];
#pragma unroll
  for (int d = 0; d < D; ++d) uw[d] = wv(w.u[d], k);
  if constexpr (DIM == 2) {
    // in_k <- pre[out_k] + {2/3 un, 1/6 (un+ut), 1/6 (un-ut)}; rho = 1 assumed   (lbm/boundary/bb.py:43-53,82-95)
    constexpr int TAX = (G::AX == 1) ? 2 : 1;
    constexpr int TD = TAX - L::A0;
    constexpr int cN[3] = {0, G::AX == 1 ? G::SIGN : 0, G::AX == 2 ? G::SIGN : 0};
    constexpr int cT[3] = {0, TAX == 1 ? 1 : 0, TAX == 2 ? 1 : 0};
    constexpr int in0 = find_dir<2>(0, cN[1], cN[2]);
    constexpr int in1 = find_dir<2>(0, cN[1] + G::SIGN * cT[1], cN[2] + G::SIGN * cT[2]);
    constexpr int in2 = find_dir<2>(0, cN[1] - G::SIGN * cT[1], cN[2] - G::SIGN * cT[2]);
    const float un = (float)G::SIGN * uw[G::ND], ut = (float)G::SIGN * uw[TD];
    const float v0 = f_pre[L::opp(in0) * ncell + cw] + (2.0f / 3.0f) * un;
    const float v1 = f_pre[L::opp(in1) * ncell + cw] + (1.0f / 6.0f) * (un + ut);
    const float v2 = f_pre[L::opp(in2) * ncell + cw] + (1.0f / 6.0f) * (un - ut);
    f[in0 * ncell + cw] = v0;
    f[(specular ? in2 : in1) * ncell + cw] = v1;
    f[(specular ? in1 : in2) * ncell + cw] = v2;
  } else {
    // in <- pre[mirror(in)] + 2 w rho_w (c_in.u_w)/cs^2, rho_w = sum_q pre; specular == bounce-back
    // in the reference                                                        (lbm3d/boundary/bb.py:9-53)
    float pre[Q], rho = 0.f;
#pragma unroll
    for (int q = 0; q < Q; ++q) { pre[q] = f_pre[q * ncell + cw]; rho += pre[q]; }
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (G::cn(q) > 0)
        f[q * ncell + cw] = pre[mirror_dir<3>(q, G::AX)] + 2.0f * L::w(q) * rho * dot_c<DIM>(q, uw) * 3.0f;
  }
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
