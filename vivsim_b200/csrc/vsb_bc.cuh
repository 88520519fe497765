// Face boundary operations on populations held in registers.
// Shared by the in-place per-function kernels (vsb_boundary.cu) and the fused wall kernel (vsb_step.cu).
#pragma once

#include "vsb_common.cuh"

namespace vsb {

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
  static constexpr int AX = LOC / 2 + L::A0;           // array axis normal to the face
  static constexpr int SIGN = (LOC % 2 == 0) ? 1 : -1;  // inward normal direction along AX
  static constexpr int ND = AX - L::A0;                // velocity component normal to the face
  static constexpr int TA = (AX == 0) ? 1 : 0;         // remaining array axes, ascending
  static constexpr int TB = (AX == 2) ? 1 : 2;
  __host__ __device__ static constexpr int cn(int q) { return L::c(q, AX) * SIGN; }  // > 0: enters the fluid
};

__host__ __device__ constexpr bool bc_needs_neighbor(int kind, int wrap) {
  return kind == VSB_BC_NEE || wrap == VSB_WRAP_PRESSURE;
}

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

// NEE / NEBB / equilibrium with the velocity / pressure / force-corrected wrappers, on the wall cell's streamed
// populations fw (updated in place); fn = streamed populations of the adjacent fluid cell (read when
// bc_needs_neighbor).
template <int DIM, int LOC>
__device__ __forceinline__ void apply_face_bc(float (&fw)[Lat<DIM>::Q], const float (&fn)[Lat<DIM>::Q], int kind, int wrap,
                                              float rho_w, float (&uw)[Lat<DIM>::D], const float (&gw)[Lat<DIM>::D]) {
  using L = Lat<DIM>;
  using G = FaceGeom<DIM, LOC>;
  constexpr int Q = L::Q, D = L::D;
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
    for (int d = 0; d < D; ++d) uw[d] -= gw[d] * 0.5f / rho_w;
  }

  if (kind == VSB_BC_EQUILIBRIUM) {            // lbm/boundary/eq.py:45-56
    equilibrium<DIM>(rho_w, uw, fw);
  } else if (kind == VSB_BC_NEE) {             // lbm/boundary/nee.py:43-60
    float fe[Q], fen[Q], rho_nb, u_nb[D];
    equilibrium<DIM>(rho_w, uw, fe);
    moments<DIM>(fn, rho_nb, u_nb);
    equilibrium<DIM>(rho_nb, u_nb, fen);
#pragma unroll
    for (int q = 0; q < Q; ++q) fw[q] = fe[q] + (fn[q] - fen[q]);
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
      const float v0 = fw[L::opp(in0)] + (2.0f / 3.0f) * un * rho_w;
      const float v1 = fw[L::opp(in1)] - shear + normal + tang;
      const float v2 = fw[L::opp(in2)] + shear + normal - tang;
      fw[in0] = v0; fw[in1] = v1; fw[in2] = v2;
    } else {
      // f_in = f_opp(in) + feq_in - feq_opp(in)    (lbm3d/boundary/nebb.py:16-32)
      float fe[Q];
      equilibrium<DIM>(rho_w, uw, fe);
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (G::cn(q) > 0) fw[q] = fw[L::opp(q)] + fe[q] - fe[L::opp(q)];
    }
  }
}

// bounce-back / specular reflection: pre = PRE-streaming populations of the wall cell.
template <int DIM, int LOC>
__device__ __forceinline__ void apply_face_reflect(float (&fw)[Lat<DIM>::Q], const float (&pre)[Lat<DIM>::Q], int specular,
                                                   const float (&uw)[Lat<DIM>::D]) {
  using L = Lat<DIM>;
  using G = FaceGeom<DIM, LOC>;
  constexpr int Q = L::Q;
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
    const float v0 = pre[L::opp(in0)] + (2.0f / 3.0f) * un;
    const float v1 = pre[L::opp(in1)] + (1.0f / 6.0f) * (un + ut);
    const float v2 = pre[L::opp(in2)] + (1.0f / 6.0f) * (un - ut);
    fw[in0] = v0;
    if (specular) { fw[in2] = v1; fw[in1] = v2; }   // diagonal targets swapped (bb.py:93)
    else { fw[in1] = v1; fw[in2] = v2; }
  } else {
    // in <- pre[mirror(in)] + 2 w rho_w (c_in.u_w)/cs^2, rho_w = sum_q pre; the reference's 3-D specular
    // reflection has the same body                                            (lbm3d/boundary/bb.py:9-53)
    float rho = 0.f;
#pragma unroll
    for (int q = 0; q < Q; ++q) rho += pre[q];
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (G::cn(q) > 0) fw[q] = pre[mirror_dir<3>(q, G::AX)] + 2.0f * L::w(q) * rho * dot_c<DIM>(q, uw) * 3.0f;
  }
}

}  // namespace vsb
