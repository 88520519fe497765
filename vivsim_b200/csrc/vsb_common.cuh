// vivsim_b200 -- shared device code for the IB-LBM hot path (sm_100a).
//
// Layout: SoA fp32, C order.  D2Q9 f[q][x][y], D3Q19 f[q][x][y][z]; the
// last spatial axis is contiguous.  Internally every field is addressed with three
// array axes (n0, n1, n2), n2 contiguous; 2-D fields use n0 = 1 so that x -> axis 1 and
// y -> axis 2.  Velocity component d (0..D-1) lives on array axis d + (3 - D).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vivsim_b200.h"

namespace vsb {

// ----------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define VSB_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) {                                \
      vsb::set_error(__VA_ARGS__);                \
      return VSB_ERR_INVALID;                     \
    }                                             \
  } while (0)

#define VSB_LAUNCH_CHECK(what)                                         \
  do {                                                                 \
    cudaError_t e__ = cudaGetLastError();                              \
    if (e__ != cudaSuccess) return vsb::cuda_fail(e__, what);          \
  } while (0)

// ----------------------------------------------------------------------------- lattices
// Direction numbering and weights: reference vivsim/lbm/lattice.py:43-63 (D2Q9) and
// vivsim/lbm3d/lattice.py:43-63 (D3Q19).  Velocities are stored on the three array axes.
template <int DIM> struct Lat;

template <> struct Lat<2> {
  static constexpr int D = 2, Q = 9, A0 = 1;  // A0: first array axis that carries a velocity component
  __host__ __device__ static constexpr int c(int q, int a) {
    constexpr int t[9][3] = {{0, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, -1, 0}, {0, 0, -1},
                             {0, 1, 1}, {0, -1, 1}, {0, -1, -1}, {0, 1, -1}};
    return t[q][a];
  }
  __host__ __device__ static constexpr float w(int q) { return q == 0 ? 4.0f / 9.0f : (q < 5 ? 1.0f / 9.0f : 1.0f / 36.0f); }
  __host__ __device__ static constexpr int opp(int q) {
    constexpr int t[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    return t[q];
  }
};

template <> struct Lat<3> {
  static constexpr int D = 3, Q = 19, A0 = 0;
  __host__ __device__ static constexpr int c(int q, int a) {
    constexpr int t[19][3] = {{0, 0, 0},  {1, 0, 0},  {-1, 0, 0},  {0, 1, 0},  {0, -1, 0}, {0, 0, 1},  {0, 0, -1},
                              {1, 1, 0},  {-1, 1, 0}, {1, -1, 0},  {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1},
                              {-1, 0, -1}, {0, 1, 1}, {0, -1, 1},  {0, 1, -1},  {0, -1, -1}};
    return t[q][a];
  }
  __host__ __device__ static constexpr float w(int q) { return q == 0 ? 1.0f / 3.0f : (q < 7 ? 1.0f / 18.0f : 1.0f / 36.0f); }
  __host__ __device__ static constexpr int opp(int q) {
    constexpr int t[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
    return t[q];
  }
};

// Direction whose velocity is c(q) mirrored across the plane normal to `axis`.
template <int DIM>
__host__ __device__ constexpr int mirror_dir(int q, int axis) {
  using L = Lat<DIM>;
  for (int r = 0; r < L::Q; ++r) {
    bool same = true;
    for (int a = 0; a < 3; ++a) same = same && (L::c(r, a) == (a == axis ? -L::c(q, a) : L::c(q, a)));
    if (same) return r;
  }
  return -1;
}

// Opposite-direction pairs (q, opp(q)) with q the smaller index: the rest population 0 plus NP pairs cover the
// lattice.  Used to share work between a direction and its opposite: c_opp = -c, so (c.u)^2, c_a c_b and the
// projected non-equilibrium part are equal for the two, and feq_q = A + B, feq_opp = A - B.
template <int DIM> struct Pairs;
template <> struct Pairs<2> {
  static constexpr int NP = 4;
  __host__ __device__ static constexpr int q(int k) { constexpr int t[4] = {1, 2, 5, 6}; return t[k]; }
};
template <> struct Pairs<3> {
  static constexpr int NP = 9;
  __host__ __device__ static constexpr int q(int k) { constexpr int t[9] = {1, 3, 5, 7, 8, 11, 12, 15, 16}; return t[k]; }
};

// ----------------------------------------------------------------------------- per-cell math
// rho = sum f, u = sum c f / rho            (reference lbm/basic.py:107-110, lbm3d/basic.py:102-105)
template <int DIM>
__device__ __forceinline__ void moments(const float (&f)[Lat<DIM>::Q], float& rho, float (&u)[Lat<DIM>::D]) {
  using L = Lat<DIM>;
  float r = 0.f;
#pragma unroll
  for (int q = 0; q < L::Q; ++q) r += f[q];
  rho = r;
  const float inv = 1.0f / r;   // one IEEE division; u = m * (1/rho) differs from m / rho by at most 1 ulp
#pragma unroll
  for (int d = 0; d < L::D; ++d) {
    float pos = 0.f, neg = 0.f;
#pragma unroll
    for (int q = 0; q < L::Q; ++q) {
      if (L::c(q, d + L::A0) > 0) pos += f[q];
      if (L::c(q, d + L::A0) < 0) neg += f[q];
    }
    u[d] = (pos - neg) * inv;
  }
}

template <int DIM>
__device__ __forceinline__ float dot_c(int q, const float (&v)[Lat<DIM>::D]) {
  using L = Lat<DIM>;
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < L::D; ++d) {
    const int c = L::c(q, d + L::A0);
    if (c > 0) s += v[d];
    if (c < 0) s -= v[d];
  }
  return s;
}

// feq_q = rho w_q (1 + 3 c.u + 4.5 (c.u)^2 - 1.5 u.u)   (lbm/basic.py:132-135, lbm3d/basic.py:121-130)
// evaluated per opposite pair: A = rho w (1 - 1.5 u.u + 4.5 (c.u)^2), B = 3 rho w (c.u); feq_q = A + B, feq_opp = A - B.
template <int DIM>
__device__ __forceinline__ void equilibrium(float rho, const float (&u)[Lat<DIM>::D], float (&feq)[Lat<DIM>::Q]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float usq = 0.f;
#pragma unroll
  for (int d = 0; d < L::D; ++d) usq += u[d] * u[d];
  const float base = 1.0f - 1.5f * usq;
  feq[0] = rho * L::w(0) * base;
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k);
    const float cu = dot_c<DIM>(q, u);
    const float rw = rho * L::w(q);
    const float A = rw * (base + 4.5f * cu * cu);
    const float B = 3.0f * rw * cu;
    feq[q] = A + B;
    feq[L::opp(q)] = A - B;
  }
}

// G_q = w_q [3 (c_q - u).g + 9 (c_q.u)(c_q.g)]           (lbm/forcing/guo.py:21-33, lbm3d/forcing/guo.py:13-38)
template <int DIM>
__device__ __forceinline__ void guo_term(const float (&g)[Lat<DIM>::D], const float (&u)[Lat<DIM>::D],
                                         float (&G)[Lat<DIM>::Q]) {
  using L = Lat<DIM>;
  float ug = 0.f;
#pragma unroll
  for (int d = 0; d < L::D; ++d) ug += u[d] * g[d];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) {
    const float cu = dot_c<DIM>(q, u), cg = dot_c<DIM>(q, g);
    G[q] = L::w(q) * (3.0f * (cg - ug) + 9.0f * cu * cg);
  }
}

// P fneq, P_qr = w_q/(2 cs^4) (c_q c_q - cs^2 I):(c_r c_r)   (lbm/collision/reg.py:23-47, lbm3d/collision/reg.py:10-41)
// The result is the same for a direction and its opposite, so only pair[k] = (P fneq)_{q(k)} and rest = (P fneq)_0
// are produced.
template <int DIM>
__device__ __forceinline__ void projection_from_sums(float (&e)[Pairs<DIM>::NP], float& rest, float (&pair)[Pairs<DIM>::NP]);

template <int DIM>
__device__ __forceinline__ void projection_pairs(const float (&fneq)[Lat<DIM>::Q], float& rest, float (&pair)[Pairs<DIM>::NP]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float e[P::NP];   // fneq_q + fneq_opp
#pragma unroll
  for (int k = 0; k < P::NP; ++k) e[k] = fneq[P::q(k)] + fneq[L::opp(P::q(k))];
  projection_from_sums<DIM>(e, rest, pair);
}

// e[k] = fneq_q + fneq_opp for pair k  ->  rest = (P fneq)_0, pair[k] = (P fneq)_{q(k)}
template <int DIM>
__device__ __forceinline__ void projection_from_sums(float (&e)[Pairs<DIM>::NP], float& rest, float (&pair)[Pairs<DIM>::NP]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float pi[L::D][L::D];
#pragma unroll
  for (int a = 0; a < L::D; ++a)
#pragma unroll
    for (int b = a; b < L::D; ++b) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < P::NP; ++k) {
        const int cc = L::c(P::q(k), a + L::A0) * L::c(P::q(k), b + L::A0);
        if (cc > 0) s += e[k];
        if (cc < 0) s -= e[k];
      }
      pi[a][b] = s;
      pi[b][a] = s;
    }
  float tr = 0.f;
#pragma unroll
  for (int a = 0; a < L::D; ++a) tr += pi[a][a];
  const float tr3 = tr * (1.0f / 3.0f);
  rest = L::w(0) * 4.5f * (-tr3);
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    float s = -tr3;
#pragma unroll
    for (int a = 0; a < L::D; ++a) {
      if (L::c(P::q(k), a + L::A0) != 0) s += pi[a][a];
#pragma unroll
      for (int b = a + 1; b < L::D; ++b) {
        const int cc = L::c(P::q(k), a + L::A0) * L::c(P::q(k), b + L::A0);
        if (cc > 0) s += 2.0f * pi[a][b];
        if (cc < 0) s -= 2.0f * pi[a][b];
      }
    }
    pair[k] = L::w(P::q(k)) * 4.5f * s;
  }
}

template <int DIM>
__device__ __forceinline__ void second_order_projection(const float (&fneq)[Lat<DIM>::Q], float (&out)[Lat<DIM>::Q]) {
  using P = Pairs<DIM>;
  float rest, pair[P::NP];
  projection_pairs<DIM>(fneq, rest, pair);
  out[0] = rest;
#pragma unroll
  for (int k = 0; k < P::NP; ++k) { out[P::q(k)] = pair[k]; out[Lat<DIM>::opp(P::q(k))] = pair[k]; }
}

// Relaxation constants prepared on the host exactly as Python evaluates them
// (double arithmetic on the scalar, one rounding to fp32).
struct Relax {
  float omega, one_minus_omega, inv_omega, one_minus_inv_omega, guo_scale;
};

inline Relax make_relax(double omega) {
  Relax r;
  r.omega = (float)omega;
  r.one_minus_omega = (float)(1.0 - omega);
  r.inv_omega = (float)(1.0 / omega);
  r.one_minus_inv_omega = (float)(1.0 - 1.0 / omega);
  r.guo_scale = (float)(1.0 - 0.5 * omega);
  return r;
}

template <int Q> struct Matrix { float a[Q * Q]; };

// (1 - omega) f + omega feq                               (lbm/basic.py:156)
template <int DIM>
__device__ __forceinline__ void collide_bgk(float (&f)[Lat<DIM>::Q], const float (&feq)[Lat<DIM>::Q], const Relax& r) {
#pragma unroll
  for (int q = 0; q < Lat<DIM>::Q; ++q) f[q] = r.one_minus_omega * f[q] + r.omega * feq[q];
}

// feq + (1 - omega) P (f - feq)                           (lbm/collision/reg.py:49, lbm3d/collision/reg.py:60-62)
template <int DIM>
__device__ __forceinline__ void collide_reg(float (&f)[Lat<DIM>::Q], const float (&feq)[Lat<DIM>::Q], const Relax& r) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  constexpr int Q = L::Q;
  float rest, pair[P::NP];
#pragma unroll
  for (int q = 0; q < Q; ++q) f[q] -= feq[q];   // fneq in place
  projection_pairs<DIM>(f, rest, pair);
  f[0] = feq[0] + r.one_minus_omega * rest;
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k), o = L::opp(q);
    const float v = r.one_minus_omega * pair[k];
    f[q] = feq[q] + v;
    f[o] = feq[o] + v;
  }
}

// Entropic KBC.  2-D: shear part from N = Pxx - Pyy and Pxy only (lbm/collision/kbc.py:37-44);
// 3-D: shear part = full second-order projection (lbm3d/collision/kbc.py:29-31).  Mixing: kbc.py:47-59 / :32-42.
// In both lattices the shear part is equal for a direction and its opposite, so it is held per pair.
template <int DIM>
__device__ __forceinline__ void collide_kbc(float (&f)[Lat<DIM>::Q], const float (&feq)[Lat<DIM>::Q], const Relax& r) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  // fneq = f - feq is recomputed where it is used instead of being held in Q more registers: with 4 cells per thread
  // the D3Q19 kernel sits at the 128-register budget
  float sh0, sh[P::NP];
  if constexpr (DIM == 2) {
    const float n4 = ((f[1] - feq[1]) - (f[2] - feq[2]) + (f[3] - feq[3]) - (f[4] - feq[4])) * 0.25f;
    const float p4 = ((f[5] - feq[5]) - (f[6] - feq[6]) + (f[7] - feq[7]) - (f[8] - feq[8])) * 0.25f;
    sh0 = 0.f;
    sh[0] = n4; sh[1] = -n4; sh[2] = p4; sh[3] = -p4;   // pairs (1,3) (2,4) (5,7) (6,8)
  } else {
    float e[P::NP];
#pragma unroll
    for (int k = 0; k < P::NP; ++k) {
      const int q = P::q(k), o = L::opp(q);
      e[k] = (f[q] - feq[q]) + (f[o] - feq[o]);
    }
    projection_from_sums<DIM>(e, sh0, sh);
  }
  float s_sh, s_hh;
  {
    const float hi = (f[0] - feq[0]) - sh0;
    const float inv = __fdividef(1.0f, feq[0] + 1e-20f);   // MUFU.RCP, <= 2 ulp
    s_sh = hi * sh0 * inv;
    s_hh = hi * hi * inv;
  }
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k), o = L::opp(q);
    const float hq = (f[q] - feq[q]) - sh[k], ho = (f[o] - feq[o]) - sh[k];
    const float tq = hq * __fdividef(1.0f, feq[q] + 1e-20f), to = ho * __fdividef(1.0f, feq[o] + 1e-20f);
    s_sh += sh[k] * (tq + to);
    s_hh += hq * tq + ho * to;
  }
  const float half_gamma = r.inv_omega - r.one_minus_inv_omega * s_sh / (s_hh + 1e-20f);
  // f -= omega (sh + hg (fneq - sh)); keep the difference (fneq - sh) explicit: hg can be large where it is tiny
  f[0] -= r.omega * (sh0 + half_gamma * ((f[0] - feq[0]) - sh0));
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k), o = L::opp(q);
    f[q] -= r.omega * (sh[k] + half_gamma * ((f[q] - feq[q]) - sh[k]));
    f[o] -= r.omega * (sh[k] + half_gamma * ((f[o] - feq[o]) - sh[k]));
  }
}

// f + A (feq - f), A given                               (lbm/collision/mrt.py:88, lbm3d/collision/mrt.py:96-98)
template <int DIM>
__device__ __forceinline__ void collide_mrt(float (&f)[Lat<DIM>::Q], const float (&feq)[Lat<DIM>::Q],
                                            const Matrix<Lat<DIM>::Q>& A) {
  constexpr int Q = Lat<DIM>::Q;
  float d[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) d[q] = feq[q] - f[q];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < Q; ++j) s = fmaf(A.a[i * Q + j], d[j], s);
    f[i] += s;
  }
}

template <int DIM>
__device__ __forceinline__ void matvec_add(float (&f)[Lat<DIM>::Q], const Matrix<Lat<DIM>::Q>& B,
                                           const float (&G)[Lat<DIM>::Q]) {
  constexpr int Q = Lat<DIM>::Q;
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < Q; ++j) s = fmaf(B.a[i * Q + j], G[j], s);
    f[i] += s;
  }
}

// ----------------------------------------------------------------------------- IB delta kernels
// ib/kernels.py:4-61 (+ the 2-point hat, which the reference names in its README but does not define)
__device__ __forceinline__ float delta(int kind, float r) {
  const float a = fabsf(r);
  switch (kind) {
    case VSB_DELTA_PESKIN3:
      if (a > 1.5f) return 0.f;
      if (a < 0.5f) return (1.0f + sqrtf(1.0f - 3.0f * a * a)) / 3.0f;
      return (5.0f - 3.0f * a - sqrtf(-2.0f + 6.0f * a - 3.0f * a * a)) / 6.0f;
    case VSB_DELTA_PESKIN4:
      if (a > 2.0f) return 0.f;
      if (a < 1.0f) return (3.0f - 2.0f * a + sqrtf(1.0f + 4.0f * a - 4.0f * a * a)) * 0.125f;
      return (5.0f - 2.0f * a - sqrtf(-7.0f + 12.0f * a - 4.0f * a * a)) * 0.125f;
    case VSB_DELTA_COSINE4:
      if (a > 2.0f) return 0.f;
      return (1.0f + cosf(3.14159265358979323846f * a * 0.5f)) * 0.25f;
    default:  // VSB_DELTA_HAT2
      return fmaxf(0.f, 1.0f - a);
  }
}

// ----------------------------------------------------------------------------- launch helpers
inline unsigned blocks_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace vsb
