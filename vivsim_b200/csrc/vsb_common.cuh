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
// 1 / x to 1 ulp (MUFU.RCP); x must be a normal number
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sum over the pairs of sign(c_q(k),axis) * v[k], without a leading "0 +"
template <int DIM>
__device__ __forceinline__ float signed_pair_sum(const float (&v)[Pairs<DIM>::NP], int axis) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float s = 0.f;
  bool have = false;
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int c = L::c(P::q(k), axis);
    if (c != 0) {
      const float t = (c > 0) ? v[k] : -v[k];
      s = have ? s + t : t;
      have = true;
    }
  }
  return s;
}

// rho = sum f, u = sum c f / rho            (reference lbm/basic.py:107-110, lbm3d/basic.py:102-105)
// summed over opposite pairs: rho = f_0 + sum_k (f_q + f_opp), momentum = sum_k c_q (f_q - f_opp)
template <int DIM>
__device__ __forceinline__ void moments(const float (&f)[Lat<DIM>::Q], float& rho, float (&u)[Lat<DIM>::D]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float d[P::NP];
  float r = f[0];
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k), o = L::opp(q);
    r += f[q] + f[o];
    d[k] = f[q] - f[o];
  }
  rho = r;
  const float inv = 1.0f / r;   // one IEEE division; u = m * (1/rho) differs from m / rho by at most 1 ulp
#pragma unroll
  for (int a = 0; a < L::D; ++a) u[a] = signed_pair_sum<DIM>(d, a + L::A0) * inv;
}

template <int DIM>
__device__ __forceinline__ float dot_c(int q, const float (&v)[Lat<DIM>::D]) {
  using L = Lat<DIM>;
  float s = 0.f;
  bool have = false;
#pragma unroll
  for (int d = 0; d < L::D; ++d) {
    const int c = L::c(q, d + L::A0);
    if (c != 0) {
      const float t = (c > 0) ? v[d] : -v[d];
      s = have ? s + t : t;
      have = true;
    }
  }
  return s;
}

// feq_q = rho w_q (1 + 3 c.u + 4.5 (c.u)^2 - 1.5 u.u)   (lbm/basic.py:132-135, lbm3d/basic.py:121-130)
// evaluated per opposite pair: A = rho w (1 - 1.5 u.u + 4.5 (c.u)^2), B = 3 rho w (c.u); feq_q = A + B, feq_opp = A - B.
template <int DIM>
__device__ __forceinline__ void equilibrium_pairs(float rho, const float (&u)[Lat<DIM>::D], float& feq0,
                                                  float (&A)[Pairs<DIM>::NP], float (&B)[Pairs<DIM>::NP]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float usq = u[0] * u[0];
#pragma unroll
  for (int d = 1; d < L::D; ++d) usq += u[d] * u[d];
  const float base = 1.0f - 1.5f * usq;
  feq0 = rho * L::w(0) * base;
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k);
    const float cu = dot_c<DIM>(q, u);
    const float rw = rho * L::w(q);
    A[k] = rw * (base + 4.5f * cu * cu);
    B[k] = 3.0f * rw * cu;
  }
}

template <int DIM>
__device__ __forceinline__ void equilibrium(float rho, const float (&u)[Lat<DIM>::D], float (&feq)[Lat<DIM>::Q]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float A[P::NP], B[P::NP];
  equilibrium_pairs<DIM>(rho, u, feq[0], A, B);
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k);
    feq[q] = A[k] + B[k];
    feq[L::opp(q)] = A[k] - B[k];
  }
}

// G_q = w_q [3 (c_q - u).g + 9 (c_q.u)(c_q.g)]           (lbm/forcing/guo.py:21-33, lbm3d/forcing/guo.py:13-38)
// per opposite pair: G_q = Hs + Ha, G_opp = Hs - Ha with Hs = w (9 (c.u)(c.g) - 3 u.g), Ha = 3 w (c.g)
template <int DIM>
__device__ __forceinline__ void guo_term_pairs(const float (&g)[Lat<DIM>::D], const float (&u)[Lat<DIM>::D], float& G0,
                                               float (&Hs)[Pairs<DIM>::NP], float (&Ha)[Pairs<DIM>::NP]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float ug = u[0] * g[0];
#pragma unroll
  for (int d = 1; d < L::D; ++d) ug += u[d] * g[d];
  G0 = L::w(0) * (-3.0f * ug);
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k);
    const float cu = dot_c<DIM>(q, u), cg = dot_c<DIM>(q, g);
    Hs[k] = L::w(q) * (9.0f * cu * cg - 3.0f * ug);
    Ha[k] = (3.0f * L::w(q)) * cg;
  }
}

template <int DIM>
__device__ __forceinline__ void guo_term(const float (&g)[Lat<DIM>::D], const float (&u)[Lat<DIM>::D],
                                         float (&G)[Lat<DIM>::Q]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float Hs[P::NP], Ha[P::NP];
  guo_term_pairs<DIM>(g, u, G[0], Hs, Ha);
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k);
    G[q] = Hs[k] + Ha[k];
    G[L::opp(q)] = Hs[k] - Ha[k];
  }
}

// P fneq, P_qr = w_q/(2 cs^4) (c_q c_q - cs^2 I):(c_r c_r)   (lbm/collision/reg.py:23-47, lbm3d/collision/reg.py:10-41)
// The result is the same for a direction and its opposite, so only pair[k] = (P fneq)_{q(k)} and rest = (P fneq)_0
// are produced.
template <int DIM, int SCALE>
__device__ __forceinline__ void projection_from_sums(const float (&e)[Pairs<DIM>::NP], float& rest, float (&pair)[Pairs<DIM>::NP]);

template <int DIM>
__device__ __forceinline__ void projection_pairs(const float (&fneq)[Lat<DIM>::Q], float& rest, float (&pair)[Pairs<DIM>::NP]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float e[P::NP];   // fneq_q + fneq_opp
#pragma unroll
  for (int k = 0; k < P::NP; ++k) e[k] = fneq[P::q(k)] + fneq[L::opp(P::q(k))];
  projection_from_sums<DIM, 1>(e, rest, pair);
}

// e[k] = (fneq_q + fneq_opp) / SCALE for pair k  ->  rest = (P fneq)_0, pair[k] = (P fneq)_{q(k)}
template <int DIM, int SCALE>
__device__ __forceinline__ void projection_from_sums(const float (&e)[Pairs<DIM>::NP], float& rest, float (&pair)[Pairs<DIM>::NP]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float pi[L::D][L::D];
#pragma unroll
  for (int a = 0; a < L::D; ++a)
#pragma unroll
    for (int b = a; b < L::D; ++b) {
      float s = 0.f;
      bool have = false;
#pragma unroll
      for (int k = 0; k < P::NP; ++k) {
        const int cc = L::c(P::q(k), a + L::A0) * L::c(P::q(k), b + L::A0);
        if (cc != 0) {
          const float t = (cc > 0) ? e[k] : -e[k];
          s = have ? s + t : t;
          have = true;
        }
      }
      pi[a][b] = s;
      pi[b][a] = s;
    }
  float tr = pi[0][0];
#pragma unroll
  for (int a = 1; a < L::D; ++a) tr += pi[a][a];
  const float tr3 = tr * (1.0f / 3.0f);
  rest = L::w(0) * (4.5f * SCALE) * (-tr3);
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
    pair[k] = L::w(P::q(k)) * (4.5f * SCALE) * s;
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
//
// Everything is held per opposite pair (q, o): with feq_q = A + B, feq_o = A - B the non-equilibrium part splits into
// the half sum es = (fneq_q + fneq_o)/2 and the half difference ea = (fneq_q - fneq_o)/2; the shear part sh is even
// (equal for q and o), so the higher-order part is h_q = hs + ea, h_o = hs - ea with hs = es - sh.  The result
// f - omega (sh + gamma/2 h) is written as feq + (1 - omega) sh + (1 - omega gamma/2) h, which needs only pair
// quantities: f_q = sym + asym, f_o = sym - asym.
template <int DIM>
__device__ __forceinline__ void collide_kbc_pairs(float (&f)[Lat<DIM>::Q], float feq0, const float (&A)[Pairs<DIM>::NP],
                                                  const float (&B)[Pairs<DIM>::NP], const Relax& r) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float es[P::NP], ea[P::NP];
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k), o = L::opp(q);
    es[k] = 0.5f * (f[q] + f[o]) - A[k];
    ea[k] = 0.5f * (f[q] - f[o]) - B[k];
  }
  const float e0 = f[0] - feq0;
  float sh0, sh[P::NP];
  if constexpr (DIM == 2) {
    // N/4 with N = fneq_1 - fneq_2 + fneq_3 - fneq_4, Pxy/4 with Pxy = fneq_5 - fneq_6 + fneq_7 - fneq_8
    const float n4 = 0.5f * (es[0] - es[1]);
    const float p4 = 0.5f * (es[2] - es[3]);
    sh0 = 0.f;
    sh[0] = n4; sh[1] = -n4; sh[2] = p4; sh[3] = -p4;   // pairs (1,3) (2,4) (5,7) (6,8)
  } else {
    projection_from_sums<DIM, 2>(es, sh0, sh);
  }
  // entropic stabiliser: gamma/2 = 1/omega - (1 - 1/omega) <sh|h> / <h|h>,  <a|b> = sum a b / feq
  const float h0 = e0 - sh0;
  float s_sh, s_hh;
  {
    const float t0 = h0 * rcp_approx(feq0 + 1e-20f);
    s_sh = sh0 * t0;
    s_hh = h0 * t0;
  }
  float hs[P::NP], sym[P::NP];
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    hs[k] = es[k] - sh[k];
    const float hq = hs[k] + ea[k], ho = hs[k] - ea[k];
    const float Ae = A[k] + 1e-20f;
    const float tq = hq * rcp_approx(Ae + B[k]), to = ho * rcp_approx(Ae - B[k]);
    s_sh += sh[k] * (tq + to);
    s_hh += hq * tq + ho * to;
    sym[k] = A[k] + r.one_minus_omega * sh[k];
  }
  const float half_gamma = r.inv_omega - r.one_minus_inv_omega * s_sh * rcp_approx(s_hh + 1e-20f);
  const float ch = 1.0f - r.omega * half_gamma;
  f[0] = feq0 + r.one_minus_omega * sh0 + ch * h0;
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k), o = L::opp(q);
    const float s = sym[k] + ch * hs[k], a = B[k] + ch * ea[k];
    f[q] = s + a;
    f[o] = s - a;
  }
}

// the same from an equilibrium given per direction (stand-alone operator)
template <int DIM>
__device__ __forceinline__ void collide_kbc(float (&f)[Lat<DIM>::Q], const float (&feq)[Lat<DIM>::Q], const Relax& r) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float A[P::NP], B[P::NP];
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k), o = L::opp(q);
    A[k] = 0.5f * (feq[q] + feq[o]);
    B[k] = 0.5f * (feq[q] - feq[o]);
  }
  collide_kbc_pairs<DIM>(f, feq[0], A, B, r);
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

// Parity-split form of a Q x Q operator M that commutes with the reflection c -> -c, M[q][r] == M[opp q][opp r]
// (every moment-space operator M^-1 S M is of this kind: the moments are even or odd polynomials of c).
// With xs_k = x_q + x_opp and xa_k = x_q - x_opp over the opposite pairs,
//   (M x)_0 = E[0][0] x_0 + sum_j E[0][1+j] xs_j
//   (M x)_q = sym_k + asym_k, (M x)_opp = sym_k - asym_k,
//   sym_k = E[1+k][0] x_0 + sum_j E[1+k][1+j] xs_j,  asym_k = sum_j O[k][j] xa_j
// i.e. (1+NP)^2 + NP^2 multiply-adds instead of Q^2 (181 instead of 361 for D3Q19, 41 instead of 81 for D2Q9).
template <int DIM> struct SplitOp {
  static constexpr int NP = Pairs<DIM>::NP, NE = NP + 1;
  float e[NE * NE];
  float o[NP * NP];
};

// Returns false (and leaves `s` undefined) when M is not parity-symmetric to within tol * max|M|.
template <int DIM>
inline bool make_split_op(const float* M, SplitOp<DIM>& s, double tol = 1e-6) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  constexpr int Q = L::Q, NP = P::NP, NE = NP + 1;
  double big = 0.0, asym = 0.0;
  for (int q = 0; q < Q; ++q)
    for (int r = 0; r < Q; ++r) {
      const double a = M[q * Q + r], b = M[L::opp(q) * Q + L::opp(r)];
      big = a > big ? a : (-a > big ? -a : big);
      const double d = a > b ? a - b : b - a;
      asym = d > asym ? d : asym;
    }
  if (asym > tol * big) return false;
  auto m = [&](int q, int r) { return 0.5 * ((double)M[q * Q + r] + (double)M[L::opp(q) * Q + L::opp(r)]); };
  s.e[0] = (float)m(0, 0);
  for (int j = 0; j < NP; ++j) {
    const int r = P::q(j), ro = L::opp(r);
    s.e[1 + j] = (float)(0.5 * (m(0, r) + m(0, ro)));
  }
  for (int k = 0; k < NP; ++k) {
    const int q = P::q(k);
    s.e[(1 + k) * NE] = (float)m(q, 0);
    for (int j = 0; j < NP; ++j) {
      const int r = P::q(j), ro = L::opp(r);
      s.e[(1 + k) * NE + 1 + j] = (float)(0.5 * (m(q, r) + m(q, ro)));
      s.o[k * NP + j] = (float)(0.5 * (m(q, r) - m(q, ro)));
    }
  }
  return true;
}

// f += M x for x given as (x_0, xs, xa)
template <int DIM>
__device__ __forceinline__ void split_matvec_add(float (&f)[Lat<DIM>::Q], const SplitOp<DIM>& M, float x0,
                                                 const float (&xs)[Pairs<DIM>::NP], const float (&xa)[Pairs<DIM>::NP]) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  constexpr int NP = P::NP, NE = NP + 1;
  {
    float s = M.e[0] * x0;
#pragma unroll
    for (int j = 0; j < NP; ++j) s = fmaf(M.e[1 + j], xs[j], s);
    f[0] += s;
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int q = P::q(k), o = L::opp(q);
    float s = M.e[(1 + k) * NE] * x0;
    float a = M.o[k * NP] * xa[0];
#pragma unroll
    for (int j = 0; j < NP; ++j) s = fmaf(M.e[(1 + k) * NE + 1 + j], xs[j], s);
#pragma unroll
    for (int j = 1; j < NP; ++j) a = fmaf(M.o[k * NP + j], xa[j], a);
    f[q] += s + a;
    f[o] += s - a;
  }
}

// MRT collision f + A (feq - f) with A in parity-split form and the equilibrium per pair
template <int DIM>
__device__ __forceinline__ void collide_mrt_split(float (&f)[Lat<DIM>::Q], float feq0, const float (&A)[Pairs<DIM>::NP],
                                                  const float (&B)[Pairs<DIM>::NP], const SplitOp<DIM>& M) {
  using L = Lat<DIM>;
  using P = Pairs<DIM>;
  float xs[P::NP], xa[P::NP];
#pragma unroll
  for (int k = 0; k < P::NP; ++k) {
    const int q = P::q(k), o = L::opp(q);
    xs[k] = 2.0f * A[k] - (f[q] + f[o]);
    xa[k] = 2.0f * B[k] - (f[q] - f[o]);
  }
  split_matvec_add<DIM>(f, M, feq0 - f[0], xs, xa);
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
