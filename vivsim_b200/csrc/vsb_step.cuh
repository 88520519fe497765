// Device code shared by the fused step kernels (vsb_step.cu) and the fused IB kernel (vsb_ibfused.cu).
#pragma once

#include <utility>

#include "vsb_bc.cuh"
#include "vsb_internal.h"
#include "vsb_mrt_moment.cuh"

namespace vsb {

// compile-time loop: body(std::integral_constant<int, I>) for I in [0, N)
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& body, std::integer_sequence<int, I...>) {
  (body(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& body) {
  static_for_impl(body, std::make_integer_sequence<int, N>{});
}

// Division of an index n < 2^31 by a launch constant: quotient = umulhi(n, mul) >> shr (divisor 1: mul = 0).
struct FastDiv { unsigned mul, shr; };

inline FastDiv make_fast_div(unsigned d) {
  FastDiv f{0u, 0u};
  if (d <= 1) return f;
  unsigned lg = 0;
  while ((1ull << lg) < d) ++lg;                       // ceil(log2 d)
  const unsigned p = 31 + lg;
  f.mul = (unsigned)(((1ull << p) + d - 1) / d);
  f.shr = p - 32;
  return f;
}

__device__ __forceinline__ unsigned fast_div(unsigned n, const FastDiv& f) {
  return f.mul ? (__umulhi(n, f.mul) >> f.shr) : n;
}

// One face operation handled inside the fused kernel's launch (extra blocks after the bulk blocks).
struct WallOpDev {
  int kind, wrap, loc, layer, mask_before;
  WallVals w;
};

// Halo hand-shake done by the edge-row launch itself (see VsbStepArgs.halo; the flag protocol is vsb_halo.cu's).
struct HaloDev {
  int mode;                        // bit 0: wait for the neighbours first; bit 1: send the crossing populations + publish
  volatile unsigned* my_flags;     // [0] written by the left neighbour, [1] by the right neighbour
  volatile unsigned* left_flags;   // left neighbour's flag words: this rank writes [1]
  volatile unsigned* right_flags;  // right neighbour's flag words: this rank writes [0]
  unsigned* counter;               // [0] step number, [1] CTA ticket, [2] set to 1 if a wait timed out
  float* left_state;               // the neighbours' copies of the buffer this launch writes (f_out)
  float* right_state;
};

template <int DIM> struct StepParams {
  int n0, n1, n2;
  int r_begin, r_end;   // physical rows of array axis A0 (the slowest real axis); ghost layers lie outside
  int s_begin, s_end;   // rows this launch updates, within [r_begin, r_end)
  int edge_rows;        // 1: only rows r_begin and r_end - 1
  const float* fin;
  float* fout;
  int do_stream, do_collide, forcing;
  Relax rx;
  float g0[3];
  const float* gwin;
  int worg[3], wsz[3];
  int wshift[3];        // added to body->origin2 (global coordinates of a slab-decomposed run -> local)
  const VsbBodyState* body;
  int parity;
  const uint8_t* mask;
  int band;             // 0 all rows, 1 skip the window's band, 2 only the band.  Band = the window's box, rounded out
                        // to whole vector groups along y (2-D), or its (x, y) footprint over all z (3-D)
  int n_skip;           // wall layers (normal to a non-contiguous axis) left to the fused wall kernel
  int skip_axis[2], skip_layer[2];
  int n_wall;           // face operations executed by the blocks appended to this launch (edges = 2)
  unsigned nb_bulk, wall_blocks0;
  WallOpDev wall[2];
  FastDiv div_nv, div_n1;   // thread index -> (row, vector column), row -> (i0, i1)
  FastDiv div_w1;           // 3-D band 2: row -> (x plane of the window, y line of the window)
  FastDiv div_nvb;          // 2-D band 2: thread index -> (row, vector group of the window's y-range)
  int nvb;                  // 2-D band 2: vector groups enumerated per row
  int prefetch_blocks;      // > 0: pull the lines of the block that many blocks ahead into L2
  HaloDev halo;
};

// MRT operators: collision A and Guo source B (lbm/collision/mrt.py:88, lbm/forcing/guo.py:60-75) as dense matrices
// and in parity-split form.  Operators that commute with the reflection c -> -c (every M^-1 S M does) run the
// kernels instantiated for VSB_COLL_MRT_SPLIT, which use only As / Bs; any other matrix runs the dense kernels.
// D3Q19 operators that are diagonal in the reference's moment basis with zero rates for the conserved moments (what
// get_mrt_collision_operator builds) run the kernels instantiated for VSB_COLL_MRT_MOMENT: no matrix at all
// (vsb_mrt_moment.cuh), and the Guo source operator I - A/2 folded into the same application.
constexpr int VSB_COLL_MRT_SPLIT = 100;   // internal collision ids, never cross the ABI
constexpr int VSB_COLL_MRT_MOMENT = 101;
__host__ __device__ constexpr bool is_mrt(int coll) {
  return coll == VSB_COLL_MRT || coll == VSB_COLL_MRT_SPLIT || coll == VSB_COLL_MRT_MOMENT;
}
// what a kernel instantiated for `coll` receives next to StepParams: 0 nothing, 1 the matrices, 2 the moment rates
__host__ __device__ constexpr int mats_kind(int coll) {
  return coll == VSB_COLL_MRT_MOMENT ? 2 : ((coll == VSB_COLL_MRT || coll == VSB_COLL_MRT_SPLIT) ? 1 : 0);
}

template <int DIM, int KIND> struct MrtMats {};
template <int DIM> struct MrtMats<DIM, 1> {
  Matrix<Lat<DIM>::Q> A, B;
  SplitOp<DIM> As, Bs;
};
template <int DIM> struct MrtMats<DIM, 2> {
  MomentOp3 mo;
};

__device__ __forceinline__ int wrap(int i, int n) {
  i += (i < 0) ? n : 0;
  i -= (i >= n) ? n : 0;
  return i;
}

// Integer origin of the force / IB window for this step.
template <int DIM>
__device__ __forceinline__ void window_origin(const StepParams<DIM>& p, int (&worg)[3]) {
  if (p.body) {
    worg[0] = p.body->origin2[p.parity][0] + p.wshift[0];
    worg[1] = p.body->origin2[p.parity][1] + p.wshift[1];
    worg[2] = p.body->origin2[p.parity][2] + p.wshift[2];
  } else {
    worg[0] = p.worg[0]; worg[1] = p.worg[1]; worg[2] = p.worg[2];
  }
}

// Window origin rule for a displaced body (host and device):
//   follow 1: astype(int32) truncation      (examples/2d/vortex_induced_vibration.py:104-105)
//   follow 2: clip(floor(.), 0, N - size)   (examples/3d/oscillating_cylinder.py:241-243)
// Both are kept inside [0, N - size]: lax.dynamic_slice clamps the start of the slice in the same way, so a body that
// drifts to the edge of the grid never makes the kernels address cells outside it.
__host__ __device__ inline int origin_rule(int follow, float origin0, float d, int grid_n, int win_n) {
  const float shifted = origin0 + (follow ? d : 0.f);
  int o = (follow == 2) ? (int)floorf(shifted) : (int)shifted;
  if (follow) {
    o = o > grid_n - win_n ? grid_n - win_n : o;
    o = o < 0 ? 0 : o;
  }
  return o;
}

// Body update: h = -force_sum + a*added_mass; Newmark-beta (gamma 1/2, beta 1/4, dt 1); next window origin.
// (dyn.py:27-51,136; examples/2d/vortex_induced_vibration.py:135-137)      One thread, host or device.
// Scalar form (dyn.py:44-46): every degree of freedom uses the same m, k, c.  Matrix form (dyn.py:36-42):
// a_next = (M + C/2 + K/4)^-1 (h - C v1 - K d1); the inverse is formed once on the host in double precision.
struct BodyUpdate {
  int n_dof, follow, dim, matrix;
  float origin0[3];
  int grid_size[3], win_size[3];
  float denom, k, c, added_mass;   // scalar form; denom = m + c/2 + k/4 evaluated in double on the host
  float minv[9], kmat[9], cmat[9], added_v[3];   // matrix form (row-major 3 x 3, leading n_dof x n_dof block)
  float* history;                  // optional (capacity, 6) ring of (d, h) per step
  int history_capacity;
};

inline BodyUpdate make_body_update(const VsbBodyParams& bp, int dim) {
  BodyUpdate u{};
  u.n_dof = bp.n_dof; u.follow = bp.follow; u.dim = dim; u.matrix = bp.matrix_form ? 1 : 0;
  for (int d = 0; d < 3; ++d) { u.origin0[d] = bp.origin0[d]; u.grid_size[d] = bp.grid_size[d]; u.win_size[d] = bp.win_size[d]; }
  u.denom = (float)(bp.m + 0.5 * bp.c + 0.25 * bp.k);
  u.k = (float)bp.k; u.c = (float)bp.c; u.added_mass = (float)bp.added_mass;
  if (u.matrix) {
    // effective mass matrix and its inverse (Gauss-Jordan with partial pivoting, double) on the n_dof x n_dof block
    const int n = bp.n_dof < 1 ? 1 : (bp.n_dof > 3 ? 3 : bp.n_dof);
    double a[3][6] = {};
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) a[i][j] = bp.mat_m[3 * i + j] + 0.5 * bp.mat_c[3 * i + j] + 0.25 * bp.mat_k[3 * i + j];
      a[i][n + i] = 1.0;
    }
    for (int col = 0; col < n; ++col) {
      int piv = col;
      for (int r = col + 1; r < n; ++r)
        if ((a[r][col] < 0 ? -a[r][col] : a[r][col]) > (a[piv][col] < 0 ? -a[piv][col] : a[piv][col])) piv = r;
      for (int j = 0; j < 2 * n; ++j) { const double t = a[col][j]; a[col][j] = a[piv][j]; a[piv][j] = t; }
      const double d = a[col][col];
      if (d == 0.0) continue;   // singular: validated by the caller (make_body_update has no error channel)
      for (int j = 0; j < 2 * n; ++j) a[col][j] /= d;
      for (int r = 0; r < n; ++r)
        if (r != col) {
          const double fct = a[r][col];
          for (int j = 0; j < 2 * n; ++j) a[r][j] -= fct * a[col][j];
        }
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        u.minv[3 * i + j] = (float)a[i][n + j];
        u.kmat[3 * i + j] = (float)bp.mat_k[3 * i + j];
        u.cmat[3 * i + j] = (float)bp.mat_c[3 * i + j];
      }
    for (int i = 0; i < 3; ++i) u.added_v[i] = (float)bp.added_mass_v[i];
  }
  u.history = bp.history_capacity > 0 ? bp.history : nullptr;
  u.history_capacity = bp.history_capacity;
  return u;
}

__host__ __device__ inline void body_update(VsbBodyState* b, const BodyUpdate& u, int parity) {
  if (!u.matrix) {
    for (int i = 0; i < u.n_dof; ++i) {
      const float h = -b->force_sum[i] + b->a[i] * u.added_mass;
      const float v1 = b->v[i] + 0.5f * b->a[i];
      const float d1 = b->d[i] + b->v[i] + 0.25f * b->a[i];
      const float a_next = (h - u.c * v1 - u.k * d1) / u.denom;
      b->h[i] = h;
      b->a[i] = a_next;
      b->v[i] = 0.5f * a_next + v1;
      b->d[i] = 0.25f * a_next + d1;
    }
  } else {
    float h[3], v1[3], d1[3], rhs[3];
    for (int i = 0; i < u.n_dof; ++i) {
      h[i] = -b->force_sum[i] + b->a[i] * u.added_v[i];
      v1[i] = b->v[i] + 0.5f * b->a[i];
      d1[i] = b->d[i] + b->v[i] + 0.25f * b->a[i];
    }
    for (int i = 0; i < u.n_dof; ++i) {
      float cv = 0.f, kd = 0.f;
      for (int j = 0; j < u.n_dof; ++j) { cv += u.cmat[3 * i + j] * v1[j]; kd += u.kmat[3 * i + j] * d1[j]; }
      rhs[i] = h[i] - cv - kd;
    }
    for (int i = 0; i < u.n_dof; ++i) {
      float a_next = 0.f;
      for (int j = 0; j < u.n_dof; ++j) a_next += u.minv[3 * i + j] * rhs[j];
      b->h[i] = h[i];
      b->a[i] = a_next;
      b->v[i] = 0.5f * a_next + v1[i];
      b->d[i] = 0.25f * a_next + d1[i];
    }
  }
  for (int i = 0; i < 3; ++i) b->force_sum[i] = 0.f;
  for (int d = 0; d < u.dim; ++d)
    b->origin2[parity ^ 1][d] = origin_rule(u.follow, u.origin0[d], b->d[d], u.grid_size[d], u.win_size[d]);
  if (u.history) {   // per-step record of (d, h), like the scan outputs of the reference's update_chunk
    float* row = u.history + 6 * (b->step % u.history_capacity);
    for (int i = 0; i < 3; ++i) { row[i] = b->d[i]; row[3 + i] = b->h[i]; }
  }
  b->step += 1;
}

// Streamed (pulled) and masked populations of one cell; scalar loads.  With do_stream = 0 the cell itself.
template <int DIM>
__device__ __forceinline__ void pull_cell(const StepParams<DIM>& p, int c0, int c1, int c2, float (&f)[Lat<DIM>::Q],
                                          bool use_mask) {
  using L = Lat<DIM>;
  const long long ncell = (long long)p.n0 * p.n1 * p.n2;
#pragma unroll
  for (int q = 0; q < L::Q; ++q) {
    const int s0 = p.do_stream ? wrap(c0 - L::c(q, 0), p.n0) : c0;
    const int s1 = p.do_stream ? wrap(c1 - L::c(q, 1), p.n1) : c1;
    const int s2 = p.do_stream ? wrap(c2 - L::c(q, 2), p.n2) : c2;
    f[q] = __ldg(p.fin + q * ncell + ((long long)s0 * p.n1 + s1) * p.n2 + s2);
  }
  if (use_mask && p.do_stream && p.mask && p.mask[((long long)c0 * p.n1 + c1) * p.n2 + c2]) {
    float t[L::Q];
#pragma unroll
    for (int q = 0; q < L::Q; ++q) t[q] = f[q];
#pragma unroll
    for (int q = 0; q < L::Q; ++q) f[q] = t[L::opp(q)];
  }
}

// Force window lookup, split so that the part that does not depend on the contiguous coordinate is done once
// per thread: `base` is the flat window index of (c0, c1, 0) and `rows_inside` tells whether the leading
// coordinates fall inside the window.
template <int DIM>
__device__ __forceinline__ void window_rows(const StepParams<DIM>& p, const int (&worg)[3], int c0, int c1,
                                            bool& rows_inside, int& base) {
  using L = Lat<DIM>;
  rows_inside = p.gwin != nullptr;
  base = 0;
  const int coord[3] = {c0, c1, 0};
#pragma unroll
  for (int d = 0; d < L::D - 1; ++d) {
    const int rel = coord[d + L::A0] - worg[d];
    rows_inside = rows_inside && (unsigned)rel < (unsigned)p.wsz[d];
    base = (base + rel) * p.wsz[d + 1];
  }
}

template <int DIM>
__device__ __forceinline__ void cell_force(const StepParams<DIM>& p, const int (&worg)[3], bool rows_inside, int base,
                                           int c2, float (&g)[Lat<DIM>::D]) {
  using L = Lat<DIM>;
#pragma unroll
  for (int d = 0; d < L::D; ++d) g[d] = p.g0[d];
  const int rel = c2 - worg[L::D - 1];
  if (rows_inside && (unsigned)rel < (unsigned)p.wsz[L::D - 1]) {
    // window fields are stored cell-major with the components packed (float2 in 2-D, float4 in 3-D) so that the IB
    // kernels can use one vector gather / one vector atomic per stencil point
    if constexpr (L::D == 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p.gwin) + base + rel);
      g[0] += v.x; g[1] += v.y;
    } else {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.gwin) + base + rel);
      g[0] += v.x; g[1] += v.y; g[2] += v.z;
    }
  }
}

// components per cell of a window field: 2 (float2) in 2-D, 4 (float4, last unused) in 3-D
template <int DIM> struct WinVec { static constexpr int NC = (DIM == 2) ? 2 : 4; };

// moments -> (Guo velocity shift) -> equilibrium -> collision -> forcing, on one cell in registers.
// Order of operations: examples/2d/poiseuille_channel.py:80-148 (EDM uses the uncorrected velocity,
// Guo shifts u by g/(2 rho) before the equilibrium).
template <int DIM, int COLL>
__device__ __forceinline__ void collide_cell(float (&f)[Lat<DIM>::Q], const float (&g)[Lat<DIM>::D], int forcing,
                                             const Relax& rx, const MrtMats<DIM, mats_kind(COLL)>& mm) {
  using L = Lat<DIM>;
  float rho, u[L::D];
  [[maybe_unused]] float feq[L::Q];
  moments<DIM>(f, rho, u);
  // a zero force contributes exactly nothing (u + 0, f + w*0): skip the work -- bit-identical
  bool has_g = false;
#pragma unroll
  for (int d = 0; d < L::D; ++d) has_g = has_g || (g[d] != 0.f);
  if (!has_g) forcing = VSB_FORCE_NONE;
  if (forcing == VSB_FORCE_GUO) {
#pragma unroll
    for (int d = 0; d < L::D; ++d) u[d] += g[d] * 0.5f / rho;
  }
  if constexpr (COLL == VSB_COLL_KBC) {
    float feq0, A[Pairs<DIM>::NP], B[Pairs<DIM>::NP];
    equilibrium_pairs<DIM>(rho, u, feq0, A, B);
    collide_kbc_pairs<DIM>(f, feq0, A, B, rx);
  } else if constexpr (COLL == VSB_COLL_MRT_MOMENT) {
    // f + A (feq - f) [+ B G with B = I - A/2]  =  f [+ G] + A (feq - f [- G/2]),  A applied in moment space
    using P = Pairs<DIM>;
    float feq0, A[P::NP], B[P::NP], xb[P::NP], xa[P::NP];
    equilibrium_pairs<DIM>(rho, u, feq0, A, B);
#pragma unroll
    for (int k = 0; k < P::NP; ++k) {
      const int q = P::q(k), o = L::opp(q);
      xb[k] = 2.0f * A[k] - (f[q] + f[o]);
      xa[k] = 2.0f * B[k] - (f[q] - f[o]);
    }
    if (forcing != VSB_FORCE_NONE) {
      float G0, Hs[P::NP], Ha[P::NP];
      guo_term_pairs<DIM>(g, u, G0, Hs, Ha);
      f[0] += G0;
#pragma unroll
      for (int k = 0; k < P::NP; ++k) {
        const int q = P::q(k), o = L::opp(q);
        f[q] += Hs[k] + Ha[k];
        f[o] += Hs[k] - Ha[k];
        if (forcing == VSB_FORCE_GUO) { xb[k] -= Hs[k]; xa[k] -= Ha[k]; }   // half of G_q + G_opp, half of G_q - G_opp
      }
    }
    float y0, yb[P::NP], ya[P::NP];
    moment_op3_apply(mm.mo, xb, xa, y0, yb, ya);
    f[0] += y0;
#pragma unroll
    for (int k = 0; k < P::NP; ++k) {
      const int q = P::q(k), o = L::opp(q);
      f[q] += yb[k] + ya[k];
      f[o] += yb[k] - ya[k];
    }
    return;
  } else if constexpr (COLL == VSB_COLL_MRT_SPLIT) {
    float feq0, A[Pairs<DIM>::NP], B[Pairs<DIM>::NP];
    equilibrium_pairs<DIM>(rho, u, feq0, A, B);
    collide_mrt_split<DIM>(f, feq0, A, B, mm.As);
  } else if constexpr (COLL == VSB_COLL_MRT) {
    equilibrium<DIM>(rho, u, feq);
    collide_mrt<DIM>(f, feq, mm.A);
  } else {
    equilibrium<DIM>(rho, u, feq);
    if constexpr (COLL == VSB_COLL_BGK) collide_bgk<DIM>(f, feq, rx);
    if constexpr (COLL == VSB_COLL_REG) collide_reg<DIM>(f, feq, rx);
  }
  if (forcing != VSB_FORCE_NONE) {
    float G[L::Q];
    guo_term<DIM>(g, u, G);
    if (forcing == VSB_FORCE_EDM) {
#pragma unroll
      for (int q = 0; q < L::Q; ++q) f[q] += G[q];
    } else {
      if constexpr (COLL == VSB_COLL_MRT_SPLIT) {
        float xs[Pairs<DIM>::NP], xa[Pairs<DIM>::NP];
#pragma unroll
        for (int k = 0; k < Pairs<DIM>::NP; ++k) {
          const int q = Pairs<DIM>::q(k), o = L::opp(q);
          xs[k] = G[q] + G[o];
          xa[k] = G[q] - G[o];
        }
        split_matvec_add<DIM>(f, mm.Bs, G[0], xs, xa);
      } else if constexpr (COLL == VSB_COLL_MRT) {
        matvec_add<DIM>(f, mm.B, G);
      } else {
#pragma unroll
        for (int q = 0; q < L::Q; ++q) f[q] += G[q] * rx.guo_scale;
      }
    }
  }
}

// Fill StepParams from the ABI struct (validation included).  Defined in vsb_step.cu.
template <int DIM> int fill_params(const VsbStepArgs& a, StepParams<DIM>& p);

}  // namespace vsb
