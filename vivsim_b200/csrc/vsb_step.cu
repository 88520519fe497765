// Fused IB-LBM time step: one pass that pulls the streamed populations,
// takes moments, applies collision (BGK / MRT / KBC / regularised) and Guo / EDM forcing and
// writes the post-collision state -- 1 read + 1 write of f per cell per step (72 B D2Q9, 152 B D3Q19).
//
// Access pattern.  SoA planes f[q][..][i2], i2 contiguous.  A thread owns VEC consecutive i2 cells
// and moves them with one 64/128-bit load/store per population.  Populations with a velocity
// component along i2 need the plane shifted by one element: the aligned vector is loaded and the
// missing element comes from the neighbouring lane by warp shuffle; only lanes at a warp or row
// edge issue one extra scalar load (which also performs the periodic wrap).  Shifts along the
// other axes only change the row that is read, so every access stays aligned and coalesced.
#include <utility>

#include "vsb_common.cuh"
#include "vsb_internal.h"

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

template <int DIM> struct StepParams {
  int n0, n1, n2;
  int r_begin, r_end;   // rows of array axis A0 (the slowest real axis) to update
  const float* fin;
  float* fout;
  int do_stream, do_collide, forcing;
  Relax rx;
  float g0[3];
  const float* gwin;
  int worg[3], wsz[3];
  const VsbBodyState* body;
  const uint8_t* mask;
};

template <int DIM, bool USED> struct MrtMats { Matrix<Lat<DIM>::Q> A, B; };
template <int DIM> struct MrtMats<DIM, false> {};

template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int VEC>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[VEC]) {
  using T = typename VecT<VEC>::type;
  const T t = __ldg(reinterpret_cast<const T*>(p));
  if constexpr (VEC == 1) v[0] = t;
  if constexpr (VEC == 2) { v[0] = t.x; v[1] = t.y; }
  if constexpr (VEC == 4) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[VEC]) {
  using T = typename VecT<VEC>::type;
  T t;
  if constexpr (VEC == 1) t = v[0];
  if constexpr (VEC == 2) { t.x = v[0]; t.y = v[1]; }
  if constexpr (VEC == 4) { t.x = v[0]; t.y = v[1]; t.z = v[2]; t.w = v[3]; }
  *reinterpret_cast<T*>(p) = t;
}

__device__ __forceinline__ int wrap(int i, int n) {
  i += (i < 0) ? n : 0;
  i -= (i >= n) ? n : 0;
  return i;
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
    int wcells = 1;
#pragma unroll
    for (int d = 0; d < L::D; ++d) wcells *= p.wsz[d];
#pragma unroll
    for (int d = 0; d < L::D; ++d) g[d] += p.gwin[d * wcells + base + rel];
  }
}

// moments -> (Guo velocity shift) -> equilibrium -> collision -> forcing, on one cell in registers.
// Order of operations: examples/2d/poiseuille_channel.py:80-148 (EDM uses the uncorrected velocity,
// Guo shifts u by g/(2 rho) before the equilibrium).
template <int DIM, int COLL>
__device__ __forceinline__ void collide_cell(float (&f)[Lat<DIM>::Q], const float (&g)[Lat<DIM>::D], int forcing,
                                             const Relax& rx, const MrtMats<DIM, COLL == VSB_COLL_MRT>& mm) {
  using L = Lat<DIM>;
  float rho, u[L::D], feq[L::Q];
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
  equilibrium<DIM>(rho, u, feq);
  if constexpr (COLL == VSB_COLL_BGK) collide_bgk<DIM>(f, feq, rx);
  if constexpr (COLL == VSB_COLL_KBC) collide_kbc<DIM>(f, feq, rx);
  if constexpr (COLL == VSB_COLL_REG) collide_reg<DIM>(f, feq, rx);
  if constexpr (COLL == VSB_COLL_MRT) collide_mrt<DIM>(f, feq, mm.A);
  if (forcing != VSB_FORCE_NONE) {
    float G[L::Q];
    guo_term<DIM>(g, u, G);
    if (forcing == VSB_FORCE_EDM) {
#pragma unroll
      for (int q = 0; q < L::Q; ++q) f[q] += G[q];
    } else {
      if constexpr (COLL == VSB_COLL_MRT) {
        matvec_add<DIM>(f, mm.B, G);
      } else {
#pragma unroll
        for (int q = 0; q < L::Q; ++q) f[q] += G[q] * rx.guo_scale;
      }
    }
  }
}

template <int DIM, int COLL, int VEC>
__global__ void __launch_bounds__(256) k_step(const StepParams<DIM> p, const MrtMats<DIM, COLL == VSB_COLL_MRT> mm) {
  using L = Lat<DIM>;
  constexpr int Q = L::Q;
  const int nv = p.n2 / VEC;
  const long long rows = (DIM == 2) ? (long long)(p.r_end - p.r_begin) : (long long)(p.r_end - p.r_begin) * p.n1;
  const long long total = rows * nv;
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = gid < total;
  if (!active) gid = total - 1;  // keep the lane in the shuffles with valid addresses; it stores nothing
  const int j = (int)(gid % nv);
  const long long row = gid / nv;
  const int i0 = (DIM == 2) ? 0 : p.r_begin + (int)(row / p.n1);
  const int i1 = (DIM == 2) ? p.r_begin + (int)row : (int)(row % p.n1);
  const int i2 = j * VEC;
  const int lane = threadIdx.x & 31;
  const long long ncell = (long long)p.n0 * p.n1 * p.n2;
  const long long cell = ((long long)i0 * p.n1 + i1) * p.n2 + i2;

  float f[VEC][Q];
  if (p.do_stream) {
    static_for<Q>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      constexpr int c2 = L::c(q, 2);
      const int s0 = (DIM == 2) ? 0 : wrap(i0 - L::c(q, 0), p.n0);
      const int s1 = wrap(i1 - L::c(q, 1), p.n1);
      const float* __restrict__ src = p.fin + q * ncell + ((long long)s0 * p.n1 + s1) * p.n2;
      if constexpr (VEC == 1) {
        f[0][q] = __ldg(src + wrap(i2 - c2, p.n2));
      } else {
        float v[VEC];
        load_vec<VEC>(src + i2, v);
        if constexpr (c2 == 0) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) f[k][q] = v[k];
        } else if constexpr (c2 > 0) {  // new[i2 + k] = old[i2 + k - 1]
          float left = __shfl_up_sync(0xffffffffu, v[VEC - 1], 1);
          if (lane == 0 || j == 0) left = __ldg(src + (i2 == 0 ? p.n2 - 1 : i2 - 1));
          f[0][q] = left;
#pragma unroll
          for (int k = 1; k < VEC; ++k) f[k][q] = v[k - 1];
        } else {                        // new[i2 + k] = old[i2 + k + 1]
          float right = __shfl_down_sync(0xffffffffu, v[0], 1);
          if (lane == 31 || j == nv - 1) right = __ldg(src + (i2 + VEC == p.n2 ? 0 : i2 + VEC));
          f[VEC - 1][q] = right;
#pragma unroll
          for (int k = 0; k < VEC - 1; ++k) f[k][q] = v[k + 1];
        }
      }
    });
    if (p.mask) {  // obstacle_bounce_back on the streamed populations (lbm/boundary/bb.py:110)
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if (p.mask[cell + k]) {
          float t[Q];
#pragma unroll
          for (int q = 0; q < Q; ++q) t[q] = f[k][q];
#pragma unroll
          for (int q = 0; q < Q; ++q) f[k][q] = t[L::opp(q)];
        }
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      float v[VEC];
      load_vec<VEC>(p.fin + q * ncell + cell, v);
#pragma unroll
      for (int k = 0; k < VEC; ++k) f[k][q] = v[k];
    }
  }

  if (p.do_collide) {
    int worg[3] = {p.worg[0], p.worg[1], p.worg[2]};
    if (p.gwin && p.body) { worg[0] = p.body->origin[0]; worg[1] = p.body->origin[1]; worg[2] = p.body->origin[2]; }
    bool rows_inside;
    int wbase;
    window_rows<DIM>(p, worg, i0, i1, rows_inside, wbase);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float g[L::D];
      cell_force<DIM>(p, worg, rows_inside, wbase, i2 + k, g);
      collide_cell<DIM, COLL>(f[k], g, p.forcing, p.rx, mm);
    }
  }

  if (active) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      float v[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = f[k][q];
      store_vec<VEC>(p.fout + q * ncell + cell, v);
    }
  }
}

// ----------------------------------------------------------------------------- wall-line fix-up
// Post-streaming operations act on whole wall lines / faces in call order and may read the adjacent
// fluid layer (NEE, pressure wrappers), including cells an earlier operation already rewrote
// (SURVEY.md appendix A6).  To reproduce that exactly the affected layers are (1) overwritten in
// f_out with the streamed populations, (2) processed in place by the ordered operations, and
// (3) collided in place; every other cell keeps the fused kernel's result.
struct LineSet {
  int n;
  int axis[12], layer[12];
};

template <int DIM>
__device__ __forceinline__ bool line_cell(const StepParams<DIM>& p, const LineSet& ls, int id, long long k, int (&c)[3]) {
  const int n[3] = {p.n0, p.n1, p.n2};
  const int ax = ls.axis[id];
  const int ta = (ax == 0) ? 1 : 0, tb = (ax == 2) ? 1 : 2;
  if (k >= (long long)n[ta] * n[tb]) return false;
  c[ta] = (int)(k / n[tb]);
  c[tb] = (int)(k % n[tb]);
  c[ax] = ls.layer[id];
  if (c[Lat<DIM>::A0] < p.r_begin || c[Lat<DIM>::A0] >= p.r_end) return false;
  for (int e = 0; e < id; ++e)   // a cell shared by several layers belongs to the first one
    if (c[ls.axis[e]] == ls.layer[e]) return false;
  return true;
}

template <int DIM>
__global__ void k_lines_restream(const StepParams<DIM> p, const LineSet ls) {
  using L = Lat<DIM>;
  int c[3];
  if (!line_cell<DIM>(p, ls, blockIdx.y, (long long)blockIdx.x * blockDim.x + threadIdx.x, c)) return;
  const long long ncell = (long long)p.n0 * p.n1 * p.n2;
  const long long cell = ((long long)c[0] * p.n1 + c[1]) * p.n2 + c[2];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) {
    const int s0 = wrap(c[0] - L::c(q, 0), p.n0), s1 = wrap(c[1] - L::c(q, 1), p.n1), s2 = wrap(c[2] - L::c(q, 2), p.n2);
    p.fout[q * ncell + cell] = p.fin[q * ncell + ((long long)s0 * p.n1 + s1) * p.n2 + s2];
  }
}

template <int DIM>
__global__ void k_lines_mask(const StepParams<DIM> p, const LineSet ls, const uint8_t* __restrict__ mask) {
  using L = Lat<DIM>;
  int c[3];
  if (!line_cell<DIM>(p, ls, blockIdx.y, (long long)blockIdx.x * blockDim.x + threadIdx.x, c)) return;
  const long long ncell = (long long)p.n0 * p.n1 * p.n2;
  const long long cell = ((long long)c[0] * p.n1 + c[1]) * p.n2 + c[2];
  if (!mask[cell]) return;
  float t[L::Q];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) t[q] = p.fout[q * ncell + cell];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) p.fout[q * ncell + cell] = t[L::opp(q)];
}

template <int DIM, int COLL>
__global__ void k_lines_collide(const StepParams<DIM> p, const MrtMats<DIM, COLL == VSB_COLL_MRT> mm, const LineSet ls) {
  using L = Lat<DIM>;
  int c[3];
  if (!line_cell<DIM>(p, ls, blockIdx.y, (long long)blockIdx.x * blockDim.x + threadIdx.x, c)) return;
  const long long ncell = (long long)p.n0 * p.n1 * p.n2;
  const long long cell = ((long long)c[0] * p.n1 + c[1]) * p.n2 + c[2];
  float f[L::Q], g[L::D];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) f[q] = p.fout[q * ncell + cell];
  int worg[3] = {p.worg[0], p.worg[1], p.worg[2]};
  if (p.gwin && p.body) { worg[0] = p.body->origin[0]; worg[1] = p.body->origin[1]; worg[2] = p.body->origin[2]; }
  bool rows_inside;
  int wbase;
  window_rows<DIM>(p, worg, c[0], c[1], rows_inside, wbase);
  cell_force<DIM>(p, worg, rows_inside, wbase, c[2], g);
  collide_cell<DIM, COLL>(f, g, p.forcing, p.rx, mm);
#pragma unroll
  for (int q = 0; q < L::Q; ++q) p.fout[q * ncell + cell] = f[q];
}

// u on the IB window from the streamed (and masked) state: feeds vsb_ib_mdf.
template <int DIM>
__global__ void k_window_moments(const StepParams<DIM> p, int follow, float o0x, float o0y, float o0z, float* __restrict__ u_win,
                                 VsbBodyState* body) {
  using L = Lat<DIM>;
  const float o0[3] = {o0x, o0y, o0z};
  int org[3] = {0, 0, 0};
  const int n[3] = {p.n0, p.n1, p.n2};
#pragma unroll
  for (int d = 0; d < L::D; ++d) {
    const float shifted = o0[d] + ((body && follow) ? body->d[d] : 0.f);
    if (follow == 2) {   // clip(floor(.)): examples/3d/oscillating_cylinder.py:241-243
      int o = (int)floorf(shifted);
      o = max(0, min(o, n[d + L::A0] - p.wsz[d]));
      org[d] = o;
    } else {             // astype(int32): examples/2d/vortex_induced_vibration.py:104-105
      org[d] = (int)shifted;
    }
  }
  long long wcells = 1;
#pragma unroll
  for (int d = 0; d < L::D; ++d) wcells *= p.wsz[d];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 && body) { body->origin[0] = org[0]; body->origin[1] = org[1]; body->origin[2] = org[2]; }
  if (t >= wcells) return;
  int rel[3] = {0, 0, 0};
  long long r = t;
#pragma unroll
  for (int d = L::D - 1; d >= 0; --d) { rel[d] = (int)(r % p.wsz[d]); r /= p.wsz[d]; }
  int c[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < L::D; ++d) c[d + L::A0] = org[d] + rel[d];
  const long long ncell = (long long)p.n0 * p.n1 * p.n2;
  float f[L::Q];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) {
    const int s0 = p.do_stream ? wrap(c[0] - L::c(q, 0), p.n0) : c[0];
    const int s1 = p.do_stream ? wrap(c[1] - L::c(q, 1), p.n1) : c[1];
    const int s2 = p.do_stream ? wrap(c[2] - L::c(q, 2), p.n2) : c[2];
    f[q] = p.fin[q * ncell + ((long long)s0 * p.n1 + s1) * p.n2 + s2];
  }
  if (p.do_stream && p.mask && p.mask[((long long)c[0] * p.n1 + c[1]) * p.n2 + c[2]]) {
    float tq[L::Q];
#pragma unroll
    for (int q = 0; q < L::Q; ++q) tq[q] = f[q];
#pragma unroll
    for (int q = 0; q < L::Q; ++q) f[q] = tq[L::opp(q)];
  }
  float rho, u[L::D];
  moments<DIM>(f, rho, u);
#pragma unroll
  for (int d = 0; d < L::D; ++d) u_win[d * wcells + t] = u[d];
}

// ----------------------------------------------------------------------------- host side
template <int DIM>
static int fill_params(const VsbStepArgs& a, StepParams<DIM>& p) {
  grid_axes(a.grid, p.n0, p.n1, p.n2);
  VSB_REQUIRE(p.n0 > 0 && p.n1 > 0 && p.n2 > 0, "vsb_step: bad grid");
  const int nrows = (DIM == 2) ? p.n1 : p.n0;
  p.r_begin = a.row_begin;
  p.r_end = a.row_end > 0 ? a.row_end : nrows;
  VSB_REQUIRE(0 <= p.r_begin && p.r_begin < p.r_end && p.r_end <= nrows, "vsb_step: bad row range [%d, %d) of %d",
              a.row_begin, a.row_end, nrows);
  VSB_REQUIRE(a.f_in && a.f_out && a.f_in != a.f_out, "vsb_step: f_in / f_out must be distinct non-null buffers");
  p.fin = a.f_in; p.fout = a.f_out;
  p.do_stream = a.do_stream; p.do_collide = a.do_collide; p.forcing = a.forcing;
  VSB_REQUIRE(a.forcing >= VSB_FORCE_NONE && a.forcing <= VSB_FORCE_GUO, "vsb_step: unknown forcing %d", a.forcing);
  p.rx = make_relax(a.omega);
  for (int d = 0; d < 3; ++d) { p.g0[d] = a.g_uniform[d]; p.worg[d] = a.win_origin[d]; p.wsz[d] = a.win_size[d]; }
  p.gwin = a.g_win; p.body = a.body; p.mask = nullptr;
  if (a.g_win) for (int d = 0; d < DIM; ++d) VSB_REQUIRE(a.win_size[d] > 0, "vsb_step: empty force window");
  int n_mask = 0;
  for (int i = 0; i < a.n_post; ++i)
    if (a.post[i].kind == VSB_POST_MASK) { p.mask = a.post[i].mask; ++n_mask; }
  VSB_REQUIRE(n_mask <= 1, "vsb_step: at most one obstacle mask per step");
  VSB_REQUIRE(n_mask == 0 || p.mask, "vsb_step: mask op without a mask");
  return VSB_OK;
}

template <int DIM, int COLL>
static int step_impl(const VsbStepArgs& a, cudaStream_t s) {
  constexpr bool MRT = (COLL == VSB_COLL_MRT);
  constexpr int Q = Lat<DIM>::Q;
  StepParams<DIM> p;
  if (int rc = fill_params<DIM>(a, p)) return rc;
  MrtMats<DIM, MRT> mm;
  if constexpr (MRT) {
    VSB_REQUIRE(a.mrt_op_host || !a.do_collide, "vsb_step: MRT collision needs mrt_op_host");
    VSB_REQUIRE(a.forcing != VSB_FORCE_GUO || a.mrt_fop_host || !a.do_collide, "vsb_step: MRT + Guo forcing needs mrt_fop_host");
    for (int i = 0; i < Q * Q; ++i) {
      mm.A.a[i] = a.mrt_op_host ? a.mrt_op_host[i] : 0.f;
      mm.B.a[i] = a.mrt_fop_host ? a.mrt_fop_host[i] : 0.f;
    }
  }
  int vec = a.vec;
  if (vec == 0) vec = (DIM == 2) ? 4 : 2;
  while (vec > 1 && (p.n2 % vec != 0 || ((uintptr_t)a.f_in % (4 * vec)) || ((uintptr_t)a.f_out % (4 * vec)))) vec >>= 1;
  VSB_REQUIRE(vec == 1 || vec == 2 || vec == 4, "vsb_step: vec must be 0, 1, 2 or 4");
  const long long rows = (DIM == 2) ? (long long)(p.r_end - p.r_begin) : (long long)(p.r_end - p.r_begin) * p.n1;
  const long long total = rows * (p.n2 / vec);
  constexpr int kBlock = 256;
  const unsigned nb = blocks_for(total, kBlock);
  if (vec == 4) k_step<DIM, COLL, 4><<<nb, kBlock, 0, s>>>(p, mm);
  else if (vec == 2) k_step<DIM, COLL, 2><<<nb, kBlock, 0, s>>>(p, mm);
  else k_step<DIM, COLL, 1><<<nb, kBlock, 0, s>>>(p, mm);
  VSB_LAUNCH_CHECK("vsb_step (fused kernel)");

  if (a.n_post == 0 || !a.do_stream) return VSB_OK;
  // layers touched by face operations: wall layer and adjacent fluid layer of each face
  LineSet ls;
  ls.n = 0;
  const int n[3] = {p.n0, p.n1, p.n2};
  for (int i = 0; i < a.n_post; ++i) {
    const VsbPostOp& op = a.post[i];
    if (op.kind == VSB_POST_MASK) continue;
    VSB_REQUIRE(op.loc >= 0 && op.loc < 2 * DIM, "vsb_step: loc %d is not a face of a %d-D grid", op.loc, DIM);
    const int ax = op.loc / 2 + Lat<DIM>::A0;
    const bool low = (op.loc % 2 == 0);
    const int lo = (ax == Lat<DIM>::A0) ? p.r_begin : 0, hi = (ax == Lat<DIM>::A0) ? p.r_end : n[ax];
    VSB_REQUIRE(hi - lo >= 2, "vsb_step: fewer than 2 layers along the normal of face %d", op.loc);
    const int layers[2] = {low ? lo : hi - 1, low ? lo + 1 : hi - 2};
    for (int l = 0; l < 2; ++l) {
      bool seen = false;
      for (int e = 0; e < ls.n; ++e) seen = seen || (ls.axis[e] == ax && ls.layer[e] == layers[l]);
      if (!seen) { ls.axis[ls.n] = ax; ls.layer[ls.n] = layers[l]; ++ls.n; }
    }
  }
  if (ls.n == 0) return VSB_OK;
  long long max_face = 0;
  for (int e = 0; e < ls.n; ++e) {
    const long long nf = (long long)n[0] * n[1] * n[2] / n[ls.axis[e]];
    max_face = nf > max_face ? nf : max_face;
  }
  const dim3 grid(blocks_for(max_face, 128), ls.n);
  k_lines_restream<DIM><<<grid, 128, 0, s>>>(p, ls);
  VSB_LAUNCH_CHECK("vsb_step (restream wall layers)");
  for (int i = 0; i < a.n_post; ++i) {
    const VsbPostOp& op = a.post[i];
    if (op.kind == VSB_POST_MASK) {
      k_lines_mask<DIM><<<grid, 128, 0, s>>>(p, ls, op.mask);
      VSB_LAUNCH_CHECK("vsb_step (mask on wall layers)");
    } else {
      if (int rc = launch_post_op(DIM, p.n0, p.n1, p.n2, op, a.f_in, a.f_out, s, p.r_begin, p.r_end)) return rc;
    }
  }
  if (a.do_collide) {
    k_lines_collide<DIM, COLL><<<grid, 128, 0, s>>>(p, mm, ls);
    VSB_LAUNCH_CHECK("vsb_step (collide wall layers)");
  }
  return VSB_OK;
}

template <int DIM>
static int step_dispatch(const VsbStepArgs& a, cudaStream_t s) {
  switch (a.collision) {
    case VSB_COLL_BGK: return step_impl<DIM, VSB_COLL_BGK>(a, s);
    case VSB_COLL_MRT: return step_impl<DIM, VSB_COLL_MRT>(a, s);
    case VSB_COLL_KBC: return step_impl<DIM, VSB_COLL_KBC>(a, s);
    case VSB_COLL_REG: return step_impl<DIM, VSB_COLL_REG>(a, s);
    default: VSB_REQUIRE(false, "vsb_step: unknown collision %d", a.collision);
  }
}

template <int DIM>
static int window_impl(const VsbStepArgs& a, int follow, const float* o0, float* u_win, VsbBodyState* body, cudaStream_t s) {
  StepParams<DIM> p;
  VsbStepArgs b = a;
  if (!b.f_out) b.f_out = u_win;  // unused by this kernel; only has to differ from f_in
  if (int rc = fill_params<DIM>(b, p)) return rc;
  long long wcells = 1;
  for (int d = 0; d < DIM; ++d) {
    VSB_REQUIRE(a.win_size[d] > 0, "vsb_ib_window_moments: empty window");
    wcells *= a.win_size[d];
  }
  k_window_moments<DIM><<<blocks_for(wcells, 128), 128, 0, s>>>(p, follow, o0[0], o0[1], o0[2], u_win, body);
  VSB_LAUNCH_CHECK("vsb_ib_window_moments");
  return VSB_OK;
}

}  // namespace vsb

using namespace vsb;

extern "C" {

int vsb_step(const VsbStepArgs* args, vsb_stream_t stream) {
  VSB_REQUIRE(args != nullptr, "vsb_step: null args");
  VSB_REQUIRE(args->grid.dim == 2 || args->grid.dim == 3, "dim must be 2 or 3, got %d", args->grid.dim);
  VSB_REQUIRE(args->n_post >= 0 && (args->n_post == 0 || args->post), "vsb_step: bad post list");
  VSB_REQUIRE(args->do_stream || args->do_collide, "vsb_step: nothing to do");
  return args->grid.dim == 2 ? step_dispatch<2>(*args, (cudaStream_t)stream) : step_dispatch<3>(*args, (cudaStream_t)stream);
}

int vsb_ib_window_moments(const VsbStepArgs* args, int follow, const float win_origin0[3], float* u_win, VsbBodyState* body,
                          vsb_stream_t stream) {
  VSB_REQUIRE(args && win_origin0 && u_win, "vsb_ib_window_moments: null argument");
  VSB_REQUIRE(args->grid.dim == 2 || args->grid.dim == 3, "dim must be 2 or 3, got %d", args->grid.dim);
  VSB_REQUIRE(follow >= 0 && follow <= 2, "follow must be 0, 1 or 2");
  VSB_REQUIRE(follow == 0 || body, "a moving window needs a body state");
  return args->grid.dim == 2 ? window_impl<2>(*args, follow, win_origin0, u_win, body, (cudaStream_t)stream)
                             : window_impl<3>(*args, follow, win_origin0, u_win, body, (cudaStream_t)stream);
}

}  // extern "C"
