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
#include <cstdlib>
#include <algorithm>
#include <cmath>

#include "vsb_step.cuh"

namespace vsb {

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

template <int DIM, int COLL, bool HALO = false>
__device__ __forceinline__ void edge_block(const StepParams<DIM>& p, const MrtMats<DIM, mats_kind(COLL)>& mm);

// Block size and register budget per lattice and collision model, measured on B200 (scripts/tune_variants.py,
// profiles/r01_summary.md section 4):
//   D3Q19: 128 registers either way; BGK / KBC / regularised and MRT in moment space run best as 4 CTAs of 128 threads
//          with 2 cells per thread, MRT in matrix form as 2 CTAs of 256 threads with 4 cells per thread.
//   D2Q9:  256 threads, 4 cells per thread; 64 registers (4 CTAs) for BGK / regularised, 80 (3 CTAs) for KBC / MRT.
// VSB_STEP3D_THREADS / VSB_STEP3D_CTAS override the D3Q19 choice for the tuning builds.
__host__ __device__ constexpr bool is_matrix_mrt(int coll) { return coll == VSB_COLL_MRT || coll == VSB_COLL_MRT_SPLIT; }
template <int DIM, int COLL> constexpr int step_max_threads() {
#ifdef VSB_STEP3D_THREADS
  return DIM == 3 ? VSB_STEP3D_THREADS : 256;
#else
  return (DIM == 3 && !is_matrix_mrt(COLL)) ? 128 : 256;
#endif
}
template <int DIM, int COLL> constexpr int step_min_ctas() {
#ifdef VSB_STEP3D_CTAS
  if (DIM == 3) return VSB_STEP3D_CTAS;
#endif
  return DIM == 3 ? (is_matrix_mrt(COLL) ? 2 : 4) : ((COLL == VSB_COLL_KBC || is_mrt(COLL)) ? 3 : 4);
}
template <int DIM, int COLL> constexpr int step_default_vec() { return (DIM == 3 && !is_matrix_mrt(COLL)) ? 2 : 4; }

// Populations of one cell (or VEC cells) that cross the cut next to an edge row go to the neighbour's ghost row as
// well (fused halo send, StepParams::halo): row r_begin sends its c = -1 populations to the LEFT neighbour's ghost row
// r_end, row r_end - 1 its c = +1 populations to the RIGHT neighbour's ghost row r_begin - 1.
template <int DIM>
__device__ __forceinline__ bool halo_target(const StepParams<DIM>& p, int i_slow, int dir, float*& base, long long& shift) {
  const long long rowstride = (DIM == 2) ? (long long)p.n2 : (long long)p.n1 * p.n2;
  if (dir < 0 && i_slow == p.r_begin) { base = p.halo.left_state; shift = (long long)(p.r_end - p.r_begin) * rowstride; return true; }
  if (dir > 0 && i_slow == p.r_end - 1) { base = p.halo.right_state; shift = -(long long)(p.r_end - p.r_begin) * rowstride; return true; }
  return false;
}

template <int DIM, int COLL, int VEC, bool HALO>
__device__ __forceinline__ void step_body(const StepParams<DIM>& p, const MrtMats<DIM, mats_kind(COLL)>& mm) {
  using L = Lat<DIM>;
  constexpr int Q = L::Q;
  if (blockIdx.x >= p.nb_bulk) {   // blocks appended after the bulk: wall layers of the face operations (edges = 2)
    edge_block<DIM, COLL, HALO>(p, mm);
    return;
  }
  const int nv = p.n2 / VEC;
  int worg[3];
  window_origin<DIM>(p, worg);
  // rows of the slowest axis handled by this launch: all of [r_begin, r_end), or only / all but the band of the force
  // window, so that the bulk can run while the IB kernels still produce the window's force.
  const int band_lo = max(p.s_begin, worg[0]), band_hi = min(p.s_end, worg[0] + p.wsz[0]);
  // The band is the window's (x, y) footprint over all z in 3-D (band 2 enumerates wsz[1] lines per x plane) and, in
  // 2-D, the window's x rows restricted to the vector groups that overlap its y-range (band 2 enumerates p.nvb groups
  // per row, a launch constant that covers any alignment of the moving window).
  const bool box = DIM == 3 && p.band == 2;
  const bool box2 = DIM == 2 && p.band == 2;
  const int jb_lo = (DIM == 2) ? worg[1] / VEC : 0;                               // vector groups of the window's y-range
  const int jb_hi = (DIM == 2) ? (worg[1] + p.wsz[1] - 1) / VEC : 0;              // (inclusive)
  const int row0 = (p.band == 2) ? band_lo : p.s_begin;
  const int nrow = p.edge_rows ? 2 : ((p.band == 2) ? max(band_hi - band_lo, 0) : p.s_end - p.s_begin);
  const int lines = box ? p.wsz[1] : p.n1;            // z lines per x plane handled by this launch
  const int nvr = box2 ? p.nvb : nv;                  // vector groups per row enumerated by this launch
  const unsigned total = (unsigned)((DIM == 2) ? nrow : nrow * lines) * (unsigned)nvr;   // < 2^31, checked by the host
  unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = gid < total;      // the lane addresses cells of the grid (it loads them, so that its neighbours' shuffles
  if (!valid) gid = 0;           // are right); `active` below: it also computes and stores them
  const unsigned row = fast_div(gid, box2 ? p.div_nvb : p.div_nv);
  const int jr = (int)(gid - row * (unsigned)nvr);    // group counter within the enumerated part of the row
  const int j = box2 ? jb_lo + jr : jr;
  if (box2 && j >= nv) valid = false;
  const int rx = (DIM == 2) ? (int)row : (int)fast_div(row, box ? p.div_w1 : p.div_n1);   // row counter along the slowest axis
  const int ix0 = p.edge_rows ? (rx == 0 ? p.r_begin : p.r_end - 1) : row0 + rx;
  const int i0 = (DIM == 2) ? 0 : ix0;
  const int i1 = (DIM == 2) ? ix0 : (int)row - rx * lines + (box ? worg[1] : 0);
  const int i2 = (valid ? j : 0) * VEC;
  const int lane = threadIdx.x & 31;
  const int n12 = p.n1 * p.n2;
  const long long ncell = (long long)p.n0 * n12;
  const long long cell = (long long)i0 * n12 + (i1 * p.n2 + i2);
  bool active = valid;
  {
    const int ix = (DIM == 2) ? i1 : i0;
    const bool in_band = ix >= band_lo && ix < band_hi &&
                         (DIM == 2 ? (j >= jb_lo && j <= jb_hi) : ((unsigned)(i1 - worg[1]) < (unsigned)p.wsz[1]));
    if (p.band == 1 && in_band) active = false;
    if (box2 && !in_band) active = false;             // padding groups of the enumeration
    // wall layers owned by the fused wall kernel (at most two)
    if (p.n_skip > 0 && (p.skip_axis[0] ? i1 : i0) == p.skip_layer[0]) active = false;
    if (p.n_skip > 1 && (p.skip_axis[1] ? i1 : i0) == p.skip_layer[1]) active = false;
  }
  // A warp without a single active lane leaves at once: nobody waits for its shuffles (they are warp-wide only), and
  // it must not load anything -- the rows excluded by `band` make up whole blocks whose lanes would otherwise all
  // read (and L2-prefetch) the same few cache lines, which serialises in one L2 slice (measured: +45 % on the bulk
  // launch of the 256^3 sphere case).
  if (__all_sync(0xffffffffu, !active)) return;
  // In 3-D (and for whole rows in 2-D) activity is uniform along a row, so an inactive lane never feeds an active
  // one.  The 2-D window box cuts rows: there an inactive lane next to an active one must still hold its real cells.
  const bool loads = (DIM == 2) ? valid : active;

  float f[VEC][Q];
  if (p.do_stream) {
    // Element offsets of the pulled rows relative to this thread's own cells, periodic in the array extents:
    // index 0 for c = +1 (source row i - 1), 1 for c = 0, 2 for c = -1 (source row i + 1).  |offset| < ncell < 2^31.
    int d0[3] = {0, 0, 0}, d1[3];
    d1[0] = (i1 == 0 ? p.n1 - 1 : -1) * p.n2;
    d1[1] = 0;
    d1[2] = (i1 == p.n1 - 1 ? 1 - p.n1 : 1) * p.n2;
    if constexpr (DIM == 3) {
      d0[0] = (i0 == 0 ? p.n0 - 1 : -1) * n12;
      d0[2] = (i0 == p.n0 - 1 ? 1 - p.n0 : 1) * n12;
    }
    const float* __restrict__ own = p.fin + cell;
    if (!loads) {   // idle lanes still take part in the shuffles: they all read the first cells of each plane (cached)
      own = p.fin;
#pragma unroll
      for (int k = 0; k < 3; ++k) { d0[k] = 0; d1[k] = 0; }
    }
    auto source = [&](auto qc) -> const float* {
      constexpr int q = decltype(qc)::value;
      return own + q * ncell + (d0[1 - L::c(q, 0)] + d1[1 - L::c(q, 1)]);
    };
    if (p.prefetch_blocks > 0) {
      // The block that will follow this one on the SM is about one resident grid ahead; its cells lie (nearly always)
      // at the same relative offsets.  Asking L2 for them now turns its DRAM latency into L2 latency.
      const unsigned ahead = (unsigned)p.prefetch_blocks * blockDim.x;
      if (active && gid + ahead < total) {
        const float* __restrict__ nxt = own + (long long)ahead * VEC;
        static_for<Q>([&](auto qc) {
          constexpr int q = decltype(qc)::value;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + q * ncell + (d0[1 - L::c(q, 0)] + d1[1 - L::c(q, 1)])));
        });
      }
    }
    if constexpr (VEC == 1) {
      static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        constexpr int c2 = L::c(q, 2);
        const int sh = (c2 > 0) ? (i2 == 0 ? p.n2 - 1 : -1) : ((c2 < 0) ? (i2 == p.n2 - 1 ? 1 - p.n2 : 1) : 0);
        f[0][q] = __ldg(source(qc) + (loads ? sh : 0));
      });
    } else {
      // all aligned vector loads first (one predicated block), then the one-element shifts of the populations that
      // move along the contiguous axis
      float v[Q][VEC];
      static_for<Q>([&](auto qc) { load_vec<VEC>(source(qc), v[decltype(qc)::value]); });
      // the neighbouring lane holds the neighbouring cells only inside one enumerated row
      const bool first = active && (lane == 0 || jr == 0), last = active && (lane == 31 || jr == nvr - 1 || j == nv - 1);
      static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        constexpr int c2 = L::c(q, 2);
        if constexpr (c2 == 0) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) f[k][q] = v[q][k];
        } else if constexpr (c2 > 0) {  // new[i2 + k] = old[i2 + k - 1]
          float left = __shfl_up_sync(0xffffffffu, v[q][VEC - 1], 1);
          if (first) left = __ldg(source(qc) + (i2 == 0 ? p.n2 - 1 : -1));
          f[0][q] = left;
#pragma unroll
          for (int k = 1; k < VEC; ++k) f[k][q] = v[q][k - 1];
        } else {                        // new[i2 + k] = old[i2 + k + 1]
          float right = __shfl_down_sync(0xffffffffu, v[q][0], 1);
          if (last) right = __ldg(source(qc) + (i2 + VEC == p.n2 ? VEC - p.n2 : VEC));
          f[VEC - 1][q] = right;
#pragma unroll
          for (int k = 0; k < VEC - 1; ++k) f[k][q] = v[q][k + 1];
        }
      });
    }
    if (p.mask && active) {  // obstacle_bounce_back on the streamed populations (lbm/boundary/bb.py:110)
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
    const float* __restrict__ own = p.fin + (active ? cell : 0);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      float v[VEC];
      load_vec<VEC>(own + q * ncell, v);
#pragma unroll
      for (int k = 0; k < VEC; ++k) f[k][q] = v[k];
    }
  }

  if (p.do_collide && active) {
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
    const bool send = HALO && (p.halo.mode & 2) != 0;
    const int i_slow = (DIM == 2) ? i1 : i0;
    static_for<Q>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      float v[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = f[k][q];
      store_vec<VEC>(p.fout + q * ncell + cell, v);
      constexpr int cx = L::c(q, L::A0);
      if constexpr (HALO && cx != 0) {
        float* base;
        long long shift;
        if (send && halo_target<DIM>(p, i_slow, cx, base, shift)) store_vec<VEC>(base + q * ncell + cell + shift, v);
      }
    });
  }
}

// HALO = true: the variant launched for the two edge rows of a slab with the fused hand-shake (VsbStepArgs.halo); the
// bulk kernels carry none of that code (it cost the 3-D kernels, which sit at their 128-register limit, 16-32 bytes
// of spills).
template <int DIM, int COLL, int VEC, bool HALO = false>
__global__ void __launch_bounds__(step_max_threads<DIM, COLL>(), step_min_ctas<DIM, COLL>()) k_step(const StepParams<DIM> p, const MrtMats<DIM, mats_kind(COLL)> mm) {
  if constexpr (!HALO) {
    step_body<DIM, COLL, VEC, false>(p, mm);
    return;
  }
  if (p.halo.mode & 1) {
    // fused halo wait: both neighbours must have published the step this rank is about to take (their sends into my
    // ghost rows are complete, and they are done reading the buffer I am about to overwrite).  Every CTA waits itself;
    // counter[0] only changes when the LAST CTA of this launch reaches the epilogue, i.e. after all of them are past here.
    if (threadIdx.x == 0) {
      const unsigned step = *reinterpret_cast<volatile unsigned*>(p.halo.counter);
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      while ((int)(p.halo.my_flags[0] - step) < 0 || (int)(p.halo.my_flags[1] - step) < 0) {
        __nanosleep(64);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) { p.halo.counter[2] = 1u; break; }   // 10 s: report, do not hang
      }
      __threadfence_system();
    }
    __syncthreads();
  }
  step_body<DIM, COLL, VEC, HALO>(p, mm);
  if (p.halo.mode & 2) {
    // fused halo send: the peer stores of this CTA are out; the last CTA to get here publishes the step
    __threadfence_system();
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&p.halo.counter[1], 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
      __threadfence_system();
      p.halo.counter[1] = 0;
      const unsigned step = p.halo.counter[0] + 1;
      p.halo.counter[0] = step;
      p.halo.left_flags[1] = step;    // "your right neighbour has delivered step `step`"
      p.halo.right_flags[0] = step;   // "your left neighbour has delivered step `step`"
      __threadfence_system();
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
__global__ void k_lines_collide(const StepParams<DIM> p, const MrtMats<DIM, mats_kind(COLL)> mm, const LineSet ls) {
  using L = Lat<DIM>;
  int c[3];
  if (!line_cell<DIM>(p, ls, blockIdx.y, (long long)blockIdx.x * blockDim.x + threadIdx.x, c)) return;
  const long long ncell = (long long)p.n0 * p.n1 * p.n2;
  const long long cell = ((long long)c[0] * p.n1 + c[1]) * p.n2 + c[2];
  float f[L::Q], g[L::D];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) f[q] = p.fout[q * ncell + cell];
  int worg[3];
  window_origin<DIM>(p, worg);
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
__global__ void k_window_moments(const StepParams<DIM> p, float* __restrict__ u_win) {
  using L = Lat<DIM>;
  int org[3];
  window_origin<DIM>(p, org);
  long long wcells = 1;
#pragma unroll
  for (int d = 0; d < L::D; ++d) wcells *= p.wsz[d];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= wcells) return;
  int rel[3] = {0, 0, 0};
  long long r = t;
#pragma unroll
  for (int d = L::D - 1; d >= 0; --d) { rel[d] = (int)(r % p.wsz[d]); r /= p.wsz[d]; }
  int c[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < L::D; ++d) c[d + L::A0] = org[d] + rel[d];
  float f[L::Q], rho, u[L::D];
  pull_cell<DIM>(p, c[0], c[1], c[2], f, true);
  moments<DIM>(f, rho, u);
#pragma unroll
  for (int d = 0; d < L::D; ++d) u_win[t * WinVec<DIM>::NC + d] = u[d];
}

// The same on a list of window cells (ascending flat window indices).
template <int DIM>
__global__ void k_window_moments_cells(const StepParams<DIM> p, float* __restrict__ u_win, const int* __restrict__ cells,
                                       const long long n_cells) {
  using L = Lat<DIM>;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cells) return;
  int org[3];
  window_origin<DIM>(p, org);
  const int flat = __ldg(cells + t);
  int rel[3] = {0, 0, 0}, r = flat;
#pragma unroll
  for (int d = L::D - 1; d >= 0; --d) { rel[d] = r % p.wsz[d]; r /= p.wsz[d]; }
  int c[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < L::D; ++d) c[d + L::A0] = org[d] + rel[d];
  float f[L::Q], rho, u[L::D];
  pull_cell<DIM>(p, c[0], c[1], c[2], f, true);
  moments<DIM>(f, rho, u);
#pragma unroll
  for (int d = 0; d < L::D; ++d) u_win[(long long)flat * WinVec<DIM>::NC + d] = u[d];
}

// One wall cell: pull, face operation, (mask), collide, store.  Face operations handled this way are independent of
// each other (see vsb_edge_fused_supported), so no ordering between them is needed.
template <int DIM, int COLL, int LOC, bool HALO = false>
__device__ __forceinline__ void edge_cell(const StepParams<DIM>& p, const MrtMats<DIM, mats_kind(COLL)>& mm, int wall_layer,
                                          int kind, int wrap_kind, const WallVals& w, int mask_before, long long k) {
  using L = Lat<DIM>;
  using G = FaceGeom<DIM, LOC>;
  constexpr int Q = L::Q, D = L::D;
  const int n[3] = {p.n0, p.n1, p.n2};
  const long long nface = (long long)n[G::TA] * n[G::TB];
  if (k >= nface) return;
  int c[3];
  c[G::TA] = (int)(k / n[G::TB]);
  c[G::TB] = (int)(k % n[G::TB]);
  c[G::AX] = wall_layer;
  if (c[L::A0] < p.r_begin || c[L::A0] >= p.r_end) return;
  const long long ncell = (long long)p.n0 * p.n1 * p.n2;
  const long long cell = ((long long)c[0] * p.n1 + c[1]) * p.n2 + c[2];
  float fw[Q], fn[Q];
  pull_cell<DIM>(p, c[0], c[1], c[2], fw, mask_before != 0);
  if (bc_needs_neighbor(kind, wrap_kind)) {
    int cn[3] = {c[0], c[1], c[2]};
    cn[G::AX] += G::SIGN;
    pull_cell<DIM>(p, cn[0], cn[1], cn[2], fn, mask_before != 0);
  } else {
#pragma unroll
    for (int q = 0; q < Q; ++q) fn[q] = 0.f;
  }
  float uw[D], gw[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { uw[d] = wv(w.u[d], k); gw[d] = wv(w.g[d], k); }
  if (kind == VSB_BC_BOUNCE_BACK || kind == VSB_BC_SPECULAR) {
    float pre[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) pre[q] = p.fin[q * ncell + cell];
    apply_face_reflect<DIM, LOC>(fw, pre, kind == VSB_BC_SPECULAR ? 1 : 0, uw);
  } else {
    apply_face_bc<DIM, LOC>(fw, fn, kind, wrap_kind, wv(w.rho, k), uw, gw);
  }
  if (!mask_before && p.mask && p.mask[cell]) {
    float t[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) t[q] = fw[q];
#pragma unroll
    for (int q = 0; q < Q; ++q) fw[q] = t[L::opp(q)];
  }
  if (p.do_collide) {
    int worg[3];
    window_origin<DIM>(p, worg);
    bool rows_inside;
    int wbase;
    float g[D];
    window_rows<DIM>(p, worg, c[0], c[1], rows_inside, wbase);
    cell_force<DIM>(p, worg, rows_inside, wbase, c[2], g);
    collide_cell<DIM, COLL>(fw, g, p.forcing, p.rx, mm);
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) p.fout[q * ncell + cell] = fw[q];
  if constexpr (HALO) {   // a wall cell on an edge row of the slab: its crossing populations travel too
    if (p.halo.mode & 2) static_for<Q>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      constexpr int cx = L::c(q, L::A0);
      if constexpr (cx != 0) {
        float* base;
        long long shift;
        if (halo_target<DIM>(p, c[L::A0], cx, base, shift)) base[q * ncell + cell + shift] = fw[q];
      }
    });
  }
}

// Wall layer of one face as a kernel of its own (vsb_edge_fused).
template <int DIM, int COLL, int LOC>
__global__ void k_edge_fused(const StepParams<DIM> p, const MrtMats<DIM, mats_kind(COLL)> mm, int wall_layer, int kind,
                             int wrap_kind, WallVals w, int mask_before) {
  edge_cell<DIM, COLL, LOC>(p, mm, wall_layer, kind, wrap_kind, w, mask_before, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}

template <int DIM, int COLL, bool HALO>
__device__ __forceinline__ void edge_block(const StepParams<DIM>& p, const MrtMats<DIM, mats_kind(COLL)>& mm) {
  unsigned b = blockIdx.x - p.nb_bulk;
  int e = 0;
  if (p.n_wall > 1 && b >= p.wall_blocks0) { e = 1; b -= p.wall_blocks0; }
  const WallOpDev& op = p.wall[e];
  const long long k = (long long)b * blockDim.x + threadIdx.x;
  if constexpr (DIM == 2) {
    if (op.loc == 0) edge_cell<2, COLL, 0, HALO>(p, mm, op.layer, op.kind, op.wrap, op.w, op.mask_before, k);
    else edge_cell<2, COLL, 1, HALO>(p, mm, op.layer, op.kind, op.wrap, op.w, op.mask_before, k);
  } else {
    switch (op.loc) {
      case 0: edge_cell<3, COLL, 0, HALO>(p, mm, op.layer, op.kind, op.wrap, op.w, op.mask_before, k); break;
      case 1: edge_cell<3, COLL, 1, HALO>(p, mm, op.layer, op.kind, op.wrap, op.w, op.mask_before, k); break;
      case 2: edge_cell<3, COLL, 2, HALO>(p, mm, op.layer, op.kind, op.wrap, op.w, op.mask_before, k); break;
      default: edge_cell<3, COLL, 3, HALO>(p, mm, op.layer, op.kind, op.wrap, op.w, op.mask_before, k); break;
    }
  }
}

// ----------------------------------------------------------------------------- host side
template <int DIM>
int fill_params(const VsbStepArgs& a, StepParams<DIM>& p) {
  grid_axes(a.grid, p.n0, p.n1, p.n2);
  VSB_REQUIRE(p.n0 > 0 && p.n1 > 0 && p.n2 > 0, "vsb_step: bad grid");
  const int nrows = (DIM == 2) ? p.n1 : p.n0;
  p.r_begin = a.row_begin;
  p.r_end = a.row_end > 0 ? a.row_end : nrows;
  VSB_REQUIRE(0 <= p.r_begin && p.r_begin < p.r_end && p.r_end <= nrows, "vsb_step: bad row range [%d, %d) of %d",
              a.row_begin, a.row_end, nrows);
  p.s_begin = a.sub_end > 0 ? a.sub_begin : p.r_begin;
  p.s_end = a.sub_end > 0 ? a.sub_end : p.r_end;
  p.edge_rows = a.edge_rows_only ? 1 : 0;
  VSB_REQUIRE(!p.edge_rows || (a.band == 0 && p.r_end - p.r_begin >= 2), "vsb_step: edge_rows_only needs band = 0 and >= 2 rows");
  VSB_REQUIRE(p.r_begin <= p.s_begin && p.s_begin < p.s_end && p.s_end <= p.r_end,
              "vsb_step: sub-range [%d, %d) outside the rows [%d, %d)", a.sub_begin, a.sub_end, p.r_begin, p.r_end);
  VSB_REQUIRE(a.f_in && a.f_out && a.f_in != a.f_out, "vsb_step: f_in / f_out must be distinct non-null buffers");
  p.fin = a.f_in; p.fout = a.f_out;
  p.do_stream = a.do_stream; p.do_collide = a.do_collide; p.forcing = a.forcing;
  VSB_REQUIRE(a.forcing >= VSB_FORCE_NONE && a.forcing <= VSB_FORCE_GUO, "vsb_step: unknown forcing %d", a.forcing);
  p.rx = make_relax(a.omega);
  for (int d = 0; d < 3; ++d) { p.g0[d] = a.g_uniform[d]; p.worg[d] = a.win_origin[d]; p.wsz[d] = a.win_size[d]; p.wshift[d] = a.win_shift[d]; }
  p.gwin = a.g_win; p.body = a.body; p.parity = a.parity & 1; p.mask = nullptr;
  if (a.g_win) for (int d = 0; d < DIM; ++d) VSB_REQUIRE(a.win_size[d] > 0, "vsb_step: empty force window");
  VSB_REQUIRE(a.band >= 0 && a.band <= 2, "vsb_step: band must be 0, 1 or 2");
  VSB_REQUIRE(a.band == 0 || a.win_size[0] > 0, "vsb_step: band mode needs a force window");
  p.band = a.band;
  p.halo = HaloDev{};
  if (a.halo_mode) {
    VSB_REQUIRE(a.halo_mode > 0 && a.halo_mode <= 3, "vsb_step: halo_mode must be 0..3");
    VSB_REQUIRE(a.halo && a.edge_rows_only, "vsb_step: the fused halo hand-shake needs halo != NULL and edge_rows_only = 1");
    const VsbHaloArgs& h = *a.halo;
    VSB_REQUIRE(h.my_flags && h.left_flags && h.right_flags && h.counter && h.left_state && h.right_state,
                "vsb_step: incomplete VsbHaloArgs");
    VSB_REQUIRE(h.state == a.f_out, "vsb_step: halo->state must be the buffer this launch writes (f_out)");
    VSB_REQUIRE(h.grid.dim == a.grid.dim && h.grid.nx == a.grid.nx && h.grid.ny == a.grid.ny &&
                (a.grid.dim == 2 || h.grid.nz == a.grid.nz), "vsb_step: halo->grid differs from the step's grid");
    VSB_REQUIRE(p.r_begin == 1 && p.r_end == nrows - 1, "vsb_step: the fused halo hand-shake expects one ghost layer per side");
    p.halo.mode = a.halo_mode;
    p.halo.my_flags = h.my_flags; p.halo.left_flags = h.left_flags; p.halo.right_flags = h.right_flags;
    p.halo.counter = h.counter; p.halo.left_state = h.left_state; p.halo.right_state = h.right_state;
  }
  p.n_skip = 0;
  p.prefetch_blocks = 0;
  p.n_wall = 0;
  p.nb_bulk = 0xffffffffu;
  p.wall_blocks0 = 0;
  int n_mask = 0;
  for (int i = 0; i < a.n_post; ++i)
    if (a.post[i].kind == VSB_POST_MASK) { p.mask = a.post[i].mask; ++n_mask; }
  VSB_REQUIRE(n_mask <= 1, "vsb_step: at most one obstacle mask per step");
  VSB_REQUIRE(n_mask == 0 || p.mask, "vsb_step: mask op without a mask");
  return VSB_OK;
}
#ifndef VSB_STEP_PART
template int fill_params<2>(const VsbStepArgs&, StepParams<2>&);
template int fill_params<3>(const VsbStepArgs&, StepParams<3>&);
#endif

template <int DIM, int COLL>
static void fill_mats(const VsbStepArgs& a, MrtMats<DIM, mats_kind(COLL)>& mm) {
  if constexpr (COLL == VSB_COLL_MRT_MOMENT) {
    int c[19][3];
    for (int q = 0; q < 19; ++q)
      for (int d = 0; d < 3; ++d) c[q][d] = Lat<3>::c(q, d);
    make_moment_op3(a.mrt_op_host, a.forcing == VSB_FORCE_GUO ? a.mrt_fop_host : nullptr, c, mm.mo);   // checked by mrt_is_moment
  } else if constexpr (is_mrt(COLL)) {
    constexpr int Q = Lat<DIM>::Q;
    for (int i = 0; i < Q * Q; ++i) {
      mm.A.a[i] = a.mrt_op_host ? a.mrt_op_host[i] : 0.f;
      mm.B.a[i] = a.mrt_fop_host ? a.mrt_fop_host[i] : 0.f;
    }
    if constexpr (COLL == VSB_COLL_MRT_SPLIT) {
      make_split_op<DIM>(mm.A.a, mm.As);
      make_split_op<DIM>(mm.B.a, mm.Bs);
    }
  }
}

// True when the MRT operators of this call commute with the reflection c -> -c (see SplitOp).
template <int DIM>
static bool mrt_is_split(const VsbStepArgs& a) {
  constexpr int Q = Lat<DIM>::Q;
  static const bool allowed = [] { const char* e = getenv("VSB_MRT_FORM"); return !e || e[0] != 'd'; }();
  if (!allowed || !a.mrt_op_host) return false;
  SplitOp<DIM> tmp;
  if (!make_split_op<DIM>(a.mrt_op_host, tmp)) return false;
  if (a.mrt_fop_host) return make_split_op<DIM>(a.mrt_fop_host, tmp);
  float zero[Q * Q] = {};
  return make_split_op<DIM>(zero, tmp);
}

// True when the D3Q19 operators of this call are diagonal in the reference's moment basis (see vsb_mrt_moment.cuh);
// VSB_MRT_FORM=split|dense in the environment keeps the matrix kernels (tuning / A-B comparisons).
template <int DIM>
static bool mrt_is_moment(const VsbStepArgs& a) {
  if constexpr (DIM != 3) return false;
  static const bool allowed = [] { const char* e = getenv("VSB_MRT_FORM"); return !e || e[0] == 'm'; }();
  if (!allowed || !a.mrt_op_host) return false;
  if (a.forcing == VSB_FORCE_GUO && !a.mrt_fop_host) return false;
  int c[19][3];
  for (int q = 0; q < 19; ++q)
    for (int d = 0; d < 3; ++d) c[q][d] = Lat<3>::c(q, d);
  MomentOp3 tmp;
  return make_moment_op3(a.mrt_op_host, a.forcing == VSB_FORCE_GUO ? a.mrt_fop_host : nullptr, c, tmp);
}

static int check_mrt(const VsbStepArgs& a) {
  if (a.collision != VSB_COLL_MRT || !a.do_collide) return VSB_OK;
  VSB_REQUIRE(a.mrt_op_host, "vsb_step: MRT collision needs mrt_op_host");
  VSB_REQUIRE(a.forcing != VSB_FORCE_GUO || a.mrt_fop_host, "vsb_step: MRT + Guo forcing needs mrt_fop_host");
  return VSB_OK;
}

// Wall layer (array axis, layer index) of a face operation for this launch's row range.
template <int DIM>
static void wall_of(const StepParams<DIM>& p, int loc, int& ax, int& wall, int& extent) {
  const int n[3] = {p.n0, p.n1, p.n2};
  ax = loc / 2 + Lat<DIM>::A0;
  const int lo = (ax == Lat<DIM>::A0) ? p.r_begin : 0, hi = (ax == Lat<DIM>::A0) ? p.r_end : n[ax];
  wall = (loc % 2 == 0) ? lo : hi - 1;
  extent = hi - lo;
}

// Face operations are independent when they sit on faces normal to ONE non-contiguous array axis (so wall cells of
// different operations never coincide and no operation reads a layer another one writes).
template <int DIM>
static bool edges_independent(const VsbStepArgs& a, const StepParams<DIM>& p, int& mask_before) {
  int axis = -1, n_bc = 0, first_mask = -1, last_bc = -1, first_bc = -1;
  bool seen[6] = {false, false, false, false, false, false};
  for (int i = 0; i < a.n_post; ++i) {
    const VsbPostOp& op = a.post[i];
    if (op.kind == VSB_POST_MASK) { first_mask = i; continue; }
    if (op.loc < 0 || op.loc >= 2 * DIM || seen[op.loc]) return false;
    seen[op.loc] = true;
    int ax, wall, extent;
    wall_of<DIM>(p, op.loc, ax, wall, extent);
    if (ax == 2 || extent < 4) return false;
    if (axis >= 0 && ax != axis) return false;
    axis = ax;
    if (first_bc < 0) first_bc = i;
    last_bc = i;
    ++n_bc;
  }
  if (n_bc == 0) return false;
  mask_before = 0;
  if (first_mask >= 0) {
    if (first_mask < first_bc) mask_before = 1;
    else if (first_mask > last_bc) mask_before = 0;
    else return false;
  }
  return true;
}

template <int DIM, int COLL, int LOC>
static int launch_edge(const StepParams<DIM>& p, const MrtMats<DIM, mats_kind(COLL)>& mm, const VsbPostOp& op,
                       int mask_before, cudaStream_t s) {
  using G = FaceGeom<DIM, LOC>;
  const int n[3] = {p.n0, p.n1, p.n2};
  int ax, wall, extent;
  wall_of<DIM>(p, op.loc, ax, wall, extent);
  WallVals w;
  w.rho = op.rho;
  for (int d = 0; d < 3; ++d) { w.u[d] = op.u[d]; w.g[d] = op.g[d]; }
  const long long nface = (long long)n[G::TA] * n[G::TB];
  k_edge_fused<DIM, COLL, LOC><<<blocks_for(nface, 128), 128, 0, s>>>(p, mm, wall, op.kind, op.wrap, w, mask_before);
  VSB_LAUNCH_CHECK("vsb_edge_fused");
  return VSB_OK;
}

template <int DIM, int COLL>
int edge_impl(const VsbStepArgs& a, cudaStream_t s, bool query_only, int* supported) {
  StepParams<DIM> p;
  if (int rc = fill_params<DIM>(a, p)) return rc;
  int mask_before = 0;
  const bool ok = a.do_stream && edges_independent<DIM>(a, p, mask_before);
  if (supported) *supported = ok ? 1 : 0;
  if (query_only) return VSB_OK;
  VSB_REQUIRE(ok, "vsb_edge_fused: the face operations are not independent; use the ordered fix-up (edges = 0)");
  if (int rc = check_mrt(a)) return rc;
  MrtMats<DIM, mats_kind(COLL)> mm;
  fill_mats<DIM, COLL>(a, mm);
  for (int i = 0; i < a.n_post; ++i) {
    const VsbPostOp& op = a.post[i];
    if (op.kind == VSB_POST_MASK) continue;
    int rc = VSB_OK;
    if constexpr (DIM == 2) {
      rc = op.loc == 0 ? launch_edge<2, COLL, 0>(p, mm, op, mask_before, s) : launch_edge<2, COLL, 1>(p, mm, op, mask_before, s);
    } else {
      switch (op.loc) {
        case 0: rc = launch_edge<3, COLL, 0>(p, mm, op, mask_before, s); break;
        case 1: rc = launch_edge<3, COLL, 1>(p, mm, op, mask_before, s); break;
        case 2: rc = launch_edge<3, COLL, 2>(p, mm, op, mask_before, s); break;
        default: rc = launch_edge<3, COLL, 3>(p, mm, op, mask_before, s); break;
      }
    }
    if (rc) return rc;
  }
  return VSB_OK;
}

template <int DIM, int COLL>
int step_impl(const VsbStepArgs& a, cudaStream_t s) {
  StepParams<DIM> p;
  if (int rc = fill_params<DIM>(a, p)) return rc;
  if (int rc = check_mrt(a)) return rc;
  MrtMats<DIM, mats_kind(COLL)> mm;
  fill_mats<DIM, COLL>(a, mm);
  const bool have_ops = a.n_post > 0 && a.do_stream;
  if ((a.edges == 1 || a.edges == 2) && have_ops) {
    int mask_before = 0;
    VSB_REQUIRE(edges_independent<DIM>(a, p, mask_before), "vsb_step: edges = 1 / 2 need independent face operations");
    for (int i = 0; i < a.n_post; ++i) {
      const VsbPostOp& op = a.post[i];
      if (op.kind == VSB_POST_MASK) continue;
      int ax, wall, extent;
      wall_of<DIM>(p, op.loc, ax, wall, extent);
      VSB_REQUIRE(p.n_skip < 2, "vsb_step: too many wall layers");
      p.skip_axis[p.n_skip] = ax; p.skip_layer[p.n_skip] = wall; ++p.n_skip;
      // edges = 2: the wall layer is processed by blocks appended to the launch that covers its rows
      const bool covered = a.band != 2 && (ax != Lat<DIM>::A0 || p.edge_rows || (wall >= p.s_begin && wall < p.s_end));
      if (a.edges == 2 && covered) {
        WallOpDev& wo = p.wall[p.n_wall++];
        wo.kind = op.kind; wo.wrap = op.wrap; wo.loc = op.loc; wo.layer = wall; wo.mask_before = mask_before;
        wo.w.rho = op.rho;
        for (int d = 0; d < 3; ++d) { wo.w.u[d] = op.u[d]; wo.w.g[d] = op.g[d]; }
      }
    }
  }
  int vec = a.vec;
  if (vec == 0) vec = step_default_vec<DIM, COLL>();
  while (vec > 1 && (p.n2 % vec != 0 || ((uintptr_t)a.f_in % (4 * vec)) || ((uintptr_t)a.f_out % (4 * vec)))) vec >>= 1;
  VSB_REQUIRE(vec == 1 || vec == 2 || vec == 4, "vsb_step: vec must be 0, 1, 2 or 4");
  if (p.halo.mode) vec = 1;   // the hand-shaking edge-row launch exists in the scalar variant only (same arithmetic per cell)
  const int nrow = p.edge_rows ? 2 : ((p.band == 2) ? std::min(p.wsz[0], p.s_end - p.s_begin) : p.s_end - p.s_begin);
  const bool box = DIM == 3 && p.band == 2;            // 3-D band 2: only the window's y-range of every x plane
  const long long rows = (DIM == 2) ? (long long)nrow : (long long)nrow * (box ? p.wsz[1] : p.n1);
  const bool box2 = DIM == 2 && p.band == 2;           // 2-D band 2: only the vector groups over the window's y-range
  p.nvb = std::min(p.n2 / vec, p.wsz[1] / vec + 2);    // any alignment of a window of wsz[1] cells touches <= this many
  const long long total = rows * (box2 ? p.nvb : p.n2 / vec);
  VSB_REQUIRE(total < (1ll << 31) && (long long)p.n0 * p.n1 * p.n2 < (1ll << 31),
              "vsb_step: more than 2^31 cells in one launch; split the rows (sub_begin / sub_end)");
  p.div_nv = make_fast_div((unsigned)(p.n2 / vec));
  p.div_n1 = make_fast_div((unsigned)p.n1);
  p.div_w1 = make_fast_div((unsigned)std::max(p.wsz[1], 1));
  p.div_nvb = make_fast_div((unsigned)std::max(p.nvb, 1));
  // Block size: for grids of only a few waves (e.g. 1024^2 = 1.73 waves of 256-thread blocks) the partly filled last
  // wave costs up to a whole wave; choose the multiple of 32 in [128, 256] that fills the last wave best.
  auto launch = [&](auto kernel) {
    static int regs = 0, n_sm = 0;
    if (regs == 0) {
      cudaFuncAttributes fa;
      int dev = 0;
      cudaGetDevice(&dev);
      if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) regs = fa.numRegs;
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      if (regs <= 0) regs = 128;
      if (n_sm <= 0) n_sm = 148;
    }
    constexpr int max_bs = step_max_threads<DIM, COLL>();
    int best_bs = max_bs;
    double best = -1.0;
    for (int bs = max_bs; bs >= max_bs / 2; bs -= 32) {
      const long long nb = (total + bs - 1) / bs;
      const int warps = bs / 32;
      const int per_sm = std::max(1, std::min(std::min(65536 / (((regs + 7) / 8 * 8) * 32) / warps, 2048 / bs), 32));
      const double waves = (double)nb / ((double)per_sm * n_sm);
      const double fill = waves / std::ceil(waves);                       // how full the average wave is
      const double occ = std::min(1.0, per_sm * bs / 768.0);              // mild preference for >= 768 threads / SM
      const double score = fill * (0.9 + 0.1 * occ) + (bs == max_bs ? 1e-3 : 0.0);
      if (score > best) { best = score; best_bs = bs; }
    }
    p.nb_bulk = blocks_for(total, best_bs);
    {
      // L2 prefetch distance: about 5.5 MB of populations ahead of the block (measured optimum on B200 for both
      // lattices, scripts/prefetch_sweep.py; beyond ~2x that the lines are evicted again before use)
      static const double pf_bytes = [] { const char* e = getenv("VSB_PREFETCH_KB"); return (e ? atof(e) : 5632.0) * 1024.0; }();
      const double block_bytes = (double)best_bs * vec * Lat<DIM>::Q * 4.0;
      p.prefetch_blocks = (pf_bytes > 0 && !box && !box2) ? std::max(1, (int)(pf_bytes / block_bytes + 0.5)) : 0;   // (box rows are not contiguous)
    }
    unsigned extra = 0;
    for (int e = 0; e < p.n_wall; ++e) {
      const int n[3] = {p.n0, p.n1, p.n2};
      const int ax = p.wall[e].loc / 2 + Lat<DIM>::A0;
      const unsigned nbw = blocks_for((long long)n[0] * n[1] * n[2] / n[ax], best_bs);
      if (e == 0) p.wall_blocks0 = nbw;
      extra += nbw;
    }
    if (a.early_launch) {
      // programmatic dependent launch: begin as soon as every CTA of the preceding kernel on this stream has started
      // (they call griddepcontrol.launch_dependents first thing) -- that kernel keeps the SMs it already holds
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(p.nb_bulk + extra);
      cfg.blockDim = dim3(best_bs);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, kernel, p, mm);
    } else {
      kernel<<<p.nb_bulk + extra, best_bs, 0, s>>>(p, mm);
    }
  };
  if (p.halo.mode) launch(k_step<DIM, COLL, 1, true>);   // two edge rows: the scalar variant is the only one instantiated
  else if (vec == 4) launch(k_step<DIM, COLL, 4>);
  else if (vec == 2) launch(k_step<DIM, COLL, 2>);
  else launch(k_step<DIM, COLL, 1>);
  VSB_LAUNCH_CHECK("vsb_step (fused kernel)");

  if (!have_ops || a.edges == 1 || a.edges == 2) return VSB_OK;
  VSB_REQUIRE(a.band == 0 && !p.edge_rows && p.s_begin == p.r_begin && p.s_end == p.r_end,
              "vsb_step: the ordered wall fix-up (edges = 0) cannot be combined with band modes or row sub-ranges");
  // layers touched by face operations: wall layer and adjacent fluid layer of each face
  LineSet ls;
  ls.n = 0;
  const int n[3] = {p.n0, p.n1, p.n2};
  for (int i = 0; i < a.n_post; ++i) {
    const VsbPostOp& op = a.post[i];
    if (op.kind == VSB_POST_MASK) continue;
    VSB_REQUIRE(op.loc >= 0 && op.loc < 2 * DIM, "vsb_step: loc %d is not a face of a %d-D grid", op.loc, DIM);
    int ax, wall, extent;
    wall_of<DIM>(p, op.loc, ax, wall, extent);
    VSB_REQUIRE(extent >= 2, "vsb_step: fewer than 2 layers along the normal of face %d", op.loc);
    const int layers[2] = {wall, wall + ((op.loc % 2 == 0) ? 1 : -1)};
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

// ----------------------------------------------------------------------------- build slicing
// The (lattice, collision) instantiations of step_impl / edge_impl -- and with them every k_step / k_edge_fused
// kernel -- live in the translation units vsb_step_part<k>.cu (each defines VSB_STEP_PART, includes this file and
// instantiates its share explicitly), so that they compile in parallel.  This file alone keeps the dispatcher and the
// C entry points and only declares the instantiations.
#ifndef VSB_STEP_PART
#define VSB_STEP_EXTERN(D, C)                                                           \
  extern template int step_impl<D, C>(const VsbStepArgs&, cudaStream_t);              \
  extern template int edge_impl<D, C>(const VsbStepArgs&, cudaStream_t, bool, int*);
VSB_STEP_EXTERN(2, VSB_COLL_BGK) VSB_STEP_EXTERN(2, VSB_COLL_REG) VSB_STEP_EXTERN(2, VSB_COLL_KBC)
VSB_STEP_EXTERN(2, VSB_COLL_MRT) VSB_STEP_EXTERN(2, VSB_COLL_MRT_SPLIT)
VSB_STEP_EXTERN(3, VSB_COLL_BGK) VSB_STEP_EXTERN(3, VSB_COLL_REG) VSB_STEP_EXTERN(3, VSB_COLL_KBC)
VSB_STEP_EXTERN(3, VSB_COLL_MRT) VSB_STEP_EXTERN(3, VSB_COLL_MRT_SPLIT) VSB_STEP_EXTERN(3, VSB_COLL_MRT_MOMENT)
#undef VSB_STEP_EXTERN

template <int DIM>
static int step_dispatch(const VsbStepArgs& a, cudaStream_t s, int what, int* supported) {
  // what: 0 step, 1 fused wall kernel, 2 query support of the fused wall kernel
  switch (a.collision) {
    case VSB_COLL_BGK: return what ? edge_impl<DIM, VSB_COLL_BGK>(a, s, what == 2, supported) : step_impl<DIM, VSB_COLL_BGK>(a, s);
    case VSB_COLL_MRT:
      if constexpr (DIM == 3) {
        if (mrt_is_moment<DIM>(a))
          return what ? edge_impl<3, VSB_COLL_MRT_MOMENT>(a, s, what == 2, supported) : step_impl<3, VSB_COLL_MRT_MOMENT>(a, s);
      }
      if (mrt_is_split<DIM>(a))
        return what ? edge_impl<DIM, VSB_COLL_MRT_SPLIT>(a, s, what == 2, supported) : step_impl<DIM, VSB_COLL_MRT_SPLIT>(a, s);
      return what ? edge_impl<DIM, VSB_COLL_MRT>(a, s, what == 2, supported) : step_impl<DIM, VSB_COLL_MRT>(a, s);
    case VSB_COLL_KBC: return what ? edge_impl<DIM, VSB_COLL_KBC>(a, s, what == 2, supported) : step_impl<DIM, VSB_COLL_KBC>(a, s);
    case VSB_COLL_REG: return what ? edge_impl<DIM, VSB_COLL_REG>(a, s, what == 2, supported) : step_impl<DIM, VSB_COLL_REG>(a, s);
    default: VSB_REQUIRE(false, "vsb_step: unknown collision %d", a.collision);
  }
}

template <int DIM>
static int window_impl(const VsbStepArgs& a, float* u_win, cudaStream_t s, const int32_t* cells = nullptr, long long n_cells = 0) {
  StepParams<DIM> p;
  VsbStepArgs b = a;
  if (!b.f_out) b.f_out = u_win;  // unused by this kernel; only has to differ from f_in
  b.band = 0;
  if (int rc = fill_params<DIM>(b, p)) return rc;
  long long wcells = 1;
  for (int d = 0; d < DIM; ++d) {
    VSB_REQUIRE(a.win_size[d] > 0, "vsb_ib_window_moments: empty window");
    wcells *= a.win_size[d];
  }
  if (cells) {
    VSB_REQUIRE(n_cells >= 0 && n_cells <= wcells, "vsb_ib_window_moments_cells: %lld cells listed, the window has %lld", n_cells, wcells);
    if (n_cells > 0) k_window_moments_cells<DIM><<<blocks_for(n_cells, 128), 128, 0, s>>>(p, u_win, cells, n_cells);
  } else {
    k_window_moments<DIM><<<blocks_for(wcells, 128), 128, 0, s>>>(p, u_win);
  }
  VSB_LAUNCH_CHECK("vsb_ib_window_moments");
  return VSB_OK;
}
#endif  // !VSB_STEP_PART

}  // namespace vsb

#ifndef VSB_STEP_PART
using namespace vsb;

static int check_step_args(const VsbStepArgs* args, const char* who) {
  VSB_REQUIRE(args != nullptr, "%s: null args", who);
  VSB_REQUIRE(args->grid.dim == 2 || args->grid.dim == 3, "dim must be 2 or 3, got %d", args->grid.dim);
  VSB_REQUIRE(args->n_post >= 0 && (args->n_post == 0 || args->post), "%s: bad post list", who);
  return VSB_OK;
}

extern "C" {

int vsb_step(const VsbStepArgs* args, vsb_stream_t stream) {
  if (int rc = check_step_args(args, "vsb_step")) return rc;
  VSB_REQUIRE(args->do_stream || args->do_collide, "vsb_step: nothing to do");
  return args->grid.dim == 2 ? step_dispatch<2>(*args, (cudaStream_t)stream, 0, nullptr)
                             : step_dispatch<3>(*args, (cudaStream_t)stream, 0, nullptr);
}

int vsb_edge_fused(const VsbStepArgs* args, vsb_stream_t stream) {
  if (int rc = check_step_args(args, "vsb_edge_fused")) return rc;
  return args->grid.dim == 2 ? step_dispatch<2>(*args, (cudaStream_t)stream, 1, nullptr)
                             : step_dispatch<3>(*args, (cudaStream_t)stream, 1, nullptr);
}

int vsb_edge_fused_supported(const VsbStepArgs* args) {
  if (check_step_args(args, "vsb_edge_fused_supported")) return 0;
  int ok = 0;
  const int rc = args->grid.dim == 2 ? step_dispatch<2>(*args, nullptr, 2, &ok) : step_dispatch<3>(*args, nullptr, 2, &ok);
  return rc == VSB_OK ? ok : 0;
}

int vsb_ib_window_moments(const VsbStepArgs* args, float* u_win, vsb_stream_t stream) {
  VSB_REQUIRE(args && u_win, "vsb_ib_window_moments: null argument");
  VSB_REQUIRE(args->grid.dim == 2 || args->grid.dim == 3, "dim must be 2 or 3, got %d", args->grid.dim);
  return args->grid.dim == 2 ? window_impl<2>(*args, u_win, (cudaStream_t)stream) : window_impl<3>(*args, u_win, (cudaStream_t)stream);
}

int vsb_ib_window_moments_cells(const VsbStepArgs* args, float* u_win, const int32_t* cells, int64_t n_cells, vsb_stream_t stream) {
  VSB_REQUIRE(args && u_win && cells, "vsb_ib_window_moments_cells: null argument");
  VSB_REQUIRE(args->grid.dim == 2 || args->grid.dim == 3, "dim must be 2 or 3, got %d", args->grid.dim);
  return args->grid.dim == 2 ? window_impl<2>(*args, u_win, (cudaStream_t)stream, cells, n_cells)
                             : window_impl<3>(*args, u_win, (cudaStream_t)stream, cells, n_cells);
}

}  // extern "C"
#endif  // !VSB_STEP_PART
