// Immersed-boundary kernels: delta kernels, stencil, interpolate,
// spread, and marker-parallel multi-direct forcing with the stencil evaluated on the fly.
// A group of lanes owns one marker; its stencil points are spread over the lanes and reduced
// with warp shuffles; spreading uses fp32 atomics (REDG) into window-sized buffers.
#include <type_traits>

#include <chrono>
#include <climits>
#include <cstdlib>
#include <string>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "vsb_mdf.cuh"

namespace vsb {

constexpr int kBlock = 128;
constexpr int kMdfCtasPerSm = 0;   // cap of the per-iteration MDF grid in CTAs per SM (0: none); VSB_MDF_GRID_CAP overrides (CTAs)

__global__ void k_delta(int kind, long long n, const float* __restrict__ r, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = delta(kind, r[i]);
}

// get_ib_stencil: offsets arange(-r+1, r+1) around floor(x), "ij" tensor product, no wrap / clamp
// (ib/stencil.py:27-51, ib3d/stencil.py:36-59)
template <int DIM>
__global__ void k_stencil(int kind, int radius, long long n_markers, const float* __restrict__ coords, int ny, int nz,
                          float* __restrict__ weights, int* __restrict__ indices) {
  const int side = 2 * radius;
  const int ns = (DIM == 2) ? side * side : side * side * side;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_markers * ns) return;
  const long long m = t / ns;
  int s = (int)(t % ns);
  int o[3] = {0, 0, 0};
  for (int d = DIM - 1; d >= 0; --d) { o[d] = s % side - radius + 1; s /= side; }
  float w = 1.f;
  int node[3] = {0, 0, 0};
  for (int d = 0; d < DIM; ++d) {
    const float x = coords[m * DIM + d];
    node[d] = (int)floorf(x) + o[d];
    w *= delta(kind, (float)node[d] - x);
  }
  weights[t] = w;
  indices[t] = (DIM == 2) ? node[0] * ny + node[1] : node[0] * (ny * nz) + node[1] * nz + node[2];
}

// interpolate: out[m,c] = sum_s w[m,s] grid[c, idx[m,s]]           (ib/stencil.py:76-78)   one warp per marker
__global__ void k_interpolate(int ncomp, long long ncell, const float* __restrict__ grid, long long n_markers, int ns,
                              const float* __restrict__ w, const int* __restrict__ idx, float* __restrict__ out) {
  const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= n_markers) return;
  for (int c = 0; c < ncomp; ++c) {
    float acc = 0.f;
    for (int s = lane; s < ns; s += 32) {
      // out-of-range stencil indices (a marker within 2 cells of the grid edge): jnp indexing wraps a negative index
      // once and clamps a gather index into range; nothing is read outside the grid
      long long i = idx[m * ns + s];
      i += (i < 0) ? ncell : 0;
      i = i < 0 ? 0 : (i >= ncell ? ncell - 1 : i);
      acc += w[m * ns + s] * grid[c * ncell + i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[m * ncomp + c] = acc;
  }
}

// spread: grid[c, idx[m,s]] += val[m,c] w[m,s]                    (ib/stencil.py:104-110)
__global__ void k_spread(int ncomp, long long ncell, float* __restrict__ grid, long long n_markers, int ns,
                         const float* __restrict__ vals, const float* __restrict__ w, const int* __restrict__ idx) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_markers * ns) return;
  const long long m = t / ns;
  const float wt = w[t];
  long long i = idx[t];
  i += (i < 0) ? ncell : 0;                 // jnp: a negative index wraps once ...
  if (i < 0 || i >= ncell) return;          // ... and an out-of-range scatter update is dropped
  for (int c = 0; c < ncomp; ++c) atomicAdd(&grid[c * ncell + i], vals[m * ncomp + c] * wt);
}

// ----------------------------------------------------------------------------- MDF stage chain
// Stage k of multi_direct_forcing (ib/mdf.py:31-64), one launch per iteration, markers spread over many CTAs:
//   k = 0     : u at the stencil points = moments of the pulled (streamed, masked) populations, u_m = interp(u)
//   k > 0     : u_m += interp(0.5 * spread(dF_{k-1}))        (buffer scratch[k-1], filled by stage k-1)
//   dF_k = (U - u_m) 2 ds ; F += dF_k
//   k < n-1   : spread dF_k -> scratch[k]
//   k = n-1   : spread F -> g_win ; total force -> body ; the last CTA to finish performs the body update
//               (the last iteration's u_m update is never used by the reference's outputs, so its spread +
//               interpolate are skipped)
// Buffers are double-buffered by step parity: while this step accumulates into its own set, every stage clears the
// matching buffer of the other set, so no memset is needed and nothing is cleared while it may still be read.
// Barrier across the (small, co-resident) grid of the fused MDF launch.  Monotonic 64-bit ticket counter: never reset.
// The launch is cooperative (cudaLaunchCooperativeKernel), so the CTAs are guaranteed to be co-resident.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(bar, 1ull);
    const unsigned long long target = (t / nblocks + 1ull) * nblocks;
    while (*reinterpret_cast<volatile unsigned long long*>(bar) < target) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

// 3-D: five CTAs of 128 threads per SM (<= 102 registers) = 740 resident CTAs, so that a body of up to ~2900 markers
// (the 2562-marker sphere of the 256^3 case needs 641 CTAs) runs as ONE wave; at 110 registers the last 49 CTAs made
// up a second wave that cost a whole CTA latency per stage.
template <int DIM, bool SHARD>
__global__ void __launch_bounds__(kBlock, DIM == 3 ? 5 : 8) k_mdf_stage(const StepParams<DIM> sp, const MdfParams p, const BodyUpdate bu,
                                                                      const ShardArg<SHARD> sh) {
  using L = Lat<DIM>;
  constexpr int NS = (DIM == 2) ? 16 : 64;   // 4^D stencil points
  constexpr int G = (DIM == 2) ? 16 : 32;    // lanes per marker
  constexpr int PPL = NS / G;                // stencil points per lane
  constexpr int NC = WinVec<DIM>::NC;
  using VecF = typename std::conditional<DIM == 2, float2, float4>::type;
  // a fused step enqueued behind this launch with early_launch = 1 may start as soon as every CTA here is resident
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ float s_force[3];
  if (threadIdx.x < 3) s_force[threadIdx.x] = 0.f;
  __syncthreads();

  const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const int gl = (int)(gthread % G);
  const long long wcells = (long long)p.wsize[0] * p.wsize[1] * (DIM == 3 ? p.wsize[2] : 1);

  int org[3] = {p.origin0[0], p.origin0[1], p.origin0[2]};
  if (p.body) { org[0] = p.body->origin2[p.parity][0]; org[1] = p.body->origin2[p.parity][1]; org[2] = p.body->origin2[p.parity][2]; }

  // A launch of ONE stage may use fewer lane groups than markers (grid capped by the host so that the chain leaves
  // most of every SM to the bulk kernel running beside it): the groups then walk the marker list in batches.  The
  // trip count is the same for every thread, so the warp shuffles below stay converged.
  const long long mstride = nthreads / G;
  const long long n_batch = (p.m_end - p.m_begin + mstride - 1) / mstride;
  for (long long batch = 0; batch < n_batch; ++batch) {
  const long long m = p.m_begin + batch * mstride + gthread / G;
  const bool active = m < p.m_end;

  // stencil of this lane's marker: the same for every iteration
  float w[PPL];
  long long idx[PPL];
  int node[PPL][DIM];
  bool ok[PPL];
  float ds2 = 0.f, tgt[DIM], u_m[DIM], F[DIM], pos[DIM], arm[2];
#pragma unroll
  for (int c = 0; c < DIM; ++c) { tgt[c] = 0.f; u_m[c] = 0.f; F[c] = 0.f; pos[c] = 0.f; }
  if (active) {
    float x[DIM];
    int base[DIM];
    marker_kinematics<DIM>(p, m, pos, tgt, arm);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      x[d] = pos[d] - (float)org[d];   // window-local coordinate, as in the reference's marker_x - ib_x0
      base[d] = (int)floorf(x[d]);
    }
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
      int s = gl * PPL + j;
      float wt = 1.f;
      bool inside = true;
#pragma unroll
      for (int d = DIM - 1; d >= 0; --d) {
        node[j][d] = base[d] + (s & 3) - 1;
        s >>= 2;
        wt *= delta(p.delta_kind, (float)node[j][d] - x[d]);
        inside = inside && node[j][d] >= 0 && node[j][d] < p.wsize[d];
      }
      w[j] = wt;
      ok[j] = inside;
      idx[j] = (DIM == 2) ? (long long)node[j][0] * p.wsize[1] + node[j][1]
                          : ((long long)node[j][0] * p.wsize[1] + node[j][1]) * p.wsize[2] + node[j][2];
    }
    ds2 = (p.ds_ptr ? p.ds_ptr[m] : p.ds_value) * 2.0f;
#pragma unroll
    for (int c = 0; c < DIM; ++c)
      if (p.stage > 0) { u_m[c] = p.marker_u[m * DIM + c]; F[c] = p.marker_force[m * DIM + c]; }   // from the previous launch
  } else {
#pragma unroll
    for (int j = 0; j < PPL; ++j) { w[j] = 0.f; ok[j] = false; idx[j] = 0; }
  }

  // iterations [stage, stage_end): one per launch, or all of them in one launch separated by grid barriers
  for (int stage = p.stage; stage < p.stage_end; ++stage) {
    const bool last = stage == p.n_iter - 1;
    // clear the other parity's buffers for the next step
    if (stage == 0 && batch == 0) clear_field<NC>(p, p.g_win_next, org[0], 0, gthread, nthreads);
    if (stage < p.n_iter - 1 && batch == 0)
      clear_field<NC>(p, p.scratch_next + (long long)stage * NC * wcells, org[0], 1, gthread, nthreads);

    float um[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) um[c] = 0.f;
    if (stage == 0 && p.u_win == nullptr) {
#pragma unroll
      for (int j = 0; j < PPL; ++j)
        if (ok[j]) {
          int cell[3] = {0, 0, 0};
#pragma unroll
          for (int d = 0; d < DIM; ++d) cell[d + L::A0] = org[d] + node[j][d];
          float f[L::Q], rho, u[DIM];
          pull_cell<DIM>(sp, cell[0], cell[1], cell[2], f, true);
          moments<DIM>(f, rho, u);
#pragma unroll
          for (int c = 0; c < DIM; ++c) um[c] += w[j] * u[c];
        }
    } else {
      const VecF* src = reinterpret_cast<const VecF*>(stage == 0 ? p.u_win : p.scratch + (long long)(stage - 1) * NC * wcells);
#pragma unroll
      for (int j = 0; j < PPL; ++j)
        if (ok[j]) {
          const VecF v = __ldcg(src + idx[j]);   // written by other SMs' reductions: read at L2
          um[0] += w[j] * v.x;
          um[1] += w[j] * v.y;
          if constexpr (DIM == 3) um[2] += w[j] * v.z;
        }
    }
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) um[c] += __shfl_xor_sync(0xffffffffu, um[c], o);
    }

    if (active) {
      float spread_val[DIM];
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        u_m[c] = (stage == 0) ? um[c] : u_m[c] + 0.5f * um[c];
        const float dF = (tgt[c] - u_m[c]) * ds2;
        F[c] = (stage == 0 ? 0.f : F[c]) + dF;
        spread_val[c] = last ? F[c] : dF;
      }
      if (gl == 0) {
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
          p.marker_u[m * DIM + c] = u_m[c];
          p.marker_force[m * DIM + c] = F[c];
        }
      }
      VecF* dst = reinterpret_cast<VecF*>(last ? p.g_win : p.scratch + (long long)stage * NC * wcells);
#pragma unroll
      for (int j = 0; j < PPL; ++j)
        if (ok[j]) {   // one vector reduction (red.global.add.v2/v4.f32) per stencil point
          VecF val;
          if constexpr (DIM == 2) val = make_float2(spread_val[0] * w[j], spread_val[1] * w[j]);
          else val = make_float4(spread_val[0] * w[j], spread_val[1] * w[j], spread_val[2] * w[j], 0.f);
          if constexpr (SHARD) shard_add<VecF>(sh, org[0], node[j][0], node[j][1], DIM == 3 ? node[j][DIM - 1] : 0, idx[j], val);
          else atomicAdd(dst + idx[j], val);
        }
      if (last && p.body && gl == 0) {
#pragma unroll
        for (int c = 0; c < DIM; ++c) atomicAdd(&s_force[c], spread_val[c]);
        if constexpr (DIM == 2) {
          if (p.rotation) atomicAdd(&s_force[2], marker_torque(p, pos, spread_val));
        }
      }
    }
    if (stage + 1 < p.stage_end) grid_barrier(p.barrier, gridDim.x);
  }
  }   // marker batches

  if (p.stage_end == p.n_iter && p.body) {
    __syncthreads();
    if (threadIdx.x < (p.rotation ? 3 : DIM)) atomicAdd(&p.body->force_sum[threadIdx.x], s_force[threadIdx.x]);
    if (p.update_body || p.host_mail) {   // the last CTA to arrive sees every contribution
      __shared__ int s_last;
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&p.body->ticket, 1) == (int)gridDim.x - 1);
      }
      __syncthreads();
      if (s_last && threadIdx.x == 0) {
        __threadfence();
        p.body->ticket = 0;
        finish_body(p, bu);
      }
    }
  }
}


// ----------------------------------------------------------------------------- tiled MDF stage (dense bodies)
// Same arithmetic as k_mdf_stage for stages that read a window field (u_win at stage 0, scratch[k-1] after), but a
// CTA owns a chunk of kTiledChunk CONSECUTIVE markers and works on the part of the window they touch in shared
// memory:
//   1. the source field of the chunk's bounding box is staged once (coalesced);
//   2. marker-centric pass (one thread per marker, no global-memory latency inside it): the 12 one-dimensional delta
//      weights are evaluated once per marker (the 64 stencil weights are their products), u_m is gathered from the
//      staged box, dF and F follow; base cell, weights and the value to spread are parked in shared memory;
//   3. cell-centric pass: every thread owns four consecutive z cells of one row of the box and sums the
//      contributions of the chunk's markers in registers, in marker order -- no shared-memory atomics (a first
//      version with fp32 shared atomics and one warp per marker was only 1.6x faster than the untiled kernel);
//   4. one coalesced vector reduction per touched cell adds the box to the window field.
// A finely meshed surface puts ~20 stencil points on every window cell, so ~64 global reductions per marker become
// ~3 per marker.  Chunks whose bounding box does not fit (markers not stored in a spatially coherent order) fall back
// to global gathers / reductions, CTA by CTA.
#ifndef VSB_TILE_CELLS
#define VSB_TILE_CELLS 2304
#endif
#ifndef VSB_TILED_CTAS
#define VSB_TILED_CTAS 3
#endif
constexpr int kTiledChunk = 256;                // markers per CTA = threads per CTA
constexpr int kTileCells = VSB_TILE_CELLS;      // cells of the staged box: 2304 x 16 B = 36 KB

// Per-marker records of the cell-centric pass, packed so that one visit costs 8 shared-memory loads, all but two of
// them warp-uniform (broadcast): base and value as one 128-bit word each, the x / y weights side by side, and the z
// weights ZERO-PADDED on both sides -- the weight of the marker for the cell dz + k of a unit is pz[3 + dz + k] for
// any -4 < dz < 4, so the inner loop needs no predicates.
struct TiledShared {
  int4 base[kTiledChunk];             // first stencil node (floor(x) - 1) per axis; .w unused
  float4 val[kTiledChunk];            // value spread by the marker (dF, or F at the last stage); .w unused
  float wxy[kTiledChunk][8];          // delta weights: x nodes 0..3, y nodes 0..3 (0 outside the window)
  float pz[kTiledChunk][12];          // 0 0 0 | z nodes 0..3 | 0 0 0 (+ 2 words of padding)
};

template <int DIM, bool SHARD>
__global__ void __launch_bounds__(kTiledChunk, VSB_TILED_CTAS) k_mdf_stage_tiled(const MdfParams p, const BodyUpdate bu, const ShardArg<SHARD> sh) {
  static_assert(DIM == 3, "the tiled stage is instantiated for D3Q19 bodies only");
  constexpr int NC = 4;
  extern __shared__ float4 s_dyn[];
  float4* s_src = s_dyn;                                             // kTileCells float4
  TiledShared& sm = *reinterpret_cast<TiledShared*>(s_dyn + kTileCells);
  __shared__ int s_lo[3], s_hi[3];
  __shared__ float s_force[3];
  const int tid = threadIdx.x;
  if (tid < 3) { s_lo[tid] = INT_MAX; s_hi[tid] = INT_MIN; s_force[tid] = 0.f; }
  __syncthreads();

  const long long gthread = (long long)blockIdx.x * blockDim.x + tid;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long wcells = (long long)p.wsize[0] * p.wsize[1] * p.wsize[2];
  const int stage = p.stage;
  const bool last = stage == p.n_iter - 1;
  int org[3] = {p.origin0[0], p.origin0[1], p.origin0[2]};
  float disp[3] = {0.f, 0.f, 0.f};
  if (p.body) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { org[d] = p.body->origin2[p.parity][d]; disp[d] = p.body->d[d]; }
  }
  // clear the other parity's buffers for the next step (as k_mdf_stage does)
  if (stage == 0) clear_field<NC>(p, p.g_win_next, org[0], 0, gthread, nthreads);
  if (stage < p.n_iter - 1) clear_field<NC>(p, p.scratch_next + (long long)stage * NC * wcells, org[0], 1, gthread, nthreads);

  const int chunk = p.chunk_begin + (int)blockIdx.x;
  const long long m_begin = p.chunk_offsets ? (long long)p.chunk_offsets[chunk] : p.m_begin + (long long)blockIdx.x * kTiledChunk;
  const int n_here = p.chunk_offsets ? min(p.chunk_offsets[chunk + 1] - (int)m_begin, kTiledChunk)
                                     : (int)min((long long)kTiledChunk, p.m_end - m_begin);
  // bounding box of the chunk's stencils (window-local cells), clipped to the window
  {
    int bmin[3] = {INT_MAX, INT_MAX, INT_MAX}, bmax[3] = {INT_MIN, INT_MIN, INT_MIN};
    if (tid < n_here) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float x = p.markers0[(m_begin + tid) * 3 + d] + disp[d] - (float)org[d];
        bmin[d] = bmax[d] = (int)floorf(x);
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      bmin[d] = __reduce_min_sync(0xffffffffu, bmin[d]);
      bmax[d] = __reduce_max_sync(0xffffffffu, bmax[d]);
      if ((tid & 31) == 0 && bmin[d] != INT_MAX) { atomicMin(&s_lo[d], bmin[d] - 1); atomicMax(&s_hi[d], bmax[d] + 2); }
    }
  }
  __syncthreads();
  int lo[3], ext[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo[d] = max(s_lo[d], 0);
    ext[d] = max(min(s_hi[d], p.wsize[d] - 1) - lo[d] + 1, 0);
  }
  const long long box = (long long)ext[0] * ext[1] * ext[2];
  const bool tiled = box > 0 && box <= kTileCells;
  const int tile_cells = tiled ? (int)box : 0;

  const float4* src = reinterpret_cast<const float4*>(stage == 0 ? p.u_win : p.scratch + (long long)(stage - 1) * NC * wcells);
  float4* dst = reinterpret_cast<float4*>(last ? p.g_win : p.scratch + (long long)stage * NC * wcells);
  for (int i = tid; i < tile_cells; i += kTiledChunk) {
    const int tz = i % ext[2], r = i / ext[2];
    const int ty = r % ext[1], tx = r / ext[1];
    s_src[i] = __ldcg(src + ((long long)(lo[0] + tx) * p.wsize[1] + (lo[1] + ty)) * p.wsize[2] + (lo[2] + tz));
  }
  __syncthreads();

  // ---- pass 2: one thread per marker -- weights, interpolation from the staged box, dF and F
  float fsum[3] = {0.f, 0.f, 0.f};
  if (tid < n_here) {
    const long long m = m_begin + tid;
    float x[3], wgt[3][4];
    int b0[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      x[d] = p.markers0[m * 3 + d] + disp[d] - (float)org[d];
      b0[d] = (int)floorf(x[d]) - 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int node = b0[d] + k;
        float wl = delta(p.delta_kind, (float)node - x[d]);
        if (node < 0 || node >= p.wsize[d]) wl = 0.f;     // node outside the window: the reference's stencil skips it
        wgt[d][k] = wl;
        if (d < 2) sm.wxy[tid][4 * d + k] = wl;
        else sm.pz[tid][3 + k] = wl;
      }
    }
    sm.base[tid] = make_int4(b0[0], b0[1], b0[2], 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) sm.pz[tid][k] = sm.pz[tid][7 + k] = 0.f;
    float um[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int jx = 0; jx < 4; ++jx) {                                // 16 points per trip keeps the register count down
      const float wx = sm.wxy[tid][jx];
#pragma unroll
      for (int jy = 0; jy < 4; ++jy)
#pragma unroll
        for (int jz = 0; jz < 4; ++jz) {
          const float w = (wgt[2][jz] * wgt[1][jy]) * wx;           // the product order of k_mdf_stage
          if (w != 0.f) {                                           // also excludes every node outside the window
            float4 v;
            if (tiled) v = s_src[((b0[0] + jx - lo[0]) * ext[1] + (b0[1] + jy - lo[1])) * ext[2] + (b0[2] + jz - lo[2])];
            else v = __ldcg(src + ((long long)(b0[0] + jx) * p.wsize[1] + (b0[1] + jy)) * p.wsize[2] + (b0[2] + jz));
            um[0] += w * v.x; um[1] += w * v.y; um[2] += w * v.z;
          }
        }
    }
    const float ds2 = (p.ds_ptr ? p.ds_ptr[m] : p.ds_value) * 2.0f;
    float spread_val[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float tgt = p.u_target ? p.u_target[m * 3 + c] : (p.body ? p.body->v[c] : 0.f);
      const float u_prev = stage > 0 ? p.marker_u[m * 3 + c] : 0.f;
      const float F_prev = stage > 0 ? p.marker_force[m * 3 + c] : 0.f;
      const float u_m = (stage == 0) ? um[c] : u_prev + 0.5f * um[c];
      const float dF = (tgt - u_m) * ds2;
      const float F = F_prev + dF;
      spread_val[c] = last ? F : dF;
      p.marker_u[m * 3 + c] = u_m;
      p.marker_force[m * 3 + c] = F;
    }
    sm.val[tid] = make_float4(spread_val[0], spread_val[1], spread_val[2], 0.f);
    if (!tiled) {                                     // box too large for the tile: global vector reductions
#pragma unroll 1
      for (int jx = 0; jx < 4; ++jx)
#pragma unroll
        for (int jy = 0; jy < 4; ++jy)
#pragma unroll
          for (int jz = 0; jz < 4; ++jz) {
            const float w = (wgt[2][jz] * wgt[1][jy]) * sm.wxy[tid][jx];
            if (w != 0.f) {
              const long long ci = ((long long)(b0[0] + jx) * p.wsize[1] + (b0[1] + jy)) * p.wsize[2] + (b0[2] + jz);
              const float4 val = make_float4(spread_val[0] * w, spread_val[1] * w, spread_val[2] * w, 0.f);
              if constexpr (SHARD) shard_add<float4>(sh, org[0], b0[0] + jx, b0[1] + jy, b0[2] + jz, ci, val);
              else atomicAdd(dst + ci, val);
            }
          }
    }
    if (last && p.body) {
#pragma unroll
      for (int c = 0; c < 3; ++c) fsum[c] = spread_val[c];
    }
  }
  if (last && p.body) {   // total force of the chunk: warp shuffles, then one shared atomic per warp and component
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) fsum[c] += __shfl_xor_sync(0xffffffffu, fsum[c], o);
      if ((tid & 31) == 0) atomicAdd(&s_force[c], fsum[c]);
    }
  }
  __syncthreads();

  // ---- pass 3 + 4: cell-centric accumulation in registers (a thread owns 4 consecutive z cells of one (x, y) row of
  // the box and sums the chunk's markers in marker order), then one vector reduction per touched cell.  The lanes of a
  // warp own different rows of the SAME z group, so the z test is warp-uniform; when the chunk's markers are sorted by
  // z (they are, for chunks cut by the host) each unit only visits the markers whose stencil reaches its z group.
  if (tiled) {
    const int my_bz = tid < n_here ? sm.base[tid].z : INT_MAX;
    const int next_bz = tid + 1 < n_here ? sm.base[tid + 1].z : INT_MAX;
    const bool z_sorted = __syncthreads_and(my_bz <= next_bz) != 0;
    const int zg = (ext[2] + 3) >> 2;
    const int n_rows = ext[0] * ext[1];
    const int n_units = n_rows * zg;
    for (int unit = tid; unit < n_units; unit += kTiledChunk) {
      const int zgi = unit / n_rows, row = unit - zgi * n_rows;
      const int cx = lo[0] + row / ext[1], cy = lo[1] + row % ext[1], cz0 = lo[2] + 4 * zgi;
      int m_lo = 0, m_hi = n_here;
      if (z_sorted) {          // markers with cz0 - 3 <= base_z <= cz0 + 3
        int a = 0, b = n_here;
        while (a < b) { const int mid = (a + b) >> 1; if (sm.base[mid].z < cz0 - 3) a = mid + 1; else b = mid; }
        m_lo = a;
        b = n_here;
        while (a < b) { const int mid = (a + b) >> 1; if (sm.base[mid].z <= cz0 + 3) a = mid + 1; else b = mid; }
        m_hi = a;
      }
      float acc[4][3];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
      for (int mi = m_lo; mi < m_hi; ++mi) {
        const int4 b = sm.base[mi];                                                                  // broadcast read
        const int dz = cz0 - b.z;                                                                    // warp-uniform
        const unsigned jx = (unsigned)(cx - b.x), jy = (unsigned)(cy - b.y);
        if (dz > -4 && dz < 4 && jx < 4u && jy < 4u) {
          const float wx = sm.wxy[mi][jx], wy = sm.wxy[mi][4 + jy];
          const float4 v = sm.val[mi];
          const float* pz = &sm.pz[mi][3 + dz];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float w = (pz[k] * wy) * wx;                                                       // 0 outside the stencil
            acc[k][0] += v.x * w; acc[k][1] += v.y * w; acc[k][2] += v.z * w;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (cz0 + k < lo[2] + ext[2] && (acc[k][0] != 0.f || acc[k][1] != 0.f || acc[k][2] != 0.f)) {
          const long long ci = ((long long)cx * p.wsize[1] + cy) * p.wsize[2] + (cz0 + k);
          const float4 val = make_float4(acc[k][0], acc[k][1], acc[k][2], 0.f);
          if constexpr (SHARD) shard_add<float4>(sh, org[0], cx, cy, cz0 + k, ci, val);
          else atomicAdd(dst + ci, val);
        }
    }
  }

  if (last && p.body) {
    if (tid < 3) atomicAdd(&p.body->force_sum[tid], s_force[tid]);
    if (p.update_body || p.host_mail) {
      __shared__ int s_last;
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        s_last = (atomicAdd(&p.body->ticket, 1) == (int)gridDim.x - 1);
      }
      __syncthreads();
      if (s_last && tid == 0) {
        __threadfence();
        p.body->ticket = 0;
        finish_body(p, bu);
      }
    }
  }
}

// body update by one thread (see body_update in vsb_step.cuh)
__global__ void k_body_newmark(VsbBodyState* b, BodyUpdate u, int parity) {
  if (threadIdx.x == 0 && blockIdx.x == 0) body_update(b, u, parity);
}

template <int DIM>
static int mdf_impl(const VsbStepArgs& sa, const VsbMdfArgs& a, const VsbBodyParams* bp, cudaStream_t stream) {
  StepParams<DIM> sp;
  VsbStepArgs b = sa;
  if (!b.f_out) b.f_out = a.g_win;   // unused by these kernels; only has to differ from f_in
  b.band = 0;
  if (int rc = fill_params<DIM>(b, sp)) return rc;
  MdfParams p;
  p.delta_kind = a.delta_kind; p.n_iter = a.n_iter; p.parity = a.parity & 1; p.n_markers = a.n_markers;
  for (int d = 0; d < 3; ++d) { p.origin0[d] = a.win_origin0[d]; p.wsize[d] = a.win_size[d]; }
  p.markers0 = a.markers0; p.u_target = a.u_target; p.ds_ptr = a.ds_ptr; p.ds_value = a.ds_value;
  p.u_win = a.u_win; p.g_win = a.g_win; p.g_win_next = a.g_win_next; p.scratch = a.scratch; p.scratch_next = a.scratch_next;
  p.marker_u = a.marker_u; p.marker_force = a.marker_force; p.body = a.body;
  p.update_body = (a.body && bp && bp->n_dof > 0) ? 1 : 0;
  p.host_mail = (a.body && !p.update_body) ? a.host_mail : nullptr;
  p.mail_seq = a.mail_seq;
  p.chunk_offsets = nullptr;
  p.m_begin = 0; p.m_end = a.n_markers; p.chunk_begin = 0; p.clear_mode = 0;
  if (a.reach_cells && a.n_reach_cells > 0) {   // clear only the cells a stencil can reach (the rest is never written)
    p.clear_mode = 2;
    p.clear_cells = a.reach_cells; p.n_clear_cells = a.n_reach_cells;
    for (int d = 0; d < 3; ++d) { p.box_lo[d] = 0; p.box_hi[d] = d < DIM ? a.win_size[d] : 1; }
  }
  p.nbr_list = a.nbr_list; p.nbr_stride = a.nbr_stride;
  p.rotation = (DIM == 2 && a.rotation) ? 1 : 0;
  p.center[0] = a.center[0]; p.center[1] = a.center[1];
  BodyUpdate bu{};
  if (p.update_body) bu = make_body_update(*bp, DIM);
  const int lanes = (DIM == 2) ? 16 : 32;
  const unsigned nb = blocks_for(a.n_markers * lanes, kBlock);
  p.barrier = reinterpret_cast<unsigned long long*>(a.barrier);
  p.stage = 0; p.stage_end = a.n_iter;
  const int mode = a.chain_mode;
  bool use_cluster = false;
  if constexpr (DIM == 2) {
    // Small 2-D bodies, on request (chain_mode 4; VSB_MDF_CTA=1 makes it the automatic choice): the whole chain in ONE
    // CTA, work field in shared memory, spread as a gather over per-cell buckets (vsb_mdf_cta.cu) -- no floating-point
    // atomics, bit-reproducible.  Measured on C2: 60 us against 14-18 us for the grid-barrier chain (one SM is not
    // enough for 8192 stencil points x 5 iterations), so it is not the default.
    static const bool cta_auto = [] { const char* e = getenv("VSB_MDF_CTA"); return e && e[0] == '1'; }();
    const bool use_cta = (mode == 4 || (mode == 0 && cta_auto)) && mdf_cta2d_supported(p);
    if (mode == 4 && !use_cta) {
      set_error("vsb_ib_mdf: chain_mode 4 (one CTA) needs a 2-D body of at most 512 markers in a window of fewer than "
                "~19000 cells");
      return VSB_ERR_INVALID;
    }
    if (use_cta) return launch_mdf_cta2d(sp, p, bu, stream);
    // Small 2-D bodies: the whole chain in one thread-block cluster, work fields in distributed shared memory,
    // hardware cluster barriers between the iterations (vsb_mdf_cluster.cu)
    use_cluster = mode == 2 && a.u_win == nullptr && mdf_cluster2d_supported(p);   // (not the automatic choice: 21.5 us against 18 for the grid-barrier chain)
    if (mode == 2 && !use_cluster) {
      set_error("vsb_ib_mdf: chain_mode 2 (cluster) needs a 2-D body of at most 512 markers, no u_win, and the "
                "neighbour list (nbr_list, nbr_stride <= 48)");
      return VSB_ERR_INVALID;
    }
    if (use_cluster) return launch_mdf_cluster2d(sp, p, bu, stream);
  } else if (mode == 2 || mode == 4) {
    set_error("vsb_ib_mdf: chain_mode 2 (cluster) and 4 (one CTA) are for 2-D bodies");
    return VSB_ERR_INVALID;
  }
  // Small bodies: every iteration in ONE launch, separated by grid barriers.  The launch is cooperative, so all of its
  // CTAs (at most 120 of 128 threads) are co-resident whatever else runs on the device.  Large bodies: one launch
  // per iteration.
  if (a.barrier && nb <= 120 && mode != 3) {   // (one marker per lane group: the kernel keeps u_m and F in registers across stages)
    ShardArg<false> none;
    void* kargs[] = {(void*)&sp, (void*)&p, (void*)&bu, (void*)&none};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_mdf_stage<DIM, false>, dim3(nb), dim3(kBlock), kargs, 0, stream);
    if (e != cudaSuccess) return cuda_fail(e, "vsb_ib_mdf (cooperative launch)");
  } else if (DIM == 3 && a.u_win != nullptr && !getenv("VSB_MDF_UNTILED")) {
    // dense body with a precomputed window velocity: every stage reads a window field -> shared-memory tiles
    if constexpr (DIM == 3) {
      constexpr size_t smem = (size_t)kTileCells * sizeof(float4) + sizeof(TiledShared);
      static bool configured = false;
      if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_mdf_stage_tiled<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "vsb_ib_mdf (shared-memory opt-in)");
        configured = true;
      }
      const bool cut = a.chunk_offsets != nullptr && a.n_chunks > 0;
      p.chunk_offsets = cut ? a.chunk_offsets : nullptr;
      const unsigned nbt = cut ? (unsigned)a.n_chunks : blocks_for(a.n_markers, kTiledChunk);
      for (int k = 0; k < a.n_iter; ++k) {
        p.stage = k; p.stage_end = k + 1;
        k_mdf_stage_tiled<3, false><<<nbt, kTiledChunk, smem, stream>>>(p, bu, ShardArg<false>{});
      }
    }
  } else {
    // One launch per iteration.  The chain is latency-bound and runs beside the bulk kernel; with one CTA per marker
    // group it would take every SM's registers for itself (118 registers x 128 threads, four or more CTAs per SM)
    // and stall the bulk.  Capped at kMdfCtasPerSm CTAs per SM, the groups walk the markers in batches instead.
    static const int cap = [] {
      const char* e = getenv("VSB_MDF_GRID_CAP");
      if (e) return atoi(e);
      int dev = 0, n_sm = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      return kMdfCtasPerSm * n_sm;
    }();
    const unsigned nbc = (cap > 0 && nb > (unsigned)cap) ? (unsigned)cap : nb;
    for (int k = 0; k < a.n_iter; ++k) {
      p.stage = k; p.stage_end = k + 1;
      k_mdf_stage<DIM, false><<<nbc, kBlock, 0, stream>>>(sp, p, bu, ShardArg<false>{});
    }
  }
  VSB_LAUNCH_CHECK("vsb_ib_mdf");
  return VSB_OK;
}

// One iteration (p.stage) of this rank's share of a sharded chain: the same kernels with the multicast spread.
// Stage 0 reads p.u_win (never the populations), so StepParams is a dummy.
int launch_mdf_stage_sharded(int dim, const MdfParams& p_in, const ShardDev& shd, bool tiled, unsigned n_chunks, cudaStream_t stream) {
  ShardArg<true> sh;
  static_cast<ShardDev&>(sh) = shd;
  BodyUpdate bu{};
  MdfParams p = p_in;
  p.update_body = 0;          // the partial sums are combined and the body advanced after the last flag barrier
  p.host_mail = nullptr;
  const long long n_mine = p.m_end - p.m_begin;
  if (n_mine <= 0) return VSB_OK;
  if (dim == 3 && tiled) {
    constexpr size_t smem = (size_t)kTileCells * sizeof(float4) + sizeof(TiledShared);
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(k_mdf_stage_tiled<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "vsb_ibshard_chain (shared-memory opt-in)");
      configured = true;
    }
    const unsigned nbt = p.chunk_offsets ? n_chunks : blocks_for(n_mine, kTiledChunk);
    if (nbt > 0) k_mdf_stage_tiled<3, true><<<nbt, kTiledChunk, smem, stream>>>(p, bu, sh);
  } else if (dim == 3) {
    StepParams<3> sp{};
    k_mdf_stage<3, true><<<blocks_for(n_mine * 32, kBlock), kBlock, 0, stream>>>(sp, p, bu, sh);
  } else {
    StepParams<2> sp{};
    k_mdf_stage<2, true><<<blocks_for(n_mine * 16, kBlock), kBlock, 0, stream>>>(sp, p, bu, sh);
  }
  VSB_LAUNCH_CHECK("vsb_ibshard_chain (iteration)");
  return VSB_OK;
}

}  // namespace vsb

using namespace vsb;

extern "C" {

int vsb_ib_delta(int kind, int64_t n, const float* r, float* out, vsb_stream_t stream) {
  VSB_REQUIRE(kind >= VSB_DELTA_PESKIN3 && kind <= VSB_DELTA_HAT2, "unknown delta kernel %d", kind);
  VSB_REQUIRE(n >= 0 && r && out, "vsb_ib_delta: bad argument");
  if (n == 0) return VSB_OK;
  k_delta<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(kind, n, r, out);
  VSB_LAUNCH_CHECK("vsb_ib_delta");
  return VSB_OK;
}

int vsb_ib_stencil(int dim, int kind, int radius, int64_t n_markers, const float* coords, int ny, int nz, float* weights,
                   int32_t* indices, vsb_stream_t stream) {
  VSB_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3, got %d", dim);
  VSB_REQUIRE(kind >= VSB_DELTA_PESKIN3 && kind <= VSB_DELTA_HAT2, "unknown delta kernel %d", kind);
  VSB_REQUIRE(radius >= 1 && radius <= 4, "stencil_radius must be in 1..4, got %d", radius);
  VSB_REQUIRE(n_markers >= 0 && coords && weights && indices, "vsb_ib_stencil: bad argument");
  if (n_markers == 0) return VSB_OK;
  const int side = 2 * radius;
  const long long total = n_markers * (dim == 2 ? side * side : side * side * side);
  if (dim == 2) k_stencil<2><<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(kind, radius, n_markers, coords, ny, nz, weights, indices);
  else k_stencil<3><<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(kind, radius, n_markers, coords, ny, nz, weights, indices);
  VSB_LAUNCH_CHECK("vsb_ib_stencil");
  return VSB_OK;
}

int vsb_ib_interpolate(int n_comp, int64_t n_cells, const float* grid, int64_t n_markers, int n_stencil, const float* weights,
                       const int32_t* indices, float* out, vsb_stream_t stream) {
  VSB_REQUIRE(n_comp > 0 && n_cells > 0 && n_markers >= 0 && n_stencil > 0 && grid && weights && indices && out,
              "vsb_ib_interpolate: bad argument");
  if (n_markers == 0) return VSB_OK;
  k_interpolate<<<blocks_for(n_markers * 32, kBlock), kBlock, 0, (cudaStream_t)stream>>>(n_comp, n_cells, grid, n_markers,
                                                                                      n_stencil, weights, indices, out);
  VSB_LAUNCH_CHECK("vsb_ib_interpolate");
  return VSB_OK;
}

int vsb_ib_spread(int n_comp, int64_t n_cells, float* grid, int64_t n_markers, int n_stencil, const float* values,
                  const float* weights, const int32_t* indices, vsb_stream_t stream) {
  VSB_REQUIRE(n_comp > 0 && n_cells > 0 && n_markers >= 0 && n_stencil > 0 && grid && values && weights && indices,
              "vsb_ib_spread: bad argument");
  if (n_markers == 0) return VSB_OK;
  k_spread<<<blocks_for(n_markers * n_stencil, 256), 256, 0, (cudaStream_t)stream>>>(n_comp, n_cells, grid, n_markers, n_stencil,
                                                                                 values, weights, indices);
  VSB_LAUNCH_CHECK("vsb_ib_spread");
  return VSB_OK;
}

int vsb_ib_mdf(const VsbStepArgs* args, const VsbMdfArgs* a, const VsbBodyParams* params, vsb_stream_t stream) {
  VSB_REQUIRE(args != nullptr && a != nullptr, "vsb_ib_mdf: null args");
  VSB_REQUIRE((a->dim == 2 || a->dim == 3) && a->dim == args->grid.dim, "vsb_ib_mdf: dim must be 2 or 3 and match the grid");
  VSB_REQUIRE(a->delta_kind >= VSB_DELTA_PESKIN3 && a->delta_kind <= VSB_DELTA_HAT2, "unknown delta kernel %d", a->delta_kind);
  VSB_REQUIRE(a->n_iter >= 1, "n_iter must be >= 1, got %d", a->n_iter);
  VSB_REQUIRE(a->n_markers >= 0 && a->markers0 && a->g_win && a->g_win_next && a->marker_u && a->marker_force,
              "vsb_ib_mdf: null buffer");
  VSB_REQUIRE(a->n_iter == 1 || (a->scratch && a->scratch_next), "vsb_ib_mdf: n_iter > 1 needs the scratch buffers");
  for (int d = 0; d < a->dim; ++d) VSB_REQUIRE(a->win_size[d] >= 4, "IB window must be at least 4 cells wide");
  if (a->n_markers == 0) return VSB_OK;
  return a->dim == 2 ? mdf_impl<2>(*args, *a, params, (cudaStream_t)stream) : mdf_impl<3>(*args, *a, params, (cudaStream_t)stream);
}

// dyn.py:27-51 on the host copy of the body state (shared by the two host-ODE entry points)
static void host_body_update(VsbBodyState* pinned, const VsbBodyParams* bp, int parity);

int vsb_body_newmark_host(VsbBodyState* body, VsbBodyState* pinned, const VsbBodyParams* bp, int parity,
                          vsb_stream_t stream) {
  VSB_REQUIRE(body && pinned && bp, "vsb_body_newmark_host: null argument");
  VSB_REQUIRE(bp->n_dof >= 1 && bp->n_dof <= 3, "n_dof must be 1..3, got %d", bp->n_dof);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(pinned, body, sizeof(VsbBodyState), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return cuda_fail(e, "vsb_body_newmark_host (device -> host)");
  host_body_update(pinned, bp, parity);
  e = cudaMemcpyAsync(body, pinned, sizeof(VsbBodyState), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return cuda_fail(e, "vsb_body_newmark_host (host -> device)");
  return VSB_OK;
}

static void host_body_update(VsbBodyState* pinned, const VsbBodyParams* bp, int parity) {
  // dyn.py:27-51 with gamma = 1/2, beta = 1/4, dt = 1, in fp32 like the reference's jnp arithmetic;
  // h = sum(-F) + a * added_mass (examples/2d/vortex_induced_vibration.py:135-136).  bp->history is a HOST ring here.
  const int dim = bp->grid_size[2] <= 1 ? 2 : 3;
  body_update(pinned, make_body_update(*bp, dim), parity & 1);
}

// One domain of vsb_run_host_ode_multi: what is enqueued for a step, and what happens when its force has arrived.
namespace {
struct HostOdeDomain {
  VsbStepArgs* a; VsbMdfArgs* mdf; const VsbBodyParams* bp; VsbBodyState* pinned;
  cudaStream_t main, ib, edge;
  cudaEvent_t fork, ib_done, edge_done;
  int has_edges, want, steps_done, in_flight, n_steps;
  cudaGraphExec_t graph[2];    // by step parity; null: launch kernel by kernel
  const void* plan;            // the caller's VsbHostPlan: key of the graph cache
  bool graphs_from_cache;
};

// The two step graphs of a domain are kept from one vsb_run_host_ode[_multi] call to the next (a driver loop calls once
// per chunk of a few hundred steps; capturing and instantiating 2 graphs per domain and chunk cost ~5 % of such a
// chunk).  Cached per VsbHostPlan; valid as long as everything a captured step depends on is unchanged -- both argument
// blocks (normalised to parity 0), the face operations, the MRT operators, streams and the page-locked state buffer.
struct HostOdeGraphs {
  std::vector<unsigned char> key;
  cudaGraphExec_t graph[2];
};
std::mutex g_host_ode_mu;
std::unordered_map<const void*, HostOdeGraphs> g_host_ode_graphs;

void key_bytes(std::vector<unsigned char>& k, const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  k.insert(k.end(), b, b + n);
}

std::vector<unsigned char> host_ode_key(const HostOdeDomain& d) {
  VsbStepArgs a = *d.a;
  VsbMdfArgs m = *d.mdf;
  if (m.parity & 1) {   // what alternates with the parity, put back to parity 0
    const float* fi = a.f_in; a.f_in = a.f_out; a.f_out = const_cast<float*>(fi);
    float* t = m.g_win; m.g_win = m.g_win_next; m.g_win_next = t;
    t = m.scratch; m.scratch = m.scratch_next; m.scratch_next = t;
  }
  m.parity = 0; m.mail_seq = 0; a.parity = 0; a.g_win = m.g_win; a.band = 0;
  std::vector<unsigned char> k;
  key_bytes(k, &a, sizeof(a));
  key_bytes(k, &m, sizeof(m));
  if (a.n_post > 0 && a.post) key_bytes(k, a.post, sizeof(VsbPostOp) * (size_t)a.n_post);
  const size_t qq = (a.grid.dim == 2 ? 81 : 361) * sizeof(float);
  if (a.mrt_op_host) key_bytes(k, a.mrt_op_host, qq);
  if (a.mrt_fop_host) key_bytes(k, a.mrt_fop_host, qq);
  const void* handles[5] = {d.main, d.ib, d.edge, d.pinned, d.mdf->host_mail};
  key_bytes(k, handles, sizeof(handles));
  return k;
}

// Graphs of an earlier call with the same plan and an identical step, if any.
bool host_ode_cached_graphs(HostOdeDomain& d) {
  const std::vector<unsigned char> key = host_ode_key(d);
  std::lock_guard<std::mutex> lock(g_host_ode_mu);
  auto it = g_host_ode_graphs.find(d.plan);
  if (it == g_host_ode_graphs.end()) return false;
  if (it->second.key != key) {   // the step changed: those graphs are stale
    for (int k = 0; k < 2; ++k)
      if (it->second.graph[k]) cudaGraphExecDestroy(it->second.graph[k]);
    g_host_ode_graphs.erase(it);
    return false;
  }
  d.graph[0] = it->second.graph[0];
  d.graph[1] = it->second.graph[1];
  d.graphs_from_cache = true;
  return true;
}

// End of a call: hand the graphs to the cache (or drop them when the domain ended on an error).
void host_ode_keep_graphs(HostOdeDomain& d, bool ok) {
  if (!d.graph[0] && !d.graph[1]) return;
  if (ok && d.graph[0] && d.graph[1] && !d.in_flight) {
    if (!d.graphs_from_cache) {
      HostOdeGraphs g;
      g.key = host_ode_key(d);
      g.graph[0] = d.graph[0]; g.graph[1] = d.graph[1];
      std::lock_guard<std::mutex> lock(g_host_ode_mu);
      auto it = g_host_ode_graphs.find(d.plan);
      if (it != g_host_ode_graphs.end()) {
        for (int k = 0; k < 2; ++k)
          if (it->second.graph[k]) cudaGraphExecDestroy(it->second.graph[k]);
        g_host_ode_graphs.erase(it);
      }
      g_host_ode_graphs.emplace(d.plan, std::move(g));
    }
  } else {
    if (d.graphs_from_cache) {
      std::lock_guard<std::mutex> lock(g_host_ode_mu);
      g_host_ode_graphs.erase(d.plan);
    }
    for (int k = 0; k < 2; ++k)
      if (d.graph[k]) cudaGraphExecDestroy(d.graph[k]);
  }
  d.graph[0] = d.graph[1] = nullptr;
}

// The device work of one step: IB chain (posts the force into the mailbox) and window band on `ib`, bulk on `main`,
// wall layers on `edge`.  join = true (graph capture) also folds the side streams back into `main`.
int host_ode_kernels(HostOdeDomain& d, bool join) {
  VsbStepArgs* a = d.a; VsbMdfArgs* mdf = d.mdf;
  cudaError_t e;
  int rc;
  const int par = mdf->parity & 1;
  a->parity = par;
  a->g_win = mdf->g_win;
  if ((e = cudaEventRecord(d.fork, d.main)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (fork)");
  if ((e = cudaStreamWaitEvent(d.ib, d.fork, 0)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (fork)");
  if ((rc = vsb_ib_mdf(a, mdf, nullptr, d.ib))) return rc;        // first: the host waits for this chain
  a->band = 1;                                                     // the bulk runs while the host advances the body
  if ((rc = vsb_step(a, d.main))) return rc;
  if (d.has_edges) {
    if ((e = cudaStreamWaitEvent(d.edge, d.fork, 0)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (edge fork)");
    if ((rc = vsb_edge_fused(a, d.edge))) return rc;
    if ((e = cudaEventRecord(d.edge_done, d.edge)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (edge join)");
  }
  // The window band of THIS step needs the chain's force field but not the body update (the window origin it reads
  // is the entry of this step's parity, written one step ago): enqueue it now, so that it runs while the force
  // travels to the host and the new body state travels back.
  a->band = 2;
  if ((rc = vsb_step(a, d.ib))) return rc;
  a->band = 0;
  if (join) {
    if ((e = cudaEventRecord(d.ib_done, d.ib)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (ib join)");
    if ((e = cudaStreamWaitEvent(d.main, d.ib_done, 0)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (join)");
    if (d.has_edges && (e = cudaStreamWaitEvent(d.main, d.edge_done, 0)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (join)");
  }
  return VSB_OK;
}

// next step: swap the population buffers and the parity-double-buffered IB fields
void host_ode_flip(HostOdeDomain& d) {
  VsbStepArgs* a = d.a; VsbMdfArgs* mdf = d.mdf;
  const float* fi = a->f_in; a->f_in = a->f_out; a->f_out = const_cast<float*>(fi);
  float* t = mdf->g_win; mdf->g_win = mdf->g_win_next; mdf->g_win_next = t;
  t = mdf->scratch; mdf->scratch = mdf->scratch_next; mdf->scratch_next = t;
  mdf->parity = (mdf->parity & 1) ^ 1;
  a->parity = mdf->parity;
  a->g_win = mdf->g_win;
}

int host_ode_enqueue(HostOdeDomain& d) {
  VsbHostMail* mail = d.mdf->host_mail;
  const int par = d.mdf->parity & 1;
  if (d.graph[par]) {
    // graph replay: the kernels post the number of the step being taken (device copy of the step counter + 1)
    d.want = d.pinned->step + 1;
    *reinterpret_cast<volatile int*>(&mail->seq) = -1;      // nothing in flight for this domain: no race with the device
    cudaError_t e = cudaGraphLaunch(d.graph[par], d.main);
    if (e != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (graph launch)");
  } else {
    d.mdf->mail_seq = mail->next;
    d.want = d.mdf->mail_seq;
    *reinterpret_cast<volatile int*>(&mail->seq) = -1;
    if (int rc = host_ode_kernels(d, false)) return rc;
  }
  d.in_flight = 1;
  return VSB_OK;
}

// The force of the step in flight has been posted: advance the body on the CPU, send the state back, close the step.
int host_ode_complete(HostOdeDomain& d) {
  VsbMdfArgs* mdf = d.mdf;
  VsbHostMail* mail = mdf->host_mail;
  cudaError_t e;
  const int par = mdf->parity & 1;
  mail->next = d.want + 1;
  for (int c = 0; c < 3; ++c) d.pinned->force_sum[c] = mail->force[c];
  host_body_update(d.pinned, d.bp, par);
  if (d.graph[par]) {
    // The graph of the NEXT step begins with the copy of the body state from this page-locked buffer (a memcpy node,
    // host_ode_capture), so a step costs the host ONE driver call -- the graph launch; the loop is bound by those
    // calls, not by the device.  Only the state after the last step of the run is sent explicitly.
    if (d.steps_done + 1 >= d.n_steps &&
        (e = cudaMemcpyAsync(mdf->body, d.pinned, sizeof(VsbBodyState), cudaMemcpyHostToDevice, d.main)) != cudaSuccess)
      return cuda_fail(e, "vsb_run_host_ode (host -> device)");
  } else {
    if ((e = cudaMemcpyAsync(mdf->body, d.pinned, sizeof(VsbBodyState), cudaMemcpyHostToDevice, d.ib)) != cudaSuccess)
      return cuda_fail(e, "vsb_run_host_ode (host -> device)");
    if ((e = cudaEventRecord(d.ib_done, d.ib)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (ib join)");
    if ((e = cudaStreamWaitEvent(d.main, d.ib_done, 0)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (join)");
    if (d.has_edges && (e = cudaStreamWaitEvent(d.main, d.edge_done, 0)) != cudaSuccess) return cuda_fail(e, "vsb_run_host_ode (join)");
  }
  host_ode_flip(d);
  d.in_flight = 0;
  d.steps_done += 1;
  return VSB_OK;
}

// Record the step of both parities as CUDA graphs (the arguments alternate with the parity and with nothing else).
// Called between steps, after the kernels have run at least once outside capture.
int host_ode_capture(HostOdeDomain& d) {
  // the legacy default stream cannot be captured: such a domain keeps launching kernel by kernel
  if (d.main == nullptr || d.main == cudaStreamLegacy || d.main == cudaStreamPerThread) return VSB_OK;
  if (host_ode_cached_graphs(d)) return VSB_OK;
  cudaError_t e;
  const int saved_seq = d.mdf->mail_seq;
  d.mdf->mail_seq = -1;
  int rc = VSB_OK, flips = 0;
  for (int k = 0; k < 2 && rc == VSB_OK; ++k) {
    const int par = d.mdf->parity & 1;
    cudaGraph_t g = nullptr;
    if ((e = cudaStreamBeginCapture(d.main, cudaStreamCaptureModeThreadLocal)) != cudaSuccess) {
      rc = cuda_fail(e, "vsb_run_host_ode (begin capture)");
      break;
    }
    // first node: the body state the host computed from the previous step's force (page-locked memory, read when the
    // node executes -- the host writes it before it launches the graph and not again until this step's force is back)
    e = cudaMemcpyAsync(d.mdf->body, d.pinned, sizeof(VsbBodyState), cudaMemcpyHostToDevice, d.main);
    rc = (e == cudaSuccess) ? host_ode_kernels(d, true) : cuda_fail(e, "vsb_run_host_ode (capture of the state copy)");
    e = cudaStreamEndCapture(d.main, &g);
    if (rc == VSB_OK && e != cudaSuccess) rc = cuda_fail(e, "vsb_run_host_ode (end capture)");
    if (rc == VSB_OK && (e = cudaGraphInstantiate(&d.graph[par], g, 0)) != cudaSuccess) {
      d.graph[par] = nullptr;
      rc = cuda_fail(e, "vsb_run_host_ode (graph instantiate)");
    }
    if (g) cudaGraphDestroy(g);
    host_ode_flip(d);
    ++flips;
  }
  if (flips & 1) host_ode_flip(d);          // leave the arguments at the parity they came with
  d.mdf->mail_seq = saved_seq;
  if (rc != VSB_OK) {
    // not capturable here (e.g. a stream that is already part of another capture): carry on kernel by kernel
    for (int k = 0; k < 2; ++k)
      if (d.graph[k]) { cudaGraphExecDestroy(d.graph[k]); d.graph[k] = nullptr; }
    cudaGetLastError();
    rc = VSB_OK;
  }
  return rc;
}

// Serve the domains i = first, first + stride, ... until each has taken n_steps steps.
int host_ode_serve(HostOdeDomain* dom, int n_domains, int stride, int first, int n_steps, int graph_after) {
  int rc;
  long long remaining = 0;
  for (int i = first; i < n_domains; i += stride) {
    if ((rc = host_ode_enqueue(dom[i]))) return rc;
    remaining += n_steps;
  }
  unsigned long long idle = 0;
  std::chrono::steady_clock::time_point t0;
  while (remaining > 0) {
    bool progress = false;
    for (int i = first; i < n_domains; i += stride) {
      HostOdeDomain& d = dom[i];
      if (!d.in_flight) continue;
      if (*reinterpret_cast<volatile int*>(&d.mdf->host_mail->seq) != d.want) continue;
      if ((rc = host_ode_complete(d))) return rc;
      --remaining;
      progress = true;
      // long runs: after both parities have run once kernel by kernel, replay them as graphs
      if (graph_after > 0 && d.steps_done == graph_after && n_steps - d.steps_done >= 2 && (rc = host_ode_capture(d))) return rc;
      if (d.steps_done < n_steps && (rc = host_ode_enqueue(d))) return rc;
    }
    if (progress) { idle = 0; continue; }
    __builtin_ia32_pause();
    if ((++idle & 0xffff) == 0) {             // every 64 K empty rounds: is the device still alive, are we out of time?
      if (idle == 0x10000) t0 = std::chrono::steady_clock::now();
      for (int i = first; i < n_domains; i += stride) {
        HostOdeDomain& d = dom[i];
        if (!d.in_flight) continue;
        // the stream the chain of the step in flight was enqueued on: `main` for a graph replay (the side streams
        // are idle then, so querying `ib` would report a finished chain while the graph is still running)
        cudaError_t e = cudaStreamQuery(d.graph[d.mdf->parity & 1] ? d.main : d.ib);
        if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "vsb_run_host_ode (waiting for the IB force)");
        if (e == cudaSuccess && *reinterpret_cast<volatile int*>(&d.mdf->host_mail->seq) != d.want) {
          set_error("vsb_run_host_ode: the IB chain finished without posting the force");
          return VSB_ERR_CUDA;
        }
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) {
        set_error("vsb_run_host_ode: timed out waiting for the IB force");
        return VSB_ERR_CUDA;
      }
    }
  }
  return VSB_OK;
}
}  // namespace

int vsb_run_host_ode_multi(int n_domains, VsbStepArgs* const* args, VsbMdfArgs* const* mdfs,
                           const VsbBodyParams* const* params, VsbBodyState* const* pinned,
                           const VsbHostPlan* const* plans, int n_steps) {
  VSB_REQUIRE(n_domains >= 1 && n_domains <= 64, "vsb_run_host_ode_multi: n_domains must be 1..64, got %d", n_domains);
  VSB_REQUIRE(args && mdfs && params && pinned && plans, "vsb_run_host_ode: null argument");
  HostOdeDomain dom[64];
  for (int i = 0; i < n_domains; ++i) {
    HostOdeDomain& d = dom[i];
    d.a = args[i]; d.mdf = mdfs[i]; d.bp = params[i]; d.pinned = pinned[i];
    const VsbHostPlan* plan = plans[i];
    VSB_REQUIRE(d.a && d.mdf && d.bp && d.pinned && plan, "vsb_run_host_ode: null argument");
    VSB_REQUIRE(d.mdf->body != nullptr && d.mdf->host_mail != nullptr, "vsb_run_host_ode: needs a body state and a host mailbox");
    VSB_REQUIRE(d.bp->n_dof >= 1 && d.bp->n_dof <= 3, "n_dof must be 1..3, got %d", d.bp->n_dof);
    VSB_REQUIRE(d.a->do_stream && d.a->do_collide, "vsb_run_host_ode: full steps only");
    d.main = (cudaStream_t)plan->main; d.ib = (cudaStream_t)plan->ib; d.edge = (cudaStream_t)plan->edge;
    d.fork = (cudaEvent_t)plan->ev_fork; d.ib_done = (cudaEvent_t)plan->ev_ib; d.edge_done = (cudaEvent_t)plan->ev_edge;
    VSB_REQUIRE(d.ib && d.fork && d.ib_done, "vsb_run_host_ode: plan needs the ib stream and the fork / ib events");
    d.has_edges = d.a->edges == 1 && d.a->n_post > 0;
    if (d.has_edges) VSB_REQUIRE(d.edge && d.edge_done, "vsb_run_host_ode: plan needs the edge stream / event");
    for (int j = 0; j < i; ++j)
      VSB_REQUIRE(dom[j].ib != d.ib && (n_domains == 1 || dom[j].main != d.main) && dom[j].mdf->host_mail != d.mdf->host_mail,
                  "vsb_run_host_ode_multi: domains %d and %d share a stream or a mailbox", j, i);
    d.want = 0; d.steps_done = 0; d.in_flight = 0; d.n_steps = n_steps;
    d.graph[0] = d.graph[1] = nullptr;
    d.plan = plan; d.graphs_from_cache = false;
  }
  if (n_steps <= 0) return VSB_OK;
  // Host threads: one thread serves all domains by default.  VSB_HOST_ODE_THREADS = T gives each of T threads a fixed
  // subset of the domains (launching kernel by kernel; measured on B200: no gain for eight 1024^2 domains, the
  // driver serialises the launches).
  int n_thr = 1;
  if (const char* e = getenv("VSB_HOST_ODE_THREADS")) n_thr = atoi(e);
  if (n_thr < 1) n_thr = 1;
  if (n_thr > n_domains) n_thr = n_domains;
  // Runs of >= 32 steps replay each domain's step as a CUDA graph (one per parity): two driver calls per step (graph
  // launch, state copy) instead of about ten.  VSB_HOST_ODE_GRAPH=0 keeps launching kernel by kernel.
  const char* ge = getenv("VSB_HOST_ODE_GRAPH");
  const int graph_after = ((ge ? atoi(ge) != 0 : true) && n_steps >= 32 && n_thr == 1) ? 2 : 0;
  auto cleanup = [&](bool ok) {
    for (int i = 0; i < n_domains; ++i) host_ode_keep_graphs(dom[i], ok);
  };
  if (n_thr == 1) {
    const int rc1 = host_ode_serve(dom, n_domains, 1, 0, n_steps, graph_after);
    cleanup(rc1 == VSB_OK);
    return rc1;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  int rcs[64];
  std::string msgs[64];
  std::vector<std::thread> workers;
  for (int t = 1; t < n_thr; ++t)
    workers.emplace_back([&, t] {
      cudaSetDevice(dev);
      rcs[t] = host_ode_serve(dom, n_domains, n_thr, t, n_steps, graph_after);
      if (rcs[t]) msgs[t] = vsb_last_error();      // the error string is thread-local: hand it to the caller's thread
    });
  rcs[0] = host_ode_serve(dom, n_domains, n_thr, 0, n_steps, graph_after);
  for (auto& w : workers) w.join();
  cleanup(false);
  for (int t = 1; t < n_thr; ++t)
    if (rcs[t]) { set_error("%s", msgs[t].c_str()); return rcs[t]; }
  return rcs[0];
}

int vsb_run_host_ode(VsbStepArgs* a, VsbMdfArgs* mdf, const VsbBodyParams* bp, VsbBodyState* pinned,
                     const VsbHostPlan* plan, int n_steps) {
  VSB_REQUIRE(a && mdf && bp && pinned && plan, "vsb_run_host_ode: null argument");
  return vsb_run_host_ode_multi(1, &a, &mdf, &bp, &pinned, &plan, n_steps);
}

// ---- non-blocking host ODE: the Newmark update runs as a stream-ordered host function
namespace {
struct HostOdeCall {            // lives from the enqueue until the last host function of the call has run
  VsbBodyParams bp;
  VsbBodyState* pinned;
  VsbHostMail* mail;
};
struct HostOdeTick { HostOdeCall* call; int parity; int last; };

void CUDART_CB host_ode_tick(void* user) {
  HostOdeTick* t = static_cast<HostOdeTick*>(user);
  HostOdeCall* c = t->call;
  for (int k = 0; k < 3; ++k) c->pinned->force_sum[k] = reinterpret_cast<volatile float*>(c->mail->force)[k];
  host_body_update(c->pinned, &c->bp, t->parity);
  if (t->last) delete c;
  delete t;
}
}  // namespace

int vsb_enqueue_host_ode(VsbStepArgs* a, VsbMdfArgs* mdf, const VsbBodyParams* bp, VsbBodyState* pinned,
                         const VsbHostPlan* plan, int n_steps) {
  VSB_REQUIRE(a && mdf && bp && pinned && plan, "vsb_enqueue_host_ode: null argument");
  VSB_REQUIRE(mdf->body != nullptr && mdf->host_mail != nullptr, "vsb_enqueue_host_ode: needs a body state and a host mailbox");
  VSB_REQUIRE(bp->n_dof >= 1 && bp->n_dof <= 3, "n_dof must be 1..3, got %d", bp->n_dof);
  VSB_REQUIRE(a->do_stream && a->do_collide, "vsb_enqueue_host_ode: full steps only");
  HostOdeDomain d;
  d.a = a; d.mdf = mdf; d.bp = bp; d.pinned = pinned;
  d.main = (cudaStream_t)plan->main; d.ib = (cudaStream_t)plan->ib; d.edge = (cudaStream_t)plan->edge;
  d.fork = (cudaEvent_t)plan->ev_fork; d.ib_done = (cudaEvent_t)plan->ev_ib; d.edge_done = (cudaEvent_t)plan->ev_edge;
  VSB_REQUIRE(d.ib && d.fork && d.ib_done, "vsb_enqueue_host_ode: plan needs the ib stream and the fork / ib events");
  d.has_edges = a->edges == 1 && a->n_post > 0;
  if (d.has_edges) VSB_REQUIRE(d.edge && d.edge_done, "vsb_enqueue_host_ode: plan needs the edge stream / event");
  d.want = 0; d.steps_done = 0; d.in_flight = 0;
  d.graph[0] = d.graph[1] = nullptr;
  if (n_steps <= 0) return VSB_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (d.main != nullptr && d.main != cudaStreamLegacy) cudaStreamIsCapturing(d.main, &cap);
  VSB_REQUIRE(cap == cudaStreamCaptureStatusNone, "vsb_enqueue_host_ode: cannot be captured into a CUDA graph");
  HostOdeCall* call = new HostOdeCall{*bp, pinned, mdf->host_mail};
  const int saved_seq = mdf->mail_seq;
  const int seq0 = mdf->host_mail->next;
  int rc = VSB_OK;
  cudaError_t e = cudaSuccess;
  int enqueued = 0;
  for (int k = 0; k < n_steps && rc == VSB_OK; ++k) {
    mdf->mail_seq = seq0 + k;                  // nobody polls it here; kept monotonic for a later vsb_run_host_ode
    if ((rc = host_ode_kernels(d, false))) break;
    HostOdeTick* tick = new HostOdeTick{call, mdf->parity & 1, k == n_steps - 1};
    if ((e = cudaLaunchHostFunc(d.ib, host_ode_tick, tick)) != cudaSuccess) {
      delete tick;
      rc = cuda_fail(e, "vsb_enqueue_host_ode (host function)");
      break;
    }
    ++enqueued;
    if ((e = cudaMemcpyAsync(mdf->body, pinned, sizeof(VsbBodyState), cudaMemcpyHostToDevice, d.ib)) != cudaSuccess ||
        (e = cudaEventRecord(d.ib_done, d.ib)) != cudaSuccess || (e = cudaStreamWaitEvent(d.main, d.ib_done, 0)) != cudaSuccess ||
        (d.has_edges && (e = cudaStreamWaitEvent(d.main, d.edge_done, 0)) != cudaSuccess)) {
      rc = cuda_fail(e, "vsb_enqueue_host_ode (state copy / join)");
      break;
    }
    host_ode_flip(d);
  }
  mdf->mail_seq = saved_seq;
  if (rc != VSB_OK) {
    // ticks already enqueued still reference `call`: let them run, then release it (none of them is marked last)
    if (enqueued > 0) cudaStreamSynchronize(d.ib);
    delete call;
  } else {
    mdf->host_mail->next = seq0 + n_steps;
  }
  return rc;
}

int vsb_step_host_ode(VsbStepArgs* a, const VsbMdfArgs* mdf, const VsbBodyParams* bp, VsbBodyState* pinned,
                      const VsbHostPlan* plan) {
  VSB_REQUIRE(a && mdf && bp && pinned && plan, "vsb_step_host_ode: null argument");
  VSB_REQUIRE(mdf->body != nullptr, "vsb_step_host_ode: needs a body state");
  cudaStream_t main = (cudaStream_t)plan->main, ib = (cudaStream_t)plan->ib, edge = (cudaStream_t)plan->edge;
  cudaEvent_t fork = (cudaEvent_t)plan->ev_fork, ib_done = (cudaEvent_t)plan->ev_ib, edge_done = (cudaEvent_t)plan->ev_edge;
  VSB_REQUIRE(ib && fork && ib_done, "vsb_step_host_ode: plan needs the ib stream and the fork / ib events (main may be the default stream)");
  cudaError_t e = cudaEventRecord(fork, main);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ib, fork, 0);
  if (e != cudaSuccess) return cuda_fail(e, "vsb_step_host_ode (fork)");
  const int has_edges = a->edges == 1 && a->n_post > 0 && a->do_stream;   // edges == 2: the walls ride in the bulk launch
  int rc;
  a->band = 1;                                     // the bulk runs while the host advances the body
  if ((rc = vsb_step(a, main))) return rc;
  if (has_edges) {
    VSB_REQUIRE(edge && edge_done, "vsb_step_host_ode: plan needs the edge stream / event");
    if ((e = cudaStreamWaitEvent(edge, fork, 0)) != cudaSuccess) return cuda_fail(e, "vsb_step_host_ode (edge fork)");
    if ((rc = vsb_edge_fused(a, edge))) return rc;
    if ((e = cudaEventRecord(edge_done, edge)) != cudaSuccess) return cuda_fail(e, "vsb_step_host_ode (edge join)");
  }
  if ((rc = vsb_ib_mdf(a, mdf, nullptr, ib))) return rc;                         // force on the markers and the window
  if ((rc = vsb_body_newmark_host(mdf->body, pinned, bp, mdf->parity, ib))) return rc;   // device -> host, ODE, host -> device
  a->band = 2;
  if ((rc = vsb_step(a, ib))) return rc;
  a->band = 0;
  if ((e = cudaEventRecord(ib_done, ib)) != cudaSuccess) return cuda_fail(e, "vsb_step_host_ode (ib join)");
  if ((e = cudaStreamWaitEvent(main, ib_done, 0)) != cudaSuccess) return cuda_fail(e, "vsb_step_host_ode (join)");
  if (has_edges && (e = cudaStreamWaitEvent(main, edge_done, 0)) != cudaSuccess) return cuda_fail(e, "vsb_step_host_ode (join)");
  return VSB_OK;
}

int vsb_body_newmark(VsbBodyState* body, const VsbBodyParams* params, int parity, vsb_stream_t stream) {
  VSB_REQUIRE(body != nullptr && params != nullptr, "vsb_body_newmark: null argument");
  VSB_REQUIRE(params->n_dof >= 1 && params->n_dof <= 3, "n_dof must be 1..3, got %d", params->n_dof);
  VSB_REQUIRE(params->follow >= 0 && params->follow <= 2, "follow must be 0, 1 or 2");
  int dim = 3;
  if (params->grid_size[2] <= 1) dim = 2;
  k_body_newmark<<<1, 32, 0, (cudaStream_t)stream>>>(body, make_body_update(*params, dim), parity & 1);
  VSB_LAUNCH_CHECK("vsb_body_newmark");
  return VSB_OK;
}

}  // extern "C"
