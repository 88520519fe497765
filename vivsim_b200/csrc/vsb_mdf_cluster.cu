// Whole multi-direct-forcing chain of a small 2-D body (<= 512 markers: the C2 cylinder) in ONE thread-block cluster
// of 8 CTAs x 512 threads, with the iterations carried out in MARKER SPACE.
//
// ib/mdf.py:31-64 iterates  u_m <- u_m + interp(0.5 spread(dF)),  dF = (U - u_m) 2 ds.  interp o spread is the linear map
//     A[m][m'] = sum over cells c of w_m(c) w_m'(c),
// and because the delta kernels are tensor products, A[m][m'] = ax[m][m'] * ay[m][m'] with one-dimensional overlap sums
// of at most four terms each.  A marker's stencil overlaps only those of the markers within 3 cells of it, and for a
// rigid body that set hardly changes, so the caller supplies it once (mdf->nbr_list).  After ONE gather of the fluid
// velocity (stage 0) every further iteration is a small sparse matrix-vector product between markers -- no trip
// through a window-sized work field, no atomics, no grid-wide barrier -- and the force is spread to the grid ONCE.
//
// History (profiles/r02_summary.md): the grid-barrier variant (k_mdf_stage<2>, all iterations in one cooperative launch)
// pays ~3.5 us per iteration (reductions into a global work field, a barrier through a global counter, a gather from
// L2): 20 us for the five iterations of C2, the critical path of a single 1024^2 domain whose bulk kernel takes 11 us.
// A cluster kernel with the work fields in distributed shared memory took 45 us (shared-memory fp32 atomics are
// compare-and-swap loops and neighbouring markers hit the same cells); dense 64 x 512 rows of A per CTA 32 us and an
// all-pairs overlap scan 22 us (8 SMs are issue-bound on 262 k pair tests).
//
//   CTA r owns markers [64 r, 64 r + 64) for the two grid-facing steps, and every CTA runs the iterations for ALL markers
//   (they are tiny and identical), so the cluster needs ONE hardware barrier:
//   1. own markers: the 16 stencil points (two per thread) -> loads of the streamed (pulled, masked) populations issued;
//   2. while they are in flight: position, stencil base and the 4 + 4 one-dimensional weights of every marker, then
//      the non-zero entries of every row of A from the neighbour list;
//   3. moments of the pulled populations, 8-lane reduction -> u_m of the own markers, written into every CTA's copy
//      through distributed shared memory; cluster barrier;
//   4. n_iter x { dF, F ; u_m += 0.5 A dF } for all 512 markers, one thread per marker, __syncthreads in between;
//   5. own markers: F spread to the global force window with vector reductions (as k_mdf_stage<2> does);
//      CTA 0: total force / torque (block reduction, deterministic) and the body update.
// Same operator as the reference; the floating-point association of the sums differs (as it does between any two
// orders of the atomic spreads), well inside the 1e-5 bound -- tests/test_gpu_step.py compares all chain modes.
#include <cooperative_groups.h>

#include <cstdlib>

#include "vsb_mdf.cuh"

namespace vsb {
namespace cg = cooperative_groups;

constexpr int kClusterCtas = 8;        // portable cluster size
constexpr int kClusterThreads = 512;
constexpr int kClusterMarkers = 512;
constexpr int kRows = kClusterMarkers / kClusterCtas;   // 64 markers per CTA face the grid
constexpr int kMaxStride = 48;         // neighbours per marker (incl. itself) the shared-memory rows can hold

struct ClusterShared {
  // one-dimensional weights, padded with four zeros on either side: the overlap of two stencils whose bases differ by
  // o in [-4, 4] is sum_k w_row[k] * w_col_padded[4 + k + o] -- no branch on o
  float wx[kClusterMarkers][12], wy[kClusterMarkers][12];
  int bx[kClusterMarkers], by[kClusterMarkers];          // first stencil node (floor(x) - 1)
  float u[kClusterMarkers][2], F[kClusterMarkers][2], tgt[kClusterMarkers][2], ds2[kClusterMarkers];
  float dF[2][kClusterMarkers][2];     // double-buffered by iteration parity
  float pos[kClusterMarkers][2];
  float red[kClusterThreads / 32][3];
};
// followed by float a[stride][512] (entries of A in the order of the neighbour list, neighbour-major so that the
// threads of a warp -- consecutive markers -- read consecutive words) and unsigned short nbr[stride][512]

__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
k_mdf_cluster2d(const StepParams<2> sp, const MdfParams p, const BodyUpdate bu, const unsigned short* __restrict__ nbr,
                const int stride, const int debug_stop) {
  using L = Lat<2>;
  extern __shared__ __align__(16) unsigned char s_raw[];
  ClusterShared& sm = *reinterpret_cast<ClusterShared*>(s_raw);
  float* s_a = reinterpret_cast<float*>(s_raw + sizeof(ClusterShared));
  unsigned short* s_nbr = reinterpret_cast<unsigned short*>(s_a + (size_t)kClusterMarkers * stride);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int m0 = rank * kRows;
  const int n_mark = (int)p.n_markers;

  // a fused step enqueued behind this launch with early_launch = 1 may start as soon as every CTA here is resident
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (debug_stop == 5) return;                      // timing aid: the launch alone
  int org[3] = {p.origin0[0], p.origin0[1], 0};
  if (p.body) { org[0] = p.body->origin2[p.parity][0]; org[1] = p.body->origin2[p.parity][1]; }

  // ---- 1. own markers: 8 lanes per marker, stencil points gl and gl + 8; issue the population loads
  const int row = tid >> 3, gl = tid & 7;
  const int m_own = m0 + row;
  const bool own = m_own < n_mark;
  float w_pt[2] = {0.f, 0.f};
  int node[2][2] = {{0, 0}, {0, 0}};
  float f[2][L::Q];
  if (own) {
    float pos[2], tgt[2], arm[2];
    marker_kinematics<2>(p, m_own, pos, tgt, arm);
    const float x = pos[0] - (float)org[0], y = pos[1] - (float)org[1];   // window-local, as marker_x - ib_x0
    const int bx = (int)floorf(x) - 1, by = (int)floorf(y) - 1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pt = gl + 8 * h, jx = pt >> 2, jy = pt & 3;
      node[h][0] = bx + jx; node[h][1] = by + jy;
      const bool inside = node[h][0] >= 0 && node[h][0] < p.wsize[0] && node[h][1] >= 0 && node[h][1] < p.wsize[1];
      // the product order of k_mdf_stage; nodes outside the window are skipped by the reference
      w_pt[h] = inside ? delta(p.delta_kind, (float)node[h][1] - y) * delta(p.delta_kind, (float)node[h][0] - x) : 0.f;
      if (w_pt[h] != 0.f) pull_cell<2>(sp, 0, org[0] + node[h][0], org[1] + node[h][1], f[h], true);
    }
  }
  if (debug_stop == 6) return;                      // timing aid: launch + own-marker set-up + population loads
  // neighbour list (stride x 512, neighbour-major) -> shared memory, coalesced
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(nbr);
    unsigned* dst = reinterpret_cast<unsigned*>(s_nbr);
    for (int i = tid; i < stride * (kClusterMarkers / 2); i += kClusterThreads) dst[i] = __ldg(src + i);
  }
  // next step's global force window is cleared here, as in k_mdf_stage (no memset on the step path)
  {
    const long long wcells = (long long)p.wsize[0] * p.wsize[1];
    for (long long i = (long long)rank * kClusterThreads + tid; i < 2 * wcells; i += kClusterCtas * kClusterThreads)
      p.g_win_next[i] = 0.f;
  }

  // ---- 2. every marker: stencil base, one-dimensional weights, target velocity (thread = marker)
  {
    const int m = tid;
    float wx[4] = {0.f, 0.f, 0.f, 0.f}, wy[4] = {0.f, 0.f, 0.f, 0.f};
    int bx = 1 << 20, by = 1 << 20;                  // padding rows overlap nobody
    float tgt[2] = {0.f, 0.f}, pos[2] = {0.f, 0.f}, ds2 = 0.f;
    if (m < n_mark) {
      float arm[2];
      marker_kinematics<2>(p, m, pos, tgt, arm);
      const float x = pos[0] - (float)org[0], y = pos[1] - (float)org[1];
      bx = (int)floorf(x) - 1;
      by = (int)floorf(y) - 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int nx = bx + k, ny = by + k;
        wx[k] = (nx >= 0 && nx < p.wsize[0]) ? delta(p.delta_kind, (float)nx - x) : 0.f;
        wy[k] = (ny >= 0 && ny < p.wsize[1]) ? delta(p.delta_kind, (float)ny - y) : 0.f;
      }
      ds2 = (p.ds_ptr ? p.ds_ptr[m] : p.ds_value) * 2.0f;
    }
    sm.bx[m] = bx; sm.by[m] = by;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sm.wx[m][k] = 0.f; sm.wx[m][8 + k] = 0.f; sm.wx[m][4 + k] = wx[k];
      sm.wy[m][k] = 0.f; sm.wy[m][8 + k] = 0.f; sm.wy[m][4 + k] = wy[k];
    }
    sm.tgt[m][0] = tgt[0]; sm.tgt[m][1] = tgt[1];
    sm.pos[m][0] = pos[0]; sm.pos[m][1] = pos[1];
    sm.ds2[m] = ds2;
    sm.F[m][0] = 0.f; sm.F[m][1] = 0.f;
  }
  __syncthreads();
  if (debug_stop == 1) return;

  // rows of A = interp o spread over the neighbour list: A[m][c] = ax * ay, one-dimensional overlaps of the two stencils
  {
    const int m = tid;
    const int rbx = sm.bx[m], rby = sm.by[m];
    float rwx[4], rwy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { rwx[k] = sm.wx[m][4 + k]; rwy[k] = sm.wy[m][4 + k]; }
#pragma unroll 2
    for (int j = 0; j < stride; ++j) {
      int c = s_nbr[j * kClusterMarkers + m];
      c = c < kClusterMarkers ? c : kClusterMarkers - 1;         // padding entry: any column, its offset is out of reach
      const int ox = max(-4, min(4, rbx - sm.bx[c])), oy = max(-4, min(4, rby - sm.by[c]));
      const float* cwx = &sm.wx[c][4 + ox];                      // node k of the row = node k + offset of the column
      const float* cwy = &sm.wy[c][4 + oy];
      float ax = 0.f, ay = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) { ax += rwx[k] * cwx[k]; ay += rwy[k] * cwy[k]; }
      s_a[j * kClusterMarkers + m] = (s_nbr[j * kClusterMarkers + m] < kClusterMarkers) ? ax * ay : 0.f;
    }
  }
  if (debug_stop == 2) return;

  // ---- 3. u_m of the own markers from the pulled populations; all-gather through distributed shared memory
  {
    float um[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (w_pt[h] != 0.f) {
        float rho, u[2];
        moments<2>(f[h], rho, u);
        um[0] += w_pt[h] * u[0];
        um[1] += w_pt[h] * u[1];
      }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) um[c] += __shfl_xor_sync(0xffffffffu, um[c], o);
    }
    // lane gl of the marker's group writes the pair (u_x, u_y) into CTA gl's copy
    float2* remote = reinterpret_cast<float2*>(cluster.map_shared_rank(&sm.u[m_own][0], gl));
    *remote = make_float2(um[0], um[1]);
  }
  cluster.sync();                                    // the only cluster barrier: every copy of u is complete
  if (debug_stop == 3) return;

  // ---- 4. iterations in marker space, all markers in every CTA (thread = marker)
  {
    const int m = tid;
    float u0 = sm.u[m][0], u1 = sm.u[m][1], F0 = 0.f, F1 = 0.f;
    const float t0 = sm.tgt[m][0], t1 = sm.tgt[m][1], ds2 = sm.ds2[m];

    for (int stage = 0; stage < p.n_iter; ++stage) {
      const float d0 = (t0 - u0) * ds2, d1 = (t1 - u1) * ds2;
      F0 += d0; F1 += d1;
      if (stage == p.n_iter - 1) break;
      const int buf = stage & 1;
      sm.dF[buf][m][0] = d0; sm.dF[buf][m][1] = d1;
      __syncthreads();
      float a0 = 0.f, a1 = 0.f;
      // four neighbours per trip with independent loads (padding entries have a = 0 and point at column 0)
      for (int j = 0; j < stride; j += 4) {
        float av[4];
        int cv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          av[k] = s_a[(j + k) * kClusterMarkers + m];
          cv[k] = s_nbr[(j + k) * kClusterMarkers + m];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = cv[k] < kClusterMarkers ? cv[k] : 0;
          const float2 d = *reinterpret_cast<const float2*>(sm.dF[buf][c]);
          a0 += av[k] * d.x;
          a1 += av[k] * d.y;
        }
      }
      u0 += 0.5f * a0; u1 += 0.5f * a1;
    }
    sm.u[m][0] = u0; sm.u[m][1] = u1;
    sm.F[m][0] = F0; sm.F[m][1] = F1;
  }
  __syncthreads();
  if (debug_stop == 4) return;

  // ---- 5. own markers: spread F to the global force window, outputs
  if (own) {
    const float F0 = sm.F[m_own][0], F1 = sm.F[m_own][1];
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (w_pt[h] != 0.f)
        atomicAdd(reinterpret_cast<float2*>(p.g_win) + ((long long)node[h][0] * p.wsize[1] + node[h][1]),
                  make_float2(F0 * w_pt[h], F1 * w_pt[h]));
    if (gl == 0) {
      p.marker_u[m_own * 2 + 0] = sm.u[m_own][0];
      p.marker_u[m_own * 2 + 1] = sm.u[m_own][1];
      p.marker_force[m_own * 2 + 0] = F0;
      p.marker_force[m_own * 2 + 1] = F1;
    }
  }
  // CTA 0: total force (and torque) over all markers, deterministic block reduction; body update
  if (rank == 0 && p.body) {
    const int m = tid;
    float s[3] = {0.f, 0.f, 0.f};
    if (m < n_mark) {
      s[0] = sm.F[m][0]; s[1] = sm.F[m][1];
      if (p.rotation) {
        const float pos[2] = {sm.pos[m][0], sm.pos[m][1]}, Fv[2] = {s[0], s[1]};
        s[2] = marker_torque(p, pos, Fv);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
      if ((tid & 31) == 0) sm.red[tid >> 5][c] = s[c];
    }
    __syncthreads();
    if (tid == 0) {
      float tot[3] = {0.f, 0.f, 0.f};
      for (int wdx = 0; wdx < kClusterThreads / 32; ++wdx)
        for (int c = 0; c < 3; ++c) tot[c] += sm.red[wdx][c];
      for (int c = 0; c < (p.rotation ? 3 : 2); ++c) p.body->force_sum[c] += tot[c];
      if (p.update_body || p.host_mail) finish_body(p, bu);
    }
  }
}

// The caller supplies, for every marker, the markers whose stencils can overlap its own (itself included), padded
// with 0xffff to `stride` <= 48 entries per marker (a multiple of 4), NEIGHBOUR-MAJOR: nbr[j][m], 512 columns.
bool mdf_cluster2d_supported(const MdfParams& p) {
  return p.n_markers > 0 && p.n_markers <= kClusterMarkers && p.u_win == nullptr && p.nbr_list != nullptr &&
         p.nbr_stride >= 4 && p.nbr_stride <= kMaxStride && p.nbr_stride % 4 == 0;
}

int launch_mdf_cluster2d(const StepParams<2>& sp, const MdfParams& p, const BodyUpdate& bu, cudaStream_t stream) {
  const size_t smem = sizeof(ClusterShared) + (size_t)kClusterMarkers * p.nbr_stride * (sizeof(float) + sizeof(unsigned short));
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_mdf_cluster2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "k_mdf_cluster2d (shared-memory opt-in)");
    configured = smem;
  }
  static const int debug_stop = getenv("VSB_CLUSTER_STOP") ? atoi(getenv("VSB_CLUSTER_STOP")) : 0;   // timing aid
  k_mdf_cluster2d<<<kClusterCtas, kClusterThreads, smem, stream>>>(sp, p, bu, p.nbr_list, p.nbr_stride, debug_stop);
  VSB_LAUNCH_CHECK("k_mdf_cluster2d");
  return VSB_OK;
}

}  // namespace vsb
