// Whole multi-direct-forcing chain of a small 2-D body (<= 512 markers: the C2 cylinder) in ONE thread-block cluster
// of 8 CTAs x 1024 threads.  The grid-barrier variant (k_mdf_stage<2>, all iterations in one cooperative launch)
// separates the iterations by barriers through a global counter and keeps the per-iteration work fields in global
// memory: ~3.5 us per iteration, 18-20 us for the five iterations of C2 -- the critical path of a single domain (the
// bulk kernel takes 11.5 us).  Here
//   * the work fields live in DISTRIBUTED SHARED MEMORY: CTA r owns a slab of window rows; spreading is a
//     red.shared::cluster.add.f32 into the owner's slab, interpolation a ld.shared::cluster from it;
//   * iterations are separated by the hardware cluster barrier (barrier.cluster.arrive / wait, ~0.2 us);
//   * three slabs per CTA rotate (spread into k % 3, gather from (k - 1) % 3, clear (k + 1) % 3), so one barrier per
//     iteration suffices: the slab being cleared was last read one barrier ago and is next written one barrier ahead.
// Arithmetic and operation order per marker are those of k_mdf_stage<2> (ib/mdf.py:31-64); the last iteration
// spreads F into the global force window with vector reductions exactly like that kernel, so vsb_step needs no change.
#include <cooperative_groups.h>

#include "vsb_mdf.cuh"

namespace vsb {
namespace cg = cooperative_groups;

constexpr int kClusterCtas = 8;        // portable cluster size
constexpr int kClusterThreads = 1024;  // 64 marker groups of 16 lanes per CTA -> 512 markers per cluster
constexpr size_t kClusterSmemMax = 200 * 1024;

__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
k_mdf_cluster2d(const StepParams<2> sp, const MdfParams p, const BodyUpdate bu, int rows_per_cta) {
  using L = Lat<2>;
  constexpr int G = 16;   // 4 x 4 stencil points, one per lane of the marker's group
  extern __shared__ float2 s_field[];            // 3 slabs of rows_per_cta x wsize[1] cells
  __shared__ float s_force[3];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int w1 = p.wsize[1];
  const int slab_cells = rows_per_cta * w1;
  if (threadIdx.x < 3) s_force[threadIdx.x] = 0.f;
  for (int i = threadIdx.x; i < 3 * slab_cells; i += blockDim.x) s_field[i] = make_float2(0.f, 0.f);

  const int gthread = (int)(rank * blockDim.x + threadIdx.x);
  const int nthreads = kClusterCtas * (int)blockDim.x;
  const long long m = gthread / G;
  const int gl = gthread % G;
  const bool active = m < p.n_markers;
  const long long wcells = (long long)p.wsize[0] * w1;

  int org[3] = {p.origin0[0], p.origin0[1], 0};
  if (p.body) { org[0] = p.body->origin2[p.parity][0]; org[1] = p.body->origin2[p.parity][1]; }

  // this lane's stencil point: the same for every iteration
  float w = 0.f, ds2 = 0.f, tgt[2] = {0.f, 0.f}, u_m[2] = {0.f, 0.f}, F[2] = {0.f, 0.f}, pos[2] = {0.f, 0.f}, arm[2];
  int node[2] = {0, 0};
  bool ok = false;
  if (active) {
    float x[2];
    int base[2];
    marker_kinematics<2>(p, m, pos, tgt, arm);
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      x[d] = pos[d] - (float)org[d];
      base[d] = (int)floorf(x[d]);
    }
    int s = gl;
    w = 1.f;
    ok = true;
#pragma unroll
    for (int d = 1; d >= 0; --d) {
      node[d] = base[d] + (s & 3) - 1;
      s >>= 2;
      w *= delta(p.delta_kind, (float)node[d] - x[d]);
      ok = ok && node[d] >= 0 && node[d] < p.wsize[d];
    }
    ds2 = (p.ds_ptr ? p.ds_ptr[m] : p.ds_value) * 2.0f;
  }
  // owner CTA and slab-local index of this lane's stencil cell
  const unsigned owner = ok ? (unsigned)(node[0] / rows_per_cta) : 0u;
  const int local = ok ? (node[0] - (int)owner * rows_per_cta) * w1 + node[1] : 0;
  float2* remote = cluster.map_shared_rank(s_field, owner);
  const long long gidx = (long long)node[0] * w1 + node[1];

  // stage 0: velocity at the stencil point from the streamed (pulled, masked) populations -- issued before the first
  // cluster barrier so that the loads are in flight while the slabs are being cleared
  float um[2] = {0.f, 0.f};
  if (ok) {
    float f[L::Q], rho, u[2];
    pull_cell<2>(sp, 0, org[0] + node[0], org[1] + node[1], f, true);
    moments<2>(f, rho, u);
    um[0] = w * u[0];
    um[1] = w * u[1];
  }
  // next step's global force window is cleared here, as in k_mdf_stage (no memset on the step path)
  for (long long i = gthread; i < 2 * wcells; i += nthreads) p.g_win_next[i] = 0.f;
  cluster.sync();                                 // every CTA's slabs are zero before anybody spreads into them

  for (int stage = 0; stage < p.n_iter; ++stage) {
    const bool last = stage == p.n_iter - 1;
    if (stage > 0) {       // 0.5 * spread(dF_{k-1}) at the stencil point, from the owner's slab
      um[0] = um[1] = 0.f;
      if (ok) {
        const float2 v = remote[((stage - 1) % 3) * slab_cells + local];
        um[0] = w * v.x;
        um[1] = w * v.y;
      }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) um[c] += __shfl_xor_sync(0xffffffffu, um[c], o);
    }
    // the slab that the NEXT iteration spreads into: last read one barrier ago
    if (!last && stage + 1 >= 3) {
      float2* z = s_field + ((stage + 1) % 3) * slab_cells;
      for (int i = threadIdx.x; i < slab_cells; i += blockDim.x) z[i] = make_float2(0.f, 0.f);
    }
    if (active) {
      float val[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        u_m[c] = (stage == 0) ? um[c] : u_m[c] + 0.5f * um[c];
        const float dF = (tgt[c] - u_m[c]) * ds2;
        F[c] = (stage == 0 ? 0.f : F[c]) + dF;
        val[c] = last ? F[c] : dF;
      }
      if (ok) {
        if (last) {        // the force field goes to global memory for vsb_step, one vector reduction per point
          atomicAdd(reinterpret_cast<float2*>(p.g_win) + gidx, make_float2(val[0] * w, val[1] * w));
        } else {           // distributed shared memory of the owner CTA
          float* dst = reinterpret_cast<float*>(remote + (stage % 3) * slab_cells + local);
          atomicAdd(dst, val[0] * w);
          atomicAdd(dst + 1, val[1] * w);
        }
      }
      if (last && gl == 0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          p.marker_u[m * 2 + c] = u_m[c];
          p.marker_force[m * 2 + c] = F[c];
          if (p.body) atomicAdd(&s_force[c], val[c]);
        }
        if (p.body && p.rotation) atomicAdd(&s_force[2], marker_torque(p, pos, val));
      }
    }
    if (!last) cluster.sync();
  }

  if (p.body) {
    __syncthreads();
    if (threadIdx.x < (p.rotation ? 3 : 2)) atomicAdd(&p.body->force_sum[threadIdx.x], s_force[threadIdx.x]);
    __threadfence();
    cluster.sync();                                // all eight partial sums are in before the update
    if (rank == 0 && threadIdx.x == 0 && (p.update_body || p.host_mail)) {
      __threadfence();
      finish_body(p, bu);
    }
  } else {
    cluster.sync();                                // nobody leaves while its shared memory may still be addressed
  }
}

static inline int cluster_rows(const MdfParams& p) { return (p.wsize[0] + kClusterCtas - 1) / kClusterCtas; }
static inline size_t cluster_smem(const MdfParams& p) { return (size_t)3 * cluster_rows(p) * p.wsize[1] * sizeof(float2); }

bool mdf_cluster2d_supported(const MdfParams& p) {
  return p.n_markers > 0 && p.n_markers <= (long long)kClusterCtas * kClusterThreads / 16 && p.u_win == nullptr &&
         cluster_smem(p) <= kClusterSmemMax;
}

// rows_per_cta = ceil(wsize[0] / 8); smem = 3 slabs (C2: 3 x 14 x 108 x 8 B = 36 KB)
int launch_mdf_cluster2d(const StepParams<2>& sp, const MdfParams& p, const BodyUpdate& bu, cudaStream_t stream) {
  const int rows_per_cta = cluster_rows(p);
  const size_t smem = cluster_smem(p);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_mdf_cluster2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmemMax);
    if (e != cudaSuccess) return cuda_fail(e, "k_mdf_cluster2d (shared-memory opt-in)");
    configured = kClusterSmemMax;
  }
  k_mdf_cluster2d<<<kClusterCtas, kClusterThreads, smem, stream>>>(sp, p, bu, rows_per_cta);
  VSB_LAUNCH_CHECK("k_mdf_cluster2d");
  return VSB_OK;
}

}  // namespace vsb
