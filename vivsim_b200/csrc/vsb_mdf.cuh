// Declarations shared by the multi-direct-forcing kernels (vsb_ib.cu, vsb_mdf_cluster.cu, vsb_ibshard.cu).
#pragma once

#include "vsb_step.cuh"

namespace vsb {

// A set of window fields on several GPUs (sharded IB chain, vsb_ibshard.cu): the same buffer on every rank, peer-mapped.
constexpr int kMaxRanks = 8;

struct MdfParams {
  int delta_kind, n_iter, stage, stage_end, parity;
  unsigned long long* barrier;
  long long n_markers;
  int origin0[3], wsize[3];
  const float* markers0;
  const float* u_target;
  const float* ds_ptr;
  float ds_value;
  const float* u_win;   // optional: precomputed window velocity (stage 0 interpolates it instead of pulling populations)
  float* g_win;         // this step's force field (zero on entry of the last stage)
  float* g_win_next;    // next step's force field: cleared by stage 0
  float* scratch;       // this step's per-iteration fields, (n_iter - 1) x dim x window
  float* scratch_next;  // next step's: buffer k is cleared by stage k
  float* marker_u;
  float* marker_force;
  VsbBodyState* body;
  int update_body;
  VsbHostMail* host_mail;   // host-ODE mode: post the total force to page-locked host memory
  int mail_seq;
  const int* chunk_offsets; // tiled kernel: marker range of every CTA (NULL: 256 consecutive markers each)
  int rotation;             // 2-D: degree of freedom 2 of the body is a rotation about `center`
  float center[2];
  long long m_begin, m_end; // markers handled by this launch (a rank's share when the chain is sharded over GPUs)
  int chunk_begin;          // tiled kernel: first chunk of this launch
  // Which part of the next step's buffers this launch clears: everything (clear_mode 0), or only the window x-planes
  // this rank's copy can receive contributions in (sharded chain): the need box's x-range.
  const unsigned short* nbr_list;   // cluster kernel: per marker, the markers whose stencils can overlap its own
  int nbr_stride;
  int clear_mode;
  int slab_x[2];            // global x-range of this rank's slab
  int need_x[2];            // window-local x-range of this rank's need box
  const int* clear_cells;   // clear_mode 2: only these window cells (ascending flat indices) inside the box below
  long long n_clear_cells;
  int box_lo[3], box_hi[3];
};

// Zero the cells of `field` (NC floats per cell) this launch is responsible for: the flat range of clear_range(), or,
// for a sharded chain, the listed reachable cells inside this rank's need box -- nothing else is ever written there.
template <int NC>
__device__ __forceinline__ void clear_field(const MdfParams& p, float* field, int org_x, int which, long long gthread,
                                            long long nthreads);

// Flat range [begin, end) of window cells to clear in a field: force field (which = 0) or work field (which = 1).
__device__ __forceinline__ void clear_range(const MdfParams& p, int org_x, int which, long long& begin, long long& end) {
  const long long plane = (long long)p.wsize[1] * (p.wsize[2] > 0 ? p.wsize[2] : 1);   // 2-D: wsize[2] unused
  int lo = 0, hi = p.wsize[0];
  if (p.clear_mode) {
    // sharded chain: all fields of a rank's copy are accumulated inside its need box only
    lo = p.need_x[0];
    hi = p.need_x[1];
    lo = lo < 0 ? 0 : lo;
    hi = hi > p.wsize[0] ? p.wsize[0] : hi;
    if (hi < lo) hi = lo;
  }
  begin = lo * plane;
  end = hi * plane;
}

// Sharded chain (vsb_ibshard.cu): where a value spread onto window cell (nx, ny, nz) has to go.  Work-field stages: the
// copy of every rank whose need box contains the cell; last stage: the copy of the rank whose slab contains the
// cell's global x.
struct ShardDev {
  int n_ranks, by_slab;
  float* dst[kMaxRanks];
  int lo[kMaxRanks][3], hi[kMaxRanks][3];   // need boxes, window-local [lo, hi)
  int x_lo[kMaxRanks], x_hi[kMaxRanks];     // slabs, global x [lo, hi)
};
template <bool SHARD> struct ShardArg {};
template <> struct ShardArg<true> : ShardDev {};

template <typename VecF>
__device__ __forceinline__ void shard_add(const ShardDev& sh, int org_x, int nx, int ny, int nz, long long idx, VecF v) {
  for (int r = 0; r < sh.n_ranks; ++r) {
    const bool in = sh.by_slab ? (org_x + nx >= sh.x_lo[r] && org_x + nx < sh.x_hi[r])
                               : (nx >= sh.lo[r][0] && nx < sh.hi[r][0] && ny >= sh.lo[r][1] && ny < sh.hi[r][1] &&
                                  nz >= sh.lo[r][2] && nz < sh.hi[r][2]);
    if (in) atomicAdd(reinterpret_cast<VecF*>(sh.dst[r]) + idx, v);
  }
}

template <int NC>
__device__ __forceinline__ void clear_field(const MdfParams& p, float* field, int org_x, int which, long long gthread,
                                            long long nthreads) {
  if (p.clear_mode == 2) {
    const int w1 = p.wsize[1], w2 = p.wsize[2] > 0 ? p.wsize[2] : 1;
    for (long long i = gthread; i < p.n_clear_cells; i += nthreads) {
      const int cell = __ldg(p.clear_cells + i);
      const int z = cell % w2, y = (cell / w2) % w1, x = cell / (w2 * w1);
      if (x >= p.box_lo[0] && x < p.box_hi[0] && y >= p.box_lo[1] && y < p.box_hi[1] && z >= p.box_lo[2] && z < p.box_hi[2]) {
#pragma unroll
        for (int c = 0; c < NC; ++c) field[(long long)cell * NC + c] = 0.f;
      }
    }
    return;
  }
  long long cb, ce;
  clear_range(p, org_x, which, cb, ce);
  for (long long i = NC * cb + gthread; i < NC * ce; i += nthreads) field[i] = 0.f;
}

// Position, target velocity and lever arm of marker m for the body state in p.body (dyn.py:69-120):
//   translation   pos = markers0 + d                                 (get_markers_coords_2dof, dyn.py:69-81)
//   rotation      pos = center + d[0:2] + R(d[2]) (markers0 - center)  (get_markers_coords_3dof, dyn.py:84-99)
//                 arm = pos - center - d[0:2];  target = (v0 - v2 arm_y, v1 + v2 arm_x)   (dyn.py:102-120)
// An explicit u_target overrides the rigid-body velocity.  `pos` is in grid coordinates (not window-local).
template <int DIM>
__device__ __forceinline__ void marker_kinematics(const MdfParams& p, long long m, float (&pos)[DIM], float (&tgt)[DIM],
                                                  float (&arm)[2]) {
  arm[0] = arm[1] = 0.f;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    pos[d] = p.markers0[m * DIM + d];
    tgt[d] = 0.f;
  }
  if (p.body) {
    if (DIM == 2 && p.rotation) {
      const float th = p.body->d[2];
      const float cs = cosf(th), sn = sinf(th);
      const float xr = pos[0] - p.center[0], yr = pos[1] - p.center[1];
      pos[0] = p.center[0] + p.body->d[0] + xr * cs - yr * sn;
      pos[1] = p.center[1] + p.body->d[1] + xr * sn + yr * cs;
      arm[0] = pos[0] - p.center[0] - p.body->d[0];
      arm[1] = pos[1] - p.center[1] - p.body->d[1];
      tgt[0] = p.body->v[0] - p.body->v[2] * arm[1];
      tgt[1] = p.body->v[1] + p.body->v[2] * arm[0];
    } else {
#pragma unroll
      for (int d = 0; d < DIM; ++d) { pos[d] += p.body->d[d]; tgt[d] = p.body->v[d]; }
    }
  }
  if (p.u_target) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) tgt[d] = p.u_target[m * DIM + d];
  }
}

// Torque of the force F on a marker about the displaced centre (get_torque_to_obj, dyn.py:139-154, sign of +F).
__device__ __forceinline__ float marker_torque(const MdfParams& p, const float (&pos)[2], const float (&F)[2]) {
  const float xr = pos[0] - (p.center[0] + p.body->d[0]);
  const float yr = pos[1] - (p.center[1] + p.body->d[1]);
  return xr * F[1] - yr * F[0];
}

// The last arrival of the last stage: device ODE or force into the host mailbox.  One thread.
__device__ __forceinline__ void finish_body(const MdfParams& p, const BodyUpdate& bu) {
  if (p.update_body) {
    body_update(p.body, bu, p.parity);          // ODE on the device
  } else if (p.host_mail) {                     // ODE on the host: post the force, the host polls for seq
    volatile VsbHostMail* mail = p.host_mail;
    for (int c = 0; c < 3; ++c) mail->force[c] = __ldcg(&p.body->force_sum[c]);
    __threadfence_system();
    mail->seq = p.mail_seq >= 0 ? p.mail_seq : p.body->step + 1;   // < 0: the step being taken (graph replays)
  }
}

// One iteration of this rank's share of a sharded chain (defined in vsb_ib.cu next to the kernels it instantiates).
int launch_mdf_stage_sharded(int dim, const MdfParams& p, const ShardDev& sh, bool tiled, unsigned n_chunks,
                             cudaStream_t stream);

// Whole chain of a small 2-D body in one thread-block cluster (vsb_mdf_cluster.cu).
bool mdf_cluster2d_supported(const MdfParams& p);
int launch_mdf_cluster2d(const StepParams<2>& sp, const MdfParams& p, const BodyUpdate& bu, cudaStream_t stream);

// Whole chain of a small 2-D body in one CTA, work field in shared memory, no floating-point atomics (vsb_mdf_cta.cu).
bool mdf_cta2d_supported(const MdfParams& p);
int launch_mdf_cta2d(const StepParams<2>& sp, const MdfParams& p, const BodyUpdate& bu, cudaStream_t stream);

}  // namespace vsb
