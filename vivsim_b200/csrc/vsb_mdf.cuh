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
};

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

// Whole chain of a small 2-D body in one thread-block cluster (vsb_mdf_cluster.cu).
bool mdf_cluster2d_supported(const MdfParams& p);
int launch_mdf_cluster2d(const StepParams<2>& sp, const MdfParams& p, const BodyUpdate& bu, cudaStream_t stream);

}  // namespace vsb
