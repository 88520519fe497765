// The whole immersed-boundary part of one time step in ONE kernel, for windows that fit shared memory.
//
// A single CTA of 1024 threads owns every marker (a group of 16 / 32 lanes per marker, several passes).  The
// window-sized scratch field of multi-direct forcing lives in shared memory, so the iterations of
// ib/mdf.py:31-64 are separated by __syncthreads() instead of kernel launches:
//
//   stage 0     u at each stencil point = moments of the pulled (streamed, masked) populations   [replaces
//               get_macroscopic + dynamic_slice], interpolate -> u_m ; dF = (U - u_m) 2 ds ; F = dF
//   stage k > 0 u_m += interp(0.5 * scratch) ; dF ; F += dF
//   between     scratch <- 0 ; spread dF (shared-memory atomics)
//   end         scratch <- spread F ; written out to g_win (every window cell: no memset needed) ;
//               body update (Newmark) by thread 0
//
// While this CTA runs on one SM the bulk of the fused step runs on the others (vsb_step band mode 1).
#include "vsb_step.cuh"

namespace vsb {

struct IbFusedParams {
  int delta_kind, n_iter;
  int n_markers;
  const float* markers0;
  const float* u_target;
  const float* ds_ptr;
  float ds_value;
  float* g_win;
  float* marker_u;
  float* marker_force;
  VsbBodyState* body;
  int update_body;
};

template <int DIM> struct Stencil {
  static constexpr int NS = (DIM == 2) ? 16 : 64;
  static constexpr int G = (DIM == 2) ? 16 : 32;
  static constexpr int PPL = NS / G;
  float w[PPL];
  int idx[PPL];
  int node[PPL][DIM];
  bool ok[PPL];
};

template <int DIM>
__device__ __forceinline__ void make_stencil(const IbFusedParams& m, const int (&org)[3], const int (&wsz)[3], int marker, int gl,
                                             bool active, Stencil<DIM>& st) {
  using S = Stencil<DIM>;
  float x[DIM];
  int base[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    float pos = active ? m.markers0[marker * DIM + d] : 2.0f;
    if (m.body) pos += m.body->d[d];
    x[d] = pos - (float)org[d];          // window-local coordinate, as the reference's marker_x - ib_x0
    base[d] = (int)floorf(x[d]);
  }
#pragma unroll
  for (int j = 0; j < S::PPL; ++j) {
    int s = gl * S::PPL + j;
    float wt = 1.f;
    bool inside = active;
#pragma unroll
    for (int d = DIM - 1; d >= 0; --d) {
      st.node[j][d] = base[d] + (s & 3) - 1;
      s >>= 2;
      wt *= delta(m.delta_kind, (float)st.node[j][d] - x[d]);
      inside = inside && st.node[j][d] >= 0 && st.node[j][d] < wsz[d];
    }
    st.w[j] = wt;
    st.ok[j] = inside;
    st.idx[j] = (DIM == 2) ? st.node[j][0] * wsz[1] + st.node[j][1]
                           : (st.node[j][0] * wsz[1] + st.node[j][1]) * wsz[2] + st.node[j][2];
  }
}

template <int DIM>
__global__ void __launch_bounds__(1024, 1) k_ib_fused(const StepParams<DIM> p, const IbFusedParams m, const BodyUpdate bu) {
  using L = Lat<DIM>;
  using S = Stencil<DIM>;
  extern __shared__ float sg[];   // [DIM][wcells] scratch field
  __shared__ float s_force[3];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int groups = nthr / S::G, grp = tid / S::G, gl = tid % S::G;
  int org[3];
  window_origin<DIM>(p, org);
  int wcells = 1;
#pragma unroll
  for (int d = 0; d < DIM; ++d) wcells *= p.wsz[d];
  const int passes = (m.n_markers + groups - 1) / groups;
  if (tid < 3) s_force[tid] = 0.f;

  for (int stage = 0; stage < m.n_iter; ++stage) {
    // ---- interpolate and update the marker state
    for (int pass = 0; pass < passes; ++pass) {
      const int marker = pass * groups + grp;
      const bool active = marker < m.n_markers;
      Stencil<DIM> st;
      make_stencil<DIM>(m, org, p.wsz, marker, gl, active, st);
      float acc[DIM];
#pragma unroll
      for (int c = 0; c < DIM; ++c) acc[c] = 0.f;
#pragma unroll
      for (int j = 0; j < S::PPL; ++j) {
        if (!st.ok[j]) continue;
        if (stage == 0) {
          int cell[3] = {0, 0, 0};
#pragma unroll
          for (int d = 0; d < DIM; ++d) cell[d + L::A0] = org[d] + st.node[j][d];
          float f[L::Q], rho, u[DIM];
          pull_cell<DIM>(p, cell[0], cell[1], cell[2], f, true);
          moments<DIM>(f, rho, u);
#pragma unroll
          for (int c = 0; c < DIM; ++c) acc[c] += st.w[j] * u[c];
        } else {
#pragma unroll
          for (int c = 0; c < DIM; ++c) acc[c] += st.w[j] * sg[c * wcells + st.idx[j]];
        }
      }
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
#pragma unroll
        for (int o = S::G / 2; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
      }
      if (active && gl == 0) {
        const float ds2 = (m.ds_ptr ? m.ds_ptr[marker] : m.ds_value) * 2.0f;
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
          const float u_m = (stage == 0) ? acc[c] : m.marker_u[marker * DIM + c] + 0.5f * acc[c];
          const float tgt = m.u_target ? m.u_target[marker * DIM + c] : (m.body ? m.body->v[c] : 0.f);
          const float dF = (tgt - u_m) * ds2;
          m.marker_u[marker * DIM + c] = u_m;
          m.marker_force[marker * DIM + c] = (stage == 0 ? 0.f : m.marker_force[marker * DIM + c]) + dF;
        }
      }
    }
    __syncthreads();
    // ---- clear the scratch field, then spread dF (or the total force after the last iteration)
    for (int i = tid; i < DIM * wcells; i += nthr) sg[i] = 0.f;
    __syncthreads();
    const bool last = stage == m.n_iter - 1;
    for (int pass = 0; pass < passes; ++pass) {
      const int marker = pass * groups + grp;
      const bool active = marker < m.n_markers;
      Stencil<DIM> st;
      make_stencil<DIM>(m, org, p.wsz, marker, gl, active, st);
      float val[DIM];
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        val[c] = 0.f;
        if (active) {
          if (last) {
            val[c] = m.marker_force[marker * DIM + c];
          } else {   // same arithmetic as above, so the value is identical to the dF that was accumulated
            const float ds2 = (m.ds_ptr ? m.ds_ptr[marker] : m.ds_value) * 2.0f;
            const float tgt = m.u_target ? m.u_target[marker * DIM + c] : (m.body ? m.body->v[c] : 0.f);
            val[c] = (tgt - m.marker_u[marker * DIM + c]) * ds2;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < S::PPL; ++j)
        if (st.ok[j]) {
#pragma unroll
          for (int c = 0; c < DIM; ++c) atomicAdd(&sg[c * wcells + st.idx[j]], val[c] * st.w[j]);
        }
      if (last && active && gl == 0) {
#pragma unroll
        for (int c = 0; c < DIM; ++c) atomicAdd(&s_force[c], val[c]);
      }
    }
    __syncthreads();
  }
  // ---- force field of the step: every window cell
  for (int i = tid; i < DIM * wcells; i += nthr) m.g_win[i] = sg[i];
  if (tid == 0 && m.body) {
#pragma unroll
    for (int c = 0; c < DIM; ++c) m.body->force_sum[c] += s_force[c];
    if (m.update_body) body_update(m.body, bu, p.parity);
  }
}

constexpr size_t kMaxFusedSmem = 200 * 1024;

static size_t fused_smem_bytes(const VsbMdfArgs& a) {
  size_t cells = 1;
  for (int d = 0; d < a.dim; ++d) cells *= (size_t)a.win_size[d];
  return cells * a.dim * sizeof(float);
}

template <int DIM>
static int ib_fused_impl(const VsbStepArgs& sa, const VsbMdfArgs& a, const VsbBodyParams* bp, cudaStream_t s) {
  StepParams<DIM> p;
  VsbStepArgs b = sa;
  if (!b.f_out) b.f_out = a.g_win;   // unused here; only has to differ from f_in
  b.band = 0;
  if (int rc = fill_params<DIM>(b, p)) return rc;
  for (int d = 0; d < DIM; ++d) {
    VSB_REQUIRE(sa.win_size[d] == a.win_size[d], "vsb_ib_fused: step and MDF windows differ");
    p.wsz[d] = a.win_size[d];
    p.worg[d] = a.win_origin0[d];
  }
  p.body = a.body;
  p.parity = a.parity & 1;
  IbFusedParams m;
  m.delta_kind = a.delta_kind; m.n_iter = a.n_iter; m.n_markers = (int)a.n_markers;
  m.markers0 = a.markers0; m.u_target = a.u_target; m.ds_ptr = a.ds_ptr; m.ds_value = a.ds_value;
  m.g_win = a.g_win; m.marker_u = a.marker_u; m.marker_force = a.marker_force; m.body = a.body;
  m.update_body = (a.body && bp && bp->n_dof > 0) ? 1 : 0;
  BodyUpdate bu{};
  if (m.update_body) bu = make_body_update(*bp, DIM);
  const size_t smem = fused_smem_bytes(a);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ib_fused<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxFusedSmem);
    if (e != cudaSuccess) return cuda_fail(e, "vsb_ib_fused (shared memory attribute)");
    attr_set = true;
  }
  k_ib_fused<DIM><<<1, 1024, smem, s>>>(p, m, bu);
  VSB_LAUNCH_CHECK("vsb_ib_fused");
  return VSB_OK;
}

}  // namespace vsb

using namespace vsb;

extern "C" {

int vsb_ib_fused_supported(const VsbMdfArgs* a) {
  if (!a || (a->dim != 2 && a->dim != 3) || a->n_markers <= 0 || a->n_markers > (1 << 24)) return 0;
  return fused_smem_bytes(*a) <= kMaxFusedSmem ? 1 : 0;
}

int vsb_ib_fused(const VsbStepArgs* args, const VsbMdfArgs* a, const VsbBodyParams* params, vsb_stream_t stream) {
  VSB_REQUIRE(args && a, "vsb_ib_fused: null argument");
  VSB_REQUIRE(a->dim == args->grid.dim && (a->dim == 2 || a->dim == 3), "vsb_ib_fused: bad dim");
  VSB_REQUIRE(a->delta_kind >= VSB_DELTA_PESKIN3 && a->delta_kind <= VSB_DELTA_HAT2, "unknown delta kernel %d", a->delta_kind);
  VSB_REQUIRE(a->n_iter >= 1, "n_iter must be >= 1, got %d", a->n_iter);
  VSB_REQUIRE(a->markers0 && a->g_win && a->marker_u && a->marker_force, "vsb_ib_fused: null buffer");
  VSB_REQUIRE(vsb_ib_fused_supported(a), "vsb_ib_fused: the IB window does not fit shared memory; use vsb_ib_mdf");
  return a->dim == 2 ? ib_fused_impl<2>(*args, *a, params, (cudaStream_t)stream)
                     : ib_fused_impl<3>(*args, *a, params, (cudaStream_t)stream);
}

}  // extern "C"
