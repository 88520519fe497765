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

// Per-axis delta weights of one marker: w(point) = prod_d wq[d][offset_d]; offsets -1..2 around floor(x).
template <int DIM> struct AxisWeights {
  float wq[DIM][4];
  int base[DIM];
};

template <int DIM>
__device__ __forceinline__ void axis_weights(int kind, const float* __restrict__ x, AxisWeights<DIM>& aw) {
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    aw.base[d] = (int)floorf(x[d]);
#pragma unroll
    for (int o = 0; o < 4; ++o) aw.wq[d][o] = delta(kind, (float)(aw.base[d] + o - 1) - x[d]);
  }
}

// stencil point s (0 .. 4^DIM - 1, last axis fastest) of a marker: weight, window index, validity
template <int DIM>
__device__ __forceinline__ bool stencil_point(const AxisWeights<DIM>& aw, const int (&wsz)[3], int s, float& w, int& idx,
                                              int (&node)[DIM]) {
  w = 1.f;
  bool inside = true;
#pragma unroll
  for (int d = DIM - 1; d >= 0; --d) {
    const int o = s & 3;
    s >>= 2;
    node[d] = aw.base[d] + o - 1;
    w *= (o == 0) ? aw.wq[d][0] : (o == 1) ? aw.wq[d][1] : (o == 2) ? aw.wq[d][2] : aw.wq[d][3];   // registers, no stack
    inside = inside && node[d] >= 0 && node[d] < wsz[d];
  }
  idx = (DIM == 2) ? node[0] * wsz[1] + node[1] : (node[0] * wsz[1] + node[1]) * wsz[2] + node[2];
  return inside;
}

// G lanes cooperate on one marker (G = 1 .. 32, chosen on the host so that one pass covers as many markers as
// possible); each lane walks NS / G stencil points and the partial sums are combined with log2(G) shuffles.
template <int DIM, int G>
__global__ void __launch_bounds__(1024, 1) k_ib_fused(const StepParams<DIM> p, const IbFusedParams m, const BodyUpdate bu) {
  using L = Lat<DIM>;
  constexpr int NS = (DIM == 2) ? 16 : 64;
  constexpr int PPL = NS / G;
  extern __shared__ float smem[];
  __shared__ float s_force[3];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int groups = nthr / G, grp = tid / G, gl = tid % G;
  int org[3];
  window_origin<DIM>(p, org);
  int wcells = 1;
#pragma unroll
  for (int d = 0; d < DIM; ++d) wcells *= p.wsz[d];
  const int M = m.n_markers;
  float* sg = smem;                      // [DIM][wcells] scratch field
  float* s_pos = sg + DIM * wcells;      // [M][DIM] window-local marker coordinates
  float* s_um = s_pos + M * DIM;         // [M][DIM] marker velocity
  float* s_F = s_um + M * DIM;           // [M][DIM] accumulated marker force
  const int passes = (M + groups - 1) / groups;
  if (tid < 3) s_force[tid] = 0.f;
  float shift[DIM], vbody[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) { shift[d] = m.body ? m.body->d[d] : 0.f; vbody[d] = m.body ? m.body->v[d] : 0.f; }
  for (int i = tid; i < M * DIM; i += nthr) {
    const int d = i % DIM;
    s_pos[i] = (m.markers0[i] + shift[d]) - (float)org[d];   // as the reference's marker_x - ib_x0
    s_F[i] = 0.f;
  }
  __syncthreads();

  for (int stage = 0; stage < m.n_iter; ++stage) {
    // ---- interpolate and update the marker state
    for (int pass = 0; pass < passes; ++pass) {
      const int marker = pass * groups + grp;
      const bool active = marker < M;
      AxisWeights<DIM> aw;
      axis_weights<DIM>(m.delta_kind, s_pos + (active ? marker : 0) * DIM, aw);
      float acc[DIM];
#pragma unroll
      for (int c = 0; c < DIM; ++c) acc[c] = 0.f;
#pragma unroll 4
      for (int j = 0; j < PPL; ++j) {
        float w;
        int idx, node[DIM];
        const bool ok = stencil_point<DIM>(aw, p.wsz, gl * PPL + j, w, idx, node) && active;
        if (!ok) continue;
        if (stage == 0) {
          int cell[3] = {0, 0, 0};
#pragma unroll
          for (int d = 0; d < DIM; ++d) cell[d + L::A0] = org[d] + node[d];
          float f[L::Q], rho, u[DIM];
          pull_cell<DIM>(p, cell[0], cell[1], cell[2], f, true);
          moments<DIM>(f, rho, u);
#pragma unroll
          for (int c = 0; c < DIM; ++c) acc[c] += w * u[c];
        } else {
#pragma unroll
          for (int c = 0; c < DIM; ++c) acc[c] += w * sg[c * wcells + idx];
        }
      }
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
      }
      if (active && gl == 0) {
        const float ds2 = (m.ds_ptr ? __ldg(m.ds_ptr + marker) : m.ds_value) * 2.0f;
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
          const float u_m = (stage == 0) ? acc[c] : s_um[marker * DIM + c] + 0.5f * acc[c];
          const float tgt = m.u_target ? __ldg(m.u_target + marker * DIM + c) : vbody[c];
          s_um[marker * DIM + c] = u_m;
          s_F[marker * DIM + c] += (tgt - u_m) * ds2;
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
      const bool active = marker < M;
      AxisWeights<DIM> aw;
      axis_weights<DIM>(m.delta_kind, s_pos + (active ? marker : 0) * DIM, aw);
      float val[DIM];
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        val[c] = 0.f;
        if (active) {
          if (last) {
            val[c] = s_F[marker * DIM + c];
          } else {   // same arithmetic as above, so this is exactly the dF that was accumulated
            const float ds2 = (m.ds_ptr ? __ldg(m.ds_ptr + marker) : m.ds_value) * 2.0f;
            const float tgt = m.u_target ? __ldg(m.u_target + marker * DIM + c) : vbody[c];
            val[c] = (tgt - s_um[marker * DIM + c]) * ds2;
          }
        }
      }
#pragma unroll 4
      for (int j = 0; j < PPL; ++j) {
        float w;
        int idx, node[DIM];
        if (stencil_point<DIM>(aw, p.wsz, gl * PPL + j, w, idx, node) && active) {
#pragma unroll
          for (int c = 0; c < DIM; ++c) atomicAdd(&sg[c * wcells + idx], val[c] * w);
        }
      }
      if (last && active && gl == 0) {
#pragma unroll
        for (int c = 0; c < DIM; ++c) atomicAdd(&s_force[c], val[c]);
      }
    }
    __syncthreads();
  }
  // ---- outputs: force field on every window cell, marker state
  for (int i = tid; i < DIM * wcells; i += nthr) {   // cell-major, components packed (see WinVec)
    const int c = i / wcells, cell = i - c * wcells;
    m.g_win[cell * WinVec<DIM>::NC + c] = sg[i];
  }
  if (DIM == 3)
    for (int i = tid; i < wcells; i += nthr) m.g_win[i * 4 + 3] = 0.f;
  for (int i = tid; i < M * DIM; i += nthr) { m.marker_force[i] = s_F[i]; m.marker_u[i] = s_um[i]; }
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
  return (cells + 3 * (size_t)a.n_markers) * a.dim * sizeof(float);   // scratch field + marker position / velocity / force
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
  // lanes per marker: as many as fit one pass of 1024 threads (more lanes = shorter serial chains)
  int g = 1;
  const int gmax = (DIM == 2) ? 16 : 32;
  while (g * 2 <= gmax && (long long)m.n_markers * (g * 2) <= 1024) g *= 2;
  auto launch = [&](auto kernel) -> int {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxFusedSmem);
    if (e != cudaSuccess) return cuda_fail(e, "vsb_ib_fused (shared memory attribute)");
    kernel<<<1, 1024, smem, s>>>(p, m, bu);
    VSB_LAUNCH_CHECK("vsb_ib_fused");
    return VSB_OK;
  };
  switch (g) {
    case 1: return launch(k_ib_fused<DIM, (DIM == 2 ? 1 : 2)>);   // 3-D keeps at least 2 lanes (32 points per lane)
    case 2: return launch(k_ib_fused<DIM, 2>);
    case 4: return launch(k_ib_fused<DIM, 4>);
    case 8: return launch(k_ib_fused<DIM, 8>);
    case 16: return launch(k_ib_fused<DIM, 16>);
    default: return launch(k_ib_fused<DIM, 32>);
  }
}

}  // namespace vsb

using namespace vsb;

extern "C" {

int vsb_ib_fused_supported(const VsbMdfArgs* a) {
  if (!a || (a->dim != 2 && a->dim != 3) || a->n_markers <= 0 || a->n_markers > (1 << 24)) return 0;
  if (a->rotation) return 0;   // marker kinematics with rotation live in the multi-CTA chain (vsb_ib_mdf)
  // Shared-memory fp32 atomics retire at ~2 cycles per lane on the one SM this kernel occupies (measured: 512 markers
  // x 16 points x 2 components x 5 iterations = 97 us), so the single-CTA form only pays for small bodies; larger ones
  // use the multi-CTA chain (vsb_ib_mdf), whose global reductions are spread over the whole chip.
  const long long atomics = a->n_markers * (a->dim == 2 ? 16 : 64) * a->dim * a->n_iter;
  if (atomics > 8192) return 0;
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
