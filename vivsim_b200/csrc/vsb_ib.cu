// Immersed-boundary kernels: delta kernels, stencil, interpolate,
// spread, and marker-parallel multi-direct forcing with the stencil evaluated on the fly.
// A group of lanes owns one marker; its stencil points are spread over the lanes and reduced
// with warp shuffles; spreading uses fp32 atomics (REDG) into window-sized buffers.
#include "vsb_step.cuh"

namespace vsb {

constexpr int kBlock = 128;

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
    for (int s = lane; s < ns; s += 32) acc += w[m * ns + s] * grid[c * ncell + idx[m * ns + s]];
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
  const int i = idx[t];
  for (int c = 0; c < ncomp; ++c) atomicAdd(&grid[c * ncell + i], vals[m * ncomp + c] * wt);
}

// ----------------------------------------------------------------------------- fused MDF stage
struct MdfParams {
  int delta_kind, n_iter, stage, parity;
  long long n_markers;
  int origin0[3], wsize[3];
  const float* markers0;
  const float* u_target;
  const float* ds_ptr;
  float ds_value;
  const float* u_win;
  float* g_win;
  float* scratch;
  float* marker_u;
  float* marker_force;
  VsbBodyState* body;
};

// Stage k of multi_direct_forcing (ib/mdf.py:31-64):
//   k = 0     : u_m = interp(u)
//   k > 0     : u_m += interp(0.5 * spread(dF_{k-1}))        (buffer scratch[k-1], filled by stage k-1)
//   dF_k = (U - u_m) 2 ds ; F += dF_k
//   k < n-1   : spread dF_k -> scratch[k]     (zeroed by the caller)
//   k = n-1   : spread F -> g_win             (the last iteration's u_m update is never used by the reference's
//               outputs, so its spread + interpolate are skipped)
template <int DIM>
__global__ void k_mdf_stage(MdfParams p) {
  constexpr int NS = (DIM == 2) ? 16 : 64;   // 4^D stencil points
  constexpr int G = (DIM == 2) ? 16 : 32;    // lanes per marker
  constexpr int PPL = NS / G;                // stencil points per lane
  __shared__ float s_force[3];
  if (threadIdx.x < 3) s_force[threadIdx.x] = 0.f;
  __syncthreads();

  const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = gthread / G;
  const int gl = (int)(gthread % G);
  const bool active = m < p.n_markers;
  const long long wcells = (long long)p.wsize[0] * p.wsize[1] * (DIM == 3 ? p.wsize[2] : 1);

  float w[PPL];
  long long idx[PPL];
  bool ok[PPL];
  if (active) {
    float x[DIM];
    int base[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      float pos = p.markers0[m * DIM + d];
      int org = p.origin0[d];
      if (p.body) { pos += p.body->d[d]; org = p.body->origin2[p.parity][d]; }
      x[d] = pos - (float)org;   // window-local coordinate, as in the reference's marker_x - ib_x0
      base[d] = (int)floorf(x[d]);
    }
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
      int s = gl * PPL + j;
      int node[DIM];
      float wt = 1.f;
      bool inside = true;
#pragma unroll
      for (int d = DIM - 1; d >= 0; --d) {
        node[d] = base[d] + (s & 3) - 1;
        s >>= 2;
        wt *= delta(p.delta_kind, (float)node[d] - x[d]);
        inside = inside && node[d] >= 0 && node[d] < p.wsize[d];
      }
      w[j] = wt;
      ok[j] = inside;
      idx[j] = (DIM == 2) ? (long long)node[0] * p.wsize[1] + node[1]
                          : ((long long)node[0] * p.wsize[1] + node[1]) * p.wsize[2] + node[2];
    }
  } else {
#pragma unroll
    for (int j = 0; j < PPL; ++j) { w[j] = 0.f; ok[j] = false; idx[j] = 0; }
  }

  const float* src = (p.stage == 0) ? p.u_win : p.scratch + (long long)(p.stage - 1) * DIM * wcells;
  float um[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < PPL; ++j)
      if (ok[j]) acc += w[j] * src[c * wcells + idx[j]];
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    um[c] = acc;
  }

  float spread_val[DIM];
  if (active) {
    const float ds2 = (p.ds_ptr ? p.ds_ptr[m] : p.ds_value) * 2.0f;
    float u_new[DIM], f_new[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      // every lane of the group reads the previous stage's marker state before lane 0 overwrites it
      const float u_m = (p.stage == 0) ? um[c] : p.marker_u[m * DIM + c] + 0.5f * um[c];
      const float tgt = p.u_target ? p.u_target[m * DIM + c] : (p.body ? p.body->v[c] : 0.f);
      const float dF = (tgt - u_m) * ds2;
      const float F = (p.stage == 0 ? 0.f : p.marker_force[m * DIM + c]) + dF;
      u_new[c] = u_m;
      f_new[c] = F;
      spread_val[c] = (p.stage == p.n_iter - 1) ? F : dF;
    }
    __syncwarp(__activemask());
    if (gl == 0) {
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        p.marker_u[m * DIM + c] = u_new[c];
        p.marker_force[m * DIM + c] = f_new[c];
      }
    }
    float* dst = (p.stage == p.n_iter - 1) ? p.g_win : p.scratch + (long long)p.stage * DIM * wcells;
#pragma unroll
    for (int j = 0; j < PPL; ++j)
      if (ok[j]) {
#pragma unroll
        for (int c = 0; c < DIM; ++c) atomicAdd(&dst[c * wcells + idx[j]], spread_val[c] * w[j]);
      }
    if (p.stage == p.n_iter - 1 && p.body && gl == 0) {
#pragma unroll
      for (int c = 0; c < DIM; ++c) atomicAdd(&s_force[c], spread_val[c]);
    }
  }
  if (p.stage == p.n_iter - 1 && p.body) {
    __syncthreads();
    if (threadIdx.x < DIM) atomicAdd(&p.body->force_sum[threadIdx.x], s_force[threadIdx.x]);
  }
}

// body update by one thread (see body_update in vsb_step.cuh)
__global__ void k_body_newmark(VsbBodyState* b, BodyUpdate u, int parity) {
  if (threadIdx.x == 0 && blockIdx.x == 0) body_update(b, u, parity);
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

int vsb_ib_mdf(const VsbMdfArgs* a, vsb_stream_t stream) {
  VSB_REQUIRE(a != nullptr, "vsb_ib_mdf: null args");
  VSB_REQUIRE(a->dim == 2 || a->dim == 3, "dim must be 2 or 3, got %d", a->dim);
  VSB_REQUIRE(a->delta_kind >= VSB_DELTA_PESKIN3 && a->delta_kind <= VSB_DELTA_HAT2, "unknown delta kernel %d", a->delta_kind);
  VSB_REQUIRE(a->n_iter >= 1, "n_iter must be >= 1, got %d", a->n_iter);
  VSB_REQUIRE(a->n_markers >= 0 && a->markers0 && a->u_win && a->g_win && a->marker_u && a->marker_force,
              "vsb_ib_mdf: null buffer");
  VSB_REQUIRE(a->n_iter == 1 || a->scratch != nullptr, "vsb_ib_mdf: n_iter > 1 needs the scratch buffer");
  for (int d = 0; d < a->dim; ++d) VSB_REQUIRE(a->win_size[d] >= 4, "IB window must be at least 4 cells wide");
  if (a->n_markers == 0) return VSB_OK;
  MdfParams p;
  p.delta_kind = a->delta_kind; p.n_iter = a->n_iter; p.parity = a->parity & 1; p.n_markers = a->n_markers;
  for (int d = 0; d < 3; ++d) { p.origin0[d] = a->win_origin0[d]; p.wsize[d] = a->win_size[d]; }
  p.markers0 = a->markers0; p.u_target = a->u_target; p.ds_ptr = a->ds_ptr; p.ds_value = a->ds_value;
  p.u_win = a->u_win; p.g_win = a->g_win; p.scratch = a->scratch; p.marker_u = a->marker_u;
  p.marker_force = a->marker_force; p.body = a->body;
  const int lanes = (a->dim == 2) ? 16 : 32;
  const unsigned nb = blocks_for(a->n_markers * lanes, kBlock);
  for (int k = 0; k < a->n_iter; ++k) {
    p.stage = k;
    if (a->dim == 2) k_mdf_stage<2><<<nb, kBlock, 0, (cudaStream_t)stream>>>(p);
    else k_mdf_stage<3><<<nb, kBlock, 0, (cudaStream_t)stream>>>(p);
  }
  VSB_LAUNCH_CHECK("vsb_ib_mdf");
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
