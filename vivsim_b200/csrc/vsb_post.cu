// Field diagnostics (reference vivsim/post.py) and grid-refinement transfers (reference vivsim/multigrid.py):
// SURVEY.md 8f rows 3 and 4 -- the callers either side of the time step.  Not on the hot path; every kernel here is a
// single streaming pass (one read of the input field, one write of the result; neighbour reads are served by L1/L2).
#include "vsb_common.cuh"

namespace vsb {

namespace {

constexpr int kPostBlock = 256;

// jnp.gradient(a, axis) with unit spacing and edge_order 1 (what post.py:29,47-55 relies on):
// interior (a[i+1] - a[i-1]) / 2, edges a[1] - a[0] and a[n-1] - a[n-2].
__device__ __forceinline__ float grad1(const float* __restrict__ a, long long i, int idx, int n, long long stride) {
  if (idx == 0) return __ldg(a + i + stride) - __ldg(a + i);
  if (idx == n - 1) return __ldg(a + i) - __ldg(a + i - stride);
  return (__ldg(a + i + stride) - __ldg(a + i - stride)) * 0.5f;
}

__host__ __device__ constexpr int post_n_out(int kind, int dim) {
  switch (kind) {
    case VSB_DIAG_VORTICITY: return dim == 2 ? 1 : 3;
    case VSB_DIAG_VELOCITY_GRADIENT:
    case VSB_DIAG_STRAIN_RATE: return dim * dim;
    default: return 1;
  }
}

__host__ __device__ constexpr bool post_needs_gradient(int kind) {
  return !(kind == VSB_DIAG_VELOCITY_MAGNITUDE || kind == VSB_DIAG_KINETIC_ENERGY || kind == VSB_DIAG_PRESSURE);
}

// One cell of one diagnostic.  `in` is u (DIM, cells), or rho (cells) for VSB_DIAG_PRESSURE.  The expressions follow
// the reference's operation order (cited per case) so that fp32 results agree to rounding of the final sums only.
template <int DIM, int KIND>
__device__ __forceinline__ void post_cell(const float* __restrict__ in, long long n, long long i, const int* idx,
                                          const int* ext, const long long* stride, float param, float* out) {
  if constexpr (KIND == VSB_DIAG_PRESSURE) {            // post.py:129-139  rho * cs2
    out[0] = __ldg(in + i) * param;
    return;
  }
  float u[DIM];
  if constexpr (!post_needs_gradient(KIND)) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) u[d] = __ldg(in + d * n + i);
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DIM; ++d) s += u[d] * u[d];
    out[0] = KIND == VSB_DIAG_VELOCITY_MAGNITUDE ? sqrtf(s)      // post.py:6-14   norm(u, axis=0)
                                                 : 0.5f * s;     // post.py:106-114  0.5 * sum(u**2)
    return;
  } else {
    float G[DIM][DIM];   // G[i][j] = d u_i / d x_j  (post.py:17-29)
#pragma unroll
    for (int c = 0; c < DIM; ++c)
#pragma unroll
      for (int a = 0; a < DIM; ++a) G[c][a] = grad1(in + c * n, i, idx[a], ext[a], stride[a]);

    if constexpr (KIND == VSB_DIAG_VELOCITY_GRADIENT) {
#pragma unroll
      for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int a = 0; a < DIM; ++a) out[c * DIM + a] = G[c][a];
    } else if constexpr (KIND == VSB_DIAG_STRAIN_RATE) {         // post.py:83-92  0.5 * (G + G^T)
#pragma unroll
      for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int a = 0; a < DIM; ++a) out[c * DIM + a] = 0.5f * (G[c][a] + G[a][c]);
    } else if constexpr (KIND == VSB_DIAG_STRAIN_RATE_MAGNITUDE) {   // post.py:95-103
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
          const float e = 0.5f * (G[c][a] + G[a][c]);
          s += e * e;
        }
      out[0] = sqrtf(s);
    } else if constexpr (KIND == VSB_DIAG_DIVERGENCE) {          // post.py:70-80  sum(gradient(u[i], axis=i))
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DIM; ++c) s += G[c][c];
      out[0] = s;
    } else if constexpr (KIND == VSB_DIAG_Q_CRITERION) {         // post.py:163-177  -0.5 * G_ij G_ji
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int a = 0; a < DIM; ++a) s += G[c][a] * G[a][c];
      out[0] = -0.5f * s;
    } else {                                                     // vorticity family, post.py:32-55
      float w[DIM == 2 ? 1 : 3];
      if constexpr (DIM == 2) {
        w[0] = G[1][0] - G[0][1];                                // dv/dx - du/dy
      } else {
        w[0] = G[2][1] - G[1][2];                                // dwdy - dvdz
        w[1] = G[0][2] - G[2][0];                                // dudz - dwdx
        w[2] = G[1][0] - G[0][1];                                // dvdx - dudy
      }
      constexpr int NW = DIM == 2 ? 1 : 3;
      if constexpr (KIND == VSB_DIAG_VORTICITY) {
#pragma unroll
        for (int k = 0; k < NW; ++k) out[k] = w[k];
      } else {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NW; ++k) s += w[k] * w[k];
        if constexpr (KIND == VSB_DIAG_VORTICITY_MAGNITUDE) out[0] = DIM == 2 ? fabsf(w[0]) : sqrtf(s);  // post.py:58-66
        else out[0] = 0.5f * s;                                  // enstrophy, post.py:142-152
      }
    }
  }
}

template <int DIM>
__device__ __forceinline__ void cell_coords(long long i, int n1, int n2, int* idx) {
  if constexpr (DIM == 2) {
    idx[0] = (int)(i / n2);
    idx[1] = (int)(i % n2);
  } else {
    idx[2] = (int)(i % n2);
    const long long r = i / n2;
    idx[1] = (int)(r % n1);
    idx[0] = (int)(r / n1);
  }
}

template <int DIM, int KIND>
__global__ void __launch_bounds__(kPostBlock) k_post_field(const float* __restrict__ in, float* __restrict__ out,
                                                           int e0, int e1, int e2, float param) {
  const int ext[3] = {e0, e1, e2};
  long long stride[3];
  long long n;
  if constexpr (DIM == 2) { stride[0] = e1; stride[1] = 1; stride[2] = 0; n = (long long)e0 * e1; }
  else { stride[0] = (long long)e1 * e2; stride[1] = e2; stride[2] = 1; n = (long long)e0 * e1 * e2; }
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int idx[3];
  cell_coords<DIM>(i, DIM == 2 ? e0 : e1, DIM == 2 ? e1 : e2, idx);
  constexpr int NO = post_n_out(KIND, DIM);
  float o[NO];
  post_cell<DIM, KIND>(in, n, i, idx, ext, stride, param, o);
#pragma unroll
  for (int k = 0; k < NO; ++k) out[k * n + i] = o[k];
}

// Domain mean of a scalar diagnostic (post.py:117-126 mean_kinetic_energy, :155-160 mean_enstrophy) without writing
// the field: grid-stride accumulation in fp32 per thread (a handful of cells), fp64 across threads and blocks.
template <int DIM, int KIND>
__global__ void __launch_bounds__(kPostBlock) k_post_sum(const float* __restrict__ in, double* __restrict__ acc,
                                                         int e0, int e1, int e2, float param) {
  const int ext[3] = {e0, e1, e2};
  long long stride[3];
  long long n;
  if constexpr (DIM == 2) { stride[0] = e1; stride[1] = 1; stride[2] = 0; n = (long long)e0 * e1; }
  else { stride[0] = (long long)e1 * e2; stride[1] = e2; stride[2] = 1; n = (long long)e0 * e1 * e2; }
  double part = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int idx[3];
    cell_coords<DIM>(i, DIM == 2 ? e0 : e1, DIM == 2 ? e1 : e2, idx);
    float o[1];
    post_cell<DIM, KIND>(in, n, i, idx, ext, stride, param, o);
    part += (double)o[0];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
  __shared__ double warp_part[kPostBlock / 32];
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    part = threadIdx.x < kPostBlock / 32 ? warp_part[threadIdx.x] : 0.0;
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
    if (threadIdx.x == 0) atomicAdd(acc, part);
  }
}

__global__ void k_post_mean_finish(const double* __restrict__ acc, float* __restrict__ out, double inv_n) {
  out[0] = (float)(acc[0] * inv_n);
}

template <int DIM, int KIND>
int launch_field(const VsbGrid& g, const float* in, float param, float* out, cudaStream_t s) {
  const long long n = DIM == 2 ? (long long)g.nx * g.ny : (long long)g.nx * g.ny * g.nz;
  k_post_field<DIM, KIND><<<blocks_for(n, kPostBlock), kPostBlock, 0, s>>>(in, out, g.nx, g.ny, DIM == 2 ? 1 : g.nz, param);
  VSB_LAUNCH_CHECK("vsb_post_field");
  return VSB_OK;
}

template <int DIM, int KIND>
int launch_mean(const VsbGrid& g, const float* in, float param, double* acc, float* out, cudaStream_t s) {
  const long long n = DIM == 2 ? (long long)g.nx * g.ny : (long long)g.nx * g.ny * g.nz;
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double), s);
  if (e != cudaSuccess) return cuda_fail(e, "vsb_post_mean: memset");
  unsigned nb = blocks_for(n, kPostBlock * 8);
  if (nb > 148u * 8u) nb = 148u * 8u;     // a few resident CTAs per SM; the loop strides over the rest
  if (nb < 1) nb = 1;
  k_post_sum<DIM, KIND><<<nb, kPostBlock, 0, s>>>(in, acc, g.nx, g.ny, DIM == 2 ? 1 : g.nz, param);
  VSB_LAUNCH_CHECK("vsb_post_mean");
  k_post_mean_finish<<<1, 1, 0, s>>>(acc, out, 1.0 / (double)n);
  VSB_LAUNCH_CHECK("vsb_post_mean: finish");
  return VSB_OK;
}

template <int DIM>
int dispatch_field(const VsbGrid& g, int kind, const float* in, float param, float* out, cudaStream_t s) {
  switch (kind) {
#define VSB_CASE(K) case K: return launch_field<DIM, K>(g, in, param, out, s);
    VSB_CASE(VSB_DIAG_VELOCITY_MAGNITUDE) VSB_CASE(VSB_DIAG_VELOCITY_GRADIENT) VSB_CASE(VSB_DIAG_VORTICITY)
    VSB_CASE(VSB_DIAG_VORTICITY_MAGNITUDE) VSB_CASE(VSB_DIAG_DIVERGENCE) VSB_CASE(VSB_DIAG_STRAIN_RATE)
    VSB_CASE(VSB_DIAG_STRAIN_RATE_MAGNITUDE) VSB_CASE(VSB_DIAG_KINETIC_ENERGY) VSB_CASE(VSB_DIAG_PRESSURE)
    VSB_CASE(VSB_DIAG_ENSTROPHY) VSB_CASE(VSB_DIAG_Q_CRITERION)
#undef VSB_CASE
  }
  set_error("vsb_post_field: unknown diagnostic %d", kind);
  return VSB_ERR_INVALID;
}

template <int DIM>
int dispatch_mean(const VsbGrid& g, int kind, const float* in, float param, double* acc, float* out, cudaStream_t s) {
  switch (kind) {
#define VSB_CASE(K) case K: return launch_mean<DIM, K>(g, in, param, acc, out, s);
    VSB_CASE(VSB_DIAG_VELOCITY_MAGNITUDE) VSB_CASE(VSB_DIAG_VORTICITY_MAGNITUDE) VSB_CASE(VSB_DIAG_DIVERGENCE)
    VSB_CASE(VSB_DIAG_STRAIN_RATE_MAGNITUDE) VSB_CASE(VSB_DIAG_KINETIC_ENERGY) VSB_CASE(VSB_DIAG_PRESSURE)
    VSB_CASE(VSB_DIAG_ENSTROPHY) VSB_CASE(VSB_DIAG_Q_CRITERION)
#undef VSB_CASE
  }
  set_error("vsb_post_mean: diagnostic %d is not a scalar field", kind);
  return VSB_ERR_INVALID;
}

int check_post_grid(const VsbGrid* g, int kind, const char* who) {
  VSB_REQUIRE(g, "%s: null grid", who);
  VSB_REQUIRE(g->dim == 2 || g->dim == 3, "%s: dim must be 2 or 3, got %d", who, g->dim);
  VSB_REQUIRE(g->nx > 0 && g->ny > 0 && (g->dim == 2 || g->nz > 0), "%s: bad grid", who);
  if (post_needs_gradient(kind))   // numpy/jax gradient needs edge_order + 1 = 2 points per axis
    VSB_REQUIRE(g->nx >= 2 && g->ny >= 2 && (g->dim == 2 || g->nz >= 2),
                "%s: every axis needs at least 2 cells for a gradient", who);
  return VSB_OK;
}

// ------------------------------------------------------------------------------------ grid refinement (D2Q9)
// Index sets of the populations crossing a block edge: reference lbm/lattice.py:58-61 (right, left, up, down).
struct MgLine {
  int q[3];
  long long coarse_base, coarse_stride;   // coarse cell (line position t): coarse_base + t * coarse_stride
  long long fine_base[2], fine_stride;    // fine cells: fine_base[layer] + (2 t + {0, 1}) * fine_stride
  long long n_coarse, n_fine;             // cells per population
  int len;                                // coarse cells along the line
};

// fine_to_coarse (multigrid.py:58-101): coarse edge line <- 0.25 * (sum of the 2 x 2 fine cells it covers), summed in
// the reference's order: (layer0[2t] + layer0[2t+1]) + layer1[2t]) + layer1[2t+1].
__global__ void k_mg_fine_to_coarse(const float* __restrict__ fine, float* __restrict__ coarse, MgLine L) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L.len) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float* p = fine + L.q[k] * L.n_fine;
    const long long a = L.fine_base[0] + 2LL * t * L.fine_stride, b = L.fine_base[1] + 2LL * t * L.fine_stride;
    const float s = ((p[a] + p[a + L.fine_stride]) + p[b]) + p[b + L.fine_stride];
    coarse[L.q[k] * L.n_coarse + L.coarse_base + t * L.coarse_stride] = 0.25f * s;
  }
}

// coarse_to_fine (multigrid.py:103-131): both fine cells along the edge take the coarse value (piecewise constant).
__global__ void k_mg_coarse_to_fine(const float* __restrict__ coarse, float* __restrict__ fine, MgLine L) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // fine cell along the line
  if (t >= 2 * L.len) return;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    fine[L.q[k] * L.n_fine + L.fine_base[0] + t * L.fine_stride] =
        coarse[L.q[k] * L.n_coarse + L.coarse_base + (t >> 1) * L.coarse_stride];
}

// Describe the edge lines of one transfer.  to_coarse selects which side of each block the reference touches.
int mg_line(int dir, bool to_coarse, int nx_f, int ny_f, int nx_c, int ny_c, MgLine& L, const char* who) {
  VSB_REQUIRE(nx_f > 0 && ny_f > 0 && nx_c > 0 && ny_c > 0, "%s: bad shapes", who);
  VSB_REQUIRE(dir >= VSB_MG_LEFT && dir <= VSB_MG_DOWN, "%s: dir must be VSB_MG_LEFT..VSB_MG_DOWN, got %d", who, dir);
  static const int dirs[4][3] = {{3, 7, 6}, {1, 5, 8}, {2, 5, 6}, {4, 7, 8}};   // left, right, up, down
  for (int k = 0; k < 3; ++k) L.q[k] = dirs[dir][k];
  L.n_coarse = (long long)nx_c * ny_c;
  L.n_fine = (long long)nx_f * ny_f;
  const bool along_y = dir == VSB_MG_LEFT || dir == VSB_MG_RIGHT;   // the edge line runs along y
  if (along_y) {
    VSB_REQUIRE(ny_f == 2 * ny_c, "%s: fine ny (%d) must be twice the coarse ny (%d)", who, ny_f, ny_c);
    if (to_coarse) VSB_REQUIRE(nx_f >= 2, "%s: the fine block needs two layers along x", who);
    L.len = ny_c;
    L.coarse_stride = 1;
    L.fine_stride = 1;
    // to_coarse: left  -> coarse[-1]  <- fine[0], fine[1];    right -> coarse[0]  <- fine[-1], fine[-2]
    // to_fine  : left  -> fine[-1]    <- coarse[0];           right -> fine[0]    <- coarse[-1]
    const bool coarse_last = to_coarse ? dir == VSB_MG_LEFT : dir == VSB_MG_RIGHT;
    L.coarse_base = coarse_last ? (long long)(nx_c - 1) * ny_c : 0;
    if (to_coarse) {
      L.fine_base[0] = dir == VSB_MG_LEFT ? 0 : (long long)(nx_f - 1) * ny_f;
      L.fine_base[1] = dir == VSB_MG_LEFT ? ny_f : (long long)(nx_f - 2) * ny_f;
    } else {
      L.fine_base[0] = L.fine_base[1] = dir == VSB_MG_LEFT ? (long long)(nx_f - 1) * ny_f : 0;
    }
  } else {
    VSB_REQUIRE(nx_f == 2 * nx_c, "%s: fine nx (%d) must be twice the coarse nx (%d)", who, nx_f, nx_c);
    if (to_coarse) VSB_REQUIRE(ny_f >= 2, "%s: the fine block needs two layers along y", who);
    L.len = nx_c;
    L.coarse_stride = ny_c;
    L.fine_stride = ny_f;
    // to_coarse: up   -> coarse[:, 0]  <- fine[:, -1], fine[:, -2];   down -> coarse[:, -1] <- fine[:, 0], fine[:, 1]
    // to_fine  : up   -> fine[:, 0]    <- coarse[:, -1];              down -> fine[:, -1]   <- coarse[:, 0]
    const bool coarse_last = to_coarse ? dir == VSB_MG_DOWN : dir == VSB_MG_UP;
    L.coarse_base = coarse_last ? ny_c - 1 : 0;
    if (to_coarse) {
      L.fine_base[0] = dir == VSB_MG_UP ? ny_f - 1 : 0;
      L.fine_base[1] = dir == VSB_MG_UP ? ny_f - 2 : 1;
    } else {
      L.fine_base[0] = L.fine_base[1] = dir == VSB_MG_UP ? 0 : ny_f - 1;
    }
  }
  return VSB_OK;
}

}  // namespace

}  // namespace vsb

using namespace vsb;

extern "C" {

int vsb_post_field(const VsbGrid* grid, int kind, const float* in, float param, float* out, vsb_stream_t stream) {
  int rc = check_post_grid(grid, kind, "vsb_post_field");
  if (rc != VSB_OK) return rc;
  VSB_REQUIRE(in && out && in != out, "vsb_post_field: null or aliased argument");
  return grid->dim == 2 ? dispatch_field<2>(*grid, kind, in, param, out, (cudaStream_t)stream)
                        : dispatch_field<3>(*grid, kind, in, param, out, (cudaStream_t)stream);
}

int vsb_post_mean(const VsbGrid* grid, int kind, const float* in, float param, double* workspace, float* out,
                  vsb_stream_t stream) {
  int rc = check_post_grid(grid, kind, "vsb_post_mean");
  if (rc != VSB_OK) return rc;
  VSB_REQUIRE(in && out && workspace, "vsb_post_mean: null argument");
  return grid->dim == 2 ? dispatch_mean<2>(*grid, kind, in, param, workspace, out, (cudaStream_t)stream)
                        : dispatch_mean<3>(*grid, kind, in, param, workspace, out, (cudaStream_t)stream);
}

int vsb_mg_fine_to_coarse(int nx_f, int ny_f, const float* f_fine, int nx_c, int ny_c, float* f_coarse, int dir,
                          vsb_stream_t stream) {
  VSB_REQUIRE(f_fine && f_coarse, "vsb_mg_fine_to_coarse: null argument");
  MgLine L;
  int rc = mg_line(dir, true, nx_f, ny_f, nx_c, ny_c, L, "vsb_mg_fine_to_coarse");
  if (rc != VSB_OK) return rc;
  k_mg_fine_to_coarse<<<blocks_for(L.len, 128), 128, 0, (cudaStream_t)stream>>>(f_fine, f_coarse, L);
  VSB_LAUNCH_CHECK("vsb_mg_fine_to_coarse");
  return VSB_OK;
}

int vsb_mg_coarse_to_fine(int nx_c, int ny_c, const float* f_coarse, int nx_f, int ny_f, float* f_fine, int dir,
                          vsb_stream_t stream) {
  VSB_REQUIRE(f_fine && f_coarse, "vsb_mg_coarse_to_fine: null argument");
  MgLine L;
  int rc = mg_line(dir, false, nx_f, ny_f, nx_c, ny_c, L, "vsb_mg_coarse_to_fine");
  if (rc != VSB_OK) return rc;
  k_mg_coarse_to_fine<<<blocks_for(2LL * L.len, 128), 128, 0, (cudaStream_t)stream>>>(f_coarse, f_fine, L);
  VSB_LAUNCH_CHECK("vsb_mg_coarse_to_fine");
  return VSB_OK;
}

}  // extern "C"
