// Per-function kernels: one C-ABI entry per reference operator.
// These exist for drop-in parity with vivsim.lbm / vivsim.lbm3d; the hot path is vsb_step.cu.
#include <cstdarg>
#include <cstdio>

#include "vsb_common.cuh"

namespace vsb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return VSB_ERR_CUDA;
}

constexpr int kBlock = 256;

// ----------------------------------------------------------------------------- streaming
template <int DIM>
__global__ void k_streaming(const float* __restrict__ f, float* __restrict__ out, int n0, int n1, int n2) {
  using L = Lat<DIM>;
  const long long n = (long long)n0 * n1 * n2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int i2 = (int)(i % n2);
  const int i1 = (int)((i / n2) % n1);
  const int i0 = (int)(i / ((long long)n1 * n2));
#pragma unroll
  for (int q = 0; q < L::Q; ++q) {
    int s0 = i0 - L::c(q, 0), s1 = i1 - L::c(q, 1), s2 = i2 - L::c(q, 2);
    s0 += (s0 < 0) ? n0 : 0; s0 -= (s0 >= n0) ? n0 : 0;
    s1 += (s1 < 0) ? n1 : 0; s1 -= (s1 >= n1) ? n1 : 0;
    s2 += (s2 < 0) ? n2 : 0; s2 -= (s2 >= n2) ? n2 : 0;
    out[q * n + i] = f[q * n + ((long long)s0 * n1 + s1) * n2 + s2];
  }
}

// ----------------------------------------------------------------------------- elementwise ops
template <int DIM>
__global__ void k_macroscopic(const float* __restrict__ f, float* __restrict__ rho, float* __restrict__ u, long long n) {
  using L = Lat<DIM>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float fl[L::Q], r, v[L::D];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) fl[q] = f[q * n + i];
  moments<DIM>(fl, r, v);
  rho[i] = r;
#pragma unroll
  for (int d = 0; d < L::D; ++d) u[d * n + i] = v[d];
}

template <int DIM>
__global__ void k_equilibrium(const float* __restrict__ rho, const float* __restrict__ u, float* __restrict__ feq, long long n) {
  using L = Lat<DIM>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v[L::D], fe[L::Q];
#pragma unroll
  for (int d = 0; d < L::D; ++d) v[d] = u[d * n + i];
  equilibrium<DIM>(rho[i], v, fe);
#pragma unroll
  for (int q = 0; q < L::Q; ++q) feq[q * n + i] = fe[q];
}

template <int DIM, int KIND>
__global__ void k_collision(const float* __restrict__ f, const float* __restrict__ feq, float* __restrict__ out,
                            long long n, Relax r, Matrix<Lat<DIM>::Q> A) {
  using L = Lat<DIM>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float fl[L::Q], fe[L::Q];
#pragma unroll
  for (int q = 0; q < L::Q; ++q) { fl[q] = f[q * n + i]; fe[q] = feq[q * n + i]; }
  if constexpr (KIND == VSB_COLL_BGK) collide_bgk<DIM>(fl, fe, r);
  if constexpr (KIND == VSB_COLL_KBC) collide_kbc<DIM>(fl, fe, r);
  if constexpr (KIND == VSB_COLL_REG) collide_reg<DIM>(fl, fe, r);
  if constexpr (KIND == VSB_COLL_MRT) collide_mrt<DIM>(fl, fe, A);
#pragma unroll
  for (int q = 0; q < L::Q; ++q) out[q * n + i] = fl[q];
}

// MODE 0: out = G; 1: out = f + scale*G; 2: out = f + B G
template <int DIM, int MODE>
__global__ void k_forcing(const float* __restrict__ f, const float* __restrict__ g, const float* __restrict__ u,
                          float* __restrict__ out, long long n, float scale, Matrix<Lat<DIM>::Q> B) {
  using L = Lat<DIM>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gv[L::D], uv[L::D], G[L::Q];
#pragma unroll
  for (int d = 0; d < L::D; ++d) { gv[d] = g[d * n + i]; uv[d] = u[d * n + i]; }
  guo_term<DIM>(gv, uv, G);
  if constexpr (MODE == 0) {
#pragma unroll
    for (int q = 0; q < L::Q; ++q) out[q * n + i] = G[q];
  } else if constexpr (MODE == 1) {
#pragma unroll
    for (int q = 0; q < L::Q; ++q) out[q * n + i] = f[q * n + i] + G[q] * scale;
  } else {
    float fl[L::Q];
#pragma unroll
    for (int q = 0; q < L::Q; ++q) fl[q] = f[q * n + i];
    matvec_add<DIM>(fl, B, G);
#pragma unroll
    for (int q = 0; q < L::Q; ++q) out[q * n + i] = fl[q];
  }
}

template <int Q>
static Matrix<Q> load_matrix(const float* host) {
  Matrix<Q> m;
  for (int i = 0; i < Q * Q; ++i) m.a[i] = host ? host[i] : 0.f;
  return m;
}

template <int DIM>
static int collision_impl(long long n, int kind, double omega, const float* op_host, const float* f, const float* feq,
                          float* out, cudaStream_t s) {
  constexpr int Q = Lat<DIM>::Q;
  const Relax r = make_relax(omega);
  const unsigned nb = blocks_for(n, kBlock);
  switch (kind) {
    case VSB_COLL_BGK: k_collision<DIM, VSB_COLL_BGK><<<nb, kBlock, 0, s>>>(f, feq, out, n, r, Matrix<Q>{}); break;
    case VSB_COLL_KBC: k_collision<DIM, VSB_COLL_KBC><<<nb, kBlock, 0, s>>>(f, feq, out, n, r, Matrix<Q>{}); break;
    case VSB_COLL_REG: k_collision<DIM, VSB_COLL_REG><<<nb, kBlock, 0, s>>>(f, feq, out, n, r, Matrix<Q>{}); break;
    case VSB_COLL_MRT:
      VSB_REQUIRE(op_host != nullptr, "vsb_collision: MRT needs op_host (Q*Q host matrix)");
      k_collision<DIM, VSB_COLL_MRT><<<nb, kBlock, 0, s>>>(f, feq, out, n, r, load_matrix<Q>(op_host));
      break;
    default: VSB_REQUIRE(false, "vsb_collision: unknown collision kind %d", kind);
  }
  VSB_LAUNCH_CHECK("vsb_collision");
  return VSB_OK;
}

template <int DIM>
static int forcing_impl(long long n, int mode, float scale, const float* fop_host, const float* f, const float* g,
                        const float* u, float* out, cudaStream_t s) {
  constexpr int Q = Lat<DIM>::Q;
  const unsigned nb = blocks_for(n, kBlock);
  if (mode == 0) k_forcing<DIM, 0><<<nb, kBlock, 0, s>>>(f, g, u, out, n, scale, Matrix<Q>{});
  else if (mode == 1) k_forcing<DIM, 1><<<nb, kBlock, 0, s>>>(f, g, u, out, n, scale, Matrix<Q>{});
  else k_forcing<DIM, 2><<<nb, kBlock, 0, s>>>(f, g, u, out, n, scale, load_matrix<Q>(fop_host));
  VSB_LAUNCH_CHECK("vsb_forcing");
  return VSB_OK;
}

}  // namespace vsb

using namespace vsb;

#define VSB_DIM_OK(dim) VSB_REQUIRE((dim) == 2 || (dim) == 3, "dim must be 2 (D2Q9) or 3 (D3Q19), got %d", (dim))

extern "C" {

int vsb_abi_version(void) { return 2; }
const char* vsb_last_error(void) { return vsb::g_err; }

int vsb_streaming(const VsbGrid* grid, const float* f, float* out, vsb_stream_t stream) {
  VSB_REQUIRE(grid && f && out, "vsb_streaming: null argument");
  VSB_DIM_OK(grid->dim);
  VSB_REQUIRE(f != out, "vsb_streaming: f and out must not alias");
  cudaStream_t s = (cudaStream_t)stream;
  if (grid->dim == 2) {
    VSB_REQUIRE(grid->nx > 0 && grid->ny > 0, "vsb_streaming: bad grid %d x %d", grid->nx, grid->ny);
    const long long n = (long long)grid->nx * grid->ny;
    k_streaming<2><<<blocks_for(n, kBlock), kBlock, 0, s>>>(f, out, 1, grid->nx, grid->ny);
  } else {
    VSB_REQUIRE(grid->nx > 0 && grid->ny > 0 && grid->nz > 0, "vsb_streaming: bad grid");
    const long long n = (long long)grid->nx * grid->ny * grid->nz;
    k_streaming<3><<<blocks_for(n, kBlock), kBlock, 0, s>>>(f, out, grid->nx, grid->ny, grid->nz);
  }
  VSB_LAUNCH_CHECK("vsb_streaming");
  return VSB_OK;
}

int vsb_macroscopic(int dim, int64_t n, const float* f, float* rho, float* u, vsb_stream_t stream) {
  VSB_DIM_OK(dim);
  if (n == 0) return VSB_OK;
  VSB_REQUIRE(n > 0 && f && rho && u, "vsb_macroscopic: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (dim == 2) k_macroscopic<2><<<blocks_for(n, kBlock), kBlock, 0, s>>>(f, rho, u, n);
  else k_macroscopic<3><<<blocks_for(n, kBlock), kBlock, 0, s>>>(f, rho, u, n);
  VSB_LAUNCH_CHECK("vsb_macroscopic");
  return VSB_OK;
}

int vsb_equilibrium(int dim, int64_t n, const float* rho, const float* u, float* feq, vsb_stream_t stream) {
  VSB_DIM_OK(dim);
  if (n == 0) return VSB_OK;
  VSB_REQUIRE(n > 0 && rho && u && feq, "vsb_equilibrium: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (dim == 2) k_equilibrium<2><<<blocks_for(n, kBlock), kBlock, 0, s>>>(rho, u, feq, n);
  else k_equilibrium<3><<<blocks_for(n, kBlock), kBlock, 0, s>>>(rho, u, feq, n);
  VSB_LAUNCH_CHECK("vsb_equilibrium");
  return VSB_OK;
}

int vsb_collision(int dim, int64_t n, int kind, double omega, const float* op_host, const float* f, const float* feq,
                  float* out, vsb_stream_t stream) {
  VSB_DIM_OK(dim);
  if (n == 0) return VSB_OK;
  VSB_REQUIRE(n > 0 && f && feq && out, "vsb_collision: bad argument");
  return dim == 2 ? collision_impl<2>(n, kind, omega, op_host, f, feq, out, (cudaStream_t)stream)
                  : collision_impl<3>(n, kind, omega, op_host, f, feq, out, (cudaStream_t)stream);
}

int vsb_guo_forcing_term(int dim, int64_t n, const float* g, const float* u, float* out, vsb_stream_t stream) {
  VSB_DIM_OK(dim);
  if (n == 0) return VSB_OK;
  VSB_REQUIRE(n > 0 && g && u && out, "vsb_guo_forcing_term: bad argument");
  return dim == 2 ? forcing_impl<2>(n, 0, 1.f, nullptr, nullptr, g, u, out, (cudaStream_t)stream)
                  : forcing_impl<3>(n, 0, 1.f, nullptr, nullptr, g, u, out, (cudaStream_t)stream);
}

int vsb_forcing(int dim, int64_t n, int kind, double omega, const float* fop_host, const float* f, const float* g,
                const float* u, float* out, vsb_stream_t stream) {
  VSB_DIM_OK(dim);
  VSB_REQUIRE(kind == VSB_FORCE_EDM || kind == VSB_FORCE_GUO, "vsb_forcing: kind must be EDM or GUO, got %d", kind);
  if (n == 0) return VSB_OK;
  VSB_REQUIRE(n > 0 && f && g && u && out, "vsb_forcing: bad argument");
  int mode = 1;
  float scale = 1.f;
  if (kind == VSB_FORCE_GUO) {
    if (fop_host) mode = 2;
    else scale = (float)(1.0 - 0.5 * omega);
  }
  return dim == 2 ? forcing_impl<2>(n, mode, scale, fop_host, f, g, u, out, (cudaStream_t)stream)
                  : forcing_impl<3>(n, mode, scale, fop_host, f, g, u, out, (cudaStream_t)stream);
}

}  // extern "C"
