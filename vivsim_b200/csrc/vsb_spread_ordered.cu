// Deterministic spread (reference ib/stencil.py:81-110): the same scatter-add as vsb_ib_spread, but every
// grid cell receives its contributions in the order of the flattened (marker, stencil point) index -- the order in
// which a sequential scatter (NumPy's add.at, XLA's CPU scatter) applies them -- so the result does not depend on how
// the hardware happens to order fp32 atomics: it is reproducible run to run and bit-identical to the CPU oracle.
//
//   1. key[t] = cell index of entry t = (m, s) (negative indices wrap once, out-of-range entries are dropped, as
//      jnp's scatter does), value[t] = t;
//   2. one STABLE radix sort of the pairs by key (cub::DeviceRadixSort over just the bits a cell index needs): equal
//      keys keep their (m, s) order;
//   3. one thread per run of equal keys walks the run and adds value * weight to the cell, product and sum rounded
//      separately (no fused multiply-add: the reference forms the products first, then adds).
// The caller provides the workspace (nothing is allocated here): vsb_ib_spread_ordered_workspace tells its size.
#include <cub/cub.cuh>

#include "vsb_common.cuh"

namespace vsb {

constexpr int kDropped = 0x7fffffff;

__global__ void k_spread_keys(long long n, long long ncell, const int* __restrict__ idx, int* __restrict__ keys,
                              int* __restrict__ ids) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  long long i = idx[t];
  i += (i < 0) ? ncell : 0;
  keys[t] = (i < 0 || i >= ncell) ? kDropped : (int)i;
  ids[t] = (int)t;
}

__global__ void k_spread_runs(long long n, int ncomp, long long ncell, int ns, float* __restrict__ grid,
                              const float* __restrict__ vals, const float* __restrict__ w, const int* __restrict__ keys,
                              const int* __restrict__ ids) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int key = keys[t];
  if (key == kDropped || (t > 0 && keys[t - 1] == key)) return;      // not the head of a run
  for (int c = 0; c < ncomp; ++c) {
    float acc = grid[c * ncell + key];
    for (long long e = t; e < n && keys[e] == key; ++e) {
      const int id = ids[e];
      acc = __fadd_rn(acc, __fmul_rn(vals[(long long)(id / ns) * ncomp + c], w[id]));
    }
    grid[c * ncell + key] = acc;
  }
}

static size_t sort_temp_bytes(long long n, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr,
                                  (int)n, 0, end_bit);
  return bytes;
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace vsb

using namespace vsb;

extern "C" {

int64_t vsb_ib_spread_ordered_workspace(int64_t n_markers, int n_stencil) {
  const long long n = (long long)n_markers * n_stencil;
  if (n <= 0) return 0;
  if (n >= (1ll << 31)) return -1;
  return (int64_t)(4 * align256((size_t)n * sizeof(int)) + align256(sort_temp_bytes(n, 32)));
}

int vsb_ib_spread_ordered(int n_comp, int64_t n_cells, float* grid, int64_t n_markers, int n_stencil, const float* values,
                          const float* weights, const int32_t* indices, void* workspace, int64_t workspace_bytes,
                          vsb_stream_t stream) {
  VSB_REQUIRE(n_comp >= 1 && n_cells >= 0 && n_markers >= 0 && n_stencil >= 1, "vsb_ib_spread_ordered: bad sizes");
  const long long n = (long long)n_markers * n_stencil;
  if (n == 0 || n_cells == 0) return VSB_OK;
  VSB_REQUIRE(n < (1ll << 31) && n_cells < (1ll << 31) - 1, "vsb_ib_spread_ordered: more than 2^31 entries or cells");
  VSB_REQUIRE(grid && values && weights && indices && workspace, "vsb_ib_spread_ordered: null argument");
  const int64_t need = vsb_ib_spread_ordered_workspace(n_markers, n_stencil);
  VSB_REQUIRE(workspace_bytes >= need, "vsb_ib_spread_ordered: workspace of %lld bytes, %lld needed", (long long)workspace_bytes,
              (long long)need);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t slab = align256((size_t)n * sizeof(int));
  char* base = static_cast<char*>(workspace);
  int* keys_in = reinterpret_cast<int*>(base);
  int* keys_out = reinterpret_cast<int*>(base + slab);
  int* ids_in = reinterpret_cast<int*>(base + 2 * slab);
  int* ids_out = reinterpret_cast<int*>(base + 3 * slab);
  void* temp = base + 4 * slab;
  size_t temp_bytes = (size_t)workspace_bytes - 4 * slab;
  k_spread_keys<<<blocks_for(n, 256), 256, 0, s>>>(n, n_cells, indices, keys_in, ids_in);
  VSB_LAUNCH_CHECK("vsb_ib_spread_ordered (keys)");
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, ids_in, ids_out, (int)n, 0, 32, s);
  if (e != cudaSuccess) return cuda_fail(e, "vsb_ib_spread_ordered (sort)");
  k_spread_runs<<<blocks_for(n, 256), 256, 0, s>>>(n, n_comp, n_cells, n_stencil, grid, values, weights, keys_out, ids_out);
  VSB_LAUNCH_CHECK("vsb_ib_spread_ordered (runs)");
  return VSB_OK;
}

}  // extern "C"
