// Halo exchange over peer memory (NVLink 5 / NVSwitch) for the slab decomposition -- no NCCL on the data path.
//
// After a step has produced the post-collision state S' on the physical rows of this rank's slab, ONE kernel
//   1. copies the edge layers of the populations that cross a cut straight into the neighbours' ghost layers
//      (stores to peer-mapped addresses):  c_x = +1 populations of the last physical row -> right neighbour's ghost
//      row 0;  c_x = -1 populations of the first physical row -> left neighbour's ghost row nx_local + 1;
//   2. after a system-scope fence, the last CTA publishes the step number in a flag word of each neighbour;
//   3. the same CTA then waits until both neighbours have published the same step number in this rank's flag words.
// Being the last kernel of the step on the stream, everything of the next step is ordered after (3): the ghost layers
// it reads are complete (RAW), and the neighbours have finished reading the buffer this rank will overwrite one step
// later (WAR, buffers ping-pong).  The step number lives in device memory, so the kernel can be replayed from a CUDA
// graph.  Every rank must run the same number of steps.
#include "vsb_common.cuh"
#include "vsb_internal.h"

namespace vsb {

struct HaloParams {
  long long row_elems;       // elements of one x-layer of one population (NY or NY*NZ)
  long long plane_elems;     // elements of one population ((nx_local + 2) * row_elems)
  int nx_local;
  int n_dirs;
  int right_dirs[5], left_dirs[5];
  const float* state;
  float* left_state;
  float* right_state;
  volatile unsigned* my_flags;     // [0] written by the left neighbour, [1] by the right neighbour
  volatile unsigned* left_flags;   // left neighbour's flag words: this rank writes [1]
  volatile unsigned* right_flags;  // right neighbour's flag words: this rank writes [0]
  unsigned* counter;               // [0] step number, [1] CTA ticket, [2] set to 1 if a wait timed out
};

// mode bit 0: copy + publish; bit 1: wait
__global__ void k_halo_push(const HaloParams p, const int mode) {
  if (mode & 1) {
  // copy: blockIdx.y selects (direction, population); x covers the layer in float4 units when aligned
  const int job = blockIdx.y;
  const bool to_right = job < p.n_dirs;
  const int q = to_right ? p.right_dirs[job] : p.left_dirs[job - p.n_dirs];
  const float* src = p.state + q * p.plane_elems + (long long)(to_right ? p.nx_local : 1) * p.row_elems;
  float* dst = (to_right ? p.right_state : p.left_state) + q * p.plane_elems +
               (long long)(to_right ? 0 : p.nx_local + 1) * p.row_elems;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((p.row_elems & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (long long i = t; i < p.row_elems / 4; i += stride) d4[i] = s4[i];
  } else {
    for (long long i = t; i < p.row_elems; i += stride) dst[i] = src[i];
  }
  }
  // publish (+ wait), by the last CTA to finish copying
  __threadfence_system();
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    s_last = (mode & 1) ? (atomicAdd(&p.counter[1], 1u) == total - 1) : 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    unsigned step = p.counter[0];
    if (mode & 1) {
      __threadfence_system();
      p.counter[1] = 0;
      step += 1;
      p.counter[0] = step;
      p.left_flags[1] = step;    // "your right neighbour has delivered step `step`"
      p.right_flags[0] = step;   // "your left neighbour has delivered step `step`"
      __threadfence_system();
    }
    if (!(mode & 2)) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(p.my_flags[0] - step) < 0 || (int)(p.my_flags[1] - step) < 0) {
      __nanosleep(64);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) { p.counter[2] = 1u; break; }   // 10 s: a neighbour is not running; report, do not hang
    }
    __threadfence_system();
  }
}

}  // namespace vsb

using namespace vsb;

static int halo_launch(const VsbHaloArgs* a, int mode, vsb_stream_t stream) {
  VSB_REQUIRE(a != nullptr, "vsb_halo: null args");
  VSB_REQUIRE(a->grid.dim == 2 || a->grid.dim == 3, "dim must be 2 or 3, got %d", a->grid.dim);
  VSB_REQUIRE(a->state && a->left_state && a->right_state && a->my_flags && a->left_flags && a->right_flags && a->counter,
              "vsb_halo: null pointer");
  VSB_REQUIRE(a->grid.nx >= 6, "vsb_halo: the local extent needs at least 4 physical layers plus 2 ghost layers");
  HaloParams p;
  p.row_elems = (long long)a->grid.ny * (a->grid.dim == 3 ? a->grid.nz : 1);
  p.plane_elems = (long long)a->grid.nx * p.row_elems;
  p.nx_local = a->grid.nx - 2;
  if (a->grid.dim == 2) {
    p.n_dirs = 3;
    const int r[3] = {1, 5, 8}, l[3] = {3, 7, 6};          // c_x = +1 / -1 (lbm/lattice.py:58-59)
    for (int i = 0; i < 3; ++i) { p.right_dirs[i] = r[i]; p.left_dirs[i] = l[i]; }
  } else {
    p.n_dirs = 5;
    const int r[5] = {1, 7, 9, 11, 13}, l[5] = {2, 8, 10, 12, 14};   // lbm3d/lattice.py:43-58
    for (int i = 0; i < 5; ++i) { p.right_dirs[i] = r[i]; p.left_dirs[i] = l[i]; }
  }
  p.state = a->state; p.left_state = a->left_state; p.right_state = a->right_state;
  p.my_flags = a->my_flags; p.left_flags = a->left_flags; p.right_flags = a->right_flags; p.counter = a->counter;
  const int block = 256;
  if (mode == 2) {   // wait only: one thread
    k_halo_push<<<dim3(1, 1), 32, 0, (cudaStream_t)stream>>>(p, mode);
  } else {
    long long per = (p.row_elems / 4 + block - 1) / block;
    const unsigned gx = (unsigned)(per < 1 ? 1 : (per > 64 ? 64 : per));
    k_halo_push<<<dim3(gx, 2 * p.n_dirs), block, 0, (cudaStream_t)stream>>>(p, mode);
  }
  VSB_LAUNCH_CHECK("vsb_halo");
  return VSB_OK;
}

extern "C" {

int vsb_halo_push(const VsbHaloArgs* a, vsb_stream_t stream) { return halo_launch(a, 3, stream); }
int vsb_halo_send(const VsbHaloArgs* a, vsb_stream_t stream) { return halo_launch(a, 1, stream); }
int vsb_halo_wait(const VsbHaloArgs* a, vsb_stream_t stream) { return halo_launch(a, 2, stream); }

int vsb_sync_status(const uint32_t* counter, vsb_stream_t stream) {
  VSB_REQUIRE(counter != nullptr, "vsb_sync_status: null counter");
  unsigned host[3] = {0u, 0u, 0u};
  cudaError_t e = cudaMemcpyAsync(host, counter, sizeof(host), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) return vsb::cuda_fail(e, "vsb_sync_status");
  if (host[2] != 0u) {
    vsb::set_error("a cross-GPU wait timed out after 10 s near step %u: a neighbouring rank never published that step "
                   "(the ranks did not run the same number of steps, or one of them stopped)", host[0]);
    return VSB_ERR_TIMEOUT;
  }
  return VSB_OK;
}

}  // extern "C"
