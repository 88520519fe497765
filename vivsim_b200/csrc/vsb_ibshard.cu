// Immersed-boundary chain shared by all ranks of a slab-decomposed run, over peer-mapped symmetric memory
// (include/vivsim_b200.h, "multi-GPU: immersed-boundary chain shared by all ranks").  Replaces "the rank whose slab
// contains the body computes everything": the markers are divided among ALL ranks (the reference's design intent --
// markers owner-computed, a small all-reduce of the total force and torque -- with the "owner" of a marker chosen for
// load balance, not by slab), a body may cross slab cuts, and the chain of a compact body (the 695 k-marker cylinder
// of BASELINE config 5) no longer serialises on one GPU.
//
// Per step, on one stream (all device code, graph-capturable):
//   k_shard_window_moments   velocity of the streamed state on the window cells of MY slab, stored into the copy of every
//                            rank whose need box contains the cell
//   k_shard_barrier          all-to-all flag barrier
//   n_iter x { k_mdf_stage / k_mdf_stage_tiled <SHARD = true> on my markers (interpolate from my copy, spread into every
//              copy that needs the cell; last iteration: into my own copy only) ; k_shard_barrier }
//   k_shard_push_force       before the last barrier: the cells of my need box -> staging slot [me] of the slab owner
//   the last barrier also exchanges the partial force / torque sums (slot [src] on every rank), adds them in rank
//   order and advances this rank's replica of the rigid body -- identical arithmetic on every rank.
//   k_shard_reduce_force     staging slots -> the force field my fluid kernels read
// Only the window cells a marker stencil can reach (a static list) are computed, sent and added: for the cylinder of
// BASELINE config 5 that is 11 % of the 118 x 118 x 500 window (measured on 8 GPUs: window velocity 0.50 -> ms, ...).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "vsb_mdf.cuh"

namespace vsb {

struct BarrierParams {
  int n_ranks, rank;
  unsigned* flags[kMaxRanks];
  float* sums[kMaxRanks];
  unsigned* counter;      // [0] barrier number, [2] time-out indicator
  VsbBodyState* body;
  int finish;             // 1: exchange the partial sums before the barrier, total + body update after it
  int n_sum;              // components of the sum (dim, +1 with rotation)
  unsigned long long* trace;
};

__global__ void k_shard_barrier(const BarrierParams b, const MdfParams p, const BodyUpdate bu) {
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = b.counter[0] + 1u;
    b.counter[0] = s_epoch;
    if (b.trace) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      b.trace[2 * (s_epoch & 4095u)] = t;
    }
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  const int r = threadIdx.x;
  if (r < b.n_ranks) {
    if (b.finish && b.body) {   // my partial sums (accumulated by the last iteration's kernel) -> slot [rank] of rank r
      const float4 part = make_float4(__ldcg(&b.body->force_sum[0]), __ldcg(&b.body->force_sum[1]),
                                      __ldcg(&b.body->force_sum[2]), 0.f);
      __stcg(reinterpret_cast<float4*>(b.sums[r] + 4 * b.rank), part);
    }
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(b.flags[r] + b.rank) = epoch;     // "rank `rank` has reached barrier `epoch`"
    const volatile unsigned* mine = b.flags[b.rank] + r;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(*mine - epoch) < 0) {
      __nanosleep(40);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) { b.counter[2] = 1u; break; }   // 10 s: a rank is not running; report, do not hang
    }
    __threadfence_system();
  }
  __syncthreads();
  if (b.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    b.trace[2 * (epoch & 4095u) + 1] = t;
  }
  if (b.finish && b.body && threadIdx.x == 0) {
    float tot[3] = {0.f, 0.f, 0.f};
    for (int s = 0; s < b.n_ranks; ++s) {            // rank order: every rank forms the same sum bit for bit
      const volatile float* slot = b.sums[b.rank] + 4 * s;
      for (int c = 0; c < b.n_sum; ++c) tot[c] += slot[c];
    }
    for (int c = 0; c < 3; ++c) b.body->force_sum[c] = tot[c];
    if (p.update_body) body_update(b.body, bu, p.parity);
    else b.body->ticket = 0;                          // fixed kinematics: the sum stays in force_sum for the caller
  }
}

struct WindowPush {
  int n_ranks;
  float* dst[kMaxRanks];
  int lo[kMaxRanks][3], hi[kMaxRanks][3];
};

// Window-local coordinates of a flat cell index.
template <int DIM>
__device__ __forceinline__ void cell_coords(int cell, const int (&wsz)[3], int (&rel)[3]) {
  rel[2] = 0;
  if constexpr (DIM == 3) { rel[2] = cell % wsz[2]; cell /= wsz[2]; }
  rel[1] = cell % wsz[1];
  rel[0] = cell / wsz[1];
}

__device__ __forceinline__ bool in_box(const int (&rel)[3], const int (&lo)[3], const int (&hi)[3], int dim) {
  return rel[0] >= lo[0] && rel[0] < hi[0] && rel[1] >= lo[1] && rel[1] < hi[1] && (dim == 2 || (rel[2] >= lo[2] && rel[2] < hi[2]));
}

// u(stream(f_in)) on the listed window cells whose x lies in this rank's rows, multicast by need box.
template <int DIM>
__global__ void __launch_bounds__(128) k_shard_window_moments(const StepParams<DIM> p, const WindowPush w,
                                                              const int* __restrict__ cells, const long long n_cells) {
  using L = Lat<DIM>;
  using VecF = typename std::conditional<DIM == 2, float2, float4>::type;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cells) return;
  int org[3];
  window_origin<DIM>(p, org);                       // local coordinates (body origin + win_shift)
  const int cell = __ldg(cells + t);
  int rel[3];
  cell_coords<DIM>(cell, p.wsz, rel);
  const int x = org[0] + rel[0];
  if (x < p.r_begin || x >= p.r_end) return;         // another rank's rows
  int c[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < L::D; ++d) c[d + L::A0] = org[d] + rel[d];
  float f[L::Q], rho, u[L::D];
  pull_cell<DIM>(p, c[0], c[1], c[2], f, true);
  moments<DIM>(f, rho, u);
  VecF v;
  v.x = u[0]; v.y = u[1];
  if constexpr (DIM == 3) { v.z = u[2]; v.w = 0.f; }
  for (int k = 0; k < w.n_ranks; ++k)
    if (in_box(rel, w.lo[k], w.hi[k], DIM)) reinterpret_cast<VecF*>(w.dst[k])[cell] = v;
  // no fence here: the flag barrier that follows on the stream fences at system scope before it publishes, and fences
  // are cumulative over everything ordered before them (the end of this kernel)
}

// The force field accumulated by my markers (my copy) -> staging slot [me] of the rank whose slab contains the cell.
struct ForcePush {
  int n_ranks, rank;
  const float* src;               // my accumulation field
  float* staging[kMaxRanks];      // base of the staging block of every rank
  int x_lo[kMaxRanks], x_hi[kMaxRanks];
  int lo[3], hi[3];               // my need box
  int org_shift;                  // global x of window x-plane 0 = body origin (global); fixed bodies: origin0
  long long field;                // floats per window field
};

template <int DIM>
__global__ void __launch_bounds__(256) k_shard_push_force(const ForcePush f, const int* __restrict__ cells, const long long n_cells,
                                                          const VsbBodyState* body, const int parity, const int wsz0,
                                                          const int wsz1, const int wsz2) {
  using VecF = typename std::conditional<DIM == 2, float2, float4>::type;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cells) return;
  const int cell = __ldg(cells + t);
  const int wsz[3] = {wsz0, wsz1, wsz2};
  int rel[3];
  cell_coords<DIM>(cell, wsz, rel);
  if (!in_box(rel, f.lo, f.hi, DIM)) return;
  const int gx = (body ? body->origin2[parity][0] : f.org_shift) + rel[0];
  const VecF v = __ldcg(reinterpret_cast<const VecF*>(f.src) + cell);
  for (int r = 0; r < f.n_ranks; ++r)
    if (gx >= f.x_lo[r] && gx < f.x_hi[r])
      reinterpret_cast<VecF*>(f.staging[r] + (long long)f.rank * f.field)[cell] = v;
}

// Staging slots -> the force field my fluid kernels read: the cells of my slab, summed over the ranks whose need box
// contains them, in rank order.
struct ForceReduce {
  int n_ranks, rank;
  const float* staging;           // my staging block
  float* out;
  int lo[kMaxRanks][3], hi[kMaxRanks][3];
  int x_lo, x_hi;
  int org_shift;
  long long field;
};

template <int DIM>
__global__ void __launch_bounds__(256) k_shard_reduce_force(const ForceReduce f, const int* __restrict__ cells, const long long n_cells,
                                                            const VsbBodyState* body, const int parity, const int wsz0,
                                                            const int wsz1, const int wsz2) {
  using VecF = typename std::conditional<DIM == 2, float2, float4>::type;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cells) return;
  const int cell = __ldg(cells + t);
  const int wsz[3] = {wsz0, wsz1, wsz2};
  int rel[3];
  cell_coords<DIM>(cell, wsz, rel);
  const int gx = (body ? body->origin2[parity][0] : f.org_shift) + rel[0];
  if (gx < f.x_lo || gx >= f.x_hi) return;
  VecF acc;
  acc.x = 0.f; acc.y = 0.f;
  if constexpr (DIM == 3) { acc.z = 0.f; acc.w = 0.f; }
  for (int r = 0; r < f.n_ranks; ++r)
    if (in_box(rel, f.lo[r], f.hi[r], DIM)) {
      const VecF v = __ldcg(reinterpret_cast<const VecF*>(f.staging + (long long)r * f.field) + cell);
      acc.x += v.x; acc.y += v.y;
      if constexpr (DIM == 3) acc.z += v.z;
    }
  reinterpret_cast<VecF*>(f.out)[cell] = acc;
}

static int launch_barrier(const VsbIbShard& sh, const MdfParams& p, const BodyUpdate& bu, int finish, int n_sum,
                          cudaStream_t stream) {
  BarrierParams b{};
  b.n_ranks = sh.n_ranks; b.rank = sh.rank;
  for (int r = 0; r < sh.n_ranks; ++r) { b.flags[r] = sh.flags[r]; b.sums[r] = sh.sums[r]; }
  b.counter = sh.counter; b.body = p.body; b.finish = finish; b.n_sum = n_sum;
  b.trace = reinterpret_cast<unsigned long long*>(sh.trace);
  k_shard_barrier<<<1, 32, 0, stream>>>(b, p, bu);
  VSB_LAUNCH_CHECK("vsb_ibshard (barrier)");
  return VSB_OK;
}

static int check_shard(const VsbIbShard* sh, const char* who) {
  VSB_REQUIRE(sh != nullptr, "%s: null shard", who);
  VSB_REQUIRE(sh->n_ranks >= 1 && sh->n_ranks <= VSB_MAX_RANKS && sh->rank >= 0 && sh->rank < sh->n_ranks,
              "%s: rank %d of %d (at most %d ranks)", who, sh->rank, sh->n_ranks, (int)VSB_MAX_RANKS);
  VSB_REQUIRE(sh->counter != nullptr, "%s: null counter", who);
  for (int r = 0; r < sh->n_ranks; ++r)
    VSB_REQUIRE(sh->flags[r] && sh->sums[r], "%s: null peer pointer for rank %d", who, r);
  return VSB_OK;
}

template <int DIM>
static int chain_impl(const VsbStepArgs& a, const VsbMdfArgs& m, const VsbIbShard& sh, const VsbBodyParams* bp, cudaStream_t stream) {
  constexpr int NC = WinVec<DIM>::NC;
  StepParams<DIM> sp;
  VsbStepArgs b = a;
  if (!b.f_out) b.f_out = sh.fields[sh.rank];   // unused by these kernels; only has to differ from f_in
  b.band = 0; b.sub_begin = 0; b.sub_end = 0; b.edge_rows_only = 0;
  if (int rc = fill_params<DIM>(b, sp)) return rc;
  long long wcells = 1;
  for (int d = 0; d < DIM; ++d) wcells *= m.win_size[d];
  const long long field = wcells * NC;
  const int par = m.parity & 1;
  auto slot = [&](int rank, int parity, int k) { return sh.fields[rank] + ((long long)parity * (m.n_iter + 1) + k) * field; };

  MdfParams p{};
  p.delta_kind = m.delta_kind; p.n_iter = m.n_iter; p.parity = par; p.n_markers = m.n_markers;
  for (int d = 0; d < 3; ++d) { p.origin0[d] = m.win_origin0[d]; p.wsize[d] = d < DIM ? m.win_size[d] : 1; }
  p.markers0 = m.markers0; p.u_target = m.u_target; p.ds_ptr = m.ds_ptr; p.ds_value = m.ds_value;
  p.u_win = slot(sh.rank, par, m.n_iter);
  p.g_win = slot(sh.rank, par, 0); p.g_win_next = slot(sh.rank, par ^ 1, 0);
  p.scratch = slot(sh.rank, par, 1); p.scratch_next = slot(sh.rank, par ^ 1, 1);
  p.marker_u = m.marker_u; p.marker_force = m.marker_force; p.body = m.body;
  p.update_body = (m.body && bp && bp->n_dof > 0) ? 1 : 0;
  p.host_mail = nullptr; p.mail_seq = 0; p.barrier = nullptr;
  p.rotation = (DIM == 2 && m.rotation) ? 1 : 0;
  p.center[0] = m.center[0]; p.center[1] = m.center[1];
  p.m_begin = sh.marker_begin; p.m_end = sh.marker_end;
  const bool tiled = DIM == 3 && m.chunk_offsets != nullptr && sh.chunk_end > sh.chunk_begin;
  p.chunk_offsets = tiled ? m.chunk_offsets : nullptr;
  p.chunk_begin = sh.chunk_begin;
  p.clear_mode = 2;                          // clear only the reachable cells of my need box
  p.clear_cells = sh.cells; p.n_clear_cells = sh.n_cells;
  for (int d = 0; d < 3; ++d) { p.box_lo[d] = sh.need_lo[sh.rank][d]; p.box_hi[d] = d < DIM ? sh.need_hi[sh.rank][d] : 1; }
  p.slab_x[0] = sh.x_lo[sh.rank]; p.slab_x[1] = sh.x_hi[sh.rank];
  p.need_x[0] = sh.need_lo[sh.rank][0]; p.need_x[1] = sh.need_hi[sh.rank][0];
  BodyUpdate bu{};
  if (p.update_body) bu = make_body_update(*bp, DIM);
  const int n_sum = DIM + (p.rotation ? 1 : 0);

  // 1. window velocity of my rows -> every copy that needs it
  WindowPush w{};
  w.n_ranks = sh.n_ranks;
  for (int r = 0; r < sh.n_ranks; ++r) {
    w.dst[r] = slot(r, par, m.n_iter);
    for (int d = 0; d < 3; ++d) { w.lo[r][d] = sh.need_lo[r][d]; w.hi[r][d] = sh.need_hi[r][d]; }
  }
  k_shard_window_moments<DIM><<<blocks_for(sh.n_cells, 128), 128, 0, stream>>>(sp, w, sh.cells, sh.n_cells);
  VSB_LAUNCH_CHECK("vsb_ibshard_chain (window velocity)");
  if (sh.ev_window_done) {
    cudaError_t e = cudaEventRecord((cudaEvent_t)sh.ev_window_done, stream);
    if (e != cudaSuccess) return cuda_fail(e, "vsb_ibshard_chain (event after the window velocity)");
  }
  if (int rc = launch_barrier(sh, p, bu, 0, n_sum, stream)) return rc;
  static const int stop = getenv("VSB_SHARD_STOP") ? atoi(getenv("VSB_SHARD_STOP")) : 1000;   // timing aid: phases to run
  if (stop <= 0) return VSB_OK;

  // 2. the iterations
  ShardDev sd{};
  sd.n_ranks = sh.n_ranks;
  for (int r = 0; r < sh.n_ranks; ++r) {
    for (int d = 0; d < 3; ++d) { sd.lo[r][d] = sh.need_lo[r][d]; sd.hi[r][d] = d < DIM ? sh.need_hi[r][d] : 1; }
    sd.x_lo[r] = sh.x_lo[r]; sd.x_hi[r] = sh.x_hi[r];
  }
  if (DIM == 2) for (int r = 0; r < sh.n_ranks; ++r) { sd.lo[r][2] = 0; sd.hi[r][2] = 1; }
  for (int k = 0; k < m.n_iter; ++k) {
    const bool last = k == m.n_iter - 1;
    p.stage = k; p.stage_end = k + 1;
    sd.by_slab = 0;
    if (!last) {
      sd.n_ranks = sh.n_ranks;
      for (int r = 0; r < sh.n_ranks; ++r) sd.dst[r] = slot(r, par, 1 + k);
    } else {
      // the force field of my markers accumulates in my own copy; the cells of my need box then travel to the ranks
      // whose slabs contain them as plain stores
      ShardDev self{};
      self.n_ranks = 1; self.by_slab = 0;
      self.dst[0] = slot(sh.rank, par, 0);
      for (int d = 0; d < 3; ++d) { self.lo[0][d] = 0; self.hi[0][d] = p.wsize[d]; }
      sd = self;
    }
    if (int rc = launch_mdf_stage_sharded(DIM, p, sd, tiled, (unsigned)(sh.chunk_end - sh.chunk_begin), stream)) return rc;
    if (last) {
      ForcePush fp{};
      fp.n_ranks = sh.n_ranks; fp.rank = sh.rank; fp.src = slot(sh.rank, par, 0); fp.field = field;
      fp.org_shift = m.win_origin0[0];
      for (int r = 0; r < sh.n_ranks; ++r) { fp.staging[r] = sh.staging[r]; fp.x_lo[r] = sh.x_lo[r]; fp.x_hi[r] = sh.x_hi[r]; }
      for (int d = 0; d < 3; ++d) { fp.lo[d] = sh.need_lo[sh.rank][d]; fp.hi[d] = d < DIM ? sh.need_hi[sh.rank][d] : 1; }
      k_shard_push_force<DIM><<<blocks_for(sh.n_cells, 256), 256, 0, stream>>>(fp, sh.cells, sh.n_cells, m.body, par, p.wsize[0],
                                                                            p.wsize[1], p.wsize[2]);
      VSB_LAUNCH_CHECK("vsb_ibshard_chain (force push)");
    }
    if (int rc = launch_barrier(sh, p, bu, last ? 1 : 0, n_sum, stream)) return rc;
    if (stop <= k + 1) return VSB_OK;
  }
  // the body update of the last barrier has written the NEXT step's origin into the other parity slot; this step's
  // origin (slot `par`) is still intact
  ForceReduce fr{};
  fr.n_ranks = sh.n_ranks; fr.rank = sh.rank; fr.staging = sh.staging[sh.rank]; fr.out = sh.force_field; fr.field = field;
  fr.x_lo = sh.x_lo[sh.rank]; fr.x_hi = sh.x_hi[sh.rank]; fr.org_shift = m.win_origin0[0];
  for (int r = 0; r < sh.n_ranks; ++r)
    for (int d = 0; d < 3; ++d) { fr.lo[r][d] = sh.need_lo[r][d]; fr.hi[r][d] = d < DIM ? sh.need_hi[r][d] : 1; }
  k_shard_reduce_force<DIM><<<blocks_for(sh.n_cells, 256), 256, 0, stream>>>(fr, sh.cells, sh.n_cells, m.body, par, p.wsize[0],
                                                                          p.wsize[1], p.wsize[2]);
  VSB_LAUNCH_CHECK("vsb_ibshard_chain (force reduce)");
  return VSB_OK;
}

}  // namespace vsb

using namespace vsb;

extern "C" {

int vsb_ibshard_barrier(const VsbIbShard* shard, vsb_stream_t stream) {
  if (int rc = check_shard(shard, "vsb_ibshard_barrier")) return rc;
  MdfParams p{};
  BodyUpdate bu{};
  return launch_barrier(*shard, p, bu, 0, 0, (cudaStream_t)stream);
}

int vsb_ibshard_chain(const VsbStepArgs* args, const VsbMdfArgs* mdf, const VsbIbShard* shard, const VsbBodyParams* params,
                      vsb_stream_t stream) {
  VSB_REQUIRE(args != nullptr && mdf != nullptr, "vsb_ibshard_chain: null args");
  if (int rc = check_shard(shard, "vsb_ibshard_chain")) return rc;
  VSB_REQUIRE((mdf->dim == 2 || mdf->dim == 3) && mdf->dim == args->grid.dim, "vsb_ibshard_chain: dim must be 2 or 3 and match the grid");
  VSB_REQUIRE(mdf->delta_kind >= VSB_DELTA_PESKIN3 && mdf->delta_kind <= VSB_DELTA_HAT2, "unknown delta kernel %d", mdf->delta_kind);
  VSB_REQUIRE(mdf->n_iter >= 1, "n_iter must be >= 1, got %d", mdf->n_iter);
  VSB_REQUIRE(mdf->n_markers >= 0 && mdf->markers0 && mdf->marker_u && mdf->marker_force, "vsb_ibshard_chain: null buffer");
  VSB_REQUIRE(mdf->host_mail == nullptr, "vsb_ibshard_chain: the rigid-body ODE of a sharded chain runs on the device");
  VSB_REQUIRE(0 <= shard->marker_begin && shard->marker_begin <= shard->marker_end && shard->marker_end <= mdf->n_markers,
              "vsb_ibshard_chain: marker share [%lld, %lld) outside [0, %lld)", (long long)shard->marker_begin,
              (long long)shard->marker_end, (long long)mdf->n_markers);
  for (int r = 0; r < shard->n_ranks; ++r)
    VSB_REQUIRE(shard->fields[r] && shard->staging[r], "vsb_ibshard_chain: null peer field pointer for rank %d", r);
  VSB_REQUIRE(shard->force_field && shard->cells && shard->n_cells >= 0 && shard->n_cells < (1ll << 31),
              "vsb_ibshard_chain: needs the local force field and the list of reachable window cells");
  VSB_REQUIRE(shard->chunk_begin >= 0 && shard->chunk_end >= shard->chunk_begin && shard->chunk_end <= mdf->n_chunks + 0,
              "vsb_ibshard_chain: chunk share outside the chunk list");
  for (int d = 0; d < mdf->dim; ++d) VSB_REQUIRE(mdf->win_size[d] >= 4, "IB window must be at least 4 cells wide");
  return mdf->dim == 2 ? chain_impl<2>(*args, *mdf, *shard, params, (cudaStream_t)stream)
                       : chain_impl<3>(*args, *mdf, *shard, params, (cudaStream_t)stream);
}

}  // extern "C"
