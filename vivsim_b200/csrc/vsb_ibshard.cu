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
//              copy that needs the cell; last iteration: force field into the slab owner's copy) ; k_shard_barrier }
//   the last barrier also exchanges the partial force / torque sums (slot [src] on every rank), adds them in rank
//   order and advances this rank's replica of the rigid body -- identical arithmetic on every rank.
#include <algorithm>
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
};

__global__ void k_shard_barrier(const BarrierParams b, const MdfParams p, const BodyUpdate bu) {
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = b.counter[0] + 1u;
    b.counter[0] = s_epoch;
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

// u(stream(f_in)) on the window cells whose x lies in this rank's rows, multicast by need box.
template <int DIM>
__global__ void __launch_bounds__(128) k_shard_window_moments(const StepParams<DIM> p, const WindowPush w) {
  using L = Lat<DIM>;
  using VecF = typename std::conditional<DIM == 2, float2, float4>::type;
  int org[3];
  window_origin<DIM>(p, org);                       // local coordinates (body origin + win_shift)
  const int x_lo = max(0, p.r_begin - org[0]), x_hi = min(p.wsz[0], p.r_end - org[0]);
  if (x_hi <= x_lo) return;
  const long long plane = (long long)p.wsz[1] * (DIM == 3 ? p.wsz[2] : 1);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)(x_hi - x_lo) * plane) return;
  int rel[3] = {0, 0, 0};
  long long r = t;
  if constexpr (DIM == 3) { rel[2] = (int)(r % p.wsz[2]); r /= p.wsz[2]; }
  rel[1] = (int)(r % p.wsz[1]); r /= p.wsz[1];
  rel[0] = x_lo + (int)r;
  int c[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < L::D; ++d) c[d + L::A0] = org[d] + rel[d];
  float f[L::Q], rho, u[L::D];
  pull_cell<DIM>(p, c[0], c[1], c[2], f, true);
  moments<DIM>(f, rho, u);
  VecF v;
  v.x = u[0]; v.y = u[1];
  if constexpr (DIM == 3) { v.z = u[2]; v.w = 0.f; }
  const long long idx = (long long)rel[0] * plane + (long long)rel[1] * (DIM == 3 ? p.wsz[2] : 1) + rel[2];
  for (int k = 0; k < w.n_ranks; ++k)
    if (rel[0] >= w.lo[k][0] && rel[0] < w.hi[k][0] && rel[1] >= w.lo[k][1] && rel[1] < w.hi[k][1] &&
        (DIM == 2 || (rel[2] >= w.lo[k][2] && rel[2] < w.hi[k][2])))
      reinterpret_cast<VecF*>(w.dst[k])[idx] = v;
  __threadfence_system();
}

static int launch_barrier(const VsbIbShard& sh, const MdfParams& p, const BodyUpdate& bu, int finish, int n_sum,
                          cudaStream_t stream) {
  BarrierParams b{};
  b.n_ranks = sh.n_ranks; b.rank = sh.rank;
  for (int r = 0; r < sh.n_ranks; ++r) { b.flags[r] = sh.flags[r]; b.sums[r] = sh.sums[r]; }
  b.counter = sh.counter; b.body = p.body; b.finish = finish; b.n_sum = n_sum;
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
    VSB_REQUIRE(sh->flags[r] && sh->sums[r] && sh->fields[r], "%s: null peer pointer for rank %d", who, r);
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
  p.clear_mode = 1;
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
  {
    // at most min(window, slab) x-planes of the window lie in this slab; the kernel derives the actual range from the
    // (possibly moving) origin
    const long long planes = std::min<long long>(m.win_size[0], sp.r_end - sp.r_begin);
    const long long n = planes * (wcells / m.win_size[0]);
    if (n > 0) k_shard_window_moments<DIM><<<blocks_for(n, 128), 128, 0, stream>>>(sp, w);
    VSB_LAUNCH_CHECK("vsb_ibshard_chain (window velocity)");
  }
  if (int rc = launch_barrier(sh, p, bu, 0, n_sum, stream)) return rc;

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
    sd.by_slab = last ? 1 : 0;
    for (int r = 0; r < sh.n_ranks; ++r) sd.dst[r] = slot(r, par, last ? 0 : 1 + k);
    if (int rc = launch_mdf_stage_sharded(DIM, p, sd, tiled, (unsigned)(sh.chunk_end - sh.chunk_begin), stream)) return rc;
    if (int rc = launch_barrier(sh, p, bu, last ? 1 : 0, n_sum, stream)) return rc;
  }
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
  VSB_REQUIRE(shard->chunk_begin >= 0 && shard->chunk_end >= shard->chunk_begin && shard->chunk_end <= mdf->n_chunks + 0,
              "vsb_ibshard_chain: chunk share outside the chunk list");
  for (int d = 0; d < mdf->dim; ++d) VSB_REQUIRE(mdf->win_size[d] >= 4, "IB window must be at least 4 cells wide");
  return mdf->dim == 2 ? chain_impl<2>(*args, *mdf, *shard, params, (cudaStream_t)stream)
                       : chain_impl<3>(*args, *mdf, *shard, params, (cudaStream_t)stream);
}

}  // extern "C"
