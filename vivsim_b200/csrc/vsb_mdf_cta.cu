// Whole multi-direct-forcing chain of a small 2-D body (<= 512 markers: the C2 cylinder) in ONE CTA of 1024 threads,
// with the work field in shared memory and NO atomics on floating-point data.
//
// Why: the atomic spreads of the other chains make the marker forces depend on the order in which the reductions
// happen to arrive (run-to-run differences at rounding level).  Here the spread is turned around: the stencil points of
// all markers are bucketed by window cell ONCE per step (a counting sort with integer shared atomics -- fp32 shared
// atomics are compare-and-swap loops, ATOMS.CAST.SPIN, and neighbouring markers hit the same cells), and every
// iteration is
//     marker phase   u_m += 0.5 * sum over the marker's 16 stencil points of w * field[cell]     (16 lanes per marker)
//                    dF = (U - u_m) 2 ds,  F += dF
//     cell phase     field[cell] = sum over the stencil points that fall on the cell of w * dF[m]  (thread per cell)
// which is ib/mdf.py:31-64 with the spread written as a gather.  The points of a cell are visited in ascending
// marker order, so the result is bit-reproducible from run to run.
//
// It was also meant to shorten the critical path of a single 1024^2 domain (two __syncthreads per iteration instead
// of a grid barrier) and does NOT: one SM needs 60 us for the 8192 stencil points of C2 -- launch 1.2, clearing and
// marker set-up 4.3, count + scan 4.5, fill + sort 4.5, velocity gather 17, iterations 5.7 each
// (profiles/r02_cta_chain_phases.txt) -- against 14-18 us for the grid-barrier chain on 64 SMs.  Not the default.
//
//   0. g_win_next cleared (plain stores); per marker: position, target velocity, stencil base, 4 + 4 weights
//   1. count the stencil points per window cell (int atomics; the slot of every point is kept in a register)
//   2. block-wide exclusive scan of (points, touched cells) packed in one word -> start of every cell's bucket and a
//      dense numbering of the touched cells (the work field only has entries for those)
//   3. fill the buckets; each cell sorts its few entries (ascending marker) -- deterministic order
//   4. stage 0: gather the fluid velocity at the stencil points from the pulled populations (or from u_win)
//   5. n_iter x { marker phase ; cell phase }, the last cell phase stores F's spread into g_win (global)
//   6. marker_u, marker_force; total force / torque by an ordered block reduction; body update (finish_body)
#include <cstdlib>

#include "vsb_mdf.cuh"

namespace vsb {

constexpr int kCtaThreads = 1024;
constexpr int kCtaMarkers = 512;
constexpr int kCtaPoints = 16 * kCtaMarkers;          // stencil points = upper bound of the touched cells
constexpr int kCtaNE = kCtaPoints / kCtaThreads;      // stencil points per thread (8)
constexpr size_t kCtaSmemMax = 227 * 1024;

struct CtaShared {
  float wx[kCtaMarkers][4], wy[kCtaMarkers][4];       // delta weights of the 4 + 4 stencil nodes (0 outside the window)
  float2 val[kCtaMarkers];                            // dF (F in the last iteration) of every marker
  float2 u[kCtaMarkers], F[kCtaMarkers], tgt[kCtaMarkers];
  float ds2[kCtaMarkers];
  int bx[kCtaMarkers], by[kCtaMarkers];               // first stencil node (floor(x) - 1), window-local
  float2 field[kCtaPoints];                           // the work field on the touched cells
  unsigned short entries[kCtaPoints];                 // buckets: (marker << 4) | (jx << 2) | jy
  unsigned short cstart[kCtaPoints + 2];              // bucket start of touched cell k (k = n_cells: total)
  unsigned short cwidx[kCtaPoints];                   // flat window index of touched cell k
  unsigned warp_tot[kCtaThreads / 32];
  float red[kCtaThreads / 32][3];
  unsigned n_cells;
};
// followed by unsigned scan[wcells]: points per window cell, then packed exclusive prefix (cells << 16 | points)

static size_t cta_smem_bytes(long long wcells) { return sizeof(CtaShared) + (size_t)wcells * sizeof(unsigned); }

__global__ void __launch_bounds__(kCtaThreads, 1)
k_mdf_cta2d(const StepParams<2> sp, const MdfParams p, const BodyUpdate bu, const int debug_stop) {
  using L = Lat<2>;
  extern __shared__ __align__(16) unsigned char s_raw[];
  CtaShared& sm = *reinterpret_cast<CtaShared*>(s_raw);
  unsigned* scan = reinterpret_cast<unsigned*>(s_raw + sizeof(CtaShared));
  const int tid = threadIdx.x;
  const int n_mark = (int)p.n_markers;
  const int w0 = p.wsize[0], w1 = p.wsize[1];
  const int wcells = w0 * w1;

  // a fused step enqueued behind this launch with early_launch = 1 may start as soon as this CTA is resident
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (debug_stop == 1) return;                      // timing aid (VSB_CTA_STOP): the launch alone
  int org[3] = {p.origin0[0], p.origin0[1], 0};
  if (p.body) { org[0] = p.body->origin2[p.parity][0]; org[1] = p.body->origin2[p.parity][1]; }

  // ---- 0. next step's force field; per-marker set-up
  clear_field<2>(p, p.g_win_next, org[0], 0, tid, kCtaThreads);
  for (int c = tid; c < wcells; c += kCtaThreads) scan[c] = 0u;
  if (tid < n_mark) {
    const int m = tid;
    float pos[2], tgt[2], arm[2];
    marker_kinematics<2>(p, m, pos, tgt, arm);
    const float x = pos[0] - (float)org[0], y = pos[1] - (float)org[1];   // window-local, as marker_x - ib_x0
    const int bx = (int)floorf(x) - 1, by = (int)floorf(y) - 1;
    sm.bx[m] = bx; sm.by[m] = by;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int nx = bx + k, ny = by + k;
      sm.wx[m][k] = (nx >= 0 && nx < w0) ? delta(p.delta_kind, (float)nx - x) : 0.f;   // nodes outside the window are skipped
      sm.wy[m][k] = (ny >= 0 && ny < w1) ? delta(p.delta_kind, (float)ny - y) : 0.f;
    }
    sm.tgt[m] = make_float2(tgt[0], tgt[1]);
    sm.ds2[m] = (p.ds_ptr ? p.ds_ptr[m] : p.ds_value) * 2.0f;
    sm.u[m] = make_float2(0.f, 0.f);
    sm.F[m] = make_float2(0.f, 0.f);
  }
  __syncthreads();
  if (debug_stop == 2) return;                      // + clearing and marker set-up

  // this thread's stencil points: point e = i * 1024 + tid belongs to marker e >> 4 (16 consecutive lanes per marker)
  float w[kCtaNE];
  int cell[kCtaNE];        // flat window index, -1: no point (beyond the markers, outside the window, or weight 0)
  int slot[kCtaNE];        // position inside the cell's bucket; later: dense number of the cell
#pragma unroll
  for (int i = 0; i < kCtaNE; ++i) {
    const int e = i * kCtaThreads + tid, m = e >> 4, jx = (e >> 2) & 3, jy = e & 3;
    w[i] = 0.f; cell[i] = -1; slot[i] = 0;
    if (m < n_mark) {
      const float wt = sm.wy[m][jy] * sm.wx[m][jx];          // the product order of k_mdf_stage
      if (wt != 0.f) {                                       // (zero also for every node outside the window)
        w[i] = wt;
        cell[i] = (sm.bx[m] + jx) * w1 + (sm.by[m] + jy);
        slot[i] = (int)atomicAdd(&scan[cell[i]], 1u);        // ---- 1. count
      }
    }
  }
  __syncthreads();

  // ---- 2. exclusive scan over the window cells of (touched ? 1 : 0) << 16 | points
  {
    const int per = (wcells + kCtaThreads - 1) / kCtaThreads;
    const int c0 = tid * per, c1 = min(c0 + per, wcells);
    unsigned sum = 0u;
    for (int c = c0; c < c1; ++c) { const unsigned n = scan[c]; sum += n | (n ? 0x10000u : 0u); }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
    if ((tid & 31) == 31) sm.warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
      unsigned t = sm.warp_tot[tid];
      unsigned it = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned s = __shfl_up_sync(0xffffffffu, it, o); if (tid >= o) it += s; }
      sm.warp_tot[tid] = it - t;                              // exclusive prefix of the warp totals
      if (tid == 31) { sm.n_cells = it >> 16; sm.cstart[it >> 16] = (unsigned short)(it & 0xffffu); }
    }
    __syncthreads();
    unsigned run = sm.warp_tot[tid >> 5] + (incl - sum);
    for (int c = c0; c < c1; ++c) {
      const unsigned n = scan[c];
      scan[c] = run;
      if (n) {
        sm.cstart[run >> 16] = (unsigned short)(run & 0xffffu);
        sm.cwidx[run >> 16] = (unsigned short)c;
        run += n | 0x10000u;
      }
    }
  }
  __syncthreads();
  const int n_cells = (int)sm.n_cells;
  if (debug_stop == 3) return;                      // + count and scan

  // ---- 3. fill the buckets; every cell orders its entries (ascending marker, then stencil point)
#pragma unroll
  for (int i = 0; i < kCtaNE; ++i)
    if (cell[i] >= 0) {
      const unsigned ex = scan[cell[i]];
      const int e = i * kCtaThreads + tid;
      sm.entries[(ex & 0xffffu) + slot[i]] = (unsigned short)e;        // e = (marker << 4) | (jx << 2) | jy
      slot[i] = (int)(ex >> 16);                                       // from here on: the dense cell number
    }
  __syncthreads();
  for (int k = tid; k < n_cells; k += kCtaThreads) {
    const int b = sm.cstart[k], e = sm.cstart[k + 1];
    for (int a = b + 1; a < e; ++a) {                                   // insertion sort: a handful of entries
      const unsigned short v = sm.entries[a];
      int j = a - 1;
      while (j >= b && sm.entries[j] > v) { sm.entries[j + 1] = sm.entries[j]; --j; }
      sm.entries[j + 1] = v;
    }
  }
  // (the cell phase reads the buckets after the next barrier)
  if (debug_stop == 4) return;                      // + fill and sort

  // ---- 4. + 5. iterations
  for (int stage = 0; stage < p.n_iter; ++stage) {
    const bool last = stage == p.n_iter - 1;
    float um[kCtaNE][2];
#pragma unroll
    for (int i = 0; i < kCtaNE; ++i) {
      um[i][0] = um[i][1] = 0.f;
      if (cell[i] >= 0) {
        if (stage == 0) {
          const int nx = cell[i] / w1, ny = cell[i] - nx * w1;
          if (p.u_win == nullptr) {
            float f[L::Q], rho, u[2];
            pull_cell<2>(sp, 0, org[0] + nx, org[1] + ny, f, true);
            moments<2>(f, rho, u);
            um[i][0] = w[i] * u[0]; um[i][1] = w[i] * u[1];
          } else {
            const float2 v = __ldcg(reinterpret_cast<const float2*>(p.u_win) + cell[i]);
            um[i][0] = w[i] * v.x; um[i][1] = w[i] * v.y;
          }
        } else {
          const float2 v = sm.field[slot[i]];
          um[i][0] = w[i] * v.x; um[i][1] = w[i] * v.y;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kCtaNE; ++i) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        um[i][0] += __shfl_xor_sync(0xffffffffu, um[i][0], o);
        um[i][1] += __shfl_xor_sync(0xffffffffu, um[i][1], o);
      }
      const int e = i * kCtaThreads + tid, m = e >> 4;
      if ((e & 15) == 0 && m < n_mark) {
        float2 u_m = sm.u[m], F = sm.F[m];
        const float2 tgt = sm.tgt[m];
        const float ds2 = sm.ds2[m];
        u_m.x = (stage == 0) ? um[i][0] : u_m.x + 0.5f * um[i][0];
        u_m.y = (stage == 0) ? um[i][1] : u_m.y + 0.5f * um[i][1];
        const float dFx = (tgt.x - u_m.x) * ds2, dFy = (tgt.y - u_m.y) * ds2;
        F.x = (stage == 0 ? 0.f : F.x) + dFx;
        F.y = (stage == 0 ? 0.f : F.y) + dFy;
        sm.u[m] = u_m; sm.F[m] = F;
        sm.val[m] = last ? F : make_float2(dFx, dFy);
      }
    }
    __syncthreads();
    if (debug_stop == 5) return;                    // + gather of the fluid velocity and the first marker phase
    // cell phase
    for (int k = tid; k < n_cells; k += kCtaThreads) {
      const int b = sm.cstart[k], e = sm.cstart[k + 1];
      float ax = 0.f, ay = 0.f;
      for (int a = b; a < e; ++a) {
        const unsigned ent = sm.entries[a];
        const unsigned m = ent >> 4;
        const float wt = sm.wy[m][ent & 3u] * sm.wx[m][(ent >> 2) & 3u];
        const float2 v = sm.val[m];
        ax += v.x * wt; ay += v.y * wt;
      }
      if (last) reinterpret_cast<float2*>(p.g_win)[sm.cwidx[k]] = make_float2(ax, ay);
      else sm.field[k] = make_float2(ax, ay);
    }
    __syncthreads();
  }

  if (debug_stop == 6) return;                      // + all iterations
  // ---- 6. outputs
  float s[3] = {0.f, 0.f, 0.f};
  if (tid < n_mark) {
    const int m = tid;
    const float2 F = sm.F[m], u_m = sm.u[m];
    p.marker_u[m * 2 + 0] = u_m.x; p.marker_u[m * 2 + 1] = u_m.y;
    p.marker_force[m * 2 + 0] = F.x; p.marker_force[m * 2 + 1] = F.y;
    s[0] = F.x; s[1] = F.y;
    if (p.body && p.rotation) {
      float pos[2], tgt[2], arm[2];
      marker_kinematics<2>(p, m, pos, tgt, arm);
      const float Fv[2] = {F.x, F.y};
      s[2] = marker_torque(p, pos, Fv);
    }
  }
  if (p.body) {   // total force (and torque): ordered block reduction, then the body update
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
      if ((tid & 31) == 0) sm.red[tid >> 5][c] = s[c];
    }
    __syncthreads();
    if (tid == 0) {
      float tot[3] = {0.f, 0.f, 0.f};
      for (int wdx = 0; wdx < kCtaThreads / 32; ++wdx)
        for (int c = 0; c < 3; ++c) tot[c] += sm.red[wdx][c];
      for (int c = 0; c < (p.rotation ? 3 : 2); ++c) p.body->force_sum[c] += tot[c];
      if (p.update_body || p.host_mail) finish_body(p, bu);
    }
  }
}

// A 2-D body of at most 512 markers whose window's cell counters fit beside the fixed tables in one CTA's shared
// memory (the C2 window of 109 x 109 cells: 195 KB in all).
bool mdf_cta2d_supported(const MdfParams& p) {
  const long long wcells = (long long)p.wsize[0] * p.wsize[1];
  return p.n_markers > 0 && p.n_markers <= kCtaMarkers && wcells > 0 && wcells < 65536 &&
         cta_smem_bytes(wcells) <= kCtaSmemMax && p.clear_mode != 1;
}

int launch_mdf_cta2d(const StepParams<2>& sp, const MdfParams& p, const BodyUpdate& bu, cudaStream_t stream) {
  const size_t smem = cta_smem_bytes((long long)p.wsize[0] * p.wsize[1]);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_mdf_cta2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCtaSmemMax);
    if (e != cudaSuccess) return cuda_fail(e, "k_mdf_cta2d (shared-memory opt-in)");
    configured = kCtaSmemMax;
  }
  static const int debug_stop = getenv("VSB_CTA_STOP") ? atoi(getenv("VSB_CTA_STOP")) : 0;   // timing aid
  k_mdf_cta2d<<<1, kCtaThreads, smem, stream>>>(sp, p, bu, debug_stop);
  VSB_LAUNCH_CHECK("k_mdf_cta2d");
  return VSB_OK;
}

}  // namespace vsb
