/* vivsim_b200 -- C ABI of the B200-native IB-LBM time step.
 *
 * The reference (haimingz/vivsim v2.0.0) is pure Python over jax.numpy and has no FFI of
 * its own; its boundary is the set of pure functions in vivsim.lbm / lbm3d / ib / ib3d.
 * Each entry point below replaces one of those functions (cited as reference file:line)
 * or the composed step the reference's examples build from them.  This is what an XLA FFI
 * custom call, a ctypes/cffi binding or any other host would bind.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless a parameter is named *_host;
 *    fp32 fields, int32 indices, uint8 masks; SoA C order: f[q][x][y] / f[q][x][y][z];
 *  - outputs are pre-allocated by the caller; nothing is allocated, freed or synchronised;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *  - returns VSB_OK or a negative code; the message is in vsb_last_error() (thread-local);
 *  - nothing throws across the ABI; calls are re-entrant per device;
 *  - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef VIVSIM_B200_H
#define VIVSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vsb_stream_t; /* cudaStream_t */

enum { VSB_OK = 0, VSB_ERR_INVALID = -1, VSB_ERR_CUDA = -2, VSB_ERR_TIMEOUT = -3 };
enum { VSB_COLL_BGK = 0, VSB_COLL_MRT = 1, VSB_COLL_KBC = 2, VSB_COLL_REG = 3 };
enum { VSB_FORCE_NONE = 0, VSB_FORCE_EDM = 1, VSB_FORCE_GUO = 2 };
enum { VSB_BC_NEE = 0, VSB_BC_NEBB = 1, VSB_BC_EQUILIBRIUM = 2, VSB_BC_BOUNCE_BACK = 3, VSB_BC_SPECULAR = 4,
       VSB_POST_MASK = 5 };
enum { VSB_WRAP_NONE = 0, VSB_WRAP_VELOCITY = 1, VSB_WRAP_PRESSURE = 2, VSB_WRAP_FORCE_CORRECTED = 3 };
/* loc: x faces, y faces, z faces (reference lbm/lattice.py:64-109, lbm3d/lattice.py:64-125) */
enum { VSB_LOC_LEFT = 0, VSB_LOC_RIGHT = 1, VSB_LOC_BOTTOM = 2, VSB_LOC_TOP = 3, VSB_LOC_BACK = 4, VSB_LOC_FRONT = 5 };
enum { VSB_DELTA_PESKIN3 = 0, VSB_DELTA_PESKIN4 = 1, VSB_DELTA_COSINE4 = 2, VSB_DELTA_HAT2 = 3 };

/* dim = 2 (D2Q9, nz ignored) or 3 (D3Q19). */
typedef struct { int dim, nx, ny, nz; } VsbGrid;

/* A wall quantity: per-face-cell device array (face shape = grid shape without the normal axis)
 * when ptr != NULL, else the scalar `value` (reference lbm/boundary/_helpers.py:53-77). */
typedef struct { const float* ptr; float value; } VsbWallValue;

/* One post-streaming operation. */
typedef struct {
  int kind;            /* VSB_BC_* or VSB_POST_MASK */
  int wrap;            /* VSB_WRAP_* (NEE / NEBB / EQUILIBRIUM only) */
  int loc;             /* VSB_LOC_* */
  VsbWallValue rho;    /* rho_wall (default 1)                                   */
  VsbWallValue u[3];   /* ux_wall, uy_wall, uz_wall                              */
  VsbWallValue g[3];   /* gx_wall ... (VSB_WRAP_FORCE_CORRECTED)                 */
  const uint8_t* mask; /* VSB_POST_MASK: obstacle mask over the grid (1 = solid) */
} VsbPostOp;

int vsb_abi_version(void);
const char* vsb_last_error(void);

/* ---- D2Q9 / D3Q19 operators (one per reference function) ------------------------------ */

/* streaming: lbm/basic.py:60-85, lbm3d/basic.py:52-87.  Periodic; bit-exact permutation. */
int vsb_streaming(const VsbGrid* grid, const float* f, float* out, vsb_stream_t stream);
/* get_macroscopic: lbm/basic.py:88-111, lbm3d/basic.py:90-106.  n_cells = any flattened spatial shape. */
int vsb_macroscopic(int dim, int64_t n_cells, const float* f, float* rho, float* u, vsb_stream_t stream);
/* get_equilibrium: lbm/basic.py:114-135, lbm3d/basic.py:109-130. */
int vsb_equilibrium(int dim, int64_t n_cells, const float* rho, const float* u, float* feq, vsb_stream_t stream);
/* collision_bgk / _mrt / _kbc / _reg: lbm/basic.py:138-156, lbm/collision/{mrt.py:66-88,kbc.py:10-59,reg.py:4-49}
 * and the lbm3d twins.  op_host: Q*Q row-major HOST matrix for VSB_COLL_MRT (else NULL). */
int vsb_collision(int dim, int64_t n_cells, int kind, double omega, const float* op_host, const float* f,
                  const float* feq, float* out, vsb_stream_t stream);
/* get_guo_forcing_term: lbm/forcing/guo.py:6-33, lbm3d/forcing/guo.py:13-38. */
int vsb_guo_forcing_term(int dim, int64_t n_cells, const float* g, const float* u, float* out, vsb_stream_t stream);
/* forcing_edm (kind EDM), forcing_guo_bgk (GUO, fop_host NULL), forcing_guo_mrt (GUO, fop_host Q*Q HOST matrix):
 * lbm/forcing/edm.py:31, lbm/forcing/guo.py:38-57,82-107 and the lbm3d twins. */
int vsb_forcing(int dim, int64_t n_cells, int kind, double omega, const float* fop_host, const float* f,
                const float* g, const float* u, float* out, vsb_stream_t stream);
/* One boundary / mask operation applied IN PLACE on post-streaming f.
 * boundary_{nee,nebb,equilibrium} and their velocity / pressure / force_corrected wrappers:
 *   lbm/boundary/{nee.py:19-65,nebb.py:20-62,eq.py:25-61,_helpers.py:80-239}, lbm3d/boundary/* twins;
 * boundary_bounce_back / boundary_specular_reflection (need f_pre = pre-streaming f):
 *   lbm/boundary/bb.py:15-95, lbm3d/boundary/bb.py:9-53;
 * obstacle_bounce_back (VSB_POST_MASK): lbm/boundary/bb.py:98-110, lbm3d/boundary/bb.py:56-59. */
int vsb_post_op(const VsbGrid* grid, const VsbPostOp* op, const float* f_pre, float* f, vsb_stream_t stream);
/* boundary_characteristic: lbm/boundary/cbc.py:14-53, lbm3d/boundary/cbc.py:16-51.
 * rho_out: face-shaped, u_out: (dim, face). */
int vsb_boundary_characteristic(const VsbGrid* grid, int loc, const float* rho, const float* u, float* rho_out,
                                float* u_out, vsb_stream_t stream);

/* ---- field diagnostics and grid-refinement transfers ------------ */

/* One diagnostic of reference vivsim/post.py per kind.  Gradients are jnp.gradient's: central differences inside,
 * one-sided at the array edges, unit spacing (post.py:17-29); every axis then needs >= 2 cells. */
enum {
  VSB_DIAG_VELOCITY_MAGNITUDE = 0,    /* post.py:6-14      u (dim, S) -> (S)                  */
  VSB_DIAG_VELOCITY_GRADIENT = 1,     /* post.py:17-29     -> (dim, dim, S), [i][j] = du_i/dx_j */
  VSB_DIAG_VORTICITY = 2,             /* post.py:32-55     -> (S) in 2-D, (3, S) in 3-D       */
  VSB_DIAG_VORTICITY_MAGNITUDE = 3,   /* post.py:58-66                                        */
  VSB_DIAG_DIVERGENCE = 4,            /* post.py:70-80                                        */
  VSB_DIAG_STRAIN_RATE = 5,           /* post.py:83-92     -> (dim, dim, S)                   */
  VSB_DIAG_STRAIN_RATE_MAGNITUDE = 6, /* post.py:95-103                                       */
  VSB_DIAG_KINETIC_ENERGY = 7,        /* post.py:106-114                                      */
  VSB_DIAG_PRESSURE = 8,              /* post.py:129-139   in = rho (S), param = cs2          */
  VSB_DIAG_ENSTROPHY = 9,             /* post.py:142-152                                      */
  VSB_DIAG_Q_CRITERION = 10           /* post.py:163-177                                      */
};
/* in: u (dim, S) (rho (S) for VSB_DIAG_PRESSURE); out: shape per kind above.  One pass, no temporaries. */
int vsb_post_field(const VsbGrid* grid, int kind, const float* in, float param, float* out, vsb_stream_t stream);
/* Domain mean of a scalar diagnostic without materialising it: mean_kinetic_energy (post.py:117-126),
 * mean_enstrophy (post.py:155-160).  workspace: one device double (cleared by the call); out: one device float. */
int vsb_post_mean(const VsbGrid* grid, int kind, const float* in, float param, double* workspace, float* out,
                  vsb_stream_t stream);

/* Grid-refinement transfers between D2Q9 blocks whose spacing differs by 2 (reference vivsim/multigrid.py).
 * dir names the side on which the populations travel: left (3,7,6), right (1,5,8), up (2,5,6), down (4,7,8). */
enum { VSB_MG_LEFT = 0, VSB_MG_RIGHT = 1, VSB_MG_UP = 2, VSB_MG_DOWN = 3 };
/* fine_to_coarse: multigrid.py:58-101.  IN PLACE on f_coarse (9, nx_c, ny_c): its receiving edge line <- mean of
 * the 2 x 2 fine cells of the two outermost fine layers.  Needs ny_f = 2 ny_c (left/right) or nx_f = 2 nx_c. */
int vsb_mg_fine_to_coarse(int nx_f, int ny_f, const float* f_fine, int nx_c, int ny_c, float* f_coarse, int dir,
                          vsb_stream_t stream);
/* coarse_to_fine: multigrid.py:103-131.  IN PLACE on f_fine (9, nx_f, ny_f): piecewise-constant copy of the coarse
 * edge line into the fine receiving line. */
int vsb_mg_coarse_to_fine(int nx_c, int ny_c, const float* f_coarse, int nx_f, int ny_f, float* f_fine, int dir,
                          vsb_stream_t stream);

/* ---- immersed boundary (reference vivsim/ib, vivsim/ib3d) ------------------------------ */

/* kernel_peskin_3pt / _4pt / kernel_cosine_4pt: ib/kernels.py:4-61 (+ 2-point hat, not in the reference). */
int vsb_ib_delta(int kind, int64_t n, const float* r, float* out, vsb_stream_t stream);
/* get_ib_stencil: ib/stencil.py:7-51 (dim 2, coords = (M,2) rows x,y; ny) and ib3d/stencil.py:7-59
 * (dim 3, coords (M,3); ny, nz).  weights, indices: (M, (2*radius)^dim). */
int vsb_ib_stencil(int dim, int kind, int radius, int64_t n_markers, const float* coords, int ny, int nz,
                   float* weights, int32_t* indices, vsb_stream_t stream);
/* interpolate: ib/stencil.py:54-78.  grid (C, n_cells) -> out (M, C). */
int vsb_ib_interpolate(int n_comp, int64_t n_cells, const float* grid, int64_t n_markers, int n_stencil,
                       const float* weights, const int32_t* indices, float* out, vsb_stream_t stream);
/* spread: ib/stencil.py:81-110.  grid (C, n_cells) += scatter of values (M, C); atomics. */
int vsb_ib_spread(int n_comp, int64_t n_cells, float* grid, int64_t n_markers, int n_stencil,
                  const float* values, const float* weights, const int32_t* indices, vsb_stream_t stream);
/* The same scatter-add with a DETERMINISTIC summation order: every cell receives its contributions in the order of
 * the flattened (marker, stencil point) index, products and sums rounded separately -- reproducible run to run and
 * bit-identical to a sequential scatter (NumPy add.at, XLA's CPU scatter).  One stable radix sort of the M * NS cell
 * indices, then one thread per run of equal indices.  `workspace`: device memory of at least
 * vsb_ib_spread_ordered_workspace(n_markers, n_stencil) bytes (a host-side size query; < 0: too many entries). */
int64_t vsb_ib_spread_ordered_workspace(int64_t n_markers, int n_stencil);
int vsb_ib_spread_ordered(int n_comp, int64_t n_cells, float* grid, int64_t n_markers, int n_stencil, const float* values,
                          const float* weights, const int32_t* indices, void* workspace, int64_t workspace_bytes,
                          vsb_stream_t stream);

/* ---- rigid body state (device resident) ------------------------------------------------ *
 * The IB window follows a moving body.  Its integer origin for the step with parity p (0/1, alternating every
 * step) is origin2[p]; it is produced by the body update of the PREVIOUS step (vsb_body_newmark / vsb_ib_fused,
 * or the host in host-ODE mode), so that all kernels of one step can run concurrently while the update writes
 * the other slot. */
typedef struct {
  float d[3], v[3], a[3]; /* displacement, velocity, acceleration of the rigid body                       */
  float h[3];             /* last total hydrodynamic force on the body: sum over markers of -F + a*added_mass */
  float force_sum[3];     /* accumulator: sum over markers of +F; cleared by the body update             */
  int origin2[2][3];      /* integer IB-window origin for step parity 0 / 1                               */
  int ticket;             /* CTA arrival counter of the last MDF stage (0 between steps)                  */
  int step;               /* number of body updates performed so far (index of the next history row)     */
} VsbBodyState;

/* Structural parameters and window rule of a rigid body (reference dyn.py:5-51 with gamma = 1/2, beta = 1/4,
 * dt = 1; coupling of examples/2d/vortex_induced_vibration.py:104-105,135-137).  Translation in 1..dim components;
 * in 2-D a third degree of freedom is the rotation about `center` (dyn.py:54-57 newmark_3dof, :84-120 marker
 * kinematics, :139-154 torque): d[2] = angle, v[2] = angular velocity, force_sum[2] = torque of +F. */
typedef struct {
  int n_dof;              /* 1..3 degrees of freedom; 0 = fixed body (no update)                           */
  int follow;             /* window rule: 0 fixed, 1 trunc(origin0 + d) (2-D VIV example :104-105),
                             2 clip(floor(origin0 + d)) (examples/3d/oscillating_cylinder.py:241-243)       */
  float origin0[3];       /* window origin for d = 0                                                       */
  int grid_size[3], win_size[3];
  double m, k, c, added_mass;
  float* history;         /* optional ring of `history_capacity` rows of 6 floats: after the update of step n, row
                             (n mod capacity) <- d[0..2], h[0..2] -- the per-step (d, h) record the reference's
                             update_chunk returns (examples/2d/vortex_induced_vibration.py:150-157).  DEVICE memory for
                             the device update, HOST memory for vsb_body_newmark_host / vsb_step_host_ode             */
  int history_capacity;
  int rotation;           /* 1 (2-D, n_dof = 3): degree of freedom 2 is the rotation about `center`         */
  float center[2];        /* x_center_init, y_center_init of dyn.py:84-154                                  */
  int matrix_form;        /* 0: the scalars m, k, c, added_mass act on every degree of freedom (dyn.py:44-46);
                             1: multi-DOF form of dyn.py:36-42 with the row-major 3 x 3 matrices below (their
                             leading n_dof x n_dof block) and one added mass per degree of freedom            */
  double mat_m[9], mat_k[9], mat_c[9];
  double added_mass_v[3];
} VsbBodyParams;

/* Mailbox in page-locked host memory (reachable from the device under unified addressing). */
typedef struct VsbHostMail {
  float force[3];   /* sum over markers of +F for the step `seq`                               */
  int seq;          /* written last by the device                                              */
  int next;         /* host side only: sequence number the next step will use (monotonic)     */
} VsbHostMail;

/* multi_direct_forcing with the stencil computed on the fly (ib/mdf.py:10-64 + ib/stencil.py:27-51 /
 * ib3d/stencil.py:36-57), marker-parallel over many CTAs, on a window of the grid: n_iter launches.
 * The velocity at the stencil points comes straight from the streamed populations of `args` (f_in, do_stream, mask),
 * so no velocity field is materialised.  The last stage adds the total force to body->force_sum and, when
 * params->n_dof > 0, the last CTA performs the body update of vsb_body_newmark.
 * Window fields are stored cell-major with the components packed per cell: float2 in 2-D, float4 (last unused) in
 * 3-D, i.e. shape (wnx, wny[, wnz], 2 | 4) -- one vector gather / vector reduction per stencil point.
 *   u_win                      optional precomputed window velocity (vsb_ib_window_moments); NULL -> stage 0 takes the
 *                              velocity at each stencil point from the streamed populations (better for sparse markers)
 *   g_win / scratch            this step's buffers: force field and (n_iter-1) work fields; ZERO on entry
 *   g_win_next / scratch_next  the buffers the NEXT step will use: cleared by this call (double-buffer by step
 *                              parity, so nothing is cleared while a kernel of this step may still read it)
 *   markers0      (M, dim) marker coordinates; the body state adds its displacement
 *   u_target      (M, dim) or NULL -> every marker targets the body velocity (body != NULL) or 0
 *   ds            (M) when ds_ptr != NULL else the scalar ds_value
 *   marker_u, marker_force (M, dim) work / output arrays (marker_force = +F; reaction = -F)
 *   body          device VsbBodyState or NULL (fixed body at markers0, window origin = win_origin0) */
typedef struct {
  int dim, delta_kind, n_iter, parity;
  int64_t n_markers;
  int win_origin0[3], win_size[3];     /* x, y, z order; unused trailing entries ignored for dim 2 */
  const float* markers0;
  const float* u_target;
  const float* ds_ptr;
  float ds_value;
  const float* u_win;
  float* g_win;
  float* g_win_next;
  float* scratch;
  float* scratch_next;
  float* marker_u;
  float* marker_force;
  VsbBodyState* body;
  void* barrier;   /* optional 8-byte device counter (zero-initialised once): lets small bodies (<= 120 CTAs) run all
                      iterations in one launch separated by grid barriers instead of one launch per iteration */
  struct VsbHostMail* host_mail;   /* optional, host-ODE mode: page-locked HOST memory the last CTA of the last stage
                      writes the total marker force to (force[0..2], then seq = mail_seq, system-scope release), so
                      that the host can pick it up by polling instead of a copy + stream synchronisation */
  int mail_seq;
  const int32_t* chunk_offsets;    /* optional, dense 3-D bodies (u_win != NULL): n_chunks + 1 device offsets into the
                      marker arrays; chunk c = markers [offsets[c], offsets[c+1]) (at most 256) is handled by one CTA
                      of the tiled kernel, which stages the chunk's bounding box of the window in shared memory.
                      The caller orders the markers and cuts the chunks so that each box (+1 cell per axis for a moving
                      body) holds at most 2304 cells; a chunk that does not fit falls back to global reductions.
                      NULL: fixed chunks of 256 consecutive markers */
  int n_chunks;
  int rotation;                    /* 2-D rigid rotation (see VsbBodyParams): marker position
                      center + d[0:2] + R(d[2]) (markers0 - center), target velocity v[0:2] + v[2] x lever arm,
                      torque of +F accumulated into body->force_sum[2]                                        */
  float center[2];
  int chain_mode;                  /* how a small body's iterations are chained inside one launch:
                      0 auto (1 for small bodies, else 3); 1 grid barriers through `barrier` (cooperative launch); 2 one thread-block cluster that
                      iterates in marker space (2-D, <= 512 markers, needs nbr_list); 3 one launch per iteration;
                      4 ONE CTA with the work field in shared memory and the spread done as a gather over per-cell
                      buckets -- no floating-point atomics, bit-reproducible (2-D, <= 512 markers, window of fewer than
                      ~19000 cells; measured 3-4x slower than 1 on BASELINE config 2, kept for reproducibility) */
  const uint16_t* nbr_list;        /* chain_mode 2: (nbr_stride, 512) device array, neighbour-major: column m lists the
                      markers whose 4 x 4 stencil can overlap that of marker m under any rigid motion of the body (m
                      itself included), padded with 0xffff; nbr_stride a multiple of 4, <= 48.  NULL: no cluster kernel */
  int nbr_stride;
  const int32_t* reach_cells;      /* optional: ascending flat window indices of the cells a marker stencil can EVER reach
                      (window following the body: the stencils at rest plus one cell of drift either way).  The call then
                      clears only those cells of g_win_next / scratch_next -- for a finely meshed surface that is a thin
                      shell of the window (BASELINE config 5: 1 / 8 of it).  NULL: the whole window */
  int64_t n_reach_cells;
} VsbMdfArgs;

/* ---- fused time step ------------------------------------------------------------------- *
 * State convention: `f_in` / `f_out` hold the POST-COLLISION populations S_n = collide(F_n),
 * where F_n is the reference's carried state (post-streaming, post-boundary).  One call does
 *     S_{n+1} = collide( post_ops( stream(S_n) ) )
 * in one pass over the grid (pull streaming + moments + collision + forcing), followed, when
 * post_ops is non-empty, by an ordered in-place fix-up of the wall lines.  With do_stream = 0
 * the call is the prologue S_0 = collide(F_0); with do_collide = 0 it is the epilogue
 * F_n = post_ops(stream(S_{n-1})).  (SURVEY.md 7 hard-part 1.)                              */
typedef struct {
  VsbGrid grid;          /* LOCAL array extent (including ghost layers when decomposed)       */
  int collision;         /* VSB_COLL_*                                                        */
  int forcing;           /* VSB_FORCE_*                                                       */
  double omega;
  const float* mrt_op_host;   /* Q*Q HOST, VSB_COLL_MRT                                        */
  const float* mrt_fop_host;  /* Q*Q HOST, VSB_COLL_MRT + VSB_FORCE_GUO                        */
  int do_stream, do_collide;
  int row_begin, row_end;     /* range of the slowest axis (x) to update; row_end = 0 -> whole extent */
  const float* f_in;
  float* f_out;
  float g_uniform[3];         /* uniform body force added everywhere                           */
  const float* g_win;         /* optional force field on a window, cell-major (wnx, wny[, wnz], 2 | 4) */
  int win_origin[3], win_size[3];
  const VsbBodyState* body;   /* optional: window origin is read from body->origin2[parity]    */
  int parity;
  int n_post;
  const VsbPostOp* post;      /* HOST array of ordered post-streaming operations (at most one
                                 VSB_POST_MASK; it is also applied to interior cells in the fused pass) */
  int vec;                    /* cells per thread along the contiguous axis: 0 = auto, 1, 2, 4 */
  int band;                   /* 0: all cells; 1: all cells except the band of the force window; 2: only that
                                 band (lets the bulk run concurrently with the IB kernels).  The band is the
                                 window's x-range in 2-D and its (x, y) footprint over all z in 3-D */
  int edges;                  /* 0: ordered wall fix-up inside vsb_step; 1: none -- the caller runs
                                 vsb_edge_fused and the fused pass leaves those wall layers untouched;
                                 2: the wall layers are processed by extra blocks of the same launch
                                 (1 and 2 need independent face operations, see vsb_edge_fused_supported) */
  int sub_begin, sub_end;     /* rows of [row_begin, row_end) this launch updates (sub_end = 0: all of them);
                                 lets edge rows, which read ghost layers, run later than the interior */
  int edge_rows_only;         /* 1: update just the first and the last physical row (ignores sub_begin / sub_end) */
  int win_shift[3];           /* added to body->origin2[parity] by the fluid kernels: the body state of a slab-decomposed
                                 run carries GLOBAL coordinates (identical on every rank) while `grid` is the local
                                 slab with its ghost layer: win_shift[0] = 1 - x0 of the slab */
  const struct VsbHaloArgs* halo;  /* optional, with edge_rows_only = 1 on a slab of a decomposed run: the launch itself does the
                                 halo hand-shake that vsb_halo_wait / vsb_halo_send would do around it (halo->state must
                                 be f_out): halo_mode bit 0 -- every CTA first waits until both neighbours have published
                                 the current step; bit 1 -- the populations crossing a cut are ALSO stored into the
                                 neighbours' ghost rows from the kernel's epilogue and the last CTA publishes the step.
                                 One launch instead of three per step */
  int halo_mode;
  int early_launch;           /* 1: the fused kernel may start while the kernel enqueued before it on `stream` is still
                                 running (programmatic dependent launch; that kernel must not produce anything this
                                 launch reads).  Used to put a body's short IB chain on the SMs FIRST and let the bulk
                                 pass (band = 1) fill the rest of the device behind it */
} VsbStepArgs;

int vsb_step(const VsbStepArgs* args, vsb_stream_t stream);

int vsb_ib_mdf(const VsbStepArgs* args, const VsbMdfArgs* mdf, const VsbBodyParams* params, vsb_stream_t stream);


/* Wall layers in ONE kernel per face: pull the streamed populations of the wall cell (and of the adjacent fluid
 * cell when the operation reads it), apply the face operation, collide, store.  Valid when the face operations are
 * independent of each other: all on faces normal to one non-contiguous axis, at least 3 layers apart.
 * vsb_edge_fused_supported returns 1 in that case, else 0 (use edges = 0). */
int vsb_edge_fused_supported(const VsbStepArgs* args);
int vsb_edge_fused(const VsbStepArgs* args, vsb_stream_t stream);

/* Velocity of the streamed state on the IB window, cell-major (wnx, wny[, wnz], 2 | 4): u_win <- u(stream(f_in)).  Uses grid, f_in,
 * do_stream, win_origin / body + parity, win_size and the mask of `args`.  The window must not contain cells of a
 * face that carries a boundary operation. */
int vsb_ib_window_moments(const VsbStepArgs* args, float* u_win, vsb_stream_t stream);
/* The same on a list of window cells only (VsbMdfArgs.reach_cells): the other cells of u_win are left untouched and are
 * never multiplied by a non-zero stencil weight. */
int vsb_ib_window_moments_cells(const VsbStepArgs* args, float* u_win, const int32_t* cells, int64_t n_cells,
                                vsb_stream_t stream);

/* Body update on the device: h = -force_sum + a*added_mass; (a,v,d) <- newmark(a,v,d,h,m,k,c); force_sum <- 0;
 * origin2[parity ^ 1] <- window origin for the next step.  `parity` is the parity of the step being completed. */
int vsb_body_newmark(VsbBodyState* body, const VsbBodyParams* params, int parity, vsb_stream_t stream);

/* The same body update with the ODE on the HOST (north_star keeps dyn.py's rigid-body ODE on the host): copies the
 * body state to `pinned` (page-locked host memory), SYNCHRONISES `stream`, advances (a, v, d) on the CPU, and copies
 * the state back asynchronously.  The only entry point that synchronises. */
int vsb_body_newmark_host(VsbBodyState* body, VsbBodyState* pinned, const VsbBodyParams* params, int parity,
                          vsb_stream_t stream);

/* The whole immersed-boundary part of one step in ONE kernel (single CTA, scratch in shared memory):
 * velocity at the stencil points from the streamed state, all multi-direct-forcing iterations, the force field
 * written to mdf->g_win (every window cell, no memset needed), the body update of vsb_body_newmark.
 * Only when the window fits shared memory: dim * window_cells * 4 B <= 200 KB (vsb_ib_fused_supported). */
int vsb_ib_fused_supported(const VsbMdfArgs* mdf);
int vsb_ib_fused(const VsbStepArgs* args, const VsbMdfArgs* mdf, const VsbBodyParams* params, vsb_stream_t stream);

/* One whole time step with the rigid-body ODE on the host, orchestrated in a single call (the per-step CPU cost of
 * issuing it from an interpreted host language would otherwise dominate): bulk of the grid on `main`; on `ib` the MDF
 * chain, vsb_body_newmark_host (synchronises `ib` only) and the window's x-range; wall kernels on `edge`; joins back
 * into `main`.  Streams and events are created by the caller (cudaStream_t / cudaEvent_t passed as void*).
 * args->edges must be 1 when there are face operations (independent ones, see vsb_edge_fused_supported). */
typedef struct {
  void* main; void* ib; void* edge;            /* cudaStream_t */
  void* ev_fork; void* ev_ib; void* ev_edge;   /* cudaEvent_t  */
} VsbHostPlan;

int vsb_step_host_ode(VsbStepArgs* args, const VsbMdfArgs* mdf, const VsbBodyParams* params, VsbBodyState* pinned,
                      const VsbHostPlan* plan);

/* n_steps whole time steps with the rigid-body ODE on the host, in ONE call: no per-step cost in the host language,
 * no stream synchronisation and no device->host copy.  Per step: bulk rows on `main`; on `ib` the MDF chain, whose
 * last CTA posts the total force into mdf->host_mail; the host polls the mailbox, advances (a, v, d) on the CPU
 * (dyn.py:5-51; `pinned` is the master copy of the body state), sends the 92-byte state back with one asynchronous
 * host->device copy and launches the window's x-range behind it.  args / mdf must be set up as for
 * vsb_step_host_ode for the first step (f_in = current state, f_out = the other buffer, parity, g_win / scratch
 * pairs); they are advanced in place (buffers and parity swapped every step), so after the call args->f_in is the
 * current state.  Returns VSB_ERR_CUDA if the device does not answer within ~10 s. */
int vsb_run_host_ode(VsbStepArgs* args, VsbMdfArgs* mdf, const VsbBodyParams* params, VsbBodyState* pinned,
                     const VsbHostPlan* plan, int n_steps);

/* NON-BLOCKING variant of vsb_run_host_ode: the n_steps steps are only ENQUEUED and the call returns at once -- nothing
 * is polled or synchronised on the calling thread, so it can sit inside a host-language custom call (an XLA FFI
 * handler must not block its stream's thread).  Per step, on `ib`: the MDF chain posts the total force into
 * mdf->host_mail, a host function (cudaLaunchHostFunc) that the driver runs once the chain has finished advances
 * (a, v, d) on the CPU (dyn.py:5-51, `pinned` is the master copy), one asynchronous 92-byte host->device copy sends the
 * state back; bulk rows on `main`, window x-range on `ib` before the host function, joined into `main`.  args / mdf as
 * for vsb_run_host_ode, advanced in place.  `pinned`, *params, mdf->host_mail and the device buffers must stay valid
 * until the work enqueued on `main` has completed.  A step pays the driver's host-function latency (tens of
 * microseconds) instead of the mailbox poll's few: prefer vsb_run_host_ode[_multi] where blocking is acceptable.
 * Not capturable into a CUDA graph (the per-call context is released by the last host function); synchronise `main`
 * before using vsb_run_host_ode / vsb_step_host_ode on the same body (they read `pinned` on the calling thread). */
int vsb_enqueue_host_ode(VsbStepArgs* args, VsbMdfArgs* mdf, const VsbBodyParams* params, VsbBodyState* pinned,
                         const VsbHostPlan* plan, int n_steps);

/* The same for n_domains (1..64) INDEPENDENT simulations on one GPU (an ensemble: e.g. the reduced-velocity sweep of a
 * VIV study, each case being one run of examples/2d/vortex_induced_vibration.py), n_steps each, in one call.  Every
 * domain brings its own argument blocks, body state, mailbox and streams (plans[i]->main / ib must differ between
 * domains).  One host thread serves all mailboxes round-robin: whichever domain's force has arrived gets its body
 * advanced (dyn.py:5-51), its state sent back and its next step enqueued, so the device always has other domains'
 * kernels to run while one domain waits for the host -- the per-step host round trip is hidden instead of paid.
 * Runs of >= 32 steps record each domain's step (both parities) as CUDA graphs after two kernel-by-kernel steps and
 * replay them: ONE driver call per step (the graph's first node is the 92-byte copy of the body state from `pinned`).
 * The graphs are kept for the next call with the same `plans[i]` and are replayed again as long as every argument a
 * captured step depends on is unchanged (both argument blocks, face operations, MRT operators, streams, `pinned`,
 * mailbox -- compared byte by byte), else recorded anew.  VSB_HOST_ODE_GRAPH=0 switches graphs off,
 * VSB_HOST_ODE_THREADS=T serves the domains from T host threads (kernel by kernel). */
int vsb_run_host_ode_multi(int n_domains, VsbStepArgs* const* args, VsbMdfArgs* const* mdfs,
                           const VsbBodyParams* const* params, VsbBodyState* const* pinned,
                           const VsbHostPlan* const* plans, int n_steps);

/* ---- multi-GPU: halo exchange over peer memory ----------------------------------------- *
 * Slab decomposition along x, one ghost layer per side (local extent grid.nx = nx_local + 2).  Replaces the
 * reference's vivsim/multidevice.py:13-38 (four lax.ppermute of the populations crossing a cut).  One kernel copies
 * the edge layers of the crossing populations (3 of 9 in D2Q9, 5 of 19 in D3Q19) of `state` directly into the ring
 * neighbours' ghost layers through peer-mapped pointers (NVLink), publishes the step number in the neighbours' flag
 * words and waits for theirs, so that everything enqueued after it sees complete ghost layers.  Graph-capturable.
 *   left_state / right_state   address of the SAME buffer in the left / right neighbour, mapped into this process
 *   my_flags (2 words)         [0] written by the left neighbour, [1] by the right neighbour
 *   left_flags / right_flags   the neighbours' flag words, peer-mapped
 *   counter (3 words, local)   step number, CTA ticket, timeout indicator (set to 1 if a neighbour never arrived) */
typedef struct VsbHaloArgs {
  VsbGrid grid;
  const float* state;
  float* left_state;
  float* right_state;
  uint32_t* my_flags;
  uint32_t* left_flags;
  uint32_t* right_flags;
  uint32_t* counter;
} VsbHaloArgs;

int vsb_halo_push(const VsbHaloArgs* args, vsb_stream_t stream);

/* The two halves of vsb_halo_push for overlapping the hand-shake with compute:
 *   vsb_halo_send   copy the edge layers into the neighbours' ghost layers and publish the step number (no wait);
 *   vsb_halo_wait   wait until both neighbours have published the current step number.
 * Per step: interior rows (no ghost dependency) run at once; vsb_halo_wait -> edge rows -> vsb_halo_send run on a
 * second stream.  vsb_halo_push == send followed by wait. */
int vsb_halo_send(const VsbHaloArgs* args, vsb_stream_t stream);
int vsb_halo_wait(const VsbHaloArgs* args, vsb_stream_t stream);

/* The cross-GPU waits above (and the flag barriers of vsb_ibshard_chain) never hang: after 10 s without an answer they
 * give up, set word [2] of their `counter` and let the stream run on with stale ghost data.  vsb_sync_status makes that
 * visible to the host: it waits for `stream` (the ONE synchronising call of this group -- call it where the host reads
 * results anyway), returns VSB_ERR_TIMEOUT with the step number in vsb_last_error() if the indicator is set, else VSB_OK.
 * `counter` is VsbHaloArgs.counter or VsbIbShard.counter. */
int vsb_sync_status(const uint32_t* counter, vsb_stream_t stream);

/* ---- multi-GPU: immersed-boundary chain shared by all ranks over peer memory --------------------------------- *
 * The slab decomposition cuts the fluid along x; an immersed body is compact, so "owner computes" would leave its
 * whole multi-direct-forcing chain (ib/mdf.py:10-64) to the one or two ranks whose slabs contain it, and a body
 * crossing a cut could not run at all.  Instead the MARKERS are divided among all ranks and the window fields
 * (velocity, per-iteration work fields, force) exist once per rank in peer-mapped symmetric memory:
 *   1. every rank computes the window velocity on the window cells of ITS slab and stores it into the copy of every
 *      rank whose markers can touch that cell (need box);
 *   2. every iteration, a rank interpolates at its markers from its own copy and adds what it spreads into the copy of
 *      every rank whose need box contains the cell (red.global.add over NVLink);
 *   3. the last iteration accumulates the force field of a rank's markers in its own copy; the rank then stores the
 *      cells of its need box into a staging slot [source rank] of the rank whose SLAB contains the cell (plain
 *      coalesced stores, every cell every step, so nothing has to be cleared), and that rank adds the slots into the
 *      force field its fluid kernel reads.  Every rank also stores its partial force (and torque) sum into a slot of
 *      every rank; all ranks then add the slots in rank order -- the "small all-reduce of total IB force and torque"
 *      of the reference's design -- and advance identical replicas of the rigid-body state;
 *   Only the window cells some marker's stencil can reach (`cells`, a static list in window-local coordinates: the
 *   body is rigid and the window follows it) are ever computed, sent or added;
 *   4. the steps are separated by an all-to-all flag barrier (one word per rank pair, system-scope release/acquire).
 * Everything is stream-ordered device code: graph-capturable, no NCCL call, no host synchronisation.
 * Coordinates of markers, window origin and body state are GLOBAL here; args->grid is the local slab and
 * args->win_shift maps global to local x. */
enum { VSB_MAX_RANKS = 8 };
typedef struct {
  int n_ranks, rank;
  float* fields[VSB_MAX_RANKS];      /* base of the window-field block of every rank (own entry = local address), laid out
                                        as [parity 2][slot n_iter + 1][window cells][2 | 4]: slot 0 force field,
                                        slots 1 .. n_iter-1 work fields, slot n_iter velocity                     */
  uint32_t* flags[VSB_MAX_RANKS];    /* VSB_MAX_RANKS words per rank: flags[r][src] = last barrier number published by src */
  float* sums[VSB_MAX_RANKS];        /* VSB_MAX_RANKS x 4 floats per rank: sums[r][src] = partial (Fx, Fy, Fz | torque, -)   */
  int need_lo[VSB_MAX_RANKS][3], need_hi[VSB_MAX_RANKS][3];   /* window-local box [lo, hi) of cells rank r's markers can touch */
  int x_lo[VSB_MAX_RANKS], x_hi[VSB_MAX_RANKS];               /* global x range [lo, hi) of the slab of rank r          */
  int64_t marker_begin, marker_end;  /* this rank's share of the marker arrays                                    */
  int chunk_begin, chunk_end;        /* ... and of mdf->chunk_offsets (tiled kernel); 0, 0: untiled                 */
  uint32_t* counter;                 /* 4 local device words: barrier number, spare, time-out indicator, spare      */
  float* staging[VSB_MAX_RANKS];     /* per rank: n_ranks window-sized fields [source rank][window cells][2 | 4]     */
  float* force_field;                /* LOCAL window-sized field the fluid kernels of this rank read (args->g_win)  */
  const int32_t* cells;              /* flat window-local indices of the cells a marker stencil can reach, ascending */
  int64_t n_cells;
  void* ev_window_done;              /* optional cudaEvent_t recorded on the chain's stream right after the window-velocity
                                        kernel: the caller can hold the bulk of the fluid pass back until then, so that the
                                        first, latency-critical kernel of the chain does not share the memory system   */
  uint64_t* trace;                   /* optional (NULL: off): 8192 device words; barrier number b stores %globaltimer at
                                        entry and exit into words 2 (b mod 4096) and 2 (b mod 4096) + 1 (timing aid)   */
} VsbIbShard;

/* One whole sharded chain of a step: window velocity -> barrier -> n_iter x (iteration -> barrier) -> force / torque
 * sums and body update (params->n_dof > 0: device ODE on every rank's replica) -> shard->force_field.
 * mdf->g_win / scratch / u_win are ignored: the fields live in shard->fields.  mdf->marker_u / marker_force are
 * filled for this rank's share only. */
int vsb_ibshard_chain(const VsbStepArgs* args, const VsbMdfArgs* mdf, const VsbIbShard* shard,
                      const VsbBodyParams* params, vsb_stream_t stream);
/* The flag barrier alone (used between set-up phases and by tests). */
int vsb_ibshard_barrier(const VsbIbShard* shard, vsb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VIVSIM_B200_H */
