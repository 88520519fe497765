/* CPU restatement (C + OpenMP) of the reference's composed IB-LBM step -- TEST INFRASTRUCTURE ONLY.
 *
 * Used as (a) a fast oracle for long / large parity runs and (b) the CPU baseline that bench.py
 * times on the GPU box's host cores (jax is not installable, so the reference's own JAX-CPU
 * backend cannot be run).  It follows the reference's UNFUSED sequence of passes, one function per
 * reference operator, exactly as an example driver calls them
 * (examples/2d/vortex_induced_vibration.py:96-148, examples/3d/flow_past_sphere.py:139-171):
 *
 *   get_macroscopic -> [window, get_ib_stencil, multi_direct_forcing, newmark] -> (u += g/2rho)
 *   -> get_equilibrium + collision_{bgk,kbc,reg} -> forcing_{edm,guo_bgk} -> streaming
 *   -> inlet NEBB (left) -> outlet equilibrium (right)
 *
 * fp32 arithmetic throughout (REF_REAL = float, the build every parity test and the CPU baseline use).  Compiled a
 * second time with -DREF_REAL=double (libiblbm_ref64.so) it is the fp64 yardstick that tells how far two correct fp32
 * evaluations of the same recipe may drift apart (tests: force tolerance over long horizons).
 * Checked against the NumPy oracle by tests/test_oracle_cport.py.  Never linked into or called by the product. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REF_REAL
#define REF_REAL float
#endif
typedef REF_REAL real;
#define R(x) ((real)(x))
#define FABS(x) ((real)fabs((double)(x)))
#define SQRT(x) (sizeof(real) == 4 ? (real)sqrtf((float)(x)) : (real)sqrt((double)(x)))
#define FLOOR(x) ((real)floor((double)(x)))

typedef struct {
  int dim, nx, ny, nz;          /* nz = 1 for dim 2 */
  int collision;                /* 0 bgk, 1 mrt (operators below), 2 kbc, 3 reg */
  int forcing;                  /* 0 none, 1 edm, 2 guo */
  double omega;
  real g0[3];                  /* uniform body force */
  int n_markers, n_iter;
  const real* markers0;        /* (M, dim) */
  const real* ds;              /* (M) */
  real worg0[3];               /* window origin (real), size */
  int wsz[3];
  int moving;                   /* 1: 2-DOF Newmark body, window follows trunc(origin0 + d) */
  double body_m, body_k, body_c, body_added;
  real d[3], v[3], a[3], h[3]; /* body state (in/out) */
  int inlet_outlet;             /* 1: left NEBB(ux=u0) + right equilibrium(ux=u0) */
  real u0;
  real* marker_force;          /* (M, dim) out: +F */
  const real* mrt_op;          /* Q x Q collision operator A = M^-1 S M (collision 1) */
  const real* mrt_fop;         /* Q x Q source operator B = M^-1 (I - S/2) M (collision 1 + forcing 2) */
  int follow;                   /* window rule of a moving body: 1 trunc(origin0 + d) (2-D VIV example :104-105),
                                   2 clip(floor(origin0 + d), 0, N - size) (examples/3d/oscillating_cylinder.py:241-243) */
} RefSpec;

static const int C2[9][3] = {{0,0,0},{1,0,0},{0,1,0},{-1,0,0},{0,-1,0},{1,1,0},{-1,1,0},{-1,-1,0},{1,-1,0}};
static const int C3[19][3] = {{0,0,0},{1,0,0},{-1,0,0},{0,1,0},{0,-1,0},{0,0,1},{0,0,-1},{1,1,0},{-1,1,0},{1,-1,0},
  {-1,-1,0},{1,0,1},{-1,0,1},{1,0,-1},{-1,0,-1},{0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1}};
static const int OPP3[19] = {0,2,1,4,3,6,5,10,9,8,7,14,13,12,11,18,17,16,15};

static inline int Q(int dim) { return dim == 2 ? 9 : 19; }
static inline const int* CV(int dim, int q) { return dim == 2 ? C2[q] : C3[q]; }
static inline real W(int dim, int q) {
  if (dim == 2) return q == 0 ? R(4.0) / R(9.0) : (q < 5 ? R(1.0) / R(9.0) : R(1.0) / R(36.0));
  return q == 0 ? R(1.0) / R(3.0) : (q < 7 ? R(1.0) / R(18.0) : R(1.0) / R(36.0));
}

/* lbm/basic.py:107-110, lbm3d/basic.py:102-105 */
static void cell_moments(int dim, const real* f, real* rho, real* u) {
  const int q_n = Q(dim);
  real r = R(0.), m[3] = {R(0.), R(0.), R(0.)};
  for (int q = 0; q < q_n; ++q) {
    r += f[q];
    const int* c = CV(dim, q);
    for (int d = 0; d < dim; ++d) m[d] += (real)c[d] * f[q];
  }
  *rho = r;
  for (int d = 0; d < dim; ++d) u[d] = m[d] / r;
}

/* lbm/basic.py:132-135 */
static void cell_feq(int dim, real rho, const real* u, real* feq) {
  real usq = R(0.);
  for (int d = 0; d < dim; ++d) usq += u[d] * u[d];
  for (int q = 0; q < Q(dim); ++q) {
    const int* c = CV(dim, q);
    real cu = R(0.);
    for (int d = 0; d < dim; ++d) cu += (real)c[d] * u[d];
    feq[q] = rho * W(dim, q) * (R(1.0) + R(3.0) * cu + R(4.5) * cu * cu - R(1.5) * usq);
  }
}

/* lbm/collision/reg.py:23-47, lbm3d/collision/reg.py:10-41 */
static void cell_proj(int dim, const real* fneq, real* out) {
  real pi[3][3] = {{0}};
  const int q_n = Q(dim);
  for (int q = 0; q < q_n; ++q) {
    const int* c = CV(dim, q);
    for (int a = 0; a < dim; ++a)
      for (int b = 0; b < dim; ++b) pi[a][b] += (real)(c[a] * c[b]) * fneq[q];
  }
  real tr = R(0.);
  for (int a = 0; a < dim; ++a) tr += pi[a][a];
  for (int q = 0; q < q_n; ++q) {
    const int* c = CV(dim, q);
    real s = R(0.);
    for (int a = 0; a < dim; ++a)
      for (int b = 0; b < dim; ++b) s += (real)(c[a] * c[b]) * pi[a][b];
    out[q] = W(dim, q) * R(4.5) * (s - tr * (R(1.0) / R(3.0)));
  }
}

static void cell_collide(const RefSpec* s, real* f, const real* feq) {
  const int dim = s->dim, q_n = Q(dim);
  const real om = (real)s->omega;
  if (s->collision == 1) { /* f + A (feq - f): lbm/collision/mrt.py:88, lbm3d/collision/mrt.py:96-98 */
    real dq[19], out[19];
    for (int q = 0; q < q_n; ++q) dq[q] = feq[q] - f[q];
    for (int q = 0; q < q_n; ++q) {
      real acc = R(0.);
      for (int k = 0; k < q_n; ++k) acc += s->mrt_op[q * q_n + k] * dq[k];
      out[q] = f[q] + acc;
    }
    for (int q = 0; q < q_n; ++q) f[q] = out[q];
    return;
  }
  if (s->collision == 0) { /* lbm/basic.py:156 */
    const real a = (real)(1.0 - s->omega);
    for (int q = 0; q < q_n; ++q) f[q] = a * f[q] + om * feq[q];
    return;
  }
  real fneq[19], sh[19];
  for (int q = 0; q < q_n; ++q) fneq[q] = f[q] - feq[q];
  if (s->collision == 3) { /* lbm/collision/reg.py:49 */
    cell_proj(dim, fneq, sh);
    const real a = (real)(1.0 - s->omega);
    for (int q = 0; q < q_n; ++q) f[q] = feq[q] + a * sh[q];
    return;
  }
  /* KBC: lbm/collision/kbc.py:29-59, lbm3d/collision/kbc.py:29-42 */
  if (dim == 2) {
    const real n4 = (fneq[1] - fneq[2] + fneq[3] - fneq[4]) / R(4.0);
    const real p4 = (fneq[5] - fneq[6] + fneq[7] - fneq[8]) / R(4.0);
    sh[0] = R(0.); sh[1] = n4; sh[2] = -n4; sh[3] = n4; sh[4] = -n4; sh[5] = p4; sh[6] = -p4; sh[7] = p4; sh[8] = -p4;
  } else {
    cell_proj(dim, fneq, sh);
  }
  real ssh = R(0.), shh = R(0.);
  for (int q = 0; q < q_n; ++q) {
    const real hi = fneq[q] - sh[q], inv = R(1.0) / (feq[q] + R(1e-20));
    ssh += hi * sh[q] * inv;
    shh += hi * hi * inv;
  }
  const real iw = (real)(1.0 / s->omega), omi = (real)(1.0 - 1.0 / s->omega);
  const real hg = iw - omi * ssh / (shh + R(1e-20));
  for (int q = 0; q < q_n; ++q) f[q] -= om * (sh[q] + hg * (fneq[q] - sh[q]));
}

/* lbm/forcing/guo.py:21-33 */
static void cell_guo(int dim, const real* g, const real* u, real* G) {
  real ug = R(0.);
  for (int d = 0; d < dim; ++d) ug += u[d] * g[d];
  for (int q = 0; q < Q(dim); ++q) {
    const int* c = CV(dim, q);
    real cu = R(0.), cg = R(0.);
    for (int d = 0; d < dim; ++d) { cu += (real)c[d] * u[d]; cg += (real)c[d] * g[d]; }
    G[q] = W(dim, q) * (R(3.0) * (cg - ug) + R(9.0) * cu * cg);
  }
}

/* ib/kernels.py:25-43 */
static real peskin4(real r) {
  const real a = FABS(r);
  if (a > R(2.0)) return R(0.);
  if (a < R(1.0)) return (R(3.0) - R(2.0) * a + SQRT(R(1.0) + R(4.0) * a - R(4.0) * a * a)) * R(0.125);
  return (R(5.0) - R(2.0) * a - SQRT(-R(7.0) + R(12.0) * a - R(4.0) * a * a)) * R(0.125);
}

typedef struct { real w[64]; int idx[64]; } Stencil;

/* ib/stencil.py:27-51, ib3d/stencil.py:36-59 (window-local coordinates) */
static void make_stencil(const RefSpec* s, const int* worg, const real* shift, int m, Stencil* st) {
  const int dim = s->dim, ns = dim == 2 ? 16 : 64;
  real x[3]; int base[3];
  for (int d = 0; d < dim; ++d) {
    x[d] = (s->markers0[m * dim + d] + shift[d]) - (real)worg[d];
    base[d] = (int)FLOOR(x[d]);
  }
  for (int k = 0; k < ns; ++k) {
    int kk = k, node[3]; real w = R(1.);
    for (int d = dim - 1; d >= 0; --d) { node[d] = base[d] + (kk & 3) - 1; kk >>= 2; w *= peskin4((real)node[d] - x[d]); }
    st->w[k] = w;
    st->idx[k] = dim == 2 ? node[0] * s->wsz[1] + node[1] : (node[0] * s->wsz[1] + node[1]) * s->wsz[2] + node[2];
  }
}

static void interp(int dim, int wcells, const real* grid, const Stencil* st, real scale, real* out) {
  const int ns = dim == 2 ? 16 : 64;
  for (int c = 0; c < dim; ++c) {
    real acc = R(0.);
    for (int k = 0; k < ns; ++k) acc += st->w[k] * (grid[c * wcells + st->idx[k]] * scale);
    out[c] = acc;
  }
}

static void spread(int dim, int wcells, real* grid, const Stencil* st, const real* val) {
  const int ns = dim == 2 ? 16 : 64;
  for (int c = 0; c < dim; ++c)
    for (int k = 0; k < ns; ++k) grid[c * wcells + st->idx[k]] += val[c] * st->w[k];
}

int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Advance n_steps reference time steps in place (f holds F_n, the reference's carried state).
 * work: f_tmp (Q*ncell), rho (ncell), u (dim*ncell). */
int ref_run(RefSpec* s, real* f, real* f_tmp, real* rho, real* u, int n_steps, int n_threads) {
  const int dim = s->dim, q_n = Q(dim);
  const int nx = s->nx, ny = s->ny, nz = dim == 3 ? s->nz : 1;
  const long ncell = (long)nx * ny * nz;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  int wcells = 1;
  for (int d = 0; d < dim; ++d) wcells *= s->wsz[d];
  const int M = s->n_markers;
  real* uw = NULL; real* gw = NULL; real* tmpw = NULL; real* um = NULL; real* Ft = NULL; Stencil* sten = NULL;
  if (M > 0) {
    uw = malloc(sizeof(real) * dim * wcells); gw = malloc(sizeof(real) * dim * wcells);
    tmpw = malloc(sizeof(real) * dim * wcells); um = malloc(sizeof(real) * M * dim);
    Ft = malloc(sizeof(real) * M * dim); sten = malloc(sizeof(Stencil) * M);
  }
  for (int step = 0; step < n_steps; ++step) {
    /* ---- get_macroscopic on the whole grid */
#pragma omp parallel for schedule(static)
    for (long i = 0; i < ncell; ++i) {
      real fl[19], uu[3];
      for (int q = 0; q < q_n; ++q) fl[q] = f[q * ncell + i];
      cell_moments(dim, fl, &rho[i], uu);
      for (int d = 0; d < dim; ++d) u[d * ncell + i] = uu[d];
    }
    /* ---- immersed boundary on the window (ib/mdf.py:31-64) */
    int worg[3] = {0, 0, 0};
    if (M > 0) {
      real shift[3] = {R(0.), R(0.), R(0.)};
      for (int d = 0; d < dim; ++d) {
        real o = s->worg0[d];
        if (s->moving && d < 2) { o = o + s->d[d]; shift[d] = s->d[d]; }
        if (s->follow == 2) {   /* clip(floor(.)): oscillating_cylinder.py:241-243 */
          const int nd = d == 0 ? nx : (d == 1 ? ny : nz);
          int oi = (int)FLOOR(o);
          oi = oi < 0 ? 0 : oi;
          worg[d] = oi > nd - s->wsz[d] ? nd - s->wsz[d] : oi;
        } else {
          worg[d] = (int)o;   /* astype(int32): truncation (vortex_induced_vibration.py:104-105) */
        }
      }
      for (int c = 0; c < dim; ++c)
        for (int ix = 0; ix < s->wsz[0]; ++ix)
          for (int iy = 0; iy < s->wsz[1]; ++iy)
            for (int iz = 0; iz < (dim == 3 ? s->wsz[2] : 1); ++iz) {
              const long gi = dim == 2 ? (long)(worg[0] + ix) * ny + (worg[1] + iy)
                                       : ((long)(worg[0] + ix) * ny + (worg[1] + iy)) * nz + (worg[2] + iz);
              const int wi = dim == 2 ? ix * s->wsz[1] + iy : (ix * s->wsz[1] + iy) * s->wsz[2] + iz;
              uw[c * wcells + wi] = u[c * ncell + gi];
            }
#pragma omp parallel for schedule(static)
      for (int m = 0; m < M; ++m) {
        make_stencil(s, worg, shift, m, &sten[m]);
        interp(dim, wcells, uw, &sten[m], R(1.0), &um[m * dim]);
        for (int c = 0; c < dim; ++c) Ft[m * dim + c] = R(0.);
      }
      for (int it = 0; it < s->n_iter; ++it) {
        memset(tmpw, 0, sizeof(real) * dim * wcells);
        for (int m = 0; m < M; ++m) {   /* scatter-add: serial, deterministic */
          real dF[3];
          for (int c = 0; c < dim; ++c) {
            const real tgt = (s->moving && c < 2) ? s->v[c] : R(0.);
            dF[c] = (tgt - um[m * dim + c]) * (s->ds[m] * R(2.0));
            Ft[m * dim + c] += dF[c];
          }
          spread(dim, wcells, tmpw, &sten[m], dF);
        }
#pragma omp parallel for schedule(static)
        for (int m = 0; m < M; ++m) {
          real du[3];
          interp(dim, wcells, tmpw, &sten[m], R(0.5), du);
          for (int c = 0; c < dim; ++c) um[m * dim + c] += du[c];
        }
      }
      memset(gw, 0, sizeof(real) * dim * wcells);
      real hsum[3] = {R(0.), R(0.), R(0.)};
      for (int m = 0; m < M; ++m) {
        spread(dim, wcells, gw, &sten[m], &Ft[m * dim]);
        for (int c = 0; c < dim; ++c) { hsum[c] += -Ft[m * dim + c]; if (s->marker_force) s->marker_force[m * dim + c] = Ft[m * dim + c]; }
      }
      if (s->moving) {   /* dyn.py:27-51 with gamma = 1/2, beta = 1/4, dt = 1 */
        const real denom = (real)(s->body_m + 0.5 * s->body_c + 0.25 * s->body_k);
        for (int c = 0; c < 2; ++c) {
          const real h = hsum[c] + s->a[c] * (real)s->body_added;
          const real v1 = s->v[c] + R(0.5) * s->a[c];
          const real d1 = s->d[c] + s->v[c] + R(0.25) * s->a[c];
          const real an = (h - (real)s->body_c * v1 - (real)s->body_k * d1) / denom;
          s->h[c] = h; s->a[c] = an; s->v[c] = R(0.5) * an + v1; s->d[c] = R(0.25) * an + d1;
        }
      }
    }
    /* ---- equilibrium + collision + forcing, in place */
#pragma omp parallel for schedule(static)
    for (long i = 0; i < ncell; ++i) {
      real fl[19], feq[19], G[19], uu[3], g[3];
      for (int q = 0; q < q_n; ++q) fl[q] = f[q * ncell + i];
      for (int d = 0; d < dim; ++d) { uu[d] = u[d * ncell + i]; g[d] = s->g0[d]; }
      if (M > 0) {
        int c3[3]; long r = i;
        c3[2] = (int)(r % nz); r /= nz; c3[1] = (int)(r % ny); c3[0] = (int)(r / ny);
        int rel[3], inside = 1;
        if (dim == 2) { rel[0] = c3[0] - worg[0]; rel[1] = c3[1] - worg[1]; rel[2] = 0; }
        else { rel[0] = c3[0] - worg[0]; rel[1] = c3[1] - worg[1]; rel[2] = c3[2] - worg[2]; }
        for (int d = 0; d < dim; ++d) inside = inside && rel[d] >= 0 && rel[d] < s->wsz[d];
        if (inside) {
          const int wi = dim == 2 ? rel[0] * s->wsz[1] + rel[1] : (rel[0] * s->wsz[1] + rel[1]) * s->wsz[2] + rel[2];
          for (int d = 0; d < dim; ++d) g[d] += gw[d * wcells + wi];
        }
      }
      if (s->forcing == 2)
        for (int d = 0; d < dim; ++d) uu[d] += g[d] * R(0.5) / rho[i];
      cell_feq(dim, rho[i], uu, feq);
      cell_collide(s, fl, feq);
      if (s->forcing) {
        cell_guo(dim, g, uu, G);
        if (s->forcing == 2 && s->collision == 1) {   /* f + B G: lbm/forcing/guo.py:106-107, lbm3d/forcing/guo.py:84-87 */
          for (int q = 0; q < q_n; ++q) {
            real acc = R(0.);
            for (int k = 0; k < q_n; ++k) acc += s->mrt_fop[q * q_n + k] * G[k];
            fl[q] += acc;
          }
        } else {
          const real sc = s->forcing == 2 ? (real)(1.0 - 0.5 * s->omega) : R(1.0);
          for (int q = 0; q < q_n; ++q) fl[q] += G[q] * sc;
        }
      }
      for (int q = 0; q < q_n; ++q) f[q * ncell + i] = fl[q];
    }
    /* ---- streaming (periodic push; lbm/basic.py:60-85) */
#pragma omp parallel for schedule(static) collapse(2)
    for (int q = 0; q < q_n; ++q)
      for (int x = 0; x < nx; ++x) {
        const int* c = CV(dim, q);
        const int sx = (x - c[0] + nx) % nx;
        for (int y = 0; y < ny; ++y) {
          const int sy = (y - c[1] + ny) % ny;
          real* dst = f_tmp + q * ncell + ((long)x * ny + y) * nz;
          const real* src = f + q * ncell + ((long)sx * ny + sy) * nz;
          if (dim == 2 || c[2] == 0) { for (int z = 0; z < nz; ++z) dst[z] = src[z]; }
          else { for (int z = 0; z < nz; ++z) dst[z] = src[(z - c[2] + nz) % nz]; }
        }
      }
    /* ---- boundary conditions on the x faces */
    if (s->inlet_outlet) {
      const long nface = (long)ny * nz;
#pragma omp parallel for schedule(static)
      for (long k = 0; k < nface; ++k) {
        const real u0 = s->u0;
        if (dim == 2) {   /* boundary_force_corrected_nebb(left, ux=u0) with g_wall = 0, rho_wall = 1: lbm/boundary/nebb.py:41-58 */
          const long cw = k;   /* x = 0 */
          const real f2 = f_tmp[2 * ncell + cw], f4 = f_tmp[4 * ncell + cw];
          const real shear = R(0.5) * (f2 - f4);
          f_tmp[1 * ncell + cw] = f_tmp[3 * ncell + cw] + (R(2.0) / R(3.0)) * u0;
          f_tmp[5 * ncell + cw] = f_tmp[7 * ncell + cw] - shear + (R(1.0) / R(6.0)) * u0;
          f_tmp[8 * ncell + cw] = f_tmp[6 * ncell + cw] + shear + (R(1.0) / R(6.0)) * u0;
        } else {          /* lbm3d/boundary/nebb.py:16-32 */
          real uw3[3] = {u0, R(0.), R(0.)}, fe[19];
          cell_feq(3, R(1.0), uw3, fe);
          for (int q = 0; q < 19; ++q)
            if (C3[q][0] > 0) f_tmp[q * ncell + k] = f_tmp[OPP3[q] * ncell + k] + fe[q] - fe[OPP3[q]];
        }
        /* boundary_equilibrium(right, ux=u0): lbm/boundary/eq.py:45-56 */
        real uwr[3] = {u0, R(0.), R(0.)}, fe[19];
        cell_feq(dim, R(1.0), uwr, fe);
        const long cr = (long)(nx - 1) * nface + k;
        for (int q = 0; q < q_n; ++q) f_tmp[q * ncell + cr] = fe[q];
      }
    }
    memcpy(f, f_tmp, sizeof(real) * q_n * ncell);
  }
  free(uw); free(gw); free(tmpw); free(um); free(Ft); free(sten);
  return 0;
}
