// D3Q19 MRT collision evaluated in MOMENT SPACE (host + device, no intrinsics: also compiled by g++ for the CPU tests).
//
// The reference applies a 19 x 19 operator A = M^-1 S M to feq - f (lbm3d/collision/mrt.py:7-30 basis M, :50-72 rates,
// :96-98 product).  When the operator handed to the library really is diagonal in that basis -- checked numerically
// on the host, the API still takes any matrix -- the product needs no matrix at all.  The rows of M are monomials of the
// lattice velocity c,
//    0 1 | 1-3 c_x c_y c_z | 4 c.c | 5 2c_x^2 - c_y^2 - c_z^2 | 6 c_y^2 - c_z^2 | 7-9 c_x c_y, c_x c_z, c_y c_z
//    10 c_x^2 c_y | 11 c_x^2 c_z | 12 c_x c_y^2 | 13 c_y^2 c_z | 14 c_x c_z^2 | 15 c_y c_z^2 | 16-18 c_x^2 c_y^2, c_x^2 c_z^2, c_y^2 c_z^2
// so in terms of the opposite-direction pairs (b = x_q + x_opp, a = x_q - x_opp) every moment is a sum of at most a
// handful of pair values and the basis inverts in closed form: ~110 flops instead of the 181 multiply-adds of the
// parity-split matrix form (vsb_common.cuh SplitOp) or the 361 of the dense product.
//
// Pair order (Pairs<3>): k = 0..8 <-> q = 1 (+x), 3 (+y), 5 (+z), 7 (+x+y), 8 (-x+y), 11 (+x+z), 12 (-x+z), 15 (+y+z), 16 (-y+z).
#pragma once

#include <cuda_runtime.h>

namespace vsb {

// Relaxation rates of the non-conserved moments (rows 4..18 of M); the conserved ones (rows 0..3) must be zero.
struct MomentOp3 {
  float s[19];
};

// y = A x for A = M^-1 diag(s) M with s[0..3] = 0, everything in pair form:
//   in : x0 (rest population), b[k] = x_q + x_opp, a[k] = x_q - x_opp
//   out: y0, yb[k], ya[k] with y_q = yb[k] + ya[k], y_opp = yb[k] - ya[k]
__host__ __device__ inline void moment_op3_apply(const MomentOp3& op, const float (&b)[9], const float (&a)[9], float& y0,
                                                 float (&yb)[9], float (&ya)[9]) {
  const float* s = op.s;
  // ---- even moments
  const float pxy = b[3] + b[4], pxz = b[5] + b[6], pyz = b[7] + b[8];       // m16, m17, m18
  const float sum_p = (pxy + pxz) + pyz;
  const float e = ((b[0] + b[1]) + b[2]) + 2.0f * sum_p;                      // m4
  const float m5 = ((2.0f * b[0] - b[1]) - b[2]) + ((pxy + pxz) - 2.0f * pyz);
  const float m6 = (b[1] - b[2]) + (pxy - pxz);
  // relaxed
  const float rxy = s[16] * pxy, rxz = s[17] * pxz, ryz = s[18] * pyz;
  const float r7 = s[7] * (b[3] - b[4]), r8 = s[8] * (b[5] - b[6]), r9 = s[9] * (b[7] - b[8]);
  const float sum_r = (rxy + rxz) + ryz;
  const float t = s[4] * e - 2.0f * sum_r;                                    // b1' + b3' + b5'
  const float r5 = s[5] * m5, r6 = s[6] * m6;
  // inverse of the basis; halves folded in because the outputs are half sums
  const float hb1 = (1.0f / 6.0f) * ((t + r5) - ((rxy + rxz) - 2.0f * ryz));  // b1' / 2
  const float hr = 0.5f * t - hb1;                                            // (b3' + b5') / 2
  const float hd = 0.5f * ((r6 - rxy) + rxz);                                 // (b3' - b5') / 2
  y0 = -(t + sum_r);
  yb[0] = hb1;
  yb[1] = 0.5f * (hr + hd);
  yb[2] = 0.5f * (hr - hd);
  yb[3] = 0.25f * (rxy + r7); yb[4] = 0.25f * (rxy - r7);
  yb[5] = 0.25f * (rxz + r8); yb[6] = 0.25f * (rxz - r8);
  yb[7] = 0.25f * (ryz + r9); yb[8] = 0.25f * (ryz - r9);
  // ---- odd moments (the momentum rows have rate zero, so they are never formed)
  const float r10 = s[10] * (a[3] + a[4]), r12 = s[12] * (a[3] - a[4]);       // c_x^2 c_y, c_x c_y^2
  const float r11 = s[11] * (a[5] + a[6]), r14 = s[14] * (a[5] - a[6]);       // c_x^2 c_z, c_x c_z^2
  const float r13 = s[13] * (a[7] + a[8]), r15 = s[15] * (a[7] - a[8]);       // c_y^2 c_z, c_y c_z^2
  ya[0] = -0.5f * (r12 + r14);
  ya[1] = -0.5f * (r10 + r15);
  ya[2] = -0.5f * (r11 + r13);
  ya[3] = 0.25f * (r10 + r12); ya[4] = 0.25f * (r10 - r12);
  ya[5] = 0.25f * (r11 + r14); ya[6] = 0.25f * (r11 - r14);
  ya[7] = 0.25f * (r13 + r15); ya[8] = 0.25f * (r13 - r15);
}

// ----------------------------------------------------------------------------- host side: recognise the operator
// The moment basis, row r evaluated on lattice velocity (cx, cy, cz).
inline double moment_row3(int r, int cx, int cy, int cz) {
  const int xx = cx * cx, yy = cy * cy, zz = cz * cz;
  switch (r) {
    case 0: return 1;
    case 1: return cx;
    case 2: return cy;
    case 3: return cz;
    case 4: return xx + yy + zz;
    case 5: return 2 * xx - yy - zz;
    case 6: return yy - zz;
    case 7: return cx * cy;
    case 8: return cx * cz;
    case 9: return cy * cz;
    case 10: return xx * cy;
    case 11: return xx * cz;
    case 12: return cx * yy;
    case 13: return yy * cz;
    case 14: return cx * zz;
    case 15: return cy * zz;
    case 16: return xx * yy;
    case 17: return xx * zz;
    default: return yy * zz;
  }
}

// True when `A` (19 x 19, row-major, directions numbered as the lattice table `c`[19][3]) equals M^-1 diag(s) M with
// s[0..3] = 0 to within tol * max|s|; fills op.s.  If `B` is given it must equal I - A / 2 (the Guo source operator of
// lbm3d/forcing/guo.py:60-75) to the same tolerance, so that  f + A (feq - f) + B G = f + G + A (feq - f - G / 2).
inline bool make_moment_op3(const float* A, const float* B, const int (*c)[3], MomentOp3& op, double tol = 2e-6) {
  constexpr int Q = 19;
  double M[Q][Q], W[Q][2 * Q];
  for (int r = 0; r < Q; ++r)
    for (int q = 0; q < Q; ++q) {
      M[r][q] = moment_row3(r, c[q][0], c[q][1], c[q][2]);
      W[r][q] = M[r][q];
      W[r][Q + q] = (r == q) ? 1.0 : 0.0;
    }
  for (int col = 0; col < Q; ++col) {          // Gauss-Jordan with partial pivoting: W = [M | I] -> [I | M^-1]
    int piv = col;
    for (int r = col + 1; r < Q; ++r)
      if ((W[r][col] < 0 ? -W[r][col] : W[r][col]) > (W[piv][col] < 0 ? -W[piv][col] : W[piv][col])) piv = r;
    if (W[piv][col] == 0.0) return false;
    for (int j = 0; j < 2 * Q; ++j) { const double tt = W[col][j]; W[col][j] = W[piv][j]; W[piv][j] = tt; }
    const double d = W[col][col];
    for (int j = 0; j < 2 * Q; ++j) W[col][j] /= d;
    for (int r = 0; r < Q; ++r)
      if (r != col) {
        const double fct = W[r][col];
        if (fct != 0.0)
          for (int j = 0; j < 2 * Q; ++j) W[r][j] -= fct * W[col][j];
      }
  }
  // D = M A M^-1
  double MA[Q][Q], D[Q][Q], big = 0.0, off = 0.0;
  for (int i = 0; i < Q; ++i)
    for (int j = 0; j < Q; ++j) {
      double acc = 0.0;
      for (int k = 0; k < Q; ++k) acc += M[i][k] * (double)A[k * Q + j];
      MA[i][j] = acc;
    }
  for (int i = 0; i < Q; ++i)
    for (int j = 0; j < Q; ++j) {
      double acc = 0.0;
      for (int k = 0; k < Q; ++k) acc += MA[i][k] * W[k][Q + j];
      D[i][j] = acc;
      const double m = acc < 0 ? -acc : acc;
      if (i == j) big = m > big ? m : big;
      else off = m > off ? m : off;
    }
  if (big == 0.0 || off > tol * big) return false;
  for (int i = 0; i < 4; ++i)
    if ((D[i][i] < 0 ? -D[i][i] : D[i][i]) > tol * big) return false;
  for (int i = 0; i < Q; ++i) op.s[i] = (i < 4) ? 0.f : (float)D[i][i];
  if (B) {
    double bad = 0.0;
    for (int i = 0; i < Q; ++i)
      for (int j = 0; j < Q; ++j) {
        const double want = (i == j ? 1.0 : 0.0) - 0.5 * (double)A[i * Q + j];
        const double d = (double)B[i * Q + j] - want;
        bad = (d < 0 ? -d : d) > bad ? (d < 0 ? -d : d) : bad;
      }
    if (bad > tol * (big > 1.0 ? big : 1.0)) return false;
  }
  return true;
}

}  // namespace vsb
