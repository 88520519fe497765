// CPU harness for vivsim_b200/csrc/vsb_mrt_moment.cuh (host + device header): compiled by g++ in
// tests/test_host_logic_cpu.py and driven through ctypes.
#include "../../vivsim_b200/csrc/vsb_mrt_moment.cuh"

static const int kC[19][3] = {{0, 0, 0},  {1, 0, 0},  {-1, 0, 0},  {0, 1, 0},  {0, -1, 0}, {0, 0, 1},  {0, 0, -1},
                              {1, 1, 0},  {-1, 1, 0}, {1, -1, 0},  {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1},
                              {-1, 0, -1}, {0, 1, 1}, {0, -1, 1},  {0, 1, -1},  {0, -1, -1}};
static const int kPairQ[9] = {1, 3, 5, 7, 8, 11, 12, 15, 16};
static const int kOpp[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};

extern "C" int mm_make(const float* A, const float* B, float* s_out) {
  vsb::MomentOp3 op;
  if (!vsb::make_moment_op3(A, B, kC, op)) return 0;
  for (int i = 0; i < 19; ++i) s_out[i] = op.s[i];
  return 1;
}

extern "C" void mm_apply(const float* s, const float* x, float* y) {
  vsb::MomentOp3 op;
  for (int i = 0; i < 19; ++i) op.s[i] = s[i];
  float b[9], a[9], yb[9], ya[9], y0;
  for (int k = 0; k < 9; ++k) {
    b[k] = x[kPairQ[k]] + x[kOpp[kPairQ[k]]];
    a[k] = x[kPairQ[k]] - x[kOpp[kPairQ[k]]];
  }
  vsb::moment_op3_apply(op, b, a, y0, yb, ya);
  y[0] = y0;
  for (int k = 0; k < 9; ++k) {
    y[kPairQ[k]] = yb[k] + ya[k];
    y[kOpp[kPairQ[k]]] = yb[k] - ya[k];
  }
}
