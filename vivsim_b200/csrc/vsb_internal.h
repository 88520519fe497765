// Internal (non-ABI) declarations shared between the .cu translation units.
#pragma once

#include <cuda_runtime.h>

#include "../../include/vivsim_b200.h"

namespace vsb {

// Map a VsbGrid onto the three array axes (2-D grids get a unit leading axis).
inline void grid_axes(const VsbGrid& g, int& n0, int& n1, int& n2) {
  if (g.dim == 2) { n0 = 1; n1 = g.nx; n2 = g.ny; }
  else { n0 = g.nx; n1 = g.ny; n2 = g.nz; }
}

// One post-streaming operation, in place on f (vsb_boundary.cu).
// rows [r_begin, r_end) of the slowest real axis are the physical domain (r_end = 0: whole extent).
int launch_post_op(int dim, int n0, int n1, int n2, const VsbPostOp& op, const float* f_pre, float* f, cudaStream_t s,
                   int r_begin, int r_end);

}  // namespace vsb
