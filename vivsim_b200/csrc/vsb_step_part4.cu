// Fused-step instantiations compiled in this translation unit: D3Q19 MRT
// (see "build slicing" in vsb_step.cu).
#define VSB_STEP_PART 4
#include "vsb_step.cu"

namespace vsb {
template int step_impl<3, VSB_COLL_MRT>(const VsbStepArgs&, cudaStream_t);
template int edge_impl<3, VSB_COLL_MRT>(const VsbStepArgs&, cudaStream_t, bool, int*);
}  // namespace vsb
