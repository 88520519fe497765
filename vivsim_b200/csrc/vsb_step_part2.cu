// Fused-step instantiations compiled in this translation unit: D3Q19 BGK, D3Q19 REG
// (see "build slicing" in vsb_step.cu).
#define VSB_STEP_PART 2
#include "vsb_step.cu"

namespace vsb {
template int step_impl<3, VSB_COLL_BGK>(const VsbStepArgs&, cudaStream_t);
template int edge_impl<3, VSB_COLL_BGK>(const VsbStepArgs&, cudaStream_t, bool, int*);
template int step_impl<3, VSB_COLL_REG>(const VsbStepArgs&, cudaStream_t);
template int edge_impl<3, VSB_COLL_REG>(const VsbStepArgs&, cudaStream_t, bool, int*);
}  // namespace vsb
