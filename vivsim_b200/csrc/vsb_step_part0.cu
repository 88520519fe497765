// Fused-step instantiations compiled in this translation unit: D2Q9 BGK, D2Q9 REG, D2Q9 KBC
// (see "build slicing" in vsb_step.cu).
#define VSB_STEP_PART 0
#include "vsb_step.cu"

namespace vsb {
template int step_impl<2, VSB_COLL_BGK>(const VsbStepArgs&, cudaStream_t);
template int edge_impl<2, VSB_COLL_BGK>(const VsbStepArgs&, cudaStream_t, bool, int*);
template int step_impl<2, VSB_COLL_REG>(const VsbStepArgs&, cudaStream_t);
template int edge_impl<2, VSB_COLL_REG>(const VsbStepArgs&, cudaStream_t, bool, int*);
template int step_impl<2, VSB_COLL_KBC>(const VsbStepArgs&, cudaStream_t);
template int edge_impl<2, VSB_COLL_KBC>(const VsbStepArgs&, cudaStream_t, bool, int*);
}  // namespace vsb
