// Fused-step instantiations compiled in this translation unit: D2Q9 MRT, D2Q9 MRT_SPLIT
// (see "build slicing" in vsb_step.cu).
#define VSB_STEP_PART 1
#include "vsb_step.cu"

namespace vsb {
template int step_impl<2, VSB_COLL_MRT>(const VsbStepArgs&, cudaStream_t);
template int edge_impl<2, VSB_COLL_MRT>(const VsbStepArgs&, cudaStream_t, bool, int*);
template int step_impl<2, VSB_COLL_MRT_SPLIT>(const VsbStepArgs&, cudaStream_t);
template int edge_impl<2, VSB_COLL_MRT_SPLIT>(const VsbStepArgs&, cudaStream_t, bool, int*);
}  // namespace vsb
