// Fused-step instantiations compiled in this translation unit: D3Q19 MRT in moment space
#define VSB_STEP_PART 6
#include "vsb_step.cu"

namespace vsb {
template int step_impl<3, VSB_COLL_MRT_MOMENT>(const VsbStepArgs&, cudaStream_t);
template int edge_impl<3, VSB_COLL_MRT_MOMENT>(const VsbStepArgs&, cudaStream_t, bool, int*);
}  // namespace vsb
