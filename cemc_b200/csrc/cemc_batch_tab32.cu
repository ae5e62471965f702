// batch_kernel instantiations with EV_TAB32 evaluation: the opt-in fp32 variant (single-precision
// product tables and sub-cluster sums).  Own translation unit: parallel build.
#include "cemc_batch_launch.cuh"

namespace cemc {
int batch_launch_tab32(const BatchLaunch &L) { return batch_launch_ev<EV_TAB32, true>(L); }
}  // namespace cemc
