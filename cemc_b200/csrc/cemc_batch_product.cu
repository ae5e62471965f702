// batch_kernel instantiations with EV_PRODUCT evaluation (own translation unit: parallel build)
#include "cemc_batch_launch.cuh"

namespace cemc {
int batch_launch_product(const BatchLaunch &L) { return batch_launch_ev<EV_PRODUCT, true>(L); }
}  // namespace cemc
