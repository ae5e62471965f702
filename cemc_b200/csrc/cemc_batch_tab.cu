// batch_kernel instantiations with EV_TAB evaluation (own translation unit: parallel build)
#include "cemc_batch_launch.cuh"

namespace cemc {
int batch_launch_tab(const BatchLaunch &L) { return batch_launch_ev<EV_TAB, true>(L); }
}  // namespace cemc
