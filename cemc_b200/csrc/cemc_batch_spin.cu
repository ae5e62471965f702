// batch_kernel instantiations with EV_SPIN evaluation (own translation unit: parallel build)
#include "cemc_batch_launch.cuh"

namespace cemc {
int batch_launch_spin(const BatchLaunch &L) { return batch_launch_ev<EV_SPIN, false>(L); }
}  // namespace cemc
