// k_sim for trees with up to 32 actions (1 register chunk per lane): see tz_sim.cuh
#include "tz_sim.cuh"

namespace tz_internal {
int launch_sim_nc1(const SimLaunch& L, cudaStream_t s) { return launch_sim_nc<1>(L, s); }
}  // namespace tz_internal
