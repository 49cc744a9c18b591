// k_sim for trees with up to 64 actions (2 register chunks per lane): see tz_sim.cuh
#include "tz_sim.cuh"

namespace tz_internal {
int launch_sim_nc2(const SimLaunch& L, cudaStream_t s) { return launch_sim_nc<2>(L, s); }
}  // namespace tz_internal
