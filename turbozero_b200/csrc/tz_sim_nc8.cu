// k_sim for trees with up to 256 actions (8 register chunks per lane): see tz_sim.cuh
#include "tz_sim.cuh"

namespace tz_internal {
int launch_sim_nc8(const SimLaunch& L, cudaStream_t s) { return launch_sim_nc<8>(L, s); }
}  // namespace tz_internal
