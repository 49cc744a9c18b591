// k_sim for trees with up to 128 actions (4 register chunks per lane): see tz_sim.cuh
#include "tz_sim.cuh"

namespace tz_internal {
int launch_sim_nc4(const SimLaunch& L, cudaStream_t s) { return launch_sim_nc<4>(L, s); }
}  // namespace tz_internal
