// k_sim for trees with up to 512 actions (16 register chunks per lane): see tz_sim.cuh
#include "tz_sim.cuh"

namespace tz_internal {
int launch_sim_nc16(const SimLaunch& L, cudaStream_t s) { return launch_sim_nc<16>(L, s); }
}  // namespace tz_internal
