// k_sim_wide, WeightedMCTS.backpropagate (softmax-weighted backup): see tz_wide.cuh
#include "tz_wide.cuh"

namespace tz_internal {
int launch_wide_weighted(const SimLaunch& L, int nc, int W, cudaStream_t s) { return launch_wide_any<true>(L, nc, W, s); }
}  // namespace tz_internal
