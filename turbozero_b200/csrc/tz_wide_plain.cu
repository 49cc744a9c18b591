// k_sim_wide, MCTS.backpropagate (plain backup): see tz_wide.cuh
#include "tz_wide.cuh"

namespace tz_internal {
int launch_wide_plain(const SimLaunch& L, int nc, int W, cudaStream_t s) { return launch_wide_any<false>(L, nc, W, s); }
}  // namespace tz_internal
