/*
 * tz_synth.h -- synthetic stand-in for the user's pgx environment + policy/value network.
 *
 * pgx and the user's JAX network cannot run in this image, so benchmarks and
 * parity tests drive the search with a deterministic hash game that has the SHAPES of the named pgx
 * games (branching factor, embedding bytes, legal-move density, episode length) and costs ~nothing.
 * It plays the role of env_step_fn (core/types.py:27), env_init_fn (core/types.py:28) and eval_fn
 * (core/types.py:31) plus the host-side glue of MCTS.iterate (core/evaluators/mcts/mcts.py:165-172)
 * and _AlphaZero.update_root (core/evaluators/alphazero.py:57-76).  It is NOT part of the product
 * path: libtz_synth.so is separate from libtz_b200.so and only bench.py / tests / smoke load it.
 *
 * The inline game definition below is shared by the device kernels (tz_synth.cu) and the C oracle's
 * driver so both are fed identical inputs; oracle/synth_numpy.py restates it independently.
 *
 * Node embedding (two pytree leaves):  core int32[4] = {h, depth, player, 0};  payload uint8[P] with
 * little-endian word w = mix32(h ^ (w+1)*0x9E3779B9) so that every row move can be verified.
 */
#ifndef TZ_SYNTH_H_
#define TZ_SYNTH_H_

#include "tz_abi.h"
#include "tz_math.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TzSynthGame {
  int32_t F;             /* branching factor */
  int32_t payload_bytes; /* P: bytes of the second embedding leaf (0 = leaf absent) */
  int32_t rho256;        /* an action a > 0 is legal with probability rho256/256; action 0 always legal */
  int32_t tau1024;       /* a state is terminal with probability tau1024/1024 ... */
  int32_t max_depth;     /* ... or when depth >= max_depth */
  uint32_t seed;
} TzSynthGame;

TZ_HD uint32_t tz_synth_init_h(uint32_t seed, uint32_t env, uint32_t episode) {
  return tz_mix32(seed * 0x9e3779b9u + env * 0x85ebca6bu + episode * 0xc2b2ae35u + 1u);
}
TZ_HD uint32_t tz_synth_step_h(uint32_t h, uint32_t a) { return tz_mix32(h * 0x9e3779b1u + a + 1u); }
TZ_HD int tz_synth_legal(uint32_t h, uint32_t a, int rho256) {
  return a == 0u || (int)(tz_mix32(h ^ ((a + 1u) * 0x85ebca6bu)) & 0xffu) < rho256;
}
TZ_HD float tz_synth_logit(uint32_t h, uint32_t a) { /* in [-2, 2), exact in fp32 */
  uint32_t u = tz_mix32(h ^ (a * 0x27d4eb2fu + 0x9e3779b9u)) >> 8;
  return (float)u * (1.0f / 4194304.0f) - 2.0f;
}
TZ_HD int tz_synth_terminal(uint32_t h, int depth, int tau1024, int max_depth) {
  return depth >= max_depth || (int)(tz_mix32(h ^ 0xc2b2ae35u) & 0x3ffu) < tau1024;
}
TZ_HD float tz_synth_reward(uint32_t h) { return (float)((int)(tz_mix32(h ^ 0x165667b1u) % 3u) - 1); }
TZ_HD float tz_synth_value(uint32_t h) { /* in [-1, 1), exact in fp32 */
  return (float)(tz_mix32(h ^ 0x27d4eb2fu) >> 8) * (1.0f / 8388608.0f) - 1.0f;
}
TZ_HD uint32_t tz_synth_payload_word(uint32_t h, uint32_t w) { return tz_mix32(h ^ ((w + 1u) * 0x9e3779b9u)); }

/* ---- device entry points (libtz_synth.so); all pointers are device pointers ---- */

/* env_init_fn for B envs: episode[b] selects the start position.  core [B,4] int32, payload [B,P]. */
int tz_synth_init_states(const TzSynthGame* g, int B, int env_offset, const int32_t* episode,
                         int32_t* core, uint8_t* payload, tz_stream_t stream);

/* Root evaluation: plain MCTS.update_root (mcts.py:137-138: unmasked softmax) when dir_noise == NULL,
 * else _AlphaZero.update_root (alphazero.py:57-76) with dir_noise [B,F] ~ Dirichlet and mixing dir_eps. */
int tz_synth_root(const TzSynthGame* g, int B, const int32_t* core, const float* dir_noise, float dir_eps,
                  float* root_policy, float* root_value, tz_stream_t stream);

/* One simulation's env_step_fn + eval_fn + mask/softmax/terminal-value (mcts.py:165-172). */
int tz_synth_leaf(const TzSynthGame* g, int B, const int32_t* parent_core, const int32_t* action,
                  float* policy, float* value, uint8_t* terminated, int32_t* new_core,
                  uint8_t* new_payload, tz_stream_t stream);

/* The real environment step after a move (core/common.py:82-99): advances core/payload in place with
 * action[b]; where the new state is terminal, sets reset_flag[b] = 1, bumps episode[b] and re-inits. */
int tz_synth_env_step(const TzSynthGame* g, int B, int env_offset, const int32_t* action, int32_t* core,
                      uint8_t* payload, int32_t* episode, uint8_t* reset_flag, tz_stream_t stream);

/* A tz_leaf_fn (tz_abi.h) whose `user` is a `const TzSynthCtx*`; the embedding leaves of `w` must be
 * {core, payload}.  Lets tz_search run whole searches without returning to the host language. */
typedef struct TzSynthCtx {
  TzSynthGame game;
  int32_t B;
} TzSynthCtx;
int tz_synth_leaf_cb(void* user, int sim, const TzWork* w, tz_stream_t stream);

/* Bench instrumentation: a tz_leaf_fn like tz_synth_leaf_cb that also records CUDA events on the launching stream
 * around every search launch (tz_synth_timed_begin(S) first, tz_synth_timed_collect after a sync). */
int tz_synth_timed_begin(int n_sims);
int tz_synth_leaf_cb_timed(void* user, int sim, const TzWork* w, tz_stream_t stream);
int tz_synth_timed_collect(float* ms_out, float* leaf_ms_out);

uint64_t tz_synth_launch_count(void);

/* 1: tz_synth_leaf launches its kernel as a programmatic dependent launch (it waits for the preceding grid before it
 * signals its own dependents -- the form TzSearchCfg.programmatic requires of a leaf kernel).  Returns the old setting. */
int tz_synth_set_programmatic(int on);

/* Measurement record of the leaf stand-in's launches, the twin of TzWork.timeline (include/tz_abi.h): while `rows_dev`
 * (device uint64 [slots,4], slots a power of two, rows initialised to {~0, 0, 0, 0}) is set, launch number q of
 * tz_synth_leaf (tz_synth_leaf_seq() = the next q) records %globaltimer ns {first warp in, last warp past its wait, last
 * warp out, -} into row q mod slots.  NULL switches it off.  Launches captured into a CUDA graph keep their row. */
int tz_synth_set_timeline(uint64_t* rows_dev, int slots);
uint64_t tz_synth_leaf_seq(void);

#ifdef __cplusplus
}
#endif
#endif /* TZ_SYNTH_H_ */
