/*
 * tz_replay.h -- C-ABI of the episode replay buffer that sits right behind the search in self-play
 * (SURVEY.md 8f rank 2): lowrollr/turbozero core/memory/replay_memory.py, as it is driven by
 * Trainer.collect (core/training/train.py:271-347).  Same conventions as tz_abi.h: device pointers, caller-owned
 * buffers, asynchronous on `stream`, in place, never allocates / synchronises / throws; returns TZ_OK, a negative TZ_E*
 * code or a positive cudaError_t.
 *
 * Layout = ReplayBufferState (replay_memory.py:24-41) with the leading env axis EpisodeReplayBuffer.init gives it
 * (:186-206): next_idx [B] i32, episode_start_idx [B] i32, populated [B,cap] bool, has_reward [B,cap] bool, and one
 * table [B,cap,row_k] per leaf of the experience pytree (BaseExperience, :8-21); bool is one byte.
 */
#ifndef TZ_REPLAY_H_
#define TZ_REPLAY_H_

#include "tz_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TzReplay {
  int32_t B;         /* envs on this device */
  int32_t capacity;  /* EpisodeReplayBuffer.capacity (replay_memory.py:50-57) */
  int32_t n_leaves;  /* leaves of the experience pytree, 1..TZ_MAX_EMB */
  int32_t reward_leaf; /* which leaf is BaseExperience.reward: float32[reward_dim] per row */
  int32_t reward_dim;  /* number of players */
  int32_t pad;
  int32_t* next_idx;          /* [B]      replay_memory.py:37 */
  int32_t* episode_start_idx; /* [B]      :38 */
  uint8_t* populated;         /* [B,cap]  :40 */
  uint8_t* has_reward;        /* [B,cap]  :41 */
  void* leaf[TZ_MAX_EMB];     /* [B,cap,leaf_row_bytes[k]]  :39 */
  int64_t leaf_row_bytes[TZ_MAX_EMB];
} TzReplay;

/* EpisodeReplayBuffer.init: replay_memory.py:186-206 (indices 0, buffer zeros, populated False, has_reward True). */
int tz_replay_init(const TzReplay* r, tz_stream_t stream);

/* The buffer half of Trainer.collect for every env (train.py:300-340), one launch:
 *   n_exp x add_experience (replay_memory.py:65-84): experiences[e * n_leaves + k] is leaf k of experience e, [B,row_k];
 *   then where terminated[b]: assign_rewards(reward[b,:]) (:87-107);  then where truncated[b]: truncate (:110-135).
 * n_exp may be 0; reward / terminated may be NULL together (no assign); truncated may be NULL (no truncate). */
int tz_replay_collect(const TzReplay* r, int n_exp, void* const* experiences, const float* reward,
                      const uint8_t* terminated, const uint8_t* truncated, tz_stream_t stream);

/* EpisodeReplayBuffer.sample, first half (replay_memory.py:157-169), for this device's [B,cap] block:
 * tz_replay_count_valid: n_valid[0] = number of slots with populated & has_reward (device int32[1]).
 * tz_replay_sample_scores: scores[i] = -gumbel[i] - log(p) with p = 1 / n_valid_total[0] for those slots -- the keys whose
 * `sample_size` smallest are what jax.random.choice(replace=False, p = mask / mask.sum()) draws -- and +inf elsewhere.
 * n_valid_total is the count over ALL devices the sample spans (the caller sums tz_replay_count_valid over ranks).
 * gumbel, scores: [B*cap] float32. */
int tz_replay_count_valid(const TzReplay* r, int32_t* n_valid, tz_stream_t stream);
int tz_replay_sample_scores(const TzReplay* r, const float* gumbel, const int32_t* n_valid_total, float* scores,
                            tz_stream_t stream);

/* Second half (replay_memory.py:171-181): out[k][j, :] = leaf[k][flat_index[j] / cap, flat_index[j] % cap, :]. */
int tz_replay_gather(const TzReplay* r, const int64_t* flat_index, int n, void* const* out, tz_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TZ_REPLAY_H_ */
