// tz_replay.cu -- sm_100a kernels + C-ABI (include/tz_replay.h) of the episode replay buffer behind the search:
// core/memory/replay_memory.py as driven by Trainer.collect (core/training/train.py:271-347).
//
// HBM-bound byte work: one warp per env, rows moved with coalesced 16-byte vectors; the three buffer updates of a
// collection step (add_experience x n, assign_rewards where terminated, truncate where truncated) are ONE launch, so the
// per-env flags are read once and the [cap] flag rows stay in registers between the phases.
#include <cuda_runtime.h>
#include <stdint.h>

#include "tz_math.h"
#include "tz_replay.h"

#include <atomic>

namespace tz_internal {
std::atomic<uint64_t> g_replay_launches{0};  // added to tz_launch_count() (tz_kernels.cu)
uint64_t replay_launches() { return g_replay_launches.load(std::memory_order_relaxed); }
}  // namespace tz_internal

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int THREADS = 128;  // 4 warps = 4 envs per CTA

__device__ __forceinline__ void warp_copy_row(uint8_t* dst, const uint8_t* src, int64_t bytes, int lane) {
  const uintptr_t a = (uintptr_t)dst | (uintptr_t)src | (uintptr_t)bytes;
  if ((a & 15) == 0) {
    for (int64_t i = lane; i < (bytes >> 4); i += 32) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  } else if ((a & 3) == 0) {
    for (int64_t i = lane; i < (bytes >> 2); i += 32) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (int64_t i = lane; i < bytes; i += 32) dst[i] = src[i];
  }
}

struct ExpPtrs {
  const uint8_t* p[4 * TZ_MAX_EMB];  // up to 4 experiences per step (the step itself + 3 transforms) per launch
};

__global__ void __launch_bounds__(THREADS) k_replay_collect(const TzReplay r, const int n_exp, const ExpPtrs ex,
                                                          const float* __restrict__ reward,
                                                          const uint8_t* __restrict__ terminated,
                                                          const uint8_t* __restrict__ truncated) {
  const int b = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= r.B) return;
  const int cap = r.capacity;
  int next = r.next_idx[b];
  int start = r.episode_start_idx[b];
  const bool term = terminated != nullptr && terminated[b] != 0;
  const bool trunc = truncated != nullptr && truncated[b] != 0;
  uint8_t* const pop = r.populated + (size_t)b * cap;
  uint8_t* const hasr = r.has_reward + (size_t)b * cap;
  // ---- add_experience (replay_memory.py:65-84), n_exp times ---------------------------------------------------
  for (int e = 0; e < n_exp; ++e) {
    const int idx = next;
    for (int k = 0; k < r.n_leaves; ++k) {
      const int64_t rb = r.leaf_row_bytes[k];
      warp_copy_row(reinterpret_cast<uint8_t*>(r.leaf[k]) + ((size_t)b * cap + (size_t)idx) * rb,
                    ex.p[e * r.n_leaves + k] + (size_t)b * rb, rb, lane);
    }
    if (lane == 0) {
      pop[idx] = 1;
      hasr[idx] = 0;
    }
    next = idx + 1 == cap ? 0 : idx + 1;  // (next_idx + 1) % capacity
  }
  __syncwarp();  // lane 0's flag writes above are read by the other lanes below
  // ---- assign_rewards where terminated (:87-107), truncate where truncated (:110-135) ----------------------------
  if (term || trunc) {
    float* const rew = reinterpret_cast<float*>(r.leaf[r.reward_leaf]) + (size_t)b * cap * r.reward_dim;
    for (int c = lane; c < cap; c += 32) {
      bool has = hasr[c] != 0;
      if (term && !has) {  // every slot without a reward, populated or not (:100-105)
        for (int p = 0; p < r.reward_dim; ++p) rew[(size_t)c * r.reward_dim + p] = reward[(size_t)b * r.reward_dim + p];
        has = true;
      }
      if (trunc && !has) pop[c] = 0;  // :130-134 (after an assign in the same step nothing is left to drop)
      hasr[c] = 1;                    // both end with has_reward = full_like(True)
    }
    if (term) start = next;   // :97
    if (trunc) next = start;  // :128 (sees the episode start an assign in the same step has just moved)
  }
  if (lane == 0) {
    r.next_idx[b] = next;
    r.episode_start_idx[b] = start;
  }
}

__global__ void __launch_bounds__(256) k_replay_count(const uint8_t* __restrict__ pop, const uint8_t* __restrict__ hasr, const size_t n,
                                                    int32_t* __restrict__ n_valid) {
  int local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    local += (pop[i] != 0 && hasr[i] != 0) ? 1 : 0;
  local = __reduce_add_sync(FULL, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_valid, local);
}

__global__ void __launch_bounds__(256) k_replay_scores(const uint8_t* __restrict__ pop, const uint8_t* __restrict__ hasr, const size_t n,
                                                     const float* __restrict__ gumbel, const int32_t* __restrict__ n_valid,
                                                     float* __restrict__ scores) {
  // p = w / sum(w) with w in {0, 1}: one value for every sampleable slot (replay_memory.py:157-169)
  const float logp = tz_logf(__fdiv_rn(1.0f, (float)*n_valid));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const bool ok = pop[i] != 0 && hasr[i] != 0;
    scores[i] = ok ? __fsub_rn(-gumbel[i], logp) : INFINITY;  // -gumbel - log(0) = +inf
  }
}

struct OutPtrs {
  uint8_t* p[TZ_MAX_EMB];
};

__global__ void __launch_bounds__(THREADS) k_replay_gather(const TzReplay r, const int64_t* __restrict__ flat_index, const int n,
                                                         const OutPtrs out) {
  const int j = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (j >= n) return;
  int64_t f = flat_index[j];
  const int64_t total = (int64_t)r.B * r.capacity;
  f = f < 0 ? 0 : (f >= total ? total - 1 : f);  // clamp like an XLA gather
  for (int k = 0; k < r.n_leaves; ++k) {
    const int64_t rb = r.leaf_row_bytes[k];
    warp_copy_row(out.p[k] + (size_t)j * rb, reinterpret_cast<const uint8_t*>(r.leaf[k]) + (size_t)f * rb, rb, lane);
  }
}

int check_replay(const TzReplay* r) {
  if (!r || r->B <= 0 || r->capacity <= 0 || r->n_leaves <= 0 || r->n_leaves > TZ_MAX_EMB) return TZ_EINVAL;
  if (!r->next_idx || !r->episode_start_idx || !r->populated || !r->has_reward) return TZ_EINVAL;
  if (r->reward_leaf < 0 || r->reward_leaf >= r->n_leaves || r->reward_dim <= 0) return TZ_EINVAL;
  for (int k = 0; k < r->n_leaves; ++k)
    if (!r->leaf[k] || r->leaf_row_bytes[k] <= 0) return TZ_EINVAL;
  if (r->leaf_row_bytes[r->reward_leaf] != 4 * (int64_t)r->reward_dim) return TZ_EINVAL;
  return TZ_OK;
}

inline int status(int launches = 1) {
  tz_internal::g_replay_launches.fetch_add((uint64_t)launches, std::memory_order_relaxed);
  const cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? TZ_OK : (int)e;
}

}  // namespace

extern "C" {

int tz_replay_init(const TzReplay* r, tz_stream_t stream) {
  const int rc = check_replay(r);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t B = (size_t)r->B, cap = (size_t)r->capacity;
  cudaError_t e = cudaMemsetAsync(r->next_idx, 0, B * 4, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(r->episode_start_idx, 0, B * 4, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(r->populated, 0, B * cap, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(r->has_reward, 1, B * cap, s);
  for (int k = 0; k < r->n_leaves && e == cudaSuccess; ++k) e = cudaMemsetAsync(r->leaf[k], 0, B * cap * (size_t)r->leaf_row_bytes[k], s);
  return e == cudaSuccess ? TZ_OK : (int)e;
}

int tz_replay_collect(const TzReplay* r, int n_exp, void* const* experiences, const float* reward, const uint8_t* terminated,
                      const uint8_t* truncated, tz_stream_t stream) {
  const int rc = check_replay(r);
  if (rc) return rc;
  if (n_exp < 0 || (n_exp > 0 && !experiences) || (terminated && !reward)) return TZ_EINVAL;
  if ((int64_t)n_exp * r->n_leaves > 4 * TZ_MAX_EMB) return TZ_ENOTSUP;
  ExpPtrs ex = {};
  for (int i = 0; i < n_exp * r->n_leaves; ++i) {
    if (!experiences[i]) return TZ_EINVAL;
    ex.p[i] = reinterpret_cast<const uint8_t*>(experiences[i]);
  }
  const int grid = (r->B * 32 + THREADS - 1) / THREADS;
  k_replay_collect<<<grid, THREADS, 0, (cudaStream_t)stream>>>(*r, n_exp, ex, reward, terminated, truncated);
  return status();
}

int tz_replay_count_valid(const TzReplay* r, int32_t* n_valid, tz_stream_t stream) {
  const int rc = check_replay(r);
  if (rc) return rc;
  if (!n_valid) return TZ_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)r->B * r->capacity;
  const cudaError_t e = cudaMemsetAsync(n_valid, 0, 4, s);
  if (e != cudaSuccess) return (int)e;
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_replay_count<<<grid, 256, 0, s>>>(r->populated, r->has_reward, n, n_valid);
  return status();
}

int tz_replay_sample_scores(const TzReplay* r, const float* gumbel, const int32_t* n_valid_total, float* scores,
                            tz_stream_t stream) {
  const int rc = check_replay(r);
  if (rc) return rc;
  if (!gumbel || !scores || !n_valid_total) return TZ_EINVAL;
  const size_t n = (size_t)r->B * r->capacity;
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_replay_scores<<<grid, 256, 0, (cudaStream_t)stream>>>(r->populated, r->has_reward, n, gumbel, n_valid_total, scores);
  return status();
}

int tz_replay_gather(const TzReplay* r, const int64_t* flat_index, int n, void* const* out, tz_stream_t stream) {
  const int rc = check_replay(r);
  if (rc) return rc;
  if (n < 0 || !flat_index || !out) return TZ_EINVAL;
  if (n == 0) return TZ_OK;
  OutPtrs o = {};
  for (int k = 0; k < r->n_leaves; ++k) {
    if (!out[k]) return TZ_EINVAL;
    o.p[k] = reinterpret_cast<uint8_t*>(out[k]);
  }
  const int grid = (n * 32 + THREADS - 1) / THREADS;
  k_replay_gather<<<grid, THREADS, 0, (cudaStream_t)stream>>>(*r, flat_index, n, o);
  return status();
}

}  // extern "C"
