// tz_synth.cu -- device-side synthetic game + "network" (standin/include/tz_synth.h): the stand-in for the user's pgx
// env_step_fn / env_init_fn and eval_fn (core/types.py:27-31) plus MCTS.iterate's host glue (mcts.py:165-172)
// and the root evaluation (mcts.py:137-138, alphazero.py:57-76).  Bench / test infrastructure only -- it is a
// separate library (libtz_synth.so) and nothing in libtz_b200.so depends on it.
//
// One warp per env; lane l owns actions l, l+32, ...  Softmax uses the path's canonical sum order and tz_expf,
// so its outputs are bit-identical to the C oracle's host version of the same game.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "tz_synth.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int THREADS = 64;
constexpr int MAX_NC = 16;

std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_programmatic{0};  // tz_synth_set_programmatic
// optional per-launch record of k_leaf in the product build (tz_synth_set_timeline), the twin of TzWork.timeline
std::atomic<unsigned long long*> g_tl_rows{nullptr};
std::atomic<uint32_t> g_tl_slots{0};
std::atomic<uint64_t> g_leaf_seq{0};
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__constant__ unsigned long long* c_tl_rows;

#ifdef TZ_PROFILE  // diagnostic build (libtz_synth_prof.so): per-launch timeline of k_leaf, see scripts/timeline.py
__device__ unsigned long long g_tl[4 * 1024];  // {first warp in, last warp past griddepcontrol.wait, last warp out, -}
__device__ __forceinline__ unsigned long long tl_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
std::atomic<uint32_t> g_tl_seq{0};
#define TZ_TL(op, slot, k) do { if ((threadIdx.x & 31) == 0) op(&g_tl[4 * (slot) + (k)], tl_now()); } while (0)
#else
#define TZ_TL(op, slot, k) do { } while (0)
#endif

__device__ __forceinline__ uint32_t fkey(float x) {
  uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
__device__ __forceinline__ float warp_max(float x) { return fkey_inv(__reduce_max_sync(FULL, fkey(x))); }
__device__ __forceinline__ float warp_canon_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(FULL, v, off));
  return v;
}

// in-place softmax over logits spread across the warp (lane l holds l, l+32, ...)
__device__ __forceinline__ void warp_softmax(float* x, int nc, int F, int lane) {
  float m = -INFINITY;
  for (int c = 0; c < nc; ++c)
    if (c * 32 + lane < F) m = fmaxf(m, x[c]);
  m = warp_max(m);
  float part = 0.0f;
  for (int c = 0; c < nc; ++c) {
    x[c] = (c * 32 + lane < F) ? tz_expf(__fsub_rn(x[c], m)) : 0.0f;
    part = __fadd_rn(part, x[c]);
  }
  const float s = warp_canon_sum(part);
  for (int c = 0; c < nc; ++c) x[c] = __fdiv_rn(x[c], s);
}

__device__ __forceinline__ void write_state(const TzSynthGame& g, uint32_t h, int depth, int player, int32_t* core,
                                            uint8_t* payload, int lane) {
  if (lane == 0) {
    core[0] = (int32_t)h;
    core[1] = depth;
    core[2] = player;
    core[3] = 0;
  }
  const int P = g.payload_bytes;
  if (P <= 0 || !payload) return;
  const int nfull = P >> 2;
  if (((uintptr_t)payload & 3) == 0) {
    for (int wd = lane; wd < nfull; wd += 32) reinterpret_cast<uint32_t*>(payload)[wd] = tz_synth_payload_word(h, (uint32_t)wd);
  } else {
    for (int wd = lane; wd < nfull; wd += 32) {
      const uint32_t v = tz_synth_payload_word(h, (uint32_t)wd);
      for (int k = 0; k < 4; ++k) payload[4 * wd + k] = (uint8_t)(v >> (8 * k));
    }
  }
  if (lane == 0 && (P & 3)) {
    const uint32_t v = tz_synth_payload_word(h, (uint32_t)nfull);
    for (int k = 0; k < (P & 3); ++k) payload[4 * nfull + k] = (uint8_t)(v >> (8 * k));
  }
}

__global__ void __launch_bounds__(THREADS) k_init_states(const TzSynthGame g, const int B, const int env_offset,
                                                       const int32_t* __restrict__ episode, int32_t* core, uint8_t* payload) {
  const int b = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const uint32_t h = tz_synth_init_h(g.seed, (uint32_t)(b + env_offset), (uint32_t)episode[b]);
  write_state(g, h, 0, 0, core + 4 * (size_t)b, payload ? payload + (size_t)b * g.payload_bytes : nullptr, lane);
}

__global__ void __launch_bounds__(THREADS) k_root(const TzSynthGame g, const int B, const int32_t* __restrict__ core,
                                                const float* __restrict__ dir_noise, const float dir_eps, const float one_minus_eps,
                                                float* __restrict__ root_policy, float* __restrict__ root_value) {
  const int b = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const int F = g.F, nc = (F + 31) >> 5;
  const uint32_t h = (uint32_t)core[4 * (size_t)b];
  float x[MAX_NC];
  for (int c = 0; c < nc; ++c) x[c] = tz_synth_logit(h, (uint32_t)(c * 32 + lane));
  warp_softmax(x, nc, F, lane);  // mcts.py:138 / alphazero.py:59: unmasked
  if (dir_noise) {               // alphazero.py:62-76
    for (int c = 0; c < nc; ++c) {
      const int a = c * 32 + lane;
      if (a < F) {
        float noisy = __fadd_rn(__fmul_rn(one_minus_eps, x[c]), __fmul_rn(dir_eps, dir_noise[(size_t)b * F + a]));
        noisy = fmaxf(noisy, TZ_FLT_MIN);
        x[c] = tz_synth_legal(h, (uint32_t)a, g.rho256) ? tz_logf(noisy) : -TZ_FLT_MAX;
      }
    }
    warp_softmax(x, nc, F, lane);
  }
  for (int c = 0; c < nc; ++c) {
    const int a = c * 32 + lane;
    if (a < F) root_policy[(size_t)b * F + a] = x[c];
  }
  if (lane == 0) root_value[b] = tz_synth_value(h);
}

__global__ void __launch_bounds__(THREADS) k_leaf(const TzSynthGame g, const int B, const int32_t* __restrict__ parent_core,
                                                const int32_t* __restrict__ action, float* __restrict__ policy,
                                                float* __restrict__ value, uint8_t* __restrict__ terminated, int32_t* new_core,
                                                uint8_t* new_payload, const int pdl_and_slot) {
  const int pdl = pdl_and_slot & 1;
  // optional launch record (tz_synth_set_timeline): bit 1 = on, the row index above it; the rows live behind a __constant__
  // pointer so that the kernel's parameter list is the same with and without the record
  unsigned long long* const tl_row = (pdl_and_slot & 2) ? c_tl_rows + 4 * (size_t)(pdl_and_slot >> 2) : nullptr;
  [[maybe_unused]] const int tl_slot = pdl_and_slot >> 2;  // diagnostic build: timeline slot of this launch
  const int b = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  TZ_TL(atomicMin, tl_slot, 0);
  if (tl_row && lane == 0) atomicMin(tl_row + 0, gtime_ns());
  const int F = g.F, nc = (F + 31) >> 5;
  const int32_t* pc = parent_core + 4 * (size_t)b;
  // Programmatic dependent launch, the form TzSearchCfg.programmatic asks of a leaf kernel: wait for the preceding
  // grid (the search kernel that produced parent_core / action) FIRST, then let the next search launch be scheduled
  // so that its tree-side prologue overlaps this kernel.  The parameter loads and index arithmetic above are inputs of
  // the wait, so they happen before it.
  if (pdl) {
    asm volatile("griddepcontrol.wait;" ::"l"(pc), "l"(action), "l"(policy), "l"(value), "l"(terminated), "l"(new_core),
                 "l"(new_payload), "r"(F), "r"(g.rho256), "r"(g.payload_bytes)
                 : "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  TZ_TL(atomicMax, tl_slot, 1);
  if (tl_row && lane == 0) atomicMax(tl_row + 1, gtime_ns());
  const uint32_t h2 = tz_synth_step_h((uint32_t)pc[0], (uint32_t)action[b]);
  const int d2 = pc[1] + 1;
  const int term = tz_synth_terminal(h2, d2, g.tau1024, g.max_depth);
  float x[MAX_NC];
  for (int c = 0; c < nc; ++c) {
    const uint32_t a = (uint32_t)(c * 32 + lane);
    x[c] = tz_synth_legal(h2, a, g.rho256) ? tz_synth_logit(h2, a) : -TZ_FLT_MAX;  // mcts.py:170
  }
  warp_softmax(x, nc, F, lane);  // mcts.py:171
  for (int c = 0; c < nc; ++c) {
    const int a = c * 32 + lane;
    if (a < F) policy[(size_t)b * F + a] = x[c];
  }
  if (lane == 0) {
    value[b] = term ? tz_synth_reward(h2) : tz_synth_value(h2);  // mcts.py:166,172
    terminated[b] = (uint8_t)term;
  }
  write_state(g, h2, d2, 1 - pc[2], new_core + 4 * (size_t)b, new_payload ? new_payload + (size_t)b * g.payload_bytes : nullptr, lane);
  TZ_TL(atomicMax, tl_slot, 2);
  if (tl_row && lane == 0) atomicMax(tl_row + 2, gtime_ns());
}

__global__ void __launch_bounds__(THREADS) k_env_step(const TzSynthGame g, const int B, const int env_offset,
                                                    const int32_t* __restrict__ action, int32_t* core, uint8_t* payload,
                                                    int32_t* episode, uint8_t* reset_flag) {
  const int b = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  int32_t* c = core + 4 * (size_t)b;
  const uint32_t h = (uint32_t)c[0];
  const int depth = c[1], player = c[2];
  const int ep = episode[b];
  __syncwarp();
  const uint32_t h2 = tz_synth_step_h(h, (uint32_t)action[b]);
  const int d2 = depth + 1;
  uint8_t* pay = payload ? payload + (size_t)b * g.payload_bytes : nullptr;
  if (tz_synth_terminal(h2, d2, g.tau1024, g.max_depth)) {  // common.py:84-99: episode over -> env_init_fn
    write_state(g, tz_synth_init_h(g.seed, (uint32_t)(b + env_offset), (uint32_t)(ep + 1)), 0, 0, c, pay, lane);
    if (lane == 0) {
      episode[b] = ep + 1;
      reset_flag[b] = 1;
    }
  } else {
    write_state(g, h2, d2, 1 - player, c, pay, lane);
    if (lane == 0) reset_flag[b] = 0;
  }
}

inline int grid_for(int B) { return (B * 32 + THREADS - 1) / THREADS; }
inline int status() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? TZ_OK : (int)e;
}
inline int check_game(const TzSynthGame* g, int B) {
  if (!g || B <= 0 || g->F <= 0 || g->payload_bytes < 0) return TZ_EINVAL;
  if (g->F > 32 * MAX_NC) return TZ_ENOTSUP;
  return TZ_OK;
}

}  // namespace

extern "C" {

uint64_t tz_synth_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

#ifdef TZ_PROFILE
int tz_synth_debug_timeline(unsigned long long* out, int reset) {  // diagnostic build only: 1024 rows of 4
  const cudaError_t e = cudaMemcpyFromSymbol(out, g_tl, sizeof(g_tl));
  if (e != cudaSuccess || !reset) return (int)e;
  static unsigned long long init[4 * 1024];
  for (int i = 0; i < 1024; ++i) {
    init[4 * i + 0] = ~0ull;
    init[4 * i + 1] = init[4 * i + 2] = init[4 * i + 3] = 0ull;
  }
  g_tl_seq.store(0);
  return (int)cudaMemcpyToSymbol(g_tl, init, sizeof(init));
}
#endif

int tz_synth_set_timeline(uint64_t* rows_dev, int slots) {
  if (rows_dev && (slots <= 0 || (slots & (slots - 1)) != 0)) return TZ_EINVAL;
  unsigned long long* rows = reinterpret_cast<unsigned long long*>(rows_dev);
  const cudaError_t e = cudaMemcpyToSymbol(c_tl_rows, &rows, sizeof(rows));
  if (e != cudaSuccess) return (int)e;
  g_tl_slots.store(rows_dev ? (uint32_t)slots : 0u, std::memory_order_relaxed);
  g_tl_rows.store(rows, std::memory_order_relaxed);
  return TZ_OK;
}

uint64_t tz_synth_leaf_seq(void) { return g_leaf_seq.load(std::memory_order_relaxed); }

int tz_synth_set_programmatic(int on) { return g_programmatic.exchange(on ? 1 : 0, std::memory_order_relaxed); }

int tz_synth_init_states(const TzSynthGame* g, int B, int env_offset, const int32_t* episode, int32_t* core,
                         uint8_t* payload, tz_stream_t stream) {
  const int rc = check_game(g, B);
  if (rc) return rc;
  if (!episode || !core || (g->payload_bytes > 0 && !payload)) return TZ_EINVAL;
  k_init_states<<<grid_for(B), THREADS, 0, (cudaStream_t)stream>>>(*g, B, env_offset, episode, core, payload);
  return status();
}

int tz_synth_root(const TzSynthGame* g, int B, const int32_t* core, const float* dir_noise, float dir_eps,
                  float* root_policy, float* root_value, tz_stream_t stream) {
  const int rc = check_game(g, B);
  if (rc) return rc;
  if (!core || !root_policy || !root_value) return TZ_EINVAL;
  const float one_minus = (float)(1.0 - (double)dir_eps);
  k_root<<<grid_for(B), THREADS, 0, (cudaStream_t)stream>>>(*g, B, core, dir_noise, dir_eps, one_minus, root_policy, root_value);
  return status();
}

int tz_synth_leaf(const TzSynthGame* g, int B, const int32_t* parent_core, const int32_t* action, float* policy,
                  float* value, uint8_t* terminated, int32_t* new_core, uint8_t* new_payload, tz_stream_t stream) {
  const int rc = check_game(g, B);
  if (rc) return rc;
  if (!parent_core || !action || !policy || !value || !terminated || !new_core) return TZ_EINVAL;
  if (g->payload_bytes > 0 && !new_payload) return TZ_EINVAL;
#ifdef TZ_PROFILE
  int tl_slot = (int)((g_tl_seq.fetch_add(1, std::memory_order_relaxed) & 1023u) << 2);
#else
  int tl_slot = 0;
  if (g_tl_rows.load(std::memory_order_relaxed))
    tl_slot = 2 | (int)((g_leaf_seq.fetch_add(1, std::memory_order_relaxed) & (uint64_t)(g_tl_slots.load(std::memory_order_relaxed) - 1)) << 2);
#endif
  if (g_programmatic.load(std::memory_order_relaxed)) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid_for(B));
    lc.blockDim = dim3(THREADS);
    lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&lc, k_leaf, *g, B, parent_core, action, policy, value, terminated, new_core, new_payload, 1 | tl_slot);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return e == cudaSuccess ? TZ_OK : (int)e;
  }
  k_leaf<<<grid_for(B), THREADS, 0, (cudaStream_t)stream>>>(*g, B, parent_core, action, policy, value, terminated, new_core,
                                                          new_payload, 0 | tl_slot);
  return status();
}

int tz_synth_env_step(const TzSynthGame* g, int B, int env_offset, const int32_t* action, int32_t* core,
                      uint8_t* payload, int32_t* episode, uint8_t* reset_flag, tz_stream_t stream) {
  const int rc = check_game(g, B);
  if (rc) return rc;
  if (!action || !core || !episode || !reset_flag || (g->payload_bytes > 0 && !payload)) return TZ_EINVAL;
  k_env_step<<<grid_for(B), THREADS, 0, (cudaStream_t)stream>>>(*g, B, env_offset, action, core, payload, episode, reset_flag);
  return status();
}

int tz_synth_leaf_cb(void* user, int sim, const TzWork* w, tz_stream_t stream) {
  (void)sim;
  const TzSynthCtx* ctx = reinterpret_cast<const TzSynthCtx*>(user);
  if (!ctx || !w) return TZ_EINVAL;
  return tz_synth_leaf(&ctx->game, ctx->B, reinterpret_cast<const int32_t*>(w->emb_parent[0]), w->action, w->policy,
                       w->value, w->terminated, reinterpret_cast<int32_t*>(w->emb_new[0]),
                       ctx->game.payload_bytes > 0 ? reinterpret_cast<uint8_t*>(w->emb_new[1]) : nullptr, stream);
}

// ---- bench instrumentation: CUDA events recorded on the launching stream around every search-kernel launch ----------
// tz_search calls the leaf callback between two search launches, so an event recorded at callback entry closes the
// previous search launch and one recorded before returning opens the next.  Everything is enqueued from C, so the
// caller can queue a whole move behind a blocker and read pure device-side durations afterwards.
static cudaEvent_t* g_ev_open = nullptr;
static cudaEvent_t* g_ev_close = nullptr;
static int g_ev_n = 0;

int tz_synth_timed_begin(int n_sims) {
  if (n_sims <= 0) return TZ_EINVAL;
  if (g_ev_n != n_sims) {
    for (int i = 0; i < g_ev_n; ++i) {
      cudaEventDestroy(g_ev_open[i]);
      cudaEventDestroy(g_ev_close[i]);
    }
    delete[] g_ev_open;
    delete[] g_ev_close;
    g_ev_open = new cudaEvent_t[n_sims];
    g_ev_close = new cudaEvent_t[n_sims];
    g_ev_n = n_sims;
    for (int i = 0; i < n_sims; ++i) {
      if (cudaEventCreate(&g_ev_open[i]) != cudaSuccess || cudaEventCreate(&g_ev_close[i]) != cudaSuccess) return TZ_EINVAL;
    }
  }
  return TZ_OK;
}

int tz_synth_leaf_cb_timed(void* user, int sim, const TzWork* w, tz_stream_t stream) {
  if (sim < 0 || sim >= g_ev_n) return TZ_EINVAL;
  cudaEventRecord(g_ev_close[sim], (cudaStream_t)stream);
  const int rc = tz_synth_leaf_cb(user, sim, w, stream);
  cudaEventRecord(g_ev_open[sim], (cudaStream_t)stream);
  return rc;
}

// ms_out[s] = duration of the search launch between leaf s and leaf s+1 (s = 0 .. n_sims-2): the fused
// expand+backprop(s) / select(s+1) launch.  leaf_ms_out[s] (optional) = duration of leaf s.  Call after a stream sync.
int tz_synth_timed_collect(float* ms_out, float* leaf_ms_out) {
  for (int s = 0; s + 1 < g_ev_n; ++s)
    if (cudaEventElapsedTime(&ms_out[s], g_ev_open[s], g_ev_close[s + 1]) != cudaSuccess) return TZ_EINVAL;
  if (leaf_ms_out)
    for (int s = 0; s < g_ev_n; ++s)
      if (cudaEventElapsedTime(&leaf_ms_out[s], g_ev_close[s], g_ev_open[s]) != cudaSuccess) return TZ_EINVAL;
  return TZ_OK;
}

}  // extern "C"
