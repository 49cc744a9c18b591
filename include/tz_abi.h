/*
 * tz_abi.h -- C-ABI of libtz_b200.so: the batched MCTS hot path of lowrollr/turbozero as
 * hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI: its search is Python/JAX traced into XLA.  Each entry point below
 * replaces the XLA computation that one reference method lowers to; the reference location is
 * cited on every declaration (paths relative to the reference repo root).  A jax.ffi / ctypes /
 * cffi binding passes raw device pointers and a stream; no framework types cross this boundary.
 *
 * Conventions
 *  - All array arguments are DEVICE pointers unless the name ends in `_host`.
 *  - Layout is the reference's batched pytree layout (core/evaluators/evaluator.py:42-45 adds
 *    the leading batch axis to core/trees/tree.py:11-19): batch-major, struct-of-arrays,
 *    row-major, dense.  bool is one byte (0/1).
 *  - Every call is asynchronous on `stream`, never allocates, never synchronises, never throws.
 *    Return value: TZ_OK (0), a negative TZ_E* code, or a positive cudaError_t from the launch.
 *  - The tree arrays are updated IN PLACE (the reference returns new pytrees; XLA aliases them).
 *  - Invariant kept by every entry point (and by the reference): rows >= next_free_idx are null
 *    (-1 in parents/edge_map, zero bytes in every data leaf).
 *  - Thread-safety: re-entrant; calls on the same tree must be stream-ordered by the caller.
 */
#ifndef TZ_ABI_H_
#define TZ_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TZ_ABI_VERSION 7
#define TZ_MAX_EMB 24 /* max number of embedding pytree leaves per node */
#define TZ_PATH_CAP 32 /* path slots kept per tree between select and backprop */
#define TZ_PATH_STRIDE (2 * TZ_PATH_CAP + 2) /* ints per tree in TzWork.path: nodes[32], actions[32], length, end child */
#define TZ_SEL_STATE_WORDS 8 /* ints per tree in TzTree.sel_state */

#define TZ_OK 0
#define TZ_EINVAL (-1)   /* bad argument (null pointer, B/N/F <= 0, unknown selector, ...) */
#define TZ_ENOTSUP (-2)  /* configuration outside what the kernels implement */

#define TZ_NULL_INDEX (-1) /* core/trees/tree.py:21 */
#define TZ_ROOT_INDEX 0    /* core/trees/tree.py:23 */

typedef void* tz_stream_t; /* cudaStream_t */

/* One batch of B fixed-capacity trees: core/trees/tree.py:11-19 (Tree) holding
 * core/evaluators/mcts/state.py:12-25 (MCTSNode) or weighted_mcts.py:14-17 (WeightedMCTSNode). */
typedef struct TzTree {
  int32_t B;                /* trees in this batch (envs on this device) */
  int32_t N;                /* max_nodes  (mcts.py:22) */
  int32_t F;                /* branching_factor (mcts.py:21) */
  int32_t n_emb;            /* number of embedding leaves, 0..TZ_MAX_EMB */
  int32_t* next_free_idx;   /* [B]      tree.py:15 */
  int32_t* parents;         /* [B,N]    tree.py:16 */
  int32_t* edge_map;        /* [B,N,F]  tree.py:17 */
  int32_t* n;               /* [B,N]    state.py:21 visit count */
  float* p;                 /* [B,N,F]  state.py:22 policy */
  float* q;                 /* [B,N]    state.py:23 value estimate */
  float* r;                 /* [B,N] raw leaf value, weighted_mcts.py:17; NULL for plain MCTS */
  uint8_t* terminated;      /* [B,N]    state.py:24 */
  int32_t* child_stats;     /* [B,N,F,4] DERIVED table, not part of the reference pytree: for every edge one 16-byte entry
                               {bit pattern of q[child], n[child] | terminated[child] << 31, bit pattern of p[node,a],
                               edge_map[node,a]} with {0, 0} for the first two words where edge_map is -1 -- exactly what
                               Tree.get_child_data (tree.py:78-98) would gather plus the node's own p and edge_map entry,
                               so that the selector reads ONE vector per child and one selection level is ONE memory
                               round trip.  edge_map and p stay authoritative; every entry point keeps the copies in
                               sync (null rows are {0, 0, 0, -1}); rebuild with tz_rebuild_child_stats after writing
                               q / n / p / terminated / edge_map from outside. */
  int32_t* best;            /* [B,N,2] DERIVED table: the selector's decision at every node, {action, next}, computed when
                               the node's statistics last changed (backprop / expansion) instead of when the walk arrives:
                               `next` >= 0 is the child the walk continues into, -1 means "no edge: expand here",
                               -(2+c) means "child c exists and is terminal: re-expand it" (mcts.py:208-213).
                               action == -1 marks the entry unknown (the walk then scores the node itself and fills it
                               in).  With it MCTS.traverse is one dependent load per level.  Maintained by every entry
                               point; reset by tz_tree_init / tz_rebuild_child_stats. */
  int32_t* sel_state;       /* [B,TZ_SEL_STATE_WORDS] the selector parameters {selector, c, c1, c2, epsilon, discount}
                               `best` was computed with; a launch with different parameters drops the tree's entries. */
  void* emb[TZ_MAX_EMB];    /* [B,N,emb_row_bytes[k]]  state.py:25, one table per pytree leaf */
  int64_t emb_row_bytes[TZ_MAX_EMB];
  uint64_t* stats;          /* optional [B,4] counters {select levels, simulations, rows before
                               re-root, rows kept by re-root}; NULL disables counting */
} TzTree;

/* Action selectors: core/evaluators/mcts/action_selection.py */
#define TZ_SEL_PUCT 0        /* PUCTSelector :61-116 */
#define TZ_SEL_MUZERO_PUCT 1 /* MuZeroPUCTSelector :119-177 (intended maths) */

/* q_transform of a selector (action_selection.py:70,128 constructor argument `q_transform`): a REGISTRY of device
 * functors, selected by id.  Adding one = a new id here, a case in q_transform_apply (csrc/tz_kernels.cu), the same case in
 * both oracles (oracle/mcts_numpy.py, oracle/tz_oracle.c) and a descriptor in turbozero_b200/action_selection.py
 * (register_q_transform); arbitrary Python callables cannot run inside the kernel. */
#define TZ_QT_NORMALIZE 0 /* normalize_q_values, action_selection.py:10-32 (the default) */
#define TZ_QT_IDENTITY 1  /* lambda q, n, parent_q, eps: q -- the discounted child values as they are (0 for missing children) */
#define TZ_QT_COUNT 2

typedef struct TzSearchCfg {
  int32_t selector;      /* TZ_SEL_* */
  float c;               /* PUCTSelector.c  (action_selection.py:67) */
  float c1, c2;          /* MuZeroPUCTSelector (action_selection.py:123-124) */
  float epsilon;         /* selector epsilon (action_selection.py:41,68) */
  float discount;        /* MCTS.discount (mcts.py:24) */
  int32_t weighted;      /* 0: MCTS.backpropagate mcts.py:231-262; 1: WeightedMCTS weighted_mcts.py:90-152 */
  float inv_q_temperature; /* float32(1/q_temperature) if q_temperature > 0, else 0 (weighted_mcts.py:113-131) */
  int32_t fma_backup;    /* 1: q update as fmaf(q, n, v) / (n+1) (XLA may contract mcts.py:322); 0: separate mul, add */
  int32_t programmatic;  /* bit 0: the per-simulation launches (tz_expand_backprop[_select]) are PROGRAMMATIC DEPENDENT LAUNCHES
                            (cudaLaunchAttributeProgrammaticStreamSerialization): the kernel may start while the preceding
                            kernel in the stream is still running; it reads the TREE state and the previous select's outputs
                            (TzWork.parent / action / path) before griddepcontrol.wait and the leaf results (TzWork.policy /
                            value / terminated / emb_new) after it, so its launch latency and first two memory round trips
                            overlap the user's last leaf kernel.
                            bit 1: the kernel also executes griddepcontrol.launch_dependents once only its walk and
                            embedding gather remain, so a following kernel that was itself launched programmatically
                            (and waits before it reads) has its launch latency hidden too.
                            Contract for bit 0: between two tz launches on the same trees the stream holds at least one
                            kernel that is launched without the programmatic attribute, or that executes
                            griddepcontrol.wait before griddepcontrol.launch_dependents (ordinary framework kernels
                            satisfy the first form).  Pays when that kernel is short AND cooperates (is itself launched
                            programmatically and waits before it reads): measured on one B200 against ordinary launches
                            (profiles/r2ak_modes.log) +16 % on configs[1] with both bits, +4 % on the 2048 shape with bit 0
                            only, +1 % on go_9x9, -6 % on the othello shape; behind an ORDINARY leaf kernel bit 0 alone
                            measured -1 ... -3 %.  0: ordinary stream-ordered launches (the Python API's default). */
  int32_t q_transform;   /* TZ_QT_*: the selector's q_transform (action_selection.py:70,109) */
  int32_t sim_warps;     /* warps that cooperate on ONE tree in the per-simulation kernel: 0 = library's choice (1 for narrow
                            trees, 4 for trees with more than 32 actions), 1 = one warp per tree, 2 / 4 / 8 = a CTA per tree
                            whose warps score the path levels side by side (results are identical for every value) */
} TzSearchCfg;

/* Per-simulation exchange buffers between the kernels and the host framework's
 * env_step_fn / eval_fn (mcts.py:160-172).  Caller-allocated. */
typedef struct TzWork {
  int32_t* parent;          /* [B]  out of select: TraversalState.parent (state.py:43) */
  int32_t* action;          /* [B]  out of select: TraversalState.action (state.py:44) */
  void* emb_parent[TZ_MAX_EMB]; /* [B,row_k] out of select: tree.data_at(parent).embedding (mcts.py:164) */
  float* policy;            /* [B,F] in to expand: masked+softmaxed policy (mcts.py:170-171) */
  float* value;             /* [B]   in: where(terminated, player_reward, value) (mcts.py:172) */
  uint8_t* terminated;      /* [B]   in: metadata.terminated (mcts.py:172,179-180) */
  void* emb_new[TZ_MAX_EMB];/* [B,row_k] in: new_embedding (mcts.py:165) */
  float* backprop_noise;    /* [B,F] in, weighted q_temperature==0 only: uniform(0,tiebreak_noise) (weighted_mcts.py:123); else NULL */
  int32_t* path;            /* [B,TZ_PATH_STRIDE] scratch owned by the library between select and the following
                               expand_backprop of the same trees (visited nodes, actions taken, path length);
                               NULL = walk parents[] and search edge rows instead */
  int32_t* path_spill;      /* optional [B,path_spill_cap,2] scratch, same ownership: {node, action} of the path levels that
                               no longer fit the 32-level ring of `path` (level l of a path of length L > 32 is kept here
                               for l < L - 32).  With it a deep backup processes 32 levels per memory round trip like the
                               ring does; without it (NULL) levels above the ring are reached by chasing parents[], one
                               dependent round trip each.  path_spill_cap >= max_nodes - 32 covers every possible path.
                               With TzSearchCfg.sim_warps > 1 (a CTA per tree) the WHOLE path is kept here linearly (level l
                               at entry l) and `path` only carries {tag, length, end child}: that form needs
                               path_spill_cap >= max_nodes (otherwise the library uses one warp per tree). */
  int32_t path_spill_cap;   /* entries per tree in path_spill */
  int32_t timeline_slots;   /* rows of `timeline` (a power of two), 0 = none */
  uint64_t* timeline;       /* optional [timeline_slots,4] measurement record of the per-simulation launches that use this
                               TzWork: row (launch sequence number mod timeline_slots, see tz_launch_seq) holds %globaltimer
                               nanoseconds {first warp in (min), last warp has its leaf results (max), last warp out (max),
                               unused}.  The caller initialises rows to {~0, 0, 0, 0}.  Costs three reductions per warp when
                               set; NULL = no record.  bench.py uses it to time the kernel INSIDE the step it describes. */
} TzWork;

int tz_abi_version(void);
const char* tz_strerror(int code);

/* Tree.reset / init_tree over the whole allocation: core/trees/tree.py:272-298. Writes every row. */
int tz_tree_init(const TzTree* t, tz_stream_t stream);

/* Recomputes TzTree.child_stats from edge_map / p / q / n / terminated (one pass over [B,N,F]) and marks every
 * TzTree.best entry unknown. */
int tz_rebuild_child_stats(const TzTree* t, tz_stream_t stream);

/* MCTS.update_root_node + Tree.set_root: mcts.py:363-384 (weighted_mcts.py:66-87), tree.py:135-150.
 * root_policy [B,F], root_value [B], root_emb[k] [B,row_k]. */
int tz_set_root(const TzTree* t, const float* root_policy, const float* root_value,
                void* const* root_emb, tz_stream_t stream);

/* MCTS.traverse with the selector inlined: mcts.py:192-228, action_selection.py:91-116, tree.py:78-98.
 * Fills w->parent, w->action, w->emb_parent (mcts.py:161-164) and w->path. */
int tz_select(const TzTree* t, const TzSearchCfg* cfg, const TzWork* w, tz_stream_t stream);

/* Second half of MCTS.iterate: node_exists test, visit_node/new_node, update_node/add_node and
 * backpropagate: mcts.py:174-189, 231-262, 299-360; tree.py:101-132,153-166; weighted_mcts.py:43-63,90-152. */
int tz_expand_backprop(const TzTree* t, const TzSearchCfg* cfg, const TzWork* w, tz_stream_t stream);

/* tz_expand_backprop for simulation i followed by tz_select for simulation i+1 in ONE launch. */
int tz_expand_backprop_select(const TzTree* t, const TzSearchCfg* cfg, const TzWork* w,
                              tz_stream_t stream);

/* MCTS.sample_root_action + MCTS.get_value: mcts.py:265-296, 111-120.
 * visits [B,F] int32 (may be NULL), policy_weights [B,F] (may be NULL), root_q [B] (may be NULL),
 * action [B] (may be NULL).  temperature == 0: argmax(policy_weights + noise[B,F]) with noise =
 * uniform(0, tiebreak_noise) supplied by the caller (mcts.py:281-285).  temperature > 0:
 * jax.random.choice(p = pw**(1/T) renormalised) driven by uniform01[B] (mcts.py:288-294). */
int tz_root_action(const TzTree* t, float temperature, const float* noise, const float* uniform01,
                   int32_t* visits, float* policy_weights, float* root_q, int32_t* action,
                   tz_stream_t stream);

/* MCTS.step / MCTS.reset fused with the caller's select (core/common.py:89-94):
 * per tree, reset_flag[b] == 2 -> tree left untouched (common.py:91, reset=False);
 * reset_flag[b] == 1 or persist_tree == 0 -> Tree.reset (tree.py:272-278);
 * else Tree.get_subtree(action[b]) (tree.py:169-269), action clamped to [0,F) like an XLA gather.
 * reset_flag may be NULL (all zero); action may be NULL only if persist_tree == 0.
 * Needs N*4 + <=64 KB of shared memory per CTA: N <= ~40000. */
int tz_reroot(const TzTree* t, const int32_t* action, const uint8_t* reset_flag, int persist_tree,
              tz_stream_t stream);

/* Callback that enqueues the host framework's env_step_fn + eval_fn + mask/softmax (mcts.py:160-172)
 * on `stream` for simulation `sim`, reading w->parent/action/emb_parent and filling
 * w->policy/value/terminated/emb_new.  Must not synchronise.  Non-zero return aborts the search. */
typedef int (*tz_leaf_fn)(void* user, int sim, const TzWork* w, tz_stream_t stream);

/* The lax.scan of MCTS.evaluate (mcts.py:99-100): num_iterations x iterate, as
 * select, then per simulation { leaf(user), expand_backprop[_select] }.  Enqueues only. */
int tz_search(const TzTree* t, const TzSearchCfg* cfg, const TzWork* w, int num_iterations,
              tz_leaf_fn leaf, void* user, tz_stream_t stream);

/* Self-test: compares the kernels' straight-line division sequence with the hardware's IEEE division (div.rn) on n
 * pseudo-random operand pairs; adds the number of differing results to *mismatches_dev (device uint64, caller-zeroed). */
int tz_selftest_div(uint64_t n, uint32_t seed, uint64_t* mismatches_dev, tz_stream_t stream);

/* Self-test: re-evaluates the selector at every allocated node whose TzTree.best entry is known and compares;
 * out_dev[0] += number of wrong entries (or non-null entries past next_free_idx), out_dev[1] += entries checked
 * (device uint64[2], caller-zeroed).  `cfg` must be the configuration the tree was last searched with. */
int tz_selftest_best(const TzTree* t, const TzSearchCfg* cfg, uint64_t* out_dev, tz_stream_t stream);

/* Number of kernels this library has launched since load (for bench accounting). */
uint64_t tz_launch_count(void);

/* Sequence number the NEXT per-simulation launch (tz_select / tz_expand_backprop[_select]) will carry; a launch writes row
 * (its sequence number mod timeline_slots) of TzWork.timeline.  Launches captured into a CUDA graph keep the number they
 * were captured with. */
uint64_t tz_launch_seq(void);

#ifdef __cplusplus
}
#endif
#endif /* TZ_ABI_H_ */
