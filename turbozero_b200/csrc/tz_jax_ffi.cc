// tz_jax_ffi.cc -- XLA FFI handlers that forward to the C-ABI of libtz_b200.so (include/tz_abi.h), so that the reference's
// JAX code can call the sm_100a kernels with `jax.ffi.ffi_call` (north star: "a thin C-ABI registered with jax.ffi").
//
// NOT COMPILED OR TESTED IN THIS IMAGE: jax / jaxlib and the XLA FFI headers (`jax.ffi.include_dir()`) are not installed and
// there is no network.  The file is mechanical -- every handler unpacks buffers into TzTree / TzWork and calls one tz_*
// entry point on XLA's stream -- and is built by `turbozero_b200.build.build_jax_ffi()` where jax is available.
// Python side: turbozero_b200/ffi_jax.py.  Reference call sites: INTEGRATION.md section 1.
//
// Operand order of every handler (all with the leading batch axis jax.vmap adds, vmap_method="broadcast_all"):
//   tree  = next_free_idx[B], parents[B,N], edge_map[B,N,F], n[B,N], p[B,N,F], q[B,N], terminated[B,N],
//           child_stats[B,N,F,4], best[B,N,2], sel_state[B,8], (r[B,N] if weighted), emb_0 .. emb_{K-1} [B,N,...]
//   The three derived tables are extra leaves the binding adds to the MCTSTree pytree (allocated by init, see ffi_jax.py).
//   Mutating handlers take the tree as operands AND as results and work on the result buffers.  With
//   input_output_aliases (jax >= 0.4.38) XLA hands the same buffer in and out; without it (jax 0.4.35-0.4.37, whose
//   ffi_call has no aliasing argument) the result buffers arrive UNINITIALISED, so every mutating handler first copies
//   each operand into its result wherever the two pointers differ (CopyIn below; a no-op when aliased).
#include <cuda_runtime.h>

#include "tz_abi.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

constexpr int kFixed = 10;  // tree leaves before the optional `r` and the embedding leaves

template <typename Get>  // Get(i) -> void* of tree leaf i and its AnyBuffer (args or rets)
ffi::Error FillTree(Get get, int weighted, int n_emb, TzTree* t) {
  auto edge = get(2);
  auto dims = edge.dimensions();
  if (dims.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "edge_map must be [B,N,F]");
  t->B = (int32_t)dims[0];
  t->N = (int32_t)dims[1];
  t->F = (int32_t)dims[2];
  t->n_emb = n_emb;
  t->next_free_idx = (int32_t*)get(0).untyped_data();
  t->parents = (int32_t*)get(1).untyped_data();
  t->edge_map = (int32_t*)edge.untyped_data();
  t->n = (int32_t*)get(3).untyped_data();
  t->p = (float*)get(4).untyped_data();
  t->q = (float*)get(5).untyped_data();
  t->terminated = (uint8_t*)get(6).untyped_data();
  t->child_stats = (int32_t*)get(7).untyped_data();
  t->best = (int32_t*)get(8).untyped_data();
  t->sel_state = (int32_t*)get(9).untyped_data();
  t->r = weighted ? (float*)get(kFixed).untyped_data() : nullptr;
  if (n_emb < 0 || n_emb > TZ_MAX_EMB) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "too many embedding leaves");
  for (int k = 0; k < n_emb; ++k) {
    auto e = get(kFixed + (weighted ? 1 : 0) + k);
    t->emb[k] = e.untyped_data();
    t->emb_row_bytes[k] = (int64_t)(e.size_bytes() / ((size_t)t->B * (size_t)t->N));
  }
  t->stats = nullptr;
  return ffi::Error::Success();
}

ffi::Error Status(int rc) {
  return rc == TZ_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, tz_strerror(rc));
}

TzSearchCfg MakeCfg(int32_t selector, float c, float c1, float c2, float epsilon, float discount, int32_t weighted,
                    float inv_q_temperature) {
  TzSearchCfg cfg = {};
  cfg.selector = selector;
  cfg.c = c;
  cfg.c1 = c1;
  cfg.c2 = c2;
  cfg.epsilon = epsilon;
  cfg.discount = discount;
  cfg.weighted = weighted;
  cfg.inv_q_temperature = inv_q_temperature;
  return cfg;  // fma_backup = 0, programmatic = 0: XLA's kernels sit between the launches
}

int TreeLeaves(int weighted, int n_emb) { return kFixed + (weighted ? 1 : 0) + n_emb; }

// operand i -> result j on `stream` unless XLA aliased them (same pointer).  Sizes must agree.
ffi::Error CopyIn(cudaStream_t stream, ffi::RemainingArgs& args, int i, ffi::RemainingRets& rets, int j) {
  auto src = *args.get<ffi::AnyBuffer>(i);
  auto dst = **rets.get<ffi::AnyBuffer>(j);
  if (src.untyped_data() == dst.untyped_data()) return ffi::Error::Success();
  if (src.size_bytes() != dst.size_bytes()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "operand / result size mismatch");
  const cudaError_t e = cudaMemcpyAsync(dst.untyped_data(), src.untyped_data(), src.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  return e == cudaSuccess ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, cudaGetErrorString(e));
}

// the first `count` operands into the first `count` results (the tree leaves of a mutating handler)
ffi::Error CopyInTree(cudaStream_t stream, ffi::RemainingArgs& args, ffi::RemainingRets& rets, int count) {
  for (int i = 0; i < count; ++i) {
    auto err = CopyIn(stream, args, i, rets, i);
    if (err.failure()) return err;
  }
  return ffi::Error::Success();
}

// ---- MCTS.update_root_node + Tree.set_root (mcts.py:363-384, tree.py:135-150) -----------------------------------------
// args: tree..., root_policy[B,F], root_value[B], root_emb_0..K-1 ; rets: tree... (aliased)
ffi::Error SetRootImpl(cudaStream_t stream, int32_t weighted, int32_t n_emb, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  const int L = TreeLeaves(weighted, n_emb);
  TzTree t = {};
  auto err = CopyInTree(stream, args, rets, L);
  if (err.failure()) return err;
  err = FillTree([&](int i) { return **rets.get<ffi::AnyBuffer>(i); }, weighted, n_emb, &t);
  if (err.failure()) return err;
  void* emb[TZ_MAX_EMB] = {};
  for (int k = 0; k < n_emb; ++k) emb[k] = args.get<ffi::AnyBuffer>(L + 2 + k)->untyped_data();
  return Status(tz_set_root(&t, (const float*)args.get<ffi::AnyBuffer>(L)->untyped_data(),
                            (const float*)args.get<ffi::AnyBuffer>(L + 1)->untyped_data(), emb, stream));
}

// ---- MCTS.traverse (mcts.py:192-228) + parent-embedding gather (mcts.py:161-164) ---------------------------------------
// args: tree..., path[B,TZ_PATH_STRIDE] ; rets: parent[B], action[B], path (aliased), best' [B,N,2] (aliased), sel_state' [B,8]
// (aliased), emb_parent_0..K-1 [B,...].  The walk fills in unknown best-table entries and a change of selector parameters
// drops the table, so `best` and `sel_state` are results of this handler too (never written through an operand).
ffi::Error SelectImpl(cudaStream_t stream, int32_t selector, float c, float c1, float c2, float epsilon, float discount,
                      int32_t weighted, int32_t n_emb, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  const int L = TreeLeaves(weighted, n_emb);
  TzTree t = {};
  auto err = FillTree([&](int i) { return *args.get<ffi::AnyBuffer>(i); }, weighted, n_emb, &t);
  if (err.failure()) return err;
  err = CopyIn(stream, args, L, rets, 2);  // path
  if (err.failure()) return err;
  err = CopyIn(stream, args, 8, rets, 3);  // best
  if (err.failure()) return err;
  err = CopyIn(stream, args, 9, rets, 4);  // sel_state
  if (err.failure()) return err;
  t.best = (int32_t*)(*rets.get<ffi::AnyBuffer>(3))->untyped_data();
  t.sel_state = (int32_t*)(*rets.get<ffi::AnyBuffer>(4))->untyped_data();
  TzWork w = {};
  w.parent = (int32_t*)(*rets.get<ffi::AnyBuffer>(0))->untyped_data();
  w.action = (int32_t*)(*rets.get<ffi::AnyBuffer>(1))->untyped_data();
  w.path = (int32_t*)(*rets.get<ffi::AnyBuffer>(2))->untyped_data();
  for (int k = 0; k < n_emb; ++k) w.emb_parent[k] = (*rets.get<ffi::AnyBuffer>(5 + k))->untyped_data();
  const TzSearchCfg cfg = MakeCfg(selector, c, c1, c2, epsilon, discount, weighted, 1.0f);
  return Status(tz_select(&t, &cfg, &w, stream));
}

// ---- second half of MCTS.iterate (mcts.py:174-189), optionally fused with the next traverse -----------------------------
// args: tree..., parent[B], action[B], path, policy[B,F], value[B], terminated[B], new_emb_0..K-1, (backprop_noise[B,F])
// rets: tree... (aliased), parent', action', path' (aliased), emb_parent_0..K-1      (the last four groups only if fused)
ffi::Error ExpandImpl(cudaStream_t stream, int32_t selector, float c, float c1, float c2, float epsilon, float discount,
                      int32_t weighted, float inv_q_temperature, int32_t n_emb, int32_t fused, int32_t has_noise,
                      ffi::RemainingArgs args, ffi::RemainingRets rets) {
  const int L = TreeLeaves(weighted, n_emb);
  TzTree t = {};
  auto err = CopyInTree(stream, args, rets, fused ? L + 3 : L);  // tree leaves (+ parent / action / path when fused)
  if (err.failure()) return err;
  err = FillTree([&](int i) { return **rets.get<ffi::AnyBuffer>(i); }, weighted, n_emb, &t);
  if (err.failure()) return err;
  TzWork w = {};
  // parent / action / path are read (this simulation) and, when fused, rewritten (the next one): aliased in -> out
  w.parent = (int32_t*)(fused ? (*rets.get<ffi::AnyBuffer>(L))->untyped_data() : args.get<ffi::AnyBuffer>(L)->untyped_data());
  w.action = (int32_t*)(fused ? (*rets.get<ffi::AnyBuffer>(L + 1))->untyped_data() : args.get<ffi::AnyBuffer>(L + 1)->untyped_data());
  w.path = (int32_t*)(fused ? (*rets.get<ffi::AnyBuffer>(L + 2))->untyped_data() : args.get<ffi::AnyBuffer>(L + 2)->untyped_data());
  w.policy = (float*)args.get<ffi::AnyBuffer>(L + 3)->untyped_data();
  w.value = (float*)args.get<ffi::AnyBuffer>(L + 4)->untyped_data();
  w.terminated = (uint8_t*)args.get<ffi::AnyBuffer>(L + 5)->untyped_data();
  for (int k = 0; k < n_emb; ++k) w.emb_new[k] = args.get<ffi::AnyBuffer>(L + 6 + k)->untyped_data();
  if (has_noise) w.backprop_noise = (float*)args.get<ffi::AnyBuffer>(L + 6 + n_emb)->untyped_data();
  if (fused)
    for (int k = 0; k < n_emb; ++k) w.emb_parent[k] = (*rets.get<ffi::AnyBuffer>(L + 3 + k))->untyped_data();
  const TzSearchCfg cfg = MakeCfg(selector, c, c1, c2, epsilon, discount, weighted, inv_q_temperature);
  return Status(fused ? tz_expand_backprop_select(&t, &cfg, &w, stream) : tz_expand_backprop(&t, &cfg, &w, stream));
}

// ---- MCTS.sample_root_action + get_value (mcts.py:265-296, 111-120) ------------------------------------------------------
// args: tree..., noise[B,F], uniform01[B] ; rets: action[B], policy_weights[B,F], root_q[B]
ffi::Error RootActionImpl(cudaStream_t stream, float temperature, int32_t weighted, int32_t n_emb, ffi::RemainingArgs args,
                          ffi::RemainingRets rets) {
  const int L = TreeLeaves(weighted, n_emb);
  TzTree t = {};
  auto err = FillTree([&](int i) { return *args.get<ffi::AnyBuffer>(i); }, weighted, n_emb, &t);
  if (err.failure()) return err;
  return Status(tz_root_action(&t, temperature, (const float*)args.get<ffi::AnyBuffer>(L)->untyped_data(),
                               (const float*)args.get<ffi::AnyBuffer>(L + 1)->untyped_data(), nullptr,
                               (float*)(*rets.get<ffi::AnyBuffer>(1))->untyped_data(), (float*)(*rets.get<ffi::AnyBuffer>(2))->untyped_data(),
                               (int32_t*)(*rets.get<ffi::AnyBuffer>(0))->untyped_data(), stream));
}

// ---- MCTS.step / reset with the caller's select (mcts.py:387-414, tree.py:169-278, common.py:89-94) ----------------------
// args: tree..., action[B], reset_flag[B] ; rets: tree... (aliased)
ffi::Error RerootImpl(cudaStream_t stream, int32_t persist_tree, int32_t weighted, int32_t n_emb, ffi::RemainingArgs args,
                      ffi::RemainingRets rets) {
  const int L = TreeLeaves(weighted, n_emb);
  TzTree t = {};
  auto err = CopyInTree(stream, args, rets, L);
  if (err.failure()) return err;
  err = FillTree([&](int i) { return **rets.get<ffi::AnyBuffer>(i); }, weighted, n_emb, &t);
  if (err.failure()) return err;
  return Status(tz_reroot(&t, (const int32_t*)args.get<ffi::AnyBuffer>(L)->untyped_data(),
                          (const uint8_t*)args.get<ffi::AnyBuffer>(L + 1)->untyped_data(), persist_tree, stream));
}

}  // namespace

#define TZ_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define TZ_SELECTOR_ATTRS \
  .Attr<int32_t>("selector").Attr<float>("c").Attr<float>("c1").Attr<float>("c2").Attr<float>("epsilon").Attr<float>("discount")

XLA_FFI_DEFINE_HANDLER_SYMBOL(TzSetRoot, SetRootImpl,
                              TZ_STREAM.Attr<int32_t>("weighted").Attr<int32_t>("n_emb").RemainingArgs().RemainingRets());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TzSelect, SelectImpl,
                              TZ_STREAM TZ_SELECTOR_ATTRS.Attr<int32_t>("weighted").Attr<int32_t>("n_emb").RemainingArgs().RemainingRets());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TzExpandBackprop, ExpandImpl,
                              TZ_STREAM TZ_SELECTOR_ATTRS.Attr<int32_t>("weighted").Attr<float>("inv_q_temperature")
                                  .Attr<int32_t>("n_emb").Attr<int32_t>("fused").Attr<int32_t>("has_noise").RemainingArgs().RemainingRets());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TzRootAction, RootActionImpl,
                              TZ_STREAM.Attr<float>("temperature").Attr<int32_t>("weighted").Attr<int32_t>("n_emb").RemainingArgs().RemainingRets());
XLA_FFI_DEFINE_HANDLER_SYMBOL(TzReroot, RerootImpl,
                              TZ_STREAM.Attr<int32_t>("persist_tree").Attr<int32_t>("weighted").Attr<int32_t>("n_emb").RemainingArgs().RemainingRets());
