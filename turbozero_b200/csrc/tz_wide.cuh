// tz_wide.cuh -- k_sim_wide: the per-simulation kernel with a CTA of W warps per tree, and its launcher template;
// instantiated for plain / weighted backups in tz_wide_plain.cu / tz_wide_weighted.cu (parallel builds).
#ifndef TZ_WIDE_CUH_
#define TZ_WIDE_CUH_

#include "tz_device.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// k_sim_wide: the per-simulation kernel with a CTA of W warps per tree (TzSearchCfg.sim_warps), for wide / deep trees.
//
// k_sim gives a tree ONE warp; on an 82-way tree every selector call costs that warp ~0.5 us and a simulation needs one
// per path level, so a launch lasts as long as its deepest path (go_9x9 shape: 46 us for the slowest warp of 1024).
// In plain MCTS the work of a simulation parallelises over the path LEVELS:
//   A  backup: level l's new statistics depend only on its old ones, the leaf value and its depth  (mcts.py:231-262)
//        -> one THREAD per level, 32 W levels per pass;
//   B  decisions: the selector at level l needs the node's rows, its new statistics and those of its path child
//        (action_selection.py:91-116) -> one WARP per level, W levels side by side, next row prefetched while one is scored;
//   C  the new walk (mcts.py:192-228) follows the old path as long as every new decision leads to the old next node: the
//        first level where it does not is found by one vote over all levels; only the remainder is walked sequentially
//        (one dependent best-table load per level, warp 0);
//   D  embedding rows (mcts.py:161-165, 354-360) are copied by the whole CTA, both copies' loads in flight together.
// WeightedMCTS (weighted_mcts.py:90-152) makes A sequential (a node's weighted value needs its child's NEW q), so there
// warp 0 runs the backup chain level by level and publishes each level's result in shared memory, and the other warps
// score the selector decisions behind it (two-stage pipeline): the chain no longer pays for the selector.
//
// Dependent memory round trips per launch: (1) everything with a static address -- scalars, the leaf results, and level
// `tid` of the path record (paths that fit one pass, the common case); (2) the path nodes' statistics, one thread per
// level, together with the first child_stats row of every warp; then arithmetic; (3) the walk's remainder; (4) the gather.
//
// The visited path is kept LINEARLY in TzWork.path_spill (level l at entry l, capacity >= max_nodes required; the library
// falls back to k_sim otherwise); TzWork.path only carries the NEGATED length and the end child, so that the two kernels
// never trust each other's record (k_sim needs a length >= 1, this kernel a length <= -1): paths of any length are handled
// 32 W levels at a time and nothing chases parents[] -- except when the record does not describe this expansion (parent /
// action not produced by this kernel's last select), where warp 0 rebuilds it from parents[] / edge_map first.
// Results are bit-identical to k_sim's (same select_core / weighted_value / backup_q on the same operands).
// ---------------------------------------------------------------------------------------------------------
#ifdef TZ_PROFILE
#define TZ_WSTAMP(i) do { if (threadIdx.x == 0 && blockIdx.x < 4096) g_prof_gt[16 * blockIdx.x + (i)] = prof_gtime(); } while (0)
#else
#define TZ_WSTAMP(i) do { } while (0)
#endif

// dst0 <- src0 and (optionally) dst1 <- src1, `bytes` each, by the whole CTA; the loads of both copies are issued before the stores
__device__ __forceinline__ void block_copy2(void* d0, const void* s0, void* d1, const void* s1, int64_t bytes, int tid, int nthr) {
  const uintptr_t a = (uintptr_t)d0 | (uintptr_t)s0 | (uintptr_t)d1 | (uintptr_t)s1 | (uintptr_t)bytes;
  if ((a & 15) == 0) {
    const int nv = (int)(bytes >> 4);
    for (int i0 = tid; i0 < nv; i0 += 2 * nthr) {  // two vectors per copy, thread and pass: their loads are in flight together
      uint4 x[2], y[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = i0 + k * nthr;
        if (i < nv) {
          if (d0) x[k] = reinterpret_cast<const uint4*>(s0)[i];
          if (d1) y[k] = reinterpret_cast<const uint4*>(s1)[i];
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = i0 + k * nthr;
        if (i < nv) {
          if (d0) reinterpret_cast<uint4*>(d0)[i] = x[k];
          if (d1) reinterpret_cast<uint4*>(d1)[i] = y[k];
        }
      }
    }
  } else if ((a & 3) == 0) {
    const int nv = (int)(bytes >> 2);
    for (int i = tid; i < nv; i += nthr) {
      uint32_t x = 0, y = 0;
      if (d0) x = reinterpret_cast<const uint32_t*>(s0)[i];
      if (d1) y = reinterpret_cast<const uint32_t*>(s1)[i];
      if (d0) reinterpret_cast<uint32_t*>(d0)[i] = x;
      if (d1) reinterpret_cast<uint32_t*>(d1)[i] = y;
    }
  } else {
    for (int64_t i = tid; i < bytes; i += nthr) {
      uint8_t x = 0, y = 0;
      if (d0) x = reinterpret_cast<const uint8_t*>(s0)[i];
      if (d1) y = reinterpret_cast<const uint8_t*>(s1)[i];
      if (d0) reinterpret_cast<uint8_t*>(d0)[i] = x;
      if (d1) reinterpret_cast<uint8_t*>(d1)[i] = y;
    }
  }
}

// CTAs per SM the register budget is capped for: every tree of a ~1 K-tree batch resident at once (7 x 148 = 1036 CTAs of four
// warps at <= 73 registers; two waves cost more than the spills)
// (trees with more than 128 actions hold 8-16 register chunks per lane: no cap there, the spills cost more; the weighted
// backup is one warp's dependent chain of exp / division sequences per level: capped at 128 registers it is 20 % faster than
// at 72, measured on the othello shape -- profiles/r2e_ab.log vs r2c_warps.log)
template <int NC, bool WEIGHTED, int W>
struct WideOcc {
  static constexpr int MIN_CTAS = NC > 4 ? 1 : (WEIGHTED ? (W == 2 ? 8 : (W == 4 ? 4 : 2)) : (W == 2 ? 14 : (W == 4 ? 7 : 3)));
};

template <int NC, bool WEIGHTED, int SEL, int W>
__global__ void __launch_bounds__(32 * W, WideOcc<NC, WEIGHTED, W>::MIN_CTAS) k_sim_wide(const __grid_constant__ SimP P, const __grid_constant__ SimLeafExtra X) {
  constexpr int NT = 32 * W;
  constexpr int WIN = NT;  // path levels per pass
  __shared__ int2 s_rec[WIN];    // {node, action taken there} of the pass's levels; index j <-> level lo + j
  __shared__ float s_q1[WIN];    // the level's q after this backup (weighted: before it, until the chain reaches the level)
  __shared__ int s_n1[WIN];      // the level's n after this backup (weighted: before)
  __shared__ float s_r[WEIGHTED ? WIN : 1];
  __shared__ int2 s_best[WIN];   // the level's new selector decision (best-table entry)
  __shared__ int s_wmin[W];
  __shared__ int s_walk[2];
  __shared__ uint64_t s_bar;       // completion of the staged best-table's bulk copy
  __shared__ volatile int s_done;  // weighted: levels of this pass whose backup is published, counted from the deepest
  extern __shared__ __align__(16) uint8_t wide_smem[];  // the tree's best-table, staged for the walk (P.best_rows > 0)
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = P.F;
  const int mode = P.mode;
  const bool do_expand = (mode & MODE_EXPAND) != 0, do_sel = (mode & MODE_SELECT) != 0;
  const TzSearchCfg& cfg = P.cfg;
  constexpr int XW = W - 1;  // the warp that writes the expansion (idle in the decisions unless the path has >= W levels)
  pdl_wait();  // (no-op unless launched programmatically) everything below may read what the preceding kernel wrote
  // (Splitting the kernel around the wait as k_sim does -- round trips 1 and 2 before it, the leaf results after -- was built
  // and measured on one box against this form: go_9x9 36.7 -> 36.0 M with programmatic launches and 36.3 -> 35.0 M with
  // ordinary ones, othello 26.5 -> 27.3 M programmatic but 28.0 -> 27.7 M ordinary, its best mode: profiles/r2ao_ab.log.)
  if (warp == 0) tl_min(P.tl_row, 0, lane);
  TZ_WSTAMP(0);
  const TV tv = make_view(P, b);
  int32_t* const path = P.w_path + (size_t)b * PATH_STRIDE;
  int2* const lin = P.w_spill + (size_t)b * P.spill_cap;  // the linear path record
  // ---- round trip 1: everything whose address is known at entry --------------------------------------------------------
  const int nfi = P.nfi[b];
  const int4 s0 = *reinterpret_cast<const int4*>(tv.sel);
  const int4 s1 = *reinterpret_cast<const int4*>(tv.sel + 4);
  int parent = 0, action = 0, termflag = 0, Lraw = 0, end_child = -1;
  float value = 0.0f;
  int2 rec = make_int2(0, 0);  // level `tid` of the recorded path, if that path fits one pass
  float pol[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) pol[c] = 0.0f;
  if (do_expand) {
    parent = P.w_parent[b];
    action = P.w_action[b];
    value = P.w_value[b];
    termflag = P.w_term[b] ? 1 : 0;
    Lraw = path[PATH_LEN];
    end_child = path[PATH_END];
    if (tid < P.spill_cap) rec = lin[tid];
    if (warp == XW) {
#pragma unroll
      for (int c = 0; c < NC; ++c) pol[c] = (c * 32 + lane < F) ? P.w_policy[(size_t)b * F + c * 32 + lane] : 0.0f;
    }
  }
  bool stale;
  {  // the best-table is only valid for the selector parameters it was computed with
    stale = s0.x != cfg.selector || s0.y != __float_as_int(cfg.c) || s0.z != __float_as_int(cfg.c1) ||
                       s0.w != __float_as_int(cfg.c2) || s1.x != __float_as_int(cfg.epsilon) ||
                       s1.y != __float_as_int(cfg.discount) || s1.z != cfg.q_transform;
    if (stale) {  // (uniform over the CTA)
      __syncthreads();  // every thread has read the old words
      for (int i = tid; i < nfi && i < tv.N; i += NT) tv.best[i] = make_int2(-1, -1);
      if (tid == 0) {
        *reinterpret_cast<int4*>(tv.sel) = make_int4(cfg.selector, __float_as_int(cfg.c), __float_as_int(cfg.c1), __float_as_int(cfg.c2));
        *reinterpret_cast<int4*>(tv.sel + 4) = make_int4(__float_as_int(cfg.epsilon), __float_as_int(cfg.discount), cfg.q_transform, 0);
      }
      __syncthreads();
    }
  }
  // stage the best-table for the walk (a fire-and-forget global -> shared copy; it lands while the backup runs): the walk's
  // remainder is then one shared-memory load per level instead of one dependent L2 / DRAM round trip
  int2* const sb = (P.best_rows > 0 && do_sel) ? reinterpret_cast<int2*>(wide_smem) : nullptr;
  // (one cp.async.bulk for the whole table when it is 16-byte aligned: the per-thread LDGSTS loop was 6 % of the kernel's
  // instructions on the go_9x9 shape -- profiles/r2x_kwide_go_source.md; measured: go_9x9 unchanged, othello +1.7 %)
  bool sb_bulk = false;
  if (sb != nullptr) {
    const int cnt = nfi + 1 < tv.N ? nfi + 1 : tv.N;
    sb_bulk = (((uintptr_t)tv.best) & 15) == 0 && cnt >= 2 && !stale;  // (a table this CTA has just rewritten: generic copies)
    if (sb_bulk) {
      if (tid == 0) {
        const unsigned bytes = (unsigned)(cnt >> 1) * 16u;
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&s_bar, bytes);
        bulk_g2s(sb, tv.best, bytes, &s_bar);
        if (cnt & 1) cp_async8(sb + cnt - 1, tv.best + cnt - 1);
      }
    } else {
      for (int i = tid; i < cnt; i += NT) cp_async8(sb + i, tv.best + i);
    }
  }
  // every thread, before it reads the staged table (a barrier follows at both call sites)
  auto staged_wait = [&]() {
    cp_async_wait_all();
    if (sb_bulk) {
      if (tid == 0) mbar_wait(&s_bar, 0);  // (the others pass the barrier after thread 0 has seen the copy complete)
    }
  };
  int L = 0;             // length of the path this expansion hangs from (levels 0 .. L-1)
  int fresh_node = -1;   // row written by this launch's expand
  if (do_expand) {
    const float* noise = (WEIGHTED && P.w_noise) ? P.w_noise + (size_t)b * F : nullptr;
    L = -Lraw;
    bool record_ok = L >= 1 && L <= P.spill_cap;  // (a length >= 1 is k_sim's ring record: not ours)
    bool spec_ok = record_ok && L <= WIN;          // `rec` is level tid of this path
    s_rec[tid] = rec;
    __syncthreads();
    if (record_ok) {
      const int2 last = spec_ok ? s_rec[L - 1] : lin[L - 1];
      record_ok = last.x == parent && last.y == action;
    }
    if (!record_ok) {
      // ---- the record does not describe this expansion: rebuild it from parents[] / edge_map (warp 0, rare) ----------
      if (warp == 0) {
        int depth = 0;
        for (int x = parent; x != TZ_NULL_INDEX && depth <= tv.N; x = tv.parents[x]) ++depth;
        depth = depth < P.spill_cap ? depth : P.spill_cap;  // (a well-formed tree has depth <= N <= capacity)
        int x = parent, act = action;
        for (int lvl = depth - 1; lvl >= 0; --lvl) {
          if (lane == 0) lin[lvl] = make_int2(x, act);
          const int up = tv.parents[x];
          if (up == TZ_NULL_INDEX) break;
          act = find_action<NC>(tv, up, x, lane);
          if (act == BIG) act = 0;  // (corrupted tree: keep going with a defined value)
          x = up;
        }
        if (lane == 0) s_walk[0] = depth;
        if (lane == 1) s_walk[1] = tv.edge[(unsigned)parent * (unsigned)F + (unsigned)action];
      }
      __syncthreads();  // (also orders warp 0's global stores to lin[] before everyone's loads)
      L = s_walk[0];
      end_child = s_walk[1];
      spec_ok = false;
      __syncthreads();
    }
    const int top = L - 1;
    const unsigned eidx = (unsigned)parent * (unsigned)F + (unsigned)action;
    const bool exists = end_child >= 0;
    const int node = exists ? end_child : (nfi < tv.N ? nfi : -1);  // full tree: nothing is written (tree.py:116-131)
    fresh_node = node;
    TZ_WSTAMP(1);

    // ---- backup + decisions, 32 W levels per pass, deepest pass first --------------------------------------------------
    float cq = value;     // the expanded child's statistics after this expansion (known once round trip 2 is back)
    int cnbits = 1 | (termflag ? TERM_BIT : 0);
    float below_q = 0.0f;  // statistics of the path child one level below the pass (first pass: the expanded child)
    int below_n = 0;
    for (int hi = top; hi >= 0; hi -= WIN) {
      const int lo = hi - (WIN - 1) > 0 ? hi - (WIN - 1) : 0;
      const int cnt = hi - lo + 1;
      const bool on = tid < cnt;
      const bool first_pass = hi == top;
      if (!(first_pass && spec_ok)) {  // the record window is not the speculative one: fetch it (deep paths, rebuilt records)
        if (!first_pass) __syncthreads();  // the previous pass is done with the shared arrays
        if (on) s_rec[tid] = lin[lo + tid];
        __syncthreads();
      }
      // -- round trip 2: the levels' statistics (one thread per level), the expanded child's, every warp's first row
      float qd = 0.0f, rd = 0.0f;
      int nd = 0;
      if (on) {
        rec = s_rec[tid];
        qd = tv.q[rec.x];
        nd = tv.n[rec.x];
        if (WEIGHTED) rd = tv.r[rec.x];
        // (prefetching the level's child_stats row into L2 here was measured 3 % SLOWER on the go_9x9 shape and neutral on
        // othello: a selector call is bound by its own ~370 dependent instructions, not by the row's latency -- r2e_ab.log)
      }
      float q_e = 0.0f;
      int n_e = 0;
      if (first_pass && exists) {
        n_e = tv.n[end_child];
        q_e = tv.q[end_child];
      }
      // decisions: warp sw of the SW scoring warps takes levels j = cnt-1-sw, cnt-1-sw-SW, ... (deepest first)
      constexpr int SW = WEIGHTED ? (W > 1 ? W - 1 : 1) : W;
      const int sw = WEIGHTED ? (W > 1 ? warp - 1 : 0) : warp;
      const bool scorer = !WEIGHTED || W == 1 || warp > 0;
      Row<NC> row, nxt;
      int j = cnt - 1 - sw;
      if (WEIGHTED && warp == 0) {
        load_row<NC, false>(tv, s_rec[cnt - 1].x, lane, row);  // the chain's first row
      } else if (scorer && j >= 0) {
        load_row<NC, true>(tv, s_rec[j].x, lane, row);
      }
      if (first_pass) {  // expand: visit an existing (terminal) child, or add_node (mcts.py:174-187, tree.py:101-132)
        int cn = 1;
        if (exists) {  // visit_node mcts.py:299-336 (only terminal children are re-expanded)
          cq = backup_q(q_e, n_e, value, cfg.fma_backup);
          cn = n_e + 1;
        }
        cnbits = cn | (termflag ? TERM_BIT : 0);
        below_q = cq;
        below_n = cnbits;
      }
      // -- A: one thread per level
      if (on) {
        if (WEIGHTED) {
          s_q1[tid] = qd;
          s_n1[tid] = nd;
          s_r[tid] = rd;
        } else {  // MCTS.backpropagate mcts.py:231-262
          const int k = top - (lo + tid) + 1;  // discounts applied on the way up to this level (mcts.py:247, once per level)
          float v = value;
          if ((cfg.discount == -1.0f || cfg.discount == 1.0f) && value == value) {
            v = (cfg.discount < 0.0f && (k & 1)) ? -value : value;  // products with +-1 are exact
          } else {
            for (int i = 0; i < k; ++i) v = __fmul_rn(v, cfg.discount);
          }
          const float q1 = backup_q(qd, nd, v, cfg.fma_backup);
          s_q1[tid] = q1;
          s_n1[tid] = nd + 1;
          tv.q[rec.x] = q1;
          tv.n[rec.x] = nd + 1;
        }
      }
      if (tid == 0) s_done = 0;
      if (first_pass && sb != nullptr) staged_wait();  // the staged best-table has landed
      __syncthreads();
      if (first_pass) TZ_WSTAMP(2);
      // -- B: one warp per level (weighted: warp 0 runs the backup chain, the others score behind it)
      if (WEIGHTED && warp == 0) {
        float bq = below_q;
        int bn = below_n;
        for (int jj = cnt - 1; jj >= 0; --jj) {
          const int2 r = s_rec[jj];
          nxt = row;
          if (jj >= 1) load_row<NC, false>(tv, s_rec[jj - 1].x, lane, nxt);
          patch_stats<NC>(row, r.y, lane, bq, bn, lo + jj < top || node >= 0);
          const float qX = s_q1[jj], rX = s_r[jj];
          const int nX = s_n1[jj];
          const float qw = weighted_value<NC>(row, F, cfg, qX, lane, noise);  // weighted_mcts.py:102-137
          const float q1 = backup_q(qw, nX, rX, cfg.fma_backup);               // :139-142
          if (lane == 0) {
            s_q1[jj] = q1;
            s_n1[jj] = nX + 1;
            asm volatile("fence.acq_rel.cta;" ::: "memory");  // release: the two stores above before the flag below
            s_done = cnt - jj;
            tv.q[r.x] = q1;
            tv.n[r.x] = nX + 1;
            if (lo + jj >= 1) {
              const int2 up = jj >= 1 ? s_rec[jj - 1] : lin[lo - 1];
              cs_set_stats(tv, (unsigned)up.x * (unsigned)F + (unsigned)up.y, q1, nX + 1);
            }
          }
          bq = q1;
          bn = nX + 1;
          row = nxt;
        }
        if (first_pass) TZ_WSTAMP(7);
      }
      if (scorer) {
        for (; j >= 0; j -= SW) {
          const int2 r = s_rec[j];
          nxt = row;
          if (j - SW >= 0) load_row<NC, true>(tv, s_rec[j - SW].x, lane, nxt);
          if (WEIGHTED && W > 1) {
            while (s_done < cnt - j) { }  // the chain has published this level (and the one below it)
            asm volatile("fence.acq_rel.cta;" ::: "memory");  // acquire
          }
          float pq;
          int pnb;
          if (j == cnt - 1) {
            pq = below_q;
            pnb = below_n;
          } else {
            pq = s_q1[j + 1];
            pnb = s_n1[j + 1];
          }
          const bool is_top = lo + j == top;
          patch_stats<NC>(row, r.y, lane, pq, pnb, !is_top || node >= 0);
          patch_edge<NC>(row, r.y, lane, node, is_top && node >= 0);
          const int2 e = select_entry<NC, SEL>(row, F, cfg, s_q1[j], s_n1[j], lane);
          if (lane == 0) {
            s_best[j] = e;
            tv.best[r.x] = e;
            if (sb != nullptr) sb[r.x] = e;
          }
          row = nxt;
        }
      }
      if (first_pass && node >= 0 && warp == XW) {
        // the expansion's writes: after this warp's share of the decisions (it has none unless the pass has >= W levels)
        int new_bx = -1, new_by = -1;
        if (!exists) {  // the new node's own selector decision
          const int2 e = fresh_entry<NC, SEL>(pol, F, cfg, cq, lane);
          new_bx = e.x;
          new_by = e.y;
        }
        if (lane == 0) {
          if (!exists) {  // new_node mcts.py:339-360 / weighted_mcts.py:43-63
            tv.parents[node] = parent;
            tv.edge[eidx] = node;
            *tv.nfi = nfi + 1;
            if (tv.r) tv.r[node] = value;
          }
          tv.q[node] = cq;
          tv.n[node] = (cnbits & BIG);
          tv.term[node] = (uint8_t)termflag;
          if (!exists) cs_set_edge(tv, eidx, node);
          tv.best[node] = make_int2(new_bx, new_by);  // (unknown for a re-expanded child: its p row changes)
          if (sb != nullptr) sb[node] = make_int2(new_bx, new_by);
        }
        const unsigned prow = (unsigned)node * (unsigned)F + (unsigned)lane;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if (c * 32 + lane < F) {
            tv.p[prow + c * 32] = pol[c];
            if (exists) cs_set_p(tv, prow + c * 32, pol[c]);
            else tv.cs[prow + c * 32] = make_int4(0, 0, __float_as_int(pol[c]), -1);
          }
        }
      }
      __syncthreads();
      // the parents on the path mirror their children's new statistics (tree.py:78-98 materialised) -- after the decisions,
      // whose row loads these stores must not race
      if (first_pass && node >= 0 && tid == 0) cs_set_stats(tv, eidx, cq, cnbits);
      if (!WEIGHTED && on && lo + tid >= 1) {
        const int2 up = tid >= 1 ? s_rec[tid - 1] : lin[lo - 1];
        cs_set_stats(tv, (unsigned)up.x * (unsigned)F + (unsigned)up.y, s_q1[tid], s_n1[tid]);
      }
      below_q = s_q1[0];
      below_n = s_n1[0];
    }
  }
  TZ_WSTAMP(3);
  if (cfg.programmatic & 2) pdl_launch_dependents();
  if (!do_sel) {  // expand-only launch (last simulation of a search): just store the new node's embedding
    if (fresh_node >= 0) {
      for (int k = 0; k < P.n_emb; ++k) {
        const SimLeaf& lf = k < SIM_LEAVES_INLINE ? P.leaf[k] : X.leaf[k - SIM_LEAVES_INLINE];
        block_copy2(lf.table + ((size_t)b * tv.N + (size_t)fresh_node) * lf.rb, lf.fresh + (size_t)b * lf.rb, nullptr, nullptr, lf.rb,
                    tid, NT);
      }
    }
    if (warp == 0) tl_max(P.tl_row, 2, lane);
    return;
  }

  // ---- C: MCTS.traverse mcts.py:192-228 -------------------------------------------------------------------------------
  // the new walk follows the old path as long as every new decision leads to the old next node; the first level where it
  // does not (level `top` always does not) is found by a vote over all levels
  int k = -1;  // the level the sequential walk starts from: its decision is known (entry kn)
  int2 kn = make_int2(-1, -1);
  int knode = TZ_ROOT_INDEX;
  if (do_expand && L >= 1) {
    const int top = L - 1;
    int first = BIG;
    if (L <= WIN) {  // everything is still in shared memory (index = level)
      if (tid <= top) {
        const bool leaves = !(tid < top && s_best[tid].y == s_rec[tid + 1 < WIN ? tid + 1 : tid].x);
        if (leaves) first = tid;
      }
    } else {
      for (int lvl = tid; lvl <= top; lvl += NT) {
        const int2 r = lin[lvl];
        const int nxt_old = lvl < top ? lin[lvl + 1].x : -1;
        const int2 e = tv.best[r.x];
        if (!(lvl < top && e.y == nxt_old)) {
          first = lvl;
          break;  // (levels are visited in increasing order per thread)
        }
      }
    }
    const int wfirst = __reduce_min_sync(FULL, first);
    if (lane == 0) s_wmin[warp] = wfirst;
    __syncthreads();
    k = s_wmin[0];
#pragma unroll
    for (int w = 1; w < W; ++w) k = min(k, s_wmin[w]);
    if (warp == 0) {
      if (L <= WIN) {
        knode = s_rec[k].x;
        kn = s_best[k];
      } else {
        knode = lin[k].x;
        kn = tv.best[knode];
      }
    }
  }
  TZ_WSTAMP(4);
  if (!do_expand && sb != nullptr) {  // select-only launch: nothing waited for the staged table yet
    staged_wait();
    __syncthreads();
  }
  if (warp == 0) {
    int node = TZ_ROOT_INDEX, levels = 0, sel_action = 0, stop_child = -1;
    int cur = TZ_ROOT_INDEX;
    bool walking = true;
    if (k >= 0) {
      node = knode;
      sel_action = kn.x;
      levels = k + 1;
      if (lane == 0) lin[k] = make_int2(knode, kn.x);  // the level keeps its node; the action taken there is the new decision
      if (kn.y < 0) {  // cond_fn mcts.py:208-213: no edge (-1), or the child is terminal (-(2 + child))
        stop_child = kn.y == -1 ? -1 : -(kn.y + 2);
        walking = false;
      } else {
        cur = kn.y;
      }
    }
    while (walking) {
      int2 e = sb != nullptr ? sb[cur] : tv.best[cur];  // the one dependent load of this level
      if (e.x < 0) {  // unknown: score the node here (PUCTSelector.__call__) and remember the decision
        Row<NC> row;
        load_row<NC, true>(tv, cur, lane, row);
        const float nq = tv.q[cur];
        const int nn = tv.n[cur];
        e = select_entry<NC, SEL>(row, F, cfg, nq, nn, lane);
        if (lane == 0) tv.best[cur] = e;  // (the staged copy is not read again at this node)
      }
      node = cur;
      sel_action = e.x;
      if (lane == 0 && levels < P.spill_cap) lin[levels] = make_int2(cur, e.x);
      ++levels;
      if (e.y < 0) {
        stop_child = e.y == -1 ? -1 : -(e.y + 2);
        break;
      }
      if (levels > tv.N) {  // never spin on a corrupted tree
        stop_child = e.y;
        break;
      }
      cur = e.y;
    }
    if (lane == 0) {
      P.w_parent[b] = node;
      P.w_action[b] = sel_action;
      path[PATH_LEN] = -levels;  // negated: this kernel's linear record, not k_sim's ring
      path[PATH_END] = stop_child;
      s_walk[0] = node;
      if (P.stats) {  // fire-and-forget reductions (RED)
        atomicAdd(reinterpret_cast<unsigned long long*>(P.stats) + 4 * (size_t)b + 0, (unsigned long long)levels);
        atomicAdd(reinterpret_cast<unsigned long long*>(P.stats) + 4 * (size_t)b + 1, 1ull);
      }
    }
  }
  __syncthreads();
  TZ_WSTAMP(5);
  // ---- D: embeddings.  store: the expanded node's rows emb[k][b, fresh_node] <- w.emb_new[k][b] (mcts.py:354-360);
  //      gather: the next parent's rows w.emb_parent[k][b] <- emb[k][b, node] (mcts.py:161-164; a node written by this very
  //      launch is read back from the caller's buffer) -- all loads of both copies in flight together ------------------------
  const int pnode = s_walk[0];
  for (int kk = 0; kk < P.n_emb; ++kk) {
    const SimLeaf& lf = kk < SIM_LEAVES_INLINE ? P.leaf[kk] : X.leaf[kk - SIM_LEAVES_INLINE];
    const uint8_t* fresh = lf.fresh + (size_t)b * lf.rb;
    uint8_t* tbl = lf.table + (size_t)b * tv.N * lf.rb;
    const uint8_t* src = pnode == fresh_node ? fresh : tbl + (size_t)pnode * lf.rb;
    block_copy2(lf.parent_out + (size_t)b * lf.rb, src, fresh_node >= 0 ? tbl + (size_t)fresh_node * lf.rb : nullptr, fresh, lf.rb, tid, NT);
  }
  TZ_WSTAMP(6);
#ifdef TZ_PROFILE
  if (tid == 0 && b < 4096) g_prof_gt[16 * b + 8] = L;
#endif
  if (warp == 0) tl_max(P.tl_row, 2, lane);
}

// ---- k_sim_wide dispatch: W warps per tree ------------------------------------------------------------------------------
template <typename K>
int launch_wide_k(K kernel, const SimLaunch& L, int W, cudaStream_t s) {
  if (use_pdl(L)) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)L.P.B);
    lc.blockDim = dim3(32 * W);
    lc.dynamicSmemBytes = L.smem;
    lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&lc, kernel, L.P, L.X);
    count_launch();
    return e == cudaSuccess ? TZ_OK : (int)e;
  }
  kernel<<<L.P.B, 32 * W, L.smem, s>>>(L.P, L.X);
  return launch_status();
}

template <int NC, bool WEIGHTED, int SEL>
int launch_wide_w(const SimLaunch& L, int W, cudaStream_t s) {
  if constexpr (NC <= 4) {
    if (W == 2) return launch_wide_k(k_sim_wide<NC, WEIGHTED, SEL, 2>, L, 2, s);
    if (W == 8) return launch_wide_k(k_sim_wide<NC, WEIGHTED, SEL, 8>, L, 8, s);
  }
  return launch_wide_k(k_sim_wide<NC, WEIGHTED, SEL, 4>, L, 4, s);
}

template <bool WEIGHTED>
int launch_wide_any(const SimLaunch& L, int nc, int W, cudaStream_t s) {
  const bool mz = L.P.cfg.selector == TZ_SEL_MUZERO_PUCT;
#define TZ_WIDE(NC_)                                                                  \
  do {                                                                                \
    if (mz) return launch_wide_w<NC_, WEIGHTED, TZ_SEL_MUZERO_PUCT | SELQ_RUNTIME>(L, W, s); \
    return launch_wide_w<NC_, WEIGHTED, TZ_SEL_PUCT | SELQ_RUNTIME>(L, W, s);         \
  } while (0)
  if (nc <= 1) TZ_WIDE(1);
  if (nc <= 2) TZ_WIDE(2);
  if (nc <= 3) TZ_WIDE(3);
  if (nc <= 4) TZ_WIDE(4);
  if (nc <= 8) TZ_WIDE(8);
  TZ_WIDE(16);
#undef TZ_WIDE
}

}  // namespace

#endif  // TZ_WIDE_CUH_
