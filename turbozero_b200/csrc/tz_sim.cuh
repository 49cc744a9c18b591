// tz_sim.cuh -- k_sim: the per-simulation kernel with ONE WARP per tree ([expand + backprop of simulation i] [select of
// simulation i+1]) and its launcher template; instantiated per register-chunk count in tz_sim_nc*.cu (parallel builds).
#ifndef TZ_SIM_CUH_
#define TZ_SIM_CUH_

#include "tz_device.cuh"

namespace {

// (Prefetching, during the walk, the rows the NEXT launch will read -- prefetch.global.L2 of q / n / the child_stats row of
// every node the walk passes -- was measured 14 % SLOWER on configs[1] (profiles/r2m_variants.log): the walk is the kernel's
// critical path and the prefetches queue ahead of its own loads.)
// (Four lanes per path level -- quads scoring eight levels per pass, min / max / argmax by two xor-shuffle steps -- was built
// and measured in round 2: the median warp's decisions fell from 0.90 to 0.67 us, but a deep path needs up to four passes
// (1.18 us for the slowest warps, which end the launch) and issuing the rows quad-wise cost 0.4 us more: 117.2 instead of
// 119.4 M simulations/s on configs[1].  Reverted; profiles/r2p_variants.log, r2q_phase_warps.log.)
// FM = 4 / 8 / 16: narrow plain-MCTS trees (F <= FM), decisions scored one lane per path level (narrow_select);
// FM = 0: one lane per child with NC register chunks per lane, U levels side by side (any F, and the weighted
// variant, whose levels are sequential).
template <int NC, bool WEIGHTED, int SEL, int FM, bool PDL>
__global__ void __launch_bounds__(SIM_THREADS) k_sim(const __grid_constant__ SimP P, const __grid_constant__ SimLeafExtra X) {
  static_assert(FM == 0 || (NC == 1 && !WEIGHTED), "the lane-per-level pass is for narrow plain-MCTS trees");
  extern __shared__ __align__(16) uint8_t sim_smem[];
  const int b = (int)((blockIdx.x * (unsigned)SIM_THREADS + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= P.B) return;  // whole warps only
  const int F = P.F;
  const int mode = P.mode;
  const bool do_expand = (mode & MODE_EXPAND) != 0, do_sel = (mode & MODE_SELECT) != 0;
  const TzSearchCfg& cfg = P.cfg;
  constexpr int U = Chunk<NC>::U;
  constexpr bool NARROW = FM > 0;

  TZ_STAMP(0);
  TZ_TL_MIN(P.pad0, 0);
  if (mode & MODE_TIMELINE) tl_min(P.tl_row, 0, lane);
#ifdef TZ_PROFILE
  const long long prof_t0 = prof_gtime();
#endif
  // ---- round trip 1: everything whose address is known at entry.  With TzSearchCfg.programmatic the tree state and
  //      the previous select's outputs (written by EARLIER tz launches) are read -- and round trip 2 is issued -- while
  //      the user's leaf kernel is still executing; the leaf results are read after griddepcontrol.wait. ----------------
  constexpr bool pdl = PDL;  // TzSearchCfg.programmatic, resolved at launch: the ordinary launch carries none of it
  if constexpr (pdl) {
    if (!do_expand) pdl_wait();  // select-only launch: the preceding kernel may still be writing this tree
    else __threadfence();        // acquire: drop L1 lines this SM may hold from before the last tz launch on this tree
  }
  int parent = 0, action = 0, termflag = 0, nfi = 0, L = 0, pn = -1, pa = 0, end_child = -1;
  float value = 0.0f;
  float pol[NC];
  uint4 pre[SIM_LEAVES_INLINE];  // the new embedding rows of the register-path leaves (see SimP.fast_mask)
  int32_t* const path = P.w_path ? P.w_path + (size_t)b * PATH_STRIDE : nullptr;
  int4 s0, s1;
  if constexpr (!pdl) {
    nfi = P.nfi[b];
    s0 = *reinterpret_cast<const int4*>(P.sel + (size_t)b * TZ_SEL_STATE_WORDS);
    s1 = *reinterpret_cast<const int4*>(P.sel + (size_t)b * TZ_SEL_STATE_WORDS + 4);
    if (do_expand) {
      parent = P.w_parent[b];
      action = P.w_action[b];
      value = P.w_value[b];
      termflag = P.w_term[b] ? 1 : 0;
      if (path) {
        L = path[PATH_LEN];
        end_child = path[PATH_END];
        pn = path[lane];
        pa = path[PATH_ACT + lane];
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) pol[c] = (c * 32 + lane < F) ? P.w_policy[(size_t)b * F + c * 32 + lane] : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < SIM_LEAVES_INLINE; ++k) {
      pre[k] = make_uint4(0u, 0u, 0u, 0u);
      if (do_expand && ((P.fast_mask >> k) & 1) && lane * 16 < (int)P.leaf[k].rb)
        pre[k] = reinterpret_cast<const uint4*>(P.leaf[k].fresh + (size_t)b * P.leaf[k].rb)[lane];
    }
  } else {  // programmatic launch: only what EARLIER tz launches wrote; the leaf results follow griddepcontrol.wait
    nfi = P.nfi[b];
    s0 = *reinterpret_cast<const int4*>(P.sel + (size_t)b * TZ_SEL_STATE_WORDS);
    s1 = *reinterpret_cast<const int4*>(P.sel + (size_t)b * TZ_SEL_STATE_WORDS + 4);
    if (do_expand) {
      parent = P.w_parent[b];
      action = P.w_action[b];
      if (path) {
        L = path[PATH_LEN];
        end_child = path[PATH_END];
        pn = path[lane];
        pa = path[PATH_ACT + lane];
      }
    }
#pragma unroll
    for (int k = 0; k < SIM_LEAVES_INLINE; ++k) pre[k] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int c = 0; c < NC; ++c) pol[c] = 0.0f;
  }
  // programmatic launch: the leaf results of this simulation (written by the user's kernels), once they are complete
  auto load_leaf_results = [&]() {
    pdl_wait();
    value = P.w_value[b];
    termflag = P.w_term[b] ? 1 : 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) pol[c] = (c * 32 + lane < F) ? P.w_policy[(size_t)b * F + c * 32 + lane] : 0.0f;
#pragma unroll
    for (int k = 0; k < SIM_LEAVES_INLINE; ++k) {
      if (((P.fast_mask >> k) & 1) && lane * 16 < (int)P.leaf[k].rb)
        pre[k] = reinterpret_cast<const uint4*>(P.leaf[k].fresh + (size_t)b * P.leaf[k].rb)[lane];
    }
  };
  const TV tv = make_view(P, b);
  int2* const sb = P.best_rows > 0 ? reinterpret_cast<int2*>(sim_smem) + (size_t)(threadIdx.x >> 5) * P.best_rows : nullptr;
  TZ_STAMP(8);
  {  // the best-table is only valid for the selector parameters it was computed with
    const bool stale = s0.x != cfg.selector || s0.y != __float_as_int(cfg.c) || s0.z != __float_as_int(cfg.c1) ||
                       s0.w != __float_as_int(cfg.c2) || s1.x != __float_as_int(cfg.epsilon) ||
                       s1.y != __float_as_int(cfg.discount) || s1.z != cfg.q_transform;
    if (stale) {  // (uniform: every lane read the same words)
      for (int i = lane; i < nfi && i < tv.N; i += 32) tv.best[i] = make_int2(-1, -1);
      if (lane == 0) {
        *reinterpret_cast<int4*>(tv.sel) =
            make_int4(cfg.selector, __float_as_int(cfg.c), __float_as_int(cfg.c1), __float_as_int(cfg.c2));
        *reinterpret_cast<int4*>(tv.sel + 4) = make_int4(__float_as_int(cfg.epsilon), __float_as_int(cfg.discount), cfg.q_transform, 0);
      }
      __syncwarp();
    }
  }
  TZ_STAMP(9);

  // state handed from the expand / backprop phase to the walk
  int my_bx = -1, my_by = -1;  // lane d: best-table entry of path level d (levels lowest..top of the ring)
  bool ring = false;           // the path ring describes this expansion: levels (top - 32, top] are in pn / pa
  bool sb_live = false;        // the shared-memory copy of the best-table is complete and current
  int top = -1, lowest = 0;
  int fresh_node = -1;         // row written by this launch's expand
  int new_bx = -1, new_by = -1;  // its best-table entry, if it is a new node

  // Path levels older than the ring (TzWork.path_spill).  Evaluated lazily, inside the rare deep-path branches only: the
  // fields sit on a parameter-bank line of their own, whose first touch the common launch must not pay for.
  auto spill_ptr = [&]() -> int2* { return P.w_spill ? P.w_spill + (size_t)b * P.spill_cap : nullptr; };
  // a deep plain backup can use the spilled levels (deep_windows) instead of chasing parents[]
  auto deep_ok = [&]() -> bool { return !WEIGHTED && P.w_spill != nullptr && L - TZ_PATH_CAP <= P.spill_cap; };
  if (do_expand) {
    top = L - 1;
    ring = path != nullptr && L >= 1 && __shfl_sync(FULL, pn, top & 31) == parent &&
           __shfl_sync(FULL, pa, top & 31) == action;  // trusted only if its deepest entry is this expansion
    const unsigned eidx = (unsigned)parent * (unsigned)F + (unsigned)action;
    const float* noise = (WEIGHTED && P.w_noise) ? P.w_noise + (size_t)b * F : nullptr;
    if (ring) {
      lowest = top - (TZ_PATH_CAP - 1) > 0 ? top - (TZ_PATH_CAP - 1) : 0;
      const int d = top - ((top - lane) & 31);  // depth held by this lane (d % 32 == lane, top-32 < d <= top)
      const bool on_path = d >= 0;
      // ---- round trip 2: every path node's statistics (one lane per level), the expanded child if it exists, and
      //      the rows of the deepest path nodes ------------------------------------------------------------------
      float qd = 0.0f, rd = 0.0f;
      int nd = 0;
      if (on_path) {
        qd = tv.q[pn];
        nd = tv.n[pn];
        if (WEIGHTED) rd = tv.r[pn];
      }
      const bool exists = end_child >= 0;
      float q_e = 0.0f;
      int n_e = 0;
      if (exists) {
        n_e = tv.n[end_child];
        q_e = tv.q[end_child];
      }
      TZ_STAMP(10);
      int4 h[NARROW ? FM : 1];  // narrow: this lane's path node's whole child_stats row
      Row<NC> rows[U];
      if constexpr (NARROW) {
        const int4* hrow = tv.cs + (unsigned)(on_path ? pn : 0) * (unsigned)F;
#pragma unroll
        for (int a = 0; a < FM; ++a) h[a] = (on_path && a < F) ? hrow[a] : make_int4(0, 0, 0, -1);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (top - u >= lowest) load_row<NC, true>(tv, __shfl_sync(FULL, pn, (top - u) & 31), lane, rows[u]);
      }
      // stage the best-table for the walk while the backprop computes (only when every change this launch makes to
      // the table is one the fast path below mirrors: the whole path is in the ring)
      if (sb != nullptr && do_sel && (L <= TZ_PATH_CAP || deep_ok())) {
        const int cnt = nfi + 1 < tv.N ? nfi + 1 : tv.N;
        if ((((uintptr_t)tv.best | (uintptr_t)sb) & 15) == 0) {
          const int pairs = cnt >> 1;
#pragma unroll 1
          for (int i = lane; i < pairs; i += 32) cp_async16(sb + 2 * i, tv.best + 2 * i);
          if ((cnt & 1) && lane == 0) cp_async8(sb + cnt - 1, tv.best + cnt - 1);
        } else {
#pragma unroll 1
          for (int i = lane; i < cnt; i += 32) cp_async8(sb + i, tv.best + i);
        }
        sb_live = true;
      }
      TZ_STAMP(1);
      if constexpr (pdl) load_leaf_results();
      TZ_TL_MAX(P.pad0, 1);
      if (mode & MODE_TIMELINE) tl_max(P.tl_row, 1, lane);

      // ---- expand: visit an existing (terminal) child, or add_node (mcts.py:174-187, tree.py:101-132) ------------
      const int node = exists ? end_child : (nfi < tv.N ? nfi : -1);  // full tree: nothing is written (tree.py:116-131)
      float cq = value;  // the child's statistics after this expansion
      int cn = 1;
      if (exists) {  // visit_node mcts.py:299-336 (only terminal children are re-expanded)
        cq = backup_q(q_e, n_e, value, cfg.fma_backup);
        cn = n_e + 1;
      }
      const int cnbits = cn | (termflag ? TERM_BIT : 0);
      if (node >= 0) {
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
          tv.n[node] = cn;
          tv.term[node] = (uint8_t)termflag;
          cs_set_stats(tv, eidx, cq, cnbits);
          if (!exists) cs_set_edge(tv, eidx, node);
          tv.best[node] = make_int2(new_bx, new_by);  // (unknown for a re-expanded child: its p row changes)
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
        fresh_node = node;
      }

      // ---- per-level quantities, one lane per level -------------------------------------------------------------
      const int n1 = nd + 1;
      const float sq1 = sqrt_count(n1);
      const float scale1 = explore_scale<SEL>(cfg, n1);
      float q1 = 0.0f;
      if (!WEIGHTED && on_path) {  // MCTS.backpropagate mcts.py:231-262: all ring levels at once
        const int k = top - d + 1;  // discounts applied on the way up to this level (mcts.py:247, once per level)
        float v = value;
        if ((cfg.discount == -1.0f || cfg.discount == 1.0f) && value == value) {
          v = (cfg.discount < 0.0f && (k & 1)) ? -value : value;  // products with +-1 are exact
        } else {
          for (int j = 0; j < k; ++j) v = __fmul_rn(v, cfg.discount);
        }
        q1 = backup_q(qd, nd, v, cfg.fma_backup);
      }
      TZ_STAMP(2);

      // ---- every path node's selector decision with the statistics it will have when the next walk arrives
      //      (weighted: preceded by the node's backup, deepest level first) ----------------------------------------
      if constexpr (NARROW) {
        // the child this path went through at this level, with its statistics as of now: from the lane one level down
        const float pq_up = __shfl_sync(FULL, q1, (lane + 1) & 31);
        const int pnb_up = __shfl_sync(FULL, n1, (lane + 1) & 31);
        const bool is_top = d == top;
        const float pq = is_top ? cq : pq_up;
        const int pnb = is_top ? cnbits : pnb_up;
        const bool patch = on_path && (!is_top || node >= 0);
#pragma unroll
        for (int a = 0; a < FM; ++a) {
          if (patch && a == pa) {
            h[a].x = __float_as_int(pq);
            h[a].y = pnb;
            if (is_top) h[a].w = node;
          }
        }
        bool unsafe = false;
        int act = narrow_select<FM, SEL, false>(h, F, cfg, q1, sq1, scale1, unsafe);
        if (__any_sync(FULL, on_path && unsafe))  // rare: operands outside div_core's proven range -> hardware division
          act = narrow_select<FM, SEL, true>(h, F, cfg, q1, sq1, scale1, unsafe);
        int child = h[0].w, cnb = h[0].y;
#pragma unroll
        for (int a = 1; a < FM; ++a) {
          if (a == act) {
            child = h[a].w;
            cnb = h[a].y;
          }
        }
        if (on_path) {  // best-table entry (see TzTree.best)
          my_bx = act;
          my_by = child < 0 ? -1 : (cnb < 0 ? -(child + 2) : child);
        }
      } else {
        float below_q = cq;  // weighted: statistics of the path child one level down, as of now
        int below_n = cnbits;
        Row<NC> ahead[U];  // the rows of the NEXT pass, loaded while this one is scored (a tree this wide is rarely in L2)
        for (int hi = top; hi >= lowest; hi -= U) {
          if (hi != top) {
#pragma unroll
            for (int u = 0; u < U; ++u) rows[u] = ahead[u];
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            ahead[u] = rows[u];
            if (hi - U - u >= lowest) load_row<NC, true>(tv, __shfl_sync(FULL, pn, (hi - U - u) & 31), lane, ahead[u]);
          }
          bool unsafe = false;
          int act_u[U];
          float nq_u[U], sq_u[U], sc_u[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int lvl = hi - u;
            act_u[u] = 0;
            nq_u[u] = sq_u[u] = sc_u[u] = 0.0f;
            if (lvl >= lowest) {
              const int sl = lvl & 31;
              const int a_here = __shfl_sync(FULL, pa, sl);
              float pq;
              int pnb;
              if (WEIGHTED) {
                pq = below_q;
                pnb = below_n;
              } else if (lvl == top) {
                pq = cq;
                pnb = cnbits;
              } else {
                pq = __shfl_sync(FULL, q1, (lvl + 1) & 31);
                pnb = __shfl_sync(FULL, n1, (lvl + 1) & 31);
              }
              patch_stats<NC>(rows[u], a_here, lane, pq, pnb, lvl < top || node >= 0);
              patch_edge<NC>(rows[u], a_here, lane, node, lvl == top && node >= 0);
              if (WEIGHTED) {  // weighted_mcts.py:102-142
                const float qX = __shfl_sync(FULL, qd, sl), rX = __shfl_sync(FULL, rd, sl);
                const int nX = __shfl_sync(FULL, nd, sl);
                const float qw = weighted_value<NC>(rows[u], F, cfg, qX, lane, noise);
                const float qn1 = backup_q(qw, nX, rX, cfg.fma_backup);
                if (lane == sl) q1 = qn1;
                below_q = qn1;
                below_n = nX + 1;
                nq_u[u] = qn1;
              } else {
                nq_u[u] = __shfl_sync(FULL, q1, sl);
              }
              sq_u[u] = __shfl_sync(FULL, sq1, sl);
              sc_u[u] = __shfl_sync(FULL, scale1, sl);
              act_u[u] = select_core<NC, SEL, false>(rows[u], F, cfg, nq_u[u], sq_u[u], sc_u[u], lane, unsafe);
            }
          }
          if (__any_sync(FULL, unsafe)) {  // rare: operands outside div_core's proven range -> hardware division
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (hi - u >= lowest) act_u[u] = select_core<NC, SEL, true>(rows[u], F, cfg, nq_u[u], sq_u[u], sc_u[u], lane, unsafe);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (hi - u >= lowest) {
              const int2 e = make_entry<NC>(rows[u], act_u[u]);
              if (lane == ((hi - u) & 31)) {
                my_bx = e.x;
                my_by = e.y;
              }
            }
          }
        }
      }
      TZ_STAMP(3);
      // ---- stores, one lane per level ------------------------------------------------------------------------
      const int ppn = __shfl_sync(FULL, pn, (lane + 31) & 31);  // the parent on the path mirrors this node's statistics
      const int ppa = __shfl_sync(FULL, pa, (lane + 31) & 31);
      if (on_path) {
        tv.q[pn] = q1;
        tv.n[pn] = n1;
        tv.best[pn] = make_int2(my_bx, my_by);
        if (d >= 1 && d > top - (TZ_PATH_CAP - 1)) cs_set_stats(tv, (unsigned)ppn * (unsigned)F + (unsigned)ppa, q1, n1);
      }
      if (sb_live) {  // mirror this launch's best-table writes into the staged copy (after the copy has landed)
        cp_async_wait_all();
        __syncwarp();
        if (on_path) sb[pn] = make_int2(my_bx, my_by);
        if (lane == 0 && node >= 0) sb[node] = make_int2(new_bx, new_by);
      }
      if (L > TZ_PATH_CAP) {  // deeper than the ring: continue above its shallowest entry by chasing parents[]
        const int sl = lowest & 31;
        const int Xn = __shfl_sync(FULL, pn, sl);
        const float qx = __shfl_sync(FULL, q1, sl);
        const int nx = __shfl_sync(FULL, n1, sl);
        if (!WEIGHTED && deep_ok()) {
          deep_windows<NC, SEL>(tv, cfg, lane, spill_ptr(), lowest, qx, nx, value, top, sb_live ? sb : nullptr);
        } else if (!WEIGHTED) {
          float val = value;
          for (int j = 0; j < TZ_PATH_CAP; ++j) val = __fmul_rn(val, cfg.discount);
          walk_up<NC>(tv, cfg, lane, Xn, qx, nx, val);
        } else {
          const int up = tv.parents[Xn];
          if (up != TZ_NULL_INDEX) {
            const int up_a = find_action<NC>(tv, up, Xn, lane);
            if (lane == 0 && up_a != BIG) cs_set_stats(tv, (unsigned)up * (unsigned)F + (unsigned)up_a, qx, nx);
            weighted_walk_up<NC>(tv, cfg, lane, up, up_a != BIG, up_a, qx, nx, noise);
          }
        }
      }
    } else {
      // ---- no usable path ring (TzWork.path == NULL, or parent / action were not produced by the last select):
      //      look the edge up, chase parents[], and leave the changed nodes' best-table entries unknown ----------
      if constexpr (pdl) load_leaf_results();
      const int enode = tv.edge[eidx];
      const bool exists = enode >= 0;
      const int node = exists ? enode : (nfi < tv.N ? nfi : -1);
      float cq = value;
      int cn = 1;
      if (exists) {
        const int n0 = tv.n[enode];
        cq = backup_q(tv.q[enode], n0, value, cfg.fma_backup);
        cn = n0 + 1;
      }
      const int cnbits = cn | (termflag ? TERM_BIT : 0);
      if (node >= 0) {
        if (lane == 0) {
          if (!exists) {
            tv.parents[node] = parent;
            tv.edge[eidx] = node;
            *tv.nfi = nfi + 1;
            if (tv.r) tv.r[node] = value;
          }
          tv.q[node] = cq;
          tv.n[node] = cn;
          tv.term[node] = (uint8_t)termflag;
          cs_set_stats(tv, eidx, cq, cnbits);
          if (!exists) cs_set_edge(tv, eidx, node);
          tv.best[node] = make_int2(-1, -1);
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
        fresh_node = node;
      }
      if (!WEIGHTED) {
        const float val = __fmul_rn(value, cfg.discount);
        const int n0 = tv.n[parent];
        const float q1 = backup_q(tv.q[parent], n0, val, cfg.fma_backup);
        if (lane == 0) {
          tv.q[parent] = q1;
          tv.n[parent] = n0 + 1;
          tv.best[parent] = make_int2(-1, -1);
        }
        walk_up<NC>(tv, cfg, lane, parent, q1, n0 + 1, val);
      } else {
        weighted_walk_up<NC>(tv, cfg, lane, parent, node >= 0, action, cq, cnbits, noise);
      }
    }
    // the expanded node's embedding rows (register-path leaves): mcts.py:354-360
    if (fresh_node >= 0) {
#pragma unroll
      for (int k = 0; k < SIM_LEAVES_INLINE; ++k) {
        if (((P.fast_mask >> k) & 1) && lane * 16 < (int)P.leaf[k].rb)
          reinterpret_cast<uint4*>(P.leaf[k].table + ((size_t)b * tv.N + (size_t)fresh_node) * P.leaf[k].rb)[lane] = pre[k];
      }
    }
    __syncwarp();  // orders this warp's tree writes before the walk's loads below
  }
  // From here on only the walk and the embedding gather remain: let the next kernel in the stream be scheduled now, so
  // that its launch latency overlaps them (it still waits for this whole grid before touching our outputs).  Not
  // earlier: a dependent grid that is resident and waiting for long takes issue slots and CTA slots from this one.
  if constexpr (pdl) {
    if (cfg.programmatic & 2) pdl_launch_dependents();
  }
  if (!do_sel) {  // expand-only launch (last simulation of a search): just store the new node's embedding
    for (int k = 0; k < P.n_emb; ++k) {
      if (k < SIM_LEAVES_INLINE && ((P.fast_mask >> k) & 1)) continue;
      move_leaf(k < SIM_LEAVES_INLINE ? P.leaf[k] : X.leaf[k - SIM_LEAVES_INLINE], b, tv.N, false, 0, fresh_node, lane);
    }
    if (mode & MODE_TIMELINE) tl_max(P.tl_row, 2, lane);
    return;
  }

  // ---- MCTS.traverse mcts.py:192-228: follow the best-table; entries computed above are still in registers -------
  TZ_STAMP(4);
  int cur = TZ_ROOT_INDEX;  // the node whose decision is needed next
  int node = TZ_ROOT_INDEX, levels = 0, sel_action = 0, stop_child = -1;
  int ring_n = -1, ring_a = 0;
  bool walking = true;
  if (do_expand && ring && lowest == 0) {
    // the new walk follows the previous path exactly as long as every decision leads to the old next node: the
    // first level where it does not is found in one vote instead of one step per level
    const int nxt_old = __shfl_sync(FULL, pn, (lane + 1) & 31);
    const bool leaves = lane <= top && !(lane < top && my_by == nxt_old);
    const int k = __ffs(__ballot_sync(FULL, leaves)) - 1;  // 0 <= k <= top (level `top` always leaves)
    if (lane <= k) {
      ring_n = pn;
      ring_a = my_bx;
    }
    node = __shfl_sync(FULL, pn, k);
    sel_action = __shfl_sync(FULL, my_bx, k);
    const int nby = __shfl_sync(FULL, my_by, k);
    levels = k + 1;
    if (nby < 0) {  // cond_fn mcts.py:208-213: no edge (-1), or the child is terminal (-(2 + child))
      stop_child = nby == -1 ? -1 : -(nby + 2);
      walking = false;
    } else {
      cur = nby;
    }
  }
  // The first 32 levels: the ring has room, nothing leaves it.  (The bound doubles as the guard against a corrupted tree:
  // a well-formed one has no path longer than N.)
  const int ring_room = tv.N + 1 < TZ_PATH_CAP ? tv.N + 1 : TZ_PATH_CAP;
  bool ring_full = false;
  while (walking) {
    if (levels >= ring_room) {  // (also when the shared prefix already fills the ring)
      ring_full = true;
      break;
    }
    int bx, by;
    if (cur == fresh_node && new_bx >= 0) {
      bx = new_bx;
      by = new_by;
    } else {
      const int2 e = sb_live ? sb[cur] : tv.best[cur];  // the one dependent load of this level
      bx = e.x;
      by = e.y;
      if (bx < 0) {  // unknown: score the node here (PUCTSelector.__call__) and remember the decision
        Row<NC> row;
        load_row<NC, true>(tv, cur, lane, row);
        const float nq = tv.q[cur];
        const int nn = tv.n[cur];
        const int2 e2 = select_entry<NC, SEL>(row, F, cfg, nq, nn, lane);
        bx = e2.x;
        by = e2.y;
        if (lane == 0) tv.best[cur] = e2;
      }
    }
    node = cur;
    sel_action = bx;
    if (lane == (levels & 31)) {
      ring_n = cur;
      ring_a = bx;
    }
    ++levels;
    if (by < 0) {
      stop_child = by == -1 ? -1 : -(by + 2);
      break;
    }
    cur = by;
  }
  if (ring_full && levels > tv.N) stop_child = cur;  // the corrupted-tree guard (N < 32): stop where we are
  if (ring_full && levels <= tv.N) {
    // Deeper than the ring (rare): every further level pushes level (levels - 32) out of it, into TzWork.path_spill when
    // the caller provided one, so that the backup of this path can process 32 levels per round trip (deep_windows).
    int2* const spill = P.w_spill ? P.w_spill + (size_t)b * P.spill_cap : nullptr;
    for (;;) {
      int bx, by;
      if (cur == fresh_node && new_bx >= 0) {
        bx = new_bx;
        by = new_by;
      } else {
        const int2 e = sb_live ? sb[cur] : tv.best[cur];
          bx = e.x;
        by = e.y;
        if (bx < 0) {
          Row<NC> row;
          load_row<NC, true>(tv, cur, lane, row);
          const float nq = tv.q[cur];
          const int nn = tv.n[cur];
          const int2 e2 = select_entry<NC, SEL>(row, F, cfg, nq, nn, lane);
          bx = e2.x;
          by = e2.y;
          if (lane == 0) tv.best[cur] = e2;
        }
      }
      node = cur;
      sel_action = bx;
      if (lane == (levels & 31)) {
        if (spill != nullptr && levels - TZ_PATH_CAP < P.spill_cap) spill[levels - TZ_PATH_CAP] = make_int2(ring_n, ring_a);
        ring_n = cur;
        ring_a = bx;
      }
      ++levels;
      if (by < 0) {
        stop_child = by == -1 ? -1 : -(by + 2);
        break;
      }
      if (levels > tv.N) {  // never spin on a corrupted tree
        stop_child = by;
        break;
      }
      cur = by;
    }
  }
  TZ_STAMP(5);
  // ---- embeddings: gather the next parent's rows (mcts.py:161-164); register-path leaves first, all loads in flight
  //      together; a node written by this very launch is read back from registers ----------------------------------
  uint4 gat[SIM_LEAVES_INLINE];
#pragma unroll
  for (int k = 0; k < SIM_LEAVES_INLINE; ++k) {
    gat[k] = pre[k];
    if (((P.fast_mask >> k) & 1) && lane * 16 < (int)P.leaf[k].rb && node != fresh_node)
      gat[k] = reinterpret_cast<const uint4*>(P.leaf[k].table + ((size_t)b * tv.N + (size_t)node) * P.leaf[k].rb)[lane];
  }
  if (lane == 0) {
    P.w_parent[b] = node;
    P.w_action[b] = sel_action;
    if (P.stats) {  // fire-and-forget reductions (RED): a load-add-store here would put two more round trips into the epilogue
      atomicAdd(reinterpret_cast<unsigned long long*>(P.stats) + 4 * (size_t)b + 0, (unsigned long long)levels);
      atomicAdd(reinterpret_cast<unsigned long long*>(P.stats) + 4 * (size_t)b + 1, 1ull);
    }
  }
  if (path) {
    path[lane] = ring_n;
    path[PATH_ACT + lane] = ring_a;
    if (lane == 0) {
      path[PATH_LEN] = levels;
      path[PATH_END] = stop_child;
    }
  }
#pragma unroll
  for (int k = 0; k < SIM_LEAVES_INLINE; ++k) {
    if (((P.fast_mask >> k) & 1) && lane * 16 < (int)P.leaf[k].rb)
      reinterpret_cast<uint4*>(P.leaf[k].parent_out + (size_t)b * P.leaf[k].rb)[lane] = gat[k];
  }
  for (int k = 0; k < P.n_emb; ++k) {
    if (k < SIM_LEAVES_INLINE && ((P.fast_mask >> k) & 1)) continue;
    move_leaf(k < SIM_LEAVES_INLINE ? P.leaf[k] : X.leaf[k - SIM_LEAVES_INLINE], b, tv.N, true, node, fresh_node, lane);
  }
  TZ_STAMP(6);
  TZ_TL_MAX(P.pad0, 2);
  if (mode & MODE_TIMELINE) tl_max(P.tl_row, 2, lane);
#ifdef TZ_PROFILE
  if (b == 0 && lane == 0) g_prof[7] = levels;
  if (b < 4096 && lane == 0) {
    g_prof_warp[4 * b + 0] = prof_t0;
    g_prof_warp[4 * b + 1] = prof_gtime();
    g_prof_warp[4 * b + 2] = L;
    g_prof_warp[4 * b + 3] = levels;
  }
#endif
}

template <typename K>
int launch_sim_k(K kernel, const SimLaunch& L, cudaStream_t s) {
  if (L.smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SIM_SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
  }
  if (use_pdl(L)) {  // programmatic dependent launch: see TzSearchCfg.programmatic
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid_for(L.P.B));
    lc.blockDim = dim3(SIM_THREADS);
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
  kernel<<<grid_for(L.P.B), SIM_THREADS, L.smem, s>>>(L.P, L.X);
  return launch_status();
}

template <int NC, bool WEIGHTED, int SEL, int FM>
int launch_sim_p(const SimLaunch& L, cudaStream_t s) {
  if (L.P.cfg.q_transform == TZ_QT_IDENTITY) {  // (the other q_transforms are instantiated for ordinary launches only)
    SimLaunch L2 = L;
    L2.P.cfg.programmatic = 0;
    return launch_sim_k(k_sim<NC, WEIGHTED, SEL | SELQ_IDENTITY, FM, false>, L2, s);
  }
  if (use_pdl(L)) return launch_sim_k(k_sim<NC, WEIGHTED, SEL, FM, true>, L, s);
  return launch_sim_k(k_sim<NC, WEIGHTED, SEL, FM, false>, L, s);
}

template <int NC, int FM>
int launch_sim_g(const SimLaunch& L, cudaStream_t s) {
  if (L.P.cfg.selector == TZ_SEL_MUZERO_PUCT) return launch_sim_p<NC, false, TZ_SEL_MUZERO_PUCT, FM>(L, s);
  return launch_sim_p<NC, false, TZ_SEL_PUCT, FM>(L, s);
}

template <int NC>
int launch_sim_nc(const SimLaunch& L, cudaStream_t s) {
  const bool mz = L.P.cfg.selector == TZ_SEL_MUZERO_PUCT;
  if (L.P.cfg.weighted) {
    if (mz) return launch_sim_p<NC, true, TZ_SEL_MUZERO_PUCT, 0>(L, s);
    return launch_sim_p<NC, true, TZ_SEL_PUCT, 0>(L, s);
  }
  if (NC == 1) {  // narrow trees: one lane per path level
    if (L.P.F <= 4) return launch_sim_g<1, 4>(L, s);
    if (L.P.F <= 8) return launch_sim_g<1, 8>(L, s);
    if (L.P.F <= 16) return launch_sim_g<1, 16>(L, s);
  }
  return launch_sim_g<NC, 0>(L, s);
}

}  // namespace

#endif  // TZ_SIM_CUH_
