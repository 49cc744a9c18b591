// tz_reroot.cu -- subtree persistence (include/tz_abi.h tz_reroot): Tree.get_subtree tree.py:169-269, Tree.reset
// tree.py:272-278 and the caller's reset-vs-step select core/common.py:89-94, as one launch per move.
// k_reroot_bulk (row moves by the bulk-copy engine) is the product path; k_reroot_all (per-thread LDGSTS gathers,
// TZ_REROOT_IMPL=ldgsts) and k_reroot (one table at a time, rows wider than the staging area) are the fallbacks.
#include "tz_device.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// re-rooting: Tree.get_subtree tree.py:169-269, Tree.reset tree.py:272-278, select common.py:89-94
// ---------------------------------------------------------------------------------------------------------
struct RerootSmem {
  int32_t* trans;   // [N]  pointer-jump scratch, then old index -> new index (or -1)
  int32_t* src_of;  // [N]  new index -> old index
  uint8_t* stage;   // [REROOT_STAGE]
};

__device__ __forceinline__ void block_fill(uint8_t* base, size_t lo, size_t hi, uint32_t pattern) {
  // fills bytes [lo, hi) with a repeated byte pattern (0x00 or 0xFF), vectorised in the aligned middle
  uint8_t* p = base + lo;
  const size_t n = hi - lo;
  const uint8_t pb = (uint8_t)pattern;
  size_t head = (16 - ((uintptr_t)p & 15)) & 15;
  if (head > n) head = n;
  for (size_t i = threadIdx.x; i < head; i += blockDim.x) p[i] = pb;
  const size_t nv = (n - head) >> 4;
  uint4* pv = reinterpret_cast<uint4*>(p + head);
  const uint4 v = make_uint4(pattern, pattern, pattern, pattern);
  for (size_t i = threadIdx.x; i < nv; i += blockDim.x) pv[i] = v;
  for (size_t i = head + (nv << 4) + threadIdx.x; i < n; i += blockDim.x) p[i] = pb;
}

// Order-preserving in-place compaction of one per-tree table with `rb`-byte rows: new row s <- old row src_of[s].
// Safe in place because src_of[s] > s for every s and chunks are processed in increasing s: a chunk's reads
// finish (barrier) before its writes, and later chunks only read rows above everything written so far.
// remap: the table holds int32 node indices that must be translated through trans[] (tree.py:247-257).
// remap == 2: the table holds best-table entries {action, next}; only `next` is an index (TzTree.best encoding).
// remap == 3: the table holds child_stats entries {q, n, p, edge}: every fourth word is an index; the tail is filled
//             with the null entry {0, 0, 0, -1} instead of a byte pattern.
__device__ void compact_table(uint8_t* base, int64_t rb, int count, int nfi, const RerootSmem& sm, int remap,
                              uint32_t null_pattern) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  if ((rb == 1 || rb == 2 || rb == 4 || rb == 8 || rb == 16) && (!remap || (remap == 1 && rb == 4) || remap == 2)) {
    // narrow rows: one thread per row, staged in registers (index tables only when a row is a single index)
    for (int s0 = 0; s0 < count; s0 += nthr) {
      const int s = s0 + tid;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (s < count) {
        const uint8_t* src = base + (size_t)sm.src_of[s] * rb;
        if (rb == 4) {
          int32_t x = *reinterpret_cast<const int32_t*>(src);
          if (remap) x = x < 0 ? -1 : sm.trans[x];
          v.x = (uint32_t)x;
        } else if (rb == 1) v.x = *src;
        else if (rb == 2) v.x = *reinterpret_cast<const uint16_t*>(src);
        else if (rb == 8) {
          const uint2 t2 = *reinterpret_cast<const uint2*>(src);
          v.x = t2.x;
          v.y = t2.y;
          if (remap == 2) {
            const int nx = (int)t2.y;
            if (nx >= 0) v.y = (uint32_t)sm.trans[nx];
            else if (nx <= -2) v.y = (uint32_t)(-(sm.trans[-(nx + 2)] + 2));
          }
        }
        else v = *reinterpret_cast<const uint4*>(src);
      }
      __syncthreads();
      if (s < count) {
        uint8_t* dst = base + (size_t)s * rb;
        if (rb == 4) *reinterpret_cast<uint32_t*>(dst) = v.x;
        else if (rb == 1) *dst = (uint8_t)v.x;
        else if (rb == 2) *reinterpret_cast<uint16_t*>(dst) = (uint16_t)v.x;
        else if (rb == 8) *reinterpret_cast<uint2*>(dst) = make_uint2(v.x, v.y);
        else *reinterpret_cast<uint4*>(dst) = v;
      }
    }
  } else {
    const int rows_per_chunk = (int)(REROOT_STAGE / rb);  // >= 1, checked on the host
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int vw = (remap == 1 || remap == 3) ? 4 : ((rb & 15) == 0 ? 16 : ((rb & 3) == 0 ? 4 : 1));
    for (int s0 = 0; s0 < count; s0 += rows_per_chunk) {
      const int rows = min(rows_per_chunk, count - s0);
      for (int s = warp; s < rows; s += nwarps) {  // gather: one warp per row, coalesced within the row
        const uint8_t* src = base + (size_t)sm.src_of[s0 + s] * rb;
        uint8_t* st = sm.stage + (size_t)s * rb;
        if (vw == 16) {
          for (int i = lane; i < (int)(rb >> 4); i += 32) reinterpret_cast<uint4*>(st)[i] = reinterpret_cast<const uint4*>(src)[i];
        } else if (vw == 4) {
          for (int i = lane; i < (int)(rb >> 2); i += 32) {
            int32_t x = reinterpret_cast<const int32_t*>(src)[i];
            if (remap == 1 || (remap == 3 && (i & 3) == 3)) x = x < 0 ? -1 : sm.trans[x];
            reinterpret_cast<int32_t*>(st)[i] = x;
          }
        } else {
          for (int i = lane; i < (int)rb; i += 32) st[i] = src[i];
        }
      }
      __syncthreads();
      {  // scatter: the chunk's destination rows are contiguous -> one flat coalesced copy
        uint8_t* dst = base + (size_t)s0 * rb;
        const size_t nbytes = (size_t)rows * rb;
        if (((uintptr_t)dst & 15) == 0 && (nbytes & 15) == 0) {
          for (size_t i = tid; i < (nbytes >> 4); i += nthr) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(sm.stage)[i];
        } else if (((uintptr_t)dst & 3) == 0 && (nbytes & 3) == 0) {
          for (size_t i = tid; i < (nbytes >> 2); i += nthr) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(sm.stage)[i];
        } else {
          for (size_t i = tid; i < nbytes; i += nthr) dst[i] = sm.stage[i];
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (remap == 3) {
    int4* e = reinterpret_cast<int4*>(base);
    for (size_t i = (size_t)count * (rb >> 4) + tid; i < (size_t)nfi * (rb >> 4); i += nthr) e[i] = make_int4(0, 0, 0, -1);
  } else {
    block_fill(base, (size_t)count * rb, (size_t)nfi * rb, null_pattern);  // tree.py:236-238,247-249
  }
}

__global__ void __launch_bounds__(REROOT_THREADS) k_reroot(const TzTree t, const int32_t* __restrict__ action,
                                                         const uint8_t* __restrict__ reset_flag, const int persist_tree) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ int wsum[REROOT_THREADS / 32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const TV tv = make_view(t, b);
  const int N = tv.N, F = tv.F;
  RerootSmem sm;
  sm.stage = smem_raw;
  sm.trans = reinterpret_cast<int32_t*>(smem_raw + REROOT_STAGE);
  sm.src_of = sm.trans + N;

  const int flag = reset_flag ? (int)reset_flag[b] : 0;
  if (flag == 2) return;  // leave this tree untouched (core/common.py:91 `lambda s: s`)
  const int nfi = *tv.nfi;
  const bool do_reset = !persist_tree || flag != 0;
  // edge_map[ROOT, action]; -1 -> nothing retained (tree.py:201-203).  Out-of-range actions clamp like an XLA gather.
  const int c = do_reset ? -1 : tv.edge[min(max(action[b], 0), F - 1)];
  int count = 0;
  if (c >= 0) {
    // (1) every node finds out whether new root c is its ancestor: pointer jumping, roots {0, c} absorb.
    //     Jacobi rounds between the two index arrays (src_of is free until the scan): race-free under racecheck.
    for (int i = tid; i < nfi; i += nthr) sm.trans[i] = (i == 0 || i == c) ? i : tv.parents[i];
    __syncthreads();
    int32_t* cur = sm.trans;
    int32_t* nxt = sm.src_of;
    for (int round = 0; round < 34; ++round) {  // ancestor distance doubles per round: <= log2(N) + 1 rounds
      int pending = 0;
      for (int i = tid; i < nfi; i += nthr) {
        const int a = cur[i];
        int g = a;
        if (a != 0 && a != c) {
          g = cur[a];
          pending |= (g != 0 && g != c);
        }
        nxt[i] = g;
      }
      int32_t* const t2 = cur;
      cur = nxt;
      nxt = t2;
      if (!__syncthreads_or(pending)) break;
    }
    if (cur != sm.trans) {
      for (int i = tid; i < nfi; i += nthr) sm.trans[i] = cur[i];
      __syncthreads();
    }
    // (2) stable compaction indices: block prefix scan over the retain flags (tree.py:204-213)
    int base = 0;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i0 = 0; i0 < nfi; i0 += nthr) {
      const int i = i0 + tid;
      const bool keep = i < nfi && i > 0 && sm.trans[i] == c;
      const unsigned bal = __ballot_sync(FULL, keep);
      if (lane == 0) wsum[warp] = __popc(bal);
      __syncthreads();
      int off = 0, total = 0;
#pragma unroll
      for (int k = 0; k < REROOT_THREADS / 32; ++k) {
        const int s = wsum[k];
        off += k < warp ? s : 0;
        total += s;
      }
      if (i < nfi) {
        const int slot = base + off + __popc(bal & ((1u << lane) - 1u));
        sm.trans[i] = keep ? slot : -1;
        if (keep) sm.src_of[slot] = i;
      }
      base += total;
      __syncthreads();
    }
    count = base;
  }
  if (tid == 0 && t.stats) {
    atomicAdd(reinterpret_cast<unsigned long long*>(t.stats) + 4 * (size_t)b + 2, (unsigned long long)nfi);
    atomicAdd(reinterpret_cast<unsigned long long*>(t.stats) + 4 * (size_t)b + 3, (unsigned long long)count);
  }
  // (3) move rows, translate indices, null the tail (tree.py:234-268)
  compact_table(reinterpret_cast<uint8_t*>(tv.parents), 4, count, nfi, sm, 1, 0xffffffffu);
  compact_table(reinterpret_cast<uint8_t*>(tv.edge), 4 * (int64_t)F, count, nfi, sm, 1, 0xffffffffu);
  compact_table(reinterpret_cast<uint8_t*>(tv.n), 4, count, nfi, sm, 0, 0u);
  compact_table(reinterpret_cast<uint8_t*>(tv.q), 4, count, nfi, sm, 0, 0u);
  if (tv.r) compact_table(reinterpret_cast<uint8_t*>(tv.r), 4, count, nfi, sm, 0, 0u);
  compact_table(reinterpret_cast<uint8_t*>(tv.term), 1, count, nfi, sm, 0, 0u);
  compact_table(reinterpret_cast<uint8_t*>(tv.p), 4 * (int64_t)F, count, nfi, sm, 0, 0u);
  compact_table(reinterpret_cast<uint8_t*>(tv.cs), 16 * (int64_t)F, count, nfi, sm, 3, 0u);  // edge word translated
  compact_table(reinterpret_cast<uint8_t*>(tv.best), 8, count, nfi, sm, 2, 0xffffffffu);    // entries move with their nodes
  for (int k = 0; k < t.n_emb; ++k) {
    const int64_t rb = t.emb_row_bytes[k];
    compact_table(reinterpret_cast<uint8_t*>(t.emb[k]) + (size_t)b * N * rb, rb, count, nfi, sm, 0, 0u);
  }
  if (tid == 0) *tv.nfi = count;
}

// ---------------------------------------------------------------------------------------------------------
// k_reroot_all: the same re-rooting with ALL tables of a tree moved together.  k_reroot above compacts one table at a
// time (two barriers and one memory round trip per table and chunk: ~30 dependent round trips per tree, which is what
// bounded it).  Here a chunk of destination rows is gathered for every table at once with fire-and-forget
// global->shared copies (LDGSTS), one wait + barrier, then scattered (index words translated on the way out): 2-3
// round trips per tree for configs[1], and the staged bytes in flight per SM are what the HBM/L2 pipe needs.
// In place for the same reason as above: src_of[s] > s, chunks in increasing s.
// ---------------------------------------------------------------------------------------------------------
constexpr int REROOT2_THREADS = 128;
constexpr int REROOT_MAX_TABS = 9 + TZ_MAX_EMB;
struct RerootTab {
  uint8_t* base;   // batch base; the tree's rows start at base + b * N * rb
  int64_t rb;      // row bytes
  int32_t kind;    // 0: opaque bytes, 1: every 32-bit word is a node index, 2: best-table entries, 3: child_stats entries,
                   // 4 / 5: p / edge_map rows, NOT gathered: rebuilt on the way out from the staged child_stats rows (their .z / .w)
  uint32_t null_pattern;  // byte pattern (replicated) of a null row for kinds 0-2
  uint32_t unit;   // bytes per gather copy: 16 / 8 / 4 by row size and base alignment, 1 = ordinary byte loads
  uint32_t units;  // rb / unit
  uint32_t magic;  // floor(2^32 / units) + 1: i / units == __umulhi(i, magic) for the index range of a chunk
  uint32_t pad;
  uint32_t st_off; // k_reroot_bulk: the table's offset in the staging area (or, preloaded tables, in the preload area)
  uint32_t code;   // k_reroot_bulk: TD_* (how the table's rows travel)
};
struct RerootP {
  int32_t B, N, F, ntab;
  int32_t* nfi;
  const int32_t* parents;
  const int32_t* edge;
  uint64_t* stats;
  int32_t stage_bytes;  // shared-memory staging area
  int32_t rpc;          // destination rows per chunk (>= 1)
  int32_t bulk_row_bytes;  // k_reroot_bulk: bytes per destination row that arrive by bulk copies (sum of rb over unit == 0 tables)
  int32_t p_tab, e_tab;    // k_reroot_bulk: indices of the p / edge_map tables (kinds 4 / 5)
  int32_t pre_first;       // k_reroot_bulk: tables [pre_first, ntab) are preloaded whole into shared memory
  RerootTab tab[REROOT_MAX_TABS];
};

__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__global__ void __launch_bounds__(REROOT2_THREADS) k_reroot_all(const __grid_constant__ RerootP P, const int32_t* __restrict__ action,
                                                              const uint8_t* __restrict__ reset_flag, const int persist_tree) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ int wsum[REROOT2_THREADS / 32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  constexpr int nthr = REROOT2_THREADS;
  const int N = P.N, F = P.F;
  uint8_t* const stage = smem_raw;
  int32_t* const trans = reinterpret_cast<int32_t*>(smem_raw + P.stage_bytes);  // old index -> new index (or -1)
  int32_t* const src_of = trans + N;                                            // new index -> old index

  const int flag = reset_flag ? (int)reset_flag[b] : 0;
  if (flag == 2) return;  // leave this tree untouched (core/common.py:91 `lambda s: s`)
  const int nfi = P.nfi[b];
  const bool do_reset = !persist_tree || flag != 0;
  const int32_t* const parents = P.parents + (size_t)b * N;
  // edge_map[ROOT, action]; -1 -> nothing retained (tree.py:201-203).  Out-of-range actions clamp like an XLA gather.
  const int c = do_reset ? -1 : P.edge[(size_t)b * N * F + min(max(action[b], 0), F - 1)];
  int count = 0;
  if (c >= 0) {
    // (1) ancestor test by pointer jumping (see k_reroot)
    //     Jacobi rounds between the two index arrays (src_of is free until the scan): no thread reads what another writes
    for (int i = tid; i < nfi; i += nthr) trans[i] = (i == 0 || i == c) ? i : parents[i];
    __syncthreads();
    int32_t* cur = trans;
    int32_t* nxt = src_of;
    for (int round = 0; round < 34; ++round) {  // ancestor distance doubles per round: <= log2(N) + 1 rounds
      int pending = 0;
      for (int i = tid; i < nfi; i += nthr) {
        const int a = cur[i];
        int g = a;
        if (a != 0 && a != c) {
          g = cur[a];
          pending |= (g != 0 && g != c);
        }
        nxt[i] = g;
      }
      int32_t* const t2 = cur;
      cur = nxt;
      nxt = t2;
      if (!__syncthreads_or(pending)) break;
    }
    if (cur != trans) {  // (uniform) the labels ended up in src_of: bring them home before the scan reuses it
      for (int i = tid; i < nfi; i += nthr) trans[i] = cur[i];
      __syncthreads();
    }
    // (2) stable compaction indices: block prefix scan over the retain flags (tree.py:204-213)
    int base = 0;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i0 = 0; i0 < nfi; i0 += nthr) {
      const int i = i0 + tid;
      const bool keep = i < nfi && i > 0 && trans[i] == c;
      const unsigned bal = __ballot_sync(FULL, keep);
      if (lane == 0) wsum[warp] = __popc(bal);
      __syncthreads();
      int off = 0, total = 0;
#pragma unroll
      for (int k = 0; k < REROOT2_THREADS / 32; ++k) {
        const int sct = wsum[k];
        off += k < warp ? sct : 0;
        total += sct;
      }
      if (i < nfi) {
        const int slot = base + off + __popc(bal & ((1u << lane) - 1u));
        trans[i] = keep ? slot : -1;
        if (keep) src_of[slot] = i;
      }
      base += total;
      __syncthreads();
    }
    count = base;
  }
  if (tid == 0 && P.stats) {
    atomicAdd(reinterpret_cast<unsigned long long*>(P.stats) + 4 * (size_t)b + 2, (unsigned long long)nfi);
    atomicAdd(reinterpret_cast<unsigned long long*>(P.stats) + 4 * (size_t)b + 3, (unsigned long long)count);
  }
  // (3) move rows, translate indices (tree.py:234-268): all tables per chunk of destination rows
  const int rpc = P.rpc;
  for (int s0 = 0; s0 < count; s0 += rpc) {
    const int rows = min(rpc, count - s0);
    size_t off = 0;
    for (int t = 0; t < P.ntab; ++t) {  // gather: fire-and-forget copies, nothing waits here
      if (P.tab[t].kind >= 4) continue;  // p / edge_map: carried by the child_stats rows
      const int64_t rb = P.tab[t].rb;
      const uint8_t* const src = P.tab[t].base + (size_t)b * N * rb;
      uint8_t* const st = stage + off;
      const uint32_t unit = P.tab[t].unit, units = P.tab[t].units, magic = P.tab[t].magic;
      const uint32_t total = (uint32_t)rows * units;
      if (unit == 16) {
        for (uint32_t i = tid; i < total; i += nthr) {
          const uint32_t r = magic ? __umulhi(i, magic) : i / units, u = i - r * units;  // (row, unit) of copy i, normally without a division
          cp_async16(st + (size_t)r * rb + 16 * u, src + (size_t)src_of[s0 + r] * rb + 16 * u);
        }
      } else if (unit == 8) {
        for (uint32_t i = tid; i < total; i += nthr) {
          const uint32_t r = magic ? __umulhi(i, magic) : i / units, u = i - r * units;
          cp_async8(st + (size_t)r * rb + 8 * u, src + (size_t)src_of[s0 + r] * rb + 8 * u);
        }
      } else if (unit == 4) {
        for (uint32_t i = tid; i < total; i += nthr) {
          const uint32_t r = magic ? __umulhi(i, magic) : i / units, u = i - r * units;
          cp_async4(st + (size_t)r * rb + 4 * u, src + (size_t)src_of[s0 + r] * rb + 4 * u);
        }
      } else {  // odd row sizes (bool / byte leaves): ordinary loads; the host orders these tables last
        for (uint32_t i = tid; i < total; i += nthr) {
          const uint32_t r = magic ? __umulhi(i, magic) : i / units, u = i - r * units;
          st[i] = src[(size_t)src_of[s0 + r] * rb + u];
        }
      }
      off += align16((size_t)rpc * rb);
    }
    cp_async_wait_all();
    __syncthreads();
    off = 0;
    for (int t = 0; t < P.ntab; ++t) {  // scatter: the chunk's destination rows are contiguous in every table
      const int64_t rb = P.tab[t].rb;
      uint8_t* const dst = P.tab[t].base + ((size_t)b * N + (size_t)s0) * rb;
      const uint8_t* const st = stage + off;
      const size_t nbytes = (size_t)rows * rb;
      const int kind = P.tab[t].kind;
      if (kind >= 4) {  // p (4) / edge_map (5) rows from the staged child_stats entries {q, n, p, edge} (table 0, offset 0)
        const int4* const cs_st = reinterpret_cast<const int4*>(stage);
        if (kind == 4) {
          for (size_t i = tid; i < (nbytes >> 2); i += nthr) reinterpret_cast<int32_t*>(dst)[i] = cs_st[i].z;
        } else {
          for (size_t i = tid; i < (nbytes >> 2); i += nthr) {
            const int32_t x = cs_st[i].w;
            reinterpret_cast<int32_t*>(dst)[i] = x < 0 ? -1 : trans[x];  // tree.py:247-257
          }
        }
        continue;  // (nothing staged for this table)
      }
      if (kind == 1) {  // every word is a node index (parents, edge_map): tree.py:247-257
        for (size_t i = tid; i < (nbytes >> 2); i += nthr) {
          const int32_t x = reinterpret_cast<const int32_t*>(st)[i];
          reinterpret_cast<int32_t*>(dst)[i] = x < 0 ? -1 : trans[x];
        }
      } else if (kind == 2) {  // best-table entries {action, next}: only `next` is an index (TzTree.best encoding)
        for (size_t i = tid; i < (nbytes >> 3); i += nthr) {
          int2 e = reinterpret_cast<const int2*>(st)[i];
          if (e.y >= 0) e.y = trans[e.y];
          else if (e.y <= -2) e.y = -(trans[-(e.y + 2)] + 2);
          reinterpret_cast<int2*>(dst)[i] = e;
        }
      } else if (kind == 3) {  // child_stats entries {q, n, p, edge}
        for (size_t i = tid; i < (nbytes >> 4); i += nthr) {
          int4 e = reinterpret_cast<const int4*>(st)[i];
          e.w = e.w < 0 ? -1 : trans[e.w];
          reinterpret_cast<int4*>(dst)[i] = e;
        }
      } else if ((((uintptr_t)dst | nbytes) & 15) == 0) {
        for (size_t i = tid; i < (nbytes >> 4); i += nthr) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(st)[i];
      } else if ((((uintptr_t)dst | nbytes) & 3) == 0) {
        for (size_t i = tid; i < (nbytes >> 2); i += nthr) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(st)[i];
      } else {
        for (size_t i = tid; i < nbytes; i += nthr) dst[i] = st[i];
      }
      off += align16((size_t)rpc * rb);
    }
    __syncthreads();  // the staging area is reused by the next chunk
  }
  // (4) null the tail rows [count, nfi) (tree.py:236-238,247-249): after every source row has been read
  for (int t = 0; t < P.ntab; ++t) {
    const int64_t rb = P.tab[t].rb;
    uint8_t* const base = P.tab[t].base + (size_t)b * N * rb;
    if (P.tab[t].kind == 3) {
      int4* e = reinterpret_cast<int4*>(base);
      for (size_t i = (size_t)count * (rb >> 4) + tid; i < (size_t)nfi * (rb >> 4); i += nthr) e[i] = make_int4(0, 0, 0, -1);
    } else {
      block_fill(base, (size_t)count * rb, (size_t)nfi * rb, P.tab[t].null_pattern);
    }
  }
  if (tid == 0) P.nfi[b] = count;
}

// ---------------------------------------------------------------------------------------------------------
// k_reroot_bulk: k_reroot_all with the row moves done by the bulk-copy engine.  The gather of k_reroot_all issues one
// LDGSTS per 16 / 8 / 4 bytes with ~10 instructions of index arithmetic each (7 copies for a 112-byte child_stats row, 17
// for a 272-byte embedding row): ncu showed it instruction-issue bound at 22 % of DRAM throughput.  Here
//  * every 16-byte-aligned row (child_stats rows, embedding leaves with rb % 16 == 0) is ONE cp.async.bulk.shared.global
//    issued by one thread per row, completing on an mbarrier armed with the chunk's expected byte count;
//  * the contiguous destination chunk of every opaque 16-byte-aligned table goes back as ONE cp.async.bulk.global.shared;
//  * child_stats rows are read from the staging area once and produce their three outputs in one pass: the translated
//    child_stats entry, the p word and the translated edge_map word (tree.py:247-257);
//  * narrow tables (best, parents, n, q, r: 4-8 bytes per row) keep one LDGSTS per row, byte-sized rows ordinary loads.
// Same chunking / in-place argument as k_reroot_all (src_of[s] > s, chunks in increasing s).
// ---------------------------------------------------------------------------------------------------------
// (mbarrier / cp.async.bulk helpers: tz_device.cuh)
// RerootTab.unit == 0 marks a table moved by bulk copies (rows and tree blocks 16-byte aligned)
#ifdef TZ_PROFILE
#define TZ_RSTAMP(i) do { if (threadIdx.x == 0 && blockIdx.x < 4096) g_prof_gt[16 * blockIdx.x + (i)] = prof_gtime(); } while (0)
#else
#define TZ_RSTAMP(i) do { } while (0)
#endif
// how a table's rows travel in k_reroot_bulk (TabDesc.code)
enum : uint32_t { TD_SKIP = 0, TD_BULK, TD_N4, TD_N8, TD_N16, TD_BYTE, TD_UNITS, TD_PRE };
struct __align__(16) TabDesc {
  uint8_t* base;     // this tree's block of the table
  uint32_t st_off;   // the table's offset in the staging area
  uint32_t rb : 24;  // row bytes (< 2^20: a row fits the staging area)
  uint32_t code : 4; // TD_*
  uint32_t kind : 4; // RerootTab.kind
};

__device__ __forceinline__ int ld_issue(const int32_t* p) {
  int v;
  asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <typename K>
inline int64_t static_shared_bytes(K kernel) {
  cudaFuncAttributes a;
  return cudaFuncGetAttributes(&a, kernel) == cudaSuccess ? (int64_t)a.sharedSizeBytes : 4096;
}

template <int NTHR>
__global__ void __launch_bounds__(NTHR) k_reroot_bulk(const __grid_constant__ RerootP P, const int32_t* __restrict__ action,
                                                               const uint8_t* __restrict__ reset_flag, const int persist_tree) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ int wsum[NTHR / 32];
  __shared__ __align__(8) uint64_t bar;
  __shared__ TabDesc s_td[REROOT_MAX_TABS];
  __shared__ int s_edge[NTHR];  // the root's edge_map row (F <= NTHR)
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  constexpr int nthr = NTHR;
  const int N = P.N, F = P.F;
  uint8_t* const stage = smem_raw;
  int32_t* const trans = reinterpret_cast<int32_t*>(smem_raw + P.stage_bytes);  // old index -> new index (or -1)
  int32_t* const src_of = trans + N;                                            // new index -> old index

  TZ_RSTAMP(0);
  // Everything the ancestor test needs travels in ONE round trip: the reset flag, next_free_idx, the action, the root's
  // edge_map row and the first PF strides of parents[] are all requested before the first of them is looked at (the flag
  // test, then the action -> edge_map -> parents chain were three dependent round trips on every tree's critical path; rows
  // beyond next_free_idx are initialised, reading them is harmless).
  const int32_t* const parents = P.parents + (size_t)b * N;
  constexpr int PF = 4;
  const bool edge_row = F <= NTHR;
  // (ld_issue: a volatile load instruction -- the compiler would otherwise sink these loads below the early return)
  const int flag = reset_flag ? (int)reset_flag[b] : 0;
  const int nfi = ld_issue(P.nfi + b);
  const int act_raw = (persist_tree && action) ? ld_issue(action + b) : 0;
  int par_pre[PF];
#pragma unroll
  for (int k = 0; k < PF; ++k) par_pre[k] = (persist_tree && tid + k * nthr < N) ? ld_issue(parents + tid + k * nthr) : TZ_NULL_INDEX;
  int my_edge = -1;
  if (edge_row && persist_tree && tid < F) my_edge = ld_issue(P.edge + (size_t)b * N * F + tid);
  if (flag == 2) return;  // leave this tree untouched (core/common.py:91 `lambda s: s`)
  const bool do_reset = !persist_tree || flag != 0;
  const int act = min(max(act_raw, 0), F - 1);  // out-of-range actions clamp like an XLA gather
  if (edge_row) s_edge[tid] = my_edge;
  if (tid == 0) mbar_init(&bar, 1);
  // edge_map[ROOT, action]; -1 -> nothing retained (tree.py:201-203)
  int c = -1;
  if (!do_reset) {  // (uniform over the CTA)
    if (edge_row) {
      __syncthreads();
      c = s_edge[act];
    } else {
      c = P.edge[(size_t)b * N * F + act];
    }
  }
  int count = 0;
  if (tid < P.ntab) {  // this tree's table descriptors (once per kernel; visible after the barriers below)
    const RerootTab& tb = P.tab[tid];
    TabDesc d;
    d.base = tb.base + (size_t)b * N * tb.rb;
    d.st_off = tb.st_off;
    d.rb = (uint32_t)tb.rb;
    d.kind = (uint32_t)tb.kind;
    d.code = tb.code;
    s_td[tid] = d;
  }
  // Narrow tables (4 / 8 / 1 bytes per row: best, parents, n, q, r, terminated) are not moved chunk by chunk: gathering them
  // costs one memory transaction per row and table (the chunk gather was transaction-bound), while the WHOLE table of this
  // tree is a few coalesced lines.  They are copied to shared memory here (fire-and-forget; they land during the ancestor test)
  // and compacted from there once the translation is known.
  uint8_t* const pre = reinterpret_cast<uint8_t*>(src_of + N);
  for (int t = P.pre_first; t < P.ntab; ++t) {  // (the host orders the preloaded tables last)
    const uint8_t* const src = P.tab[t].base + (size_t)b * N * P.tab[t].rb;
    uint8_t* const dstp = pre + P.tab[t].st_off;
    const int words = (int)(((size_t)nfi * P.tab[t].rb + 3) >> 2);
    for (int i = tid; i < words; i += nthr) cp_async4(dstp + 4 * (size_t)i, src + 4 * (size_t)i);
  }
  TZ_RSTAMP(1);
  if (c >= 0) {
    // (1) ancestor test by pointer jumping (see k_reroot): Jacobi rounds between the two index arrays
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int i = tid + k * nthr;
      if (i < nfi) trans[i] = (i == 0 || i == c) ? i : par_pre[k];
    }
    for (int i = tid + PF * nthr; i < nfi; i += nthr) trans[i] = (i == 0 || i == c) ? i : parents[i];
    __syncthreads();
    int32_t* cur = trans;
    int32_t* nxt = src_of;
    for (int round = 0; round < 34; ++round) {  // ancestor distance doubles per round: <= log2(N) + 1 rounds
      int pending = 0;
      for (int i = tid; i < nfi; i += nthr) {
        const int a = cur[i];
        int g = a;
        if (a != 0 && a != c) {
          g = cur[a];
          pending |= (g != 0 && g != c);
        }
        nxt[i] = g;
      }
      int32_t* const t2 = cur;
      cur = nxt;
      nxt = t2;
      if (!__syncthreads_or(pending)) break;
    }
    if (cur != trans) {  // (uniform) the labels ended up in src_of: bring them home before the scan reuses it
      for (int i = tid; i < nfi; i += nthr) trans[i] = cur[i];
      __syncthreads();
    }
    TZ_RSTAMP(2);
    // (2) stable compaction indices: block prefix scan over the retain flags (tree.py:204-213)
    int base = 0;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i0 = 0; i0 < nfi; i0 += nthr) {
      const int i = i0 + tid;
      const bool keep = i < nfi && i > 0 && trans[i] == c;
      const unsigned bal = __ballot_sync(FULL, keep);
      if (lane == 0) wsum[warp] = __popc(bal);
      __syncthreads();
      int off = 0, total = 0;
#pragma unroll
      for (int k = 0; k < NTHR / 32; ++k) {
        const int sct = wsum[k];
        off += k < warp ? sct : 0;
        total += sct;
      }
      if (i < nfi) {
        const int slot = base + off + __popc(bal & ((1u << lane) - 1u));
        trans[i] = keep ? slot : -1;
        if (keep) src_of[slot] = i;
      }
      base += total;
      __syncthreads();
    }
    count = base;
  } else {
    __syncthreads();  // the mbarrier initialisation is visible to every thread on both paths
  }
  if (tid == 0 && P.stats) {
    atomicAdd(reinterpret_cast<unsigned long long*>(P.stats) + 4 * (size_t)b + 2, (unsigned long long)nfi);
    atomicAdd(reinterpret_cast<unsigned long long*>(P.stats) + 4 * (size_t)b + 3, (unsigned long long)count);
  }
  // (2b) the preloaded narrow tables: new row s <- old row src_of[s], indices translated; rows [count, nfi) nulled
  cp_async_wait_all();
  __syncthreads();
  for (int g = P.pre_first; g < P.ntab; ++g) {
    const TabDesc d = s_td[g];
    const uint8_t* const sp = pre + d.st_off;
    if (d.rb == 4) {
      int32_t* const out = reinterpret_cast<int32_t*>(d.base);
      const int32_t nullv = (int32_t)P.tab[g].null_pattern;
      for (int sidx = tid; sidx < nfi; sidx += nthr) {
        int32_t x = nullv;
        if (sidx < count) {
          x = reinterpret_cast<const int32_t*>(sp)[src_of[sidx]];
          if (d.kind == 1) x = x < 0 ? -1 : trans[x];  // a node index (parents): tree.py:247-257
        }
        out[sidx] = x;
      }
    } else if (d.rb == 8) {
      int2* const out = reinterpret_cast<int2*>(d.base);
      const int32_t nullv = (int32_t)P.tab[g].null_pattern;
      for (int sidx = tid; sidx < nfi; sidx += nthr) {
        int2 e = make_int2(nullv, nullv);
        if (sidx < count) {
          e = reinterpret_cast<const int2*>(sp)[src_of[sidx]];
          if (d.kind == 2) {  // best-table entries {action, next}: only `next` is an index (TzTree.best encoding)
            if (e.y >= 0) e.y = trans[e.y];
            else if (e.y <= -2) e.y = -(trans[-(e.y + 2)] + 2);
          }
        }
        out[sidx] = e;
      }
    } else {  // one byte per row
      for (int sidx = tid; sidx < nfi; sidx += nthr) d.base[sidx] = sidx < count ? sp[src_of[sidx]] : (uint8_t)P.tab[g].null_pattern;
    }
  }
  TZ_RSTAMP(3);
  // (3) move rows, translate indices (tree.py:234-268): all tables per chunk of destination rows.  The per-table facts a chunk
  // needs (this tree's block, the table's offset in the staging area, how its rows travel) are worked out ONCE into shared
  // memory (s_td, below); per chunk every WARP then takes whole tables (lanes over the rows), so a thread interprets 2-3
  // descriptors per chunk instead of walking the kernel-parameter table list (~5 us per chunk when every thread did).
  const int rpc = P.rpc;
  constexpr int nwarps = NTHR / 32;
  const int warp = tid >> 5, lane = tid & 31;
  const TabDesc* const cs_td = &s_td[0];  // child_stats is table 0 (the host orders it first)
  unsigned parity = 0;
#ifdef TZ_PROFILE
  long long acc_issue = 0, acc_wait = 0, acc_scatter = 0, t_a = prof_gtime();
#endif
  for (int s0 = 0; s0 < count; s0 += rpc) {
    const int rows = min(rpc, count - s0);
    // ---- gather --------------------------------------------------------------------------------------------------
    if (tid == 0) mbar_expect_tx(&bar, (unsigned)rows * (unsigned)P.bulk_row_bytes);  // arms this chunk's phase
    for (int g = warp; g < P.pre_first; g += nwarps) {
      const TabDesc d = s_td[g];
      uint8_t* const st = stage + d.st_off;
      const uint32_t rb = d.rb;
      switch (d.code) {
        case TD_BULK:  // one bulk copy per RUN of adjacent source rows (their destinations are adjacent by construction): the
                       // bulk-copy unit takes ~3 ns per copy whatever its size, and nodes allocated by consecutive simulations
                       // into the same subtree are neighbours
          for (int r0 = 0; r0 < rows; r0 += 32) {
            const int r = r0 + lane;
            const bool in = r < rows;
            const int srow = in ? src_of[s0 + r] : -2;
            const int prev = __shfl_up_sync(FULL, srow, 1);
            const bool start = in && (lane == 0 || prev != srow - 1);
            const unsigned starts = __ballot_sync(FULL, start);
            const unsigned inmask = __ballot_sync(FULL, in);
            if (start) {
              const unsigned above = starts & ~((2u << lane) - 1u);                    // runs that start above this lane
              const int end = above ? __ffs(above) - 1 : 32 - __clz(inmask);           // first lane past this run
              bulk_g2s(st + (size_t)r * rb, d.base + (size_t)srow * rb, (unsigned)(end - lane) * rb, &bar);
            }
          }
          break;
        case TD_N4:
          for (int r = lane; r < rows; r += 32) cp_async4(st + 4 * (size_t)r, d.base + 4 * (size_t)src_of[s0 + r]);
          break;
        case TD_N8:
          for (int r = lane; r < rows; r += 32) cp_async8(st + 8 * (size_t)r, d.base + 8 * (size_t)src_of[s0 + r]);
          break;
        case TD_N16:
          for (int r = lane; r < rows; r += 32) cp_async16(st + 16 * (size_t)r, d.base + 16 * (size_t)src_of[s0 + r]);
          break;
        case TD_BYTE:  // one-byte rows: the aligned 32-bit word that holds the byte (no blocking load in the gather)
          for (int r = lane; r < rows; r += 32) {
            const uint8_t* a = d.base + src_of[s0 + r];
            cp_async4(st + 4 * (size_t)r, a - ((uintptr_t)a & 3));
          }
          break;
        case TD_UNITS: {  // wider rows that are not 16-byte aligned: `unit`-byte copies (the table list has unit / units)
          const uint32_t unit = P.tab[g].unit, units = P.tab[g].units, magic = P.tab[g].magic;
          const uint32_t total = (uint32_t)rows * units;
          for (uint32_t i = lane; i < total; i += 32) {
            const uint32_t r = magic ? __umulhi(i, magic) : i / units, u = i - r * units;
            uint8_t* const sp = st + (size_t)r * rb + unit * u;
            const uint8_t* const gp = d.base + (size_t)src_of[s0 + r] * rb + unit * u;
            if (unit == 16) cp_async16(sp, gp);
            else if (unit == 8) cp_async8(sp, gp);
            else if (unit == 4) cp_async4(sp, gp);
            else *sp = *gp;  // odd row sizes (byte leaves): ordinary loads
          }
          break;
        }
        default:  // TD_SKIP: p / edge_map, carried by the child_stats rows
          break;
      }
    }
    if (s0 == 0) TZ_RSTAMP(4);
#ifdef TZ_PROFILE
    { const long long t_b = prof_gtime(); acc_issue += t_b - t_a; t_a = t_b; }
#endif
    cp_async_wait_all();
    mbar_wait(&bar, parity);
    parity ^= 1u;
    __syncthreads();
    if (s0 == 0) TZ_RSTAMP(5);
#ifdef TZ_PROFILE
    { const long long t_b = prof_gtime(); acc_wait += t_b - t_a; t_a = t_b; }
#endif
    // ---- scatter: the chunk's destination rows are contiguous in every table -----------------------------------------
    {  // child_stats entries {q, n, p, edge}, all threads: one read of the staged entry -> the translated entry, the p word and
       // the translated edge_map word (tree.py:247-257)
      const int4* const st = reinterpret_cast<const int4*>(stage + cs_td->st_off);
      int4* const dst = reinterpret_cast<int4*>(cs_td->base) + (size_t)s0 * F;
      int32_t* const p_dst = reinterpret_cast<int32_t*>(s_td[P.p_tab].base) + (size_t)s0 * F;
      int32_t* const e_dst = reinterpret_cast<int32_t*>(s_td[P.e_tab].base) + (size_t)s0 * F;
      const int n_ent = rows * F;
      for (int i = tid; i < n_ent; i += nthr) {
        int4 e = st[i];
        e.w = e.w < 0 ? -1 : trans[e.w];
        dst[i] = e;
        p_dst[i] = e.z;
        e_dst[i] = e.w;
      }
    }
    bool stored_bulk = false;
    for (int g = 1 + warp; g < P.pre_first; g += nwarps) {
      const TabDesc d = s_td[g];
      const uint8_t* const st = stage + d.st_off;
      uint8_t* const dst = d.base + (size_t)s0 * d.rb;
      switch (d.code) {
        case TD_BULK:  // opaque rows that came in by bulk copies go out as ONE bulk copy
          if (lane == 0) {
            fence_async_smem();
            bulk_s2g(dst, st, (unsigned)rows * d.rb);
            stored_bulk = true;
          }
          break;
        case TD_N4:
          if (d.kind == 1) {  // a node index (parents): tree.py:247-257
            for (int r = lane; r < rows; r += 32) {
              const int32_t x = reinterpret_cast<const int32_t*>(st)[r];
              reinterpret_cast<int32_t*>(dst)[r] = x < 0 ? -1 : trans[x];
            }
          } else {
            for (int r = lane; r < rows; r += 32) reinterpret_cast<uint32_t*>(dst)[r] = reinterpret_cast<const uint32_t*>(st)[r];
          }
          break;
        case TD_N8:
          if (d.kind == 2) {  // best-table entries {action, next}: only `next` is an index (TzTree.best encoding)
            for (int r = lane; r < rows; r += 32) {
              int2 e = reinterpret_cast<const int2*>(st)[r];
              if (e.y >= 0) e.y = trans[e.y];
              else if (e.y <= -2) e.y = -(trans[-(e.y + 2)] + 2);
              reinterpret_cast<int2*>(dst)[r] = e;
            }
          } else {
            for (int r = lane; r < rows; r += 32) reinterpret_cast<uint2*>(dst)[r] = reinterpret_cast<const uint2*>(st)[r];
          }
          break;
        case TD_N16:
          for (int r = lane; r < rows; r += 32) reinterpret_cast<uint4*>(dst)[r] = reinterpret_cast<const uint4*>(st)[r];
          break;
        case TD_BYTE:  // one-byte rows staged as the words that hold them
          for (int r = lane; r < rows; r += 32) dst[r] = st[4 * (size_t)r + ((uintptr_t)(d.base + src_of[s0 + r]) & 3)];
          break;
        case TD_UNITS: {
          const size_t nbytes = (size_t)rows * d.rb;
          if (d.kind == 1) {  // every word is a node index
            for (size_t i = lane; i < (nbytes >> 2); i += 32) {
              const int32_t x = reinterpret_cast<const int32_t*>(st)[i];
              reinterpret_cast<int32_t*>(dst)[i] = x < 0 ? -1 : trans[x];
            }
          } else if ((((uintptr_t)dst | nbytes) & 15) == 0) {
            for (size_t i = lane; i < (nbytes >> 4); i += 32) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(st)[i];
          } else if ((((uintptr_t)dst | nbytes) & 3) == 0) {
            for (size_t i = lane; i < (nbytes >> 2); i += 32) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(st)[i];
          } else {
            for (size_t i = lane; i < nbytes; i += 32) dst[i] = st[i];
          }
          break;
        }
        default:
          break;
      }
    }
    if (stored_bulk) {  // (the issuing lanes) the staging area may be overwritten once the bulk stores have READ it
      bulk_commit();
      bulk_wait_read0();
    }
    __syncthreads();  // the staging area is reused by the next chunk
    if (s0 == 0) TZ_RSTAMP(6);
#ifdef TZ_PROFILE
    { const long long t_b = prof_gtime(); acc_scatter += t_b - t_a; t_a = t_b; }
#endif
  }
  TZ_RSTAMP(7);
  // (4) null the tail rows [count, nfi) (tree.py:236-238,247-249): after every source row has been read
  for (int t = 0; t < P.pre_first; ++t) {  // (the preloaded narrow tables were done in (2b))
    const int64_t rb = P.tab[t].rb;
    uint8_t* const base = P.tab[t].base + (size_t)b * N * rb;
    if (P.tab[t].kind == 3) {
      int4* e = reinterpret_cast<int4*>(base);
      for (size_t i = (size_t)count * (rb >> 4) + tid; i < (size_t)nfi * (rb >> 4); i += nthr) e[i] = make_int4(0, 0, 0, -1);
    } else {
      block_fill(base, (size_t)count * rb, (size_t)nfi * rb, P.tab[t].null_pattern);
    }
  }
  if (tid == 0) P.nfi[b] = count;
  TZ_RSTAMP(8);
#ifdef TZ_PROFILE
  if (tid == 0 && b < 4096) {
    g_prof_gt[16 * b + 9] = nfi;
    g_prof_gt[16 * b + 10] = count;
    g_prof_gt[16 * b + 11] = acc_issue;
    g_prof_gt[16 * b + 12] = acc_wait;
    g_prof_gt[16 * b + 13] = acc_scatter;
    g_prof_gt[16 * b + 14] = (count + rpc - 1) / rpc;
  }
#endif
}

}  // namespace

extern "C" {

int tz_reroot(const TzTree* t, const int32_t* action, const uint8_t* reset_flag, int persist_tree, tz_stream_t stream) {
  const int rc = check_tree(t);
  if (rc) return rc;
  if (persist_tree && !action) return TZ_EINVAL;
  // ---- all tables of a tree moved together (k_reroot_all) whenever one row of every table fits the staging area ----
  RerootP P = {};
  P.B = t->B;
  P.N = t->N;
  P.F = t->F;
  P.nfi = t->next_free_idx;
  P.parents = t->parents;
  P.edge = t->edge_map;
  P.stats = t->stats;
  int nt = 0;
  auto add = [&](void* base, int64_t rb, int kind, uint32_t pat) {
    P.tab[nt].base = reinterpret_cast<uint8_t*>(base);
    P.tab[nt].rb = rb;
    P.tab[nt].kind = kind;
    P.tab[nt].null_pattern = pat;
    ++nt;
  };
  const int64_t F = t->F;
  // widest first; rows that need ordinary loads (size not a multiple of 4) last, so that their loads overlap the copies in flight
  add(t->child_stats, 16 * F, 3, 0u);
  for (int k = 0; k < t->n_emb; ++k)
    if ((t->emb_row_bytes[k] & 3) == 0) add(t->emb[k], t->emb_row_bytes[k], 0, 0u);
  add(t->p, 4 * F, 4, 0u);                  // rebuilt from the child_stats rows: not gathered, not staged
  add(t->edge_map, 4 * F, 5, 0xffffffffu);  // (child_stats is table 0, so its chunk sits at the start of the staging area)
  add(t->best, 8, 2, 0xffffffffu);
  add(t->parents, 4, 1, 0xffffffffu);
  add(t->n, 4, 0, 0u);
  add(t->q, 4, 0, 0u);
  if (t->r) add(t->r, 4, 0, 0u);
  for (int k = 0; k < t->n_emb; ++k)
    if ((t->emb_row_bytes[k] & 3) != 0) add(t->emb[k], t->emb_row_bytes[k], 0, 0u);
  add(t->terminated, 1, 0, 0u);
  P.ntab = nt;
  for (int k = 0; k < nt; ++k) {  // gather granularity per table: by row size and base alignment (every tree's block starts at
    RerootTab& tb = P.tab[k];     // base + b * N * rb, so rb's alignment covers all of them)
    const uintptr_t al = (uintptr_t)tb.base | (uintptr_t)tb.rb;
    tb.unit = (al & 15) == 0 ? 16u : ((al & 7) == 0 ? 8u : ((al & 3) == 0 ? 4u : 1u));
    tb.units = (uint32_t)(tb.rb / tb.unit);
    // exact for every copy index of a chunk (i < units * N) iff units^2 * N < 2^32; otherwise the kernel divides
    const bool exact = (uint64_t)tb.units * tb.units * (uint64_t)t->N < (1ull << 32);
    tb.magic = (exact && tb.units > 1) ? (uint32_t)(0x100000000ull / tb.units) + 1u : 0u;
    tb.pad = 0;
  }
  for (int k = 0; k < nt; ++k) P.tab[k].pad = (uint32_t)P.tab[k].rb;  // staged bytes per row
  // TZ_REROOT_IMPL=ldgsts selects the previous gather (per-thread LDGSTS copies) for A/B measurements; default: bulk copies
  static const bool use_ldgsts = [] {
    const char* e = getenv("TZ_REROOT_IMPL");
    return e != nullptr && strcmp(e, "ldgsts") == 0;
  }();
  const int64_t per_sm = 227 * 1024;
  if (!use_ldgsts) {
    // ---- k_reroot_bulk -----------------------------------------------------------------------------------------------
    int64_t bulk_bytes = 0, row_total = 0, pre_bytes = 0;
    // narrow tables are preloaded whole into shared memory when that costs little of it (N * 21 bytes: trees of <= 512 nodes)
    int64_t narrow = 0;
    for (int k = 0; k < nt; ++k)
      if (P.tab[k].kind <= 2 && (P.tab[k].rb == 4 || P.tab[k].rb == 8 || P.tab[k].rb == 1)) narrow += (P.tab[k].rb * (int64_t)t->N + 15) & ~(int64_t)15;
    const bool preload = narrow <= 12 * 1024 && (t->N & 3) == 0;
    for (int k = 0; k < nt; ++k) {
      RerootTab& tb = P.tab[k];
      if (tb.kind == 4) P.p_tab = k;
      if (tb.kind == 5) P.e_tab = k;
      if (tb.kind >= 4) continue;
      if (preload && tb.kind <= 2 && (tb.rb == 4 || tb.rb == 8 || tb.rb == 1) && (((uintptr_t)tb.base) & 3) == 0) {
        tb.unit = 6;  // copied whole to shared memory, compacted from there
        tb.magic = (uint32_t)pre_bytes;
        tb.pad = 0;
        pre_bytes += (tb.rb * (int64_t)t->N + 15) & ~(int64_t)15;
        continue;
      }
      const bool aligned = (((uintptr_t)tb.base | (uintptr_t)tb.rb) & 15) == 0;
      // (16-byte rows stay with one LDGSTS each: the bulk-copy unit takes ~4 cycles per copy whatever its size)
      if ((tb.kind == 3 || tb.kind == 0) && aligned && tb.rb >= 32 && tb.rb < (1 << 20)) {
        tb.unit = 0;  // moved by cp.async.bulk
        bulk_bytes += tb.rb;
      } else if (tb.kind == 0 && tb.rb == 1) {
        tb.unit = 5;  // one-byte rows travel as the aligned word that holds them
        tb.pad = 4;
      }
      row_total += tb.pad;
    }
    {  // preloaded tables last (stable), so that the kernel's loops need no per-table test
      RerootTab tmp[REROOT_MAX_TABS];
      int m = 0;
      for (int k = 0; k < nt; ++k)
        if (P.tab[k].unit != 6) tmp[m++] = P.tab[k];
      P.pre_first = m;
      for (int k = 0; k < nt; ++k)
        if (P.tab[k].unit == 6) tmp[m++] = P.tab[k];
      for (int k = 0; k < nt; ++k) {
        P.tab[k] = tmp[k];
        if (P.tab[k].kind == 4) P.p_tab = k;
        if (P.tab[k].kind == 5) P.e_tab = k;
      }
    }
    P.bulk_row_bytes = (int32_t)bulk_bytes;
    // Staging area: as many CTAs per SM as let every tree of the batch be resident at once (one wave over the 148 SMs, at most
    // 7) -- unless a chunk would then hold fewer than 8 destination rows (wide rows: go_9x9's are 5.4 KB): every chunk costs a
    // memory round trip, two barriers and a pass over the table list, so fewer, larger CTAs (and several waves) move more
    // bytes per second.  Measured on the go_9x9 shape: 3 rows per chunk at 7 CTAs per SM = 2.4 TB/s (profiles/r2d_phase_reroot_go.log).
    int ctas = (t->B + 147) / 148;
    ctas = ctas < 1 ? 1 : (ctas > 7 ? 7 : ctas);
    // (the kernels' static shared memory -- descriptors, the root's edge row -- counts against the SM like the dynamic part)
    static const int64_t static_smem[3] = {static_shared_bytes(k_reroot_bulk<128>), static_shared_bytes(k_reroot_bulk<256>),
                                           static_shared_bytes(k_reroot_bulk<512>)};
    auto threads_for = [](int c) { return c <= 3 ? 512 : (c <= 5 ? 256 : 128); };
    auto stage_for = [&](int c) {
      const int64_t stat = static_smem[threads_for(c) == 128 ? 0 : (threads_for(c) == 256 ? 1 : 2)];
      int64_t st = per_sm / c - 1024 - stat - 8 * (int64_t)t->N - pre_bytes - 64;  // (1 KB per-CTA reserve, static shared memory, index scratch, preloaded tables)
      st = st > 160 * 1024 ? 160 * 1024 : st;
      return st & ~(int64_t)15;
    };
    while (ctas > 2 && (stage_for(ctas) - 16 * nt) / row_total < 8) --ctas;
    const int64_t stage = stage_for(ctas);
    int64_t rpc = stage > 0 ? (stage - 16 * nt) / row_total : 0;
    // the mbarrier's transaction count is 20 bits: a chunk's bulk bytes stay below 1 MiB (the staging area is <= 160 KB)
    if (rpc >= 1) {
      P.stage_bytes = (int32_t)stage;
      P.rpc = (int32_t)(rpc > t->N ? t->N : rpc);
      size_t off = 0;
      for (int k = 0; k < nt; ++k) {  // how every table's rows travel, and where they are staged
        RerootTab& tb = P.tab[k];
        if (tb.unit == 6) {
          tb.st_off = tb.magic;
          tb.code = TD_PRE;
          continue;
        }
        tb.st_off = (uint32_t)off;
        tb.code = tb.kind >= 4 ? TD_SKIP
                  : tb.unit == 0 ? TD_BULK
                  : tb.unit == 5 ? TD_BYTE
                  : (tb.units == 1 && tb.unit == 4) ? TD_N4
                  : (tb.units == 1 && tb.unit == 8) ? TD_N8
                  : (tb.units == 1 && tb.unit == 16) ? TD_N16
                                                     : TD_UNITS;
        if (tb.kind < 4) off += ((size_t)P.rpc * tb.pad + 15) & ~(size_t)15;
      }
      const size_t smem = (size_t)stage + 8 * (size_t)t->N + (size_t)pre_bytes;
      // threads per CTA by CTAs per SM (pointer jumping and the scan over up to N nodes, the child_stats pass of the scatter):
      // 512 when at most 3 CTAs share an SM, 256 up to 5, else 128
      const int nthreads = threads_for(ctas);
      auto kernel = nthreads == 512 ? k_reroot_bulk<512> : (nthreads == 256 ? k_reroot_bulk<256> : k_reroot_bulk<128>);
      if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
      }
      kernel<<<t->B, nthreads, smem, (cudaStream_t)stream>>>(P, action, reset_flag, persist_tree);
      return launch_status();
    }
    for (int k = 0; k < nt; ++k) {  // (fall through to the one-table-at-a-time kernel below with the original descriptors)
      P.tab[k].pad = (uint32_t)P.tab[k].rb;
    }
  }
  int64_t row_total = 0;
  for (int k = 0; k < nt; ++k) row_total += P.tab[k].kind >= 4 ? 0 : P.tab[k].rb;
  // k_reroot_all: staging area as large as lets every tree of the batch be resident at once (one wave over the 148 SMs), within
  // [16 KB, 64 KB]; the index scratch (8 N bytes) and 1 KB of per-CTA reserve come on top
  const int ctas_wanted = (t->B + 147) / 148;
  int64_t stage = per_sm / (ctas_wanted < 1 ? 1 : ctas_wanted) - 1024 - 8 * (int64_t)t->N - 64;
  stage = stage > 64 * 1024 ? 64 * 1024 : stage;
  stage = stage < 16 * 1024 ? 16 * 1024 : stage;
  stage &= ~(int64_t)15;
  const int64_t rpc = (stage - 16 * nt) / row_total;
  if (use_ldgsts && rpc >= 1 && stage + 8 * (int64_t)t->N <= 200 * 1024) {
    P.stage_bytes = (int32_t)stage;
    P.rpc = (int32_t)(rpc > t->N ? t->N : rpc);
    const size_t smem = (size_t)stage + 8 * (size_t)t->N;
    if (smem > 48 * 1024) {
      const cudaError_t e = cudaFuncSetAttribute(k_reroot_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    k_reroot_all<<<t->B, REROOT2_THREADS, smem, (cudaStream_t)stream>>>(P, action, reset_flag, persist_tree);
    return launch_status();
  }
  // ---- very wide rows: one table at a time (k_reroot) -------------------------------------------------------------
  int64_t max_rb = 16 * (int64_t)t->F;
  for (int k = 0; k < t->n_emb; ++k) max_rb = t->emb_row_bytes[k] > max_rb ? t->emb_row_bytes[k] : max_rb;
  if (max_rb > REROOT_STAGE) return TZ_ENOTSUP;
  const size_t smem = (size_t)REROOT_STAGE + 8 * (size_t)t->N;
  if (smem > 227 * 1024) return TZ_ENOTSUP;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(k_reroot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  k_reroot<<<t->B, REROOT_THREADS, smem, (cudaStream_t)stream>>>(*t, action, reset_flag, persist_tree);
  return launch_status();
}

}  // extern "C"
