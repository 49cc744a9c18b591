/*
 * tz_oracle.c -- ORACLE (test infrastructure, NOT product code): plain-C CPU restatement of the reference's
 * batched MCTS path, one tree at a time, OpenMP over trees.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * oracle/_build/libtz_oracle.so.  The product library (libtz_b200.so) neither links nor calls it.
 *
 * PARITY PINNING: the reference (lowrollr/turbozero) has no tests or golden vectors and JAX cannot be installed here.
 * This file is pinned by golden fixtures produced by the reference's own unmodified source executed on a NumPy
 * emulation of the jax API (tests/golden/, oracle/jaxshim/README.md), by hand-derived known answers
 * and by agreement with the independently written literal NumPy restatement oracle/mcts_numpy.py (different
 * algorithms where the reference allows: e.g. re-rooting here is a single forward sweep using parents[i] < i, there it
 * is the reference's N-1 label-propagation rounds + scatter).  XLA's own float code generation (FMA contraction,
 * exp/log/pow) remains unpinned -- "parity unpinned" for those bits (DESIGN.md "Residual risk").
 *
 * Host arrays use the SAME struct (TzTree/TzWork/TzSearchCfg, include/tz_abi.h) and layout as the device path.
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math (oracle/build.py).  Reference citations are relative
 * to the reference repo root.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "tz_abi.h"
#include "tz_math.h"
#include "tz_synth.h"

#define EXPORT __attribute__((visibility("default")))
#define MAXF 1024
#define PAR_MIN 64 /* below this many trees a parallel region costs more than it saves */

/* ---- canonical float sum (32 strided partials + xor butterfly), see oracle/mcts_numpy.py canon_sum ---- */
static float canon_sum(const float* x, int n) {
  float v[32], t[32];
  for (int l = 0; l < 32; ++l) v[l] = 0.0f;
  for (int j = 0; j < n; ++j) v[j & 31] = v[j & 31] + x[j];
  for (int off = 16; off >= 1; off >>= 1) {
    for (int l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ off];
    memcpy(v, t, sizeof v);
  }
  return v[0];
}

static void softmax(const float* logit, float* out, int F) {
  float m = logit[0];
  for (int a = 1; a < F; ++a) m = logit[a] > m ? logit[a] : m;
  for (int a = 0; a < F; ++a) out[a] = tz_expf(logit[a] - m);
  float s = canon_sum(out, F);
  for (int a = 0; a < F; ++a) out[a] = out[a] / s;
}

/* ---- per-tree view ---- */
typedef struct {
  int N, F;
  int32_t* nfi;
  int32_t* parents;
  int32_t* edge;
  int32_t* n;
  float* p;
  float* q;
  float* r;
  uint8_t* term;
  int n_emb;
  uint8_t* emb[TZ_MAX_EMB];
  int64_t rb[TZ_MAX_EMB];
  uint64_t* stats;
} View;

static View view(const TzTree* t, int b) {
  View v;
  size_t N = (size_t)t->N, F = (size_t)t->F;
  v.N = t->N;
  v.F = t->F;
  v.nfi = t->next_free_idx + b;
  v.parents = t->parents + b * N;
  v.edge = t->edge_map + b * N * F;
  v.n = t->n + b * N;
  v.p = t->p + b * N * F;
  v.q = t->q + b * N;
  v.r = t->r ? t->r + b * N : NULL;
  v.term = t->terminated + b * N;
  v.n_emb = t->n_emb;
  for (int k = 0; k < t->n_emb; ++k) {
    v.rb[k] = t->emb_row_bytes[k];
    v.emb[k] = (uint8_t*)t->emb[k] + (size_t)b * N * (size_t)v.rb[k];
  }
  v.stats = t->stats ? t->stats + 4 * (size_t)b : NULL;
  return v;
}

/* action_selection.py:10-32 normalize_q_values on an F-vector */
static void normalize_q(const float* dq, const int32_t* cn, float parent_q, float eps, float* out, int F) {
  float mn = dq[0], mx = dq[0];
  for (int a = 1; a < F; ++a) {
    mn = dq[a] < mn ? dq[a] : mn;
    mx = dq[a] > mx ? dq[a] : mx;
  }
  mn = parent_q < mn ? parent_q : mn;
  mx = parent_q > mx ? parent_q : mx;
  float denom = mx - mn;
  denom = denom > eps ? denom : eps;
  for (int a = 0; a < F; ++a) {
    float c = cn[a] > 0 ? dq[a] : mn;
    out[a] = (c - mn) / denom;
  }
}

/* tree.py:78-98 get_child_data for q and n */
static void child_qn(const View* v, int i, float discount, float* dq, int32_t* cn) {
  const int32_t* row = v->edge + (size_t)i * v->F;
  for (int a = 0; a < v->F; ++a) {
    int c = row[a];
    float cq = c < 0 ? 0.0f : v->q[c];
    cn[a] = c < 0 ? 0 : v->n[c];
    dq[a] = cq * discount;
  }
}

/* action_selection.py:91-116 (PUCT) and :150-177 (MuZero PUCT, intended arity) */
static int select_action(const View* v, int i, const TzSearchCfg* cfg) {
  float dq[MAXF], qn[MAXF];
  int32_t cn[MAXF];
  int F = v->F;
  child_qn(v, i, cfg->discount, dq, cn);
  normalize_q(dq, cn, v->q[i], cfg->epsilon, qn, F);
  if (cfg->q_transform == TZ_QT_IDENTITY) /* the q_transform registry (include/tz_abi.h TZ_QT_*), action_selection.py:70,109 */
    for (int a = 0; a < F; ++a) qn[a] = dq[a];
  float sq = sqrtf((float)v->n[i]);
  const float* p = v->p + (size_t)i * F;
  float log_term = 0.0f;
  if (cfg->selector == TZ_SEL_MUZERO_PUCT) {
    float t = ((float)v->n[i] + cfg->c2) + 1.0f;
    log_term = tz_logf(t / cfg->c2) + cfg->c1;
  }
  int best = 0;
  float best_s = 0.0f;
  for (int a = 0; a < F; ++a) {
    float denom = (float)(cn[a] + 1);
    float u;
    if (cfg->selector == TZ_SEL_MUZERO_PUCT)
      u = ((p[a] * sq) / denom) * log_term;
    else
      u = ((cfg->c * p[a]) * sq) / denom;
    float s = qn[a] + u;
    if (a == 0 || s > best_s) {
      best_s = s;
      best = a;
    }
  }
  return best;
}

/* mcts.py:322 */
static float backup_q(float q, int n, float value, int fma) {
  float fn = (float)n;
  float num = fma ? fmaf(q, fn, value) : (q * fn) + value;
  return num / (float)(n + 1);
}

/* mcts.py:192-228 */
static void traverse(const View* v, const TzSearchCfg* cfg, int* parent_out, int* action_out, int* levels_out) {
  int parent = TZ_ROOT_INDEX;
  int action = select_action(v, parent, cfg);
  int levels = 1;
  for (;;) {
    int child = v->edge[(size_t)parent * v->F + action];
    if (child < 0 || v->term[child]) break;
    parent = child;
    action = select_action(v, parent, cfg);
    ++levels;
  }
  *parent_out = parent;
  *action_out = action;
  *levels_out = levels;
}

#ifdef TZO_TRACE
/* histogram over simulations of (path length L, first level d at which the walk's node differs from the previous walk's),
   both clamped to 255; per thread: the previous path */
EXPORT long long tzo_trace_hist[256][256];
static __thread int trace_prev[4096], trace_prev_len;
static void tzo_trace_walk(const View* v, int parent, int levels, int first) {
  int cur[4096], L = levels < 4096 ? levels : 4096, x = parent;
  for (int l = L - 1; l >= 0 && x >= 0; --l) {
    cur[l] = x;
    x = v->parents[x];
  }
  int d = 0;
  if (!first)
    while (d < L && d < trace_prev_len && cur[d] == trace_prev[d]) ++d;
  /* d = levels shared with the previous walk (level 0, the root, is always shared) */
  if (!first) {
    int pl = trace_prev_len < 255 ? trace_prev_len : 255, dd = d < 255 ? d : 255;
#pragma omp atomic
    tzo_trace_hist[pl][dd] += 1;
  }
  for (int l = 0; l < L; ++l) trace_prev[l] = cur[l];
  trace_prev_len = L;
}
#endif

/* mcts.py:174-187 + tree.py:101-132,153-166 */
static void expand(View* v, int parent, int action, const float* policy, float value, uint8_t term,
                   void* const* new_emb, int b, const TzSearchCfg* cfg) {
  int F = v->F;
  int node = v->edge[(size_t)parent * F + action];
  if (node >= 0) {
    v->q[node] = backup_q(v->q[node], v->n[node], value, cfg->fma_backup);
    v->n[node] += 1;
  } else {
    int nfi = *v->nfi;
    if (nfi >= v->N) return; /* full: edge stays -1, row write dropped (tree.py:116-131) */
    node = nfi;
    v->parents[node] = parent;
    v->n[node] = 1;
    v->q[node] = value;
    if (v->r) v->r[node] = value;
    v->edge[(size_t)parent * F + action] = node;
    *v->nfi = nfi + 1;
  }
  memcpy(v->p + (size_t)node * F, policy, sizeof(float) * F);
  v->term[node] = term ? 1 : 0;
  for (int k = 0; k < v->n_emb; ++k)
    memcpy(v->emb[k] + (size_t)node * v->rb[k], (const uint8_t*)new_emb[k] + (size_t)b * v->rb[k], (size_t)v->rb[k]);
}

/* mcts.py:231-262 */
static void backprop(View* v, int parent, float value, const TzSearchCfg* cfg) {
  int node = parent;
  float val = value;
  while (node != TZ_NULL_INDEX) {
    val = val * cfg->discount;
    v->q[node] = backup_q(v->q[node], v->n[node], val, cfg->fma_backup);
    v->n[node] += 1;
    node = v->parents[node];
  }
}

/* weighted_mcts.py:90-152 */
static void weighted_backprop(View* v, int parent, const TzSearchCfg* cfg, const float* noise) {
  float dq[MAXF], nq[MAXF], vals[MAXF], logit[MAXF], w[MAXF];
  int32_t cn[MAXF];
  int F = v->F;
  int node = parent;
  while (node != TZ_NULL_INDEX) {
    child_qn(v, node, cfg->discount, dq, cn);
    normalize_q(dq, cn, v->q[node], TZ_FLT_EPS, nq, F);
    if (cfg->inv_q_temperature > 0.0f) {
      for (int a = 0; a < F; ++a) {
        vals[a] = tz_powf(nq[a], cfg->inv_q_temperature);
        logit[a] = cn[a] > 0 ? nq[a] : -TZ_FLT_MAX;
      }
    } else {
      int imax = 0;
      float best = 0.0f;
      for (int a = 0; a < F; ++a) {
        float s = nq[a] + noise[a];
        if (a == 0 || s > best) {
          best = s;
          imax = a;
        }
        logit[a] = -TZ_FLT_MAX;
        vals[a] = nq[a];
      }
      logit[imax] = 1.0f;
    }
    softmax(logit, w, F);
    for (int a = 0; a < F; ++a) w[a] = w[a] * vals[a];
    float qw = canon_sum(w, F);
    v->q[node] = backup_q(qw, v->n[node], v->r[node], cfg->fma_backup);
    v->n[node] += 1;
    node = v->parents[node];
  }
}

/* mcts.py:363-384 / weighted_mcts.py:66-87 + tree.py:135-150 */
static void set_root(View* v, const float* pol, float val, void* const* root_emb, int b) {
  int visited = v->n[0] > 0;
  memcpy(v->p, pol, sizeof(float) * v->F);
  if (!visited) {
    v->q[0] = val;
    v->n[0] = 1;
    if (v->r) v->r[0] = val;
  }
  for (int k = 0; k < v->n_emb; ++k)
    memcpy(v->emb[k], (const uint8_t*)root_emb[k] + (size_t)b * v->rb[k], (size_t)v->rb[k]);
  if (*v->nfi < 1) *v->nfi = 1;
}

/* mcts.py:265-296 + 111-120 */
static int root_action(const View* v, float temperature, const float* noise, const float* uniform01,
                       int32_t* visits_out, float* pw_out, float* q_out) {
  int F = v->F;
  int32_t vis[MAXF];
  float pw[MAXF], pwt[MAXF];
  int total = 0;
  for (int a = 0; a < F; ++a) {
    int c = v->edge[a];
    vis[a] = c < 0 ? 0 : v->n[c];
    total += vis[a];
  }
  float ftot = (float)(total > 1 ? total : 1);
  float unif = (float)(1.0 / (double)F);
  for (int a = 0; a < F; ++a) pw[a] = total > 0 ? (float)vis[a] / ftot : unif;
  if (visits_out) memcpy(visits_out, vis, sizeof(int32_t) * F);
  if (pw_out) memcpy(pw_out, pw, sizeof(float) * F);
  if (q_out) *q_out = v->q[0];
  int action = 0;
  if (temperature == 0.0f) {
    if (!noise) return -1;
    float best = 0.0f;
    for (int a = 0; a < F; ++a) {
      float s = pw[a] + noise[a];
      if (a == 0 || s > best) {
        best = s;
        action = a;
      }
    }
  } else {
    if (!uniform01) return -1;
    float inv_t = (float)(1.0 / (double)temperature);
    for (int a = 0; a < F; ++a) pwt[a] = tz_powf(pw[a], inv_t);
    float s = canon_sum(pwt, F);
    float acc = 0.0f, last = 0.0f;
    for (int a = 0; a < F; ++a) {
      pwt[a] = pwt[a] / s;
      acc = acc + pwt[a];
      pwt[a] = acc; /* cumsum */
      last = acc;
    }
    float rr = last * (1.0f - *uniform01);
    for (int a = 0; a < F; ++a) action += pwt[a] < rr ? 1 : 0;
  }
  return action;
}

static void clear_rows(View* v, int lo, int hi) {
  if (hi <= lo) return;
  size_t cnt = (size_t)(hi - lo), F = (size_t)v->F;
  for (size_t i = 0; i < cnt; ++i) v->parents[lo + i] = -1;
  for (size_t i = 0; i < cnt * F; ++i) v->edge[(size_t)lo * F + i] = -1;
  memset(v->n + lo, 0, cnt * 4);
  memset(v->p + (size_t)lo * F, 0, cnt * F * 4);
  memset(v->q + lo, 0, cnt * 4);
  if (v->r) memset(v->r + lo, 0, cnt * 4);
  memset(v->term + lo, 0, cnt);
  for (int k = 0; k < v->n_emb; ++k) memset(v->emb[k] + (size_t)lo * v->rb[k], 0, cnt * (size_t)v->rb[k]);
}

/* tree.py:169-269 get_subtree; tree.py:272-278 reset; common.py:89-94 select between them */
static void reroot(View* v, int action, int do_reset, int32_t* label /*N*/, int32_t* trans /*N*/) {
  int nfi = *v->nfi, F = v->F;
  int c = do_reset ? -1 : v->edge[action];
  if (v->stats) v->stats[2] += (uint64_t)nfi;
  if (c < 0) { /* absent child or reset: everything erased */
    clear_rows(v, 0, nfi);
    *v->nfi = 0;
    return;
  }
  /* forward label sweep: valid because every allocated node has parents[i] < i */
  label[0] = 0;
  int count = 0;
  trans[0] = -1;
  for (int i = 1; i < nfi; ++i) {
    int par = v->parents[i];
    label[i] = par == 0 ? i : label[par];
    trans[i] = label[i] == c ? count++ : -1;
  }
  for (int i = 1; i < nfi; ++i) {
    int d = trans[i];
    if (d < 0) continue;
    int par = v->parents[i];
    v->parents[d] = par < 0 ? -1 : trans[par];
    for (int a = 0; a < F; ++a) {
      int e = v->edge[(size_t)i * F + a];
      v->edge[(size_t)d * F + a] = e < 0 ? -1 : trans[e];
    }
    v->n[d] = v->n[i];
    v->q[d] = v->q[i];
    if (v->r) v->r[d] = v->r[i];
    v->term[d] = v->term[i];
    memmove(v->p + (size_t)d * F, v->p + (size_t)i * F, sizeof(float) * F);
    for (int k = 0; k < v->n_emb; ++k)
      memmove(v->emb[k] + (size_t)d * v->rb[k], v->emb[k] + (size_t)i * v->rb[k], (size_t)v->rb[k]);
  }
  clear_rows(v, count, nfi);
  *v->nfi = count;
  if (v->stats) v->stats[3] += (uint64_t)count;
}

/* ------------------------------------------------------------------------------------------------ */
/* exported host-memory mirror of the device C-ABI                                                    */
/* ------------------------------------------------------------------------------------------------ */
EXPORT int tzo_tree_init(const TzTree* t) {
#pragma omp parallel for schedule(static) if (t->B >= PAR_MIN)
  for (int b = 0; b < t->B; ++b) {
    View v = view(t, b);
    clear_rows(&v, 0, v.N);
    *v.nfi = 0;
  }
  return 0;
}

EXPORT int tzo_set_root(const TzTree* t, const float* root_policy, const float* root_value, void* const* root_emb) {
#pragma omp parallel for schedule(static) if (t->B >= PAR_MIN)
  for (int b = 0; b < t->B; ++b) {
    View v = view(t, b);
    set_root(&v, root_policy + (size_t)b * t->F, root_value[b], root_emb, b);
  }
  return 0;
}

EXPORT int tzo_select(const TzTree* t, const TzSearchCfg* cfg, const TzWork* w) {
  if (t->F > MAXF) return TZ_ENOTSUP;
#pragma omp parallel for schedule(static) if (t->B >= PAR_MIN)
  for (int b = 0; b < t->B; ++b) {
    View v = view(t, b);
    int parent, action, levels;
    traverse(&v, cfg, &parent, &action, &levels);
    w->parent[b] = parent;
    w->action[b] = action;
    for (int k = 0; k < v.n_emb; ++k)
      memcpy((uint8_t*)w->emb_parent[k] + (size_t)b * v.rb[k], v.emb[k] + (size_t)parent * v.rb[k], (size_t)v.rb[k]);
    if (v.stats) {
      v.stats[0] += (uint64_t)levels;
      v.stats[1] += 1;
    }
  }
  return 0;
}

EXPORT int tzo_expand_backprop(const TzTree* t, const TzSearchCfg* cfg, const TzWork* w) {
  if (t->F > MAXF) return TZ_ENOTSUP;
#pragma omp parallel for schedule(static) if (t->B >= PAR_MIN)
  for (int b = 0; b < t->B; ++b) {
    View v = view(t, b);
    int parent = w->parent[b], action = w->action[b];
    float value = w->value[b];
    expand(&v, parent, action, w->policy + (size_t)b * t->F, value, w->terminated[b], w->emb_new, b, cfg);
    if (cfg->weighted)
      weighted_backprop(&v, parent, cfg, w->backprop_noise ? w->backprop_noise + (size_t)b * t->F : NULL);
    else
      backprop(&v, parent, value, cfg);
  }
  return 0;
}

EXPORT int tzo_root_action(const TzTree* t, float temperature, const float* noise, const float* uniform01,
                           int32_t* visits, float* policy_weights, float* root_q, int32_t* action) {
  if (t->F > MAXF) return TZ_ENOTSUP;
  int bad = 0;
#pragma omp parallel for schedule(static) if (t->B >= PAR_MIN)
  for (int b = 0; b < t->B; ++b) {
    View v = view(t, b);
    size_t F = (size_t)t->F;
    int a = root_action(&v, temperature, noise ? noise + b * F : NULL, uniform01 ? uniform01 + b : NULL,
                        visits ? visits + b * F : NULL, policy_weights ? policy_weights + b * F : NULL,
                        root_q ? root_q + b : NULL);
    if (action) {
      if (a < 0) bad = 1;
      action[b] = a;
    }
  }
  return bad ? TZ_EINVAL : 0;
}

EXPORT int tzo_reroot(const TzTree* t, const int32_t* action, const uint8_t* reset_flag, int persist_tree) {
#pragma omp parallel if (t->B >= PAR_MIN)
  {
    int32_t* label = (int32_t*)malloc(sizeof(int32_t) * (size_t)t->N * 2);
    int32_t* trans = label + t->N;
#pragma omp for schedule(static)
    for (int b = 0; b < t->B; ++b) {
      View v = view(t, b);
      int flag = reset_flag ? reset_flag[b] : 0;
      if (flag == 2) continue; /* untouched: common.py:91 `lambda s: s` */
      int do_reset = !persist_tree || flag != 0;
      int a = do_reset ? 0 : action[b];
      a = a < 0 ? 0 : (a >= t->F ? t->F - 1 : a); /* XLA gather clamps */
      reroot(&v, a, do_reset, label, trans);
    }
    free(label);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* synthetic game on the host (same inline definition as the device stand-in, standin/include/tz_synth.h)     */
/* ------------------------------------------------------------------------------------------------ */
static void synth_write_state(const TzSynthGame* g, uint32_t h, int depth, int player, int32_t* core, uint8_t* payload) {
  core[0] = (int32_t)h;
  core[1] = depth;
  core[2] = player;
  core[3] = 0;
  int P = g->payload_bytes;
  for (int off = 0; off < P; off += 4) {
    uint32_t wv = tz_synth_payload_word(h, (uint32_t)(off / 4));
    int nb = P - off < 4 ? P - off : 4;
    memcpy(payload + off, &wv, (size_t)nb); /* little-endian host */
  }
}

static void synth_policy(const TzSynthGame* g, uint32_t h, int masked, float* pol) {
  float logit[MAXF];
  for (int a = 0; a < g->F; ++a)
    logit[a] = (!masked || tz_synth_legal(h, (uint32_t)a, g->rho256)) ? tz_synth_logit(h, (uint32_t)a) : -TZ_FLT_MAX;
  softmax(logit, pol, g->F);
}

static void synth_root_one(const TzSynthGame* g, const int32_t* core, const float* noise, float eps, float* pol, float* val) {
  uint32_t h = (uint32_t)core[0];
  int F = g->F;
  synth_policy(g, h, 0, pol);
  *val = tz_synth_value(h);
  if (!noise) return;
  float logit[MAXF];
  float one_minus = (float)(1.0 - (double)eps);
  for (int a = 0; a < F; ++a) {
    float noisy = (one_minus * pol[a]) + (eps * noise[a]);
    noisy = noisy > TZ_FLT_MIN ? noisy : TZ_FLT_MIN;
    float l = tz_logf(noisy);
    logit[a] = tz_synth_legal(h, (uint32_t)a, g->rho256) ? l : -TZ_FLT_MAX;
  }
  softmax(logit, pol, F);
}

static void synth_leaf_one(const TzSynthGame* g, const int32_t* pcore, int action, float* pol, float* val,
                           uint8_t* term, int32_t* ncore, uint8_t* npayload) {
  uint32_t h2 = tz_synth_step_h((uint32_t)pcore[0], (uint32_t)action);
  int d2 = pcore[1] + 1;
  int t = tz_synth_terminal(h2, d2, g->tau1024, g->max_depth);
  synth_policy(g, h2, 1, pol);
  *val = t ? tz_synth_reward(h2) : tz_synth_value(h2);
  *term = (uint8_t)t;
  synth_write_state(g, h2, d2, 1 - pcore[2], ncore, npayload);
}

EXPORT int tzo_synth_init_states(const TzSynthGame* g, int B, int env_offset, const int32_t* episode, int32_t* core, uint8_t* payload) {
#pragma omp parallel for schedule(static) if (B >= PAR_MIN)
  for (int b = 0; b < B; ++b)
    synth_write_state(g, tz_synth_init_h(g->seed, (uint32_t)(b + env_offset), (uint32_t)episode[b]), 0, 0,
                      core + 4 * (size_t)b, payload ? payload + (size_t)b * g->payload_bytes : NULL);
  return 0;
}

EXPORT int tzo_synth_root(const TzSynthGame* g, int B, const int32_t* core, const float* dir_noise, float dir_eps,
                          float* root_policy, float* root_value) {
  if (g->F > MAXF) return TZ_ENOTSUP;
#pragma omp parallel for schedule(static) if (B >= PAR_MIN)
  for (int b = 0; b < B; ++b)
    synth_root_one(g, core + 4 * (size_t)b, dir_noise ? dir_noise + (size_t)b * g->F : NULL, dir_eps,
                   root_policy + (size_t)b * g->F, root_value + b);
  return 0;
}

EXPORT int tzo_synth_leaf(const TzSynthGame* g, int B, const int32_t* parent_core, const int32_t* action, float* policy,
                          float* value, uint8_t* terminated, int32_t* new_core, uint8_t* new_payload) {
  if (g->F > MAXF) return TZ_ENOTSUP;
#pragma omp parallel for schedule(static) if (B >= PAR_MIN)
  for (int b = 0; b < B; ++b)
    synth_leaf_one(g, parent_core + 4 * (size_t)b, action[b], policy + (size_t)b * g->F, value + b, terminated + b,
                   new_core + 4 * (size_t)b, new_payload ? new_payload + (size_t)b * g->payload_bytes : NULL);
  return 0;
}

static void synth_env_step_one(const TzSynthGame* g, int env, int action, int32_t* core, uint8_t* payload,
                               int32_t* episode, uint8_t* reset_flag) {
  uint32_t h2 = tz_synth_step_h((uint32_t)core[0], (uint32_t)action);
  int d2 = core[1] + 1;
  if (tz_synth_terminal(h2, d2, g->tau1024, g->max_depth)) {
    *episode += 1;
    synth_write_state(g, tz_synth_init_h(g->seed, (uint32_t)env, (uint32_t)*episode), 0, 0, core, payload);
    *reset_flag = 1;
  } else {
    synth_write_state(g, h2, d2, 1 - core[2], core, payload);
    *reset_flag = 0;
  }
}

EXPORT int tzo_synth_env_step(const TzSynthGame* g, int B, int env_offset, const int32_t* action, int32_t* core,
                              uint8_t* payload, int32_t* episode, uint8_t* reset_flag) {
#pragma omp parallel for schedule(static) if (B >= PAR_MIN)
  for (int b = 0; b < B; ++b)
    synth_env_step_one(g, b + env_offset, action[b], core + 4 * (size_t)b,
                       payload ? payload + (size_t)b * g->payload_bytes : NULL, episode + b, reset_flag + b);
  return 0;
}

/*
 * Whole self-play on the CPU, tree-major: each thread owns a tree and runs `moves` x (set_root, S simulations,
 * root action, env step, re-root) on it.  This is the cache-friendliest CPU schedule (the reference's vmapped
 * dataflow is far slower) and is what bench.py times as cpu_baseline / --impl reference.
 *   dir_noise   [moves,B,F] or NULL     root_noise [moves,B,F] (temperature==0) or NULL
 *   uniform01   [moves,B]   (temperature>0) or NULL
 *   core [B,4], payload [B,P], episode [B]: env state, advanced in place
 *   actions_out [moves,B], pw_out [moves,B,F] optional
 */
EXPORT int tzo_selfplay(const TzTree* t, const TzSearchCfg* cfg, const TzSynthGame* g, int num_iterations, int moves,
                        float temperature, int persist_tree, int env_offset, const float* dir_noise, float dir_eps,
                        const float* root_noise, const float* uniform01, int32_t* core, uint8_t* payload,
                        int32_t* episode, int32_t* actions_out, float* pw_out, int nthreads) {
  if (t->F > MAXF || t->n_emb != (g->payload_bytes > 0 ? 2 : 1)) return TZ_ENOTSUP;
  if (cfg->weighted && (!(cfg->inv_q_temperature > 0.0f) || !t->r)) return TZ_ENOTSUP;
  int B = t->B, F = t->F, P = g->payload_bytes;
  int bad = 0;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    int32_t* label = (int32_t*)malloc(sizeof(int32_t) * (size_t)t->N * 2);
    float* pol = (float*)malloc(sizeof(float) * (size_t)F);
    uint8_t* npay = (uint8_t*)malloc((size_t)(P > 0 ? P : 1));
    int32_t ncore[4];
#pragma omp for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
      View v = view(t, b);
      int32_t* mycore = core + 4 * (size_t)b;
      uint8_t* mypay = payload ? payload + (size_t)b * P : NULL;
      for (int m = 0; m < moves; ++m) {
        size_t mb = (size_t)m * B + b;
        float rv;
        synth_root_one(g, mycore, dir_noise ? dir_noise + mb * F : NULL, dir_eps, pol, &rv);
        void* root_emb[2] = {core, payload}; /* set_root indexes by b */
        set_root(&v, pol, rv, root_emb, b);
        for (int s = 0; s < num_iterations; ++s) {
          int parent, action, levels;
          traverse(&v, cfg, &parent, &action, &levels);
#ifdef TZO_TRACE /* diagnostic build (scripts/trace_divergence.py): where does a walk leave the previous one? */
          tzo_trace_walk(&v, parent, levels, s == 0);
#endif
          float val;
          uint8_t term;
          const int32_t* pcore = (const int32_t*)(v.emb[0] + (size_t)parent * 16);
          synth_leaf_one(g, pcore, action, pol, &val, &term, ncore, npay);
          void* new_emb[2] = {ncore, npay};
          expand(&v, parent, action, pol, val, term, new_emb, 0, cfg);
          if (cfg->weighted)
            weighted_backprop(&v, parent, cfg, NULL);
          else
            backprop(&v, parent, val, cfg);
          if (v.stats) {
            v.stats[0] += (uint64_t)levels;
            v.stats[1] += 1;
          }
        }
        int a = root_action(&v, temperature, root_noise ? root_noise + mb * F : NULL, uniform01 ? uniform01 + mb : NULL,
                            NULL, pw_out ? pw_out + mb * F : NULL, NULL);
        if (a < 0) {
          bad = 1;
          a = 0;
        }
        if (actions_out) actions_out[mb] = a;
        uint8_t rf;
        synth_env_step_one(g, b + env_offset, a, mycore, mypay, episode + b, &rf);
        reroot(&v, a, rf || !persist_tree, label, label + t->N);
      }
    }
    free(label);
    free(pol);
    free(npay);
  }
  return bad ? TZ_EINVAL : 0;
}

EXPORT int tzo_num_threads(void) { return omp_get_max_threads(); }

/* the path's own exp / log / pow (include/tz_math.h), exported so tests can pin the NumPy restatement to them */
EXPORT int tzo_math(int which, const float* x, float y, float* out, int n) {
  for (int i = 0; i < n; ++i) out[i] = which == 0 ? tz_expf(x[i]) : which == 1 ? tz_logf(x[i]) : tz_powf(x[i], y);
  return 0;
}
