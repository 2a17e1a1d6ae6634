/*
 * hmm_oracle.c — CPU restatement of the Jstacs HMM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this file.  The product (jstacs_b200/) never links, imports or calls it.
 *
 * PARITY STATUS: "parity unpinned by reference tests".  The reference ships no golden vectors,
 * known-answer tests or fixtures for this path and no JVM exists in the build
 * container, so this restatement is pinned by (tests/test_oracle.py): brute-force enumeration of
 * all state paths on tiny models, the fwd/bwd consistency checks the reference keeps as comments
 * (AbstractHMM.java:1053-1065, HigherOrderHMM.java:1061-1095), a hand-computed 2-state example and
 * EM monotonicity.  java/Harness.java dumps the same vectors from Jstacs itself on any box with a JDK.
 *
 * Every function cites the reference lines it follows.  Abbreviations (relative to the reference root):
 *   HOH   = de/jstacs/sequenceScores/statisticalModels/trainable/hmm/models/HigherOrderHMM.java
 *   AHMM  = de/jstacs/sequenceScores/statisticalModels/trainable/hmm/AbstractHMM.java
 *   BHOT  = de/jstacs/sequenceScores/statisticalModels/trainable/hmm/transitions/BasicHigherOrderTransition.java
 *   ACDE  = de/jstacs/sequenceScores/statisticalModels/trainable/hmm/states/emissions/discrete/AbstractConditionalDiscreteEmission.java
 *   HODE  = .../states/emissions/discrete/HigherOrderDiscreteEmission.java
 *   NORM  = de/jstacs/utils/Normalisation.java
 *   TB    = de/jstacs/utils/ToolBox.java
 *
 * Arithmetic: fp64, libm exp/log (JDK Math.exp/log are <=1 ulp intrinsics; differences are
 * ~1e-16 relative, far inside the 1e-9 / 1e-6 gates; Viterbi uses only add/compare => exact).
 * Compile with -ffp-contract=off so no FMA contraction changes the association order.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/jstacs_hmm.h"

#define NEG_INF (-INFINITY)

typedef struct jo_model {
  int S, a, m;
  int *lc_ctx_ptr; /* [m+2] */
  int n_ctx_total, max_ctx;
  int *ctx_elem, *ctx_last_state;
  int n_elem, n_child_total;
  int *elem_child_ptr, *child_state, *child_desc, *elem_ctx, *trans_index;
  double *trans_param, *trans_lognorm, *trans_hyper;
  int n_emis, n_cond_total;
  int *state_emis, *emis_order, *emis_cond_ptr;
  double *emis_param, *emis_lognorm, *emis_hyper;
  uint8_t *state_silent, *state_final;
  double log_uniform; /* ACDE:298  logUniform = -Math.log(a) */
  int max_children, max_in; /* helper sizes, HOH:189-203 */
  /* allowed states groups (AHMM:146-228): preComputedContext as a flag per (group, context of layer class m) */
  int n_groups;
  uint8_t *allowed;              /* [n_groups * max_ctx] */
  const uint8_t *groups_codes;   /* the code buffer the attached per-position groups are aligned with */
  const int32_t *groups;         /* [n_res] group of every position, -1 = not constrained; NULL = no constraint at all */
} jo_model;

static void *dup_mem(const void *src, size_t bytes) {
  void *p = malloc(bytes ? bytes : 1);
  if (bytes) memcpy(p, src, bytes);
  return p;
}

static inline int lc_of(const jo_model *m, long layer) { return layer < m->m ? (int)layer : m->m; } /* BHOT:558-560 */
static inline int n_ctx(const jo_model *m, long layer) {                                           /* BHOT:592-594 */
  int lc = lc_of(m, layer);
  return m->lc_ctx_ptr[lc + 1] - m->lc_ctx_ptr[lc];
}
static inline int elem_of(const jo_model *m, long layer, int ctx) { /* BHOT:571-573 getTransitionElementIndex */
  return m->ctx_elem[m->lc_ctx_ptr[lc_of(m, layer)] + ctx];
}
static inline int n_children(const jo_model *m, long layer, int ctx) { /* BHOT:576-578 */
  int e = elem_of(m, layer, ctx);
  return m->elem_child_ptr[e + 1] - m->elem_child_ptr[e];
}
static inline int last_state(const jo_model *m, long layer, int ctx) { /* BHOT:627-634 */
  return m->ctx_last_state[m->lc_ctx_ptr[lc_of(m, layer)] + ctx];
}
/* BHOT:547-555 fillTransitionInformation: container = {state, next context, advance} */
static inline void trans_info(const jo_model *m, long layer, int ctx, int child, int c[3]) {
  int e = elem_of(m, layer, ctx);
  int p = m->elem_child_ptr[e] + child;
  c[0] = m->child_state[p];
  c[2] = m->state_silent[c[0]] ? 0 : 1;
  c[1] = m->elem_ctx[(size_t)lc_of(m, layer + c[2]) * m->n_elem + m->child_desc[p]];
}
/* BHOT:581-586 + :1116-1118  parameters[index] - logNorm (no transIndex indirection, "omitted for runtime reasons") */
static inline double trans_score(const jo_model *m, long layer, int ctx, int child) {
  int e = elem_of(m, layer, ctx);
  return m->trans_param[m->elem_child_ptr[e] + child] - m->trans_lognorm[e];
}

/* NORM:65-99 getLogSum(start,end,lnVal) */
static double log_sum(int start, int end, const double *v) {
  double offset = NEG_INF, sum = 0;
  for (int i = start; i < end; i++)
    if (v[i] > offset) offset = v[i];
  if (isinf(offset)) return NEG_INF;
  for (int i = start; i < end; i++) sum += exp(v[i] - offset);
  return offset + log(sum);
}
static double log_sum2(double x, double y) { /* NORM:45-47 with two arguments */
  double v[2] = {x, y};
  return log_sum(0, 2, v);
}
/* TB:124-138 ToolBox.max(start,end,array) */
static double tb_max(int start, int end, const double *v) {
  double mx = v[start];
  for (int i = start + 1; i < end; i++)
    if (mx < v[i]) mx = v[i];
  return mx;
}
/* NORM:259-271, 375-403 logSumNormalisation(d,0,n,dest,0): returns log of the sum */
static double log_sum_normalisation(const double *d, int n) {
  double offset = NEG_INF, sum = 0;
  for (int i = 0; i < n; i++) offset = fmax(offset, d[i]);
  if (offset == NEG_INF) return NEG_INF;
  for (int i = 0; i < n; i++) sum += exp(d[i] - offset);
  return offset + log(sum);
}

/* HODE:138-146 getConditionIndex, ACDE:431-457 getLogProbFor (forward strand) */
static inline int cond_index(const jo_model *m, int e, const uint8_t *x, long pos) {
  int k = m->emis_order[e], res = 0, pw = 1;
  for (int i = 0; i < k; i++) {
    long p = pos - (i + 1);
    if (p < 0) return -1;
    res += pw * x[p];
    pw *= m->a;
  }
  return res;
}
static double emis_score(const jo_model *m, int state, const uint8_t *x, long pos) {
  int e = m->state_emis[state];
  if (m->emis_order[e] < 0) return 0; /* SilentEmission.getLogProbFor returns 0 */
  double res = 0;
  int cond = cond_index(m, e, x, pos);
  if (cond < 0) {
    res += m->log_uniform;
  } else {
    int row = m->emis_cond_ptr[e] + cond;
    res -= m->emis_lognorm[row];
    res += m->emis_param[(size_t)row * m->a + x[pos]];
  }
  return res;
}
/* ACDE:637-659 addToStatistic */
static inline void emis_add(const jo_model *m, double *emis_stat, int state, const uint8_t *x, long pos, double w) {
  int e = m->state_emis[state];
  if (m->emis_order[e] < 0) return;
  int cond = cond_index(m, e, x, pos);
  if (cond >= 0) emis_stat[(size_t)(m->emis_cond_ptr[e] + cond) * m->a + x[pos]] += w;
}
/* BHOT:453-455 + :1170-1172 addToStatistic: goes to transitions[transIndex[element]] */
static inline void trans_add(const jo_model *m, double *trans_stat, long layer, int ctx, int child, double w) {
  int e = m->trans_index[elem_of(m, layer, ctx)];
  trans_stat[m->elem_child_ptr[e] + child] += w;
}

/* ------------------------------------------------------------------------------------------- */
typedef struct jo_work {
  long cap_len;       /* matrices hold cap_len+1 layers */
  double *fwd, *bwd;  /* [(cap_len+1) * max_ctx] */
  double *log_em;     /* [S] */
  double *f_int[2];   /* [max_ctx * (max_in+1)] forwardIntermediate, HOH:197 */
  int *n_sum[2];      /* [max_ctx] numberOfSummands */
  double *b_int;      /* backwardIntermediate */
} jo_work;

static void work_init(jo_work *w) { memset(w, 0, sizeof(*w)); w->cap_len = -1; }
static void work_free(jo_work *w) {
  free(w->fwd); free(w->bwd); free(w->log_em); free(w->f_int[0]); free(w->f_int[1]);
  free(w->n_sum[0]); free(w->n_sum[1]); free(w->b_int);
  work_init(w);
}
static void work_reserve(jo_work *w, const jo_model *m, long len) {
  if (!w->log_em) {
    w->log_em = malloc(sizeof(double) * (m->S + 1));
    for (int h = 0; h < 2; h++) {
      w->f_int[h] = malloc(sizeof(double) * (size_t)m->max_ctx * (m->max_in + 1));
      w->n_sum[h] = malloc(sizeof(int) * m->max_ctx);
    }
    w->b_int = malloc(sizeof(double) * (m->max_children + m->S + 2));
  }
  if (len > w->cap_len) {
    free(w->fwd); free(w->bwd);
    w->fwd = malloc(sizeof(double) * (size_t)(len + 1) * m->max_ctx);
    w->bwd = malloc(sizeof(double) * (size_t)(len + 1) * m->max_ctx);
    w->cap_len = len;
  }
}
#define FWD(w, m, l, c) ((w)->fwd[(size_t)(l) * (m)->max_ctx + (c)])
#define BWD(w, m, l, c) ((w)->bwd[(size_t)(l) * (m)->max_ctx + (c)])
#define FINT(w, m, h, c) ((w)->f_int[h] + (size_t)(c) * ((m)->max_in + 1))

/* HOH:331-340 getAllowedContext(pos, start, allowedStatesGroup, maxOrder) as a membership test: layer pos-start, the
 * context's last state emitted position pos-1.  grp: the groups of the sequence x (indexed by position), or NULL. */
static inline const int32_t *groups_of(const jo_model *m, const uint8_t *x) {
  return m->groups ? m->groups + (x - m->groups_codes) : NULL;
}
static inline int ctx_allowed(const jo_model *m, const int32_t *grp, long pos, long start, int ctx) {
  if (!grp || pos - start < m->m || grp[pos - 1] == -1) return 1; /* defContext[min(maxOrder, pos-start)]: every context */
  return m->allowed[(size_t)grp[pos - 1] * m->max_ctx + ctx];     /* preComputedContext[group] */
}

static void fill_log_emission(jo_work *w, const jo_model *m, const uint8_t *x, long pos) { /* HOH:342-346 */
  for (int s = 0; s < m->S; s++) w->log_em[s] = emis_score(m, s, x, pos);
}

/* HOH:359-434 fillFwdMatrix (all filters accept; allowedStatesGroup = the groups attached for x, if any) */
static void fill_fwd(jo_work *w, const jo_model *m, const uint8_t *x, long start_pos, long end_pos) {
  long len = end_pos - start_pos + 1, l = 0, pos = start_pos;
  int c[3];
  const int32_t *grp = groups_of(m, x);
  work_reserve(w, m, len);
  for (size_t i = 0; i < (size_t)(len + 1) * m->max_ctx; i++) w->fwd[i] = NEG_INF; /* AHMM:961-963 */
  memset(w->n_sum[0], 0, sizeof(int) * m->max_ctx);
  w->n_sum[0][0] = 1;
  FINT(w, m, 0, 0)[0] = 0;
  while (pos <= end_pos) {
    fill_log_emission(w, m, x, pos);
    int h = (int)(l % 2);
    memset(w->n_sum[1 - h], 0, sizeof(int) * m->max_ctx);
    int nc = n_ctx(m, l);
    for (int ctx = 0; ctx < nc; ctx++) {
      if (!ctx_allowed(m, grp, pos, start_pos, ctx)) continue; /* allContext = getAllowedContext(startPos, s, ...), HOH:386-388 */
      int n = n_children(m, l, ctx);
      if (w->n_sum[h][ctx] > 0) {
        FWD(w, m, l, ctx) = log_sum(0, w->n_sum[h][ctx], FINT(w, m, h, ctx));
        for (int ch = 0; ch < n; ch++) {
          trans_info(m, l, ctx, ch, c);
          int hh = (h + c[2]) % 2;
          FINT(w, m, hh, c[1])[w->n_sum[hh][c[1]]] = FWD(w, m, l, ctx) + w->log_em[c[0]] + trans_score(m, l, ctx, ch);
          w->n_sum[hh][c[1]]++;
        }
      } else {
        FWD(w, m, l, ctx) = NEG_INF;
      }
    }
    l++;
    pos++;
  }
  /* final summing and silent states, HOH:407-432 */
  int h = (int)(l % 2);
  int nc = n_ctx(m, l);
  for (int ctx = 0; ctx < nc; ctx++) {
    if (!ctx_allowed(m, grp, pos, start_pos, ctx)) continue; /* HOH:409 */
    int n = n_children(m, l, ctx);
    if (w->n_sum[h][ctx] > 0) {
      FWD(w, m, l, ctx) = log_sum(0, w->n_sum[h][ctx], FINT(w, m, h, ctx));
      for (int ch = 0; ch < n; ch++) {
        trans_info(m, l, ctx, ch, c);
        if (m->state_silent[c[0]]) {
          FINT(w, m, h, c[1])[w->n_sum[h][c[1]]++] = FWD(w, m, l, ctx) + trans_score(m, l, ctx, ch);
        }
      }
    } else {
      FWD(w, m, l, ctx) = NEG_INF;
    }
  }
}

/* HOH:1046-1058 computeLogScoreFromForward */
static double log_score_from_forward(const jo_work *w, const jo_model *m, long l) {
  double res = NEG_INF;
  int nc = n_ctx(m, l);
  if (m->m > 0) {
    for (int i = 0; i < nc; i++)
      if (m->state_final[last_state(m, l, i)]) res = log_sum2(res, FWD(w, m, l, i));
  } else {
    res = log_sum(0, nc, &FWD(w, m, l, 0));
  }
  return res;
}

#define TYPE_LIKELIHOOD 0
#define TYPE_VITERBI 1

/* HOH:471-571 fillBwdOrViterbiMatrix */
static void fill_bwd_or_viterbi(jo_work *w, const jo_model *m, int type, const uint8_t *x, long start_pos,
                                long end_pos, double weight, int add, double *trans_stat, double *emis_stat) {
  int zero = m->m == 0;
  long l = end_pos - start_pos + 1;
  int c[3];
  double res, val, new_weight;
  const int32_t *grp = groups_of(m, x);
  work_reserve(w, m, l);
  for (size_t i = 0; i < (size_t)(l + 1) * m->max_ctx; i++) w->bwd[i] = NEG_INF; /* provideMatrix(1,...) */
  if (add) {
    fill_fwd(w, m, x, start_pos, end_pos);
    res = log_score_from_forward(w, m, l);
  } else {
    res = NAN;
  }
  /* init, HOH:485-527 */
  int nc = n_ctx(m, l);
  for (int i = 0; i < m->max_ctx; i++) BWD(w, m, l, i) = zero ? 0 : NEG_INF;
  for (int ctx = nc - 1; ctx >= 0; ctx--) {
    if (!ctx_allowed(m, grp, end_pos + 1, start_pos, ctx)) continue; /* getAllowedContext(endPos+1, startPos, ...), HOH:493 */
    int n = n_children(m, l, ctx), ns = 0;
    if (zero || m->state_final[last_state(m, l, ctx)]) val = 0; else val = NEG_INF;
    for (int ch = 0; ch < n; ch++) {
      trans_info(m, l, ctx, ch, c);
      if (m->state_silent[c[0]]) {
        w->b_int[ns] = BWD(w, m, l, c[1]) + trans_score(m, l, ctx, ch);
        if (add) {
          new_weight = weight * exp(FWD(w, m, l, ctx) + w->b_int[ns] - res);
          trans_add(m, trans_stat, l, ctx, ch, new_weight);
        }
        ns++;
      }
    }
    if (ns == 0) {
      BWD(w, m, l, ctx) = val;
    } else {
      BWD(w, m, l, ctx) = type == TYPE_VITERBI ? fmax(val, tb_max(0, ns, w->b_int))
                                               : log_sum2(val, log_sum(0, ns, w->b_int));
    }
  }
  /* all positions backward, HOH:530-570 */
  long pos = end_pos;
  while (--l >= 0) {
    fill_log_emission(w, m, x, pos);
    nc = n_ctx(m, l);
    for (int ctx = nc - 1; ctx >= 0; ctx--) {
      if (!ctx_allowed(m, grp, pos, start_pos, ctx)) continue; /* getAllowedContext(endPos, startPos, ...), HOH:537 */
      int n = n_children(m, l, ctx);
      if (n > 0) {
        for (int ch = 0; ch < n; ch++) {
          trans_info(m, l, ctx, ch, c);
          w->b_int[ch] = BWD(w, m, l + c[2], c[1]) + w->log_em[c[0]] + trans_score(m, l, ctx, ch);
          if (add) {
            new_weight = weight * exp(FWD(w, m, l, ctx) + w->b_int[ch] - res);
            emis_add(m, emis_stat, c[0], x, pos, new_weight);
            trans_add(m, trans_stat, l, ctx, ch, new_weight);
          }
        }
        BWD(w, m, l, ctx) = type == TYPE_VITERBI ? tb_max(0, n, w->b_int) : log_sum(0, n, w->b_int);
      }
    }
    pos--;
    /* HOH:567 Arrays.fill(bwdMatrix[l-1], -inf) — already -inf from provideMatrix */
  }
}

/* HOH:617-708 viterbi. path may be NULL (Viterbi training: adds `weight` to the statistics).
 * Returns bwd[0][0]; *path_len receives the number of emitted states, or -1 when no child could be
 * chosen (Java would fail with an index error on childIdx == -1000). */
static double viterbi(jo_work *w, const jo_model *m, const uint8_t *x, long start_pos, long end_pos, double weight,
                      int32_t *path, long path_cap, long *path_len, double *trans_stat, double *emis_stat) {
  fill_bwd_or_viterbi(w, m, TYPE_VITERBI, x, start_pos, end_pos, 0, 0, NULL, NULL);
  long l = end_pos - start_pos + 1, layer = 0, np = 0, pos = start_pos;
  int ctx = 0, c[3];
  double current, dist, best;
  int bad = 0;
  while (layer < l) {
    int n = n_children(m, layer, ctx);
    best = INFINITY;
    int child = -1000, state = -1000, new_ctx = -1000, adv = -1000;
    for (int ch = 0; ch < n; ch++) {
      trans_info(m, layer, ctx, ch, c);
      current = BWD(w, m, layer + c[2], c[1]) + emis_score(m, c[0], x, pos) + trans_score(m, layer, ctx, ch);
      dist = current - BWD(w, m, layer, ctx);
      dist *= dist;
      if (dist < best) { child = ch; state = c[0]; new_ctx = c[1]; adv = c[2]; best = dist; }
    }
    if (child < 0) { bad = 1; break; }
    if (!path) {
      trans_add(m, trans_stat, layer, ctx, child, weight);
      emis_add(m, emis_stat, state, x, pos, weight);
    } else if (np < path_cap) {
      path[np] = state;
    }
    np++;
    pos += adv;
    layer += adv;
    ctx = new_ctx;
  }
  /* silent states at the end, HOH:667-705 */
  while (!bad) {
    int n = n_children(m, layer, ctx);
    int ls = last_state(m, layer, ctx);
    dist = (ls >= 0 && m->state_final[ls]) ? 0 - BWD(w, m, layer, ctx) : NEG_INF;
    best = dist * dist;
    int child = -1000, state = -1000, new_ctx = -1000;
    for (int ch = 0; ch < n; ch++) {
      trans_info(m, layer, ctx, ch, c);
      if (c[2] == 0) {
        current = BWD(w, m, layer, c[1]) + emis_score(m, c[0], x, pos) + trans_score(m, layer, ctx, ch);
        dist = current - BWD(w, m, layer, ctx);
        dist *= dist;
        if (dist < best) { child = ch; state = c[0]; new_ctx = c[1]; best = dist; }
      }
    }
    if (state >= 0) {
      if (!path) {
        trans_add(m, trans_stat, layer, ctx, child, weight);
        emis_add(m, emis_stat, state, x, pos, weight);
      } else if (np < path_cap) {
        path[np] = state;
      }
      np++;
      ctx = new_ctx;
    } else {
      break;
    }
  }
  if (path_len) *path_len = bad ? -1 : np;
  return BWD(w, m, 0, 0);
}

/* ------------------------------------------------------------------------------------------- */
/* model life cycle                                                                              */
jo_model *jo_model_create(const jh_model_desc *d) {
  jo_model *m = calloc(1, sizeof(*m));
  m->S = d->n_states; m->a = d->alphabet; m->m = d->max_order;
  m->lc_ctx_ptr = dup_mem(d->lc_ctx_ptr, sizeof(int) * (m->m + 2));
  m->n_ctx_total = m->lc_ctx_ptr[m->m + 1];
  m->max_ctx = 0;
  for (int lc = 0; lc <= m->m; lc++) {
    int n = m->lc_ctx_ptr[lc + 1] - m->lc_ctx_ptr[lc];
    if (n > m->max_ctx) m->max_ctx = n;
  }
  m->ctx_elem = dup_mem(d->ctx_elem, sizeof(int) * m->n_ctx_total);
  m->ctx_last_state = dup_mem(d->ctx_last_state, sizeof(int) * m->n_ctx_total);
  m->n_elem = d->n_elem;
  m->elem_child_ptr = dup_mem(d->elem_child_ptr, sizeof(int) * (m->n_elem + 1));
  m->n_child_total = m->elem_child_ptr[m->n_elem];
  m->child_state = dup_mem(d->child_state, sizeof(int) * m->n_child_total);
  m->child_desc = dup_mem(d->child_desc, sizeof(int) * m->n_child_total);
  m->elem_ctx = dup_mem(d->elem_ctx, sizeof(int) * (size_t)(m->m + 1) * m->n_elem);
  m->trans_index = dup_mem(d->trans_index, sizeof(int) * m->n_elem);
  m->trans_param = dup_mem(d->trans_param, sizeof(double) * m->n_child_total);
  m->trans_lognorm = dup_mem(d->trans_lognorm, sizeof(double) * m->n_elem);
  m->trans_hyper = dup_mem(d->trans_hyper, sizeof(double) * m->n_child_total);
  m->n_emis = d->n_emis;
  m->state_emis = dup_mem(d->state_emis, sizeof(int) * m->S);
  m->emis_order = dup_mem(d->emis_order, sizeof(int) * m->n_emis);
  m->emis_cond_ptr = dup_mem(d->emis_cond_ptr, sizeof(int) * (m->n_emis + 1));
  m->n_cond_total = m->emis_cond_ptr[m->n_emis];
  m->emis_param = dup_mem(d->emis_param, sizeof(double) * (size_t)m->n_cond_total * m->a);
  m->emis_lognorm = dup_mem(d->emis_lognorm, sizeof(double) * m->n_cond_total);
  m->emis_hyper = dup_mem(d->emis_hyper, sizeof(double) * (size_t)m->n_cond_total * m->a);
  m->state_silent = dup_mem(d->state_silent, m->S);
  m->state_final = dup_mem(d->state_final, m->S);
  m->log_uniform = -log((double)m->a);
  m->max_children = 0;
  for (int e = 0; e < m->n_elem; e++) {
    int n = m->elem_child_ptr[e + 1] - m->elem_child_ptr[e];
    if (n > m->max_children) m->max_children = n;
  }
  /* bound for forwardIntermediate's last dimension (HOH:197 uses getMaximalInDegree()+1) */
  int *indeg = calloc((size_t)(m->m + 1) * m->max_ctx + 1, sizeof(int));
  m->max_in = 1;
  for (int lc = 0; lc <= m->m; lc++) {
    int nc = m->lc_ctx_ptr[lc + 1] - m->lc_ctx_ptr[lc], c[3];
    for (int ctx = 0; ctx < nc; ctx++) {
      int n = n_children(m, lc, ctx);
      for (int ch = 0; ch < n; ch++) {
        trans_info(m, lc, ctx, ch, c);
        if (c[1] < 0) continue;
        int t = lc_of(m, lc + c[2]) * m->max_ctx + c[1];
        if (++indeg[t] > m->max_in) m->max_in = indeg[t];
      }
    }
  }
  free(indeg);
  m->max_in += 1;
  return m;
}

/* AHMM:173-228 fillPreComputedContext: a context of layer class m is allowed for group g if some context of that class has a
 * child whose state belongs to g and which leads to it.  group_ptr[n_groups+1], group_states: the states of every group. */
void jo_model_set_groups(jo_model *m, int n_groups, const int32_t *group_ptr, const int32_t *group_states) {
  free(m->allowed);
  m->allowed = NULL;
  m->n_groups = n_groups;
  if (n_groups <= 0) return;
  m->allowed = calloc((size_t)n_groups * m->max_ctx, 1);
  int nc = n_ctx(m, m->m), c[3];
  uint8_t *in_group = malloc(m->S);
  for (int g = 0; g < n_groups; g++) {
    memset(in_group, 0, m->S);
    for (int k = group_ptr[g]; k < group_ptr[g + 1]; k++) in_group[group_states[k]] = 1;
    for (int ctx = 0; ctx < nc; ctx++) {
      int n = n_children(m, m->m, ctx);
      for (int ch = 0; ch < n; ch++) {
        trans_info(m, m->m, ctx, ch, c);
        if (in_group[c[0]]) m->allowed[(size_t)g * m->max_ctx + c[1]] = 1;
      }
    }
  }
  free(in_group);
}
/* The allowedStatesGroup arrays of the sequences (HOH:812-820: one int per position, -1 = unconstrained), aligned with the
 * code buffer of the following calls; NULL detaches (every call then behaves as with allowedStatesGroup == null). */
void jo_model_attach_groups(jo_model *m, const uint8_t *codes, const int32_t *groups) {
  m->groups_codes = codes;
  m->groups = groups;
}

void jo_model_destroy(jo_model *m) {
  if (!m) return;
  free(m->allowed);
  free(m->lc_ctx_ptr); free(m->ctx_elem); free(m->ctx_last_state); free(m->elem_child_ptr);
  free(m->child_state); free(m->child_desc); free(m->elem_ctx); free(m->trans_index);
  free(m->trans_param); free(m->trans_lognorm); free(m->trans_hyper); free(m->state_emis);
  free(m->emis_order); free(m->emis_cond_ptr); free(m->emis_param); free(m->emis_lognorm);
  free(m->emis_hyper); free(m->state_silent); free(m->state_final);
  free(m);
}

void jo_model_get_params(const jo_model *m, double *tp, double *tl, double *ep, double *el) {
  if (tp) memcpy(tp, m->trans_param, sizeof(double) * m->n_child_total);
  if (tl) memcpy(tl, m->trans_lognorm, sizeof(double) * m->n_elem);
  if (ep) memcpy(ep, m->emis_param, sizeof(double) * (size_t)m->n_cond_total * m->a);
  if (el) memcpy(el, m->emis_lognorm, sizeof(double) * m->n_cond_total);
}

/* precompute(): TransitionElement.java:119-126 / BHOT:935-941 and ACDE:466-480 */
void jo_model_set_params(jo_model *m, const double *tp, const double *tl, const double *ep, const double *el) {
  if (tp) memcpy(m->trans_param, tp, sizeof(double) * m->n_child_total);
  if (ep) memcpy(m->emis_param, ep, sizeof(double) * (size_t)m->n_cond_total * m->a);
  if (tl) memcpy(m->trans_lognorm, tl, sizeof(double) * m->n_elem);
  else if (tp)
    for (int e = 0; e < m->n_elem; e++)
      m->trans_lognorm[e] = log_sum_normalisation(m->trans_param + m->elem_child_ptr[e],
                                                  m->elem_child_ptr[e + 1] - m->elem_child_ptr[e]);
  if (el) memcpy(m->emis_lognorm, el, sizeof(double) * m->n_cond_total);
  else if (ep)
    for (int r = 0; r < m->n_cond_total; r++)
      m->emis_lognorm[r] = log_sum_normalisation(m->emis_param + (size_t)r * m->a, m->a);
}

/* HOH:253-259 getLogPriorTerm = BHOT:501-509 (+:1128-1143) + sum ACDE:368-379 */
double jo_log_prior(const jo_model *m) {
  double pt = 0;
  for (int t = 0; t < m->n_elem; t++) {
    if (m->trans_index[t] != t) continue;
    int p0 = m->elem_child_ptr[t], n = m->elem_child_ptr[t + 1] - p0;
    double term = 0;
    if (n > 1) {
      double r = 0, sum_hyper = 0;
      for (int d = 0; d < n; d++) {
        sum_hyper += m->trans_hyper[p0 + d];
        r += m->trans_hyper[p0 + d] * m->trans_param[p0 + d];
      }
      term = sum_hyper == 0 ? 0 : r - sum_hyper * m->trans_lognorm[t];
    }
    pt += term;
  }
  double res = pt;
  for (int e = 0; e < m->n_emis; e++) {
    if (m->emis_order[e] < 0) continue; /* SilentEmission.getLogPriorTerm() == 0 */
    double r = 0;
    for (int row = m->emis_cond_ptr[e]; row < m->emis_cond_ptr[e + 1]; row++) {
      double ess = 0; /* ACDE:289-291 ess[i] = sum of hyper-parameters */
      for (int j = 0; j < m->a; j++) ess += m->emis_hyper[(size_t)row * m->a + j];
      if (ess > 0) {
        r += -ess * m->emis_lognorm[row];
        for (int j = 0; j < m->a; j++) r += m->emis_hyper[(size_t)row * m->a + j] * m->emis_param[(size_t)row * m->a + j];
      }
    }
    res += r;
  }
  return res;
}

/* M-step: BHOT:458-466 + :1242-1262 (transitions), ACDE:682-699 (emissions).
 * The statistics arrays are modified in place (hyper-parameters are added) like in the reference. */
void jo_mstep(jo_model *m, double *trans_stat, double *emis_stat) {
  for (int t = 0; t < m->n_elem; t++) {
    int p0 = m->elem_child_ptr[t], n = m->elem_child_ptr[t + 1] - p0;
    if (m->trans_index[t] == t) {
      if (n > 0) {
        double ln = 0;
        for (int i = 0; i < n; i++) {
          trans_stat[p0 + i] += m->trans_hyper[p0 + i];
          ln += trans_stat[p0 + i];
          m->trans_param[p0 + i] = log(trans_stat[p0 + i]);
        }
        if (ln == 0) {
          for (int i = 0; i < n; i++) m->trans_param[p0 + i] = 0;
          ln = n;
        }
        ln = log(ln);
        for (int i = 0; i < n; i++) m->trans_param[p0 + i] -= ln;
        m->trans_lognorm[t] = log_sum(0, n, m->trans_param + p0);
      }
    } else { /* setParameters(transitions[transIndex[t]]) + precompute, BHOT:1400-1406 */
      int q0 = m->elem_child_ptr[m->trans_index[t]];
      for (int i = 0; i < n; i++) m->trans_param[p0 + i] = m->trans_param[q0 + i];
      m->trans_lognorm[t] = log_sum_normalisation(m->trans_param + p0, n);
    }
  }
  for (int e = 0; e < m->n_emis; e++) {
    if (m->emis_order[e] < 0) continue;
    for (int row = m->emis_cond_ptr[e]; row < m->emis_cond_ptr[e + 1]; row++) {
      double *st = emis_stat + (size_t)row * m->a, *hy = m->emis_hyper + (size_t)row * m->a;
      double sum = 0;
      for (int i = 0; i < m->a; i++) { st[i] += hy[i]; sum += st[i]; }
      if (sum == 0) {
        for (int i = 0; i < m->a; i++) st[i] = 1;
        sum = m->a;
      }
      for (int i = 0; i < m->a; i++) m->emis_param[(size_t)row * m->a + i] = log(st[i] / sum);
      m->emis_lognorm[row] = 0;
    }
  }
}

/* ------------------------------------------------------------------------------------------- */
/* per data set entry points                                                                     */

/* AHMM:1049-1070 logProb for every sequence == HOH:927-941 */
int jo_loglik(const jo_model *m, const uint8_t *codes, const int64_t *off, int64_t n, double *out) {
  jo_work w; work_init(&w);
  for (int64_t i = 0; i < n; i++) {
    long L = (long)(off[i + 1] - off[i]);
    if (L <= 0) { work_free(&w); return -1; }
    fill_bwd_or_viterbi(&w, m, TYPE_LIKELIHOOD, codes + off[i], 0, L - 1, 1, 0, NULL, NULL);
    out[i] = BWD(&w, m, 0, 0);
  }
  work_free(&w);
  return 0;
}

/* AHMM:985-998 getLogProbFor(seq,start,end): a window of one sequence, conditions use absolute positions */
double jo_loglik_window(const jo_model *m, const uint8_t *x, int64_t start, int64_t end) {
  jo_work w; work_init(&w);
  fill_bwd_or_viterbi(&w, m, TYPE_LIKELIHOOD, x, (long)start, (long)end, 1, 0, NULL, NULL);
  double r = BWD(&w, m, 0, 0);
  work_free(&w);
  return r;
}

/* log-likelihood through the forward matrix, the "sanity check" of AHMM:1053-1065 */
double jo_loglik_forward(const jo_model *m, const uint8_t *x, int64_t L) {
  jo_work w; work_init(&w);
  fill_fwd(&w, m, x, 0, (long)L - 1);
  double r = log_score_from_forward(&w, m, (long)L);
  work_free(&w);
  return r;
}

/* full matrices for tests: fwd/bwd [(L+1)*max_ctx], row l holds n_ctx(l) valid entries */
int jo_matrices(const jo_model *m, const uint8_t *x, int64_t L, int type, double *fwd, double *bwd) {
  jo_work w; work_init(&w);
  fill_fwd(&w, m, x, 0, (long)L - 1);
  fill_bwd_or_viterbi(&w, m, type, x, 0, (long)L - 1, 1, 0, NULL, NULL);
  if (fwd) memcpy(fwd, w.fwd, sizeof(double) * (size_t)(L + 1) * m->max_ctx);
  if (bwd) memcpy(bwd, w.bwd, sizeof(double) * (size_t)(L + 1) * m->max_ctx);
  work_free(&w);
  return m->max_ctx;
}

/* AHMM:894-900 getViterbiPathsFor. paths is CSR by path_ptr (capacity per sequence), -1 padded. */
int jo_viterbi(const jo_model *m, const uint8_t *codes, const int64_t *off, int64_t n, int32_t *paths,
               const int64_t *path_ptr, int64_t *path_len, double *scores) {
  jo_work w; work_init(&w);
  for (int64_t i = 0; i < n; i++) {
    long L = (long)(off[i + 1] - off[i]), np = 0;
    if (L <= 0) { work_free(&w); return -1; }
    long cap = (long)(path_ptr[i + 1] - path_ptr[i]);
    for (long k = 0; k < cap; k++) paths[path_ptr[i] + k] = -1;
    scores[i] = viterbi(&w, m, codes + off[i], 0, L - 1, 0, paths + path_ptr[i], cap, &np, NULL, NULL);
    if (path_len) path_len[i] = np;
  }
  work_free(&w);
  return 0;
}

/* HOH:590-598 getViterbiPathFor(startPos, endPos, seq): emissions see the residues before startPos (absolute positions,
 * HigherOrderDiscreteEmission.java:138-146). */
double jo_viterbi_window(const jo_model *m, const uint8_t *x, int64_t start, int64_t end, int32_t *path, int64_t cap, int64_t *path_len) {
  jo_work w; work_init(&w);
  long np = 0;
  for (long k = 0; k < cap; k++) path[k] = -1;
  double sc = viterbi(&w, m, x, (long)start, (long)end, 0, path, (long)cap, &np, NULL, NULL);
  if (path_len) *path_len = np;
  work_free(&w);
  return sc;
}

/* HOH:294-329 fillLogStatePosteriorMatrix(silentZero=true) + AHMM:790-797 (first column dropped) for positions
 * start..end of x (AHMM:776-797 getLogStatePosteriorMatrixFor(startPos, endPos, seq)).  out[S][L] row-major, L = end-start+1. */
int jo_posterior_window(const jo_model *m, const uint8_t *x, int64_t start, int64_t end, double *out) {
  jo_work w; work_init(&w);
  const int64_t L = end - start + 1;
  if (m->m == 0) { /* HOH:296-308: every position stands alone; statePosterior comes zero-initialised (AHMM:745-757) */
    int c[3];
    double *le = calloc((size_t)m->S, sizeof(double));
    for (long l = 0; l < L; l++) {
      for (int s = 0; s < m->S; s++) out[(size_t)s * L + l] = 0.0;
      for (int s = 0; s < m->S; s++) {
        trans_info(m, 0, 0, s, c);
        le[s] = out[(size_t)c[0] * L + l] = trans_score(m, 0, 0, s) + emis_score(m, c[0], x, (long)start + l);
      }
      double d = log_sum(0, m->S, le);
      for (int s = 0; s < m->S; s++) out[(size_t)s * L + l] -= d;
    }
    free(le);
    work_free(&w);
    return 0;
  }
  fill_fwd(&w, m, x, (long)start, (long)end);
  fill_bwd_or_viterbi(&w, m, TYPE_LIKELIHOOD, x, (long)start, (long)end, 1, 0, NULL, NULL);
  double log_prob = BWD(&w, m, 0, 0);
  double *col = malloc(sizeof(double) * m->S);
  for (long l = 0; l <= L; l++) {
    for (int s = 0; s < m->S; s++) col[s] = NEG_INF;
    int nc = n_ctx(m, l);
    for (int c = 0; c < nc; c++) {
      int st = last_state(m, l, c);
      if (st >= 0 && !m->state_silent[st]) col[st] = log_sum2(col[st], FWD(&w, m, l, c) + BWD(&w, m, l, c));
    }
    if (l >= 1)
      for (int s = 0; s < m->S; s++) out[(size_t)s * L + (l - 1)] = col[s] - log_prob;
  }
  free(col);
  work_free(&w);
  return 0;
}

int jo_posterior(const jo_model *m, const uint8_t *x, int64_t L, double *out) { return jo_posterior_window(m, x, 0, L - 1, out); }

/* AHMM:1125-1139 decodeStatePosterior: strict '<' => lowest index wins ties */
void jo_decode_posterior(const double *post, int S, int64_t L, int32_t *path) {
  for (int64_t l = 0; l < L; l++) {
    int best = 0;
    for (int j = 1; j < S; j++)
      if (post[(size_t)best * L + l] < post[(size_t)j * L + l]) best = j;
    path[l] = best;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* E-step: HOH:791-855 doOneStep over [start,end) with the static partition of HOH:1241-1248      */
typedef struct estep_job {
  const jo_model *m;
  const uint8_t *codes;
  const int64_t *off;
  const double *weights;
  int64_t start, end;
  int mode;
  double *trans_stat, *emis_stat;
  double score;
  int err;
} estep_job;

static void *estep_worker(void *arg) {
  estep_job *j = arg;
  const jo_model *m = j->m;
  jo_work w; work_init(&w);
  double new_value = 0, weight = 1, score;
  for (int64_t n = j->start; n < j->end; n++) {
    long L = (long)(j->off[n + 1] - j->off[n]);
    if (L <= 0) { j->err = 1; break; }
    if (j->weights) weight = j->weights[n];
    if (j->mode == JH_MODE_VITERBI) {
      long np;
      score = viterbi(&w, m, j->codes + j->off[n], 0, L - 1, weight, NULL, 0, &np, j->trans_stat, j->emis_stat);
    } else { /* HOH:723-726 baumWelch */
      fill_bwd_or_viterbi(&w, m, TYPE_LIKELIHOOD, j->codes + j->off[n], 0, L - 1, weight, 1, j->trans_stat, j->emis_stat);
      score = BWD(&w, m, 0, 0);
    }
    new_value += weight * score;
  }
  j->score = new_value;
  work_free(&w);
  return NULL;
}

/* Compute.oneIteration (HOH:1250-1261) with `threads` workers, followed by joinStatistics
 * (HOH:1263-1279, BHOT:1181-1194, ACDE:618-635): worker 0 += worker i in index order.
 * trans_stat[n_child_total], emis_stat[n_cond_total*a] receive the joined statistics. */
int jo_estep(const jo_model *m, const uint8_t *codes, const int64_t *off, int64_t n, const double *weights, int mode,
             int threads, double *trans_stat, double *emis_stat, double *score) {
  if (threads < 1) threads = 1;
  size_t nt = (size_t)m->n_child_total, ne = (size_t)m->n_cond_total * m->a;
  estep_job *jobs = calloc(threads, sizeof(estep_job));
  pthread_t *tid = calloc(threads, sizeof(pthread_t));
  int64_t last = 0;
  for (int i = 0; i < threads; i++) {
    jobs[i].m = m; jobs[i].codes = codes; jobs[i].off = off; jobs[i].weights = weights; jobs[i].mode = mode;
    jobs[i].start = last;
    jobs[i].end = i == threads - 1 ? n : (int64_t)(i + 1) * n / threads; /* HOH:1241-1248 */
    last = jobs[i].end;
    if (i == 0) { jobs[i].trans_stat = trans_stat; jobs[i].emis_stat = emis_stat; }
    else { jobs[i].trans_stat = malloc(sizeof(double) * (nt + 1)); jobs[i].emis_stat = malloc(sizeof(double) * (ne + 1)); }
    memset(jobs[i].trans_stat, 0, sizeof(double) * nt); /* resetStatistics, HOH:890-895 */
    memset(jobs[i].emis_stat, 0, sizeof(double) * ne);
  }
  for (int i = 1; i < threads; i++) pthread_create(&tid[i], NULL, estep_worker, &jobs[i]);
  estep_worker(&jobs[0]);
  for (int i = 1; i < threads; i++) pthread_join(tid[i], NULL);
  double res = 0;
  int err = 0;
  for (int i = 0; i < threads; i++) { res += jobs[i].score; err |= jobs[i].err; }
  for (int i = 1; i < threads; i++) {
    for (size_t k = 0; k < nt; k++) trans_stat[k] += jobs[i].trans_stat[k];
    for (size_t k = 0; k < ne; k++) emis_stat[k] += jobs[i].emis_stat[k];
    free(jobs[i].trans_stat); free(jobs[i].emis_stat);
  }
  *score = res;
  free(jobs); free(tid);
  return err ? -1 : 0;
}

/* HOH:733-789 train, one start, skipInit == true:
 *   do { new = prior + oneIteration(); it++; if tc(it, old, new) estimateFromStatistics() else break; }
 * term_kind/max_iter/eps as in jh_train_opts. Returns the number of E-steps; objective[cap]. */
int jo_train(jo_model *m, const uint8_t *codes, const int64_t *off, int64_t n, const double *weights, int mode,
             int term_kind, int max_iter, double eps, int threads, double *objective, int cap) {
  size_t nt = (size_t)m->n_child_total, ne = (size_t)m->n_cond_total * m->a;
  double *ts = malloc(sizeof(double) * (nt + 1)), *es = malloc(sizeof(double) * (ne + 1));
  double old_value, new_value = NEG_INF, score = 0;
  int it = 0;
  for (;;) {
    old_value = new_value;
    if (jo_estep(m, codes, off, n, weights, mode, threads, ts, es, &score) != 0) { it = -1; break; }
    new_value = jo_log_prior(m) + score;
    if (it < cap) objective[it] = new_value;
    it++;
    int go_on;
    if (term_kind == JH_TERM_ITERATIONS) go_on = it < max_iter;
    else go_on = (eps <= fabs(old_value - new_value)) && (max_iter <= 0 || it < max_iter);
    if (go_on) jo_mstep(m, ts, es); else break;
  }
  free(ts); free(es);
  return it;
}

/* brute force over all state paths (tests only, tiny models without silent states, m == 1):
 * HOH:265-292 getLogProbForPath for every path; returns log-sum and the best path score. */
double jo_path_logprob(const jo_model *m, const uint8_t *x, int64_t L, const int32_t *path, int64_t path_len) {
  double res = 0;
  long l = 0, layer = 0, pos = 0;
  int ctx = 0, c[3];
  (void)L;
  if (!m->state_final[path[path_len - 1]]) return NAN;
  while (l < path_len) {
    int state = path[l], n = n_children(m, layer, ctx), child = -1;
    int e = elem_of(m, layer, ctx);
    for (int i = 0; i < n; i++)
      if (m->child_state[m->elem_child_ptr[e] + i] == state) { child = i; break; } /* BHOT:637-643 getChildIdx */
    if (child < 0) return NEG_INF; /* "Impossible path" */
    res += trans_score(m, layer, ctx, child) + emis_score(m, state, x, pos);
    trans_info(m, layer, ctx, child, c);
    ctx = c[1];
    if (c[2] == 1) { pos++; layer++; }
    l++;
  }
  return res;
}
