// kernels_exact.cuh — reference-order log-space kernels (general context graph, transition order m >= 1,
// no silent states).  One GROUP of G lanes per sequence (G = 2..32, 32/G sequences per warp), contexts striped
// over the lanes of the group, per-group DP vectors in shared memory, residues streamed with 128-bit loads.
//
// These kernels keep the reference's arithmetic order so that
//   * Viterbi paths are bit-exact (only fp64 add/compare, association (bwd + em) + trans, no FMA contraction,
//     tie rule of HigherOrderHMM.java:633-650),
//   * log-likelihoods differ from Java only through exp/log ulps (n-ary max-offset log-sum-exp in child order,
//     Normalisation.java:65-99).
// They are the production path for Viterbi and the fallback / cross-check of the scaled-space fast kernels.
#pragma once
#include "common.cuh"

namespace jhx {

// Launch condition decided on the device: the scaled kernels hand sequences they cannot finish to these kernels through a
// device-side list, and an E-step whose statistics are not finite is redone here — without a host round trip, so the EM
// loop stays stream ordered (and graph capturable).  A launch whose condition does not hold exits at once.
struct Cond {
  const int32_t *flag = nullptr;        // work only if *flag == want (nullptr: always)
  int want = 0;
  const int32_t *n_list_dev = nullptr;  // length of seq_list, read on the device (nullptr: the n_list argument)
};

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// ---- residue stream: 16-byte chunks cached in registers, every lane of a group reads the same address ----
struct ResReader {
  const uint8_t *seq;
  unsigned long long cur;
  uint4 buf;
  __device__ __forceinline__ void init(const uint8_t *s) {
    seq = s;
    cur = ~0ull;
    buf = make_uint4(0, 0, 0, 0);
  }
  __device__ __forceinline__ int get(int64_t q) {
    unsigned long long a = (unsigned long long)(seq + q);
    unsigned long long ch = a >> 4;
    if (ch != cur) {
      buf = __ldg(reinterpret_cast<const uint4 *>(ch << 4));
      cur = ch;
    }
    unsigned idx = (unsigned)(a >> 2) & 3u;
    unsigned w = idx == 0 ? buf.x : idx == 1 ? buf.y : idx == 2 ? buf.z : buf.w;
    return (int)((w >> (((unsigned)a & 3u) * 8u)) & 0xffu);
  }
};

// rolling context hash h(pos) = sum_{i=0..K} a^i x[pos-i]; the table index of an order-k emission is h mod a^(k+1)
// (cond*a + sym with cond = sum_{i<k} a^i x[pos-1-i], HigherOrderDiscreteEmission.java:138-146).
__device__ __forceinline__ int hash_at(const DevModel &M, ResReader &rd, int64_t pos) {
  int h = 0, pw = 1;
  for (int i = 0; i <= M.Kmax; i++) {
    int64_t q = pos - i;
    if (q < 0) break;
    h += pw * rd.get(q);
    pw *= M.a;
  }
  return h;
}
__device__ __forceinline__ int hash_back(const DevModel &M, ResReader &rd, int h, int64_t pos) {  // pos -> pos-1
  int64_t q = pos - 1 - M.Kmax;
  return h / M.a + (q >= 0 ? M.hpow * rd.get(q) : 0);
}
__device__ __forceinline__ int hash_fwd(const DevModel &M, ResReader &rd, int h, int64_t pos) {  // pos -> pos+1
  return (h * M.a + rd.get(pos + 1)) % M.hmod;
}
__device__ __forceinline__ int emis_index(const DevModel &M, int s, int h) {
  int mod = M.state_mod[s];
  return M.state_tab[s] + (M.a_pow2 ? (h & (mod - 1)) : (h % mod));
}

template <int G>
struct Group {
  int lane, lead, slot;
  unsigned mask;
  double *v0, *v1, *em;
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  __device__ __forceinline__ long long bcast(long long v) const { return __shfl_sync(mask, v, lead); }
  __device__ __forceinline__ double bcastd(double v) const { return __shfl_sync(mask, v, lead); }
};

template <int G>
__device__ __forceinline__ Group<G> make_group(const DevModel &M, double *group_smem) {
  Group<G> g;
  int wl = threadIdx.x & 31;
  g.lane = wl % G;
  g.lead = wl - g.lane;
  g.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << g.lead);
  int gpc = blockDim.x / G;
  int gid = threadIdx.x / G;
  g.slot = blockIdx.x * gpc + gid;
  int per = 2 * M.Cmax + M.S;
  g.v0 = group_smem + (size_t)gid * per;
  g.v1 = g.v0 + M.Cmax;
  g.em = g.v1 + M.Cmax;
  return g;
}

// HigherOrderHMM.fillLogEmission (:342-346) for position pos.
template <int G>
__device__ __forceinline__ void fill_em(const DevModel &M, const Group<G> &g, int64_t pos, int h) {
  for (int s = g.lane; s < M.S; s += G) {
    int k = M.state_order[s];
    double v;
    if (k < 0) v = 0.0;
    else if (pos < k) v = dadd(0.0, M.log_uniform);
    else v = M.escore[emis_index(M, s, h)];
    g.em[s] = v;
  }
}

__device__ __forceinline__ double edge_value(const DevModel &M, const int *cn, int p, const double *nxt, const double *em) {
  return dadd(dadd(nxt[cn[p]], em[M.child_state[p]]), M.tscore[p]);  // (bwd + emission) + transition, HOH:546-549
}

// bwd[l][ctx] = getLogSum(0,n,backwardIntermediate)  (HigherOrderHMM.java:560-562, Normalisation.java:65-99)
__device__ __forceinline__ double ctx_bwd_lse(const DevModel &M, int lc, int c, const double *nxt, const double *em, double *offset_out) {
  int e = M.ctx_elem[M.lc_ctx_ptr[lc] + c];
  int p0 = M.elem_child_ptr[e], p1 = M.elem_child_ptr[e + 1];
  *offset_out = JH_NEG_INF;
  if (p0 == p1) return JH_NEG_INF;
  const int *cn = M.child_next + (size_t)lc * M.n_child;
  double off = JH_NEG_INF;
  for (int p = p0; p < p1; p++) {
    double v = edge_value(M, cn, p, nxt, em);
    if (v > off) off = v;
  }
  if (isinf(off)) return JH_NEG_INF;
  double sum = 0;
  for (int p = p0; p < p1; p++) sum = dadd(sum, exp(dsub(edge_value(M, cn, p, nxt, em), off)));
  *offset_out = off;
  return dadd(off, log(sum));
}

// ToolBox.max (ToolBox.java:124-138) + the forward-trace choice of HigherOrderHMM.java:633-650:
// first child (child order) minimising (current - bwd[l][ctx])^2 with strict '<'.
__device__ __forceinline__ double ctx_bwd_max(const DevModel &M, int lc, int c, const double *nxt, const double *em, int *choice) {
  int e = M.ctx_elem[M.lc_ctx_ptr[lc] + c];
  int p0 = M.elem_child_ptr[e], p1 = M.elem_child_ptr[e + 1];
  *choice = 255;
  if (p0 == p1) return JH_NEG_INF;
  const int *cn = M.child_next + (size_t)lc * M.n_child;
  double mx = edge_value(M, cn, p0, nxt, em);
  for (int p = p0 + 1; p < p1; p++) {
    double v = edge_value(M, cn, p, nxt, em);
    if (mx < v) mx = v;
  }
  double best = INFINITY;
  for (int p = p0; p < p1; p++) {
    double dist = dsub(edge_value(M, cn, p, nxt, em), mx);
    dist = dmul(dist, dist);
    if (dist < best) {
      best = dist;
      *choice = p - p0;
    }
  }
  return mx;
}

__device__ __forceinline__ double lse2(double x, double y) {  // Normalisation.getLogSum(x, y)
  double off = JH_NEG_INF;
  if (x > off) off = x;
  if (y > off) off = y;
  if (isinf(off)) return JH_NEG_INF;
  double sum = dadd(dadd(0.0, exp(dsub(x, off))), exp(dsub(y, off)));
  return dadd(off, log(sum));
}

// grp / s0: allowed states groups of the sequence and the window start (layer L <-> absolute position s0 + L)
template <int G>
__device__ __forceinline__ void init_last_layer(const DevModel &M, const Group<G> &g, int64_t L, double *v, const int32_t *grp = nullptr, int64_t s0 = 0) {
  int lc = L < M.m ? (int)L : M.m;  // bwd[L][c] = finalState[lastState(c)] ? 0 : -inf  (HOH:487-527 without silent states)
  int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
  for (int c = g.lane; c < nc; c += G) {
    int ls = M.ctx_last[M.lc_ctx_ptr[lc] + c];
    v[c] = (ls >= 0 && M.is_final[ls] && ctx_allowed(M, grp, s0 + L, L, c)) ? 0.0 : JH_NEG_INF;  // only allowed contexts are filled (HOH:493-496)
  }
}

// ------------------------------------------------------------------------------------------------
// AbstractHMM.logProb (AbstractHMM.java:1049-1070): backward sweep, result bwd[0][0]; nothing is stored.
template <int G>
__global__ void __launch_bounds__(256) k_exact_loglik(DevModel M, DevData D, unsigned long long *counter, double *out,
                                                      const int32_t *seq_list, int64_t n_list) {
  extern __shared__ double smem[];
  Group<G> g = make_group<G>(M, smem);
  ResReader rd;
  const int64_t n_work = seq_list ? n_list : D.n_seq;
  for (;;) {
    long long it = 0;
    if (g.lane == 0) it = (long long)atomicAdd(counter, 1ull);
    it = g.bcast(it);
    if (it >= n_work) break;
    int sid = seq_list ? seq_list[it] : D.order[it];
    int64_t off = D.offsets[sid], L = D.offsets[sid + 1] - off;
    rd.init(D.codes + off);
    const int32_t *grp = D.groups ? D.groups + off : nullptr;
    double *nxt = g.v0, *cur = g.v1;
    init_last_layer<G>(M, g, L, nxt, grp);
    int h = hash_at(M, rd, L - 1);
    for (int64_t l = L - 1; l >= 0; --l) {
      fill_em<G>(M, g, l, h);
      g.sync();
      int lc = l < M.m ? (int)l : M.m;
      int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
      double dummy;
      for (int c = g.lane; c < nc; c += G) cur[c] = ctx_allowed(M, grp, l, l, c) ? ctx_bwd_lse(M, lc, c, nxt, g.em, &dummy) : JH_NEG_INF;
      g.sync();
      double *t = cur; cur = nxt; nxt = t;
      if (l > 0) h = hash_back(M, rd, h, l);
    }
    if (g.lane == 0) out[sid] = nxt[0];
    g.sync();
  }
}

// ------------------------------------------------------------------------------------------------
// HigherOrderHMM.viterbi (:617-708): backward max sweep + forward trace.  The trace decision of the reference
// depends only on (layer, context), so it is taken during the sweep and stored as one byte per (layer, context).
// stats != nullptr: Viterbi training (path == null in the reference): `weight` is added along the path.
// W.seq != nullptr: windows of sequences (getViterbiPathFor(startPos, endPos, seq), HOH:590-598): layer 0 is the window start,
// the emissions keep their absolute positions; outputs are indexed by the window then.  path_ptr: CSR of the output paths
// (nullptr: the data set offsets).
template <int G>
__global__ void __launch_bounds__(256) k_exact_viterbi(DevModel M, DevData D, WorkList W, unsigned long long *counter, uint8_t *choice_scratch,
                                                       int64_t scratch_rows, PathOut paths, const int64_t *path_ptr, double *scores, double *stats) {
  extern __shared__ double smem[];
  Group<G> g = make_group<G>(M, smem);
  ResReader rd;
  uint8_t *choice = choice_scratch + (size_t)g.slot * scratch_rows * M.Cmax;
  double score_sum = 0;
  for (;;) {
    long long it = 0;
    if (g.lane == 0) it = (long long)atomicAdd(counter, 1ull);
    it = g.bcast(it);
    if (it >= (W.seq ? W.n : D.n_seq)) break;
    const int sid = W.seq ? W.seq[it] : D.order[it];
    const int64_t oid = W.seq ? it : sid;
    const int64_t off = D.offsets[sid], Lseq = D.offsets[sid + 1] - off;
    const int64_t s0 = W.start ? W.start[it] : 0, s1 = W.end ? W.end[it] : Lseq - 1;
    const int64_t L = s1 - s0 + 1;
    const int64_t pbase = path_ptr ? path_ptr[oid] : off + s0;
    rd.init(D.codes + off);
    const int32_t *grp = D.groups ? D.groups + off : nullptr;
    double *nxt = g.v0, *cur = g.v1;
    init_last_layer<G>(M, g, L, nxt, grp, s0);
    int h = hash_at(M, rd, s1);
    for (int64_t l = L - 1; l >= 0; --l) {
      fill_em<G>(M, g, s0 + l, h);
      g.sync();
      int lc = l < M.m ? (int)l : M.m;
      int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
      for (int c = g.lane; c < nc; c += G) {
        int ch;
        cur[c] = ctx_bwd_max(M, lc, c, nxt, g.em, &ch);
        if (!ctx_allowed(M, grp, s0 + l, l, c)) cur[c] = JH_NEG_INF;  // bwdMatrix stays -inf outside the allowed contexts (HOH:537, 567)
        choice[(size_t)l * M.Cmax + c] = (uint8_t)ch;
      }
      g.sync();
      double *t = cur; cur = nxt; nxt = t;
      if (l > 0) h = hash_back(M, rd, h, s0 + l);
    }
    if (g.lane == 0) {
      double sc = nxt[0];
      if (scores) scores[oid] = sc;
      double w = D.weights ? D.weights[sid] : 1.0;
      score_sum = dadd(score_sum, dmul(w, sc));
      int ctx = 0;
      bool ok = true;
      int hh = 0;
      if (stats) hh = hash_at(M, rd, s0);
      for (int64_t l = 0; l < L; l++) {
        int lc = l < M.m ? (int)l : M.m;
        int ch = ok ? choice[(size_t)l * M.Cmax + ctx] : 255;
        if (ch == 255) {
          ok = false;
          if (paths) path_store(paths, pbase + l, -1);
          continue;
        }
        int p = M.elem_child_ptr[M.ctx_elem[M.lc_ctx_ptr[lc] + ctx]] + ch;
        int st = M.child_state[p];
        if (paths) path_store(paths, pbase + l, st);
        if (stats) {
          atomicAdd(&stats[M.child_stat[p]], w);
          if (M.state_order[st] >= 0 && s0 + l >= M.state_order[st]) atomicAdd(&stats[M.n_child + emis_index(M, st, hh)], w);
          if (l + 1 < L) hh = hash_fwd(M, rd, hh, s0 + l);
        }
        ctx = M.child_next[(size_t)lc * M.n_child + p];
      }
    }
    g.sync();
  }
  if (stats && g.lane == 0 && score_sum != 0) atomicAdd(&stats[M.n_child + M.n_cond * M.a], score_sum);
}

// ------------------------------------------------------------------------------------------------
// Baum-Welch E-step (HigherOrderHMM.java:471-571 with add == true), posterior decoding (:294-329 +
// AbstractHMM.java:1125-1139) and the full posterior matrix of one sequence.
//   forward  : pull form of fillFwdMatrix (:359-434); the summands of a context arrive in (source context,
//              child) order, exactly the push order of the reference, and are reduced by the same n-ary LSE.
//   res      : computeLogScoreFromForward (:1046-1058), pairwise getLogSum over the final contexts.
//   counts   : weight * exp((fwd + bI) - res) per edge, evaluated as exp(bI - offset) * exp((fwd + offset) - res)
//              (one exp per context instead of per edge; SURVEY §8d "legitimate flop reductions");
//              emission statistics from the context posterior exp((fwd + bwd) - res).
#define JHX_MODE_COUNTS 0
#define JHX_MODE_DECODE 1
#define JHX_MODE_POSTERIOR 2

template <int G>
__global__ void __launch_bounds__(256) k_exact_fb(DevModel M, DevData D, unsigned long long *counter, double *fwd_scratch,
                                                  int64_t scratch_rows, int mode, double *stats_global, int stats_in_smem,
                                                  PathOut paths, double *loglik, double *post_out, const int32_t *seq_list,
                                                  int64_t n_list, Cond cond) {
  extern __shared__ double smem[];
  if (cond.flag && *cond.flag != cond.want) return;
  if (cond.n_list_dev) n_list = *cond.n_list_dev;
  const int n_stats = M.n_child + M.n_cond * M.a;
  double *sstats = smem;
  double *group_smem = smem + (stats_in_smem ? n_stats : 0);
  double *stats = stats_in_smem ? sstats : stats_global;
  if (stats_in_smem) {
    for (int i = threadIdx.x; i < n_stats; i += blockDim.x) sstats[i] = 0;
    __syncthreads();
  }
  Group<G> g = make_group<G>(M, group_smem);
  ResReader rd;
  double *fwd = fwd_scratch + (size_t)g.slot * scratch_rows * M.Cmax;
  double score_sum = 0;
  const int64_t n_work = seq_list ? n_list : D.n_seq;
  for (;;) {
    long long it = 0;
    if (g.lane == 0) it = (long long)atomicAdd(counter, 1ull);
    it = g.bcast(it);
    if (it >= n_work) break;
    int sid = seq_list ? seq_list[it] : D.order[it];
    int64_t off = D.offsets[sid], L = D.offsets[sid + 1] - off;
    double w = D.weights ? D.weights[sid] : 1.0;
    rd.init(D.codes + off);
    // allowed states groups constrain training (baumWelch(..., allowedStatesGroup), HOH:723-726); the state posteriors of the
    // reference are computed without them (fillLogStatePosteriorMatrix -> fillFwdMatrix(startPos, endPos, seq), HOH:309-310)
    const int32_t *grp = (D.groups && mode == JHX_MODE_COUNTS) ? D.groups + off : nullptr;
    // ---------------- forward ----------------
    double *cur = g.v0, *nxt = g.v1;
    if (g.lane == 0) {
      cur[0] = 0.0;  // getLogSum of the single seed 0 (HOH:364-366): 0 + log(exp(0)) == 0
      fwd[0] = 0.0;
    }
    int h = hash_at(M, rd, 0);
    for (int64_t l = 0; l < L; l++) {
      fill_em<G>(M, g, l, h);
      g.sync();
      int lc = l < M.m ? (int)l : M.m, lct = (l + 1) < M.m ? (int)(l + 1) : M.m;
      int nct = M.lc_ctx_ptr[lct + 1] - M.lc_ctx_ptr[lct];
      const int *ip = M.in_ptr + (size_t)lc * (M.Cmax + 1);
      const int *isrc = M.in_src + (size_t)lc * M.n_child;
      const int *ipp = M.in_p + (size_t)lc * M.n_child;
      for (int c = g.lane; c < nct; c += G) {
        int q0 = ip[c], q1 = ip[c + 1];
        double r = JH_NEG_INF;
        if (q1 > q0) {
          double offm = JH_NEG_INF;
          for (int q = q0; q < q1; q++) {
            int p = ipp[q];
            double v = dadd(dadd(cur[isrc[q]], g.em[M.child_state[p]]), M.tscore[p]);  // HOH:391-393
            if (v > offm) offm = v;
          }
          if (!isinf(offm)) {
            double sum = 0;
            for (int q = q0; q < q1; q++) {
              int p = ipp[q];
              double v = dadd(dadd(cur[isrc[q]], g.em[M.child_state[p]]), M.tscore[p]);
              sum = dadd(sum, exp(dsub(v, offm)));
            }
            r = dadd(offm, log(sum));
          }
        }
        if (!ctx_allowed(M, grp, l + 1, l + 1, c)) r = JH_NEG_INF;  // fwdMatrix is only filled for the allowed contexts (HOH:386-388)
        nxt[c] = r;
        fwd[(size_t)(l + 1) * M.Cmax + c] = r;
      }
      g.sync();
      double *t = cur; cur = nxt; nxt = t;
      if (l + 1 < L) h = hash_fwd(M, rd, h, l);
    }
    // res = computeLogScoreFromForward(L)
    double res = JH_NEG_INF;
    {
      int lcL = L < M.m ? (int)L : M.m;
      int ncL = M.lc_ctx_ptr[lcL + 1] - M.lc_ctx_ptr[lcL];
      if (g.lane == 0) {
        for (int i = 0; i < ncL; i++) {
          int ls = M.ctx_last[M.lc_ctx_ptr[lcL] + i];
          if (ls >= 0 && M.is_final[ls]) res = lse2(res, cur[i]);
        }
      }
      res = g.bcastd(res);
    }
    // ---------------- backward ----------------
    double *bn = g.v0, *bc = g.v1;
    g.sync();
    init_last_layer<G>(M, g, L, bn, grp);
    h = hash_at(M, rd, L - 1);
    g.sync();
    const bool counts = mode == JHX_MODE_COUNTS && !isinf(res);
    for (int64_t l = L - 1; l >= 0; --l) {
      fill_em<G>(M, g, l, h);
      g.sync();
      int lc = l < M.m ? (int)l : M.m, lct = (l + 1) < M.m ? (int)(l + 1) : M.m;
      int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
      int nct = M.lc_ctx_ptr[lct + 1] - M.lc_ctx_ptr[lct];
      if (counts) {
        // emission statistic of position l: posterior of the contexts of layer l+1 (their last state emitted x[l])
        for (int c = g.lane; c < nct; c += G) {
          int st = M.ctx_last[M.lc_ctx_ptr[lct] + c];
          if (M.state_order[st] >= 0 && l >= M.state_order[st]) {
            double pw = dmul(w, exp(dsub(dadd(fwd[(size_t)(l + 1) * M.Cmax + c], bn[c]), res)));
            if (pw != 0) atomicAdd(&stats[M.n_child + emis_index(M, st, h)], pw);
          }
        }
      } else if (mode != JHX_MODE_COUNTS) {
        // state posterior of position l (layer l+1), HOH:311-323: pairwise getLogSum over the contexts ending in s
        double best = JH_NEG_INF;
        int best_s = 0x7fffffff;
        for (int s = g.lane; s < M.S; s += G) {
          double acc = JH_NEG_INF;
          for (int c = 0; c < nct; c++)
            if (M.ctx_last[M.lc_ctx_ptr[lct] + c] == s && M.state_order[s] >= 0)
              acc = lse2(acc, dadd(fwd[(size_t)(l + 1) * M.Cmax + c], bn[c]));
          if (post_out) post_out[(size_t)s * L + l] = acc;  // - logProb is applied when the sweep has finished
          if (acc > best) { best = acc; best_s = s; }  // ascending s per lane: first maximum
        }
        if (paths) {
#pragma unroll
          for (int d = G / 2; d >= 1; d >>= 1) {
            double ob = __shfl_xor_sync(g.mask, best, d);
            int os = __shfl_xor_sync(g.mask, best_s, d);
            if (ob > best || (ob == best && os < best_s)) { best = ob; best_s = os; }
          }
          if (g.lane == 0) path_store(paths, off + l, best_s == 0x7fffffff ? 0 : best_s);  // all -inf: index 0 (AHMM:1130-1137)
        }
      }
      for (int c = g.lane; c < nc; c += G) {
        double offm;
        double b = ctx_bwd_lse(M, lc, c, bn, g.em, &offm);
        if (!ctx_allowed(M, grp, l, l, c)) { b = JH_NEG_INF; offm = JH_NEG_INF; }
        bc[c] = b;
        if (counts && !isinf(offm)) {
          double f = fwd[(size_t)l * M.Cmax + c];
          double cf = dmul(w, exp(dsub(dadd(f, offm), res)));
          if (cf != 0) {
            int e = M.ctx_elem[M.lc_ctx_ptr[lc] + c];
            int p0 = M.elem_child_ptr[e], p1 = M.elem_child_ptr[e + 1];
            const int *cn = M.child_next + (size_t)lc * M.n_child;
            for (int p = p0; p < p1; p++) {
              double t = exp(dsub(edge_value(M, cn, p, bn, g.em), offm));
              if (t != 0) atomicAdd(&stats[M.child_stat[p]], dmul(t, cf));
            }
          }
        }
      }
      g.sync();
      double *t = bc; bc = bn; bn = t;
      if (l > 0) h = hash_back(M, rd, h, l);
    }
    double lp = bn[0];  // bwd[0][0]
    g.sync();
    if (g.lane == 0) {
      if (loglik) loglik[sid] = lp;
      score_sum = dadd(score_sum, dmul(w, lp));
    }
    if (post_out) {
      g.sync();
      for (int64_t i = g.lane; i < (int64_t)M.S * L; i += G) post_out[i] = dsub(post_out[i], lp);
    }
    g.sync();
  }
  if (mode == JHX_MODE_COUNTS) {
    if (g.lane == 0 && score_sum != 0) atomicAdd(&stats_global[n_stats], score_sum);
    if (stats_in_smem) {
      __syncthreads();
      for (int i = threadIdx.x; i < n_stats; i += blockDim.x)
        if (sstats[i] != 0) atomicAdd(&stats_global[i], sstats[i]);
    }
  }
}

}  // namespace jhx
