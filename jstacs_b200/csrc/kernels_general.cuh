// kernels_general.cuh — reference-order log-space kernels for the GENERAL context graph: silent states (adv = 0 edges
// inside a layer, BasicHigherOrderTransition.java:191-257), any transition order m >= 1, trailing silent states
// (HigherOrderHMM.java:407-432, 487-527, 667-705), and scoring of sub-sequences [start,end] whose emissions see the
// residues before `start` (getLogProbFor(seq,start,end); HigherOrderDiscreteEmission.java:138-146 uses absolute positions).
//
// Mapping: ONE SEQUENCE (or window) PER THREAD.  Silent edges make the contexts of a layer depend on each other in
// topological order, so the contexts of one sequence are swept sequentially exactly like the reference does; the
// parallelism is over sequences — profile HMMs (HMMFactory.createProfileHMM, parseProfileHMMFromHMMer) are run over many
// short sequences or windows (projects/xanthogenomes/NHMMer.java:225-244).  DP rows live in global scratch with the
// thread index fastest ([row][context][thread]) so that the lanes of a warp, which walk the same context graph in lock
// step, touch consecutive addresses.  Arithmetic and summation order are the reference's: (bwd + em) + trans, n-ary
// max-offset log-sum-exp in child order, ToolBox.max, first child minimising (current - bwd)^2 with strict '<'.
#pragma once
#include "kernels_exact.cuh"

namespace jhg {

using jhx::dadd;
using jhx::dmul;
using jhx::dsub;
using jhx::lse2;
using jhx::ResReader;

struct GenModel {            // additions to DevModel for silent states
  const unsigned char *silent;  // [S]
  const int *sin_ptr;           // [(m+1)*(Cmax+1)] silent in-edges of the contexts of class lc (sources in the same layer)
  const int *sin_src;           // [(m+1)*n_child]
  const int *sin_p;             // [(m+1)*n_child]
};

using ::WorkList;

struct Rows {  // per-thread DP rows in global memory, thread index fastest
  double *p;
  int64_t stride;  // number of threads of the grid
  int Cmax;
  __device__ __forceinline__ double &at(int64_t row, int c) const { return p[((size_t)row * Cmax + c) * stride]; }
};

__device__ __forceinline__ int lc_of(const DevModel &M, int64_t layer) { return layer < M.m ? (int)layer : M.m; }

// emission score of `state` at absolute position pos (context hash h of that position); SilentEmission scores 0
__device__ __forceinline__ double emis_of(const DevModel &M, int state, int64_t pos, int h) {
  const int k = M.state_order[state];
  if (k < 0) return 0.0;
  if (pos < k) return dadd(0.0, M.log_uniform);
  return M.escore[jhx::emis_index(M, state, h)];
}

// value of child p of a context in layer l: (bwd[l+adv][next] + em) + trans   (HOH:546-549); silent children read the
// row of the same layer, which the descending sweep has already written
__device__ __forceinline__ double child_value(const DevModel &M, const GenModel &X, int lc, int p, const Rows &B, int64_t row_same, int64_t row_next,
                                              int64_t pos, int h, bool with_emission) {
  const int st = M.child_state[p];
  const int nxt = M.child_next[(size_t)lc * M.n_child + p];
  const bool sil = X.silent[st] != 0;
  const double b = B.at(sil ? row_same : row_next, nxt);
  return dadd(with_emission ? dadd(b, emis_of(M, st, pos, h)) : b, M.tscore[p]);
}

// One layer of the backward sweep (HOH:530-570): contexts descending; LIKELIHOOD: n-ary getLogSum, VITERBI: ToolBox.max
// plus the forward-trace decision (first child, strict '<' on the squared distance) as one byte.
template <bool VITERBI>
__device__ __forceinline__ void bwd_layer(const DevModel &M, const GenModel &X, int64_t l, const Rows &B, int64_t row_cur, int64_t row_next, int64_t pos,
                                          int h, uint8_t *choice, int64_t cstride) {
  const int lc = lc_of(M, l);
  const int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
  for (int c = nc - 1; c >= 0; --c) {
    const int e = M.ctx_elem[M.lc_ctx_ptr[lc] + c];
    const int p0 = M.elem_child_ptr[e], p1 = M.elem_child_ptr[e + 1];
    double res = JH_NEG_INF;
    int ch = 255;
    if (p1 > p0) {
      if (VITERBI) {
        double mx = child_value(M, X, lc, p0, B, row_cur, row_next, pos, h, true);
        for (int p = p0 + 1; p < p1; p++) {
          const double v = child_value(M, X, lc, p, B, row_cur, row_next, pos, h, true);
          if (mx < v) mx = v;
        }
        double best = INFINITY;
        for (int p = p0; p < p1; p++) {
          double dist = dsub(child_value(M, X, lc, p, B, row_cur, row_next, pos, h, true), mx);
          dist = dmul(dist, dist);
          if (dist < best) { best = dist; ch = p - p0; }
        }
        res = mx;
      } else {
        double off = JH_NEG_INF;
        for (int p = p0; p < p1; p++) {
          const double v = child_value(M, X, lc, p, B, row_cur, row_next, pos, h, true);
          if (v > off) off = v;
        }
        if (!isinf(off)) {
          double sum = 0;
          for (int p = p0; p < p1; p++) sum = dadd(sum, exp(dsub(child_value(M, X, lc, p, B, row_cur, row_next, pos, h, true), off)));
          res = dadd(off, log(sum));
        }
      }
    }
    B.at(row_cur, c) = res;
    if (VITERBI) choice[(size_t)c * cstride] = (uint8_t)ch;
  }
}

// Layer L (HOH:485-527): bwd[L][c] = final ? 0 : -inf, combined with the silent children of c.
template <bool VITERBI>
__device__ __forceinline__ void bwd_last_layer(const DevModel &M, const GenModel &X, int64_t L, const Rows &B, int64_t row, uint8_t *choice,
                                               int64_t cstride) {
  const int lc = lc_of(M, L);
  const int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
  for (int c = nc - 1; c >= 0; --c) {
    const int e = M.ctx_elem[M.lc_ctx_ptr[lc] + c];
    const int p0 = M.elem_child_ptr[e], p1 = M.elem_child_ptr[e + 1];
    const int ls = M.ctx_last[M.lc_ctx_ptr[lc] + c];
    const double val = (M.m == 0 || (ls >= 0 && M.is_final[ls])) ? 0.0 : JH_NEG_INF;
    int ns = 0;
    double off = JH_NEG_INF, mx = JH_NEG_INF;
    for (int p = p0; p < p1; p++) {
      if (!X.silent[M.child_state[p]]) continue;
      const double v = child_value(M, X, lc, p, B, row, row, 0, 0, false);
      if (ns == 0) mx = v;
      else if (mx < v) mx = v;  // ToolBox.max
      if (v > off) off = v;
      ns++;
    }
    double res = val;
    int ch = 255;
    if (ns > 0) {
      if (VITERBI) {
        res = fmax(val, mx);
        // trailing silent states of the trace (HOH:667-705): staying has distance (0 - bwd)^2 for a final state, inf otherwise
        double dist = (M.m == 0 || (ls >= 0 && M.is_final[ls])) ? dsub(0.0, res) : JH_NEG_INF;
        double best = dmul(dist, dist);
        for (int p = p0; p < p1; p++) {
          if (!X.silent[M.child_state[p]]) continue;
          dist = dsub(child_value(M, X, lc, p, B, row, row, 0, 0, false), res);
          dist = dmul(dist, dist);
          if (dist < best) { best = dist; ch = p - p0; }
        }
      } else {
        double inner = JH_NEG_INF;
        if (!isinf(off)) {
          double sum = 0;
          for (int p = p0; p < p1; p++)
            if (X.silent[M.child_state[p]]) sum = dadd(sum, exp(dsub(child_value(M, X, lc, p, B, row, row, 0, 0, false), off)));
          inner = dadd(off, log(sum));
        }
        res = lse2(val, inner);
      }
    }
    B.at(row, c) = res;
    if (VITERBI) choice[(size_t)c * cstride] = (uint8_t)ch;
  }
}

// ------------------------------------------------------------------------------------------------
// AbstractHMM.logProb (AbstractHMM.java:1049-1070) for sequences or windows: backward sweep with two rolling rows.
__global__ void __launch_bounds__(128) k_general_loglik(DevModel M, GenModel X, DevData D, WorkList W, unsigned long long *counter, double *rows,
                                                        double *out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  Rows B{rows + tid, nthreads, M.Cmax};
  ResReader rd;
  for (;;) {
    const long long it = (long long)atomicAdd(counter, 1ull);
    if (it >= W.n) break;
    const int sid = W.seq ? W.seq[it] : D.order[it];
    const int64_t off = D.offsets[sid], Lseq = D.offsets[sid + 1] - off;
    const int64_t s0 = W.start ? W.start[it] : 0, s1 = W.end ? W.end[it] : Lseq - 1;
    const int64_t L = s1 - s0 + 1;
    rd.init(D.codes + off);
    bwd_last_layer<false>(M, X, L, B, L & 1, nullptr, 0);
    int h = L > 0 ? jhx::hash_at(M, rd, s1) : 0;
    for (int64_t l = L - 1; l >= 0; --l) {
      bwd_layer<false>(M, X, l, B, l & 1, (l + 1) & 1, s0 + l, h, nullptr, 0);
      if (l > 0) h = jhx::hash_back(M, rd, h, s0 + l);
    }
    out[W.seq ? it : sid] = B.at(0, 0);
  }
}

// HigherOrderHMM.viterbi (:617-708) for the general graph.  choice bytes: [(L+1) layers][Cmax][thread].
// paths: CSR by path_ptr (capacity per sequence), path_len = number of states written, -1 if no child could be chosen.
// stats != nullptr: Viterbi training (the sequence weight is added along the path instead of writing it).
__global__ void __launch_bounds__(128) k_general_viterbi(DevModel M, GenModel X, DevData D, WorkList W, unsigned long long *counter, double *rows,
                                                         uint8_t *choices, int64_t max_layers, int32_t *paths, const int64_t *path_ptr,
                                                         int64_t *path_len, double *scores, double *stats) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  Rows B{rows + tid, nthreads, M.Cmax};
  uint8_t *const cbase = choices + tid;  // + (layer*Cmax + ctx)*nthreads
  ResReader rd;
  double score_sum = 0;
  for (;;) {
    const long long it = (long long)atomicAdd(counter, 1ull);
    if (it >= W.n) break;
    // work item: a whole sequence, or the window start..end of one (getViterbiPathFor(startPos, endPos, seq), HOH:590-598:
    // layer 0 is the window start, the emissions keep their absolute positions); outputs are indexed by the work item then
    const int sid = W.seq ? W.seq[it] : D.order[it];
    const int64_t oid = W.seq ? it : sid;
    const int64_t off = D.offsets[sid], Lseq = D.offsets[sid + 1] - off;
    const int64_t s0 = W.start ? W.start[it] : 0, s1 = W.end ? W.end[it] : Lseq - 1;
    const int64_t L = s1 - s0 + 1;
    const double w = D.weights ? D.weights[sid] : 1.0;
    rd.init(D.codes + off);
    bwd_last_layer<true>(M, X, L, B, L & 1, cbase + (size_t)L * M.Cmax * nthreads, nthreads);
    int h = L > 0 ? jhx::hash_at(M, rd, s1) : 0;
    for (int64_t l = L - 1; l >= 0; --l) {
      bwd_layer<true>(M, X, l, B, l & 1, (l + 1) & 1, s0 + l, h, cbase + (size_t)l * M.Cmax * nthreads, nthreads);
      if (l > 0) h = jhx::hash_back(M, rd, h, s0 + l);
    }
    const double sc = B.at(0, 0);
    if (scores) scores[oid] = sc;
    score_sum = dadd(score_sum, dmul(w, sc));
    // forward trace
    int32_t *p_out = paths ? paths + path_ptr[oid] : nullptr;
    const int64_t cap = paths ? path_ptr[oid + 1] - path_ptr[oid] : 0;
    int64_t np = 0, layer = 0;
    int ctx = 0;
    bool bad = false;
    int hh = (stats && L > 0) ? jhx::hash_at(M, rd, s0) : 0;
    for (;;) {
      const int lc = lc_of(M, layer);
      const int ch = cbase[((size_t)layer * M.Cmax + ctx) * nthreads];
      if (ch == 255) {
        bad = layer < L;  // below layer L a missing choice is the reference's index error; at layer L it ends the path
        break;
      }
      const int p = M.elem_child_ptr[M.ctx_elem[M.lc_ctx_ptr[lc] + ctx]] + ch;
      const int st = M.child_state[p];
      const bool sil = X.silent[st] != 0;
      if (stats) {
        atomicAdd(&stats[M.child_stat[p]], w);
        if (!sil && M.state_order[st] >= 0 && layer >= M.state_order[st]) atomicAdd(&stats[M.n_child + jhx::emis_index(M, st, hh)], w);
      } else if (np < cap) {
        p_out[np] = st;
      }
      np++;
      ctx = M.child_next[(size_t)lc * M.n_child + p];
      if (!sil) {
        if (stats && layer + 1 < L) hh = jhx::hash_fwd(M, rd, hh, s0 + layer);
        layer++;
      }
      if (np > (L + 1) * (int64_t)(M.S + 1)) { bad = true; break; }  // cannot happen for a valid (acyclic silent) graph
    }
    if (path_len) path_len[oid] = bad ? -1 : np;
    if (paths && !bad)
      for (int64_t i = np; i < cap; i++) p_out[i] = -1;
  }
  if (stats && score_sum != 0) atomicAdd(&stats[M.n_child + M.n_cond * M.a], score_sum);
}

// HOH:359-434 fillFwdMatrix in pull form: fwd[l][c] for l = 0..L.  Summand order of the reference's push form: emitting
// in-edges from the previous layer by (source context, child), then silent in-edges of the same layer.
__device__ __forceinline__ void general_forward(const DevModel &M, const GenModel &X, ResReader &rd, int64_t L, const Rows &Fw, int64_t s0 = 0) {  // s0: window start
  int h = 0;  // context hash of position l-1 (emission of the children that entered layer l)
  for (int64_t l = 0; l <= L; l++) {
    const int lc = lc_of(M, l);
    const int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
    const int lcs = lc_of(M, l - 1);  // class of the source layer
    const int *ip = M.in_ptr + (size_t)lcs * (M.Cmax + 1), *isrc = M.in_src + (size_t)lcs * M.n_child, *ipp = M.in_p + (size_t)lcs * M.n_child;
    const int *sp = X.sin_ptr + (size_t)lc * (M.Cmax + 1), *ssrc = X.sin_src + (size_t)lc * M.n_child, *spp = X.sin_p + (size_t)lc * M.n_child;
    if (l > 0) h = l == 1 ? jhx::hash_at(M, rd, s0) : jhx::hash_fwd(M, rd, h, s0 + l - 2);
    for (int c = 0; c < nc; c++) {
      double offm = JH_NEG_INF;
      int n = 0;
      if (l == 0 && c == 0) { offm = 0.0; n = 1; }  // the seed summand 0 (HOH:364-366)
      if (l > 0)
        for (int q = ip[c]; q < ip[c + 1]; q++) {
          const int p = ipp[q];
          const double v = dadd(dadd(Fw.at(l - 1, isrc[q]), emis_of(M, M.child_state[p], s0 + l - 1, h)), M.tscore[p]);
          if (v > offm) offm = v;
          n++;
        }
      for (int q = sp[c]; q < sp[c + 1]; q++) {
        const double v = dadd(dadd(Fw.at(l, ssrc[q]), 0.0), M.tscore[spp[q]]);
        if (v > offm) offm = v;
        n++;
      }
      double r = JH_NEG_INF;
      if (n > 0 && !isinf(offm)) {
        double sum = 0;
        if (l == 0 && c == 0) sum = dadd(sum, exp(dsub(0.0, offm)));
        if (l > 0)
          for (int q = ip[c]; q < ip[c + 1]; q++) {
            const int p = ipp[q];
            const double v = dadd(dadd(Fw.at(l - 1, isrc[q]), emis_of(M, M.child_state[p], s0 + l - 1, h)), M.tscore[p]);
            sum = dadd(sum, exp(dsub(v, offm)));
          }
        for (int q = sp[c]; q < sp[c + 1]; q++) {
          const double v = dadd(dadd(Fw.at(l, ssrc[q]), 0.0), M.tscore[spp[q]]);
          sum = dadd(sum, exp(dsub(v, offm)));
        }
        r = dadd(offm, log(sum));
      }
      Fw.at(l, c) = r;
    }
  }
}

// Baum-Welch E-step for the general graph (HOH:471-571 with add == true): full forward matrix in pull form
// (summand order of the reference's push form: emitting in-edges from the previous layer by (source context, child),
// then silent in-edges of the same layer by (source context, child)), then the backward sweep with
// weight * exp((fwd + bI) - res) per edge.  fwd: [(L+1) layers][Cmax][thread]; bwd: two rolling rows.
__global__ void __launch_bounds__(128) k_general_estep(DevModel M, GenModel X, DevData D, unsigned long long *counter, double *rows, double *fwd,
                                                       int64_t max_layers, double *stats) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  Rows B{rows + tid, nthreads, M.Cmax};
  Rows Fw{fwd + tid, nthreads, M.Cmax};
  ResReader rd;
  double score_sum = 0;
  const int n_stats = M.n_child + M.n_cond * M.a;
  for (;;) {
    const long long it = (long long)atomicAdd(counter, 1ull);
    if (it >= D.n_seq) break;
    const int sid = D.order[it];
    const int64_t off = D.offsets[sid], L = D.offsets[sid + 1] - off;
    const double w = D.weights ? D.weights[sid] : 1.0;
    rd.init(D.codes + off);
    general_forward(M, X, rd, L, Fw);
    int h = 0;
    // res = computeLogScoreFromForward(L) (HOH:1046-1058)
    double res = JH_NEG_INF;
    {
      const int lcL = lc_of(M, L);
      const int ncL = M.lc_ctx_ptr[lcL + 1] - M.lc_ctx_ptr[lcL];
      for (int i = 0; i < ncL; i++) {
        const int ls = M.ctx_last[M.lc_ctx_ptr[lcL] + i];
        if (M.m == 0 || (ls >= 0 && M.is_final[ls])) res = lse2(res, Fw.at(L, i));
      }
    }
    const bool counts = !isinf(res);
    // ---------------- backward + counts ----------------
    bwd_last_layer<false>(M, X, L, B, L & 1, nullptr, 0);
    if (counts) {  // statistics of the silent edges of layer L (HOH:510-513)
      const int lc = lc_of(M, L);
      const int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
      for (int c = nc - 1; c >= 0; --c) {
        const int e = M.ctx_elem[M.lc_ctx_ptr[lc] + c];
        for (int p = M.elem_child_ptr[e]; p < M.elem_child_ptr[e + 1]; p++) {
          if (!X.silent[M.child_state[p]]) continue;
          const double bI = child_value(M, X, lc, p, B, L & 1, L & 1, 0, 0, false);
          const double nw = dmul(w, exp(dsub(dadd(Fw.at(L, c), bI), res)));
          if (nw != 0) atomicAdd(&stats[M.child_stat[p]], nw);
        }
      }
    }
    h = L > 0 ? jhx::hash_at(M, rd, L - 1) : 0;
    for (int64_t l = L - 1; l >= 0; --l) {
      bwd_layer<false>(M, X, l, B, l & 1, (l + 1) & 1, l, h, nullptr, 0);
      if (counts) {
        const int lc = lc_of(M, l);
        const int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
        for (int c = nc - 1; c >= 0; --c) {
          const double f = Fw.at(l, c);
          if (isinf(f)) continue;
          const int e = M.ctx_elem[M.lc_ctx_ptr[lc] + c];
          for (int p = M.elem_child_ptr[e]; p < M.elem_child_ptr[e + 1]; p++) {
            const double bI = child_value(M, X, lc, p, B, l & 1, (l + 1) & 1, l, h, true);
            const double nw = dmul(w, exp(dsub(dadd(f, bI), res)));
            if (nw != 0) {
              const int st = M.child_state[p];
              if (M.state_order[st] >= 0 && l >= M.state_order[st]) atomicAdd(&stats[M.n_child + jhx::emis_index(M, st, h)], nw);
              atomicAdd(&stats[M.child_stat[p]], nw);
            }
          }
        }
      }
      if (l > 0) h = jhx::hash_back(M, rd, h, l);
    }
    score_sum = dadd(score_sum, dmul(w, B.at(0, 0)));
  }
  if (score_sum != 0) atomicAdd(&stats[n_stats], score_sum);
}

// HOH:294-329 fillLogStatePosteriorMatrix(silentZero = true) + AHMM:790-797 for the general graph: full forward and
// backward matrices per thread; post[s][l] = pairwise getLogSum over the contexts of layer l+1 that end in the emitting
// state s of (fwd + bwd), minus bwd[0][0].  post_out != nullptr: the S x L matrix of work item 0's sequence (row-major);
// paths != nullptr: AbstractHMM.decodeStatePosterior (AHMM:1125-1139) per position, lowest index on ties.
// col: [S][thread] scratch column.
__global__ void __launch_bounds__(128) k_general_posterior(DevModel M, GenModel X, DevData D, WorkList W, unsigned long long *counter, double *fwd,
                                                           double *bwd, double *col, double *post_out, PathOut paths, double *loglik) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  Rows Fw{fwd + tid, nthreads, M.Cmax}, B{bwd + tid, nthreads, M.Cmax};
  double *const cl = col + tid;  // + s*nthreads
  ResReader rd;
  for (;;) {
    const long long it = (long long)atomicAdd(counter, 1ull);
    if (it >= W.n) break;
    const int sid = W.seq ? W.seq[it] : D.order[it];
    const int64_t off = D.offsets[sid], Lseq = D.offsets[sid + 1] - off;
    const int64_t s0 = W.start ? W.start[it] : 0, s1 = W.end ? W.end[it] : Lseq - 1;  // window (layer 0 = position s0)
    const int64_t L = s1 - s0 + 1;
    rd.init(D.codes + off);
    if (M.m == 0) {
      // Transition order 0 (HOH:296-308): every position stands alone, post[state][l] = (trans + em) - getLogSum over the
      // children of the single element (child order).  The log-likelihood is the sum of the per-position normalisers only if
      // every state is final, which order 0 implies (HOH:498): it is computed by the backward sweep as for every model.
      const int e0 = M.ctx_elem[M.lc_ctx_ptr[0]];
      const int p0 = M.elem_child_ptr[e0], p1 = M.elem_child_ptr[e0 + 1];
      int h = L > 0 ? jhx::hash_at(M, rd, s0) : 0;
      if (loglik) {
        bwd_last_layer<false>(M, X, L, B, L & 1, nullptr, 0);
        int hb = L > 0 ? jhx::hash_at(M, rd, s1) : 0;
        for (int64_t l = L - 1; l >= 0; --l) {
          bwd_layer<false>(M, X, l, B, l & 1, (l + 1) & 1, s0 + l, hb, nullptr, 0);
          if (l > 0) hb = jhx::hash_back(M, rd, hb, s0 + l);
        }
        loglik[sid] = B.at(0, 0);
      }
      for (int64_t l = 0; l < L; l++) {
        for (int st = 0; st < M.S; st++) cl[(size_t)st * nthreads] = 0.0;  // states that are no child keep the array's initial 0
        double o = JH_NEG_INF;
        for (int p = p0; p < p1; p++) {
          const int st = M.child_state[p];
          const double v = dadd(M.tscore[p], emis_of(M, st, s0 + l, h));
          cl[(size_t)st * nthreads] = v;
          if (v > o) o = v;
        }
        double dn = JH_NEG_INF;
        if (!isinf(o)) {
          double sum = 0;
          for (int p = p0; p < p1; p++) sum = dadd(sum, exp(dsub(cl[(size_t)M.child_state[p] * nthreads], o)));
          dn = dadd(o, log(sum));
        }
        int best = 0;
        double bv = dsub(cl[0], dn);
        if (post_out) post_out[l] = bv;
        for (int st = 1; st < M.S; st++) {
          const double v = dsub(cl[(size_t)st * nthreads], dn);
          if (post_out) post_out[(size_t)st * L + l] = v;
          if (bv < v) { bv = v; best = st; }
        }
        if (paths) path_store(paths, off + s0 + l, best);
        if (l + 1 < L) h = jhx::hash_fwd(M, rd, h, s0 + l);
      }
      continue;
    }
    general_forward(M, X, rd, L, Fw, s0);
    bwd_last_layer<false>(M, X, L, B, L, nullptr, 0);
    int h = L > 0 ? jhx::hash_at(M, rd, s1) : 0;
    for (int64_t l = L - 1; l >= 0; --l) {
      bwd_layer<false>(M, X, l, B, l, l + 1, s0 + l, h, nullptr, 0);
      if (l > 0) h = jhx::hash_back(M, rd, h, s0 + l);
    }
    const double lp = B.at(0, 0);
    if (loglik) loglik[sid] = lp;
    for (int64_t l = 1; l <= L; l++) {
      for (int st = 0; st < M.S; st++) cl[(size_t)st * nthreads] = JH_NEG_INF;
      const int lc = lc_of(M, l);
      const int nc = M.lc_ctx_ptr[lc + 1] - M.lc_ctx_ptr[lc];
      for (int c = 0; c < nc; c++) {
        const int st = M.ctx_last[M.lc_ctx_ptr[lc] + c];
        if (st >= 0 && !X.silent[st]) cl[(size_t)st * nthreads] = lse2(cl[(size_t)st * nthreads], dadd(Fw.at(l, c), B.at(l, c)));
      }
      int best = 0;
      double bv = dsub(cl[0], lp);
      if (post_out) post_out[l - 1] = bv;
      for (int st = 1; st < M.S; st++) {
        const double v = dsub(cl[(size_t)st * nthreads], lp);
        if (post_out) post_out[(size_t)st * L + (l - 1)] = v;
        if (bv < v) { bv = v; best = st; }
      }
      if (paths) path_store(paths, off + s0 + l - 1, best);
    }
  }
}

}  // namespace jhg
