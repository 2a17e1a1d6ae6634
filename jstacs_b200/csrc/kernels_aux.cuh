// kernels_aux.cuh — parameter preparation, M-step, prior, data checks, FP64 peak probe.
#pragma once
#include "common.cuh"

namespace jha {

struct DevParams {
  int n_elem, n_child, n_cond, a, S;
  const int *elem_child_ptr;  // [n_elem+1]
  const int *child_elem;      // [n_child]
  const int *trans_index;     // [n_elem]
  const int *row_emitting;    // [n_cond] 1 (all rows belong to emitting emissions)
  double *trans_param, *trans_lognorm, *emis_param, *emis_lognorm;
  const double *trans_hyper, *emis_hyper;
  double *tscore, *escore;  // log-space score tables of the exact kernels
};

// Normalisation.getLogSum(0,n,v) (Normalisation.java:65-99)
__device__ inline double log_sum_n(const double *v, int n) {
  double off = JH_NEG_INF, sum = 0;
  for (int i = 0; i < n; i++)
    if (v[i] > off) off = v[i];
  if (isinf(off)) return JH_NEG_INF;
  for (int i = 0; i < n; i++) sum = __dadd_rn(sum, exp(__dsub_rn(v[i], off)));
  return __dadd_rn(off, log(sum));
}

// score tables: parameters[c] - logNorm (BasicHigherOrderTransition.java:1116-1118) and
// (0 - logNorm[cond]) + params[cond][sym] (AbstractConditionalDiscreteEmission.java:445-456)
__global__ void k_prep_scores(DevParams P) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n_child) P.tscore[i] = __dsub_rn(P.trans_param[i], P.trans_lognorm[P.child_elem[i]]);
  if (i < P.n_cond * P.a) P.escore[i] = __dadd_rn(__dsub_rn(0.0, P.emis_lognorm[i / P.a]), P.emis_param[i]);
}

// precompute() of both parameter families: logNorm = logSumNormalisation(parameters)
// (TransitionElement.java:119-126, AbstractConditionalDiscreteEmission.java:466-480)
__global__ void k_precompute_lognorm(DevParams P, int do_trans, int do_emis) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (do_trans && i < P.n_elem) {
    int p0 = P.elem_child_ptr[i], n = P.elem_child_ptr[i + 1] - p0;
    P.trans_lognorm[i] = log_sum_n(P.trans_param + p0, n);
  }
  if (do_emis && i < P.n_cond) P.emis_lognorm[i] = log_sum_n(P.emis_param + (size_t)i * P.a, P.a);
}

// M-step == estimateFromStatistic of every transition element and emission
// (BasicHigherOrderTransition.java:458-466 + :1242-1262, AbstractConditionalDiscreteEmission.java:682-699).
// One block; the statistics are modified in place (hyper-parameters added) like in the reference.
__global__ void k_mstep(DevParams P, double *trans_stat, double *emis_stat) {
  for (int t = threadIdx.x; t < P.n_elem; t += blockDim.x) {
    if (P.trans_index[t] != t) continue;
    int p0 = P.elem_child_ptr[t], n = P.elem_child_ptr[t + 1] - p0;
    if (n <= 0) continue;
    double ln = 0;
    for (int i = 0; i < n; i++) {
      double s = __dadd_rn(trans_stat[p0 + i], P.trans_hyper[p0 + i]);
      trans_stat[p0 + i] = s;
      ln = __dadd_rn(ln, s);
      P.trans_param[p0 + i] = log(s);
    }
    if (ln == 0) {
      for (int i = 0; i < n; i++) P.trans_param[p0 + i] = 0;
      ln = n;
    }
    ln = log(ln);
    for (int i = 0; i < n; i++) P.trans_param[p0 + i] = __dsub_rn(P.trans_param[p0 + i], ln);
    P.trans_lognorm[t] = log_sum_n(P.trans_param + p0, n);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < P.n_elem; t += blockDim.x) {  // tied elements: setParameters + precompute
    int r = P.trans_index[t];
    if (r == t) continue;
    int p0 = P.elem_child_ptr[t], q0 = P.elem_child_ptr[r], n = P.elem_child_ptr[t + 1] - p0;
    for (int i = 0; i < n; i++) P.trans_param[p0 + i] = P.trans_param[q0 + i];
    P.trans_lognorm[t] = log_sum_n(P.trans_param + p0, n);
  }
  for (int r = threadIdx.x; r < P.n_cond; r += blockDim.x) {
    double *st = emis_stat + (size_t)r * P.a;
    const double *hy = P.emis_hyper + (size_t)r * P.a;
    double sum = 0;
    for (int i = 0; i < P.a; i++) {
      st[i] = __dadd_rn(st[i], hy[i]);
      sum = __dadd_rn(sum, st[i]);
    }
    if (sum == 0) {
      for (int i = 0; i < P.a; i++) st[i] = 1;
      sum = P.a;
    }
    for (int i = 0; i < P.a; i++) P.emis_param[(size_t)r * P.a + i] = log(st[i] / sum);
    P.emis_lognorm[r] = 0;
  }
}

// HigherOrderHMM.getLogPriorTerm (:253-259): the sums run in the reference's order on ONE thread (the value is reproduced bit
// for bit), but the operands are staged through shared memory by the whole block first — a lone thread chasing 18 000
// global loads took 1.7 ms per EM iteration for BASELINE configs[2] (0.8 % of the step, more for higher emission orders).
// out[0] = prior (+ add[0] if add != nullptr).  One block of LOG_PRIOR_THREADS threads, LOG_PRIOR_SMEM doubles of shared memory.
constexpr int LOG_PRIOR_THREADS = 256, LOG_PRIOR_SMEM = 5120;  // 40 KB
__global__ void __launch_bounds__(LOG_PRIOR_THREADS) k_log_prior(DevParams P, const double *add, double *out) {
  __shared__ double sh[LOG_PRIOR_SMEM];
  __shared__ double s_pt, s_res, s_r;
  __shared__ int s_open;
  const int tid = threadIdx.x, a = P.a;
  // ---- transitions: children of the elements in tiles of whole elements ----
  if (tid == 0) s_pt = 0;
  for (int t0 = 0; t0 < P.n_elem;) {
    int t1 = t0;
    const int p_base = P.elem_child_ptr[t0];
    while (t1 < P.n_elem && 2 * (P.elem_child_ptr[t1 + 1] - p_base) <= LOG_PRIOR_SMEM) t1++;
    if (t1 == t0) {  // an element with more children than the tile holds: straight from global memory
      if (tid == 0 && P.trans_index[t0] == t0) {
        int p0 = P.elem_child_ptr[t0], n = P.elem_child_ptr[t0 + 1] - p0;
        double r = 0, shy = 0;
        for (int d = 0; d < n; d++) {
          shy = __dadd_rn(shy, P.trans_hyper[p0 + d]);
          r = __dadd_rn(r, __dmul_rn(P.trans_hyper[p0 + d], P.trans_param[p0 + d]));
        }
        s_pt = __dadd_rn(s_pt, n > 1 && shy != 0 ? __dsub_rn(r, __dmul_rn(shy, P.trans_lognorm[t0])) : 0.0);
      }
      t0++;
      continue;
    }
    const int nn = P.elem_child_ptr[t1] - p_base;
    __syncthreads();
    for (int i = tid; i < nn; i += LOG_PRIOR_THREADS) {
      sh[i] = P.trans_hyper[p_base + i];
      sh[nn + i] = P.trans_param[p_base + i];
    }
    __syncthreads();
    if (tid == 0) {
      double pt = s_pt;
      for (int t = t0; t < t1; t++) {
        if (P.trans_index[t] != t) continue;
        const int p0 = P.elem_child_ptr[t] - p_base, n = P.elem_child_ptr[t + 1] - P.elem_child_ptr[t];
        double term = 0;
        if (n > 1) {
          double r = 0, shy = 0;
          for (int d = 0; d < n; d++) {
            shy = __dadd_rn(shy, sh[p0 + d]);
            r = __dadd_rn(r, __dmul_rn(sh[p0 + d], sh[nn + p0 + d]));
          }
          term = shy == 0 ? 0 : __dsub_rn(r, __dmul_rn(shy, P.trans_lognorm[t]));
        }
        pt = __dadd_rn(pt, term);
      }
      s_pt = pt;
    }
    t0 = t1;
  }
  __syncthreads();
  // ---- emissions: rows are grouped by emission; the reference sums per emission and then adds to the total.  Rows of one
  // emission are contiguous; row_emitting[] == 2 marks the first row of each emission (restores the association order).
  if (tid == 0) { s_res = s_pt; s_r = 0; s_open = 0; }
  const int per_row = 2 * a + 1;
  const int tile = LOG_PRIOR_SMEM / per_row;  // rows per tile (tile == 0: alphabet too large for the tile, global path)
  for (int row0 = 0; row0 < P.n_cond; row0 += max(tile, 1)) {
    const int rows = min(max(tile, 1), P.n_cond - row0);
    __syncthreads();
    if (tile > 0) {
      for (int i = tid; i < rows * a; i += LOG_PRIOR_THREADS) {
        sh[i] = P.emis_hyper[(size_t)row0 * a + i];
        sh[rows * a + i] = P.emis_param[(size_t)row0 * a + i];
      }
      for (int i = tid; i < rows; i += LOG_PRIOR_THREADS) sh[2 * rows * a + i] = P.emis_lognorm[row0 + i];
    }
    __syncthreads();
    if (tid == 0) {
      double res = s_res, r = s_r;
      bool open = s_open != 0;
      for (int k = 0; k < rows; k++) {
        const int row = row0 + k;
        const double *hy = tile > 0 ? sh + (size_t)k * a : P.emis_hyper + (size_t)row * a;
        const double *pa = tile > 0 ? sh + (size_t)rows * a + (size_t)k * a : P.emis_param + (size_t)row * a;
        const double ln = tile > 0 ? sh[2 * rows * a + k] : P.emis_lognorm[row];
        if (P.row_emitting[row] == 2) {
          if (open) res = __dadd_rn(res, r);
          r = 0;
          open = true;
        }
        double ess = 0;
        for (int j = 0; j < a; j++) ess = __dadd_rn(ess, hy[j]);
        if (ess > 0) {
          r = __dadd_rn(r, __dmul_rn(-ess, ln));
          for (int j = 0; j < a; j++) r = __dadd_rn(r, __dmul_rn(hy[j], pa[j]));
        }
      }
      s_res = res; s_r = r; s_open = open ? 1 : 0;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double res = s_res;
    if (s_open) res = __dadd_rn(res, s_r);
    out[0] = add ? __dadd_rn(res, add[0]) : res;
  }
}

// 2-bit packed residues (four per byte, lowest bits first, every sequence on a byte boundary) -> uint8 codes; also the
// largest code (alphabet check).  One block per sequence (grid-stride), a thread widens 4 packed bytes = 16 residues.
__global__ void k_unpack2(const uint8_t *packed, const int64_t *byte_ptr, const int64_t *offsets, int64_t n_seq, uint8_t *codes, int alphabet,
                          unsigned int *max_out) {
  unsigned int mx = 0;
  for (int64_t n = blockIdx.x; n < n_seq; n += gridDim.x) {
    const uint8_t *src = packed + byte_ptr[n];
    uint8_t *dst = codes + offsets[n];
    const int64_t L = offsets[n + 1] - offsets[n];
    for (int64_t i = (int64_t)threadIdx.x * 16; i < L; i += (int64_t)blockDim.x * 16) {
      unsigned w[4] = {0, 0, 0, 0};
      const int64_t nb = min((int64_t)4, (L - i + 3) / 4);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (k < nb) {
          const unsigned b = src[i / 4 + k];
          w[k] = (b & 3u) | ((b >> 2 & 3u) << 8) | ((b >> 4 & 3u) << 16) | ((b >> 6 & 3u) << 24);
        }
      }
      if (i + 16 <= L && (((unsigned long long)(dst + i)) & 15ull) == 0) {
        *reinterpret_cast<uint4 *>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
        if (alphabet < 4)
          for (int k = 0; k < 16; k++) mx = max(mx, (w[k >> 2] >> ((k & 3) * 8)) & 0xffu);
      } else {
        for (int64_t k = 0; k < 16 && i + k < L; k++) {
          const unsigned c = (w[k >> 2] >> ((k & 3) * 8)) & 0xffu;
          dst[i + k] = (uint8_t)c;
          mx = max(mx, c);
        }
      }
    }
  }
  if (alphabet < 4) {  // codes are < 4 by construction
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_out, mx);
  }
}

// E-step whose statistics came out non-finite (flag set by k_check_finite): zero them before the log-space kernels redo
// the whole E-step — decided on the device, the EM loop has no host round trip.
__global__ void k_zero_if(const int32_t *flag, double *p, int n) {
  if (*flag == 0) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0.0;
}

// objective of EM iteration *iter -> objective[*iter % cap]; increments the device-side iteration counter (graph replay)
__global__ void k_store_objective(const double *value, double *objective, int32_t *iter, int cap) {
  objective[*iter % cap] = *value;
  *iter += 1;
}

// int32 state paths -> one byte per residue (255 = no state: the int32 path holds -1), 16 residues per thread; only the
// thread-per-sequence kernels of the general context graph still produce int32 paths that have to be narrowed
__global__ void k_narrow_paths(const int32_t *paths, int64_t n, uint8_t *out) {
  int64_t i = (int64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
  for (; i < n; i += stride) {
    if (i + 16 <= n) {
      unsigned w[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int4 v = *reinterpret_cast<const int4 *>(paths + i + 4 * k);
        w[k] = (unsigned)(v.x & 0xff) | ((unsigned)(v.y & 0xff) << 8) | ((unsigned)(v.z & 0xff) << 16) | ((unsigned)(v.w & 0xff) << 24);
      }
      *reinterpret_cast<uint4 *>(out + i) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      for (int64_t k = i; k < n; k++) out[k] = (uint8_t)(paths[k] & 0xff);
    }
  }
}

// Fixed-length paths (models without silent states) behind the variable-length interface of jh_viterbi_paths /
// jh_viterbi_windows: path_len[i] = L_i, or -1 when no path has a probability > 0 (the reference throws "Impossible
// path", HigherOrderHMM.java:659); the unused capacity behind a path is filled with -1.
// lens: a = offsets (b == nullptr: L_i = a[i+1] - a[i]) or a = start, b = end (L_i = b[i] - a[i] + 1).
__global__ void k_finish_paths(const double *scores, const int64_t *a, const int64_t *b, int64_t n, int32_t *paths, const int64_t *path_ptr,
                               int64_t *path_len) {
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const int64_t L = b ? b[i] - a[i] + 1 : a[i + 1] - a[i];
    if (threadIdx.x == 0 && path_len) path_len[i] = isinf(scores[i]) ? -1 : L;
    if (paths && path_ptr)
      for (int64_t k = path_ptr[i] + L + threadIdx.x; k < path_ptr[i + 1]; k += blockDim.x) paths[k] = -1;
  }
}

// largest residue code of the packed data (alphabet check of AbstractHMM.java:991-993 done once per data set)
__global__ void k_max_code(const uint8_t *codes, int64_t n, unsigned int *out) {
  unsigned int mx = 0;
  int64_t i = (int64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 16, stride = (int64_t)gridDim.x * blockDim.x * 16;
  for (; i + 16 <= n; i += stride) {
    uint4 v = *reinterpret_cast<const uint4 *>(codes + i);
    unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      unsigned int x = w[k];
      unsigned int m = max(max(x & 0xff, (x >> 8) & 0xff), max((x >> 16) & 0xff, x >> 24));
      mx = max(mx, m);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t j = n - (n % 16); j < n; j++) mx = max(mx, (unsigned int)codes[j]);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}

// FP64 FMA peak probe: 8 independent register-resident DFMA chains per thread.
__global__ void k_fp64_peak(double *out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 0.9999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;
}

// DFMA latency/throughput probe: C independent chains per thread
template <int C>
__global__ void k_fp64_chains(double *out, int iters, double seed) {
  double a[C];
#pragma unroll
  for (int c = 0; c < C; c++) a[c] = seed + c;
  const double m = 0.9999999, k = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < C; c++) a[c] = fma(a[c], m, k);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; c++) s += a[c];
  if (s == 12345.678) out[0] = s;
}

// FP64 tensor-core probe (measurement only; the product kernels are SIMT): C independent m8n8k4 DMMA accumulator chains
// per warp.  One DMMA = 8*8*4 FMAs per warp.
template <int C>
__global__ void k_dmma_chains(double *out, int iters, double seed) {
  double c0[C], c1[C];
#pragma unroll
  for (int c = 0; c < C; c++) { c0[c] = seed + c; c1[c] = seed - c; }
  double a = 0.9999999 + 1e-9 * (threadIdx.x & 31), b = 1.0000001;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < C; c++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; c++) s += c0[c] + c1[c];
  if (s == 12345.678) out[0] = s;
}

}  // namespace jha
