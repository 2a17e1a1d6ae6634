// kernels_scan.cuh — checkpoint pass of the long-sequence path, PARALLEL IN THE POSITION, for small dense models
// (at most 16 states after padding: the classic "CpG island HMM over a chromosome" shape).
//
// The forward recursion is linear: the vector after a chunk is the vector before it times the chunk's S x S transfer
// matrix.  Row i of that matrix is nothing but the ordinary scaled recursion started from the unit vector e_i, so
//   1. k_scan_chunks: one THREAD per (chunk, direction, start state) runs the recursion over its chunk with its own exact
//      power-of-two rescaling (S times the arithmetic of the sequential pass — but all chunks of all sequences at once:
//      122 000 chunks x 2 x S threads for a 250 Mb sequence instead of one dependent chain of 250e6 steps);
//   2. k_scan_combine: one warp per (sequence, direction) walks over the chunk matrices (n_chunks dependent steps of S^2
//      multiply-adds), producing exactly the checkpoints, exponents, P(x) and objective that k_sparse_pass1 produces;
//   3. the chunk kernels of kernels_sparse.cuh (posterior decisions / expected counts) run unchanged.
// Rows keep separate exponents, and rows whose start state cannot explain the chunk are flagged as zero, so no mass is
// lost to a common scale; terms more than 1000 binades below the largest one are dropped in the combination, where
// they could not change an fp64 sum anyway.  The result differs from the sequential pass only by the association of
// the sums (~1e-13 relative), far inside the 1e-9 bound.
#pragma once
#include "kernels_sparse.cuh"

namespace jhp {

using jhf::FastDev;
using jhs::Pass1Extra;

constexpr int ROW_ZERO = -(1 << 28);  // exponent of an all-zero row
constexpr int ROW_BAD = -(1 << 29);   // the row's recursion came near the denormal range: the sequence goes to the fallback list

// exact rescale of a thread's vector; returns false when the vector is zero (expo untouched)
template <int G>
__device__ __forceinline__ bool rescale_thread(double (&v)[G], int &expo, bool &ok) {
  unsigned hi = (unsigned)__double2hiint(v[0]);
#pragma unroll
  for (int j = 1; j < G; j++) hi = max(hi, (unsigned)__double2hiint(v[j]));
  const int be = (int)(hi >> 20);
  if (be == 0) return false;  // zero (or denormal: below anything that matters)
  const double f = __hiloint2double((2046 - be) << 20, 0);
#pragma unroll
  for (int j = 0; j < G; j++) v[j] *= f;
  expo += be - 1023;
  ok &= (unsigned)(be - 81) < (0x7feu - 81u);
  return true;
}

__device__ __forceinline__ int hash_direct(const FastDev &F, const uint8_t *x, int t) {  // sum_{k<=K, t-k>=0} a^k x[t-k]
  int h = 0, pw = 1;
  for (int k = 0; k <= F.Kmax && t - k >= 0; k++) {
    h += pw * (int)__ldg(x + t - k);
    pw *= F.a;
  }
  return h;
}

// R: [chunk][direction][start state][G] rows, RE: [chunk][direction][start state] exponents (ROW_ZERO / ROW_BAD)
template <int G>
__global__ void __launch_bounds__(128) k_scan_chunks(FastDev F, DevData D, int CK, int64_t n_lc, int n_dir, const int32_t *chunk_seq,
                                                     const int32_t *chunk_idx, double *R, int *RE) {
  extern __shared__ double sm[];
  double *sT = sm, *sE = sm + G * G;
  for (int i = threadIdx.x; i < G * G; i += blockDim.x) sT[i] = F.T[i];
  for (int i = threadIdx.x; i < F.H * G; i += blockDim.x) sE[i] = F.E[i];
  __syncthreads();
  const int64_t n_items = n_lc * n_dir * G;  // n_dir = 1: forward only (scoring); start state fastest, then the chunk, direction slowest: a warp runs one direction
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(item % G), dir = (int)(item / (n_lc * G));
    const int64_t g = (item / G) % n_lc;
    const int sid = chunk_seq[g], c = chunk_idx[g];
    const int64_t off = D.offsets[sid];
    const int L = (int)(D.offsets[sid + 1] - off);
    const int t0 = c * CK, t1 = min(L, t0 + CK) - 1;
    const uint8_t *x = D.codes + off;
    double *const Rrow = R + ((size_t)(g * 2 + dir) * G + i) * G;
    int *const REp = RE + (size_t)(g * 2 + dir) * G + i;
    double v[G];
    int ex = 0;
    bool ok = true, live = true;
    auto emis = [&](int h, int t, double (&e)[G]) {
#pragma unroll
      for (int j = 0; j < G; j++) e[j] = sE[(size_t)h * G + j];
      if (t < F.Kmax) {
#pragma unroll
        for (int j = 0; j < G; j++)
          if (t < F.korder[j]) e[j] = F.inv_a;
      }
    };
    if (dir == 0) {  // row i of the forward transfer of chunk c: alpha_{t0-1} = e_i  ->  alpha_{t1}
      int h, t;
      double e[G];
      if (c == 0) {  // the start element takes the place of the vector before the chunk: only "row" 0 exists
        live = i == 0;
        h = hash_direct(F, x, 0);
        emis(h, 0, e);
#pragma unroll
        for (int j = 0; j < G; j++) v[j] = F.pi[j] * e[j];
        t = 1;
      } else {
        h = hash_direct(F, x, t0);
        emis(h, t0, e);
#pragma unroll
        for (int j = 0; j < G; j++) v[j] = sT[i * G + j] * e[j];
        t = t0 + 1;
      }
      live = live && rescale_thread<G>(v, ex, ok);
      for (; live && t <= t1; t++) {
        const int xt = (int)__ldg(x + t);
        h = F.a_pow2 ? (((h * F.a) | xt) & (F.H - 1)) : ((h * F.a + xt) % F.H);
        emis(h, t, e);
        double nv[G];
#pragma unroll
        for (int j = 0; j < G; j++) {
          double s0 = 0, s1 = 0;
#pragma unroll
          for (int k = 0; k < G; k += 2) {
            s0 = fma(v[k], sT[k * G + j], s0);
            s1 = fma(v[k + 1], sT[(k + 1) * G + j], s1);
          }
          nv[j] = (s0 + s1) * e[j];
        }
#pragma unroll
        for (int j = 0; j < G; j++) v[j] = nv[j];
        if ((t & 7) == 0) live = rescale_thread<G>(v, ex, ok);
      }
      if (live) live = rescale_thread<G>(v, ex, ok);
    } else {  // row i of the backward transfer of chunk c: beta_{t1} = e_i  ->  beta_{t0-1}   (chunk 0 has no predecessor)
      live = c > 0;
      if (t1 == L - 1) {  // the last chunk starts from the final-state vector: only "row" 0 exists
        live = live && i == 0;
#pragma unroll
        for (int j = 0; j < G; j++) v[j] = F.fin[j];
      } else {
#pragma unroll
        for (int j = 0; j < G; j++) v[j] = j == i ? 1.0 : 0.0;
      }
      int h = live ? hash_direct(F, x, t1) : 0;
      for (int t = t1; live && t >= t0; t--) {  // beta_t -> beta_{t-1}
        double e[G];
        emis(h, t, e);
        double xb[G], nv[G];
#pragma unroll
        for (int j = 0; j < G; j++) xb[j] = e[j] * v[j];
#pragma unroll
        for (int k = 0; k < G; k++) {
          double s0 = 0, s1 = 0;
#pragma unroll
          for (int j = 0; j < G; j += 2) {
            s0 = fma(sT[k * G + j], xb[j], s0);
            s1 = fma(sT[k * G + j + 1], xb[j + 1], s1);
          }
          nv[k] = s0 + s1;
        }
#pragma unroll
        for (int j = 0; j < G; j++) v[j] = nv[j];
        if ((t & 7) == 0) live = rescale_thread<G>(v, ex, ok);
        const int q = t - 1 - F.Kmax;  // residue that completes the hash of t-1
        h = h / F.a + F.hpow * (q >= 0 ? (int)__ldg(x + q) : 0);
      }
      if (live) live = rescale_thread<G>(v, ex, ok);
    }
#pragma unroll
    for (int j = 0; j < G; j++) Rrow[j] = live ? v[j] : 0.0;
    *REp = !live ? ROW_ZERO : (ok ? ex : ROW_BAD);
  }
}

__device__ __forceinline__ double pow2_neg(int k) {  // 2^k for k <= 0, 0 below the normal range
  return k < -1022 ? 0.0 : __hiloint2double((1023 + k) << 20, 0);
}

// One warp per (long sequence, direction), lane = state.  long_seg[k] = first chunk (in the chunk list) of the k-th long sequence.
// Produces what k_sparse_pass1 produces: ck_alpha / ck_beta rows [chunk_ptr[seq] + c][SP] (+ exponents, P(x), objective in X).
template <int G>
__global__ void __launch_bounds__(128) k_scan_combine(FastDev F, DevData D, int64_t n_long, int n_dir, const int32_t *long_ids, const int64_t *long_seg,
                                                      const int64_t *chunk_ptr, const double *R, const int *RE, int SP, double *ck_alpha,
                                                      double *ck_beta, double *loglik, int32_t *fallback, Pass1Extra X) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_dir * n_long) return;
  const int64_t k = w / n_dir;
  const int dir = (int)(w % n_dir);
  const int sid = long_ids[k];
  const int64_t g0 = long_seg[k], nck = long_seg[k + 1] - g0, cbase = chunk_ptr[sid];
  const bool act = lane < G;
  bool ok = true;
  // The walk over the chunks is one dependent chain; the rows and exponents of a chunk do not depend on it, so the next
  // chunk's are fetched (into registers) while the current one is combined: the global-memory latency leaves the chain.
  struct ChunkRows {
    int e;          // exponent of row `lane`
    double r[G];    // r[i] = row i, entry `lane`
  };
  auto fetch = [&](int64_t g, ChunkRows &C) {
    C.e = act ? RE[(size_t)(g * 2 + dir) * G + lane] : ROW_ZERO;
    const double *rows = R + (size_t)(g * 2 + dir) * G * G + lane;
#pragma unroll
    for (int i = 0; i < G; i++) C.r[i] = act ? rows[(size_t)i * G] : 0.0;
  };
  // v <- sum_i v[i] 2^(e_i - E) row_i; returns false when nothing is left
  auto apply = [&](const ChunkRows &C, double &v, long long &expo) {
    const int my_e = C.e;
    if (__any_sync(FULL, act && v > 0 && my_e == ROW_BAD)) ok = false;
    const bool valid = act && v > 0 && my_e > ROW_ZERO / 2;
    int emax = valid ? my_e : INT_MIN;
#pragma unroll
    for (int dlt = 16; dlt >= 1; dlt >>= 1) emax = max(emax, __shfl_xor_sync(FULL, emax, dlt));
    double res = 0;
    if (emax != INT_MIN) {
      const double sc = valid ? v * pow2_neg(my_e - emax) : 0.0;
#pragma unroll
      for (int i = 0; i < G; i++) res = fma(__shfl_sync(FULL, sc, i), C.r[i], res);
      expo += emax;
    }
    v = res;
    // exact rescale by the exponent of the warp-wide maximum
    unsigned hi = __reduce_max_sync(FULL, (unsigned)__double2hiint(v));
    const int be = (int)(hi >> 20);
    if (be == 0) return false;
    v *= __hiloint2double((2046 - be) << 20, 0);
    expo += be - 1023;
    ok &= (unsigned)(be - 81) < (0x7feu - 81u);
    return true;
  };
  ChunkRows cur, nxt;
  if (dir == 0) {
    const int e0 = RE[(size_t)(g0 * 2) * G];
    double a = act ? R[(size_t)(g0 * 2) * G * G + lane] : 0.0;  // chunk 0, "row" 0: alpha at the end of chunk 0
    long long A = e0 > ROW_ZERO / 2 ? e0 : 0;
    bool alive = e0 > ROW_ZERO / 2;
    if (e0 == ROW_BAD) ok = false;
    if (nck > 1) fetch(g0 + 1, nxt);
    for (int64_t c = 1; c < nck && alive; c++) {
      cur = nxt;
      if (c + 1 < nck) fetch(g0 + c + 1, nxt);
      if (ck_alpha && lane < SP) ck_alpha[(size_t)(cbase + c) * SP + lane] = a;
      if (X.ck_alpha_exp && lane == 0) X.ck_alpha_exp[cbase + c] = A;
      alive = apply(cur, a, A);
    }
    double sfin = act && alive ? a * F.fin[lane] : 0.0;
#pragma unroll
    for (int dlt = 16; dlt >= 1; dlt >>= 1) sfin += __shfl_xor_sync(FULL, sfin, dlt);
    if (lane == 0) {
      const bool good = alive && ok && sfin > 0;
      const double lp = good ? log(sfin) + (double)A * 0.6931471805599453094 : JH_NEG_INF;
      if (loglik) loglik[sid] = lp;
      if (X.seq_sfin) {
        X.seq_sfin[sid] = good ? sfin : 0.0;
        X.seq_AL[sid] = A;
        if (good) atomicAdd(X.score, (D.weights ? D.weights[sid] : 1.0) * lp);
      }
      if (!good) {
        int pos = atomicAdd(fallback, 1);
        fallback[1 + pos] = sid;
      }
    }
  } else {
    const int64_t gl = g0 + nck - 1;  // last chunk, "row" 0: beta at the last position of chunk nck-2
    const int e0 = RE[(size_t)(gl * 2 + 1) * G];
    double b = act ? R[(size_t)(gl * 2 + 1) * G * G + lane] : 0.0;
    long long Bx = e0 > ROW_ZERO / 2 ? e0 : 0;
    bool alive = e0 > ROW_ZERO / 2;
    if (nck > 2) fetch(g0 + nck - 2, nxt);
    for (int64_t c = nck - 2; c >= 0; c--) {  // ck_beta[c] = beta at the last position of chunk c
      if (lane < SP) ck_beta[(size_t)(cbase + c) * SP + lane] = alive ? b : 0.0;
      if (X.ck_beta_exp && lane == 0) X.ck_beta_exp[cbase + c] = Bx;
      if (c == 0) break;
      cur = nxt;
      if (c - 1 >= 1) fetch(g0 + c - 1, nxt);
      if (alive) alive = apply(cur, b, Bx);
    }
  }
}

}  // namespace jhp
