// kernels_fast.cuh — scaled linear-space forward/backward kernels (the production E-step and scoring path).
//
// Scope: first-order transitions (m == 1), no silent states, S <= 32 states, emission table a^(Kmax+1) x G doubles
// <= 64 KB.  Everything else runs on the reference-order kernels of kernels_exact.cuh.
//
// Why a different formulation keeps the reference's results.  The reference evaluates every DP cell as an n-ary
// log-sum-exp (one exp per edge, one log per context; HigherOrderHMM.java:371-405, 531-570).  The same quantities
// are computed here in probability space with an EXACT power-of-two rescaling per position (the scale is 2^-e,
// e = exponent of the largest entry, so rescaling introduces no rounding at all); one log per SEQUENCE turns the
// result back into a log-likelihood.  Per position a state needs S fused multiply-adds instead of S exp.  Results
// agree with the log-space recursion to ~1e-13 relative (tests/test_gpu_parity.py checks both formulations);
// the 1e-9 / 1e-6 gates of BASELINE.json are met with 4-7 digits of margin.  If the scaled recursion loses all mass
// (possible only with zero-probability parameters and > 2^1022 dynamic range between lineages) the sequence is
// handed to the log-space kernels on the GPU; there is no CPU fallback.
//
// Mapping: one group of G lanes per sequence (G = 2,4,8,16,32 >= S; 32/G sequences per warp), lane = state.
//   forward : lane j holds column T[:,j] in registers; alpha is broadcast through shared memory with 128-bit
//             loads; per position G FMAs per lane.  Scaled rows go to an HBM scratch (G*8 bytes per position).
//   backward: lane i holds row T[i,:] and the expected-count accumulators O[i,:] in registers; the broadcast
//             vector b[j] = e_j(x_l)*beta_{l+1}[j] feeds BOTH the beta recursion and the rank-1 count update
//             O[i][j] += alpha_l[i]*g_l*b[j]  (2 FMAs per loaded value, 4 per 128-bit load).
//             xi_l(i,j) = T[i][j]*O-term, so T is multiplied in once per kernel, not per position.
//   emission counts: gamma_l(j) = alpha_{l+1}[j]*beta_{l+1}[j]*scale into a per-CTA shared table [context hash][state].
#pragma once
#include <type_traits>

#include "common.cuh"
#include "kernels_exact.cuh"

struct jh_dataset;
template <typename T>
struct DevBuf;

namespace jhf {

// Copies of the global emission-count table [H][G].  The fp64 reductions of a position go to one copy, chosen by the warp
// slot: the copies spread them over the L2 slices, and — what matters for small tables — over ADDRESSES: the L2 serialises
// reductions to one address, and a 16-state model with first-order emissions (256 cells) on 16 copies spent two thirds of its
// E-step waiting for them (same kernel, 4096 cells per copy: 11 ms instead of 37).  The number of copies is a power of two
// chosen so that all copies together have about 2^19 cells (4 MB, L2 resident), at least 16, at most 2048.
constexpr int EC_REPLICAS_MIN = 16, EC_REPLICAS_MAX = 2048;
constexpr int64_t EC_CELLS_TARGET = (int64_t)1 << 19;
inline int ec_replicas_for(int64_t cells) {
  int r = EC_REPLICAS_MIN;
  while (r < EC_REPLICAS_MAX && (int64_t)r * cells < EC_CELLS_TARGET) r *= 2;
  return r;
}

using jhx::ResReader;

struct FastDev {
  int S, G, a, Kmax, H;  // H = a^(Kmax+1) rows of the emission table
  int hpow, a_pow2, n_child, n_stats;
  int ec_mask;          // number of emission-count table copies - 1 (a power of two)
  double inv_a;
  const double *T;      // [G*G] T[i*G+j] = P(i -> j)
  const double *Tt;     // [G*G] transposed
  const double *pi;     // [G]
  const double *E;      // [H*G] E[h*G+j] = P(x | context) of state j for context hash h
  const double *fin;    // [G] 1.0 for final states
  const int *korder;    // [G] emission order per state (padding lanes: 0)
  const int *tstat;     // [G*G] statistics slot of edge i->j or -1
  const int *pistat;    // [G]
  const int *estat_tab; // [G] state_tab (offset into the emission statistics), -1 for padding lanes
  const int *estat_mod; // [G] a^(k+1)
};

struct FastTables {
  bool usable = false, e_global = false;  // e_global: emission table too large for shared memory (> 64 KB)
  int S = 0, G = 0, a = 0, Kmax = 0, H = 0, n_child = 0, n_cond = 0;
  double *d_T = nullptr, *d_Tt = nullptr, *d_pi = nullptr, *d_E = nullptr, *d_fin = nullptr;
  int *d_korder = nullptr, *d_tstat = nullptr, *d_pistat = nullptr, *d_estat_tab = nullptr, *d_estat_mod = nullptr;
  int *d_tmap = nullptr, *d_pimap = nullptr;  // child index p of edge i->j / start->j, -1 if absent
  bool sorted_children = false;               // every context lists its children in ascending state order (kernels_viterbi.cuh)
  double *d_vT = nullptr, *d_vpi = nullptr, *d_vfin = nullptr, *d_vE = nullptr;  // log-space dense tables of the lane-per-sequence Viterbi
  int *d_fexp = nullptr;                      // per-row exponent scratch
  double *d_ec = nullptr;                     // [ec_replicas][H*G] global emission-count tables
  int ec_replicas = EC_REPLICAS_MIN;
  size_t ec_bytes() const { return sizeof(double) * (size_t)ec_replicas * H * G; }
  size_t fexp_n = 0;
  FastDev dev() const {
    FastDev F;
    F.S = S; F.G = G; F.a = a; F.Kmax = Kmax; F.H = H;
    int pw = 1;
    for (int i = 0; i < Kmax; i++) pw *= a;
    F.hpow = pw; F.a_pow2 = (a & (a - 1)) == 0; F.n_child = n_child; F.n_stats = n_child + n_cond * a;
    F.inv_a = 1.0 / a;
    F.ec_mask = ec_replicas - 1;
    F.T = d_T; F.Tt = d_Tt; F.pi = d_pi; F.E = d_E; F.fin = d_fin; F.korder = d_korder; F.tstat = d_tstat; F.pistat = d_pistat;
    F.estat_tab = d_estat_tab; F.estat_mod = d_estat_mod;
    return F;
  }
};

inline void destroy(FastTables &F) {
  void *ptrs[] = {F.d_T, F.d_Tt, F.d_pi, F.d_E, F.d_fin, F.d_korder, F.d_tstat, F.d_pistat, F.d_estat_tab, F.d_estat_mod, F.d_tmap, F.d_pimap, F.d_fexp, F.d_ec, F.d_vT, F.d_vpi, F.d_vfin, F.d_vE};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  F = FastTables();
}

template <typename T>
static int up(T **dst, const std::vector<T> &h, cudaStream_t s) {
  if (cudaMalloc((void **)dst, std::max<size_t>(h.size(), 1) * sizeof(T)) != cudaSuccess) {
    jh_set_error("out of device memory (fast tables)");
    return JH_ERR_NOMEM;
  }
  if (!h.empty()) JH_CUDA(cudaMemcpyAsync(*dst, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return JH_OK;
}

// Decide whether the model is in scope and build the index maps.
inline int build(FastTables &F, int S, int a, int m, bool has_silent, int Kmax, const std::vector<int> &lc_ctx_ptr,
                 const std::vector<int> &ctx_elem, const std::vector<int> &ctx_last, const std::vector<int> &elem_child_ptr,
                 const std::vector<int> &child_state, const std::vector<int> &child_stat, const std::vector<int> &state_order,
                 const std::vector<int> &state_tab, const std::vector<int> &state_mod, const std::vector<uint8_t> &state_final,
                 int n_child, int n_cond, cudaStream_t s) {
  F.usable = false;
  if (m != 1 || has_silent || S > 32) return JH_OK;
  if (lc_ctx_ptr[1] - lc_ctx_ptr[0] != 1) return JH_OK;
  int G = 2;
  while (G < S) G *= 2;
  int64_t H = a;
  for (int i = 0; i < Kmax; i++) H *= a;
  if (H > 65536 || H * G * 8 > ((int64_t)16 << 20)) return JH_OK;  // context hashes are uint16 records
  F.e_global = H * G * 8 > 64 * 1024;  // the emission table stays in global memory (kernels with the EG flag)
  std::vector<int> tmap((size_t)G * G, -1), pimap(G, -1), tstat((size_t)G * G, -1), pistat(G, -1), korder(G, 0), etab(G, -1), emod(G, 1);
  std::vector<double> fin(G, 0.0);
  std::vector<char> seen(S, 0);
  int e0 = ctx_elem[lc_ctx_ptr[0]];
  bool sorted = true;
  for (int p = elem_child_ptr[e0]; p < elem_child_ptr[e0 + 1]; p++) {
    pimap[child_state[p]] = p;
    pistat[child_state[p]] = child_stat[p];
    if (p > elem_child_ptr[e0] && child_state[p] <= child_state[p - 1]) sorted = false;
  }
  for (int c = lc_ctx_ptr[1]; c < lc_ctx_ptr[2]; c++) {
    int i = ctx_last[c], e = ctx_elem[c];
    if (i < 0 || i >= S || seen[i]) return JH_OK;  // two contexts for one state: not a plain first-order chain
    seen[i] = 1;
    for (int p = elem_child_ptr[e]; p < elem_child_ptr[e + 1]; p++) {
      tmap[(size_t)i * G + child_state[p]] = p;
      tstat[(size_t)i * G + child_state[p]] = child_stat[p];
      if (p > elem_child_ptr[e] && child_state[p] <= child_state[p - 1]) sorted = false;
    }
  }
  for (int j = 0; j < S; j++) {
    korder[j] = state_order[j];
    etab[j] = state_tab[j];
    emod[j] = state_mod[j];
    fin[j] = state_final[j] ? 1.0 : 0.0;
  }
  F.S = S; F.G = G; F.a = a; F.Kmax = Kmax; F.H = (int)H; F.n_child = n_child; F.n_cond = n_cond;
  int r;
  if ((r = up(&F.d_tmap, tmap, s)) || (r = up(&F.d_pimap, pimap, s)) || (r = up(&F.d_tstat, tstat, s)) || (r = up(&F.d_pistat, pistat, s)) ||
      (r = up(&F.d_korder, korder, s)) || (r = up(&F.d_estat_tab, etab, s)) || (r = up(&F.d_estat_mod, emod, s)) || (r = up(&F.d_fin, fin, s)))
    return r;
  std::vector<double> z((size_t)std::max<int64_t>(H * G, (int64_t)G * G), 0.0);
  std::vector<double> zT((size_t)G * G, 0.0), zpi(G, 0.0), zE((size_t)H * G, 0.0);
  if ((r = up(&F.d_T, zT, s)) || (r = up(&F.d_Tt, zT, s)) || (r = up(&F.d_pi, zpi, s)) || (r = up(&F.d_E, zE, s))) return r;
  F.ec_replicas = ec_replicas_for((int64_t)H * G);
  std::vector<double> zEC((size_t)F.ec_replicas * H * G, 0.0);
  if ((r = up(&F.d_ec, zEC, s))) return r;
  if ((r = up(&F.d_vT, zT, s)) || (r = up(&F.d_vpi, zpi, s)) || (r = up(&F.d_vfin, zpi, s)) || (r = up(&F.d_vE, zE, s))) return r;
  F.sorted_children = sorted;
  F.usable = true;
  return JH_OK;
}

// linear-space tables from the log-space score tables (after every parameter update)
__global__ void k_fast_prepare(FastDev F, const int *tmap, const int *pimap, const double *tscore, const double *escore, double *T,
                               double *Tt, double *pi, double *E) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int G = F.G;
  if (idx < G * G) {
    int i = idx / G, j = idx % G, p = tmap[idx];
    double v = p >= 0 ? exp(tscore[p]) : 0.0;
    T[i * G + j] = v;
    Tt[j * G + i] = v;
  }
  if (idx < G) pi[idx] = pimap[idx] >= 0 ? exp(tscore[pimap[idx]]) : 0.0;
  if (idx < F.H * G) {
    int h = idx / G, j = idx % G;
    E[idx] = F.estat_tab[j] >= 0 ? exp(escore[F.estat_tab[j] + h % F.estat_mod[j]]) : 0.0;
  }
}

inline int prepare(FastTables &F, const DevModel &M, cudaStream_t s) {
  if (!F.usable) return JH_OK;
  int n = std::max(F.G * F.G, F.H * F.G);
  k_fast_prepare<<<(n + 255) / 256, 256, 0, s>>>(F.dev(), F.d_tmap, F.d_pimap, M.tscore, M.escore, F.d_T, F.d_Tt, F.d_pi, F.d_E);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

// ------------------------------------------------------------------------------------------------
template <int G>
struct FGroup {
  int lane, lead;
  unsigned mask;
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};
template <int G>
__device__ __forceinline__ FGroup<G> fgroup() {
  FGroup<G> g;
  int wl = threadIdx.x & 31;
  g.lane = wl % G;
  g.lead = wl - g.lane;
  g.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << g.lead);
  return g;
}

__device__ __forceinline__ double pow2i(int k) {  // 2^k, exact; 0 below the normal range
  if (k < -1022) return 0.0;
  if (k > 1023) return INFINITY;
  return __hiloint2double((1023 + k) << 20, 0);
}

// Exact rescale: divide by 2^e with e = exponent of the group maximum (branch free).
// Returns false when the maximum is zero/denormal or not finite; the value is then left to die as 0 / NaN and the
// caller's sticky flag (or the final finite check) routes the sequence to the log-space kernels.
template <int G>
__device__ __forceinline__ bool rescale(double &v, int &expo, unsigned mask) {
  unsigned hi = (unsigned)__double2hiint(v);
  unsigned mx = __reduce_max_sync(mask, hi);
  int be = (int)(mx >> 20);
  v *= __hiloint2double((2046 - be) << 20, 0);  // 2^-(be-1023); be == 0 keeps zeros zero
  expo += be - 1023;
  return (unsigned)(be - 1) < 0x7fdu;
}

template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
  for (int d = G / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(mask, v, d);
  return v;
}

__device__ __forceinline__ int hmod_of(const FastDev &F, int h) { return F.a_pow2 ? (h & (F.H - 1)) : (h % F.H); }

// dot product of the broadcast vector (shared memory, 128-bit loads) with a register-resident row/column;
// 8 independent accumulators so that the DFMA chain never waits on its own latency
template <int G>
__device__ __forceinline__ double bcast_dot(const double *vec, const double (&R)[G]) {
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  if (G >= 8) {
#pragma unroll
    for (int i = 0; i < G; i += 8) {
      double2 x = *reinterpret_cast<const double2 *>(vec + i);
      double2 y = *reinterpret_cast<const double2 *>(vec + i + 2);
      double2 z = *reinterpret_cast<const double2 *>(vec + i + 4);
      double2 u = *reinterpret_cast<const double2 *>(vec + i + 6);
      a0 = fma(x.x, R[i], a0);
      a1 = fma(x.y, R[i + 1], a1);
      a2 = fma(y.x, R[i + 2], a2);
      a3 = fma(y.y, R[i + 3], a3);
      a4 = fma(z.x, R[i + 4], a4);
      a5 = fma(z.y, R[i + 5], a5);
      a6 = fma(u.x, R[i + 6], a6);
      a7 = fma(u.y, R[i + 7], a7);
    }
  } else if (G == 4) {
    double2 x = *reinterpret_cast<const double2 *>(vec);
    double2 y = *reinterpret_cast<const double2 *>(vec + 2);
    a0 = x.x * R[0];
    a1 = x.y * R[1];
    a2 = y.x * R[2];
    a3 = y.y * R[3];
  } else {
    double2 x = *reinterpret_cast<const double2 *>(vec);
    a0 = x.x * R[0];
    a1 = x.y * R[1];
  }
  return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// Scoring: forward recursion only, nothing stored (AbstractHMM.logProb semantics, value of bwd[0][0]).
template <int G>
__global__ void __launch_bounds__(256) k_fast_loglik(FastDev F, DevData D, unsigned long long *counter, double *out) {
  extern __shared__ double sm[];
  double *sT = sm, *sE = sT + G * G, *svec = sE + (size_t)F.H * G;
  for (int i = threadIdx.x; i < G * G; i += blockDim.x) sT[i] = F.T[i];
  for (int i = threadIdx.x; i < F.H * G; i += blockDim.x) sE[i] = F.E[i];
  __syncthreads();
  FGroup<G> g = fgroup<G>();
  double *vec = svec + (size_t)(threadIdx.x / G) * 2 * G;
  const int lane = g.lane;
  double Tcol[G];
#pragma unroll
  for (int i = 0; i < G; i++) Tcol[i] = sT[i * G + lane];
  const int kl = F.korder[lane];
  const double pil = F.pi[lane], finl = F.fin[lane];
  ResReader rd;
  for (;;) {
    long long it = 0;
    if (lane == 0) it = (long long)atomicAdd(counter, 1ull);
    it = __shfl_sync(g.mask, it, g.lead);
    if (it >= D.n_seq) break;
    int sid = D.order[it];
    int64_t off = D.offsets[sid], L = D.offsets[sid + 1] - off;
    rd.init(D.codes + off);
    int h = rd.get(0), A = 0;
    double e = 0 < kl ? F.inv_a : sE[h * G + lane];
    double f = pil * e;
    bool ok = rescale<G>(f, A, g.mask);
    for (int64_t l = 1; l < L; l++) {
      double *v = vec + (l & 1) * G;
      v[lane] = f;
      h = hmod_of(F, h * F.a + rd.get(l));
      e = l < kl ? F.inv_a : sE[h * G + lane];
      g.sync();
      f = bcast_dot<G>(v, Tcol) * e;
      ok &= rescale<G>(f, A, g.mask);
    }
    double sfin = group_sum<G>(ok ? f * finl : 0.0, g.mask);
    if (lane == 0) out[sid] = (ok && sfin > 0) ? log(sfin) + A * 0.6931471805599453094 : JH_NEG_INF;
    g.sync();
  }
}

// ------------------------------------------------------------------------------------------------
// Residue staging: a group loads G aligned 16-byte chunks (128-bit loads, one per lane) of its sequence into shared
// memory, turns them into context hashes h(p) = sum_{i<=K} a^i x[p-i] (uint16) once, and the position loops then
// read one broadcast uint16 per position in either direction — no per-position hash arithmetic.
template <int G>
struct HashTile {
  static constexpr int T = 16 * G;  // positions per tile
  uint8_t *raw;                     // [16 halo + T]
  uint16_t *hs;                     // [T]
  const uint8_t *base;              // 16-byte aligned address at or before the first residue
  int mis;                          // residue 0 sits at base[mis]
  int cur;                          // tile currently staged (-1: none)
  __device__ __forceinline__ void init(uint8_t *r, uint16_t *h, const uint8_t *seq) {
    raw = r; hs = h;
    mis = (int)((unsigned long long)seq & 15ull);
    base = seq - mis;
    cur = -1;
  }
  // stage tile tau: positions [tau*T - mis, (tau+1)*T - mis)
  template <bool POW2>
  __device__ __forceinline__ void load(const FastDev &F, int tau, int lane, unsigned mask, int ashift, int hmask, int K) {
    const int64_t start = (int64_t)tau * T - mis;  // position of raw[16]
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(base + (int64_t)tau * T) + lane);
    *reinterpret_cast<uint4 *>(raw + 16 + 16 * lane) = v;
    if (lane == 0) {
      uint4 hv = make_uint4(0, 0, 0, 0);
      if (start > 0) hv = __ldg(reinterpret_cast<const uint4 *>(base + (int64_t)tau * T) - 1);
      *reinterpret_cast<uint4 *>(raw) = hv;
    }
    __syncwarp(mask);
    // rolling hash over this lane's 16 positions, seeded with the K residues before them
    const int64_t p0 = start + 16 * lane;
    int h = 0;
    for (int i = K; i >= 1; i--) {
      int64_t q = p0 - i;
      int x = q >= 0 ? raw[16 + 16 * lane - i] : 0;
      h = POW2 ? ((h << ashift) | x) : (h * F.a + x);
    }
    if (POW2) {
      const unsigned w[4] = {v.x, v.y, v.z, v.w};
      unsigned out[8];
#pragma unroll
      for (int u = 0; u < 16; u++) {
        int x = (int)((w[u >> 2] >> ((u & 3) * 8)) & 0xffu);
        if (p0 + u < 0) x = 0;
        h = ((h << ashift) | x) & hmask;
        if (u & 1) out[u >> 1] |= (unsigned)h << 16;
        else out[u >> 1] = (unsigned)h;
      }
      uint4 *dst = reinterpret_cast<uint4 *>(hs + 16 * lane);
      dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
      dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    } else {  // general alphabets: rolled loop, runtime modulo
      for (int u = 0; u < 16; u++) {
        int x = p0 + u < 0 ? 0 : raw[16 + 16 * lane + u];
        h = (h * F.a + x) % F.H;
        hs[16 * lane + u] = (uint16_t)h;
      }
    }
    __syncwarp(mask);
    cur = tau;
  }
};

// Baum-Welch E-step: forward with stored scaled rows, backward with fused expected-count accumulation.
// stats layout: [transition statistics n_child | emission statistics n_cond*a | sum_n w_n*logP_n].
// gEC: global [H*G] emission-count table (fire-and-forget fp64 reductions), folded into stats by k_fast_fold_ec.
//
// NS sequences per group advance in lock step (NS = 1 or 2): they share the register-resident T row/column and the
// count accumulators O, and their independent dependency chains are interleaved instruction by instruction, which is
// what keeps the FP64 pipe fed with only 8 warps per SM (244 registers per thread).  Sequences arrive sorted by
// decreasing length, so sequence 0 of a group is the longer one: the head of its backward sweep and the tail of its
// forward sweep run alone (NA = 1), the common part runs joint (NA = NS).
template <int G, int NT, bool POW2, int NS>
__global__ void __launch_bounds__(NT, 1) k_fast_estep(FastDev F, DevData D, unsigned long long *counter, double *fwd_scratch,
                                                     int *fexp_scratch, int64_t rows, double *stats, double *gEC, int32_t *fallback) {
  extern __shared__ double sm[];
  constexpr int GPC = NT / G;  // groups per CTA
  constexpr int T = HashTile<G>::T;
  constexpr int DEPTH = 6, RING = 8, SLOTB = G * 8 + 16;
  const int HG = F.H * G;
  double *sT = sm, *sTt = sT + G * G, *sE = sTt + G * G, *sO = sE + HG, *sPi = sO + G * G, *svec = sPi + G;
  uint16_t *shs = reinterpret_cast<uint16_t *>(svec + (size_t)GPC * 2 * NS * G);
  uint8_t *sraw = reinterpret_cast<uint8_t *>(shs + (size_t)GPC * NS * T);
  uint8_t *sring = sraw + (size_t)GPC * NS * (T + 16);  // [GPC][NS][RING][SLOTB]
  for (int i = threadIdx.x; i < G * G; i += NT) { sT[i] = F.T[i]; sTt[i] = F.Tt[i]; sO[i] = 0; }
  for (int i = threadIdx.x; i < HG; i += NT) sE[i] = F.E[i];
  for (int i = threadIdx.x; i < G; i += NT) sPi[i] = 0;
  __syncthreads();
  FGroup<G> g = fgroup<G>();
  const int lane = g.lane;
  const int gid = threadIdx.x / G;
  double *const vec = svec + (size_t)gid * 2 * NS * G;  // [parity][NS][G]
  const int64_t slot0 = ((int64_t)blockIdx.x * GPC + gid) * NS;
  const int kl = F.korder[lane];
  const double pil = F.pi[lane], finl = F.fin[lane], inv_a = F.inv_a;
  const double *const sEl = sE + lane;
  double *const gECl = gEC + lane;
  int ashift = 0;
  while ((1 << ashift) < F.a) ashift++;
  const int hmask = F.H - 1;
  const bool lane0 = lane == 0;
  double O[G];
#pragma unroll
  for (int j = 0; j < G; j++) O[j] = 0;
  double SC = 0, score = 0;
  HashTile<G> tile[NS];
  double *fwd[NS];
  int *fexp[NS];
  char *ring[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    fwd[s] = fwd_scratch + (size_t)(slot0 + s) * rows * G + lane;
    fexp[s] = fexp_scratch + (size_t)(slot0 + s) * rows;
    ring[s] = reinterpret_cast<char *>(sring) + (size_t)(gid * NS + s) * RING * SLOTB;
  }
  for (;;) {
    long long it = 0;
    if (lane0) it = (long long)atomicAdd(counter, (unsigned long long)NS);
    it = __shfl_sync(g.mask, it, g.lead);
    if (it >= D.n_seq) break;
    int L[NS], sid[NS], mis[NS];
    double w[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
      const bool valid = it + s < D.n_seq;
      sid[s] = valid ? D.order[it + s] : -1;
      int64_t off = 0;
      L[s] = 0;
      w[s] = 0;
      if (valid) {
        off = D.offsets[sid[s]];
        L[s] = (int)(D.offsets[sid[s] + 1] - off);
        w[s] = D.weights ? D.weights[sid[s]] : 1.0;
      }
      tile[s].init(sraw + (size_t)(gid * NS + s) * (T + 16), shs + (size_t)(gid * NS + s) * T, D.codes + off);
      mis[s] = tile[s].mis;
    }
    const int Ljoint = L[NS - 1];  // the shortest one (order is by decreasing length)
    // ======================= forward =======================
    int A[NS];
    double f[NS];
    bool ok[NS];
    {
      double R[G];
#pragma unroll
      for (int i = 0; i < G; i++) R[i] = sT[i * G + lane];  // column T[:,lane]
      double e_next[NS];
      double *fp[NS];
      int *ep[NS];
#pragma unroll
      for (int s = 0; s < NS; s++) {
        A[s] = 0; f[s] = 0; ok[s] = true; e_next[s] = inv_a;
        fp[s] = fwd[s] + G;
        ep[s] = fexp[s] + 1;
        if (L[s] > 0) {
          tile[s].template load<POW2>(F, 0, lane, g.mask, ashift, hmask, F.Kmax);
          int h = tile[s].hs[mis[s]];
          double e = 0 < kl ? inv_a : sEl[h * G];
          f[s] = pil * e;
          ok[s] = rescale<G>(f[s], A[s], g.mask);
          *fp[s] = f[s];
          if (lane0) *ep[s] = A[s];
          if (L[s] > 1) {
            const int tp = 1 + mis[s];
            if ((tp & (T - 1)) == 0) tile[s].template load<POW2>(F, tp / T, lane, g.mask, ashift, hmask, F.Kmax);
            h = tile[s].hs[tp & (T - 1)];
            e_next[s] = 1 < kl ? inv_a : sEl[h * G];
          }
        }
      }
      int par = 0;
      auto fstep = [&](auto na_tag, int t) {  // rows t -> t+1 of the first NA sequences
        constexpr int NA = decltype(na_tag)::value;
        double *v = vec + par;
        par ^= NS * G;
        double e[NA];
#pragma unroll
        for (int s = 0; s < NA; s++) {
          v[s * G + lane] = f[s];
          e[s] = e_next[s];
        }
        g.sync();
#pragma unroll
        for (int s = 0; s < NA; s++) {
          if (t + 1 < L[s]) {  // emission of the next position, one iteration ahead (hash -> table -> register)
            const int tp = t + 1 + mis[s];
            if ((tp & (T - 1)) == 0) tile[s].template load<POW2>(F, tp / T, lane, g.mask, ashift, hmask, F.Kmax);
            const int h = tile[s].hs[tp & (T - 1)];
            e_next[s] = (t + 1) < kl ? inv_a : sEl[h * G];
          }
        }
#pragma unroll
        for (int s = 0; s < NA; s++) f[s] = bcast_dot<G>(v + s * G, R) * e[s];
#pragma unroll
        for (int s = 0; s < NA; s++) {
          ok[s] &= rescale<G>(f[s], A[s], g.mask);
          fp[s] += G;
          ep[s] += 1;
          *fp[s] = f[s];
          if (lane0) *ep[s] = A[s];
        }
      };
      for (int t = 1; t < Ljoint; t++) fstep(std::integral_constant<int, NS>(), t);
      if (NS > 1)
        for (int t = Ljoint < 1 ? 1 : Ljoint; t < L[0]; t++) fstep(std::integral_constant<int, 1>(), t);
    }
    double invS[NS];
    int AL[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
      double sfin = group_sum<G>((ok[s] && L[s] > 0) ? f[s] * finl : 0.0, g.mask);
      AL[s] = A[s];
      invS[s] = 0;
      if (L[s] > 0) {
        if (ok[s] && sfin > 0) {
          invS[s] = w[s] / sfin;
          if (lane0) score += w[s] * (log(sfin) + A[s] * 0.6931471805599453094);
        } else if (lane0) {  // the scaled recursion lost all mass: the reference-order kernels take this sequence
          int pos = atomicAdd(fallback, 1);
          fallback[1 + pos] = sid[s];
        }
      }
    }
    // ======================= backward + counts =======================
    // A sequence that fell back (or an empty slot) runs with weight 0: every posterior it produces is exactly 0.
    if (L[0] > 0) {
      double R[G];
#pragma unroll
      for (int j = 0; j < G; j++) R[j] = sTt[j * G + lane];  // row T[lane,:]
      double Bcur[NS], Fnext[NS], e_next[NS];
      int Bx[NS], Anext[NS], h_next[NS];
      // forward rows come back through a shared-memory ring filled by cp.async DEPTH positions ahead
      auto issue = [&](int t) {  // rows t - DEPTH of every sequence that has them
        const int r = t - DEPTH;
#pragma unroll
        for (int s = 0; s < NS; s++) {
          if (r >= 1 && r <= L[s] - 1) {
            char *sl = ring[s] + (size_t)(r & (RING - 1)) * SLOTB;
            cp_async8(sl + lane * 8, fwd[s] + (size_t)r * G);
            if (lane0) cp_async4(sl + G * 8, fexp[s] + r);
          }
        }
        cp_async_commit();
      };
#pragma unroll
      for (int s = 0; s < NS; s++) {
        // invS = w / P_scaled is folded into the backward vector: every posterior below is already weighted
        Bcur[s] = finl * invS[s];
        Bx[s] = 0;
        rescale<G>(Bcur[s], Bx[s], g.mask);
        Fnext[s] = f[s];
        Anext[s] = AL[s];
        h_next[s] = 0;
        e_next[s] = inv_a;
        if (L[s] > 0) {
          const int tp = L[s] - 1 + mis[s];
          if (tile[s].cur != tp / T) tile[s].template load<POW2>(F, tp / T, lane, g.mask, ashift, hmask, F.Kmax);
          h_next[s] = tile[s].hs[tp & (T - 1)];
          e_next[s] = (L[s] - 1) < kl ? inv_a : sEl[h_next[s] * G];
        }
      }
      for (int d = 0; d < DEPTH; d++) issue(L[0] - 1 + DEPTH - d);
      int par = 0;
      g.sync();
      auto bstep = [&](auto na_tag, int t) {  // position t >= 1 of the first NA sequences
        constexpr int NA = decltype(na_tag)::value;
        double *v = vec + par;
        par ^= NS * G;
#pragma unroll
        for (int s = 0; s < NA; s++) {
          const int k1 = max(-1023, min(1024, Anext[s] + Bx[s] - AL[s]));
          const double gam = Fnext[s] * Bcur[s] * __hiloint2double((1023 + k1) << 20, 0);  // posterior of state `lane` at t
          if (t >= kl) atomicAdd(gECl + h_next[s] * G, gam);
          v[s * G + lane] = e_next[s] * Bcur[s];
        }
        cp_async_wait<DEPTH - 1>();
        g.sync();
        double aw[NA], Fl[NA];
        int Al[NA];
#pragma unroll
        for (int s = 0; s < NA; s++) {
          const char *sl = ring[s] + (size_t)(t & (RING - 1)) * SLOTB;
          Fl[s] = *reinterpret_cast<const double *>(sl + lane * 8);
          Al[s] = *reinterpret_cast<const int *>(sl + G * 8);
          {  // next position: tile, hash, emission
            const int tp = t - 1 + mis[s];
            if (tile[s].cur != tp / T) tile[s].template load<POW2>(F, tp / T, lane, g.mask, ashift, hmask, F.Kmax);
            h_next[s] = tile[s].hs[tp & (T - 1)];
            e_next[s] = (t - 1) < kl ? inv_a : sEl[h_next[s] * G];
          }
          const int k2 = max(-1023, min(1024, Al[s] + Bx[s] - AL[s]));
          aw[s] = Fl[s] * __hiloint2double((1023 + k2) << 20, 0);
        }
        double b[NA][4];
#pragma unroll
        for (int s = 0; s < NA; s++) b[s][0] = b[s][1] = b[s][2] = b[s][3] = 0;
        if (G >= 4) {
#pragma unroll
          for (int j = 0; j < G; j += 4) {
            double2 x[NA], y[NA];
#pragma unroll
            for (int s = 0; s < NA; s++) {
              x[s] = *reinterpret_cast<const double2 *>(v + s * G + j);
              y[s] = *reinterpret_cast<const double2 *>(v + s * G + j + 2);
            }
#pragma unroll
            for (int s = 0; s < NA; s++) {
              b[s][0] = fma(R[j], x[s].x, b[s][0]);
              b[s][1] = fma(R[j + 1], x[s].y, b[s][1]);
              b[s][2] = fma(R[j + 2], y[s].x, b[s][2]);
              b[s][3] = fma(R[j + 3], y[s].y, b[s][3]);
            }
#pragma unroll
            for (int s = 0; s < NA; s++) {
              O[j] = fma(aw[s], x[s].x, O[j]);
              O[j + 1] = fma(aw[s], x[s].y, O[j + 1]);
              O[j + 2] = fma(aw[s], y[s].x, O[j + 2]);
              O[j + 3] = fma(aw[s], y[s].y, O[j + 3]);
            }
          }
        } else {
#pragma unroll
          for (int s = 0; s < NA; s++) {
            double2 x = *reinterpret_cast<const double2 *>(v + s * G);
            b[s][0] = fma(R[0], x.x, b[s][0]);
            b[s][1] = fma(R[1], x.y, b[s][1]);
            O[0] = fma(aw[s], x.x, O[0]);
            O[1] = fma(aw[s], x.y, O[1]);
          }
        }
#pragma unroll
        for (int s = 0; s < NA; s++) {
          Bcur[s] = (b[s][0] + b[s][1]) + (b[s][2] + b[s][3]);
          rescale<G>(Bcur[s], Bx[s], g.mask);  // cannot fail when the forward mass is positive; NaN reaches the finite check
          Fnext[s] = Fl[s];
          Anext[s] = Al[s];
        }
        issue(t);
      };
      if (NS > 1)
        for (int t = L[0] - 1; t >= (Ljoint < 1 ? 1 : Ljoint); --t) bstep(std::integral_constant<int, 1>(), t);
      for (int t = (NS > 1 ? Ljoint : L[0]) - 1; t >= 1; --t) bstep(std::integral_constant<int, NS>(), t);
      cp_async_wait<0>();
#pragma unroll
      for (int s = 0; s < NS; s++) {  // position 0: posterior of the first state = start-element statistic
        if (L[s] > 0) {
          const int k1 = max(-1023, min(1024, Anext[s] + Bx[s] - AL[s]));
          const double gam = Fnext[s] * Bcur[s] * __hiloint2double((1023 + k1) << 20, 0);
          if (0 >= kl) atomicAdd(gECl + h_next[s] * G, gam);
          SC += gam;
        }
      }
    }
    g.sync();
  }
  // ---------------- flush: T is multiplied in once; CTA-level reduction, then one atomic per cell ----------------
  {
#pragma unroll
    for (int j = 0; j < G; j++) {
      double v = O[j] * sTt[j * G + lane];
      if (v != 0) atomicAdd(&sO[lane * G + j], v);
    }
    if (SC != 0) atomicAdd(&sPi[lane], SC);
    if (lane0 && score != 0) atomicAdd(&stats[F.n_stats], score);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * G; i += NT)
    if (sO[i] != 0 && F.tstat[i] >= 0) atomicAdd(&stats[F.tstat[i]], sO[i]);
  for (int i = threadIdx.x; i < G; i += NT)
    if (sPi[i] != 0 && F.pistat[i] >= 0) atomicAdd(&stats[F.pistat[i]], sPi[i]);
}

// emission-count table [H][G] -> emission statistics (several context hashes share a cell when a state's order < Kmax)
// One warp per cell: the lanes stride over the copies, then a shuffle tree (a fixed summation order).
__global__ void k_fast_fold_ec(FastDev F, const double *gEC, double *stats) {
  const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= F.H * F.G) return;
  const int j = i % F.G, hh = i / F.G;
  double v = 0;
  for (int rep = lane; rep <= F.ec_mask; rep += 32) v += gEC[(size_t)rep * F.H * F.G + i];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if (lane == 0 && v != 0 && F.estat_tab[j] >= 0) atomicAdd(&stats[F.n_child + F.estat_tab[j] + hh % F.estat_mod[j]], v);
}

// NaN/Inf guard over the statistics block: a non-finite value makes the driver redo the E-step in log space.
__global__ void k_check_finite(const double *stats, int n, int *flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !isfinite(stats[i])) atomicExch(flag, 1);
}

// ------------------------------------------------------------------------------------------------
#define JHF_DISPATCH(G, ...)                                \
  switch (G) {                                              \
    case 2: { constexpr int GG = 2; __VA_ARGS__; } break;   \
    case 4: { constexpr int GG = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int GG = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int GG = 16; __VA_ARGS__; } break; \
    default: { constexpr int GG = 32; __VA_ARGS__; } break; \
  }

inline int run_loglik(FastTables &F, const DevModel &, const DevData &D, int64_t n_seq, int64_t, double *d_out, unsigned long long *counter,
                      int sm_count, cudaStream_t s) {
  JH_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
  const int NT = 256;
  size_t smem = ((size_t)F.G * F.G + (size_t)F.H * F.G + (size_t)(NT / F.G) * 2 * F.G) * sizeof(double);
  int r = JH_OK;
  JHF_DISPATCH(F.G, {
    auto kern = k_fast_loglik<GG>;
    if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) { jh_set_error("fast scoring kernel does not fit"); r = JH_ERR_UNSUPPORTED; }
    else {
      int64_t grid = std::min<int64_t>((int64_t)sm_count * per_sm, (n_seq + NT / GG - 1) / (NT / GG));
      kern<<<(int)std::max<int64_t>(grid, 1), NT, smem, s>>>(F.dev(), D, counter, d_out);
      JH_LAUNCHED();
    }
  });
  if (r) return r;
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

}  // namespace jhf
