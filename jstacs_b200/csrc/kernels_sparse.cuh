// kernels_sparse.cuh — scaled linear-space kernels for SPARSE first-order models with 33..256 states and for
// chromosome-scale sequences (BASELINE cfg5: 128 states, 4 children per state, 24 sequences of 10-250 Mb, forward-backward +
// posterior decoding).  Same exact power-of-two rescaling as kernels_fast.cuh; the transition structure is kept as
// padded in-/out-edge lists (ELL, at most IND edges per state) held in registers.
//
// Mapping: one WARP per sequence or chunk, SPL = S_pad/32 consecutive states per lane, the DP vector double-buffered in
// shared memory (gathered by edge list, written back with 128-bit stores), the context hash rolled warp-uniformly
// straight from the residues.
//
// A chromosome cannot be decoded the way the reference does it (two full (L+1) x C matrices: 256 GB for 250 Mb x 128
// states).  The forward/backward recursions are inherently sequential in the position, so the decoding is
// split in two exact phases:
//   1. k_sparse_pass1: per sequence one warp runs the forward recursion and a second warp the backward recursion over the
//      WHOLE sequence, concurrently, storing only a checkpoint vector every CK positions (1 KB per 4096 residues) — and the
//      log-likelihood.  This phase is latency bound (one dependent step per position).
//   2. k_sparse_decode_chunks: every chunk of CK positions is independent given its two checkpoints: a warp recomputes the
//      forward rows of the chunk into a scratch, runs the backward recursion through the chunk and writes the state with
//      the largest posterior per position (alpha_t[j] * beta_t[j], lowest index on ties, AbstractHMM.java:1125-1139).
//      Thousands of chunks run in parallel, whatever the number of sequences.
// Scoring alone is phase 1 without the backward warp and without checkpoints.
#pragma once
#include "kernels_exact.cuh"

namespace jhs {

using jhx::ResReader;

// copies of the emission-count table of the E-step: a power of two sized like jhf::ec_replicas_for (reductions to one address
// are serialised by the L2: few cells need many copies), at least 8
inline int ec_replicas_sparse_for(int64_t cells) {
  int r = 8;
  while (r < 2048 && (int64_t)r * cells < ((int64_t)1 << 19)) r *= 2;
  return r;
}

struct SparseDev {
  int S, SP, a, Kmax, H, hpow, a_pow2, ashift;  // ashift = log2(a) when a is a power of two
  int ec_mask;           // copies of the emission-count table - 1
  double inv_a;
  const int *in_src;     // [IND][SP] source state of in-edge k of state j (padding: j itself, weight 0)
  const double *in_T;    // [IND][SP] P(src -> j)
  const int *out_dst;    // [IND][SP] target state of out-edge k of state i
  const double *out_T;   // [IND][SP] P(i -> dst)
  const double *pi;      // [SP]
  const double *fin;     // [SP] 1.0 for final states
  const double *E;       // [H][SP] emission probability by context hash
  const int *korder;     // [SP]
};

struct SparseTables {
  bool usable = false, e_global = false;
  int S = 0, SP = 0, SPL = 0, IND = 0, a = 0, Kmax = 0, H = 0, ec_replicas = 8;
  size_t ec_bytes() const { return sizeof(double) * (size_t)ec_replicas * H * SP; }
  int *d_in_src = nullptr, *d_out_dst = nullptr, *d_in_p = nullptr, *d_out_p = nullptr, *d_pimap = nullptr, *d_korder = nullptr,
      *d_etab = nullptr, *d_emod = nullptr, *d_out_stat = nullptr, *d_pistat = nullptr;
  double *d_in_T = nullptr, *d_out_T = nullptr, *d_pi = nullptr, *d_fin = nullptr, *d_E = nullptr, *d_ec = nullptr;
  // log-space tables of the chromosome-scale Viterbi (kernels_vitlong.cuh) and band structure (kernels_band.cuh)
  double *d_out_lT = nullptr, *d_lpi = nullptr, *d_lfin = nullptr, *d_lE = nullptr;
  int *d_pi_rank = nullptr;
  int band_lo = 0, band_hi = -1;  // every edge i -> j has (j - i) mod S in [band_lo, band_hi] (cyclic, |.| <= SPL); hi < lo: not banded
  SparseDev dev() const {
    SparseDev F;
    F.S = S; F.SP = SP; F.a = a; F.Kmax = Kmax; F.H = H;
    int pw = 1;
    for (int i = 0; i < Kmax; i++) pw *= a;
    F.hpow = pw; F.a_pow2 = (a & (a - 1)) == 0; F.inv_a = 1.0 / a;
    F.ashift = 0;
    while ((1 << F.ashift) < a) F.ashift++;
    F.ec_mask = ec_replicas - 1;
    F.in_src = d_in_src; F.in_T = d_in_T; F.out_dst = d_out_dst; F.out_T = d_out_T; F.pi = d_pi; F.fin = d_fin; F.E = d_E; F.korder = d_korder;
    return F;
  }
};

inline void destroy(SparseTables &F) {
  void *ptrs[] = {F.d_in_src, F.d_out_dst, F.d_in_p, F.d_out_p, F.d_pimap, F.d_korder, F.d_etab, F.d_emod, F.d_out_stat, F.d_pistat,
                  F.d_in_T,   F.d_out_T,   F.d_pi,   F.d_fin,   F.d_E,     F.d_ec,     F.d_out_lT, F.d_lpi, F.d_lfin, F.d_lE, F.d_pi_rank};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  F = SparseTables();
}

template <typename T>
static int up(T **dst, const std::vector<T> &h, cudaStream_t s) {
  if (cudaMalloc((void **)dst, std::max<size_t>(h.size(), 1) * sizeof(T)) != cudaSuccess) {
    jh_set_error("out of device memory (sparse tables)");
    return JH_ERR_NOMEM;
  }
  if (!h.empty()) JH_CUDA(cudaMemcpyAsync(*dst, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return JH_OK;
}

// Scope test + edge lists: plain first-order chain, no silent states, up to 256 states, SPL * IND <= 32 (S/32 states per lane times
// the padded in-/out-degree): dense models up to 32 states, <= 16 edges per state up to 64 states, <= 8 up to 128, <= 4 up to 256.
inline int build(SparseTables &F, int S, int a, int m, bool has_silent, int Kmax, const std::vector<int> &lc_ctx_ptr, const std::vector<int> &ctx_elem,
                 const std::vector<int> &ctx_last, const std::vector<int> &elem_child_ptr, const std::vector<int> &child_state,
                 const std::vector<int> &child_stat, const std::vector<int> &state_order, const std::vector<int> &state_tab, const std::vector<int> &state_mod,
                 const std::vector<uint8_t> &state_final, cudaStream_t s) {
  F.usable = false;
  if (m != 1 || has_silent || S > 256) return JH_OK;
  if (lc_ctx_ptr[1] - lc_ctx_ptr[0] != 1) return JH_OK;
  const int SPL = S <= 32 ? 1 : S <= 64 ? 2 : S <= 128 ? 4 : 8, SP = 32 * SPL;
  int64_t H = a;
  for (int i = 0; i < Kmax; i++) H *= a;
  if (H * SP * 8 > ((int64_t)16 << 20)) return JH_OK;
  F.e_global = H * SP * 8 > 96 * 1024;
  std::vector<std::vector<std::pair<int, int>>> in(S), out(S);  // (other state, child index)
  std::vector<char> seen(S, 0);
  for (int c = lc_ctx_ptr[1]; c < lc_ctx_ptr[2]; c++) {
    int i = ctx_last[c], e = ctx_elem[c];
    if (i < 0 || i >= S || seen[i]) return JH_OK;
    seen[i] = 1;
    for (int p = elem_child_ptr[e]; p < elem_child_ptr[e + 1]; p++) {
      out[i].push_back({child_state[p], p});
      in[child_state[p]].push_back({i, p});
    }
  }
  size_t deg = 1;
  for (int j = 0; j < S; j++) deg = std::max(deg, std::max(in[j].size(), out[j].size()));
  if (deg > 32) return JH_OK;
  const int IND = deg <= 4 ? 4 : deg <= 8 ? 8 : deg <= 16 ? 16 : 32;
  if (SPL * IND > 32) return JH_OK;  // both edge lists of a lane have to fit into registers
  std::vector<int> in_src((size_t)IND * SP), out_dst((size_t)IND * SP), in_p((size_t)IND * SP, -1), out_p((size_t)IND * SP, -1), pimap(SP, -1),
      korder(SP, 0), etab(SP, -1), emod(SP, 1), out_stat((size_t)IND * SP, -1), pistat(SP, -1);
  std::vector<double> fin(SP, 0.0);
  for (int k = 0; k < IND; k++)
    for (int j = 0; j < SP; j++) {
      in_src[(size_t)k * SP + j] = j;
      out_dst[(size_t)k * SP + j] = j;
      if (j < S && k < (int)in[j].size()) { in_src[(size_t)k * SP + j] = in[j][k].first; in_p[(size_t)k * SP + j] = in[j][k].second; }
      if (j < S && k < (int)out[j].size()) {
        out_dst[(size_t)k * SP + j] = out[j][k].first;
        out_p[(size_t)k * SP + j] = out[j][k].second;
        out_stat[(size_t)k * SP + j] = child_stat[out[j][k].second];
      }
    }
  int e0 = ctx_elem[lc_ctx_ptr[0]];
  std::vector<int> pi_rank(SP, 0x7fffffff);
  std::vector<double> lfin(SP, -INFINITY);
  for (int p = elem_child_ptr[e0]; p < elem_child_ptr[e0 + 1]; p++) {
    pimap[child_state[p]] = p;
    pistat[child_state[p]] = child_stat[p];
    pi_rank[child_state[p]] = p - elem_child_ptr[e0];
  }
  for (int j = 0; j < S; j++) lfin[j] = state_final[j] ? 0.0 : -INFINITY;
  // band structure: all edges within SPL states (cyclically) of their source, S a multiple of 32
  {
    int lo = 0x7fffffff, hi = -0x7fffffff;
    bool banded = S == SP;
    for (int i = 0; i < S && banded; i++)
      for (auto &ed : out[i]) {
        int dlt = ((ed.first - i) % S + S) % S;
        if (dlt > S / 2) dlt -= S;
        if (dlt > SPL || dlt < -SPL) { banded = false; break; }
        lo = std::min(lo, dlt);
        hi = std::max(hi, dlt);
      }
    F.band_lo = banded && lo <= hi ? lo : 0;
    F.band_hi = banded && lo <= hi ? hi : -1;
  }
  for (int j = 0; j < S; j++) {
    korder[j] = state_order[j];
    etab[j] = state_tab[j];
    emod[j] = state_mod[j];
    fin[j] = state_final[j] ? 1.0 : 0.0;
  }
  F.S = S; F.SP = SP; F.SPL = SPL; F.IND = IND; F.a = a; F.Kmax = Kmax; F.H = (int)H;
  int r;
  F.ec_replicas = ec_replicas_sparse_for((int64_t)H * SP);
  std::vector<double> zE((size_t)H * SP, 0.0), zT((size_t)IND * SP, 0.0), zp(SP, 0.0), zEC((size_t)F.ec_replicas * H * SP, 0.0);
  if ((r = up(&F.d_in_src, in_src, s)) || (r = up(&F.d_out_dst, out_dst, s)) || (r = up(&F.d_in_p, in_p, s)) || (r = up(&F.d_out_p, out_p, s)) ||
      (r = up(&F.d_pimap, pimap, s)) || (r = up(&F.d_korder, korder, s)) || (r = up(&F.d_etab, etab, s)) || (r = up(&F.d_emod, emod, s)) ||
      (r = up(&F.d_fin, fin, s)) || (r = up(&F.d_in_T, zT, s)) || (r = up(&F.d_out_T, zT, s)) || (r = up(&F.d_pi, zp, s)) || (r = up(&F.d_E, zE, s)) ||
      (r = up(&F.d_out_stat, out_stat, s)) || (r = up(&F.d_pistat, pistat, s)) || (r = up(&F.d_ec, zEC, s)) || (r = up(&F.d_out_lT, zT, s)) ||
      (r = up(&F.d_lpi, zp, s)) || (r = up(&F.d_lfin, lfin, s)) || (r = up(&F.d_lE, zE, s)) || (r = up(&F.d_pi_rank, pi_rank, s)))
    return r;
  F.usable = true;
  return JH_OK;
}

// probabilities from the log-space score tables (after every parameter update)
__global__ void k_sparse_prepare(SparseDev F, int IND, const int *in_p, const int *out_p, const int *pimap, const int *etab, const int *emod,
                                 const double *tscore, const double *escore, double *in_T, double *out_T, double *pi, double *E) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < IND * F.SP) {
    in_T[idx] = in_p[idx] >= 0 ? exp(tscore[in_p[idx]]) : 0.0;
    out_T[idx] = out_p[idx] >= 0 ? exp(tscore[out_p[idx]]) : 0.0;
  }
  if (idx < F.SP) pi[idx] = pimap[idx] >= 0 ? exp(tscore[pimap[idx]]) : 0.0;
  if (idx < F.H * F.SP) {
    int h = idx / F.SP, j = idx % F.SP;
    E[idx] = etab[j] >= 0 ? exp(escore[etab[j] + h % emod[j]]) : 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int hmod_of(const SparseDev &F, int h) { return F.a_pow2 ? (h & (F.H - 1)) : (h % F.H); }
__device__ __forceinline__ int hash_at(const SparseDev &F, ResReader &rd, int64_t pos) {
  int h = 0, pw = 1;
  for (int i = 0; i <= F.Kmax; i++) {
    int64_t q = pos - i;
    if (q < 0) break;
    h += pw * rd.get(q);
    pw *= F.a;
  }
  return h;
}
__device__ __forceinline__ int hash_back(const SparseDev &F, ResReader &rd, int h, int64_t pos) {  // pos -> pos-1
  int64_t q = pos - 1 - F.Kmax;
  return (F.a_pow2 ? (h >> F.ashift) : (h / F.a)) + (q >= 0 ? F.hpow * rd.get(q) : 0);
}
__device__ __forceinline__ int hash_fwd(const SparseDev &F, ResReader &rd, int h, int64_t pos) {  // pos-1 -> pos
  return F.a_pow2 ? (((h << F.ashift) | rd.get(pos)) & (F.H - 1)) : ((h * F.a + rd.get(pos)) % F.H);
}
// The DP vector lives in shared memory "lane fastest": state j of lane j / SPL sits at slot (j % SPL) * 32 + j / SPL, so
// the lanes' accesses to their u-th state (and to a neighbour state at a fixed distance, the common banded case) are
// consecutive doubles: no bank conflicts.
template <int SPL>
__device__ __forceinline__ int slot_of(int j) { return (j % SPL) * 32 + j / SPL; }

// exact rescale of the warp's vector: divide by 2^e, e = exponent of the maximum; false when the maximum has left the
// safe range (80 binades above the denormals, see kernels_mma.cuh)
__device__ __forceinline__ double jhs_pow2(int k) {  // 2^k, 0 below the normal range, largest power above it
  k = max(-1023, min(1023, k));
  return __hiloint2double((1023 + k) << 20, 0);
}

template <int SPL>
__device__ __forceinline__ bool rescale_warp(double (&v)[SPL], long long &expo) {
  unsigned hi = (unsigned)__double2hiint(v[0]);
#pragma unroll
  for (int i = 1; i < SPL; i++) hi = max(hi, (unsigned)__double2hiint(v[i]));
  hi = __reduce_max_sync(0xffffffffu, hi);
  const int be = (int)(hi >> 20);
  const double f = __hiloint2double((2046 - be) << 20, 0);
#pragma unroll
  for (int i = 0; i < SPL; i++) v[i] *= f;
  expo += be - 1023;
  return (unsigned)(be - 81) < (0x7feu - 81u);
}

template <int SPL>
__device__ __forceinline__ void store_vec(double *dst, const double (&v)[SPL]) {  // dst: the lane's SPL consecutive slots
  if (SPL == 1) dst[0] = v[0];
  else {
#pragma unroll
    for (int u = 0; u + 1 < SPL; u += 2) *reinterpret_cast<double2 *>(dst + u) = make_double2(v[u], v[u + 1]);
  }
}
template <int SPL>
__device__ __forceinline__ void load_vec(const double *src, double (&v)[SPL]) {
  if (SPL == 1) v[0] = src[0];
  else {
#pragma unroll
    for (int u = 0; u + 1 < SPL; u += 2) {
      const double2 x = *reinterpret_cast<const double2 *>(src + u);
      v[u] = x.x;
      v[u + 1] = x.y;
    }
  }
}

template <int SPL>
__device__ __forceinline__ void put_shared(double *vec, int lane, const double (&v)[SPL]) {
#pragma unroll
  for (int u = 0; u < SPL; u++) vec[u * 32 + lane] = v[u];
}

// emission probabilities of the lane's states at position pos (context hash h)
template <int SPL>
__device__ __forceinline__ void emis(const SparseDev &F, const double *sE, int lane, double (&e)[SPL], int h, int64_t pos) {
  load_vec<SPL>(sE + (size_t)h * F.SP + SPL * lane, e);
  if (pos < F.Kmax) {
#pragma unroll
    for (int u = 0; u < SPL; u++)
      if (pos < F.korder[SPL * lane + u]) e[u] = F.inv_a;
  }
}

template <int SPL, int IND>
struct Edges {  // the lane's edge lists in registers: byte offsets into the shared vector and probabilities
  int off[IND][SPL];
  double w[IND][SPL];
  __device__ __forceinline__ void load(const int *idx, const double *T, int SP, int lane) {
#pragma unroll
    for (int k = 0; k < IND; k++)
#pragma unroll
      for (int u = 0; u < SPL; u++) {
        off[k][u] = slot_of<SPL>(idx[(size_t)k * SP + SPL * lane + u]) * 8;
        w[k][u] = T[(size_t)k * SP + SPL * lane + u];
      }
  }
  // out[u] = sum_k vec[idx[k][u]] * w[k][u]
  __device__ __forceinline__ void gather(const double *vec, double (&out)[SPL]) const {
    const char *base = reinterpret_cast<const char *>(vec);
#pragma unroll
    for (int u = 0; u < SPL; u++) {
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int k = 0; k < IND; k += 2) {
        a0 = fma(*reinterpret_cast<const double *>(base + off[k][u]), w[k][u], a0);
        a1 = fma(*reinterpret_cast<const double *>(base + off[k + 1][u]), w[k + 1][u], a1);
      }
      out[u] = a0 + a1;
    }
  }
  // the same gather, and O[k][u] += a[u] * vec[idx[k][u]] with the values it has loaded anyway (expected edge counts)
  __device__ __forceinline__ void gather_count(const double *vec, double (&out)[SPL], const double (&a)[SPL], double (&O)[IND][SPL]) const {
    const char *base = reinterpret_cast<const char *>(vec);
#pragma unroll
    for (int u = 0; u < SPL; u++) {
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int k = 0; k < IND; k += 2) {
        const double x0 = *reinterpret_cast<const double *>(base + off[k][u]), x1 = *reinterpret_cast<const double *>(base + off[k + 1][u]);
        a0 = fma(x0, w[k][u], a0);
        a1 = fma(x1, w[k + 1][u], a1);
        O[k][u] = fma(a[u], x0, O[k][u]);
        O[k + 1][u] = fma(a[u], x1, O[k + 1][u]);
      }
      out[u] = a0 + a1;
    }
  }
};

constexpr int RS_SPARSE = 8;  // positions between two rescales

// ---- residue streams ----------------------------------------------------------------------------------------------
// The position loops are latency bound (one dependent step per position and warp), so the residues are pulled through
// registers 16 at a time (one 128-bit load per block, the next block prefetched) and extracted with compile-time shifts;
// all index arithmetic inside a sequence is 32 bit (Jstacs sequence lengths are ints).
struct SeqBlocks {
  const uint4 *base;  // 16-byte aligned address at or before residue 0
  int mis;            // residue t sits at byte t + mis
  __device__ __forceinline__ void init(const uint8_t *seq) {
    mis = (int)((unsigned long long)seq & 15ull);
    base = reinterpret_cast<const uint4 *>(seq - mis);
  }
  __device__ __forceinline__ uint4 load(int blk) const { return blk >= 0 ? __ldg(base + blk) : make_uint4(0, 0, 0, 0); }
};

// step(t, x[t]) for t = ta .. tb ascending.  Whole 16-byte blocks run word by word (the four words rotate through one
// register, the four bytes of a word are unrolled), so the loop body stays small enough for the instruction cache — a
// warp that is alone on its SM cannot hide instruction fetches; the ragged head and tail go position by position.
template <typename Step>
__device__ __forceinline__ void for_residues_up(const SeqBlocks &B, int ta, int tb, Step step) {
  if (tb < ta) return;
  const uint8_t *bytes = reinterpret_cast<const uint8_t *>(B.base) + B.mis;  // residue t = bytes[t]
  int t = ta;
  const int head_end = min(tb + 1, ((ta + B.mis + 15) & ~15) - B.mis);  // first block boundary at or after ta
  for (; t < head_end; t++) step(t, (int)__ldg(bytes + t));
  int blk = (t + B.mis) >> 4;
  uint4 cur = B.load(blk);
  for (; t + 15 <= tb; t += 16, blk++) {
    const uint4 nxt = B.load(blk + 1);  // the code buffer is padded: one block past the end is readable
    unsigned w0 = cur.x, w1 = cur.y, w2 = cur.z, w3 = cur.w;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      step(t + 4 * k, (int)(w0 & 0xffu));
      step(t + 4 * k + 1, (int)((w0 >> 8) & 0xffu));
      step(t + 4 * k + 2, (int)((w0 >> 16) & 0xffu));
      step(t + 4 * k + 3, (int)(w0 >> 24));
      w0 = w1; w1 = w2; w2 = w3;
    }
    cur = nxt;
  }
  for (; t <= tb; t++) step(t, (int)__ldg(bytes + t));
}
// step(p, x[p]) for p = pb .. pa descending; positions below 0 deliver 0
template <typename Step>
__device__ __forceinline__ void for_residues_down(const SeqBlocks &B, int pb, int pa, Step step) {
  if (pb < pa) return;
  const uint8_t *bytes = reinterpret_cast<const uint8_t *>(B.base) + B.mis;
  int p = pb;
  const int head_end = max(pa - 1, max(-1, ((pb + B.mis + 1) & ~15) - B.mis - 1));  // last position of the block below, >= -1
  for (; p > head_end; p--) step(p, p >= 0 ? (int)__ldg(bytes + p) : 0);
  if (p >= 15) {
    int blk = (p + B.mis) >> 4;
    uint4 cur = B.load(blk);
    for (; p - 15 >= pa && p - 15 >= 0; p -= 16, blk--) {
      const uint4 nxt = B.load(blk - 1);
      unsigned w0 = cur.w, w1 = cur.z, w2 = cur.y, w3 = cur.x;
#pragma unroll 1
      for (int k = 0; k < 4; k++) {
        step(p - 4 * k, (int)(w0 >> 24));
        step(p - 4 * k - 1, (int)((w0 >> 16) & 0xffu));
        step(p - 4 * k - 2, (int)((w0 >> 8) & 0xffu));
        step(p - 4 * k - 3, (int)(w0 & 0xffu));
        w0 = w1; w1 = w2; w2 = w3;
      }
      cur = nxt;
    }
  }
  for (; p >= pa; p--) step(p, p >= 0 ? (int)__ldg(bytes + p) : 0);
}
__device__ __forceinline__ int roll_up(const SparseDev &F, int h, int x) {  // h(t-1), x[t] -> h(t)
  return F.a_pow2 ? (((h << F.ashift) | x) & (F.H - 1)) : ((h * F.a + x) % F.H);
}
__device__ __forceinline__ int roll_down(const SparseDev &F, int h, int x) {  // h(t), x[t-1-K] (0 before the start) -> h(t-1)
  return (F.a_pow2 ? (h >> F.ashift) : (h / F.a)) + F.hpow * x;
}

// Phase 1 / scoring.  Work item w: sequence w>>1 ... direction w&1 when both directions run (n_dir == 2), else forward only.
// ck_alpha / ck_beta: [chunk_ptr[seq] + c][SP]; loglik[seq]; fallback: sequences the scaled recursion could not finish.
struct Pass1Extra {        // E-step mode (all pointers null otherwise)
  long long *ck_alpha_exp;   // [n_chunks+1] exponent of each alpha checkpoint
  long long *ck_beta_exp;    // [n_chunks+1] exponent of each beta checkpoint
  double *seq_sfin;          // [n_seq] P(x) = seq_sfin * 2^seq_AL, 0 for sequences the scaled recursion could not finish
  long long *seq_AL;         // [n_seq]
  const int32_t *list;       // sequences to process (nullptr: all, in D.order)
  int64_t n_list;
  double *score;             // += sum_n w_n log P(x_n) over the processed sequences
  // scoring that meets in the middle (kernels_band.cuh): the forward warp stops at position L/2 and the backward warp at
  // the same position; both leave their vector, its exponent and their "ok" flag here, indexed by (list position, direction)
  double *meet_vec;          // [n_list][2][SP]
  long long *meet_exp;       // [n_list][2]
  int32_t *meet_ok;          // [n_list][2]
};

// EG (all kernels below): the emission table (> 96 KB: high emission orders) is read from global memory (L1/L2) instead of shared memory
template <int SPL, int IND, bool EG>
__global__ void __launch_bounds__(128) k_sparse_pass1(SparseDev F, DevData D, int n_dir, int CK, const int64_t *chunk_ptr, double *ck_alpha,
                                                      double *ck_beta, unsigned long long *counter, double *loglik, int32_t *fallback, Pass1Extra X) {
  extern __shared__ double sm[];
  const int SP = F.SP;
  const double *sE = EG ? F.E : sm;
  double *vec = sm + (EG ? 0 : (size_t)F.H * SP) + (size_t)(threadIdx.x >> 5) * 2 * SP;  // two buffers per warp
  if (!EG)
    for (int i = threadIdx.x; i < F.H * SP; i += blockDim.x) sm[i] = F.E[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  ResReader rd;
  SeqBlocks blocks;
  for (;;) {
    long long it = 0;
    if (lane == 0) it = (long long)atomicAdd(counter, 1ull);
    it = __shfl_sync(FULL, it, 0);
    if (it >= (X.list ? X.n_list : D.n_seq) * n_dir) break;
    const int sid = X.list ? X.list[it / n_dir] : D.order[it / n_dir];
    const bool backward = n_dir == 2 && (it & 1);
    const int64_t off = D.offsets[sid];
    const int L = (int)(D.offsets[sid + 1] - off);
    rd.init(D.codes + off);
    blocks.init(D.codes + off);
    const int64_t cbase = chunk_ptr ? chunk_ptr[sid] : 0;
    if (!backward) {
      Edges<SPL, IND> G;
      G.load(F.in_src, F.in_T, SP, lane);
      double f[SPL];
      long long A = 0;
      bool ok = true;
      int h = hash_at(F, rd, 0);
      {
        double e[SPL];
        emis<SPL>(F, sE, lane, e, hmod_of(F, h), 0);
#pragma unroll
        for (int u = 0; u < SPL; u++) f[u] = F.pi[SPL * lane + u] * e[u];
      }
      ok = rescale_warp<SPL>(f, A);
      int par = 0, ck_left = CK;  // steps until the next checkpoint (the vector BEFORE chunk c = alpha_{c*CK-1}, stored in step t = c*CK)
      double *ck_dst = ck_alpha ? ck_alpha + (size_t)(cbase + 1) * SP + SPL * lane : nullptr;
      long long *cke_dst = X.ck_alpha_exp ? X.ck_alpha_exp + cbase + 1 : nullptr;
      auto step = [&](int t, int x) {  // alpha_{t-1} -> alpha_t
        h = roll_up(F, h, x);
        double e[SPL];
        emis<SPL>(F, sE, lane, e, h, t);  // issued first: its latency hides behind the exchange of the vector
        double *v = vec + par;
        par ^= SP;
        put_shared<SPL>(v, lane, f);
        if (ck_dst && --ck_left == 0) {
          store_vec<SPL>(ck_dst, f);
          ck_dst += SP;
          ck_left = CK;
          if (cke_dst) {
            if (lane == 0) *cke_dst = A;
            cke_dst++;
          }
        }
        __syncwarp();
        G.gather(v, f);
#pragma unroll
        for (int u = 0; u < SPL; u++) f[u] *= e[u];
        if ((t & (RS_SPARSE - 1)) == 0) ok &= rescale_warp<SPL>(f, A);
      };
      for_residues_up(blocks, 1, L - 1, step);
      ok &= rescale_warp<SPL>(f, A);
      double sfin = 0;
#pragma unroll
      for (int u = 0; u < SPL; u++) sfin += f[u] * F.fin[SPL * lane + u];
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) sfin += __shfl_xor_sync(FULL, sfin, d);
      if (lane == 0) {
        const bool good = ok && sfin > 0;
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
      Edges<SPL, IND> G;
      G.load(F.out_dst, F.out_T, SP, lane);
      double b[SPL], e[SPL];
      long long Bx = 0;
#pragma unroll
      for (int u = 0; u < SPL; u++) b[u] = F.fin[SPL * lane + u];
      int h = hash_at(F, rd, L - 1);
      int par = 0;
      emis<SPL>(F, sE, lane, e, hmod_of(F, h), L - 1);
      // checkpoints: beta at the last position of chunk c, t = (c+1)*CK - 1 < L - 1
      const int n_full = (L - 1) / CK;            // chunks that end before position L-1
      int ck_left = n_full > 0 ? (L - 1) - (n_full * CK - 1) : 0x7fffffff;  // steps until t reaches n_full*CK-1
      double *ck_dst = ck_beta + (size_t)(cbase + n_full - 1) * SP + SPL * lane;
      long long *cke_dst = X.ck_beta_exp ? X.ck_beta_exp + cbase + n_full - 1 : nullptr;
      const int K1 = F.Kmax + 1;
      auto step = [&](int p, int x) {  // t = p + K + 1: beta_t -> beta_{t-1}; x = x[t-1-K] completes the hash of t-1
        const int t = p + K1;
        if (ck_left == 0) {
          store_vec<SPL>(ck_dst, b);
          ck_dst -= SP;
          ck_left = CK;
          if (cke_dst) {
            if (lane == 0) *cke_dst = Bx;
            cke_dst--;
          }
        }
        --ck_left;
        double *v = vec + par;
        par ^= SP;
        double xb[SPL];
#pragma unroll
        for (int u = 0; u < SPL; u++) xb[u] = e[u] * b[u];
        put_shared<SPL>(v, lane, xb);
        h = roll_down(F, h, x);
        emis<SPL>(F, sE, lane, e, h, t - 1);  // for the next step, in flight during the exchange
        __syncwarp();
        G.gather(v, b);
        if ((t & (RS_SPARSE - 1)) == 0) rescale_warp<SPL>(b, Bx);
      };
      for_residues_down(blocks, L - 2 - F.Kmax, -F.Kmax, step);  // t = L-1 .. 1
    }
    __syncwarp();
  }
}

// Phase 1 for FEW long sequences (fewer work items than SMs): the recursion is one dependent step per position, so what
// counts is the latency of a step, not throughput.  One CTA of SP/32 warps per (sequence, direction), ONE STATE PER THREAD.
// Critical path of a step: shared-memory store of the vector -> CTA barrier -> IND loads -> multiply/FMA -> add; the
// emission probability of the step is folded into the edge weights before the barrier (off the path: the context hash is
// known one step ahead).  The power-of-two rescale is lazy: the maximum of the vector of a flagged step travels with the
// next exchange and the factor is applied to the vector after it (any power of two is exact; the growth per step is
// bounded), so a rescale costs no barrier of its own.  Aligned 16-residue blocks run as straight-line code (no branches:
// with four warps on the SM every taken branch is an exposed instruction fetch); blocks that contain a checkpoint, the
// ragged ends and alphabets that are not a power of two go step by step.
template <int IND, bool EG>
__global__ void __launch_bounds__(256) k_sparse_pass1_cta(SparseDev F, DevData D, int n_dir, int CK, const int64_t *chunk_ptr, double *ck_alpha,
                                                          double *ck_beta, double *loglik, int32_t *fallback, Pass1Extra X) {
  extern __shared__ double sm[];
  const int SP = F.SP, j = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;  // blockDim.x == SP
  const double *sE = EG ? F.E : sm;
  double *vbuf = sm + (EG ? 0 : (size_t)F.H * SP);          // [2][SP]
  unsigned *mbuf = reinterpret_cast<unsigned *>(vbuf + 2 * SP);     // [2][8] row maxima (high words) of the lazy rescale
  double *red = reinterpret_cast<double *>(mbuf + 16);              // [8]
  if (!EG)
    for (int i = threadIdx.x; i < F.H * SP; i += blockDim.x) sm[i] = F.E[i];
  __syncthreads();
  const unsigned FULL = 0xffffffffu;
  const int kord = F.korder[j], ash = F.ashift, Hm = F.H - 1, hpow = F.hpow, SPB = SP * 8;
  const char *const sEb = reinterpret_cast<const char *>(sE);
  ResReader rd;
  SeqBlocks blocks;
  // exact rescale by the exponent of the CTA-wide maximum (two barriers; used at the ends of a sequence only)
  auto cta_rescale = [&](double &v, long long &expo) {
    unsigned hi = __reduce_max_sync(FULL, (unsigned)__double2hiint(v));
    __syncthreads();
    if (lane == 0) mbuf[warp] = hi;
    __syncthreads();
    hi = mbuf[0];
    for (int w = 1; w < NW; w++) hi = max(hi, mbuf[w]);
    const int be = (int)(hi >> 20);
    v *= __hiloint2double((2046 - be) << 20, 0);
    expo += be - 1023;
    return (unsigned)(be - 81) < (0x7feu - 81u);
  };
  const long long items = (X.list ? X.n_list : D.n_seq) * n_dir;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const int sid = X.list ? X.list[it / n_dir] : D.order[it / n_dir];
    const bool backward = n_dir == 2 && (it & 1);
    const int64_t off = D.offsets[sid];
    const int L = (int)(D.offsets[sid + 1] - off);
    rd.init(D.codes + off);
    blocks.init(D.codes + off);
    const int mis = blocks.mis;
    const uint8_t *const bytes = reinterpret_cast<const uint8_t *>(blocks.base) + mis;  // residue t = bytes[t]
    const int64_t cbase = chunk_ptr ? chunk_ptr[sid] : 0;
    int eo[IND];
    double ew[IND];
    {
      const int *idx = backward ? F.out_dst : F.in_src;
      const double *T = backward ? F.out_T : F.in_T;
#pragma unroll
      for (int k = 0; k < IND; k++) {
        eo[k] = idx[(size_t)k * SP + j] * 8;
        ew[k] = T[(size_t)k * SP + j];
      }
    }
    auto gather = [&](const double *v, const double (&w)[IND]) {  // NA independent accumulator chains (dependent FMAs are the critical path)
      constexpr int NA = IND >= 16 ? 8 : IND >= 8 ? 4 : 2;
      const char *base = reinterpret_cast<const char *>(v);
      double acc[NA];
#pragma unroll
      for (int k = 0; k < NA; k++) acc[k] = *reinterpret_cast<const double *>(base + eo[k]) * w[k];
#pragma unroll
      for (int k = NA; k < IND; k++) acc[k % NA] = fma(*reinterpret_cast<const double *>(base + eo[k]), w[k], acc[k % NA]);
#pragma unroll
      for (int d = NA / 2; d >= 1; d >>= 1)
#pragma unroll
        for (int k = 0; k < d; k++) acc[k] += acc[k + d];
      return acc[0];
    };
    // factor of a published maximum; false when it has left the safe range
    auto lazy_scale = [&](const unsigned *mb, double &v, long long &expo) {
      unsigned hi = mb[0];
      for (int w = 1; w < NW; w++) hi = max(hi, mb[w]);
      const int be = (int)(hi >> 20);
      v *= __hiloint2double((2046 - be) << 20, 0);
      expo += be - 1023;
      return (unsigned)(be - 81) < (0x7feu - 81u);
    };
    auto publish = [&](unsigned *mb, double v) {
      const unsigned hi = __reduce_max_sync(FULL, (unsigned)__double2hiint(v));
      if (lane == 0) mb[warp] = hi;
    };
    bool pend = false;  // the maximum of the current vector is to be published with the next exchange
    if (!backward) {
      double f;
      long long A = 0;
      int h = hash_at(F, rd, 0);
      {
        double e = sE[(size_t)hmod_of(F, h) * SP + j];
        if (0 < kord) e = F.inv_a;
        f = F.pi[j] * e;
      }
      bool ok = cta_rescale(f, A);
      int ck_left = CK;
      double *ck_dst = ck_alpha ? ck_alpha + (size_t)(cbase + 1) * SP + j : nullptr;
      long long *cke_dst = X.ck_alpha_exp ? X.ck_alpha_exp + cbase + 1 : nullptr;
      // exchange buffers and rescale cadence follow the residue ADDRESS (t + mis), so that aligned blocks start in a known state
      auto step = [&](int t, int x) {  // alpha_{t-1} -> alpha_t
        h = roll_up(F, h, x);
        double e = *reinterpret_cast<const double *>(sEb + (size_t)h * SPB + j * 8);
        if (t < kord) e = F.inv_a;
        double we[IND];
#pragma unroll
        for (int k = 0; k < IND; k++) we[k] = ew[k] * e;
        const int pr = (t + mis) & 1;
        double *v = vbuf + pr * SP;
        unsigned *mb = mbuf + pr * 8;
        v[j] = f;
        if (pend) publish(mb, f);
        if (ck_dst && --ck_left == 0) {
          *ck_dst = f;
          ck_dst += SP;
          ck_left = CK;
          if (cke_dst) {
            if (j == 0) *cke_dst = A;
            cke_dst++;
          }
        }
        __syncthreads();
        f = gather(v, we);
        if (pend) ok &= lazy_scale(mb, f, A);
        pend = ((t + mis) & 7) == 6;
      };
      auto fast16 = [&](const uint4 blk) {  // 16 steps of an aligned block: positions >= Kmax, no checkpoint inside, pend false on entry and exit
        const unsigned wds[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const int x = (int)((wds[k >> 2] >> ((k & 3) * 8)) & 0xffu);
          h = ((h << ash) | x) & Hm;
          const double e = *reinterpret_cast<const double *>(sEb + (size_t)h * SPB + j * 8);
          double we[IND];
#pragma unroll
          for (int i = 0; i < IND; i++) we[i] = ew[i] * e;
          double *v = vbuf + (k & 1) * SP;
          v[j] = f;
          if (k == 7 || k == 15) publish(mbuf + (k & 1) * 8, f);
          __syncthreads();
          f = gather(v, we);
          if (k == 7 || k == 15) ok &= lazy_scale(mbuf + (k & 1) * 8, f, A);
        }
      };
      int t = 1;
      if (F.a_pow2) {
#pragma unroll 1
        for (; t <= L - 1 && ((((t + mis) & 15) != 0) || t <= F.Kmax); t++) step(t, (int)__ldg(bytes + t));
        int blk = (t + mis) >> 4;
        uint4 cur = t + 15 <= L - 1 ? blocks.load(blk) : make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (; t + 15 <= L - 1; t += 16, blk++) {
          const uint4 nxt = blocks.load(blk + 1);  // the code buffer is padded: one block past the end is readable
          if (ck_dst && ck_left <= 16) {
#pragma unroll 1
            for (int k = 0; k < 16; k++) step(t + k, (int)__ldg(bytes + t + k));
          } else {
            fast16(cur);
            ck_left -= 16;
          }
          cur = nxt;
        }
      }
#pragma unroll 1
      for (; t <= L - 1; t++) step(t, (int)__ldg(bytes + t));
      ok &= cta_rescale(f, A);
      double sfin = f * F.fin[j];
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) sfin += __shfl_xor_sync(FULL, sfin, d);
      __syncthreads();
      if (lane == 0) red[warp] = sfin;
      __syncthreads();
      if (j == 0) {
        sfin = 0;
        for (int w = 0; w < NW; w++) sfin += red[w];
        const bool good = ok && sfin > 0;
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
      // beta_{t-1}[i] = sum_k T[i][dst_k] (e_t[dst_k] beta_t[dst_k]): the exchanged vector is e_t * beta_t (one multiply on the
      // dependent chain; folding e into the edge weights instead costs IND loads + multiplies + register rotation per step
      // and measured 80 ns per position against 66 ns forward)
      double b = F.fin[j];
      long long Bx = 0;
      int h = hmod_of(F, hash_at(F, rd, L - 1));
      double e = sE[(size_t)h * SP + j];
      if (L - 1 < kord) e = F.inv_a;
      const int n_full = (L - 1) / CK;
      int ck_left = n_full > 0 ? (L - 1) - (n_full * CK - 1) : 0x7fffffff;
      double *ck_dst = ck_beta + (size_t)(cbase + n_full - 1) * SP + j;
      long long *cke_dst = X.ck_beta_exp ? X.ck_beta_exp + cbase + n_full - 1 : nullptr;
      const int K1 = F.Kmax + 1;
      auto step = [&](int p, int x) {  // t = p + K + 1: beta_t -> beta_{t-1}; x = x[t-1-K] completes the hash of t-1
        const int t = p + K1;
        if (ck_left == 0) {
          *ck_dst = b;
          ck_dst -= SP;
          ck_left = CK;
          if (cke_dst) {
            if (j == 0) *cke_dst = Bx;
            cke_dst--;
          }
        }
        --ck_left;
        const int pr = (p + mis) & 1;
        double *v = vbuf + pr * SP;
        unsigned *mb = mbuf + pr * 8;
        v[j] = e * b;
        if (pend) publish(mb, b);
        h = roll_down(F, h, x);
        e = *reinterpret_cast<const double *>(sEb + (size_t)h * SPB + j * 8);
        if (t - 1 < kord) e = F.inv_a;
        __syncthreads();
        b = gather(v, ew);
        if (pend) lazy_scale(mb, b, Bx);
        pend = ((p + mis) & 7) == 1;
      };
      auto fast16 = [&](const uint4 blk) {  // 16 steps down an aligned block: p >= 0, no checkpoint inside
        const unsigned wds[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
        for (int k = 15; k >= 0; k--) {
          const int x = (int)((wds[k >> 2] >> ((k & 3) * 8)) & 0xffu);
          double *v = vbuf + (k & 1) * SP;
          v[j] = e * b;
          if (k == 8 || k == 0) publish(mbuf + (k & 1) * 8, b);
          h = (h >> ash) + hpow * x;
          e = *reinterpret_cast<const double *>(sEb + (size_t)h * SPB + j * 8);
          __syncthreads();
          b = gather(v, ew);
          if (k == 8 || k == 0) lazy_scale(mbuf + (k & 1) * 8, b, Bx);
        }
      };
      int p = L - 2 - F.Kmax;
      const int pa = -F.Kmax;
      if (F.a_pow2) {
#pragma unroll 1
        for (; p >= pa && (((p + mis) & 15) != 15 || p < 15); p--) step(p, p >= 0 ? (int)__ldg(bytes + p) : 0);
        int blk = (p + mis) >> 4;
        uint4 cur = p >= 15 ? blocks.load(blk) : make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (; p >= 15; p -= 16, blk--) {  // (p + mis) & 15 == 15: the block is bytes[p-15 .. p]
          const uint4 nxt = blocks.load(blk - 1);
          if (ck_left <= 16) {
#pragma unroll 1
            for (int k = 0; k < 16; k++) step(p - k, (int)__ldg(bytes + p - k));
          } else {
            fast16(cur);
            ck_left -= 16;
          }
          cur = nxt;
        }
      }
#pragma unroll 1
      for (; p >= pa; p--) step(p, p >= 0 ? (int)__ldg(bytes + p) : 0);
    }
    __syncthreads();
  }
}

// Phase 2: posterior decoding of one chunk per warp.  chunk_seq / chunk_idx: the chunk list; rows: [warp slot][CK][SP].
template <int SPL, int IND, bool EG>
__global__ void __launch_bounds__(128) k_sparse_decode_chunks(SparseDev F, DevData D, int CK, int ROWS, int64_t n_chunks, const int32_t *chunk_seq,
                                                              const int32_t *chunk_idx, const int64_t *chunk_ptr, const double *ck_alpha,
                                                              const double *ck_beta, unsigned long long *counter, double *rows, PathOut paths) {
  extern __shared__ double sm[];
  const int SP = F.SP;
  const double *sE = EG ? F.E : sm;
  double *vec = sm + (EG ? 0 : (size_t)F.H * SP) + (size_t)(threadIdx.x >> 5) * 2 * SP;
  if (!EG)
    for (int i = threadIdx.x; i < F.H * SP; i += blockDim.x) sm[i] = F.E[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int64_t slot = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double *const R = rows + (size_t)slot * ROWS * SP + SPL * lane;  // ROWS = min(CK, longest sequence)
  ResReader rd;
  SeqBlocks blocks;
  Edges<SPL, IND> Gin, Gout;
  Gin.load(F.in_src, F.in_T, SP, lane);
  Gout.load(F.out_dst, F.out_T, SP, lane);
  const int K1 = F.Kmax + 1;
  for (;;) {
    long long g = 0;
    if (lane == 0) g = (long long)atomicAdd(counter, 1ull);
    g = __shfl_sync(FULL, g, 0);
    if (g >= n_chunks) break;
    const int sid = chunk_seq[g], c = chunk_idx[g];
    const int64_t off = D.offsets[sid];
    const int L = (int)(D.offsets[sid + 1] - off);
    const int t0 = c * CK, t1 = min(L, t0 + CK) - 1;
    const int64_t cb = chunk_ptr[sid] + c;
    rd.init(D.codes + off);
    blocks.init(D.codes + off);
    // ---- forward rows of the chunk ----
    double f[SPL];
    long long dummy = 0;
    int h = hmod_of(F, hash_at(F, rd, t0));
    int par = 0;
    {
      double e[SPL];
      emis<SPL>(F, sE, lane, e, h, t0);
      if (c == 0) {
#pragma unroll
        for (int u = 0; u < SPL; u++) f[u] = F.pi[SPL * lane + u] * e[u];
      } else {
        double *v = vec + par;
        par ^= SP;
        double a0[SPL];
        load_vec<SPL>(ck_alpha + (size_t)cb * SP + SPL * lane, a0);
        put_shared<SPL>(v, lane, a0);
        __syncwarp();
        Gin.gather(v, f);
#pragma unroll
        for (int u = 0; u < SPL; u++) f[u] *= e[u];
      }
    }
    rescale_warp<SPL>(f, dummy);
    store_vec<SPL>(R, f);
    double *rdst = R;
    auto fstep = [&](int t, int x) {
      h = roll_up(F, h, x);
      double e[SPL];
      emis<SPL>(F, sE, lane, e, h, t);
      double *v = vec + par;
      par ^= SP;
      put_shared<SPL>(v, lane, f);
      __syncwarp();
      Gin.gather(v, f);
#pragma unroll
      for (int u = 0; u < SPL; u++) f[u] *= e[u];
      if ((t & (RS_SPARSE - 1)) == 0) rescale_warp<SPL>(f, dummy);
      rdst += SP;
      store_vec<SPL>(rdst, f);
    };
    for_residues_up(blocks, t0 + 1, t1, fstep);
    // ---- backward through the chunk, decision per position ----
    double b[SPL], e[SPL], a[SPL];
    if (t1 == L - 1) {
#pragma unroll
      for (int u = 0; u < SPL; u++) b[u] = F.fin[SPL * lane + u];
    } else {
      load_vec<SPL>(ck_beta + (size_t)cb * SP + SPL * lane, b);
    }
    int hb = hash_at(F, rd, t1);
    emis<SPL>(F, sE, lane, e, hmod_of(F, hb), t1);
    load_vec<SPL>(rdst, a);
    auto decide = [&](int t) {  // largest alpha_t[j] * beta_t[j], lowest state index on ties
      double best = a[0] * b[0];
      int arg = SPL * lane;
#pragma unroll
      for (int u = 1; u < SPL; u++) {
        const double gm = a[u] * b[u];
        if (gm > best) { best = gm; arg = SPL * lane + u; }
      }
#pragma unroll
      for (int d = 1; d <= 16; d <<= 1) {
        const double ob = __shfl_xor_sync(FULL, best, d);
        const int oa = __shfl_xor_sync(FULL, arg, d);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
      }
      if (lane == 0) path_store(paths, off + t, arg);
    };
    auto bstep = [&](int p, int x) {  // t = p + K + 1 > t0: decision of position t, then beta_t -> beta_{t-1}
      const int t = p + K1;
      double *v = vec + par;
      par ^= SP;
      double xb[SPL];
#pragma unroll
      for (int u = 0; u < SPL; u++) xb[u] = e[u] * b[u];
      put_shared<SPL>(v, lane, xb);
      hb = roll_down(F, hb, x);
      emis<SPL>(F, sE, lane, e, hb, t - 1);
      decide(t);  // its shuffles overlap the exchange
      rdst -= SP;
      load_vec<SPL>(rdst, a);
      __syncwarp();
      Gout.gather(v, b);
      if ((t & (RS_SPARSE - 1)) == 0) rescale_warp<SPL>(b, dummy);
    };
    for_residues_down(blocks, t1 - K1, t0 + 1 - K1, bstep);  // t = t1 .. t0+1
    decide(t0);
    __syncwarp();
  }
}

// Baum-Welch E-step, one chunk per warp (the count analogue of k_sparse_decode_chunks).  A sequence of one chunk is done
// entirely here (forward from the start element, P(x), backward with counts); longer sequences come with the checkpoints,
// their exponents and P(x) = sfin * 2^AL from k_sparse_pass1.  Per position t of the chunk:
//   gamma_t[j]    = alpha_t[j] beta_t[j] 2^(A_t + B_t - AL) w / sfin           -> emission statistic [hash of t][j], start statistic
//   xi_t(i -> j)  = alpha_{t-1}[i] T[i][j] vb_t[j] 2^(A_{t-1} + B_t - AL) w / sfin,  vb_t = e_t * beta_t
// accumulated per out-edge in registers (T is multiplied in at the end), with the values the beta recursion gathers anyway.
// rows: [slot][ROWS][SP] doubles, rexp: [slot][ROWS] exponents relative to the chunk's base exponent (ROWS = min(CK, longest sequence)).
// gEC: [copies][H][SP] emission-count tables; stats: [transition | emission | score].

struct EstepTables {
  const int *out_stat;     // [IND][SP] statistics slot of out-edge k of state i, -1 for padding
  const int *pistat;       // [SP]
  int n_stats;             // index of the score entry
};

template <int SPL, int IND, bool EG>
__global__ void __launch_bounds__(128) k_sparse_estep_chunks(SparseDev F, EstepTables ET, DevData D, int CK, int ROWS, int64_t n_chunks, const int32_t *chunk_seq,
                                                             const int32_t *chunk_idx, const int64_t *chunk_ptr, const double *ck_alpha,
                                                             const double *ck_beta, const long long *ck_alpha_exp, const long long *ck_beta_exp,
                                                             const double *seq_sfin, const long long *seq_AL, unsigned long long *counter,
                                                             double *rows, int *rexp, double *stats, double *gEC, int32_t *fallback) {
  extern __shared__ double sm[];
  const int SP = F.SP;
  const double *sE = EG ? F.E : sm;
  double *vec = sm + (EG ? 0 : (size_t)F.H * SP) + (size_t)(threadIdx.x >> 5) * 2 * SP;
  if (!EG)
    for (int i = threadIdx.x; i < F.H * SP; i += blockDim.x) sm[i] = F.E[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int64_t slot = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double *const R = rows + (size_t)slot * ROWS * SP + SPL * lane;
  int *const RX = rexp + (size_t)slot * ROWS;
  double *const gECw = gEC + (size_t)(slot & F.ec_mask) * F.H * SP + SPL * lane;
  ResReader rd;
  SeqBlocks blocks;
  double O[IND][SPL], SC[SPL];
#pragma unroll
  for (int u = 0; u < SPL; u++) {
    SC[u] = 0;
#pragma unroll
    for (int k = 0; k < IND; k++) O[k][u] = 0;
  }
  double score = 0;
  const int K1 = F.Kmax + 1;
  for (;;) {
    long long g = 0;
    if (lane == 0) g = (long long)atomicAdd(counter, 1ull);
    g = __shfl_sync(FULL, g, 0);
    if (g >= n_chunks) break;
    const int sid = chunk_seq[g], c = chunk_idx[g];
    const int64_t off = D.offsets[sid];
    const int L = (int)(D.offsets[sid + 1] - off);
    const int t0 = c * CK, t1 = min(L, t0 + CK) - 1;
    const int64_t cb = chunk_ptr[sid] + c;
    const bool single = L <= CK;
    const double w = D.weights ? D.weights[sid] : 1.0;
    double sfin = single ? 0.0 : seq_sfin[sid];
    long long AL = single ? 0 : seq_AL[sid];
    if (!single && !(sfin > 0)) continue;  // pass 1 could not finish this sequence: the log-space kernels take it
    rd.init(D.codes + off);
    blocks.init(D.codes + off);
    // ---- forward rows of the chunk (exponents relative to A0) ----
    double f[SPL];
    long long A = (c == 0) ? 0 : ck_alpha_exp[cb];
    const long long A0 = A;
    bool ok = true;
    int h = hmod_of(F, hash_at(F, rd, t0));
    int par = 0;
    double a0[SPL];  // alpha_{t0-1} (the checkpoint), needed again for the first transition of the chunk
#pragma unroll
    for (int u = 0; u < SPL; u++) a0[u] = 0;
    {
      Edges<SPL, IND> Gin;
      Gin.load(F.in_src, F.in_T, SP, lane);
      {
        double e[SPL];
        emis<SPL>(F, sE, lane, e, h, t0);
        if (c == 0) {
#pragma unroll
          for (int u = 0; u < SPL; u++) f[u] = F.pi[SPL * lane + u] * e[u];
        } else {
          double *v = vec + par;
          par ^= SP;
          load_vec<SPL>(ck_alpha + (size_t)cb * SP + SPL * lane, a0);
          put_shared<SPL>(v, lane, a0);
          __syncwarp();
          Gin.gather(v, f);
#pragma unroll
          for (int u = 0; u < SPL; u++) f[u] *= e[u];
        }
      }
      ok &= rescale_warp<SPL>(f, A);
      store_vec<SPL>(R, f);
      if (lane == 0) RX[0] = (int)(A - A0);
      double *rdst = R;
      int *xdst = RX;
      auto fstep = [&](int t, int x) {
        h = roll_up(F, h, x);
        double e[SPL];
        emis<SPL>(F, sE, lane, e, h, t);
        double *v = vec + par;
        par ^= SP;
        put_shared<SPL>(v, lane, f);
        __syncwarp();
        Gin.gather(v, f);
#pragma unroll
        for (int u = 0; u < SPL; u++) f[u] *= e[u];
        if ((t & (RS_SPARSE - 1)) == 0) ok &= rescale_warp<SPL>(f, A);
        rdst += SP;
        store_vec<SPL>(rdst, f);
        xdst++;
        if (lane == 0) *xdst = (int)(A - A0);
      };
      for_residues_up(blocks, t0 + 1, t1, fstep);
    }
    if (single) {  // the whole sequence is this chunk: P(x) from the last row
      double s_ = 0;
#pragma unroll
      for (int u = 0; u < SPL; u++) s_ += f[u] * F.fin[SPL * lane + u];
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) s_ += __shfl_xor_sync(FULL, s_, d);
      sfin = s_;
      AL = A;
      if (!(ok && sfin > 0)) {
        if (lane == 0) {
          int pos = atomicAdd(fallback, 1);
          fallback[1 + pos] = sid;
        }
        continue;
      }
      if (lane == 0) score += w * (log(sfin) + (double)AL * 0.6931471805599453094);
    }
    const double winv = w / sfin;
    __syncwarp();
    // ---- backward through the chunk with counts ----
    Edges<SPL, IND> Gout;
    Gout.load(F.out_dst, F.out_T, SP, lane);
    double b[SPL], e[SPL], a[SPL];
    long long Bx = 0;
    if (t1 == L - 1) {
#pragma unroll
      for (int u = 0; u < SPL; u++) b[u] = F.fin[SPL * lane + u];
    } else {
      load_vec<SPL>(ck_beta + (size_t)cb * SP + SPL * lane, b);
      Bx = ck_beta_exp[cb];
    }
    int hb = hash_at(F, rd, t1);
    emis<SPL>(F, sE, lane, e, hmod_of(F, hb), t1);
    const double *rsrc = R + (size_t)(t1 - t0) * SP;
    const int *xsrc = RX + (t1 - t0);
    load_vec<SPL>(rsrc, a);
    long long At = A0 + *xsrc;
    // the rows below come back from L2 / HBM: rows t-1 and t-2 are kept in flight in registers (pre0, pre1), row t-3 is
    // requested at the start of step t
    double pre0[SPL], pre1[SPL];
    int pe0 = 0, pe1 = 0;
#pragma unroll
    for (int u = 0; u < SPL; u++) pre0[u] = pre1[u] = 0;
    if (t1 - 1 >= t0) { load_vec<SPL>(rsrc - SP, pre0); pe0 = xsrc[-1]; }
    if (t1 - 2 >= t0) { load_vec<SPL>(rsrc - 2 * SP, pre1); pe1 = xsrc[-2]; }
    rsrc -= 2 * SP;  // row of pre1
    xsrc -= 2;
    // one position: emission statistic of t, then (with the previous row ap / exponent Ap) the edge statistic and beta_{t-1}
    auto consume = [&](int t, const double (&ap)[SPL], long long Ap, bool recurse, int x) {
      const long long kg = At + Bx - AL, kx = Ap + Bx - AL;
      const double sg = jhs_pow2((int)max(-1100ll, min(1100ll, kg))) * winv, sx = jhs_pow2((int)max(-1100ll, min(1100ll, kx))) * winv;
      const int hm = hmod_of(F, hb);
#pragma unroll
      for (int u = 0; u < SPL; u++) {
        const double gm = a[u] * b[u] * sg;  // posterior of the lane's state u at position t
        if (t >= F.Kmax || t >= F.korder[SPL * lane + u]) atomicAdd(gECw + (size_t)hm * SP + u, gm);
        if (t == 0) SC[u] += gm;
      }
      if (t == 0) return;
      double *v = vec + par;
      par ^= SP;
      double xb[SPL], as[SPL];
#pragma unroll
      for (int u = 0; u < SPL; u++) {
        xb[u] = e[u] * b[u];
        as[u] = ap[u] * sx;
      }
      put_shared<SPL>(v, lane, xb);
      if (recurse) {
        hb = roll_down(F, hb, x);
        emis<SPL>(F, sE, lane, e, hb, t - 1);
      }
      __syncwarp();
      Gout.gather_count(v, b, as, O);
      if ((t & (RS_SPARSE - 1)) == 0) rescale_warp<SPL>(b, Bx);
    };
    auto bstep = [&](int p, int x) {  // t = p + K + 1 > t0
      const int t = p + K1;
      double ap[SPL];
#pragma unroll
      for (int u = 0; u < SPL; u++) {
        ap[u] = pre0[u];
        pre0[u] = pre1[u];
      }
      const long long Ap = A0 + pe0;
      pe0 = pe1;
      rsrc -= SP;
      xsrc--;
      if (t - 3 >= t0) {
        load_vec<SPL>(rsrc, pre1);
        pe1 = *xsrc;
      }
      consume(t, ap, Ap, true, x);
#pragma unroll
      for (int u = 0; u < SPL; u++) a[u] = ap[u];
      At = Ap;
    };
    for_residues_down(blocks, t1 - K1, t0 + 1 - K1, bstep);  // t = t1 .. t0+1
    consume(t0, a0, A0, false, 0);                           // t = t0: the transition out of the previous chunk (or the start)
    __syncwarp();
  }
  // ---- flush ----
  {
    Edges<SPL, IND> Gout;
    Gout.load(F.out_dst, F.out_T, SP, lane);
#pragma unroll
    for (int k = 0; k < IND; k++)
#pragma unroll
      for (int u = 0; u < SPL; u++) {
        const int st = ET.out_stat[(size_t)k * SP + SPL * lane + u];
        const double v = O[k][u] * Gout.w[k][u];
        if (st >= 0 && v != 0) atomicAdd(&stats[st], v);
      }
#pragma unroll
    for (int u = 0; u < SPL; u++) {
      const int st = ET.pistat[SPL * lane + u];
      if (st >= 0 && SC[u] != 0) atomicAdd(&stats[st], SC[u]);
    }
    if (score != 0) atomicAdd(&stats[ET.n_stats], score);
  }
}

// emission-count tables [replica][H][SP] -> emission statistics
// (one warp per cell: the lanes stride over the copies, then a shuffle tree)
__global__ void k_sparse_fold_ec(SparseDev F, const int *etab, const int *emod, int n_child, const double *gEC, double *stats) {
  const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= F.H * F.SP) return;
  const int hh = i / F.SP, j = i % F.SP;              // rows are written at SPL*lane + u == state index
  double v = 0;
  for (int rep = lane; rep <= F.ec_mask; rep += 32) v += gEC[(size_t)rep * F.H * F.SP + i];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if (lane == 0 && v != 0 && etab[j] >= 0) atomicAdd(&stats[n_child + etab[j] + hh % emod[j]], v);
}

}  // namespace jhs
