// kernels_band.cuh — the latency-bound whole-sequence passes for BANDED models, warp resident.
//
// The checkpoint pass of the long-sequence path (kernels_sparse.cuh, phase 1) and the max sweep of the chromosome-scale
// Viterbi (kernels_vitlong.cuh, phase A) are ONE dependent step per position: for BASELINE cfg5 the 248 Mb sequence alone
// is 248e6 steps, and what a step costs in latency is what the whole call costs.  The general kernels exchange the state
// vector through shared memory: store -> barrier -> edge-list loads -> FMAs (66-70 ns per step with one state per thread
// and a CTA barrier, 160 ns with a warp and __syncwarp).
//
// Many HMMs over long sequences are banded: every edge i -> j stays within a few states of its source (left-to-right
// duration / profile-like chains, rings; cfg5: i -> i, i+1, i+2, i+3 mod 128).  With S = 32 * SPL states and SPL
// consecutive states per lane, a band of half-width <= SPL touches the two neighbour lanes only, so a step is
//     shuffle the neighbour lane's values (SHFL, register to register) -> FMAs with weights held in registers,
// no shared-memory round trip and no barrier at all; the rescale is lazy (the exponent of the vector's maximum is reduced
// off the dependent chain and applied two steps later — any power of two is exact), the emission row of the next
// position is fetched a step ahead.  The band [LO, HI] (offsets of the children relative to their source) is a template
// parameter, so the neighbour values that are never used are never shuffled and the FMA count is the band width.
//
// Results are those of the general kernels up to the summation order inside a step (sum-product) and bit-identical for the
// max sweep (max is order independent; the adds keep the reference's association (bwd + em) + trans).
#pragma once
#include "kernels_vitlong.cuh"

namespace jhb {

using jhs::Pass1Extra;
using jhs::SeqBlocks;
using jhs::SparseDev;
using jhx::ResReader;

template <int SPL, int LO, int HI>
struct Band {
  static_assert(LO <= HI && -SPL <= LO && HI <= SPL, "band outside the neighbour lanes");
  static constexpr int W = HI - LO + 1;
  // values of the previous / next lane that a lane needs when it reads index (u - d) [SRC_MINUS] or (u + d) for d in [LO, HI]
  // forward (sources j - d): previous lane's last max(HI,0) values, next lane's first max(-LO,0) values
  // backward (targets i + d): next lane's first max(HI,0) values, previous lane's last max(-LO,0) values
};

// weights of the lane's states per band offset: w[u][d - LO]; idx / T: padded edge lists [IND][SP] (kernels_sparse.cuh).
// sign = +1: the other end of edge k of state j is j + d (out-edges); sign = -1: j - d (in-edges).
template <int SPL, int LO, int HI, typename T>
__device__ __forceinline__ void band_weights(const int *idx, const T *w_in, int IND, int SP, int S, int lane, int sign, T zero, T (&w)[SPL][HI - LO + 1]) {
#pragma unroll
  for (int u = 0; u < SPL; u++)
#pragma unroll
    for (int x = 0; x <= HI - LO; x++) w[u][x] = zero;
  for (int k = 0; k < IND; k++) {
#pragma unroll
    for (int u = 0; u < SPL; u++) {
      const int j = SPL * lane + u;
      const int o = idx[(size_t)k * SP + j];
      const T wt = w_in[(size_t)k * SP + j];
      int d = sign * (o - j);
      d = ((d % S) + S) % S;
      if (d > S / 2) d -= S;
#pragma unroll
      for (int x = 0; x <= HI - LO; x++)
        if (d == x + LO && !(wt == zero)) w[u][x] = wt;  // padding edges carry the zero element
    }
  }
}

// c[0 .. 3 SPL): [previous lane | own | next lane]; only the parts a band can reach are shuffled
template <int SPL, int NPREV, int NNEXT>
__device__ __forceinline__ void exchange(const double (&v)[SPL], int lane, double (&c)[3 * SPL]) {
  const unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int u = 0; u < SPL; u++) c[SPL + u] = v[u];
#pragma unroll
  for (int u = 0; u < SPL; u++) {
    if (u >= SPL - NPREV) c[u] = __shfl_sync(FULL, v[u], (lane + 31) & 31);
    else c[u] = 0;
    if (u < NNEXT) c[2 * SPL + u] = __shfl_sync(FULL, v[u], (lane + 1) & 31);
    else c[2 * SPL + u] = 0;
  }
}

// lazy rescale: `probe` reduces the exponent of the vector's maximum (off the dependent chain), `apply` scales a later vector
struct LazyScale {
  unsigned hi = 0;
  template <int SPL>
  __device__ __forceinline__ void probe(const double (&v)[SPL]) {
    unsigned h = (unsigned)__double2hiint(v[0]);
#pragma unroll
    for (int i = 1; i < SPL; i++) h = max(h, (unsigned)__double2hiint(v[i]));
    hi = __reduce_max_sync(0xffffffffu, h);
  }
  template <int SPL>
  __device__ __forceinline__ bool apply(double (&v)[SPL], long long &expo) const {
    const int be = (int)(hi >> 20);
    const double f = __hiloint2double((2046 - be) << 20, 0);
#pragma unroll
    for (int i = 0; i < SPL; i++) v[i] *= f;
    expo += be - 1023;
    return (unsigned)(be - 81) < (0x7feu - 81u);
  }
};

// Checkpoint pass / scoring for banded models: one WARP per (sequence, direction), one warp per CTA.  Interface and outputs
// of jhs::k_sparse_pass1_cta.  The emission table lives in shared memory (models with larger tables keep the general kernels).
// A lone warp cannot hide instruction fetches, so (like k_sparse_pass1_cta) aligned blocks of 16 residues run as
// straight-line code without any branch — hash by shift/mask (power-of-two alphabets), no checkpoint inside, the rescale
// at fixed places (probe after steps 5 and 13, applied after steps 7 and 15) — and everything else (ragged ends, the first
// K positions, blocks that contain a checkpoint, other alphabets) goes through the general step.
template <int SPL, int LO, int HI>
__global__ void __launch_bounds__(32) k_band_pass1(SparseDev F, int IND, DevData D, int n_dir, int CK, const int64_t *chunk_ptr, double *ck_alpha,
                                                   double *ck_beta, double *loglik, int32_t *fallback, Pass1Extra X) {
  constexpr int W = HI - LO + 1;
  constexpr int FP = HI > 0 ? HI : 0, FN = LO < 0 ? -LO : 0;  // forward: values needed from the previous / next lane
  constexpr int BN = HI > 0 ? HI : 0, BP = LO < 0 ? -LO : 0;  // backward
  extern __shared__ double sm[];
  const int SP = F.SP, lane = threadIdx.x & 31;
  const double *sE = sm;
  for (int i = threadIdx.x; i < F.H * SP; i += blockDim.x) sm[i] = F.E[i];
  __syncthreads();
  const unsigned FULL = 0xffffffffu;
  const int ash = F.ashift, Hm = F.H - 1, hpow = F.hpow;
  const double *const sEl = sE + SPL * lane;  // + h * SP
  ResReader rd;
  SeqBlocks blocks;
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
    double w[SPL][W];
    LazyScale lz;
    if (!backward) {
      band_weights<SPL, LO, HI, double>(F.in_src, F.in_T, IND, SP, F.S, lane, -1, 0.0, w);
      double f[SPL];
      long long A = 0;
      bool ok = true;
      int h = jhs::hash_at(F, rd, 0);
      {
        double e[SPL];
        jhs::emis<SPL>(F, sE, lane, e, jhs::hmod_of(F, h), 0);
#pragma unroll
        for (int u = 0; u < SPL; u++) f[u] = F.pi[SPL * lane + u] * e[u];
      }
      ok = jhs::rescale_warp<SPL>(f, A);
      int ck_left = CK;  // steps until the next checkpoint (the vector BEFORE chunk c = alpha_{c*CK-1}, stored in step t = c*CK)
      double *ck_dst = ck_alpha ? ck_alpha + (size_t)(cbase + 1) * SP + SPL * lane : nullptr;
      long long *cke_dst = X.ck_alpha_exp ? X.ck_alpha_exp + cbase + 1 : nullptr;
      auto advance = [&](const double (&e)[SPL]) {  // alpha_{t-1} -> alpha_t given the emission row of t
        double c[3 * SPL];
        exchange<SPL, FP, FN>(f, lane, c);
#pragma unroll
        for (int u = 0; u < SPL; u++) {
          double a0 = 0;
#pragma unroll
          for (int x = 0; x < W; x++) a0 = fma(c[SPL + u - (x + LO)], w[u][x], a0);  // own values first for LO >= 0
          f[u] = a0 * e[u];
        }
      };
      auto step = [&](int t, int x) {
        h = jhs::roll_up(F, h, x);
        double e[SPL];
        jhs::emis<SPL>(F, sE, lane, e, h, t);
        if (ck_dst && --ck_left == 0) {
          jhs::store_vec<SPL>(ck_dst, f);
          ck_dst += SP;
          ck_left = CK;
          if (cke_dst) {
            if (lane == 0) *cke_dst = A;
            cke_dst++;
          }
        }
        advance(e);
        if ((t & 7) == 0) ok &= jhs::rescale_warp<SPL>(f, A);
      };
      auto fast16 = [&](const uint4 blk) {  // 16 steps of an aligned block: positions > Kmax, no checkpoint inside
        const unsigned wds[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const int x = (int)((wds[k >> 2] >> ((k & 3) * 8)) & 0xffu);
          h = ((h << ash) | x) & Hm;
          double e[SPL];
          jhs::load_vec<SPL>(sEl + (size_t)h * SP, e);
          advance(e);
          if (k == 5 || k == 13) lz.probe<SPL>(f);
          if (k == 7 || k == 15) ok &= lz.apply<SPL>(f, A);
        }
      };
      const int mid = L >> 1;                        // meeting point of the two directions (X.meet_vec)
      const int t_end = X.meet_vec ? mid : L - 1;    // last position of the forward recursion
      int t = 1;
      if (F.a_pow2) {
#pragma unroll 1
        for (; t <= t_end && ((((t + mis) & 15) != 0) || t <= F.Kmax); t++) step(t, (int)__ldg(bytes + t));
        int blk = (t + mis) >> 4;
        uint4 cur = t + 15 <= t_end ? blocks.load(blk) : make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (; t + 15 <= t_end; t += 16, blk++) {
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
      for (; t <= t_end; t++) step(t, (int)__ldg(bytes + t));
      ok &= jhs::rescale_warp<SPL>(f, A);
      if (X.meet_vec) {  // alpha_mid, its exponent and the flag; k_band_meet finishes the sequence
        const size_t slot = (size_t)(it / n_dir) * 2;
        jhs::store_vec<SPL>(X.meet_vec + slot * SP + SPL * lane, f);
        if (lane == 0) {
          X.meet_exp[slot] = A;
          X.meet_ok[slot] = ok ? 1 : 0;
        }
        __syncwarp();
        continue;
      }
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
      band_weights<SPL, LO, HI, double>(F.out_dst, F.out_T, IND, SP, F.S, lane, +1, 0.0, w);
      double b[SPL], e[SPL];
      long long Bx = 0;
#pragma unroll
      for (int u = 0; u < SPL; u++) b[u] = F.fin[SPL * lane + u];
      int h = jhs::hmod_of(F, jhs::hash_at(F, rd, L - 1));
      jhs::emis<SPL>(F, sE, lane, e, h, L - 1);
      // checkpoints: beta at the last position of chunk c, t = (c+1)*CK - 1 < L - 1
      const int n_full = (L - 1) / CK;            // chunks that end before position L-1
      int ck_left = n_full > 0 ? (L - 1) - (n_full * CK - 1) : 0x7fffffff;  // steps until t reaches n_full*CK-1
      double *ck_dst = ck_beta + (size_t)(cbase + n_full - 1) * SP + SPL * lane;
      long long *cke_dst = X.ck_beta_exp ? X.ck_beta_exp + cbase + n_full - 1 : nullptr;
      const int K1 = F.Kmax + 1;
      auto advance = [&]() {  // beta_t -> beta_{t-1}: the exchanged vector is e_t * beta_t
        double xb[SPL];
#pragma unroll
        for (int u = 0; u < SPL; u++) xb[u] = e[u] * b[u];
        double c[3 * SPL];
        exchange<SPL, BP, BN>(xb, lane, c);
#pragma unroll
        for (int u = 0; u < SPL; u++) {
          double a0 = 0;
#pragma unroll
          for (int xo = 0; xo < W; xo++) a0 = fma(c[SPL + u + (xo + LO)], w[u][xo], a0);
          b[u] = a0;
        }
      };
      auto step = [&](int p, int x) {  // t = p + K + 1: beta_t -> beta_{t-1}; x = x[t-1-K] completes the hash of t-1
        const int t = p + K1;
        if (ck_left == 0) {
          jhs::store_vec<SPL>(ck_dst, b);
          ck_dst -= SP;
          ck_left = CK;
          if (cke_dst) {
            if (lane == 0) *cke_dst = Bx;
            cke_dst--;
          }
        }
        --ck_left;
        advance();
        h = jhs::roll_down(F, h, x);
        jhs::emis<SPL>(F, sE, lane, e, h, t - 1);
        if ((t & 7) == 0) jhs::rescale_warp<SPL>(b, Bx);
      };
      auto fast16 = [&](const uint4 blk) {  // 16 steps down an aligned block: p >= 0 (all positions >= Kmax), no checkpoint inside
        const unsigned wds[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
        for (int k = 15; k >= 0; k--) {
          const int x = (int)((wds[k >> 2] >> ((k & 3) * 8)) & 0xffu);
          advance();
          h = (h >> ash) + hpow * x;
          jhs::load_vec<SPL>(sEl + (size_t)h * SP, e);
          if (k == 10 || k == 2) lz.probe<SPL>(b);
          if (k == 8 || k == 0) lz.apply<SPL>(b, Bx);
        }
      };
      int p = L - 2 - F.Kmax;
      // the last step goes from t = p + Kmax + 1 to t - 1: down to position 0, or to the meeting point L/2 (X.meet_vec)
      const int pa = X.meet_vec ? (L >> 1) - F.Kmax : -F.Kmax;
      const int pb = pa > 0 ? pa : 0;  // blocks must not reach below the last step
      if (F.a_pow2) {
#pragma unroll 1
        for (; p >= pa && (((p + mis) & 15) != 15 || p - 15 < pb); p--) step(p, p >= 0 ? (int)__ldg(bytes + p) : 0);
        int blk = (p + mis) >> 4;
        uint4 cur = p - 15 >= pb ? blocks.load(blk) : make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (; p - 15 >= pb; p -= 16, blk--) {  // (p + mis) & 15 == 15: the block is bytes[p-15 .. p]
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
      if (X.meet_vec) {  // beta_mid
        const bool okb = jhs::rescale_warp<SPL>(b, Bx);
        const size_t slot = (size_t)(it / n_dir) * 2 + 1;
        jhs::store_vec<SPL>(X.meet_vec + slot * SP + SPL * lane, b);
        if (lane == 0) {
          X.meet_exp[slot] = Bx;
          X.meet_ok[slot] = okb ? 1 : 0;
        }
      }
    }
    __syncwarp();
  }
}

// Scoring that meets in the middle: P(x) = sum_i alpha_mid[i] beta_mid[i].  The forward chain over [0, L/2] and the backward
// chain over [L/2, L) of a sequence run at the same time on two warps (k_band_pass1 with X.meet_vec), which halves the
// dependent chain that a long sequence costs (cfg5: 248e6 -> 124e6 steps for scoring).  One warp per sequence finishes it;
// outputs as in the forward-only pass (loglik, fallback list).  The value differs from the one-directional pass only by
// the association of the sums.
__global__ void __launch_bounds__(128) k_band_meet(int SP, DevData D, const int32_t *list, int64_t n_list, const double *meet_vec,
                                                   const long long *meet_exp, const int32_t *meet_ok, double *loglik, int32_t *fallback) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n_list) return;
  const int sid = list ? list[i] : D.order[i];
  const double *a = meet_vec + (size_t)i * 2 * SP, *b = a + SP;
  double s = 0;
  for (int j = lane; j < SP; j += 32) s += a[j] * b[j];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) {
    const bool good = meet_ok[i * 2] && meet_ok[i * 2 + 1] && s > 0;
    if (loglik) loglik[sid] = good ? log(s) + (double)(meet_exp[i * 2] + meet_exp[i * 2 + 1]) * 0.6931471805599453094 : JH_NEG_INF;
    if (!good) {
      int pos = atomicAdd(fallback, 1);
      fallback[1 + pos] = sid;
    }
  }
}

// Phase A of the chromosome-scale Viterbi for banded models (jhw::k_vit_sweep with the shuffle exchange): bit-identical rows.
// Same loop structure as k_band_pass1 (straight-line aligned blocks); the log emission table is read through L1.
template <int SPL, int LO, int HI>
__global__ void __launch_bounds__(32) k_band_vit_sweep(SparseDev F, jhw::VitLongDev V, int IND, DevData D, int CK, const int32_t *list, int64_t n_list,
                                                       const int64_t *chunk_ptr, double *ck_v, double *scores) {
  constexpr int W = HI - LO + 1;
  constexpr int BN = HI > 0 ? HI : 0, BP = LO < 0 ? -LO : 0;
  const int SP = F.SP, lane = threadIdx.x & 31;
  const int ash = F.ashift, hpow = F.hpow;
  const double *const lEl = V.lE + SPL * lane;  // + h * SP
  ResReader rd;
  SeqBlocks blocks;
  double w[SPL][W];
  band_weights<SPL, LO, HI, double>(F.out_dst, V.out_lT, IND, SP, F.S, lane, +1, JH_NEG_INF, w);
  for (long long it = blockIdx.x; it < n_list; it += gridDim.x) {
    const int sid = list[it];
    const int64_t off = D.offsets[sid];
    const int L = (int)(D.offsets[sid + 1] - off);
    rd.init(D.codes + off);
    blocks.init(D.codes + off);
    const int mis = blocks.mis;
    const uint8_t *const bytes = reinterpret_cast<const uint8_t *>(blocks.base) + mis;
    const int64_t cbase = chunk_ptr[sid];
    double b[SPL], e[SPL];
#pragma unroll
    for (int u = 0; u < SPL; u++) b[u] = V.lfin[SPL * lane + u];
    int h = jhs::hmod_of(F, jhs::hash_at(F, rd, L - 1));
    jhw::lemis<SPL>(F, V, lane, e, h, L - 1);
    const int K1 = F.Kmax + 1;
    int ck_left = (L - 1) % CK;
    double *ck_dst = ck_v + (size_t)(cbase + (L - 1) / CK) * SP + SPL * lane;
    auto advance = [&]() {  // bwd[t+1] -> bwd[t]
      double s[SPL];
#pragma unroll
      for (int u = 0; u < SPL; u++) s[u] = __dadd_rn(b[u], e[u]);
      double c[3 * SPL];
      exchange<SPL, BP, BN>(s, lane, c);
#pragma unroll
      for (int u = 0; u < SPL; u++) {
        double mx = __dadd_rn(c[SPL + u + LO], w[u][0]);  // (bwd + em) + trans; no NaN can occur (no +inf anywhere):
#pragma unroll
        for (int xo = 1; xo < W; xo++) {                   // a plain compare-and-select instead of fmax's NaN handling
          const double v = __dadd_rn(c[SPL + u + (xo + LO)], w[u][xo]);
          mx = v > mx ? v : mx;
        }
        b[u] = mx;
      }
    };
    auto step = [&](int p, int x) {
      const int t = p + K1;
      advance();
      h = jhs::roll_down(F, h, x);
      jhw::lemis<SPL>(F, V, lane, e, h, t - 1);
      if (ck_left == 0) {  // t is a multiple of CK: the row above chunk t/CK - 1
        jhs::store_vec<SPL>(ck_dst, b);
        ck_dst -= SP;
        ck_left = CK;
      }
      --ck_left;
    };
    auto fast16 = [&](const uint4 blk) {
      const unsigned wds[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
      for (int k = 15; k >= 0; k--) {
        const int x = (int)((wds[k >> 2] >> ((k & 3) * 8)) & 0xffu);
        advance();
        h = (h >> ash) + hpow * x;
        jhs::load_vec<SPL>(lEl + (size_t)h * SP, e);
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
      for (; p >= 15; p -= 16, blk--) {
        const uint4 nxt = blocks.load(blk - 1);
        if (ck_left < 16) {
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
    double score;
    int st;
    jhw::start_choice<SPL>(F, V, lane, b, e, score, st);
    if (lane == 0 && scores) scores[sid] = score;
    __syncwarp();
  }
}

}  // namespace jhb
