// kernels_vitlong.cuh — bit-exact Viterbi for CHROMOSOME-SCALE sequences (BASELINE cfg5: 128 states x 250 Mb).
//
// HigherOrderHMM.viterbi (HigherOrderHMM.java:617-708) fills the whole (L+1) x C matrix bwd[l][c] (256 GB here) and then
// traces forward.  The decision at (layer, context) depends on the matrix only through bwd[l+1][.] and is taken with the
// very same expression the sweep evaluates, so it can be recomputed from any stored row.  Three exact phases:
//   A. k_vit_sweep    — per long sequence ONE warp runs the backward max sweep over the whole sequence and stores the row
//                       bwd[c*CK] at every chunk boundary (S doubles per CK positions) and the score bwd[0][0].  The values
//                       are accumulated strictly from the sequence end with the reference's association
//                       (bwd + em) + trans and __dadd_rn (no contraction), so every stored row is bit-identical to the
//                       reference's matrix row.  One dependent step per position: latency bound, like the checkpoint pass
//                       of the E-step (kernels_sparse.cuh).
//   B. k_vit_chunks<0> — every chunk is independent given the row above it: a warp recomputes the rows of its chunk (the
//                       same arithmetic from the same bits = the same rows), turns every decision into a "next state" byte
//                       nxt[t][i] (first child in CHILD ORDER whose value equals the maximum: strict '>' scan, the reference's
//                       strict '<' on the squared distance, :633-650) in a scratch of CK x S bytes, and follows them from
//                       every possible entry state: map_c[i] = state at the chunk's last position given state i before it.
//   C. k_vit_compose  — one thread per sequence walks over the chunk maps: entry state of every chunk (n_chunks steps).
//   D. k_vit_chunks<1> — the chunks again, now with their entry state known: recompute nxt, follow it, write the path.
// Sequences of a single chunk skip A and C.  Work per residue: one sweep step + two chunk steps, against one sweep step and
// a trace in the reference; memory: S doubles per CK positions + CK*S bytes per resident warp, independent of L.
//
// Scope: the sparse-table scope of kernels_sparse.cuh (first-order chain, no silent states, <= 255 states here because a
// state is one byte with 255 = "no state", bounded degree).  Viterbi training and windows stay on kernels_exact.cuh.
#pragma once
#include "kernels_sparse.cuh"

namespace jhw {

using jhs::SparseDev;
using jhs::SeqBlocks;
using jhx::ResReader;

struct VitLongDev {
  const double *out_lT;   // [IND][SP] log score of out-edge k of state i (child order), -inf for padding
  const double *lpi;      // [SP] log score of the start element's edge to state j, -inf if absent
  const int *pi_rank;     // [SP] position of state j among the start element's children, INT_MAX if absent
  const double *lfin;     // [SP] 0 for final states, -inf otherwise
  const double *lE;       // [H][SP] log emission score by context hash
  double log_uniform;     // 0 + (-log a), AbstractConditionalDiscreteEmission.java:447-450
};

// log-space tables from the score tables of the exact kernels (after every parameter update); tscore / escore hold the
// reference's values bit for bit (parameters - logNorm; (0 - logNorm[cond]) + params[cond][sym])
__global__ void k_vitlong_prepare(SparseDev F, int IND, const int *out_p, const int *pimap, const int *etab, const int *emod, const double *tscore,
                                  const double *escore, double *out_lT, double *lpi, double *lE) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < IND * F.SP) out_lT[idx] = out_p[idx] >= 0 ? tscore[out_p[idx]] : JH_NEG_INF;
  if (idx < F.SP) lpi[idx] = pimap[idx] >= 0 ? tscore[pimap[idx]] : JH_NEG_INF;
  if (idx < F.H * F.SP) {
    int h = idx / F.SP, j = idx % F.SP;
    lE[idx] = etab[j] >= 0 ? escore[etab[j] + h % emod[j]] : JH_NEG_INF;
  }
}

template <int SPL>
__device__ __forceinline__ void lemis(const SparseDev &F, const VitLongDev &V, int lane, double (&e)[SPL], int h, int64_t pos) {
  jhs::load_vec<SPL>(V.lE + (size_t)h * F.SP + SPL * lane, e);
  if (pos < F.Kmax) {
#pragma unroll
    for (int u = 0; u < SPL; u++)
      if (pos < F.korder[SPL * lane + u]) e[u] = V.log_uniform;
  }
}

template <int SPL, int IND>
struct MaxEdges {  // the lane's out-edge lists in registers: byte offsets into the shared vector, target states, log scores
  int off[IND][SPL];
  double w[IND][SPL];
  __device__ __forceinline__ void load(const int *idx, const double *T, int SP, int lane) {
#pragma unroll
    for (int k = 0; k < IND; k++)
#pragma unroll
      for (int u = 0; u < SPL; u++) {
        off[k][u] = jhs::slot_of<SPL>(idx[(size_t)k * SP + SPL * lane + u]) * 8;
        w[k][u] = T[(size_t)k * SP + SPL * lane + u];
      }
  }
  // out[u] = max_k ( vec[dst[k][u]] + w[k][u] ), arg[u] = first k (child order) that attains it, 255 if every value is -inf
  __device__ __forceinline__ void gather_max(const double *vec, double (&out)[SPL], int (&arg)[SPL]) const {
    const char *base = reinterpret_cast<const char *>(vec);
#pragma unroll
    for (int u = 0; u < SPL; u++) {
      double mx = JH_NEG_INF;
      int a = 255;
#pragma unroll
      for (int k = 0; k < IND; k++) {
        const double v = __dadd_rn(*reinterpret_cast<const double *>(base + off[k][u]), w[k][u]);  // (bwd + em) + trans
        if (v > mx) { mx = v; a = k; }
      }
      out[u] = mx;
      arg[u] = a;
    }
  }
  __device__ __forceinline__ void gather_max(const double *vec, double (&out)[SPL]) const {
    const char *base = reinterpret_cast<const char *>(vec);
#pragma unroll
    for (int u = 0; u < SPL; u++) {
      double mx = __dadd_rn(*reinterpret_cast<const double *>(base + off[0][u]), w[0][u]);
#pragma unroll
      for (int k = 1; k < IND; k++) {  // compare-and-select: no NaN can occur (no +inf anywhere), fmax's handling is not needed
        const double v = __dadd_rn(*reinterpret_cast<const double *>(base + off[k][u]), w[k][u]);
        mx = v > mx ? v : mx;
      }
      out[u] = mx;
    }
  }
};

// layer 0: the start context.  value = max_j ((bwd[1][j] + em(j,0)) + t(start, j)); state = first child of the start
// element (child order) that attains it, 255 if none.
template <int SPL>
__device__ __forceinline__ void start_choice(const SparseDev &F, const VitLongDev &V, int lane, const double (&b)[SPL], const double (&e)[SPL], double &score,
                                             int &state) {
  double best = JH_NEG_INF;
  int rk = 0x7fffffff, st = 255;
#pragma unroll
  for (int u = 0; u < SPL; u++) {
    const int j = SPL * lane + u;
    const double v = __dadd_rn(__dadd_rn(b[u], e[u]), V.lpi[j]);
    const int r = V.pi_rank[j];
    if (r != 0x7fffffff && (v > best || (v == best && r < rk && v > JH_NEG_INF))) { best = v; rk = r; st = j; }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, d);
    const int orr = __shfl_xor_sync(0xffffffffu, rk, d), os = __shfl_xor_sync(0xffffffffu, st, d);
    if (ob > best || (ob == best && orr < rk)) { best = ob; rk = orr; st = os; }
  }
  score = best;
  state = best > JH_NEG_INF ? st : 255;
}

// ---- phase A ---------------------------------------------------------------------------------------------------------
// list: the sequences of more than one chunk (longest first); ck_v[(chunk_ptr[sid] + c) * SP + j] = bwd[c*CK][j], c >= 1.
template <int SPL, int IND>
__global__ void __launch_bounds__(32) k_vit_sweep(SparseDev F, VitLongDev V, DevData D, int CK, const int32_t *list, int64_t n_list,
                                                  const int64_t *chunk_ptr, double *ck_v, unsigned long long *counter, double *scores) {
  extern __shared__ double sm[];
  const int SP = F.SP, lane = threadIdx.x & 31;
  double *vec = sm;  // [2][SP]
  const unsigned FULL = 0xffffffffu;
  ResReader rd;
  SeqBlocks blocks;
  MaxEdges<SPL, IND> G;
  G.load(F.out_dst, V.out_lT, SP, lane);
  for (;;) {
    long long it = 0;
    if (lane == 0) it = (long long)atomicAdd(counter, 1ull);
    it = __shfl_sync(FULL, it, 0);
    if (it >= n_list) break;
    const int sid = list[it];
    const int64_t off = D.offsets[sid];
    const int L = (int)(D.offsets[sid + 1] - off);
    rd.init(D.codes + off);
    blocks.init(D.codes + off);
    const int64_t cbase = chunk_ptr[sid];
    double b[SPL], e[SPL];
#pragma unroll
    for (int u = 0; u < SPL; u++) b[u] = V.lfin[SPL * lane + u];  // bwd[L][i] = finalState[i] ? 0 : -inf (HOH:487-527)
    int h = jhs::hash_at(F, rd, L - 1);
    int par = 0;
    lemis<SPL>(F, V, lane, e, jhs::hmod_of(F, h), L - 1);
    const int K1 = F.Kmax + 1;
    int ck_left = (L - 1) % CK;  // steps until t reaches a multiple of CK (t starts at L-1)
    double *ck_dst = ck_v + (size_t)(cbase + (L - 1) / CK) * SP + SPL * lane;
    auto step = [&](int p, int x) {  // t = p + K + 1: bwd[t+1] -> bwd[t]; x = x[t-1-K] completes the hash of t-1
      const int t = p + K1;
      double *v = vec + par;
      par ^= SP;
      double s[SPL];
#pragma unroll
      for (int u = 0; u < SPL; u++) s[u] = __dadd_rn(b[u], e[u]);
      jhs::put_shared<SPL>(v, lane, s);
      h = jhs::roll_down(F, h, x);
      lemis<SPL>(F, V, lane, e, h, t - 1);  // for the next step, in flight during the exchange
      __syncwarp();
      G.gather_max(v, b);
      if (ck_left == 0) {  // t is a multiple of CK: the row above chunk t/CK - 1
        jhs::store_vec<SPL>(ck_dst, b);
        ck_dst -= SP;
        ck_left = CK;
      }
      --ck_left;
    };
    jhs::for_residues_down(blocks, L - 2 - F.Kmax, -F.Kmax, step);  // t = L-1 .. 1
    double score;
    int st;
    start_choice<SPL>(F, V, lane, b, e, score, st);
    if (lane == 0 && scores) scores[sid] = score;
    __syncwarp();
  }
}

// ---- phases B and D ----------------------------------------------------------------------------------------------------
// nxt scratch: [CK][SP] bytes per resident warp.  TRACE = 0: map[(chunk_ptr[sid] + c) * SP + i] = state at the chunk's
// last position given state i at the position before the chunk (chunk 0: given the start context, every i).
// TRACE = 1: path of the chunk from entry[chunk_ptr[sid] + c] -> paths; single-chunk sequences also deliver their score.
template <int SPL, int IND, int TRACE>
__global__ void __launch_bounds__(128) k_vit_chunks(SparseDev F, VitLongDev V, DevData D, int CK, int64_t n_chunks, const int32_t *chunk_seq,
                                                    const int32_t *chunk_idx, const int64_t *chunk_ptr, const double *ck_v, unsigned long long *counter,
                                                    uint8_t *nxt_scratch, uint8_t *map, const uint8_t *entry, PathOut paths, const int64_t *path_ptr,
                                                    double *scores) {
  extern __shared__ double sm[];
  const int SP = F.SP, lane = threadIdx.x & 31;
  double *vec = sm + (size_t)(threadIdx.x >> 5) * 2 * SP;  // two buffers per warp
  const unsigned FULL = 0xffffffffu;
  uint8_t *const nxt = nxt_scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (size_t)CK * SP;
  ResReader rd;
  SeqBlocks blocks;
  MaxEdges<SPL, IND> G;
  G.load(F.out_dst, V.out_lT, SP, lane);
  int dst[IND][SPL];  // target state of the lane's out-edges (the byte written for a decision)
#pragma unroll
  for (int k = 0; k < IND; k++)
#pragma unroll
    for (int u = 0; u < SPL; u++) dst[k][u] = F.out_dst[(size_t)k * SP + SPL * lane + u];
  for (;;) {
    long long g = 0;
    if (lane == 0) g = (long long)atomicAdd(counter, 1ull);
    g = __shfl_sync(FULL, g, 0);
    if (g >= n_chunks) break;
    const int sid = chunk_seq[g], c = chunk_idx[g];
    const int64_t off = D.offsets[sid];
    const int L = (int)(D.offsets[sid + 1] - off);
    const int64_t cbase = chunk_ptr[sid];
    const int t0 = c * CK, t1 = min(t0 + CK, L) - 1;
    rd.init(D.codes + off);
    blocks.init(D.codes + off);
    double b[SPL], e[SPL];
    if (t1 == L - 1) {
#pragma unroll
      for (int u = 0; u < SPL; u++) b[u] = V.lfin[SPL * lane + u];
    } else {
      jhs::load_vec<SPL>(ck_v + (size_t)(cbase + c + 1) * SP + SPL * lane, b);  // bwd[t1 + 1]
    }
    int h = jhs::hash_at(F, rd, t1);
    int par = 0;
    lemis<SPL>(F, V, lane, e, jhs::hmod_of(F, h), t1);
    const int K1 = F.Kmax + 1;
    const int tlo = max(t0, 1);
    auto step = [&](int p, int x) {  // t = p + K + 1 in t1 .. tlo: decisions of layer t (state at position t given the state at t-1)
      const int t = p + K1;
      double *v = vec + par;
      par ^= SP;
      double s[SPL];
#pragma unroll
      for (int u = 0; u < SPL; u++) s[u] = __dadd_rn(b[u], e[u]);
      jhs::put_shared<SPL>(v, lane, s);
      h = jhs::roll_down(F, h, x);
      lemis<SPL>(F, V, lane, e, h, t - 1);
      __syncwarp();
      int arg[SPL];
      G.gather_max(v, b, arg);
      unsigned packed = 0;
#pragma unroll
      for (int u = 0; u < SPL; u++) {
        int n = 255;
#pragma unroll
        for (int k = 0; k < IND; k++) n = arg[u] == k ? dst[k][u] : n;
        if (SPL >= 4) packed |= (unsigned)n << (8 * (u & 3));
        if (SPL < 4) nxt[(size_t)(t - t0) * SP + SPL * lane + u] = (uint8_t)n;
        if (SPL >= 4 && (u & 3) == 3) {
          *reinterpret_cast<unsigned *>(nxt + (size_t)(t - t0) * SP + SPL * lane + (u & ~3)) = packed;
          packed = 0;
        }
      }
    };
    jhs::for_residues_down(blocks, t1 - 1 - F.Kmax, tlo - 1 - F.Kmax, step);  // t = t1 .. tlo
    int st0 = 255;
    if (t0 == 0) {  // b = bwd[1], e = emission scores of position 0
      double score;
      start_choice<SPL>(F, V, lane, b, e, score, st0);
      if (TRACE && lane == 0 && scores && t1 == L - 1) scores[sid] = score;
    }
    __syncwarp();  // the decision bytes of this warp are read back by its own lanes
    if (TRACE == 0) {
      int st[SPL];
#pragma unroll
      for (int u = 0; u < SPL; u++) st[u] = t0 == 0 ? st0 : SPL * lane + u;
      if (t0 == 0) {  // every entry of chunk 0 is the start context: one walk, by every lane
        for (int t = tlo; t <= t1; t++) st[0] = st[0] == 255 ? 255 : nxt[(size_t)(t - t0) * SP + st[0]];
#pragma unroll
        for (int u = 1; u < SPL; u++) st[u] = st[0];
      } else {
        for (int t = tlo; t <= t1; t++) {
#pragma unroll
          for (int u = 0; u < SPL; u++) st[u] = st[u] == 255 ? 255 : nxt[(size_t)(t - t0) * SP + st[u]];  // SPL independent walks per lane
        }
      }
#pragma unroll
      for (int u = 0; u < SPL; u++) map[(size_t)(cbase + c) * SP + SPL * lane + u] = (uint8_t)st[u];
    } else if (paths) {
      const int64_t pbase = path_ptr ? path_ptr[sid] : off;
      if (lane == 0) {
        int st = t0 == 0 ? st0 : (int)entry[cbase + c];
        if (t0 == 0) path_store(paths, pbase, st == 255 ? -1 : st);
        for (int t = tlo; t <= t1; t++) {
          st = st == 255 ? 255 : nxt[(size_t)(t - t0) * SP + st];
          path_store(paths, pbase + t, st == 255 ? -1 : st);
        }
      }
    }
    __syncwarp();
  }
}

// ---- phase C -----------------------------------------------------------------------------------------------------------
// entry[chunk_ptr[sid] + c] = state at position c*CK - 1 (c >= 1), from the chunk maps; one thread per long sequence
__global__ void k_vit_compose(int SP, int CK, const int32_t *list, int64_t n_list, const int64_t *offsets, const int64_t *chunk_ptr, const uint8_t *map,
                              uint8_t *entry) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_list) return;
  const int sid = list[i];
  const int64_t L = offsets[sid + 1] - offsets[sid], n_ck = (L + CK - 1) / CK, cbase = chunk_ptr[sid];
  int x = map[(size_t)cbase * SP];  // chunk 0: state at its last position
  for (int64_t c = 1; c < n_ck; c++) {
    entry[cbase + c] = (uint8_t)x;
    if (c + 1 < n_ck) x = x == 255 ? 255 : map[(size_t)(cbase + c) * SP + x];
  }
}

}  // namespace jhw
