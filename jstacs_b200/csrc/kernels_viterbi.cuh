// kernels_viterbi.cuh — bit-exact Viterbi for plain first-order models (m == 1, no silent states, S <= 32, every
// context's children sorted by state), ONE SEQUENCE PER LANE.
//
// HigherOrderHMM.viterbi (HigherOrderHMM.java:617-708): backward max sweep
//     bwd[l][i] = max_ch ( (bwd[l+1][j_ch] + em(j_ch, l)) + t(i, ch) )                      (:546-549, :560-562)
// and a forward trace that picks, at every (layer, context), the FIRST child in child order whose value equals the
// maximum bit for bit (:633-650, strict '<' on the squared distance).  Only fp64 adds and compares are involved, so the
// result is reproduced exactly as long as the association (bwd + em) + trans is kept: s_j = bwd[j] + em_j is formed
// once per state and position, every edge is then ONE add `s_j + t(i,j)` and one compare.  With children sorted by
// state the first maximum in child order is the first maximum in state order; absent edges are -inf entries of a dense
// table (never selected; a context whose candidates are all -inf yields choice 255 = "impossible", like the reference's
// NaN distances).  Models with unsorted child lists or higher transition order stay on kernels_exact.cuh.
//
// Mapping: a warp takes four octets (32 sequences of similar length, kernels_mma.cuh), lane = sequence: no shuffles, no
// shared-memory exchange of DP values, no per-position barrier; the transition table is read from shared memory at a
// warp-uniform address (broadcast), the emission row by the lane's own context hash (octet hash records).  The choice
// of every (position, state) is one byte; the S bytes of a lane and position are packed and written as one coalesced
// store ([position][lane][S] layout); the trace re-reads them with loads that do not depend on the path.
#pragma once
#include "kernels_mma.cuh"

namespace jhv {

using jhm::OctData;

struct VitTables {
  const double *T;    // [S*S] T[i*S+j] = log score of edge i -> j, -inf if absent
  const double *pi;   // [S]   start element
  const double *fin;  // [S]   0 for final states, -inf otherwise
  const double *E;    // [H*S] E[h*S+j] = log emission score of state j for context hash h
  const int *korder;  // [S]
  int H, Kmax;
  double log_uniform;
};

template <int S>
struct VitCfg {
  static constexpr int CB = S < 4 ? 4 : S;  // choice bytes per (position, lane)
  static constexpr int NW = CB / 4;         // 32-bit words
};

// log-space dense tables from the score tables of the exact kernels (after every parameter update)
__global__ void k_viterbi_prepare(jhf::FastDev F, const int *tmap, const int *pimap, const double *tscore, const double *escore, double *T,
                                  double *pi, double *fin, double *E) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int G = F.G;
  if (idx < G * G) T[idx] = tmap[idx] >= 0 ? tscore[tmap[idx]] : JH_NEG_INF;
  if (idx < G) {
    pi[idx] = pimap[idx] >= 0 ? tscore[pimap[idx]] : JH_NEG_INF;
    fin[idx] = F.fin[idx] != 0.0 ? 0.0 : JH_NEG_INF;
  }
  if (idx < F.H * G) {
    int h = idx / G, j = idx % G;
    E[idx] = F.estat_tab[j] >= 0 ? escore[F.estat_tab[j] + h % F.estat_mod[j]] : JH_NEG_INF;
  }
}

// U8: the path leaves the kernel as one byte per residue (255 = no state), eight states per 64-bit store where the
// sequence's bytes are aligned; otherwise int32 per residue.
template <int S, bool EG, bool U8>
__global__ void __launch_bounds__(128) k_viterbi_lanes(VitTables V, OctData D, unsigned long long *counter, uint8_t *choice_scratch,
                                                       int64_t rows, void *paths, const int64_t *path_ptr, double *scores) {
  using C = VitCfg<S>;
  constexpr int CB = C::CB, NW = C::NW;
  extern __shared__ double sm[];
  double *sT = sm, *sPi = sT + S * S;
  const double *sE = EG ? V.E : sPi + S;  // EG: emission table too large for shared memory, read through L1/L2
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) sT[i] = V.T[i];
  for (int i = threadIdx.x; i < S; i += blockDim.x) sPi[i] = V.pi[i];
  if (!EG)
    for (int i = threadIdx.x; i < V.H * S; i += blockDim.x) sPi[S + i] = V.E[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int64_t slot = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint8_t *const ch_base = choice_scratch + ((size_t)slot * rows * 32 + lane) * CB;  // + l*32*CB
  const int64_t n_quads = (D.n_oct + 3) >> 2;
  double finv[S];
#pragma unroll
  for (int j = 0; j < S; j++) finv[j] = V.fin[j];

  for (;;) {
    long long qo = 0;
    if (lane == 0) qo = (long long)atomicAdd(counter, 1ull);
    qo = __shfl_sync(FULL, qo, 0);
    if (qo >= n_quads) break;
    const int64_t oct = qo * 4 + (lane >> 3);
    int sid = -1;
    if (oct < D.n_oct) sid = D.oct_seq[oct * 8 + (lane & 7)];
    int Lr = 0;
    int64_t off = 0;
    if (sid >= 0) {
      off = D.offsets[sid];
      Lr = (int)(D.offsets[sid + 1] - off);
    }
    const int Lmax = __reduce_max_sync(FULL, Lr), Lmin = __reduce_min_sync(FULL, Lr);
    // hash record stream of the lane's sequence (one zero record in front of the buffer, kernels_mma.cuh)
    const uint16_t *const hrow = D.hashes + ((oct < D.n_oct ? D.oct_hbase[oct] : 0) + 1) * 8 + (lane & 7);
    double bw[S];
#pragma unroll
    for (int j = 0; j < S; j++) bw[j] = finv[j];  // bwd[L][i] = finalState[i] ? 0 : -inf  (HOH:487-527)
    int hn = Lmax > 0 ? hrow[(size_t)(Lmax - 1) * 8] : 0;
    for (int l = Lmax - 1; l >= 1; --l) {
      const int h = hn;
      hn = hrow[(size_t)(l - 1) * 8];
      const bool act = l < Lr;
      // s_j = bwd[l+1][j] + em(j, l)
      double s[S];
      const double *erow = sE + (size_t)(act ? h : 0) * S;
      if (S >= 2) {
#pragma unroll
        for (int j = 0; j < S; j += 2) {
          const double2 e2 = *reinterpret_cast<const double2 *>(erow + j);
          s[j] = e2.x;
          s[j + 1] = e2.y;
        }
      }
      if (l < V.Kmax) {  // condition not available: logUniform (ACDE:447-450); 0 + logUniform == logUniform
#pragma unroll
        for (int j = 0; j < S; j++)
          if (l < V.korder[j]) s[j] = V.log_uniform;
      }
#pragma unroll
      for (int j = 0; j < S; j++) s[j] = __dadd_rn(bw[j], s[j]);
      unsigned cw[NW];
#pragma unroll
      for (int w = 0; w < NW; w++) cw[w] = 0;
      double nb[S];
#pragma unroll
      for (int i = 0; i < S; i++) {
        double mx = JH_NEG_INF;
        unsigned arg = 255u;
#pragma unroll
        for (int j = 0; j < S; j += 2) {
          const double2 t2 = *reinterpret_cast<const double2 *>(sT + i * S + j);  // warp-uniform address: broadcast
          const double v0 = __dadd_rn(s[j], t2.x), v1 = __dadd_rn(s[j + 1], t2.y);
          if (v0 > mx) { mx = v0; arg = j; }
          if (v1 > mx) { mx = v1; arg = j + 1; }
        }
        nb[i] = mx;
        cw[i >> 2] |= arg << ((i & 3) * 8);
      }
      if (l < Lmin) {
#pragma unroll
        for (int i = 0; i < S; i++) bw[i] = nb[i];
      } else {
#pragma unroll
        for (int i = 0; i < S; i++) bw[i] = act ? nb[i] : bw[i];
      }
      uint8_t *dst = ch_base + (size_t)l * 32 * CB;
      if (NW == 1) *reinterpret_cast<unsigned *>(dst) = cw[0];
      else if (NW == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(cw[0], cw[1]);
      else {
#pragma unroll
        for (int w = 0; w < NW; w += 4) *reinterpret_cast<uint4 *>(dst + 4 * w) = make_uint4(cw[w], cw[w + 1], cw[w + 2], cw[w + 3]);
      }
    }
    // layer 0: the start context
    double score = JH_NEG_INF;
    unsigned ctx = 255u;
    if (Lr > 0) {
      const double *erow = sE + (size_t)hn * S;
#pragma unroll
      for (int j = 0; j < S; j++) {
        double e = erow[j];
        if (0 < V.korder[j]) e = V.log_uniform;
        const double v = __dadd_rn(__dadd_rn(bw[j], e), sPi[j]);
        if (v > score) { score = v; ctx = j; }
      }
      if (scores) scores[sid] = score;
    }
    __syncwarp();  // the choice bytes of this warp are read back by the same lanes only, but keep the stores ordered
    // forward trace (HOH:627-664): the loads do not depend on the path
    if (paths && Lr > 0) {
      const int64_t pbase = path_ptr ? path_ptr[sid] : off;
      int32_t *p = U8 ? nullptr : reinterpret_cast<int32_t *>(paths) + pbase;
      uint8_t *p8 = U8 ? reinterpret_cast<uint8_t *>(paths) + pbase : nullptr;
      unsigned long long acc = 0;  // U8: states of the current aligned 8-byte group, flushed when its last byte is known
      int nacc = 0;
      auto emit = [&](int l, unsigned st) {  // st: state or 255
        if (!U8) {
          p[l] = st == 255u ? -1 : (int32_t)st;
          return;
        }
        acc |= (unsigned long long)st << (8 * nacc);
        nacc++;
        const bool last = l == Lr - 1;
        if ((((unsigned long long)(p8 + l) + 1) & 7ull) == 0 || last) {
          if (nacc == 8) *reinterpret_cast<unsigned long long *>(p8 + l - 7) = acc;
          else
            for (int k = 0; k < nacc; k++) p8[l - nacc + 1 + k] = (uint8_t)(acc >> (8 * k));
          acc = 0;
          nacc = 0;
        }
      };
      bool ok = ctx != 255u;
      emit(0, ok ? ctx : 255u);
      constexpr int TU = NW <= 2 ? 8 : (NW <= 4 ? 4 : 2);  // positions whose choice words are fetched together
      for (int l0 = 1; l0 < Lr; l0 += TU) {
        unsigned cw[TU][NW];
#pragma unroll
        for (int k = 0; k < TU; k++) {  // independent loads first, the dependent walk afterwards
          const uint8_t *src = ch_base + (size_t)min(l0 + k, Lr - 1) * 32 * CB;
          if (NW == 1) cw[k][0] = *reinterpret_cast<const unsigned *>(src);
          else if (NW == 2) {
            const uint2 v = *reinterpret_cast<const uint2 *>(src);
            cw[k][0] = v.x; cw[k][1] = v.y;
          } else {
#pragma unroll
            for (int w = 0; w < NW; w += 4) {
              const uint4 v = *reinterpret_cast<const uint4 *>(src + 4 * w);
              cw[k][w] = v.x; cw[k][w + 1] = v.y; cw[k][w + 2] = v.z; cw[k][w + 3] = v.w;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < TU; k++) {
          if (l0 + k < Lr) {
            unsigned word = cw[k][0];
#pragma unroll
            for (int w = 1; w < NW; w++) word = (ctx >> 2) == (unsigned)w ? cw[k][w] : word;
            const unsigned nxt = (word >> ((ctx & 3u) * 8u)) & 0xffu;
            ok = ok && nxt != 255u;
            ctx = ok ? nxt : 0u;
            emit(l0 + k, ok ? ctx : 255u);
          }
        }
      }
    }
    __syncwarp();
  }
}

// Viterbi training (HigherOrderHMM.viterbi with path == null, :617-708 + :723-726 of the E-step loop): the statistics are the
// weight of every sequence added along its best path.  The paths come from the fast decoders above (byte per residue, 255 =
// no state), so the E-step of Viterbi training is `decode + count` instead of the reference-order kernel with its choice
// matrix: one warp per sequence, lanes stride over the positions, counts go to a shared-memory copy of the statistics block
// when it fits (fp64 atomics on shared memory; a global table of a few hundred hot cells would serialise in the L2).
// stats: [transition | emission | score] as in kernels_mma.cuh.
__global__ void __launch_bounds__(256) k_viterbi_path_counts(jhf::FastDev F, DevData D, const uint8_t *paths, const double *scores, int n_cond_a,
                                                             int in_smem, double *stats) {
  extern __shared__ double sst[];
  const int n_stats = F.n_child + n_cond_a, G = F.G;
  double *const st_ = in_smem ? sst : stats;
  if (in_smem) {
    for (int i = threadIdx.x; i < n_stats; i += blockDim.x) sst[i] = 0;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  double score_sum = 0;
  for (int64_t n = (int64_t)blockIdx.x * wpc + (threadIdx.x >> 5); n < D.n_seq; n += (int64_t)gridDim.x * wpc) {
    const int64_t off = D.offsets[n];
    const int L = (int)(D.offsets[n + 1] - off);
    const double w = D.weights ? D.weights[n] : 1.0;
    const uint8_t *x = D.codes + off, *p = paths + off;
    if (lane == 0) score_sum += w * scores[n];
    for (int l = lane; l < L; l += 32) {
      const int s = p[l];
      if (s == 255) continue;  // the trace stopped: nothing is counted from here on (all later states are 255 as well)
      const int slot = l == 0 ? F.pistat[s] : F.tstat[(int)p[l - 1] * G + s];
      if (slot >= 0) atomicAdd(&st_[slot], w);
      if (l >= F.korder[s] && F.estat_tab[s] >= 0) {
        int h = 0, pw = 1;
        for (int k = 0; k <= F.Kmax && l - k >= 0; k++) {
          h += pw * (int)__ldg(x + l - k);
          pw *= F.a;
        }
        atomicAdd(&st_[F.n_child + F.estat_tab[s] + h % F.estat_mod[s]], w);
      }
    }
  }
  if (lane == 0 && score_sum != 0) atomicAdd(&stats[n_stats], score_sum);
  if (in_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_stats; i += blockDim.x)
      if (sst[i] != 0) atomicAdd(&stats[i], sst[i]);
  }
}

}  // namespace jhv
