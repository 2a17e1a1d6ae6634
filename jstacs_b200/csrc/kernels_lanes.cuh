// kernels_lanes.cuh — scaled linear-space E-step and scoring for SMALL models (2..4 states, e.g. the two-state
// CpG-island HMM of BASELINE cfg1), ONE SEQUENCE PER LANE.
//
// With S <= 4 the per-position work of a sequence is a handful of FMAs; spreading the states over lanes (kernels_fast.cuh)
// spends its time in shuffles and barriers.  Here a lane keeps alpha / beta (S doubles), the S x S count accumulators and
// the transition matrix in registers and runs its own sequence: no shuffles, no shared-memory exchange, the rescale is
// lane-local.  A warp takes four octets (32 sequences of similar length, kernels_mma.cuh); forward rows go to an HBM
// scratch laid out [position][lane][S] so that every store and load of a warp is one contiguous segment.  Emission counts
// go to a lane-private column of a shared-memory table ([context hash][state][lane], conflict free, no atomics) when it
// fits, else to the replicated global table of the octet kernel.
#pragma once
#include "kernels_mma.cuh"

namespace jhl {

using jhf::FastDev;
using jhm::OctData;
using jhm::pow2_clamped;

template <int S>
__device__ __forceinline__ bool rescale_lane(double (&v)[S], int &expo) {
  unsigned hi = (unsigned)__double2hiint(v[0]);
#pragma unroll
  for (int i = 1; i < S; i++) hi = max(hi, (unsigned)__double2hiint(v[i]));
  const int be = (int)(hi >> 20);
  const double f = __hiloint2double((2046 - be) << 20, 0);
#pragma unroll
  for (int i = 0; i < S; i++) v[i] *= f;
  expo += be - 1023;
  return (unsigned)(be - 81) < (0x7feu - 81u);
}

template <int S>
__device__ __forceinline__ void emis_lane(const FastDev &F, const double *sE, double (&e)[S], int h, int pos) {
  const double *row = sE + (size_t)h * S;
#pragma unroll
  for (int j = 0; j < S; j += 2) {
    const double2 x = *reinterpret_cast<const double2 *>(row + j);
    e[j] = x.x;
    e[j + 1] = x.y;
  }
  if (pos < F.Kmax) {
#pragma unroll
    for (int j = 0; j < S; j++)
      if (pos < F.korder[j]) e[j] = F.inv_a;
  }
}

template <int S>
__device__ __forceinline__ void store_row(double *dst, const double (&v)[S]) {
#pragma unroll
  for (int j = 0; j < S; j += 2) *reinterpret_cast<double2 *>(dst + j) = make_double2(v[j], v[j + 1]);
}
template <int S>
__device__ __forceinline__ void load_row(const double *src, double (&v)[S]) {
#pragma unroll
  for (int j = 0; j < S; j += 2) {
    const double2 x = *reinterpret_cast<const double2 *>(src + j);
    v[j] = x.x;
    v[j + 1] = x.y;
  }
}

// MODE 0: scoring only (out = log-likelihood per sequence); MODE 1: Baum-Welch E-step.
// stats layout and fallback list as in kernels_mma.cuh.  ec_smem != 0: emission counts in the shared lane-private table.
template <int S, int MODE, bool EG>
__global__ void __launch_bounds__(128) k_lanes(FastDev F, OctData D, unsigned long long *counter, double *fwd_scratch, int *fexp_scratch, int64_t rows,
                                               double *stats, double *gEC, int32_t *fallback, double *out, int ec_smem) {
  extern __shared__ double sm[];
  const int HS = F.H * S;
  const double *sE = EG ? F.E : sm;  // EG: emission table too large for shared memory (then ec_smem == 0 as well)
  double *ecw = sm + HS + (size_t)(threadIdx.x >> 5) * HS * 32;  // [H*S][32 lanes], this warp's table
  if (!EG)
    for (int i = threadIdx.x; i < HS; i += blockDim.x) sm[i] = F.E[i];
  if (MODE == 1 && ec_smem)
    for (int i = threadIdx.x; i < HS * 32 * (int)(blockDim.x >> 5); i += blockDim.x) sm[HS + i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int64_t slot = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double *const fw = MODE == 1 ? fwd_scratch + ((size_t)slot * rows * 32 + lane) * S : nullptr;  // + t*32*S
  int *const fe = MODE == 1 ? fexp_scratch + (size_t)slot * rows * 32 + lane : nullptr;           // + t*32
  const int64_t n_quads = (D.n_oct + 3) >> 2;
  double T[S][S], pi[S], fin[S];
#pragma unroll
  for (int i = 0; i < S; i++) {
    pi[i] = F.pi[i];
    fin[i] = F.fin[i];
#pragma unroll
    for (int j = 0; j < S; j++) T[i][j] = F.T[i * S + j];
  }
  double O[S][S], SC[S];
#pragma unroll
  for (int i = 0; i < S; i++) {
    SC[i] = 0;
#pragma unroll
    for (int j = 0; j < S; j++) O[i][j] = 0;
  }
  double score = 0;
  double *const gECw = gEC ? gEC + (size_t)(slot & F.ec_mask) * HS : nullptr;
  auto ec_add = [&](int h, int j, double g) {
    if (ec_smem) ecw[(size_t)(h * S + j) * 32 + lane] += g;
    else atomicAdd(gECw + (size_t)h * S + j, g);
  };

  for (;;) {
    long long qo = 0;
    if (lane == 0) qo = (long long)atomicAdd(counter, 1ull);
    qo = __shfl_sync(FULL, qo, 0);
    if (qo >= n_quads) break;
    const int64_t oct = qo * 4 + (lane >> 3);
    int sid = -1;
    if (oct < D.n_oct) sid = D.oct_seq[oct * 8 + (lane & 7)];
    int Lr = 0;
    double w = 0;
    if (sid >= 0) {
      Lr = (int)(D.offsets[sid + 1] - D.offsets[sid]);
      w = D.weights ? D.weights[sid] : 1.0;
    }
    const int Lmax = __reduce_max_sync(FULL, Lr);
    const uint16_t *const hrow = D.hashes + ((oct < D.n_oct ? D.oct_hbase[oct] : 0) + 1) * 8 + (lane & 7);
    // ================================ forward ================================
    double a[S], e[S];
    int A = 0;
    bool ok = true;
    emis_lane<S>(F, sE, e, Lr > 0 ? hrow[0] : 0, 0);
#pragma unroll
    for (int j = 0; j < S; j++) a[j] = Lr > 0 ? pi[j] * e[j] : 0.0;
    ok = rescale_lane<S>(a, A);
    if (MODE == 1) {
      store_row<S>(fw, a);
      fe[0] = A;
    }
    int hn = Lmax > 1 ? hrow[8] : 0;
    for (int t = 1; t < Lmax; t++) {
      const int h = hn;
      hn = hrow[(size_t)(t + 1) * 8];  // one record past the octet is readable (zero pad / next octet)
      const bool act = t < Lr;
      emis_lane<S>(F, sE, e, act ? h : 0, t);
      double n[S];
#pragma unroll
      for (int j = 0; j < S; j++) {
        double acc = 0;
#pragma unroll
        for (int i = 0; i < S; i++) acc = fma(a[i], T[i][j], acc);
        n[j] = acc * e[j];
      }
      int nA = A;
      const bool nok = rescale_lane<S>(n, nA);
      if (act) {
#pragma unroll
        for (int j = 0; j < S; j++) a[j] = n[j];
        A = nA;
        ok &= nok;
      } else {
#pragma unroll
        for (int j = 0; j < S; j++) n[j] = 0.0;
      }
      if (MODE == 1) {
        store_row<S>(fw + (size_t)t * 32 * S, n);
        fe[(size_t)t * 32] = nA;
      }
    }
    double sfin = 0;
#pragma unroll
    for (int j = 0; j < S; j++) sfin += a[j] * fin[j];
    const int AL = A;
    const bool good = Lr > 0 && ok && sfin > 0;
    double invS = 0;
    if (Lr > 0) {
      const double lp = good ? log(sfin) + AL * 0.6931471805599453094 : JH_NEG_INF;
      if (MODE == 0) out[sid] = lp;
      if (good) {
        invS = w / sfin;
        score += w * lp;
      } else {
        int pos = atomicAdd(fallback, 1);
        fallback[1 + pos] = sid;
      }
    }
    if (MODE == 0) continue;
    // ================================ backward + counts ================================
    __syncwarp();
    double beta[S];
    int Bx = 0;
#pragma unroll
    for (int j = 0; j < S; j++) beta[j] = fin[j] * invS;
    rescale_lane<S>(beta, Bx);
    double at[S];  // alpha_t of the position being consumed
    int At = 0;
    if (Lmax > 0) {
      load_row<S>(fw + (size_t)(Lmax - 1) * 32 * S, at);
      At = fe[(size_t)(Lmax - 1) * 32];
    }
    int h = Lmax > 0 ? hrow[(size_t)(Lmax - 1) * 8] : 0;
    for (int t = Lmax - 1; t >= 1; --t) {
      const bool act = t < Lr;
      const int hp = hrow[(size_t)(t - 1) * 8];
      double ap[S];  // alpha_{t-1}
      load_row<S>(fw + (size_t)(t - 1) * 32 * S, ap);
      const int Ap = fe[(size_t)(t - 1) * 32];
      emis_lane<S>(F, sE, e, act ? h : 0, t);
      if (act) {
        const double sc = pow2_clamped(At + Bx - AL), sp = pow2_clamped(Ap + Bx - AL);
        double b[S];
#pragma unroll
        for (int j = 0; j < S; j++) {
          const double g = at[j] * sc * beta[j];  // posterior of state j at position t (weighted)
          if (t >= F.Kmax || t >= F.korder[j]) ec_add(h, j, g);
          b[j] = e[j] * beta[j];
        }
        double nb[S];
#pragma unroll
        for (int i = 0; i < S; i++) {
          const double ai = ap[i] * sp;
          double acc = 0;
#pragma unroll
          for (int j = 0; j < S; j++) {
            O[i][j] = fma(ai, b[j], O[i][j]);  // T[i][j] is multiplied in at the end
            acc = fma(T[i][j], b[j], acc);
          }
          nb[i] = acc;
        }
        rescale_lane<S>(nb, Bx);
#pragma unroll
        for (int j = 0; j < S; j++) beta[j] = nb[j];
      }
#pragma unroll
      for (int j = 0; j < S; j++) at[j] = ap[j];
      At = Ap;
      h = hp;
    }
    if (Lr > 0) {  // position 0: emission statistic and the statistic of the start element
      const double sc = pow2_clamped(At + Bx - AL);
#pragma unroll
      for (int j = 0; j < S; j++) {
        const double g = at[j] * sc * beta[j];
        if (0 >= F.korder[j]) ec_add(h, j, g);
        SC[j] += g;
      }
    }
    __syncwarp();
  }
  if (MODE == 0) return;
  // ---------------- flush: sum over the lanes of the warp, one atomic per cell ----------------
  auto warp_sum = [&](double v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
  };
#pragma unroll
  for (int i = 0; i < S; i++) {
#pragma unroll
    for (int j = 0; j < S; j++) {
      const double v = warp_sum(O[i][j] * T[i][j]);
      if (lane == 0 && v != 0 && F.tstat[i * S + j] >= 0) atomicAdd(&stats[F.tstat[i * S + j]], v);
    }
    const double s = warp_sum(SC[i]);
    if (lane == 0 && s != 0 && F.pistat[i] >= 0) atomicAdd(&stats[F.pistat[i]], s);
  }
  const double sc = warp_sum(score);
  if (lane == 0 && sc != 0) atomicAdd(&stats[F.n_stats], sc);
  if (ec_smem) {
    __syncwarp();
    for (int c = 0; c < HS; c++) {
      const double v = warp_sum(ecw[(size_t)c * 32 + lane]);
      const int hh = c / S, j = c % S;
      if (lane == 0 && v != 0 && F.estat_tab[j] >= 0) atomicAdd(&stats[F.n_child + F.estat_tab[j] + hh % F.estat_mod[j]], v);
    }
  }
}

}  // namespace jhl
