// kernels_lanechunks.cuh — chunk pass of the long-sequence path for SMALL dense models (2..8 states after padding), ONE CHUNK
// PER LANE.  The warp-per-chunk kernels of kernels_sparse.cuh keep one state per lane: for a 2-state model 30 of 32 lanes
// idle.  Here a lane owns a whole chunk (like kernels_lanes.cuh owns a whole sequence): state vector, T and the S x S
// count accumulators in registers, the chunk's forward rows in an HBM scratch laid out [position][lane][S] (coalesced
// across the lanes of a warp), the emission table in shared memory.  Inputs are the checkpoints of the parallel checkpoint
// pass (kernels_scan.cuh); the arithmetic per position is the one of k_sparse_decode_chunks / k_sparse_estep_chunks.
//   MODE 0: posterior decoding (per position the state with the largest alpha*beta, lowest index on ties)
//   MODE 1: Baum-Welch expected counts
#pragma once
#include "kernels_scan.cuh"

namespace jhc {

using jhf::FastDev;

struct LaneChunkIn {
  int CK, ROWS, SP;              // positions per chunk, rows of scratch per lane, row stride of the checkpoints
  int64_t n_chunks;
  const int32_t *chunk_seq, *chunk_idx;
  const int64_t *chunk_ptr;
  const double *ck_alpha, *ck_beta;            // [chunk_ptr[seq] + c][SP]
  const long long *ck_alpha_exp, *ck_beta_exp; // MODE 1
  const double *seq_sfin;                      // MODE 1: P(x) = seq_sfin * 2^seq_AL of the sequences with several chunks
  const long long *seq_AL;
};

template <int G>
__device__ __forceinline__ bool rescale_lane(double (&v)[G], long long &expo, bool &ok) {
  unsigned hi = (unsigned)__double2hiint(v[0]);
#pragma unroll
  for (int j = 1; j < G; j++) hi = max(hi, (unsigned)__double2hiint(v[j]));
  const int be = (int)(hi >> 20);
  if (be == 0) return false;
  const double f = __hiloint2double((2046 - be) << 20, 0);
#pragma unroll
  for (int j = 0; j < G; j++) v[j] *= f;
  expo += be - 1023;
  ok &= (unsigned)(be - 81) < (0x7feu - 81u);
  return true;
}

__device__ __forceinline__ double pow2_any(long long k) {  // 2^k, 0 below the normal range, largest power above it
  const int kk = (int)max(-1023ll, min(1023ll, k));
  return __hiloint2double((1023 + kk) << 20, 0);
}

// rows: [slot][ROWS][32][G], rexp: [slot][ROWS][32] (MODE 1); gEC: [copies][H][G]; stats as in kernels_mma.cuh
template <int G, int MODE>
__global__ void __launch_bounds__(128) k_lane_chunks(FastDev F, DevData D, LaneChunkIn I, unsigned long long *counter, double *rows, int *rexp,
                                                     PathOut paths, double *stats, double *gEC, int32_t *fallback, int ec_smem) {
  extern __shared__ double sm[];
  double *sE = sm;
  const int HG = F.H * G;
  // ec_smem: the emission statistic of a lane goes to its private column of a shared-memory table [H*G][32] per warp (no
  // atomics; small tables would otherwise funnel 2e8 reductions into a few hundred L2 addresses), flushed at the end
  double *const ecw = sm + HG + (size_t)(threadIdx.x >> 5) * HG * 32;
  for (int i = threadIdx.x; i < HG; i += blockDim.x) sE[i] = F.E[i];
  if (MODE == 1 && ec_smem)
    for (int i = threadIdx.x; i < HG * 32 * (int)(blockDim.x >> 5); i += blockDim.x) sm[HG + i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int64_t slot = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double *const R = rows + ((size_t)slot * I.ROWS * 32 + lane) * G;  // + r*32*G
  int *const RX = MODE == 1 ? rexp + (size_t)slot * I.ROWS * 32 + lane : nullptr;  // + r*32
  double *const gECw = gEC + (size_t)(slot & F.ec_mask) * HG;
  double T[G][G], fin[G];
#pragma unroll
  for (int i = 0; i < G; i++) {
    fin[i] = F.fin[i];
#pragma unroll
    for (int j = 0; j < G; j++) T[i][j] = F.T[i * G + j];
  }
  double O[G][G], SC[G];
#pragma unroll
  for (int i = 0; i < G; i++) {
    SC[i] = 0;
#pragma unroll
    for (int j = 0; j < G; j++) O[i][j] = 0;
  }
  double score = 0;
  auto emis = [&](int h, int t, double (&e)[G]) {
#pragma unroll
    for (int j = 0; j < G; j++) e[j] = sE[(size_t)h * G + j];
    if (t < F.Kmax) {
#pragma unroll
      for (int j = 0; j < G; j++)
        if (t < F.korder[j]) e[j] = F.inv_a;
    }
  };
  for (;;) {
    long long base = 0;
    if (lane == 0) base = (long long)atomicAdd(counter, 32ull);
    base = __shfl_sync(FULL, base, 0);
    if (base >= I.n_chunks) break;
    const int64_t g = base + lane;
    if (g < I.n_chunks) {
      const int sid = I.chunk_seq[g], c = I.chunk_idx[g];
      const int64_t off = D.offsets[sid];
      const int L = (int)(D.offsets[sid + 1] - off);
      const int t0 = c * I.CK, t1 = min(L, t0 + I.CK) - 1;
      const int64_t cb = I.chunk_ptr[sid] + c;
      const uint8_t *x = D.codes + off;
      const bool single = L <= I.CK;
      const double w = D.weights ? D.weights[sid] : 1.0;
      double sfin = (MODE == 1 && !single) ? I.seq_sfin[sid] : 1.0;
      long long AL = (MODE == 1 && !single) ? I.seq_AL[sid] : 0;
      bool ok = true, live = !(MODE == 1 && !single && !(sfin > 0));  // sequences the checkpoint pass could not finish are skipped
      // ---- forward rows of the chunk ----
      double a[G], a0[G];
      long long A = (MODE == 1 && c > 0) ? I.ck_alpha_exp[cb] : 0;
      const long long A0 = A;
      int h = jhp::hash_direct(F, x, t0);
      {
        double e[G];
        emis(h, t0, e);
        if (c == 0) {
#pragma unroll
          for (int j = 0; j < G; j++) {
            a0[j] = 0;
            a[j] = F.pi[j] * e[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < G; j++) a0[j] = I.ck_alpha[(size_t)cb * I.SP + j];
#pragma unroll
          for (int j = 0; j < G; j++) {
            double s = 0;
#pragma unroll
            for (int i = 0; i < G; i++) s = fma(a0[i], T[i][j], s);
            a[j] = s * e[j];
          }
        }
      }
      live = live && rescale_lane<G>(a, A, ok);
#pragma unroll
      for (int j = 0; j < G; j++) R[j] = a[j];
      if (MODE == 1) RX[0] = (int)(A - A0);
      for (int t = t0 + 1; live && t <= t1; t++) {
        const int xt = (int)__ldg(x + t);
        h = F.a_pow2 ? (((h * F.a) | xt) & (F.H - 1)) : ((h * F.a + xt) % F.H);
        double e[G], na[G];
        emis(h, t, e);
#pragma unroll
        for (int j = 0; j < G; j++) {
          double s = 0;
#pragma unroll
          for (int i = 0; i < G; i++) s = fma(a[i], T[i][j], s);
          na[j] = s * e[j];
        }
#pragma unroll
        for (int j = 0; j < G; j++) a[j] = na[j];
        if ((t & 7) == 0) live = rescale_lane<G>(a, A, ok);
        double *dst = R + (size_t)(t - t0) * 32 * G;
#pragma unroll
        for (int j = 0; j < G; j++) dst[j] = a[j];
        if (MODE == 1) RX[(size_t)(t - t0) * 32] = (int)(A - A0);
      }
      if (MODE == 1 && single && live) {  // the whole sequence is this chunk: P(x) from the last row
        double s = 0;
#pragma unroll
        for (int j = 0; j < G; j++) s += a[j] * fin[j];
        sfin = s;
        AL = A;
        live = ok && sfin > 0;
        if (live) score += w * (log(sfin) + (double)AL * 0.6931471805599453094);
      }
      if (MODE == 1 && single && !live) {
        int pos = atomicAdd(fallback, 1);
        fallback[1 + pos] = sid;
      }
      // ---- backward through the chunk ----
      if (live) {
        const double winv = MODE == 1 ? w / sfin : 1.0;
        double b[G];
        long long Bx = 0;
        if (t1 == L - 1) {
#pragma unroll
          for (int j = 0; j < G; j++) b[j] = fin[j];
        } else {
#pragma unroll
          for (int j = 0; j < G; j++) b[j] = I.ck_beta[(size_t)cb * I.SP + j];
          if (MODE == 1) Bx = I.ck_beta_exp[cb];
        }
        int hb = jhp::hash_direct(F, x, t1);
        bool okb = true;
        // The rows come back from L2 / HBM: a rolling window keeps rows t, t-1, t-2 in registers and requests row t-3 at the
        // top of step t, so a load has two steps to arrive (the row in front of the chunk is the checkpoint a0)
        auto load_row = [&](int t, double (&r)[G], int &rx) {
          if (t >= t0) {
            const double *src = R + (size_t)(t - t0) * 32 * G;
#pragma unroll
            for (int j = 0; j < G; j++) r[j] = src[j];
            rx = MODE == 1 ? RX[(size_t)(t - t0) * 32] : 0;
          } else {
#pragma unroll
            for (int j = 0; j < G; j++) r[j] = a0[j];
            rx = 0;
          }
        };
        double r0[G], r1[G], r2[G], r3[G];
        int x0, x1, x2, x3 = 0;
        load_row(t1, r0, x0);
        load_row(t1 - 1, r1, x1);
        load_row(t1 - 2, r2, x2);
#pragma unroll
        for (int j = 0; j < G; j++) r3[j] = 0;
        for (int t = t1; t >= t0; t--) {
          if (t - 3 >= t0 - 1) load_row(t - 3, r3, x3);
          double e[G];
          emis(hb, t, e);
          if (MODE == 0) {  // largest alpha_t[j] * beta_t[j], lowest index on ties
            double best = r0[0] * b[0];
            int arg = 0;
#pragma unroll
            for (int j = 1; j < G; j++) {
              const double gm = r0[j] * b[j];
              if (gm > best) { best = gm; arg = j; }
            }
            path_store(paths, off + t, arg);
          } else {
            const long long At = A0 + x0;
            const double sg = pow2_any(At + Bx - AL) * winv;
            double *gdst = gECw + (size_t)hb * G;
#pragma unroll
            for (int j = 0; j < G; j++) {
              const double gm = r0[j] * b[j] * sg;
              if (t >= F.Kmax || t >= F.korder[j]) {
                if (ec_smem) ecw[(size_t)(hb * G + j) * 32 + lane] += gm;
                else atomicAdd(gdst + j, gm);
              }
              if (t == 0) SC[j] += gm;
            }
          }
          if (t == 0) break;
          // alpha_{t-1} = r1: the previous row of the chunk, or the checkpoint in front of it
          double xb[G], nb[G];
#pragma unroll
          for (int j = 0; j < G; j++) xb[j] = e[j] * b[j];
          if (MODE == 1) {
            const double sx = pow2_any(A0 + x1 + Bx - AL) * winv;
#pragma unroll
            for (int i = 0; i < G; i++) {
              const double as = r1[i] * sx;
#pragma unroll
              for (int j = 0; j < G; j++) O[i][j] = fma(as, xb[j], O[i][j]);
            }
          }
          if (t == t0) break;  // the transition out of the previous chunk is counted, beta is not needed below t0
#pragma unroll
          for (int i = 0; i < G; i++) {
            double s = 0;
#pragma unroll
            for (int j = 0; j < G; j++) s = fma(T[i][j], xb[j], s);
            nb[i] = s;
          }
#pragma unroll
          for (int j = 0; j < G; j++) {
            b[j] = nb[j];
            r0[j] = r1[j];
            r1[j] = r2[j];
            r2[j] = r3[j];
          }
          x0 = x1; x1 = x2; x2 = x3;
          if ((t & 7) == 0) rescale_lane<G>(b, Bx, okb);
          const int q = t - 1 - F.Kmax;
          hb = hb / F.a + F.hpow * (q >= 0 ? (int)__ldg(x + q) : 0);
        }
      }
    }
    __syncwarp();
  }
  if (MODE == 0) return;
  // ---------------- flush: T is multiplied in once; sum over the lanes of the warp, one atomic per cell ----------------
  auto warp_sum = [&](double v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
  };
#pragma unroll
  for (int i = 0; i < G; i++) {
#pragma unroll
    for (int j = 0; j < G; j++) {
      const double v = warp_sum(O[i][j] * T[i][j]);
      if (lane == 0 && v != 0 && F.tstat[i * G + j] >= 0) atomicAdd(&stats[F.tstat[i * G + j]], v);
    }
    const double s = warp_sum(SC[i]);
    if (lane == 0 && s != 0 && F.pistat[i] >= 0) atomicAdd(&stats[F.pistat[i]], s);
  }
  const double sc = warp_sum(score);
  if (lane == 0 && sc != 0) atomicAdd(&stats[F.n_stats], sc);
  if (ec_smem) {
    __syncwarp();
    for (int c = 0; c < HG; c++) {
      const double v = warp_sum(ecw[(size_t)c * 32 + lane]);
      const int hh = c / G, j = c % G;
      if (lane == 0 && v != 0 && F.estat_tab[j] >= 0) atomicAdd(&stats[F.n_child + F.estat_tab[j] + hh % F.estat_mod[j]], v);
    }
  }
}

}  // namespace jhc
