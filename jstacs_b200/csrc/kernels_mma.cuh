// kernels_mma.cuh — Baum-Welch E-step with EIGHT sequences per warp in lock step ("octets"), the production E-step
// for dense first-order models with 5..32 states.
//
// Same scaled linear-space formulation as kernels_fast.cuh (exact power-of-two rescaling per position, one log per
// sequence, results identical to ~1e-13), different mapping.  Per position the three products of an octet
//     forward   alpha'[n][j]  = sum_i alpha[n][i] T[i][j]              ( 8 x G  by  G x G )
//     backward  beta' [n][i]  = sum_j b[n][j] T[i][j],  b = e * beta   ( 8 x G  by  G x G )
//     counts    O[i][j]      += sum_n alpha[n][i] b[n][j]              ( G x 8  by  8 x G )
// are small fp64 matrix products.  They are issued as warp-wide m8n8k4 FP64 MMA instructions: on B200 the measured
// DMMA rate equals the DFMA rate (37.0 vs 36.9 TFLOP/s, profiles/probe_fp64.py), so the ceiling is the same FP64
// pipe, but one DMMA replaces 8 DFMA + 4 shared-memory broadcast loads, which removes the issue-slot and
// latency limits of the lane-per-state kernel (36 % of the FP64 peak, profiles/r1_estep_simt.md).
//
// Register layouts (lane l, r = l>>2 the sequence of the octet, q = l&3):
//   RL  lane holds X[n = r][state 8t + 2q + b], t < NT, b < 2.  This is the D-fragment layout of the recursion
//       products AND, with the state permutation sigma(k,q) = 8(k>>1) + 2q + (k&1), their A-fragment layout:
//       the result of one step feeds the next without any data movement.
//   TL  lane holds X[n = 4kk + q][state 8t + r], kk < 2: A- and B-fragment layout of the count product, whose
//       reduction index is the SEQUENCE.  alpha comes from the forward scratch anyway (cp.async ring in shared
//       memory, read in either layout), b makes one round trip through shared memory per position.
//
// Residues are not touched here: the context hashes h(p) = sum_{i<=K} a^i x[p-i] depend only on the data, so they
// are computed ONCE per data set into octet-interleaved uint16 records (k_hash_octets) and streamed (2 B/residue/pass).
#pragma once
#include "kernels_fast.cuh"

namespace jhm {


using jhf::FastDev;

struct OctData {
  const int32_t *oct_seq;    // [n_oct*8] sequence ids of the octet rows, -1 = empty row; row 0 is the longest
  const int64_t *oct_hbase;  // [n_oct] first hash record of the octet
  const uint16_t *hashes;    // records of 8 uint16: [oct_hbase[o] + position][row]
  const int64_t *offsets;    // [n_seq+1]
  const double *weights;     // [n_seq] or nullptr
  int64_t n_oct;
};

// One thread per hash record (8 sequences at one position).  h(p) uses the K residues before p (fewer at the start).
__global__ void k_hash_octets(const uint8_t *codes, const int64_t *offsets, const int32_t *oct_seq, const int64_t *oct_hbase,
                              int64_t n_oct, int a, int K, uint16_t *hashes) {
  for (int64_t o = blockIdx.x; o < n_oct; o += gridDim.x) {
    int64_t off[8];
    int len[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      int sid = oct_seq[o * 8 + r];
      off[r] = sid >= 0 ? offsets[sid] : 0;
      len[r] = sid >= 0 ? (int)(offsets[sid + 1] - offsets[sid]) : 0;
    }
    uint4 *rec = reinterpret_cast<uint4 *>(hashes + (oct_hbase[o] + 1) * 8);  // record 0 of the buffer is a zero pad
    for (int p = threadIdx.x; p < len[0]; p += blockDim.x) {
      unsigned hv[8];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        int h = 0;
        if (p < len[r]) {
          int pw = 1;
          for (int i = 0; i <= K && p - i >= 0; i++) {
            h += pw * codes[off[r] + p - i];
            pw *= a;
          }
        }
        hv[r] = (unsigned)h;
      }
      rec[p] = make_uint4(hv[0] | (hv[1] << 16), hv[2] | (hv[3] << 16), hv[4] | (hv[5] << 16), hv[6] | (hv[7] << 16));
    }
  }
}

__device__ __forceinline__ void dmma(double &c0, double &c1, const double a, const double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
// predicated forms (no divergent branch in the position loops)
__device__ __forceinline__ void cp_async16_if(bool p, void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %2, 0;\n @p cp.async.cg.shared.global [%0], [%1], 16;\n}\n" ::"r"(sa), "l"(gmem), "r"((int)p) : "memory");
}
__device__ __forceinline__ void cp_async4_if(bool p, void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %2, 0;\n @p cp.async.ca.shared.global [%0], [%1], 4;\n}\n" ::"r"(sa), "l"(gmem), "r"((int)p) : "memory");
}
__device__ __forceinline__ double pow2_clamped(int k) {  // 2^k for k in [-1023, 1024): 0 below, finite above
  k = max(-1023, min(1023, k));
  return __hiloint2double((1023 + k) << 20, 0);
}

// exact rescale of the NV values a lane holds of its sequence: divide by 2^e, e = exponent of the row maximum
template <int NV>
__device__ __forceinline__ bool rescale_row(double (&v)[NV], int &expo) {
  unsigned hi = (unsigned)__double2hiint(v[0]);
#pragma unroll
  for (int i = 1; i < NV; i++) hi = max(hi, (unsigned)__double2hiint(v[i]));
  hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
  hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
  const int be = (int)(hi >> 20);
  const double f = __hiloint2double((2046 - be) << 20, 0);  // 2^-(be-1023); be == 0 keeps zeros zero
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] *= f;
  expo += be - 1023;
  // "ok": the row maximum is a normal number with 80 binades of head room below it, so that every component within
  // 2^-60 of the maximum was still normal at every position since the last rescale (a row maximum grows by at most
  // a factor 32 = 2^5 per position, RS = 4 positions between rescales): nothing that matters at 1e-16 was flushed
  return (unsigned)(be - 81) < (0x7feu - 81u);
}

__device__ __forceinline__ double piv_of(const FastDev &F, int s) { return F.pi[s]; }

// emission probabilities of position pos (context hash h) for the NV states of a lane in the RL layout
template <int NT, bool GEN>
__device__ __forceinline__ void emis_rl(const FastDev &F, const double *sE, int q, double (&e)[2 * NT], int h, int pos) {
  const double *row = sE + (size_t)h * (8 * NT) + 2 * q;
#pragma unroll
  for (int t = 0; t < NT; t++) {
    const double2 x = *reinterpret_cast<const double2 *>(row + 8 * t);
    e[2 * t] = x.x;
    e[2 * t + 1] = x.y;
  }
  if (GEN && pos < F.Kmax) {  // HigherOrderDiscreteEmission: condition not available -> uniform (ACDE:447-450)
#pragma unroll
    for (int v = 0; v < 2 * NT; v++)
      if (pos < F.korder[8 * (v >> 1) + 2 * q + (v & 1)]) e[v] = F.inv_a;
  }
}

// step kinds: general (ragged octet or position below the emission order, rescales), uniform with / without rescale.
// Uniform steps rescale every RS-th position only: between two rescales a row maximum grows by at
// most a factor G (column sums of T), so RS positions without rescaling stay far inside the 60 + 5*RS binades of head
// room that rescale_row's "ok" test demands.
using GeneralStep = std::integral_constant<int, 2>;
using UniformStep = std::integral_constant<int, 1>;
using UniformStepNoRescale = std::integral_constant<int, 0>;
constexpr int RS = 4;
// Hash records are 16 bytes per position: every eighth position opens a new 128-byte line, and with small models a
// position takes less time than a trip to L2 / HBM.  The line HPF positions ahead is prefetched into L1 once per RS steps.
constexpr int HPF = 48;
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }

// Scaled forward recursion of one octet (lane (r,q) of the warp) over the rows t_from+1 .. t_to.  t_from < 0: the
// recursion starts with the start element (row 0); otherwise A / Aexp hold row t_from on entry.  On return A holds row
// min(t_to, Lr-1) of the lane's sequence (scaled by 2^-Aexp); ok is cleared when a rescale came near the denormal range.
// STORE: rows and exponents go to the scratch, row t at fw + t*8*G and fe + t*8 (the caller shifts the base pointers).
template <int NT, bool STORE>
__device__ __forceinline__ void octet_forward(const FastDev &F, const double *sE, const int r, const int q, const uint16_t *hrow, const int Lr,
                                              const int Lmax, const int Lmin, double *fw, int *fe, double (&A)[2 * NT], int &Aexp, bool &ok,
                                              const int t_from = -1, const int t_to = 0x7fffffff) {
  constexpr int G = 8 * NT, NV = 2 * NT;
  const int kmin = F.Kmax;
  {
      double Bf[NV][NT];  // B fragments of T: Bf[k][t] = T[sigma(k,q)][8t + r]
#pragma unroll
      for (int k = 0; k < NV; k++)
#pragma unroll
        for (int t = 0; t < NT; t++) Bf[k][t] = F.T[(8 * (k >> 1) + 2 * q + (k & 1)) * G + 8 * t + r];
      double e[NV];
      const int ts = max(t_from, 0), te = min(t_to, Lmax - 1);  // steps ts .. te-1 (row t -> t+1)
      if (t_from < 0) {
        emis_rl<NT, true>(F, sE, q, e, hrow[0], 0);
#pragma unroll
        for (int v = 0; v < NV; v++) A[v] = Lr > 0 ? piv_of(F, 8 * (v >> 1) + 2 * q + (v & 1)) * e[v] : 0.0;
        ok = rescale_row<NV>(A, Aexp);
        if (STORE) {
#pragma unroll
          for (int t = 0; t < NT; t++) *reinterpret_cast<double2 *>(fw + 8 * t) = make_double2(A[2 * t], A[2 * t + 1]);
          if (q == 0) fe[r] = Aexp;
        }
      }
      emis_rl<NT, true>(F, sE, q, e, hrow[(size_t)(ts + 1) * 8], ts + 1);
      int h2 = hrow[(size_t)(ts + 2) * 8];  // hash of position t+2, loaded one step ahead of its use
      auto fstep = [&](auto general, int t) {  // row t -> t+1
        constexpr bool GEN = decltype(general)::value == 2, RESC = decltype(general)::value != 0;
        const int h3 = hrow[(size_t)(t + 3) * 8];
        double Dv[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) Dv[v] = 0;
#pragma unroll
        for (int k = 0; k < NV; k++)
#pragma unroll
          for (int tp = 0; tp < NT; tp++) dmma(Dv[2 * tp], Dv[2 * tp + 1], A[k], Bf[k][tp]);
#pragma unroll
        for (int v = 0; v < NV; v++) Dv[v] *= e[v];
        int nexp = Aexp;
        const bool nok = RESC ? rescale_row<NV>(Dv, nexp) : true;
        if (!GEN) {  // every sequence of the octet is still running
#pragma unroll
          for (int v = 0; v < NV; v++) A[v] = Dv[v];
          Aexp = nexp;
          ok &= nok;
        } else {  // ragged tail: finished sequences keep their last row in registers and store zeros
          const bool act = t + 1 < Lr;
#pragma unroll
          for (int v = 0; v < NV; v++) {
            A[v] = act ? Dv[v] : A[v];
            Dv[v] = act ? Dv[v] : 0.0;
          }
          Aexp = act ? nexp : Aexp;
          ok &= nok || !act;
        }
        if (STORE) {
          double *dst = fw + (size_t)(t + 1) * 8 * G;
#pragma unroll
          for (int tp = 0; tp < NT; tp++) *reinterpret_cast<double2 *>(dst + 8 * tp) = make_double2(Dv[2 * tp], Dv[2 * tp + 1]);
          if (q == 0) fe[(size_t)(t + 1) * 8 + r] = nexp;
        }
        emis_rl<NT, GEN>(F, sE, q, e, h2, t + 2);
        h2 = h3;
      };
      // uniform steps: t+1 < Lmin (all sequences running) and t+2 >= Kmax (no order check for the emission fetched)
      int t = ts;
      const int t_lo = min(max(kmin - 2, ts), te), t_hi = max(min(Lmin - 1, te), t_lo);
      for (; t < t_lo; t++) fstep(GeneralStep(), t);
      for (; t + RS <= t_hi; t += RS) {
        prefetch_l1(hrow + (size_t)min(t + HPF, Lmax) * 8);
#pragma unroll
        for (int u = 0; u < RS - 1; u++) fstep(UniformStepNoRescale(), t + u);
        fstep(UniformStep(), t + RS - 1);
      }
      for (; t < t_hi; t++) fstep(UniformStep(), t);
      for (; t < te; t++) fstep(GeneralStep(), t);
  }
}

template <int NT>
struct MmaCfg {
  static constexpr int G = 8 * NT, NV = 2 * NT, WARPS = 8, NTHREADS = 256;
  static constexpr int ROWB = G * 8;              // bytes of one sequence's row of a tile (no padding: XOR swizzle below)
  // forward rows come back from the HBM scratch through a ring of RING slots, issued PD = RING - 1 positions ahead: the
  // smaller the model the shorter a position, so the distance grows as the rows shrink (about 8 KB per warp for every NT)
  static constexpr int RING = NT == 1 ? 16 : NT == 2 ? 8 : 4, PD = RING - 1;
  static constexpr int SLOTB = 8 * ROWB + 32;     // 8 rows + 8 exponents
  static constexpr int WARPB = RING * SLOTB + 2 * 8 * ROWB + 64;
  static size_t smem_bytes(int H) { return ((size_t)H * G + (size_t)G * G + G) * sizeof(double) + (size_t)WARPS * WARPB; }
  // Tiles [sequence n][state] live in shared memory in 16-byte chunks; chunk c of row n is stored at c ^ swz(n).
  // With this both access patterns are bank-conflict free: RL (lane (r,q): 16 bytes at chunk 4t+q of row r) because the
  // two rows of a quarter warp differ in bit 2 of the chunk, TL (lane (r,q): 8 bytes of state 8t+r of row 4kk+q) because
  // the four rows of a half warp land in four different 32-byte bank groups.
  __device__ static __forceinline__ int swz(int n) { return NT >= 2 ? ((n & 2) | ((n & 1) << 2)) : (n & 2); }
};

// stats layout: [transition statistics n_child | emission statistics n_cond*a | sum_n w_n*logP_n]; gEC as in kernels_fast.cuh
//
// CHUNKED (long sequences: the (L+1) x 8 x G forward rows of an octet would not leave enough scratch for a full grid): the
// forward recursion first runs over the whole octet storing only a checkpoint row every CK positions (and P(x)); then the
// chunks are taken from the last to the first: forward rows of the chunk recomputed from its checkpoint into a scratch of
// CK+1 rows, backward recursion and counts through the chunk, beta carried in registers across the chunk boundary.  One
// extra forward pass (+25 % work), scratch independent of the sequence length.  `rows` = rows per slot (CK+1 when chunked),
// ck_scratch: [slot][ck_rows][8][G] checkpoint rows, ck_exp: [slot][ck_rows][8].
struct OctChunks {
  int CK;             // positions per chunk
  int64_t ck_rows;    // checkpoints per slot
  double *ck_scratch;
  int *ck_exp;
};

// DET (option "deterministic"): a fixed-order E-step — bit-identical statistics from run to run.  The three
// sources of run-to-run differences are removed: octets are assigned to the warp slots statically (octet o -> slot o mod
// slots, in increasing order) instead of through the work queue; emission counts are added as 64-bit FIXED-POINT integers
// (integer reductions commute; scale = a power of two chosen from the weighted residue count so that no cell can overflow,
// resolution comparable to an fp64 accumulator of that magnitude); transition / start counts and the score leave every
// warp as a row of partial sums that k_det_fold_parts adds in slot order.  Costs the dynamic load balance on ragged data.
struct DetOut {
  double *part;   // [slots][G*G + G + 1]: O * T | start statistic | score
  double scale;   // fixed-point scale of the emission counts
};

// EG: the emission table (more than 64 KB: high emission orders) stays in global memory (L1/L2) instead of shared memory;
// the fetch for position t+2 is issued at step t, so its latency stays off the recursion.
template <int NT, bool CHUNKED, bool EG, bool DET = false>
__global__ void __launch_bounds__(256, 1) k_mma_estep(FastDev F, OctData D, unsigned long long *counter, double *fwd_scratch,
                                                      int *fexp_scratch, int64_t rows, double *stats, double *gEC, int32_t *fallback,
                                                      OctChunks X, DetOut det) {
  using C = MmaCfg<NT>;
  constexpr int G = C::G, NV = C::NV, ROWB = C::ROWB, RING = C::RING, SLOTB = C::SLOTB;
  extern __shared__ double sm[];
  const int HG = F.H * G;
  const double *sE = EG ? F.E : sm;
  double *sO = sm + (EG ? 0 : HG), *sPi = sO + G * G;
  char *swarp = reinterpret_cast<char *>(sPi + G) + (size_t)(threadIdx.x >> 5) * C::WARPB;
  char *const ring = swarp;                        // [RING][SLOTB]
  char *const bbuf = swarp + RING * SLOTB;         // [2][8*ROWB]
  int *const sbx = reinterpret_cast<int *>(bbuf + 2 * 8 * ROWB);  // [2][8]
  if (!EG)
    for (int i = threadIdx.x; i < HG; i += blockDim.x) sm[i] = F.E[i];
  for (int i = threadIdx.x; i < G * G + G; i += blockDim.x) sO[i] = 0;
  __syncthreads();
  const int l = threadIdx.x & 31, r = l >> 2, q = l & 3;
  const unsigned FULL = 0xffffffffu;
  const int64_t slot = (int64_t)blockIdx.x * C::WARPS + (threadIdx.x >> 5);
  double *const fw = fwd_scratch + (size_t)slot * rows * 8 * G + r * G + 2 * q;  // + t*8*G + 8*tp (+b)
  int *const fe = fexp_scratch + (size_t)slot * rows * 8;                       // + t*8 + row
  const int kmin = F.Kmax;  // positions below Kmax need the per-state order check
  // emission-count table: the copies (kernels_fast.cuh) spread the fp64 reductions over the L2 slices (one 64 KB table kept a few
  // slices at 80 % atomic-unit load while the average was 28 %, profiles/r1_estep_octets.md); k_fast_fold_ec sums them
  double *const gECw = gEC + (size_t)(slot & F.ec_mask) * HG;
  // byte offsets inside a swizzled tile: RL chunk of (row r, tile tp), TL element of (row 4kk+q, state 8tp+r)
  int rl_off[NT], tl_off[2][NT];
#pragma unroll
  for (int tp = 0; tp < NT; tp++) {
    rl_off[tp] = r * ROWB + (((4 * tp + q) ^ C::swz(r)) << 4);
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
      const int n = 4 * kk + q, s = 8 * tp + r;
      tl_off[kk][tp] = n * ROWB + ((((s >> 1) ^ C::swz(n)) << 4) | ((s & 1) << 3));
    }
  }
  double O[NT][NT][2];
#pragma unroll
  for (int i = 0; i < NT; i++)
#pragma unroll
    for (int j = 0; j < NT; j++) O[i][j][0] = O[i][j][1] = 0;
  double SC[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) SC[v] = 0;
  double score = 0;
  auto ec_add = [&](double *p, double v) {  // emission statistic: fp64 reduction, or fixed point when the order must not matter
    if (DET) atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)__double2ll_rn(v * det.scale));
    else atomicAdd(p, v);
  };
  long long o_static = slot;

  for (;;) {
    long long o = 0;
    if (DET) {
      o = o_static;
      o_static += (long long)gridDim.x * C::WARPS;
    } else {
      if (l == 0) o = (long long)atomicAdd(counter, 1ull);
      o = __shfl_sync(FULL, o, 0);
    }
    if (o >= D.n_oct) break;
    const int sid = D.oct_seq[o * 8 + r];
    int Lr = 0;
    double w = 0;
    if (sid >= 0) {
      Lr = (int)(D.offsets[sid + 1] - D.offsets[sid]);
      w = D.weights ? D.weights[sid] : 1.0;
    }
    const int Lmax = __shfl_sync(FULL, Lr, 0);
    const int Lmin = __reduce_min_sync(FULL, Lr);
    // hash records: one zero record in front of the buffer and one behind it, so the prefetches below may run one or two
    // positions past either end of the octet
    const uint16_t *const hrow = D.hashes + (D.oct_hbase[o] + 1) * 8 + r;  // + position*8
    // ================================ forward ================================
    double A[NV];
    int Aexp = 0;
    bool ok = true;
    const int CK = CHUNKED ? X.CK : 0x40000000;
    const int n_ck = CHUNKED ? (Lmax + CK - 1) / CK : 1;
    double *const ckw = CHUNKED ? X.ck_scratch + (size_t)slot * X.ck_rows * 8 * G + r * G + 2 * q : nullptr;  // + c*8*G + 8*tp
    int *const cke = CHUNKED ? X.ck_exp + (size_t)slot * X.ck_rows * 8 : nullptr;                            // + c*8 + row
    if (!CHUNKED) octet_forward<NT, true>(F, sE, r, q, hrow, Lr, Lmax, Lmin, fw, fe, A, Aexp, ok);
    else {
      for (int c = 0; c < n_ck; c++) {  // checkpoint c = row c*CK - 1, the row in front of chunk c
        octet_forward<NT, false>(F, sE, r, q, hrow, Lr, Lmax, Lmin, nullptr, nullptr, A, Aexp, ok, c * CK - 1, (c + 1) * CK - 1);
        if (c + 1 < n_ck) {
#pragma unroll
          for (int tp = 0; tp < NT; tp++) *reinterpret_cast<double2 *>(ckw + (size_t)(c + 1) * 8 * G + 8 * tp) = make_double2(A[2 * tp], A[2 * tp + 1]);
          if (q == 0) cke[(c + 1) * 8 + r] = Aexp;
        }
      }
    }
    // ---- termination: P(x) = sum over final states, one log per sequence ----
    double sfin = 0;
#pragma unroll
    for (int v = 0; v < NV; v++) sfin += A[v] * F.fin[8 * (v >> 1) + 2 * q + (v & 1)];
    sfin += __shfl_xor_sync(FULL, sfin, 1);
    sfin += __shfl_xor_sync(FULL, sfin, 2);
    const int AL = Aexp;
    double invS = 0;
    if (Lr > 0) {
      if (ok && sfin > 0) {
        invS = w / sfin;
        if (q == 0) score += w * (log(sfin) + AL * 0.6931471805599453094);
      } else if (q == 0) {  // the scaled recursion lost (or nearly lost) its mass: the log-space kernels take this sequence
        int pos = atomicAdd(fallback, 1);
        fallback[1 + pos] = sid;
      }
    }
    // ================================ backward + counts ================================
    {
      double Bb[NV][NT];  // B fragments of T^T: Bb[k][t] = T[8t + r][sigma(k,q)]
#pragma unroll
      for (int k = 0; k < NV; k++)
#pragma unroll
        for (int t = 0; t < NT; t++) Bb[k][t] = F.T[(8 * t + r) * G + 8 * (k >> 1) + 2 * q + (k & 1)];
      double beta[NV];
      int Bx = 0;
#pragma unroll
      for (int v = 0; v < NV; v++) beta[v] = F.fin[8 * (v >> 1) + 2 * q + (v & 1)] * invS;  // 1/P(x) and the weight are folded in
      rescale_row<NV>(beta, Bx);
      int row_lo = 0;  // rows below row_lo are not in the scratch (chunked: the scratch holds rows row_lo .. of the current chunk)
      auto issue = [&](int row) {  // forward row `row` (8 sequences x G doubles + 8 exponents) -> ring slot row & 3
        const bool have = row >= (CHUNKED ? row_lo : 0);
        char *sl = ring + (size_t)(row & (RING - 1)) * SLOTB;
        const int64_t rel = CHUNKED ? row - row_lo : row;
        const double *src = fw + rel * 8 * G;
#pragma unroll
        for (int tp = 0; tp < NT; tp++) cp_async16_if(have, sl + rl_off[tp], src + 8 * tp);
        cp_async4_if(have && l < 8, sl + 8 * ROWB + l * 4, fe + rel * 8 + l);
        jhf::cp_async_commit();
      };
      __syncwarp();
      if (!CHUNKED) {
#pragma unroll
        for (int u = 0; u < C::PD; u++) issue(Lmax - 1 - u);
      }
      int h = hrow[(size_t)(Lmax - 1) * 8];
      int h1 = hrow[(size_t)(Lmax - 2) * 8];  // hash of position t-1, loaded one step ahead of its use
      double e[NV];
      emis_rl<NT, true>(F, sE, q, e, h, Lmax - 1);
      int par = 0;
      auto bstep = [&](auto general, int t) {  // consumes position t: beta_t -> beta_{t-1}, counts of the edges (t-1 -> t)
        constexpr bool GEN = decltype(general)::value == 2, RESC = decltype(general)::value != 0;
        const int h2 = hrow[(size_t)(t - 2) * 8];
        const bool act = !GEN || t < Lr;
        // critical path first: b = e * beta feeds the recursion product straight from registers
        double b[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) b[v] = act ? e[v] * beta[v] : 0.0;
        char *bb = bbuf + (size_t)par * 8 * ROWB;
#pragma unroll
        for (int tp = 0; tp < NT; tp++) *reinterpret_cast<double2 *>(bb + rl_off[tp]) = make_double2(b[2 * tp], b[2 * tp + 1]);
        if (q == 0) sbx[par * 8 + r] = Bx - AL;
        jhf::cp_async_wait<C::PD - 2>();  // rows t and t-1 have landed; rows t-2 .. t-PD+1 may still be in flight
        __syncwarp();
        issue(t - C::PD);  // PD - 1 steps ahead of its first use; its slot held row t+1, last read before this __syncwarp
        const char *slt = ring + (size_t)(t & (RING - 1)) * SLOTB;
        const char *slp = ring + (size_t)((t - 1) & (RING - 1)) * SLOTB;
        // beta_{t-1}[n][i] = sum_j b[n][j] T[i][j]
        double Dv[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) Dv[v] = 0;
#pragma unroll
        for (int k = 0; k < NV; k++)
#pragma unroll
          for (int tp = 0; tp < NT; tp++) dmma(Dv[2 * tp], Dv[2 * tp + 1], b[k], Bb[k][tp]);
        {  // emission statistic of position t: gamma_t = alpha_t * beta_t * 2^(A_t + Bx - AL)
          const int At = *reinterpret_cast<const int *>(slt + 8 * ROWB + r * 4);
          double sc = pow2_clamped(At + Bx - AL);
          if (!act) sc = 0.0;
          double *gdst = gECw + (size_t)h * G + 2 * q;
#pragma unroll
          for (int tp = 0; tp < NT; tp++) {
            const double2 a2 = *reinterpret_cast<const double2 *>(slt + rl_off[tp]);
            const double g0 = a2.x * sc * beta[2 * tp], g1 = a2.y * sc * beta[2 * tp + 1];
            if (!GEN || t >= kmin) {
              ec_add(gdst + 8 * tp, g0);
              ec_add(gdst + 8 * tp + 1, g1);
            } else {
              if (t >= F.korder[8 * tp + 2 * q]) ec_add(gdst + 8 * tp, g0);
              if (t >= F.korder[8 * tp + 2 * q + 1]) ec_add(gdst + 8 * tp + 1, g1);
            }
          }
        }
        // O[i][j] += sum_n alpha_{t-1}[n][i] 2^(A_{t-1}[n] + Bx[n] - AL[n]) b[n][j]   (T[i][j] is multiplied in at the end)
        {
          double aT[2][NT], bT[2][NT];
#pragma unroll
          for (int kk = 0; kk < 2; kk++) {
            const int n = 4 * kk + q;
            const int ea = *reinterpret_cast<const int *>(slp + 8 * ROWB + n * 4);
            const double f = pow2_clamped(ea + sbx[par * 8 + n]);
#pragma unroll
            for (int tp = 0; tp < NT; tp++) {
              aT[kk][tp] = *reinterpret_cast<const double *>(slp + tl_off[kk][tp]) * f;
              bT[kk][tp] = *reinterpret_cast<const double *>(bb + tl_off[kk][tp]);
            }
          }
#pragma unroll
          for (int kk = 0; kk < 2; kk++)
#pragma unroll
            for (int ti = 0; ti < NT; ti++)
#pragma unroll
              for (int tj = 0; tj < NT; tj++) dmma(O[ti][tj][0], O[ti][tj][1], aT[kk][ti], bT[kk][tj]);
        }
        int nbx = Bx;
        if (RESC) rescale_row<NV>(Dv, nbx);  // cannot fail while the forward mass is positive; NaN reaches the finite check
        if (act) {
#pragma unroll
          for (int v = 0; v < NV; v++) beta[v] = Dv[v];
          Bx = nbx;
        }
        h = h1;
        h1 = h2;
        emis_rl<NT, GEN>(F, sE, q, e, h, t - 1);
        par ^= 1;
      };
      for (int c = n_ck - 1; c >= 0; --c) {
        const int t0 = CHUNKED ? c * CK : 0, t1 = CHUNKED ? min(t0 + CK, Lmax) - 1 : Lmax - 1;
        if (CHUNKED) {  // forward rows t0-1 .. t1 of the chunk -> scratch rows 0 .. (c == 0: rows 0 .. t1)
          row_lo = max(t0 - 1, 0);
          if (c > 0) {
#pragma unroll
            for (int tp = 0; tp < NT; tp++) {
              const double2 a2 = *reinterpret_cast<const double2 *>(ckw + (size_t)c * 8 * G + 8 * tp);
              A[2 * tp] = a2.x;
              A[2 * tp + 1] = a2.y;
              *reinterpret_cast<double2 *>(fw + 8 * tp) = a2;
            }
            Aexp = cke[c * 8 + r];
            if (q == 0) fe[r] = Aexp;
          } else Aexp = 0;
          bool ok2 = true;
          octet_forward<NT, true>(F, sE, r, q, hrow, Lr, Lmax, Lmin, fw - (int64_t)row_lo * 8 * G, fe - (int64_t)row_lo * 8, A, Aexp, ok2, t0 - 1, t1);
          __syncwarp();
#pragma unroll
          for (int u = 0; u < C::PD; u++) issue(t1 - u);
        }
        // uniform steps: t < Lmin (all sequences running) and t-1 >= Kmax (no order checks)
        int t = t1;
        const int tl = max(t0, 1), t_hi = min(Lmin - 1, t1), t_lo = max(kmin + 1, tl);
        for (; t > t_hi && t >= tl; --t) bstep(GeneralStep(), t);
        for (; t - RS + 1 >= t_lo; t -= RS) {
          prefetch_l1(hrow + (size_t)max(t - HPF, 0) * 8);
#pragma unroll
          for (int u = 0; u < RS - 1; u++) bstep(UniformStepNoRescale(), t - u);
          bstep(UniformStep(), t - RS + 1);
        }
        for (; t >= t_lo; --t) bstep(UniformStep(), t);
        for (; t >= tl; --t) bstep(GeneralStep(), t);
        jhf::cp_async_wait<0>();
        __syncwarp();
      }
      {  // position 0: emission statistic and the statistic of the start element
        const char *sl0 = ring;
        const int A0 = *reinterpret_cast<const int *>(sl0 + 8 * ROWB + r * 4);
        const double sc = Lr > 0 ? pow2_clamped(A0 + Bx - AL) : 0.0;
        double *gdst = gECw + (size_t)h * G + 2 * q;
#pragma unroll
        for (int tp = 0; tp < NT; tp++) {
          const double2 a2 = *reinterpret_cast<const double2 *>(sl0 + rl_off[tp]);
          const double g0 = a2.x * sc * beta[2 * tp], g1 = a2.y * sc * beta[2 * tp + 1];
          if (0 >= F.korder[8 * tp + 2 * q]) ec_add(gdst + 8 * tp, g0);
          if (0 >= F.korder[8 * tp + 2 * q + 1]) ec_add(gdst + 8 * tp + 1, g1);
          SC[2 * tp] += g0;
          SC[2 * tp + 1] += g1;
        }
      }
      __syncwarp();
    }
  }
  if (DET) {  // ---------------- flush, fixed order: one row of partial sums per warp slot, added by k_det_fold_parts ----------------
    double *const P = det.part + (size_t)slot * (G * G + G + 1);
#pragma unroll
    for (int ti = 0; ti < NT; ti++)
#pragma unroll
      for (int tj = 0; tj < NT; tj++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const int i = 8 * ti + r, j = 8 * tj + 2 * q + b;
          P[i * G + j] = O[ti][tj][b] * F.T[i * G + j];
        }
#pragma unroll
    for (int v = 0; v < NV; v++) {
      double s = SC[v];
      s += __shfl_xor_sync(FULL, s, 4);
      s += __shfl_xor_sync(FULL, s, 8);
      s += __shfl_xor_sync(FULL, s, 16);
      if (r == 0) P[G * G + 8 * (v >> 1) + 2 * q + (v & 1)] = s;
    }
    double s = q == 0 ? score : 0.0;
    s += __shfl_xor_sync(FULL, s, 4);
    s += __shfl_xor_sync(FULL, s, 8);
    s += __shfl_xor_sync(FULL, s, 16);
    if (l == 0) P[G * G + G] = s;
  }
  if (!DET) {
  // ---------------- flush: T is multiplied in once; CTA-level reduction, then one atomic per cell ----------------
#pragma unroll
  for (int ti = 0; ti < NT; ti++)
#pragma unroll
    for (int tj = 0; tj < NT; tj++)
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int i = 8 * ti + r, j = 8 * tj + 2 * q + b;
        const double v = O[ti][tj][b] * F.T[i * G + j];
        if (v != 0) atomicAdd(&sO[i * G + j], v);
      }
#pragma unroll
  for (int v = 0; v < NV; v++) {  // start statistic: sum over the sequences of the octet (lanes with equal q), then one atomic
    double s = SC[v];
    s += __shfl_xor_sync(FULL, s, 4);
    s += __shfl_xor_sync(FULL, s, 8);
    s += __shfl_xor_sync(FULL, s, 16);
    if (r == 0 && s != 0) atomicAdd(&sPi[8 * (v >> 1) + 2 * q + (v & 1)], s);
  }
  if (score != 0) atomicAdd(&stats[F.n_stats], score);
  __syncthreads();
  for (int i = threadIdx.x; i < G * G; i += blockDim.x)
    if (sO[i] != 0 && F.tstat[i] >= 0) atomicAdd(&stats[F.tstat[i]], sO[i]);
  for (int i = threadIdx.x; i < G; i += blockDim.x)
    if (sPi[i] != 0 && F.pistat[i] >= 0) atomicAdd(&stats[F.pistat[i]], sPi[i]);
  }
}

// Deterministic mode: rows of per-slot partial sums [slots][G*G + G + 1] -> statistics, in slot order per cell and in cell order
// per statistics slot (tied transition elements share slots).  One block.
__global__ void __launch_bounds__(256) k_det_fold_parts(FastDev F, const double *part, int64_t slots, double *tmp, double *stats) {
  const int G = F.G, NC = G * G + G + 1;
  for (int c = threadIdx.x; c < NC; c += blockDim.x) {
    double v = 0;
    for (int64_t sl = 0; sl < slots; sl++) v += part[(size_t)sl * NC + c];
    tmp[c] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c = 0; c < G * G; c++)
      if (tmp[c] != 0 && F.tstat[c] >= 0) stats[F.tstat[c]] += tmp[c];
    for (int c = 0; c < G; c++)
      if (tmp[G * G + c] != 0 && F.pistat[c] >= 0) stats[F.pistat[c]] += tmp[G * G + c];
    stats[F.n_stats] += tmp[G * G + G];
  }
}

// Deterministic mode: fixed-point emission-count copies -> emission statistics.  Integer sums are exact, so the copies and
// the context hashes that share a cell (state order < Kmax) can be added in any order; the cells are then added to the
// statistics in a fixed order by one thread (states that share an emission share cells).  One block; tmp: [G * H] doubles.
__global__ void __launch_bounds__(256) k_det_fold_ec(FastDev F, const unsigned long long *gEC, double inv_scale, double *tmp, double *stats) {
  const int G = F.G, H = F.H;
  for (int idx = threadIdx.x; idx < G * H; idx += blockDim.x) {
    const int j = idx / H, ci = idx % H;
    double v = 0;
    if (F.estat_tab[j] >= 0 && ci < F.estat_mod[j]) {
      long long tot = 0;
      for (int hh = ci; hh < H; hh += F.estat_mod[j])
        for (int rep = 0; rep <= F.ec_mask; rep++) tot += (long long)gEC[((size_t)rep * H + hh) * G + j];
      v = (double)tot * inv_scale;
    }
    tmp[idx] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int j = 0; j < G; j++) {
      if (F.estat_tab[j] < 0) continue;
      for (int ci = 0; ci < F.estat_mod[j]; ci++)
        if (tmp[j * H + ci] != 0) stats[F.n_child + F.estat_tab[j] + ci] += tmp[j * H + ci];
    }
}

// Scoring (AbstractHMM.logProb, AbstractHMM.java:1049-1070: the value of bwd[0][0]) with the octet forward recursion;
// nothing is stored.  A sequence whose scaled recursion came near the denormal range is listed in `fallback`
// ([0] = count, then ids) and scored by the log-space kernel afterwards.
template <int NT, bool EG>
__global__ void __launch_bounds__(256) k_mma_loglik(FastDev F, OctData D, unsigned long long *counter, double *out, int32_t *fallback) {
  constexpr int G = 8 * NT, NV = 2 * NT;
  extern __shared__ double sm[];
  const int HG = F.H * G;
  const double *sE = EG ? F.E : sm;
  if (!EG)
    for (int i = threadIdx.x; i < HG; i += blockDim.x) sm[i] = F.E[i];
  __syncthreads();
  const int l = threadIdx.x & 31, r = l >> 2, q = l & 3;
  const unsigned FULL = 0xffffffffu;
  for (;;) {
    long long o = 0;
    if (l == 0) o = (long long)atomicAdd(counter, 1ull);
    o = __shfl_sync(FULL, o, 0);
    if (o >= D.n_oct) break;
    const int sid = D.oct_seq[o * 8 + r];
    int Lr = 0;
    if (sid >= 0) Lr = (int)(D.offsets[sid + 1] - D.offsets[sid]);
    const int Lmax = __shfl_sync(FULL, Lr, 0);
    const int Lmin = __reduce_min_sync(FULL, Lr);
    const uint16_t *const hrow = D.hashes + (D.oct_hbase[o] + 1) * 8 + r;
    double A[NV];
    int Aexp = 0;
    bool ok = true;
    octet_forward<NT, false>(F, sE, r, q, hrow, Lr, Lmax, Lmin, nullptr, nullptr, A, Aexp, ok);
    double sfin = 0;
#pragma unroll
    for (int v = 0; v < NV; v++) sfin += A[v] * F.fin[8 * (v >> 1) + 2 * q + (v & 1)];
    sfin += __shfl_xor_sync(FULL, sfin, 1);
    sfin += __shfl_xor_sync(FULL, sfin, 2);
    if (q == 0 && Lr > 0) {
      if (ok && sfin > 0) out[sid] = log(sfin) + Aexp * 0.6931471805599453094;
      else {
        out[sid] = JH_NEG_INF;
        int pos = atomicAdd(fallback, 1);
        fallback[1 + pos] = sid;
      }
    }
    __syncwarp();
  }
}

// Posterior decoding (HigherOrderHMM.fillLogStatePosteriorMatrix :294-329 + AbstractHMM.decodeStatePosterior
// AbstractHMM.java:1125-1139): per position the state with the largest posterior, lowest index on ties.  The posterior
// of state j at position t is alpha_t[j] * beta_t[j] up to a positive factor that is constant over the states of a
// position (the scales and 1/P(x)), so the decision needs neither exponents nor a logarithm.  Same octet mapping as the
// E-step: forward rows through the HBM scratch and a cp.async ring, backward recursion on the FP64 MMA path, no count
// product.  paths: int32 per residue (CSR by the data set offsets), loglik optional.
// CHUNKED: as in k_mma_estep (checkpoint rows every CK positions, rows of one chunk per resident warp).
template <int NT, bool CHUNKED, bool EG>
__global__ void __launch_bounds__(256, 1) k_mma_decode(FastDev F, OctData D, unsigned long long *counter, double *fwd_scratch, int *fexp_scratch,
                                                       int64_t rows, PathOut paths, double *loglik, int32_t *fallback, OctChunks X) {
  using C = MmaCfg<NT>;
  constexpr int G = C::G, NV = C::NV, ROWB = C::ROWB, RING = C::RING, SLOTB = C::SLOTB;
  extern __shared__ double sm[];
  const int HG = F.H * G;
  const double *sE = EG ? F.E : sm;
  char *const ring = reinterpret_cast<char *>(sm + (EG ? 0 : HG)) + (size_t)(threadIdx.x >> 5) * (RING * SLOTB);
  if (!EG)
    for (int i = threadIdx.x; i < HG; i += blockDim.x) sm[i] = F.E[i];
  __syncthreads();
  const int l = threadIdx.x & 31, r = l >> 2, q = l & 3;
  const unsigned FULL = 0xffffffffu;
  const int64_t slot = (int64_t)blockIdx.x * C::WARPS + (threadIdx.x >> 5);
  double *const fw = fwd_scratch + (size_t)slot * rows * 8 * G + r * G + 2 * q;
  int *const fe = fexp_scratch + (size_t)slot * rows * 8;
  int rl_off[NT];
#pragma unroll
  for (int tp = 0; tp < NT; tp++) rl_off[tp] = r * ROWB + (((4 * tp + q) ^ C::swz(r)) << 4);
  // state with the largest product among the lane's NV values, then over the four lanes of the sequence
  auto best_state = [&](const double (&a)[NV], const double (&b)[NV]) {
    double best = a[0] * b[0];
    int arg = 2 * q;
#pragma unroll
    for (int v = 1; v < NV; v++) {
      const double g = a[v] * b[v];
      const int st = 8 * (v >> 1) + 2 * q + (v & 1);
      if (g > best) { best = g; arg = st; }  // ascending state order inside the lane: strict '>' keeps the lowest index
    }
#pragma unroll
    for (int d = 1; d <= 2; d <<= 1) {
      const double ob = __shfl_xor_sync(FULL, best, d);
      const int oa = __shfl_xor_sync(FULL, arg, d);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    return arg;
  };
  for (;;) {
    long long o = 0;
    if (l == 0) o = (long long)atomicAdd(counter, 1ull);
    o = __shfl_sync(FULL, o, 0);
    if (o >= D.n_oct) break;
    const int sid = D.oct_seq[o * 8 + r];
    int Lr = 0;
    int64_t off = 0;
    if (sid >= 0) {
      off = D.offsets[sid];
      Lr = (int)(D.offsets[sid + 1] - off);
    }
    const int Lmax = __shfl_sync(FULL, Lr, 0);
    const int Lmin = __reduce_min_sync(FULL, Lr);
    const uint16_t *const hrow = D.hashes + (D.oct_hbase[o] + 1) * 8 + r;
    double A[NV];
    int Aexp = 0;
    bool ok = true;
    const int CK = CHUNKED ? X.CK : 0x40000000;
    const int n_ck = CHUNKED ? (Lmax + CK - 1) / CK : 1;
    double *const ckw = CHUNKED ? X.ck_scratch + (size_t)slot * X.ck_rows * 8 * G + r * G + 2 * q : nullptr;
    int *const cke = CHUNKED ? X.ck_exp + (size_t)slot * X.ck_rows * 8 : nullptr;
    if (!CHUNKED) octet_forward<NT, true>(F, sE, r, q, hrow, Lr, Lmax, Lmin, fw, fe, A, Aexp, ok);
    else {
      for (int c = 0; c < n_ck; c++) {
        octet_forward<NT, false>(F, sE, r, q, hrow, Lr, Lmax, Lmin, nullptr, nullptr, A, Aexp, ok, c * CK - 1, (c + 1) * CK - 1);
        if (c + 1 < n_ck) {
#pragma unroll
          for (int tp = 0; tp < NT; tp++) *reinterpret_cast<double2 *>(ckw + (size_t)(c + 1) * 8 * G + 8 * tp) = make_double2(A[2 * tp], A[2 * tp + 1]);
          if (q == 0) cke[(c + 1) * 8 + r] = Aexp;
        }
      }
    }
    double sfin = 0;
#pragma unroll
    for (int v = 0; v < NV; v++) sfin += A[v] * F.fin[8 * (v >> 1) + 2 * q + (v & 1)];
    sfin += __shfl_xor_sync(FULL, sfin, 1);
    sfin += __shfl_xor_sync(FULL, sfin, 2);
    const bool good = Lr > 0 && ok && sfin > 0;
    if (q == 0 && Lr > 0) {
      if (loglik) loglik[sid] = good ? log(sfin) + Aexp * 0.6931471805599453094 : JH_NEG_INF;
      if (!good) {  // decoded by the log-space kernel afterwards
        int pos = atomicAdd(fallback, 1);
        fallback[1 + pos] = sid;
      }
    }
    // ---- backward: beta only, decision per position ----
    double Bb[NV][NT];
#pragma unroll
    for (int k = 0; k < NV; k++)
#pragma unroll
      for (int t = 0; t < NT; t++) Bb[k][t] = F.T[(8 * t + r) * G + 8 * (k >> 1) + 2 * q + (k & 1)];
    double beta[NV];
    int Bx = 0;
#pragma unroll
    for (int v = 0; v < NV; v++) beta[v] = F.fin[8 * (v >> 1) + 2 * q + (v & 1)];
    int row_lo = 0;
    auto issue = [&](int row) {
      const bool have = row >= (CHUNKED ? row_lo : 0);
      char *sl = ring + (size_t)(row & (RING - 1)) * SLOTB;
      const double *src = fw + (int64_t)(CHUNKED ? row - row_lo : row) * 8 * G;
#pragma unroll
      for (int tp = 0; tp < NT; tp++) cp_async16_if(have, sl + rl_off[tp], src + 8 * tp);
      jhf::cp_async_commit();
    };
    __syncwarp();
    if (!CHUNKED) {
#pragma unroll
      for (int u = 0; u < C::PD; u++) issue(Lmax - 1 - u);
    }
    int h = hrow[(size_t)(Lmax - 1) * 8];
    int h1 = hrow[(size_t)(Lmax - 2) * 8];
    double e[NV];
    emis_rl<NT, true>(F, sE, q, e, h, Lmax - 1);
    for (int c = n_ck - 1; c >= 0; --c) {
    const int t0 = CHUNKED ? c * CK : 0, t1 = CHUNKED ? min(t0 + CK, Lmax) - 1 : Lmax - 1;
    if (CHUNKED) {  // forward rows t0 .. t1 of the chunk -> scratch (row t0 first)
      row_lo = t0;
      if (c > 0) {
#pragma unroll
        for (int tp = 0; tp < NT; tp++) {
          const double2 a2 = *reinterpret_cast<const double2 *>(ckw + (size_t)c * 8 * G + 8 * tp);
          A[2 * tp] = a2.x;
          A[2 * tp + 1] = a2.y;
        }
        Aexp = cke[c * 8 + r];
      } else Aexp = 0;
      bool ok2 = true;
      octet_forward<NT, true>(F, sE, r, q, hrow, Lr, Lmax, Lmin, fw - (int64_t)row_lo * 8 * G, fe - (int64_t)row_lo * 8, A, Aexp, ok2, t0 - 1, t1);
      __syncwarp();
#pragma unroll
      for (int u = 0; u < C::PD; u++) issue(t1 - u);
    }
    for (int t = t1; t >= t0; --t) {
      if ((t & (RS - 1)) == 0) prefetch_l1(hrow + (size_t)max(t - HPF, 0) * 8);
      const int h2 = t >= 2 ? hrow[(size_t)(t - 2) * 8] : 0;
      const bool act = t < Lr;
      double b[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) b[v] = e[v] * beta[v];
      jhf::cp_async_wait<C::PD - 1>();  // row t has landed
      __syncwarp();
      const char *slt = ring + (size_t)(t & (RING - 1)) * SLOTB;
      double Dv[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) Dv[v] = 0;
      if (t > 0) {
#pragma unroll
        for (int k = 0; k < NV; k++)
#pragma unroll
          for (int tp = 0; tp < NT; tp++) dmma(Dv[2 * tp], Dv[2 * tp + 1], b[k], Bb[k][tp]);
      }
      double Fa[NV];
#pragma unroll
      for (int tp = 0; tp < NT; tp++) {
        const double2 a2 = *reinterpret_cast<const double2 *>(slt + rl_off[tp]);
        Fa[2 * tp] = a2.x;
        Fa[2 * tp + 1] = a2.y;
      }
      const int arg = best_state(Fa, beta);
      if (q == 0 && act && good) path_store(paths, off + t, arg);
      __syncwarp();
      issue(t - C::PD);
      if (t > 0) {
        int nbx = Bx;
        rescale_row<NV>(Dv, nbx);
        if (t >= Lmin) {
#pragma unroll
          for (int v = 0; v < NV; v++) beta[v] = act ? Dv[v] : beta[v];
        } else {
#pragma unroll
          for (int v = 0; v < NV; v++) beta[v] = Dv[v];
        }
        h = h1;
        h1 = h2;
        emis_rl<NT, true>(F, sE, q, e, h, t - 1);
      }
    }
    jhf::cp_async_wait<0>();
    __syncwarp();
    }
  }
}

}  // namespace jhm
