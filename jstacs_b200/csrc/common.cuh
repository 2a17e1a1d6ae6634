// common.cuh — shared declarations of the B200-native HMM engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/jstacs_hmm.h"

#define JH_NEG_INF (-INFINITY)

// ------------------------------------------------------------------------------------------------
// error plumbing
void jh_set_error(const char *fmt, ...);
#define JH_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      jh_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      return JH_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

extern std::atomic<int64_t> g_kernel_launches;
#define JH_LAUNCHED() (g_kernel_launches.fetch_add(1, std::memory_order_relaxed))

// ------------------------------------------------------------------------------------------------
// Device view of the flattened model (all pointers are device memory, read-only during an E-step).
struct DevModel {
  int S, a, m, Cmax, n_elem, n_child, n_ctx_total, n_cond;
  int Kmax;   // maximal emission order over the emitting states
  int hpow;   // a^Kmax
  int hmod;   // a^(Kmax+1)
  int a_pow2; // 1 if a is a power of two (h % mod == h & (mod-1))
  double log_uniform;  // -log(a), AbstractConditionalDiscreteEmission.java:298
  const int *lc_ctx_ptr;      // [m+2]
  const int *ctx_elem;        // [n_ctx_total]
  const int *ctx_last;        // [n_ctx_total]
  const int *elem_child_ptr;  // [n_elem+1]
  const int *child_state;     // [n_child]
  const int *child_next;      // [(m+1)*n_child] local context index in layer class min(lc+1,m)
  const int *child_stat;      // [n_child] slot of the edge in the transition statistics (tying applied)
  const double *tscore;       // [n_child] parameters[c] - logNorm
  const int *in_ptr;          // [(m+1)*(Cmax+1)] in-edge CSR of the contexts of class min(lc+1,m), per source class lc
  const int *in_src;          // [(m+1)*n_child]  source context (ascending), then child index (ascending)
  const int *in_p;            // [(m+1)*n_child]  flat child index
  const int *state_order;     // [S] emission order of the state, -1 silent
  const int *state_tab;       // [S] first entry of the state's emission table in escore
  const int *state_mod;       // [S] a^(order+1)
  const double *escore;       // [n_cond*a] (0 - logNorm[cond]) + params[cond][sym]
  const unsigned char *is_final;  // [S]
  // allowed states groups (AbstractHMM.java:146-228): preComputedContext as a flag per (group, context of layer class m)
  const unsigned char *allowed;   // [n_groups * Cmax] or nullptr
  int n_groups;
};

// Device view of a packed data set.
struct DevData {
  const uint8_t *codes;     // padded by 32 bytes at the end
  const int64_t *offsets;   // [n+1]
  const double *weights;    // [n] or nullptr
  const int32_t *order;     // [n] sequence ids sorted by decreasing length (length bucketing)
  int64_t n_seq;
  const int32_t *groups;    // [n_res] allowedStatesGroup of every position (-1: unconstrained), nullptr: no constraint
};

// HigherOrderHMM.getAllowedContext (HigherOrderHMM.java:331-340) as a membership test: contexts of layer l (l >= m) whose
// last state emitted the absolute position pos - 1 are restricted to preComputedContext[group[pos - 1]].
// grp: groups of the sequence (indexed by absolute position) or nullptr.
__device__ __forceinline__ bool ctx_allowed(const DevModel &M, const int32_t *grp, int64_t pos, int64_t layer, int c) {
  if (!grp || !M.allowed || layer < M.m) return true;
  const int g = grp[pos - 1];
  return g < 0 || M.allowed[(size_t)g * M.Cmax + c] != 0;
}

struct WorkList {   // sequences, or windows of sequences
  const int32_t *seq;     // [n] sequence ids (nullptr: D.order)
  const int64_t *start;   // [n] first position (nullptr: 0)
  const int64_t *end;     // [n] last position inclusive (nullptr: L-1)
  int64_t n;
};

// Sink of a state path: int32 per residue (IntList of the reference) or one byte per residue (255 = no state), decided by
// the entry point (jh_viterbi / jh_viterbi_u8, ...): the kernels emit the narrow form directly, nothing is converted afterwards.
struct PathOut {
  void *p;
  int u8;
  __host__ __device__ explicit operator bool() const { return p != nullptr; }
};
__device__ __forceinline__ void path_store(const PathOut &o, int64_t idx, int v) {
  if (o.u8) reinterpret_cast<uint8_t *>(o.p)[idx] = (uint8_t)v;  // -1 -> 255
  else reinterpret_cast<int32_t *>(o.p)[idx] = v;
}

__device__ __forceinline__ double lse_finish(double offset, double sum) { return offset + log(sum); }
