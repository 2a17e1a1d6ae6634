// jstacs_hmm.cu — C ABI (include/jstacs_hmm.h) of the B200-native HMM engine: handles, dispatch, EM driver.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (see __graft_entry__.py).
#include <dlfcn.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>

#include "common.cuh"
#include "kernels_aux.cuh"
#include "kernels_exact.cuh"
#include "kernels_fast.cuh"
#include "kernels_general.cuh"
#include "kernels_mma.cuh"
#include "kernels_lanes.cuh"
#include "kernels_sparse.cuh"
#include "kernels_scan.cuh"
#include "kernels_lanechunks.cuh"
#include "kernels_viterbi.cuh"
#include "kernels_vitlong.cuh"
#include "kernels_band.cuh"

// ------------------------------------------------------------------------------------------------
// Process-wide state is limited to DEFAULTS for handles created afterwards (device, options) and to the launch counter;
// everything a call works with (device, stream, options, scratch) lives in the handle, so two models on two devices,
// or two threads with different options, do not interfere (a handle is still used by one thread at a time).
static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_kernel_launches{0};
static std::atomic<int> g_default_device{0};
static std::mutex g_mu;                      // guards the option defaults, the legacy stream default and the stagers
static cudaStream_t g_user_stream = nullptr;  // jh_set_stream (legacy): default stream of handles on g_user_stream_device
static bool g_have_user_stream = false;
static int g_user_stream_device = 0;

struct Options {
  int64_t fast = 1;            // use the scaled linear-space kernels where the model allows it
  int64_t scratch_mb = 0;      // upper bound of DP scratch memory; 0 = automatic: 55 % of the free device memory, at most 96 GB
  int64_t blocks_per_sm = 0;   // 0 = occupancy API
  int64_t fast_variant = 0;
  int64_t chunk = 2048;        // positions per checkpoint of the sparse decoding path
  int64_t split_short = 1;     // 2..4 states, few short sequences: cut them into chunks for the position-parallel path
  int64_t lane_chunks = 1;     // 2..8 states: chunk pass with one chunk per lane (kernels_lanechunks.cuh)
  int64_t scan = 1;            // small dense models: checkpoint pass parallel in the position (kernels_scan.cuh)
  int64_t overlap = 1;         // long-sequence path: per-sequence streams, a sequence's chunk pass starts when its checkpoint pass ends
  int64_t pass1_cta = 1;       // few long sequences: one CTA per (sequence, direction) in the checkpoint pass
  int64_t oct_chunk = -1;      // positions per checkpoint of the chunked octet kernels (0: never chunk; < 0: automatic, see oct_chunk_of)
  int64_t oct_chunk_force = 0; // tests: chunk every bucket longer than 2 * oct_chunk
  int64_t graph = 1;           // jh_baum_welch with IterationCondition: replay the EM iteration as a CUDA graph
  int64_t vit_chunk = 4096;    // positions per checkpoint of the chromosome-scale Viterbi
  int64_t vit_chunk_force = 0; // tests: checkpointed Viterbi for every sequence longer than 2 * vit_chunk
  int64_t stage = 1;           // host transfers of pageable caller memory through the library's pinned double buffers
  int64_t band = 1;            // banded models: warp-resident whole-sequence passes (kernels_band.cuh)
  int64_t oct_full_rows_mb = 24576;  // octet kernels: buckets whose full forward rows (whole grid) exceed this are chunked
  int64_t vit_train_fast = 1;  // Viterbi training: decode with the fast kernels, then count along the paths (0: reference-order kernel)
  int64_t band_meet = 1;       // banded models, scoring of long sequences: forward and backward chains meet in the middle
  int64_t deterministic = 0;   // Baum-Welch E-step in a fixed order (bit-identical statistics run to run): octet kernels only (5..32 states)
};
static Options g_defaults;
static std::atomic<uint64_t> g_opt_gen{1};

// returns false for an unknown name
static bool set_option_field(Options &o, const std::string &n, int64_t value) {
  if (n == "fast") o.fast = value;
  else if (n == "scratch_mb") o.scratch_mb = value;
  else if (n == "blocks_per_sm") o.blocks_per_sm = value;
  else if (n == "fast_variant") o.fast_variant = value;
  else if (n == "oct_chunk") o.oct_chunk = value < 0 ? -1 : value == 0 ? 0 : std::min<int64_t>(1 << 20, std::max<int64_t>(8, value));
  else if (n == "pass1_cta") o.pass1_cta = value;
  else if (n == "overlap") o.overlap = value;
  else if (n == "scan") o.scan = value;
  else if (n == "lane_chunks") o.lane_chunks = value;
  else if (n == "split_short") o.split_short = value;
  else if (n == "oct_chunk_force") o.oct_chunk_force = value;
  else if (n == "chunk") o.chunk = std::min<int64_t>(1 << 20, std::max<int64_t>(16, value));
  else if (n == "graph") o.graph = value;
  else if (n == "vit_chunk") o.vit_chunk = std::min<int64_t>(1 << 20, std::max<int64_t>(16, value));
  else if (n == "vit_chunk_force") o.vit_chunk_force = value;
  else if (n == "stage") o.stage = value;
  else if (n == "band") o.band = value;
  else if (n == "oct_full_rows_mb") o.oct_full_rows_mb = value;
  else if (n == "deterministic") o.deterministic = value;
  else if (n == "band_meet") o.band_meet = value;
  else if (n == "vit_train_fast") o.vit_train_fast = value;
  else return false;
  return true;
}

void jh_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Device buffers come from the device's stream-ordered memory pool with an unlimited release threshold: a data set that
// is created and destroyed per call (the e2e path: pack -> upload -> train -> free) reuses the pool's memory instead of
// paying cudaFree's unmap (measured: up to 360 ms for the 3 GB of a 1M x 1 kb data set).  Allocation and release are
// ordered on the stream of the call that is running (tl_stream, set by the entry guard of every API function): nothing
// waits for the device, and work of other handles on other streams is not stalled.
static thread_local cudaStream_t tl_stream = nullptr;
static thread_local bool tl_have_stream = false;
static void pool_setup(int dev) {
  static std::atomic<uint64_t> done{0};
  if (dev < 0 || dev > 63 || (done.load() >> dev) & 1) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done.fetch_or(1ull << dev);
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaStream_t home = nullptr;  // stream the buffer was allocated on
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFreeAsync(p, tl_have_stream ? tl_stream : home);  // ordered after the work already queued on that stream
    p = nullptr;
    n = 0;
  }
  int alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    home = tl_stream;
    if (cudaMallocAsync((void **)&p, count * sizeof(T), home) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      jh_set_error("out of device memory allocating %zu bytes", count * sizeof(T));
      return JH_ERR_NOMEM;
    }
    n = count;
    return JH_OK;
  }
  int reserve(size_t count) { return count <= n ? (int)JH_OK : alloc(count); }
  int upload(const std::vector<T> &h, cudaStream_t s) {
    int r = alloc(h.size());
    if (r) return r;
    if (!h.empty()) JH_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return JH_OK;
  }
};

// ------------------------------------------------------------------------------------------------
// Host transfers.  Callers hand over ordinary (pageable) memory — a Java off-heap segment, a numpy array — which the
// driver would copy through its own small bounce buffer at ~5 GB/s (measured: 1 GB of byte paths = 210 ms).  The library
// stages such transfers through its own PINNED double buffers: a few host threads copy slice k+1 between the caller's
// memory and a pinned buffer while the DMA engine moves slice k, so the transfer runs at the host's memcpy rate or at
// the PCIe rate, whichever is lower.  Memory that is already pinned (cudaHostAlloc / cudaHostRegister) is copied directly.
class CopyPool {  // persistent helper threads for slice-parallel memcpy
 public:
  explicit CopyPool(int n) {
    for (int i = 0; i < n; i++) th_.emplace_back([this, i] { run(i); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  void copy(void *dst, const void *src, size_t bytes) {
    const int parts = (int)th_.size() + 1;
    if (bytes < (size_t)1 << 20 || parts == 1) {
      memcpy(dst, src, bytes);
      return;
    }
    const size_t slice = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> g(mu_);
      dst_ = (char *)dst; src_ = (const char *)src; bytes_ = bytes; slice_ = slice;
      pending_ = (int)th_.size();
      gen_++;
    }
    cv_.notify_all();
    const size_t o = slice * th_.size();  // the caller copies the last slice itself
    if (o < bytes) memcpy((char *)dst + o, (const char *)src + o, bytes - o);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void run(int i) {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return gen_ != seen; });
      seen = gen_;
      if (stop_) return;
      char *d = dst_;
      const char *s = src_;
      const size_t o = slice_ * i, n = o < bytes_ ? std::min(slice_, bytes_ - o) : 0;
      lk.unlock();
      if (n) memcpy(d + o, s + o, n);
      lk.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  uint64_t gen_ = 0;
  bool stop_ = false;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
  int pending_ = 0;
};

struct Stager {
  static constexpr size_t CHUNK = (size_t)16 << 20;
  static constexpr int NB = 3;
  void *buf[NB] = {};
  cudaEvent_t ev[NB] = {};
  bool ready = false;
  int init() {
    if (ready) return JH_OK;
    for (int i = 0; i < NB; i++) {
      JH_CUDA(cudaHostAlloc(&buf[i], CHUNK, cudaHostAllocDefault));
      JH_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    ready = true;
    return JH_OK;
  }
};
static Stager g_stager[64];
static std::mutex g_stage_mu;
static CopyPool &copy_pool() {
  static CopyPool pool((int)std::min<unsigned>(7, std::max<unsigned>(1, std::thread::hardware_concurrency() / 2)) );
  return pool;
}
static bool is_pinned(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
// host -> device; returns when the host buffer may be reused (the device copy is ordered on s)
static int stage_h2d(void *dst, const void *src, size_t bytes, cudaStream_t s, int dev, bool stage) {
  if (bytes == 0) return JH_OK;
  if (!stage || bytes < ((size_t)4 << 20) || is_pinned(src)) {
    JH_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    JH_CUDA(cudaStreamSynchronize(s));
    return JH_OK;
  }
  std::lock_guard<std::mutex> g(g_stage_mu);
  Stager &S = g_stager[dev & 63];
  int r = S.init();
  if (r) return r;
  CopyPool &P = copy_pool();
  size_t k = 0;
  for (size_t o = 0; o < bytes; o += Stager::CHUNK, k++) {
    const int b = (int)(k % Stager::NB);
    const size_t n = std::min(Stager::CHUNK, bytes - o);
    if (k >= Stager::NB) JH_CUDA(cudaEventSynchronize(S.ev[b]));
    P.copy(S.buf[b], (const char *)src + o, n);
    JH_CUDA(cudaMemcpyAsync((char *)dst + o, S.buf[b], n, cudaMemcpyHostToDevice, s));
    JH_CUDA(cudaEventRecord(S.ev[b], s));
  }
  JH_CUDA(cudaStreamSynchronize(s));
  return JH_OK;
}
// device -> host, ordered after the work queued on s; returns when dst holds the data
static int stage_d2h(void *dst, const void *src, size_t bytes, cudaStream_t s, int dev, bool stage) {
  if (bytes == 0) return JH_OK;
  if (!stage || bytes < ((size_t)4 << 20) || is_pinned(dst)) {
    JH_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    JH_CUDA(cudaStreamSynchronize(s));
    return JH_OK;
  }
  std::lock_guard<std::mutex> g(g_stage_mu);
  Stager &S = g_stager[dev & 63];
  int r = S.init();
  if (r) return r;
  CopyPool &P = copy_pool();
  const size_t n_chunks = (bytes + Stager::CHUNK - 1) / Stager::CHUNK;
  auto drain = [&](size_t k) -> int {  // chunk k: pinned -> caller memory
    const int b = (int)(k % Stager::NB);
    const size_t o = k * Stager::CHUNK, n = std::min(Stager::CHUNK, bytes - o);
    JH_CUDA(cudaEventSynchronize(S.ev[b]));
    P.copy((char *)dst + o, S.buf[b], n);
    return JH_OK;
  };
  for (size_t k = 0; k < n_chunks; k++) {
    const int b = (int)(k % Stager::NB);
    const size_t o = k * Stager::CHUNK, n = std::min(Stager::CHUNK, bytes - o);
    if (k >= Stager::NB && (r = drain(k - Stager::NB))) return r;  // frees buffer b
    JH_CUDA(cudaMemcpyAsync(S.buf[b], (const char *)src + o, n, cudaMemcpyDeviceToHost, s));
    JH_CUDA(cudaEventRecord(S.ev[b], s));
  }
  for (size_t k = n_chunks > Stager::NB ? n_chunks - Stager::NB : 0; k < n_chunks; k++)
    if ((r = drain(k))) return r;
  return JH_OK;
}

// ------------------------------------------------------------------------------------------------
struct Id128 {  // ncclUniqueId (128 bytes, passed by value)
  char b[128];
};
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, Id128, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.lib) return JH_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    jh_set_error("libnccl.so.2 not found: %s", dlerror());
    return JH_ERR_COMM;
  }
  g_nccl.GetUniqueId = (int (*)(void *))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void **, int, Id128, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.Broadcast = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(g_nccl.lib, "ncclBroadcast");
  g_nccl.CommDestroy = (int (*)(void *))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    jh_set_error("libnccl.so.2 lacks a required symbol");
    g_nccl.lib = nullptr;
    return JH_ERR_COMM;
  }
  return JH_OK;
}

// ------------------------------------------------------------------------------------------------
struct jh_model {
  // host IR
  int S = 0, a = 0, m = 0, Cmax = 0, n_elem = 0, n_child = 0, n_ctx_total = 0, n_cond = 0, n_emis = 0, Kmax = 0;
  bool has_silent = false;
  int max_children = 0;
  std::vector<int> lc_ctx_ptr, ctx_elem, ctx_last, elem_child_ptr, child_state, child_desc, elem_ctx, trans_index;
  std::vector<int> state_emis, emis_order, emis_cond_ptr;
  std::vector<uint8_t> state_silent, state_final;
  std::vector<int> child_elem, child_next, child_stat, in_ptr, in_src, in_p, state_order, state_tab, state_mod, row_flag;
  std::vector<int> sin_ptr, sin_src, sin_p;  // silent in-edges inside a layer (kernels_general.cuh)
  // device
  DevBuf<int> d_lc_ctx_ptr, d_ctx_elem, d_ctx_last, d_elem_child_ptr, d_child_state, d_child_next, d_child_stat, d_in_ptr,
      d_in_src, d_in_p, d_state_order, d_state_tab, d_state_mod, d_child_elem, d_trans_index, d_row_flag;
  DevBuf<unsigned char> d_is_final, d_silent, d_allowed;
  int n_groups = 0;
  DevBuf<int> d_sin_ptr, d_sin_src, d_sin_p;
  DevBuf<double> d_rows;       // rolling DP rows of the thread-per-sequence kernels
  DevBuf<double> d_trans_param, d_trans_lognorm, d_trans_hyper, d_emis_param, d_emis_lognorm, d_emis_hyper, d_tscore, d_escore;
  DevBuf<double> d_stats;      // [n_child | n_cond*a | score | pad]
  DevBuf<double> d_objective;  // per-iteration objective
  DevBuf<double> d_fwd;        // forward scratch
  DevBuf<uint8_t> d_choice;    // Viterbi choice scratch
  DevBuf<unsigned long long> d_counter;
  DevBuf<double> d_det;        // deterministic mode: per-slot partial sums + fold scratch
  DevBuf<uint8_t> d_vt_paths;  // Viterbi training on the fast decoders: byte paths and scores of the data set
  DevBuf<double> d_vt_scores;
  DevBuf<double> d_meet_vec;   // scoring that meets in the middle: [n][2][SP] vectors, [n][2] exponents and flags
  DevBuf<long long> d_meet_exp;
  DevBuf<int32_t> d_meet_ok;
  DevBuf<int32_t> d_fallback;  // [1 + n] list of sequences the fast path handed to the exact kernels
  jhf::FastTables fast;        // linear-space tables of the fast path
  jhs::SparseTables sparse;    // edge lists of the sparse path (33..256 states)
  DevBuf<double> d_ck_alpha, d_ck_beta;  // checkpoint vectors of chromosome-scale decoding
  DevBuf<long long> d_ck_exp, d_seq_AL;  // E-step: exponents of the checkpoints [2][n_chunks+1], of P(x) [n_seq]
  DevBuf<double> d_seq_sfin;             // E-step: mantissa of P(x) [n_seq]
  DevBuf<int> d_rexp;                    // E-step: exponents of the recomputed forward rows [slot][chunk]
  DevBuf<double> d_scanR;                // chunk transfer rows of the parallel checkpoint pass (kernels_scan.cuh)
  DevBuf<int> d_scanE;
  DevBuf<double> d_vit_ck;               // chromosome-scale Viterbi: rows bwd[c*CK] (kernels_vitlong.cuh)
  DevBuf<uint8_t> d_vit_map, d_vit_entry, d_vit_nxt;
  DevBuf<unsigned long long> d_counters; // one work counter per concurrent chunk kernel
  std::vector<cudaStream_t> seq_streams; // one stream per long sequence: its chunk pass starts when ITS checkpoint pass is done
  std::vector<cudaEvent_t> seq_events;
  cudaEvent_t ev_fork = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
  bool timed = false, ktimed = false;
  int sm_count = 148;
  void *comm = nullptr;
  int n_ranks = 1, rank = 0;
  // per-handle execution context
  int device = 0;
  cudaStream_t user_stream = nullptr;
  bool have_user_stream = false;
  Options opt;                                             // defaults (jh_set_option) + overrides (jh_model_set_option)
  uint64_t opt_seen = 0;
  int64_t auto_scratch_mb = 0;
  std::vector<std::pair<std::string, int64_t>> overrides;
  int failed = 0;                                          // multi-rank: error of this rank's E-step, returned after the collectives
  char failed_msg[1024] = "";
  bool capturing = false;                                  // stream capture of the EM iteration in progress: no timing events
  double last_objective = JH_NEG_INF;                      // final objective of the last jh_baum_welch
  DevBuf<int32_t> d_iter;                                  // device-side iteration counter of the graph-replayed EM loop
  cudaGraphExec_t em_graph = nullptr;
  const void *em_graph_data = nullptr;
  int em_graph_mode = -1;
  int64_t n_stats() const { return (int64_t)n_child + (int64_t)n_cond * a; }
  cudaStream_t st() const {
    if (have_user_stream) return user_stream;
    if (g_have_user_stream && g_user_stream_device == device) return g_user_stream;
    return stream;
  }
  void refresh_options() {
    const uint64_t gen = g_opt_gen.load();
    if (gen == opt_seen) return;
    std::lock_guard<std::mutex> g(g_mu);
    opt = g_defaults;
    for (auto &kv : overrides) set_option_field(opt, kv.first, kv.second);
    if (opt.scratch_mb <= 0) {  // automatic budget, sized once per model from what the device has free (180 GB HBM3e on a B200)
      if (auto_scratch_mb <= 0) {
        size_t free_b = 0, total_b = 0;
        cudaSetDevice(device);
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)32 << 30; }
        auto_scratch_mb = std::max<int64_t>(1024, std::min<int64_t>(96 * 1024, (int64_t)(free_b >> 20) * 55 / 100));
      }
      opt.scratch_mb = auto_scratch_mb;
    }
    opt_seen = gen;
  }
  DevModel dev() const {
    DevModel M;
    M.S = S; M.a = a; M.m = m; M.Cmax = Cmax; M.n_elem = n_elem; M.n_child = n_child; M.n_ctx_total = n_ctx_total; M.n_cond = n_cond;
    M.Kmax = Kmax;
    int pw = 1;
    for (int i = 0; i < Kmax; i++) pw *= a;
    M.hpow = pw; M.hmod = pw * a;
    M.a_pow2 = (a & (a - 1)) == 0;
    M.log_uniform = -log((double)a);
    M.lc_ctx_ptr = d_lc_ctx_ptr.p; M.ctx_elem = d_ctx_elem.p; M.ctx_last = d_ctx_last.p; M.elem_child_ptr = d_elem_child_ptr.p;
    M.child_state = d_child_state.p; M.child_next = d_child_next.p; M.child_stat = d_child_stat.p; M.tscore = d_tscore.p;
    M.in_ptr = d_in_ptr.p; M.in_src = d_in_src.p; M.in_p = d_in_p.p; M.state_order = d_state_order.p; M.state_tab = d_state_tab.p;
    M.state_mod = d_state_mod.p; M.escore = d_escore.p; M.is_final = d_is_final.p;
    M.allowed = n_groups > 0 ? d_allowed.p : nullptr; M.n_groups = n_groups;
    return M;
  }
  jhg::GenModel gen() const {
    jhg::GenModel X;
    X.silent = d_silent.p; X.sin_ptr = d_sin_ptr.p; X.sin_src = d_sin_src.p; X.sin_p = d_sin_p.p;
    return X;
  }
  jha::DevParams params() const {
    jha::DevParams P;
    P.n_elem = n_elem; P.n_child = n_child; P.n_cond = n_cond; P.a = a; P.S = S;
    P.elem_child_ptr = d_elem_child_ptr.p; P.child_elem = d_child_elem.p; P.trans_index = d_trans_index.p; P.row_emitting = d_row_flag.p;
    P.trans_param = d_trans_param.p; P.trans_lognorm = d_trans_lognorm.p; P.emis_param = d_emis_param.p; P.emis_lognorm = d_emis_lognorm.p;
    P.trans_hyper = d_trans_hyper.p; P.emis_hyper = d_emis_hyper.p; P.tscore = d_tscore.p; P.escore = d_escore.p;
    return P;
  }
};

struct jh_dataset {
  int64_t n_seq = 0, n_res = 0, max_len = 0, min_len = 0;
  int alphabet = 0;
  int device = 0;
  cudaStream_t stream = nullptr;   // uploads of this data set
  cudaEvent_t ev_last = nullptr;   // recorded after every call that reads the data set: destroy waits for it, not for the device
  std::vector<int64_t> offsets;
  std::vector<int32_t> order;
  // length buckets over `order` (decreasing length): [bucket_ptr[b], bucket_ptr[b+1]) with max length bucket_len[b]
  std::vector<int64_t> bucket_ptr, bucket_len;
  DevBuf<uint8_t> d_codes;
  DevBuf<int64_t> d_offsets;
  DevBuf<int32_t> d_order;
  DevBuf<double> d_weights;
  bool has_weights = false;
  double weighted_total = 0;   // sum_n w_n * L_n of the current weights (host side; bounds every count cell)
  DevBuf<int32_t> d_groups;   // allowedStatesGroup of every position (jh_dataset_set_groups)
  bool has_groups = false;
  // octets: 8 consecutive sequences of a bucket advance in lock step through one warp (kernels_mma.cuh);
  // bucket b owns octets [bucket_oct_ptr[b], bucket_oct_ptr[b+1]); hash records are built lazily per (alphabet, Kmax)
  std::vector<int32_t> oct_seq;
  std::vector<int64_t> oct_hbase, bucket_oct_ptr;
  DevBuf<int32_t> d_oct_seq;
  DevBuf<int64_t> d_oct_hbase;
  DevBuf<uint16_t> d_hashes;
  int hash_a = -1, hash_K = -1;
  // chunk list of the sparse decoding path (kernels_sparse.cuh), built per chunk length
  int64_t chunk_len = 0, n_chunks = 0, n_long = 0;
  std::vector<int64_t> long_seg;  // [n_long+1] first chunk (in the chunk list) of the k-th long sequence; [n_long] = first single-chunk sequence
  DevBuf<int64_t> d_chunk_ptr;
  DevBuf<int32_t> d_chunk_seq, d_chunk_idx, d_long;  // d_long: sequences of more than one chunk, longest first
  DevBuf<int64_t> d_long_seg;
  DevData dev(int64_t first = 0, int64_t count = -1) const {
    DevData D;
    D.codes = d_codes.p; D.offsets = d_offsets.p; D.weights = has_weights ? d_weights.p : nullptr;
    D.order = d_order.p + first; D.n_seq = count < 0 ? n_seq : count;
    D.groups = has_groups ? d_groups.p : nullptr;
    return D;
  }
};

// ------------------------------------------------------------------------------------------------
extern "C" int jh_abi_version(void) { return JH_ABI_VERSION; }
extern "C" const char *jh_last_error(void) { return g_err; }
extern "C" int jh_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    jh_set_error("no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(cudaGetLastError()));
    return JH_ERR_CUDA;
  }
  return n;
}
extern "C" int jh_set_device(int device) {
  JH_CUDA(cudaSetDevice(device));
  g_default_device = device;
  return JH_OK;
}
extern "C" int64_t jh_kernel_launches(void) { return g_kernel_launches.load(); }
extern "C" int jh_set_stream(void *s) {
  std::lock_guard<std::mutex> g(g_mu);
  g_user_stream = (cudaStream_t)s;
  g_have_user_stream = s != nullptr;
  g_user_stream_device = g_default_device;
  return JH_OK;
}
extern "C" int jh_synchronize(void) {
  JH_CUDA(cudaDeviceSynchronize());
  return JH_OK;
}
extern "C" int jh_set_option(const char *name, int64_t value) {
  std::string n(name ? name : "");
  if (n == "trim_pool") {  // give the unused memory of the device pool back to the driver (keep `value` bytes)
    cudaMemPool_t pool;
    JH_CUDA(cudaSetDevice(g_default_device));
    JH_CUDA(cudaDeviceSynchronize());
    JH_CUDA(cudaDeviceGetDefaultMemPool(&pool, g_default_device));
    JH_CUDA(cudaMemPoolTrimTo(pool, (size_t)std::max<int64_t>(0, value)));
    return JH_OK;
  }
  std::lock_guard<std::mutex> g(g_mu);
  if (!set_option_field(g_defaults, n, value)) {
    jh_set_error("unknown option '%s'", n.c_str());
    return JH_ERR_INVALID;
  }
  g_opt_gen++;
  return JH_OK;
}
extern "C" int jh_model_set_option(jh_model *m, const char *name, int64_t value) {
  if (!m) { jh_set_error("null model"); return JH_ERR_INVALID; }
  std::string n(name ? name : "");
  Options probe;
  if (!set_option_field(probe, n, value)) {
    jh_set_error("unknown option '%s'", n.c_str());
    return JH_ERR_INVALID;
  }
  bool found = false;
  for (auto &kv : m->overrides)
    if (kv.first == n) { kv.second = value; found = true; }
  if (!found) m->overrides.emplace_back(n, value);
  m->opt_seen = 0;
  return JH_OK;
}
extern "C" int jh_model_set_stream(jh_model *m, void *s) {
  if (!m) { jh_set_error("null model"); return JH_ERR_INVALID; }
  m->user_stream = (cudaStream_t)s;
  m->have_user_stream = s != nullptr;
  return JH_OK;
}

static int ensure_device(int device) {
  int n = jh_device_count();
  if (n <= 0) {
    if (n == 0) jh_set_error("no CUDA device visible (there is no CPU fallback)");
    return JH_ERR_CUDA;
  }
  JH_CUDA(cudaSetDevice(device));
  pool_setup(device);
  return JH_OK;
}

// Entry guard of every API call on a model: the handle's device becomes current, its options are refreshed, buffers that
// the call allocates or frees are ordered on the handle's stream.
struct Scope {
  bool prev_have;
  cudaStream_t prev;
  Scope(int device, cudaStream_t s) : prev_have(tl_have_stream), prev(tl_stream) {
    cudaSetDevice(device);
    tl_stream = s;
    tl_have_stream = true;
  }
  ~Scope() {
    tl_stream = prev;
    tl_have_stream = prev_have;
  }
};
#define JH_ENTER(m)                                              \
  if (!(m)) { jh_set_error("null model"); return JH_ERR_INVALID; } \
  (m)->refresh_options();                                        \
  Scope scope__((m)->device, (m)->st())

// ------------------------------------------------------------------------------------------------
static int prep_scores(jh_model *m) {
  int n = std::max(m->n_child, m->n_cond * m->a);
  jha::k_prep_scores<<<(n + 255) / 256, 256, 0, m->st()>>>(m->params());
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  if (m->sparse.usable) {
    jhs::SparseTables &P = m->sparse;
    int np = std::max(P.IND * P.SP, P.H * P.SP);
    jhs::k_sparse_prepare<<<(np + 255) / 256, 256, 0, m->st()>>>(P.dev(), P.IND, P.d_in_p, P.d_out_p, P.d_pimap, P.d_etab, P.d_emod, m->d_tscore.p,
                                                               m->d_escore.p, P.d_in_T, P.d_out_T, P.d_pi, P.d_E);
    JH_LAUNCHED();
    jhw::k_vitlong_prepare<<<(np + 255) / 256, 256, 0, m->st()>>>(P.dev(), P.IND, P.d_out_p, P.d_pimap, P.d_etab, P.d_emod, m->d_tscore.p, m->d_escore.p,
                                                                P.d_out_lT, P.d_lpi, P.d_lE);
    JH_LAUNCHED();
    JH_CUDA(cudaGetLastError());
  }
  int r = jhf::prepare(m->fast, m->dev(), m->st());
  if (r || !m->fast.usable) return r;
  jhf::FastTables &F = m->fast;
  int nv = std::max(F.G * F.G, F.H * F.G);
  jhv::k_viterbi_prepare<<<(nv + 255) / 256, 256, 0, m->st()>>>(F.dev(), F.d_tmap, F.d_pimap, m->d_tscore.p, m->d_escore.p, F.d_vT, F.d_vpi, F.d_vfin, F.d_vE);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

static void model_free(jh_model *m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  if (m->have_user_stream) cudaStreamSynchronize(m->user_stream);
  for (cudaStream_t st : m->seq_streams) cudaStreamSynchronize(st);
  Scope scope(m->device, cudaStreamPerThread);  // everything queued has finished: the buffers go back on a stream that outlives the handle
  if (m->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
  if (m->em_graph) cudaGraphExecDestroy(m->em_graph);
  jhf::destroy(m->fast);
  jhs::destroy(m->sparse);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->evk0) cudaEventDestroy(m->evk0);
  if (m->evk1) cudaEventDestroy(m->evk1);
  for (cudaStream_t st : m->seq_streams) cudaStreamDestroy(st);
  for (cudaEvent_t ev : m->seq_events) cudaEventDestroy(ev);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

extern "C" int jh_model_create(const jh_model_desc *d, jh_model **out) {
  if (!d || !out) { jh_set_error("null argument"); return JH_ERR_INVALID; }
  if (d->abi_version != JH_ABI_VERSION) { jh_set_error("jh_model_desc.abi_version %d != %d", d->abi_version, JH_ABI_VERSION); return JH_ERR_INVALID; }
  if (d->n_states < 1 || d->alphabet < 2 || d->alphabet > 255 || d->max_order < 0 || d->n_elem < 1 || d->n_emis < 1) {
    jh_set_error("invalid model sizes");
    return JH_ERR_INVALID;
  }
  const int device = g_default_device;
  int r = ensure_device(device);
  if (r) return r;
  jh_model *m = new jh_model();
  m->device = device;
  m->refresh_options();
  m->S = d->n_states; m->a = d->alphabet; m->m = d->max_order; m->n_elem = d->n_elem; m->n_emis = d->n_emis;
  m->lc_ctx_ptr.assign(d->lc_ctx_ptr, d->lc_ctx_ptr + m->m + 2);
  m->n_ctx_total = m->lc_ctx_ptr[m->m + 1];
  for (int lc = 0; lc <= m->m; lc++) m->Cmax = std::max(m->Cmax, m->lc_ctx_ptr[lc + 1] - m->lc_ctx_ptr[lc]);
  m->ctx_elem.assign(d->ctx_elem, d->ctx_elem + m->n_ctx_total);
  m->ctx_last.assign(d->ctx_last_state, d->ctx_last_state + m->n_ctx_total);
  m->elem_child_ptr.assign(d->elem_child_ptr, d->elem_child_ptr + m->n_elem + 1);
  m->n_child = m->elem_child_ptr[m->n_elem];
  m->child_state.assign(d->child_state, d->child_state + m->n_child);
  m->child_desc.assign(d->child_desc, d->child_desc + m->n_child);
  m->elem_ctx.assign(d->elem_ctx, d->elem_ctx + (size_t)(m->m + 1) * m->n_elem);
  m->trans_index.assign(d->trans_index, d->trans_index + m->n_elem);
  m->state_emis.assign(d->state_emis, d->state_emis + m->S);
  m->emis_order.assign(d->emis_order, d->emis_order + m->n_emis);
  m->emis_cond_ptr.assign(d->emis_cond_ptr, d->emis_cond_ptr + m->n_emis + 1);
  m->n_cond = m->emis_cond_ptr[m->n_emis];
  m->state_silent.assign(d->state_silent, d->state_silent + m->S);
  m->state_final.assign(d->state_final, d->state_final + m->S);
  auto fail = [&](int code, const char *msg) {
    jh_set_error("%s", msg);
    model_free(m);
    return code;
  };
  // validation (the reference checks these in BasicHigherOrderTransition.init and the AbstractHMM constructor)
  for (int t = 0; t < m->n_elem; t++) {
    if (m->trans_index[t] < 0 || m->trans_index[t] > t) return fail(JH_ERR_INVALID, "transIndex constraint 0<=transIndex[i]<=i violated");
    int n = m->elem_child_ptr[t + 1] - m->elem_child_ptr[t];
    if (n < 0) return fail(JH_ERR_INVALID, "elem_child_ptr must be non-decreasing");
    int rn = m->elem_child_ptr[m->trans_index[t] + 1] - m->elem_child_ptr[m->trans_index[t]];
    if (rn != n) return fail(JH_ERR_INVALID, "tied TransitionElements don't have the same number of children");
    m->max_children = std::max(m->max_children, n);
  }
  for (int p = 0; p < m->n_child; p++) {
    if (m->child_state[p] < 0 || m->child_state[p] >= m->S) return fail(JH_ERR_INVALID, "child state out of range");
    if (m->child_desc[p] < 0 || m->child_desc[p] >= m->n_elem) return fail(JH_ERR_INVALID, "descendant element out of range");
  }
  for (int s = 0; s < m->S; s++) {
    if (m->state_emis[s] < 0 || m->state_emis[s] >= m->n_emis) return fail(JH_ERR_INVALID, "emission index out of range");
    if (m->state_silent[s]) m->has_silent = true;
    if ((m->emis_order[m->state_emis[s]] < 0) != (m->state_silent[s] != 0)) return fail(JH_ERR_INVALID, "state_silent does not match the emission");
  }
  // derived tables -----------------------------------------------------------------------------
  m->child_elem.resize(m->n_child);
  m->child_stat.resize(m->n_child);
  for (int t = 0; t < m->n_elem; t++)
    for (int p = m->elem_child_ptr[t]; p < m->elem_child_ptr[t + 1]; p++) {
      m->child_elem[p] = t;
      m->child_stat[p] = m->elem_child_ptr[m->trans_index[t]] + (p - m->elem_child_ptr[t]);
    }
  m->child_next.assign((size_t)(m->m + 1) * m->n_child, -1);
  for (int lc = 0; lc <= m->m; lc++) {
    int lct = std::min(lc + 1, m->m);
    for (int p = 0; p < m->n_child; p++) {
      int adv = m->state_silent[m->child_state[p]] ? 0 : 1;
      m->child_next[(size_t)lc * m->n_child + p] = m->elem_ctx[(size_t)(adv ? lct : lc) * m->n_elem + m->child_desc[p]];
    }
  }
  // in-edge CSR per source layer class (forward recursion in pull form; order = source context, then child)
  m->in_ptr.assign((size_t)(m->m + 1) * (m->Cmax + 1), 0);
  m->in_src.assign((size_t)(m->m + 1) * std::max(m->n_child, 1), 0);
  m->in_p.assign((size_t)(m->m + 1) * std::max(m->n_child, 1), 0);
  for (int lc = 0; lc <= m->m; lc++) {
    int lct = std::min(lc + 1, m->m);
    int nc = m->lc_ctx_ptr[lc + 1] - m->lc_ctx_ptr[lc], nct = m->lc_ctx_ptr[lct + 1] - m->lc_ctx_ptr[lct];
    std::vector<std::vector<std::pair<int, int>>> in(nct);
    for (int c = 0; c < nc; c++) {
      int e = m->ctx_elem[m->lc_ctx_ptr[lc] + c];
      if (e < 0 || e >= m->n_elem) return fail(JH_ERR_INVALID, "context element out of range");
      for (int p = m->elem_child_ptr[e]; p < m->elem_child_ptr[e + 1]; p++) {
        if (m->state_silent[m->child_state[p]]) continue;
        int t = m->child_next[(size_t)lc * m->n_child + p];
        if (t < 0 || t >= nct) return fail(JH_ERR_INVALID, "a child leads to a context that does not exist in the next layer");
        in[t].push_back({c, p});
      }
    }
    int *ip = m->in_ptr.data() + (size_t)lc * (m->Cmax + 1);
    int q = 0;
    for (int t = 0; t < nct; t++) {
      ip[t] = q;
      for (auto &sp : in[t]) {
        m->in_src[(size_t)lc * m->n_child + q] = sp.first;
        m->in_p[(size_t)lc * m->n_child + q] = sp.second;
        q++;
      }
    }
    for (int t = nct; t <= m->Cmax; t++) ip[t] = q;
  }
  // silent in-edges per layer class (sources in the same layer; order = source context, then child)
  m->sin_ptr.assign((size_t)(m->m + 1) * (m->Cmax + 1), 0);
  m->sin_src.assign((size_t)(m->m + 1) * std::max(m->n_child, 1), 0);
  m->sin_p.assign((size_t)(m->m + 1) * std::max(m->n_child, 1), 0);
  for (int lc = 0; lc <= m->m; lc++) {
    int nc = m->lc_ctx_ptr[lc + 1] - m->lc_ctx_ptr[lc];
    std::vector<std::vector<std::pair<int, int>>> in(nc);
    for (int c = 0; c < nc; c++) {
      int e = m->ctx_elem[m->lc_ctx_ptr[lc] + c];
      for (int p = m->elem_child_ptr[e]; p < m->elem_child_ptr[e + 1]; p++) {
        if (!m->state_silent[m->child_state[p]]) continue;
        int t = m->child_next[(size_t)lc * m->n_child + p];
        if (t < 0 || t >= nc) return fail(JH_ERR_INVALID, "a silent child leads to a context that does not exist in its layer");
        if (t <= c) return fail(JH_ERR_INVALID, "silent states are not in topological order");
        in[t].push_back({c, p});
      }
    }
    int *ip = m->sin_ptr.data() + (size_t)lc * (m->Cmax + 1);
    int q = 0;
    for (int t = 0; t < nc; t++) {
      ip[t] = q;
      for (auto &sp : in[t]) {
        m->sin_src[(size_t)lc * m->n_child + q] = sp.first;
        m->sin_p[(size_t)lc * m->n_child + q] = sp.second;
        q++;
      }
    }
    for (int t = nc; t <= m->Cmax; t++) ip[t] = q;
  }
  m->state_order.resize(m->S); m->state_tab.resize(m->S); m->state_mod.resize(m->S);
  m->row_flag.assign(std::max(m->n_cond, 1), 1);
  for (int e = 0; e < m->n_emis; e++)
    if (m->emis_order[e] >= 0 && m->emis_cond_ptr[e + 1] > m->emis_cond_ptr[e]) m->row_flag[m->emis_cond_ptr[e]] = 2;
  for (int s = 0; s < m->S; s++) {
    int e = m->state_emis[s], k = m->emis_order[e];
    m->state_order[s] = k;
    m->state_tab[s] = m->emis_cond_ptr[e] * m->a;
    int64_t mod = m->a;
    for (int i = 0; i < k; i++) mod *= m->a;
    if (k >= 0) {
      if (mod * m->a > (int64_t)1 << 30) return fail(JH_ERR_UNSUPPORTED, "emission order too large for the context hash");
      if (m->emis_cond_ptr[e + 1] - m->emis_cond_ptr[e] != mod / m->a) return fail(JH_ERR_INVALID, "emission table does not have alphabet^order conditions");
      m->Kmax = std::max(m->Kmax, k);
    }
    m->state_mod[s] = k >= 0 ? (int)mod : 1;
  }
  // device ---------------------------------------------------------------------------------------
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, m->device) == cudaSuccess) m->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(JH_ERR_CUDA, "cudaStreamCreate failed");
  Scope scope(m->device, m->st());
  cudaEventCreate(&m->ev0);
  cudaEventCreate(&m->ev1);
  cudaEventCreate(&m->evk0);
  cudaEventCreate(&m->evk1);
  cudaStream_t s = m->st();
  std::vector<unsigned char> fin(m->state_final.begin(), m->state_final.end());
  std::vector<double> tp(d->trans_param, d->trans_param + m->n_child), tl(d->trans_lognorm, d->trans_lognorm + m->n_elem),
      th(d->trans_hyper, d->trans_hyper + m->n_child), ep(d->emis_param, d->emis_param + (size_t)m->n_cond * m->a),
      el(d->emis_lognorm, d->emis_lognorm + m->n_cond), eh(d->emis_hyper, d->emis_hyper + (size_t)m->n_cond * m->a);
#define UP(buf, vec) if ((r = m->buf.upload(vec, s)) != 0) { model_free(m); return r; }
  UP(d_lc_ctx_ptr, m->lc_ctx_ptr) UP(d_ctx_elem, m->ctx_elem) UP(d_ctx_last, m->ctx_last) UP(d_elem_child_ptr, m->elem_child_ptr)
  UP(d_child_state, m->child_state) UP(d_child_next, m->child_next) UP(d_child_stat, m->child_stat) UP(d_in_ptr, m->in_ptr)
  UP(d_in_src, m->in_src) UP(d_in_p, m->in_p) UP(d_state_order, m->state_order) UP(d_state_tab, m->state_tab)
  UP(d_state_mod, m->state_mod) UP(d_child_elem, m->child_elem) UP(d_trans_index, m->trans_index) UP(d_row_flag, m->row_flag)
  UP(d_is_final, fin) UP(d_silent, m->state_silent) UP(d_sin_ptr, m->sin_ptr) UP(d_sin_src, m->sin_src) UP(d_sin_p, m->sin_p) UP(d_trans_param, tp) UP(d_trans_lognorm, tl) UP(d_trans_hyper, th) UP(d_emis_param, ep)
  UP(d_emis_lognorm, el) UP(d_emis_hyper, eh)
#undef UP
  if ((r = m->d_tscore.alloc(m->n_child)) || (r = m->d_escore.alloc((size_t)m->n_cond * m->a)) ||
      (r = m->d_stats.alloc(m->n_stats() + 8)) || (r = m->d_counter.alloc(8)) || (r = m->d_objective.alloc(4096)) || (r = m->d_iter.alloc(4))) {
    model_free(m);
    return r;
  }
  if ((r = jhf::build(m->fast, m->S, m->a, m->m, m->has_silent, m->Kmax, m->lc_ctx_ptr, m->ctx_elem, m->ctx_last, m->elem_child_ptr,
                      m->child_state, m->child_stat, m->state_order, m->state_tab, m->state_mod, m->state_final, m->n_child, m->n_cond, s)) != 0) {
    model_free(m);
    return r;
  }
  if ((r = jhs::build(m->sparse, m->S, m->a, m->m, m->has_silent, m->Kmax, m->lc_ctx_ptr, m->ctx_elem, m->ctx_last, m->elem_child_ptr, m->child_state,
                      m->child_stat, m->state_order, m->state_tab, m->state_mod, m->state_final, s)) != 0) {
    model_free(m);
    return r;
  }
  if ((r = prep_scores(m)) != 0) { model_free(m); return r; }
  if (cudaStreamSynchronize(s) != cudaSuccess) return fail(JH_ERR_CUDA, "model upload failed");
  *out = m;
  return JH_OK;
}

extern "C" void jh_model_destroy(jh_model *m) { model_free(m); }

extern "C" int jh_model_sizes(const jh_model *m, int64_t sizes[4]) {
  if (!m || !sizes) { jh_set_error("null argument"); return JH_ERR_INVALID; }
  sizes[0] = m->n_child; sizes[1] = m->n_elem; sizes[2] = (int64_t)m->n_cond * m->a; sizes[3] = m->n_cond;
  return JH_OK;
}

extern "C" int jh_model_set_params(jh_model *m, const double *tp, const double *tl, const double *ep, const double *el) {
  JH_ENTER(m);
  cudaStream_t s = m->st();
  if (tp) JH_CUDA(cudaMemcpyAsync(m->d_trans_param.p, tp, sizeof(double) * m->n_child, cudaMemcpyHostToDevice, s));
  if (tl) JH_CUDA(cudaMemcpyAsync(m->d_trans_lognorm.p, tl, sizeof(double) * m->n_elem, cudaMemcpyHostToDevice, s));
  if (ep) JH_CUDA(cudaMemcpyAsync(m->d_emis_param.p, ep, sizeof(double) * m->n_cond * m->a, cudaMemcpyHostToDevice, s));
  if (el) JH_CUDA(cudaMemcpyAsync(m->d_emis_lognorm.p, el, sizeof(double) * m->n_cond, cudaMemcpyHostToDevice, s));
  int do_t = tp && !tl, do_e = ep && !el;
  if (do_t || do_e) {
    int n = std::max(m->n_elem, m->n_cond);
    jha::k_precompute_lognorm<<<(n + 127) / 128, 128, 0, s>>>(m->params(), do_t, do_e);
    JH_LAUNCHED();
  }
  int r = prep_scores(m);
  if (r) return r;
  JH_CUDA(cudaStreamSynchronize(s));  // the host buffers are borrowed for the call only
  return JH_OK;
}

extern "C" int jh_model_get_params(jh_model *m, double *tp, double *tl, double *ep, double *el) {
  JH_ENTER(m);
  cudaStream_t s = m->st();
  if (tp) JH_CUDA(cudaMemcpyAsync(tp, m->d_trans_param.p, sizeof(double) * m->n_child, cudaMemcpyDeviceToHost, s));
  if (tl) JH_CUDA(cudaMemcpyAsync(tl, m->d_trans_lognorm.p, sizeof(double) * m->n_elem, cudaMemcpyDeviceToHost, s));
  if (ep) JH_CUDA(cudaMemcpyAsync(ep, m->d_emis_param.p, sizeof(double) * m->n_cond * m->a, cudaMemcpyDeviceToHost, s));
  if (el) JH_CUDA(cudaMemcpyAsync(el, m->d_emis_lognorm.p, sizeof(double) * m->n_cond, cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  return JH_OK;
}

extern "C" int jh_model_log_prior(jh_model *m, double *out) {
  JH_ENTER(m);
  if (!out) { jh_set_error("null argument"); return JH_ERR_INVALID; }
  jha::k_log_prior<<<1, jha::LOG_PRIOR_THREADS, 0, m->st()>>>(m->params(), nullptr, m->d_objective.p + 4095);
  JH_LAUNCHED();
  JH_CUDA(cudaMemcpyAsync(out, m->d_objective.p + 4095, sizeof(double), cudaMemcpyDeviceToHost, m->st()));
  JH_CUDA(cudaStreamSynchronize(m->st()));
  return JH_OK;
}

// ------------------------------------------------------------------------------------------------
static void dataset_free(jh_dataset *d) {
  if (!d) return;
  cudaSetDevice(d->device);
  if (d->ev_last) cudaEventSynchronize(d->ev_last);  // the last call that read the data set, on whatever stream it ran
  if (d->stream) cudaStreamSynchronize(d->stream);
  Scope scope(d->device, cudaStreamPerThread);
  if (d->ev_last) cudaEventDestroy(d->ev_last);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
}
struct DatasetGuard {  // error paths of the constructors release everything
  jh_dataset *d;
  ~DatasetGuard() { dataset_free(d); }
  jh_dataset *release() { jh_dataset *x = d; d = nullptr; return x; }
};

// host-side layout of a data set: length buckets (decreasing length), octets
static int dataset_layout(jh_dataset *d, const int64_t *offsets, int64_t n_seq) {
  if (offsets[0] != 0) { jh_set_error("offsets[0] must be 0"); return JH_ERR_INVALID; }
  d->n_seq = n_seq;
  d->offsets.assign(offsets, offsets + n_seq + 1);
  d->n_res = offsets[n_seq];
  d->min_len = n_seq ? INT64_MAX : 0;
  bool sorted = true;
  for (int64_t i = 0; i < n_seq; i++) {
    int64_t L = offsets[i + 1] - offsets[i];
    if (L <= 0) {  // the reference fails on an empty sequence too (finalState[-1], HigherOrderHMM.java:498)
      jh_set_error("sequence %lld is empty", (long long)i);
      return JH_ERR_INVALID;
    }
    if (i > 0 && L > offsets[i] - offsets[i - 1]) sorted = false;
    d->max_len = std::max(d->max_len, L);
    d->min_len = std::min(d->min_len, L);
  }
  d->order.resize(n_seq);
  std::iota(d->order.begin(), d->order.end(), 0);
  if (!sorted)  // length bucketing: decreasing length, stable
    std::stable_sort(d->order.begin(), d->order.end(), [&](int32_t x, int32_t y) {
      return offsets[x + 1] - offsets[x] > offsets[y + 1] - offsets[y];
    });
  // buckets: consecutive runs whose length is within a factor 2 of the run's first (longest) sequence
  d->bucket_ptr.push_back(0);
  for (int64_t i = 0; i < n_seq;) {
    int64_t L0 = offsets[d->order[i] + 1] - offsets[d->order[i]];
    int64_t j = i;
    while (j < n_seq && 2 * (offsets[d->order[j] + 1] - offsets[d->order[j]]) > L0) j++;
    d->bucket_len.push_back(L0);
    d->bucket_ptr.push_back(j);
    i = j;
  }
  d->bucket_oct_ptr.push_back(0);
  int64_t rec = 0;
  for (size_t b = 0; b + 1 < d->bucket_ptr.size(); b++) {
    for (int64_t i = d->bucket_ptr[b]; i < d->bucket_ptr[b + 1]; i += 8) {
      for (int64_t k = i; k < i + 8; k++) d->oct_seq.push_back(k < d->bucket_ptr[b + 1] ? d->order[k] : -1);
      d->oct_hbase.push_back(rec);
      rec += offsets[d->order[i] + 1] - offsets[d->order[i]];  // the first sequence of an octet is its longest
    }
    d->bucket_oct_ptr.push_back((int64_t)d->oct_hbase.size());
  }
  d->oct_hbase.push_back(rec);  // total number of records
  return JH_OK;
}

// codes: uint8 per residue (packed2 == 0), or 2 bits per residue, four residues per byte, lowest bits first, every
// sequence starting on a byte boundary at byte pack_ptr[n] (packed2 == 1; SparseSequence-style storage of DNA).
static int dataset_create_impl(const uint8_t *codes, const int64_t *pack_ptr, int packed2, const int64_t *offsets, int64_t n_seq,
                               const double *weights, int32_t alphabet, jh_dataset **out) {
  if (!offsets || !out || n_seq < 0 || (n_seq > 0 && !codes)) { jh_set_error("null argument"); return JH_ERR_INVALID; }
  if (n_seq > 0x7fffffff) { jh_set_error("more than 2^31-1 sequences"); return JH_ERR_UNSUPPORTED; }
  if (packed2 && (alphabet > 4 || !pack_ptr)) { jh_set_error("2-bit packing needs an alphabet of at most 4 symbols and byte offsets"); return JH_ERR_INVALID; }
  const int device = g_default_device;
  int r = ensure_device(device);
  if (r) return r;
  DatasetGuard G{new jh_dataset()};
  jh_dataset *d = G.d;
  d->device = device;
  d->alphabet = alphabet;
  Options opt;
  {
    std::lock_guard<std::mutex> g(g_mu);
    opt = g_defaults;
  }
  JH_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  JH_CUDA(cudaEventCreateWithFlags(&d->ev_last, cudaEventDisableTiming));
  Scope scope(device, d->stream);
  cudaStream_t s = d->stream;
  if ((r = dataset_layout(d, offsets, n_seq))) return r;
  if ((r = d->d_codes.alloc((size_t)d->n_res + 64)) || (r = d->d_offsets.alloc(n_seq + 1)) || (r = d->d_order.alloc(n_seq)) ||
      (r = d->d_weights.alloc(n_seq)))
    return r;
  JH_CUDA(cudaMemsetAsync(d->d_codes.p + d->n_res, 0, 64, s));
  JH_CUDA(cudaMemcpyAsync(d->d_offsets.p, offsets, sizeof(int64_t) * (n_seq + 1), cudaMemcpyHostToDevice, s));
  if (n_seq) JH_CUDA(cudaMemcpyAsync(d->d_order.p, d->order.data(), sizeof(int32_t) * n_seq, cudaMemcpyHostToDevice, s));
  if (weights && n_seq) {
    JH_CUDA(cudaMemcpyAsync(d->d_weights.p, weights, sizeof(double) * n_seq, cudaMemcpyHostToDevice, s));
    d->has_weights = true;
    d->weighted_total = 0;
    for (int64_t i = 0; i < n_seq; i++) d->weighted_total += fabs(weights[i]) * (double)(offsets[i + 1] - offsets[i]);
  }
  DevBuf<unsigned int> d_mx;
  if ((r = d_mx.alloc(1))) return r;
  JH_CUDA(cudaMemsetAsync(d_mx.p, 0, sizeof(unsigned int), s));
  if (d->n_res && !packed2) {
    if ((r = stage_h2d(d->d_codes.p, codes, (size_t)d->n_res, s, device, opt.stage != 0))) return r;
    // alphabet check on the device (WrongAlphabetException, AbstractHMM.java:991-993)
    jha::k_max_code<<<592, 256, 0, s>>>(d->d_codes.p, d->n_res, d_mx.p);
    JH_LAUNCHED();
  } else if (d->n_res) {  // 2 bits per residue over PCIe (a quarter of the bytes), widened on the device
    DevBuf<uint8_t> d_pk;
    DevBuf<int64_t> d_pp;
    const int64_t pk_bytes = pack_ptr[n_seq];
    if ((r = d_pk.alloc((size_t)pk_bytes + 16)) || (r = d_pp.alloc(n_seq + 1))) return r;
    JH_CUDA(cudaMemcpyAsync(d_pp.p, pack_ptr, sizeof(int64_t) * (n_seq + 1), cudaMemcpyHostToDevice, s));
    if ((r = stage_h2d(d_pk.p, codes, (size_t)pk_bytes, s, device, opt.stage != 0))) return r;
    jha::k_unpack2<<<(int)std::min<int64_t>(n_seq, 148 * 32), 256, 0, s>>>(d_pk.p, d_pp.p, d->d_offsets.p, n_seq, d->d_codes.p, alphabet, d_mx.p);
    JH_LAUNCHED();
    JH_CUDA(cudaGetLastError());
    JH_CUDA(cudaStreamSynchronize(s));  // d_pk / d_pp are locals
  }
  unsigned int h_mx = 0;
  JH_CUDA(cudaMemcpyAsync(&h_mx, d_mx.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    jh_set_error("CUDA error %s while uploading the data set", cudaGetErrorString(e));
    return JH_ERR_CUDA;
  }
  if ((int)h_mx >= alphabet) {
    jh_set_error("The AlphabetContainer of the data set and the model do not match (code %u >= alphabet size %d).", h_mx, alphabet);
    return JH_ERR_ALPHABET;
  }
  *out = G.release();
  return JH_OK;
}

extern "C" int jh_dataset_create(const uint8_t *codes, const int64_t *offsets, int64_t n_seq, const double *weights,
                                 int32_t alphabet, jh_dataset **out) {
  return dataset_create_impl(codes, nullptr, 0, offsets, n_seq, weights, alphabet, out);
}

extern "C" int jh_dataset_create_packed2(const uint8_t *packed, const int64_t *byte_ptr, const int64_t *offsets, int64_t n_seq,
                                         const double *weights, int32_t alphabet, jh_dataset **out) {
  return dataset_create_impl(packed, byte_ptr, 1, offsets, n_seq, weights, alphabet, out);
}

// ---- streamed construction: the residues arrive shard by shard (memory-mapped files), never as one host array ----------
struct jh_dataset_builder {
  jh_dataset *d = nullptr;
  int64_t n_seq_total = 0, n_res_total = 0, n_seq = 0, n_res = 0;
  std::vector<int64_t> offsets;
  DevBuf<unsigned int> d_mx;
  Options opt;
};

extern "C" int jh_dataset_builder_create(int64_t n_seq, int64_t n_residues, int32_t alphabet, jh_dataset_builder **out) {
  if (!out || n_seq < 0 || n_residues < 0) { jh_set_error("invalid builder arguments"); return JH_ERR_INVALID; }
  if (n_seq > 0x7fffffff) { jh_set_error("more than 2^31-1 sequences"); return JH_ERR_UNSUPPORTED; }
  const int device = g_default_device;
  int r = ensure_device(device);
  if (r) return r;
  DatasetGuard G{new jh_dataset()};
  jh_dataset *d = G.d;
  d->device = device;
  d->alphabet = alphabet;
  JH_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  JH_CUDA(cudaEventCreateWithFlags(&d->ev_last, cudaEventDisableTiming));
  Scope scope(device, d->stream);
  std::unique_ptr<jh_dataset_builder> b(new jh_dataset_builder());
  {
    std::lock_guard<std::mutex> g(g_mu);
    b->opt = g_defaults;
  }
  b->n_seq_total = n_seq; b->n_res_total = n_residues;
  b->offsets.reserve(n_seq + 1);
  b->offsets.push_back(0);
  if ((r = d->d_codes.alloc((size_t)n_residues + 64)) || (r = b->d_mx.alloc(1))) return r;
  JH_CUDA(cudaMemsetAsync(d->d_codes.p + n_residues, 0, 64, d->stream));
  JH_CUDA(cudaMemsetAsync(b->d_mx.p, 0, sizeof(unsigned int), d->stream));
  b->d = G.release();
  *out = b.release();
  return JH_OK;
}

extern "C" void jh_dataset_builder_destroy(jh_dataset_builder *b) {
  if (!b) return;
  if (b->d) {
    Scope scope(b->d->device, b->d->stream);
    b->d_mx.release();
  }
  dataset_free(b->d);
  b->d = nullptr;
  delete b;
}

extern "C" int jh_dataset_builder_append(jh_dataset_builder *b, const uint8_t *codes, const int64_t *lengths, int64_t n_seq, int32_t packed2) {
  if (!b || !b->d || n_seq < 0 || (n_seq > 0 && (!codes || !lengths))) { jh_set_error("invalid append arguments"); return JH_ERR_INVALID; }
  jh_dataset *d = b->d;
  if (packed2 && d->alphabet > 4) { jh_set_error("2-bit packing needs an alphabet of at most 4 symbols"); return JH_ERR_INVALID; }
  Scope scope(d->device, d->stream);
  cudaStream_t s = d->stream;
  int64_t res = 0, bytes = 0;
  std::vector<int64_t> bp(n_seq + 1, 0), off(n_seq + 1, 0);
  for (int64_t i = 0; i < n_seq; i++) {
    if (lengths[i] <= 0) { jh_set_error("sequence %lld is empty", (long long)(b->n_seq + i)); return JH_ERR_INVALID; }
    off[i] = b->n_res + res;
    bp[i] = bytes;
    res += lengths[i];
    bytes += packed2 ? (lengths[i] + 3) / 4 : lengths[i];
  }
  off[n_seq] = b->n_res + res;
  bp[n_seq] = bytes;
  if (b->n_seq + n_seq > b->n_seq_total || b->n_res + res > b->n_res_total) { jh_set_error("more sequences or residues appended than announced"); return JH_ERR_INVALID; }
  int r;
  if (!packed2) {
    if ((r = stage_h2d(d->d_codes.p + b->n_res, codes, (size_t)res, s, d->device, b->opt.stage != 0))) return r;
    if (res) {
      jha::k_max_code<<<592, 256, 0, s>>>(d->d_codes.p + b->n_res, res, b->d_mx.p);
      JH_LAUNCHED();
    }
  } else if (n_seq) {
    DevBuf<uint8_t> d_pk;
    DevBuf<int64_t> d_pp, d_off;
    if ((r = d_pk.alloc((size_t)bytes + 16)) || (r = d_pp.alloc(n_seq + 1)) || (r = d_off.alloc(n_seq + 1))) return r;
    JH_CUDA(cudaMemcpyAsync(d_pp.p, bp.data(), sizeof(int64_t) * (n_seq + 1), cudaMemcpyHostToDevice, s));
    JH_CUDA(cudaMemcpyAsync(d_off.p, off.data(), sizeof(int64_t) * (n_seq + 1), cudaMemcpyHostToDevice, s));
    if ((r = stage_h2d(d_pk.p, codes, (size_t)bytes, s, d->device, b->opt.stage != 0))) return r;
    jha::k_unpack2<<<(int)std::min<int64_t>(n_seq, 148 * 32), 256, 0, s>>>(d_pk.p, d_pp.p, d_off.p, n_seq, d->d_codes.p, d->alphabet, b->d_mx.p);
    JH_LAUNCHED();
    JH_CUDA(cudaGetLastError());
    JH_CUDA(cudaStreamSynchronize(s));  // the packed staging buffers are locals
  }
  JH_CUDA(cudaStreamSynchronize(s));
  for (int64_t i = 0; i < n_seq; i++) b->offsets.push_back(off[i + 1]);
  b->n_seq += n_seq;
  b->n_res += res;
  return JH_OK;
}

extern "C" int jh_dataset_builder_finish(jh_dataset_builder *b, const double *weights, jh_dataset **out) {
  if (!b || !b->d || !out) { jh_set_error("invalid finish arguments"); return JH_ERR_INVALID; }
  if (b->n_seq != b->n_seq_total || b->n_res != b->n_res_total) {
    jh_set_error("the builder holds %lld of %lld sequences and %lld of %lld residues", (long long)b->n_seq, (long long)b->n_seq_total, (long long)b->n_res,
                 (long long)b->n_res_total);
    return JH_ERR_INVALID;
  }
  jh_dataset *d = b->d;
  Scope scope(d->device, d->stream);
  cudaStream_t s = d->stream;
  const int64_t n_seq = b->n_seq;
  int r = dataset_layout(d, b->offsets.data(), n_seq);
  if (r) return r;
  if ((r = d->d_offsets.alloc(n_seq + 1)) || (r = d->d_order.alloc(n_seq)) || (r = d->d_weights.alloc(n_seq))) return r;
  JH_CUDA(cudaMemcpyAsync(d->d_offsets.p, b->offsets.data(), sizeof(int64_t) * (n_seq + 1), cudaMemcpyHostToDevice, s));
  if (n_seq) JH_CUDA(cudaMemcpyAsync(d->d_order.p, d->order.data(), sizeof(int32_t) * n_seq, cudaMemcpyHostToDevice, s));
  if (weights && n_seq) {
    JH_CUDA(cudaMemcpyAsync(d->d_weights.p, weights, sizeof(double) * n_seq, cudaMemcpyHostToDevice, s));
    d->has_weights = true;
    d->weighted_total = 0;
    for (int64_t i = 0; i < n_seq; i++) d->weighted_total += fabs(weights[i]) * (double)(d->offsets[i + 1] - d->offsets[i]);
  }
  unsigned int h_mx = 0;
  JH_CUDA(cudaMemcpyAsync(&h_mx, b->d_mx.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  if ((int)h_mx >= d->alphabet) {
    jh_set_error("The AlphabetContainer of the data set and the model do not match (code %u >= alphabet size %d).", h_mx, d->alphabet);
    return JH_ERR_ALPHABET;
  }
  b->d_mx.release();
  *out = d;
  b->d = nullptr;
  delete b;
  return JH_OK;
}

extern "C" int jh_dataset_set_weights(jh_dataset *d, const double *weights) {
  if (!d) { jh_set_error("null data set"); return JH_ERR_INVALID; }
  if (!weights) { d->has_weights = false; return JH_OK; }
  d->weighted_total = 0;
  for (int64_t i = 0; i < d->n_seq; i++) d->weighted_total += fabs(weights[i]) * (double)(d->offsets[i + 1] - d->offsets[i]);
  JH_CUDA(cudaSetDevice(d->device));
  if (d->ev_last) JH_CUDA(cudaEventSynchronize(d->ev_last));  // a running E-step may still read the old weights
  if (d->n_seq) {
    JH_CUDA(cudaMemcpyAsync(d->d_weights.p, weights, sizeof(double) * d->n_seq, cudaMemcpyHostToDevice, d->stream));
    JH_CUDA(cudaStreamSynchronize(d->stream));
  }
  d->has_weights = true;
  return JH_OK;
}

extern "C" void jh_dataset_destroy(jh_dataset *d) { dataset_free(d); }

static bool general_model(const jh_model *m);

extern "C" int jh_dataset_set_groups(jh_dataset *d, const int32_t *groups) {
  if (!d) { jh_set_error("null data set"); return JH_ERR_INVALID; }
  if (!groups) { d->has_groups = false; return JH_OK; }
  JH_CUDA(cudaSetDevice(d->device));
  Scope scope(d->device, d->stream);
  if (d->ev_last) JH_CUDA(cudaEventSynchronize(d->ev_last));
  int r = d->d_groups.reserve((size_t)d->n_res + 1);
  if (r) return r;
  if ((r = stage_h2d(d->d_groups.p, groups, sizeof(int32_t) * (size_t)d->n_res, d->stream, d->device, true))) return r;
  JH_CUDA(cudaStreamSynchronize(d->stream));
  d->has_groups = true;
  return JH_OK;
}

// AbstractHMM.fillPreComputedContext (AbstractHMM.java:173-228): a context of layer class m is allowed for group g if some
// context of that class has a child whose state belongs to g and which leads to it.
extern "C" int jh_model_set_states_groups(jh_model *m, int32_t n_groups, const int32_t *group_ptr, const int32_t *group_states) {
  JH_ENTER(m);
  if (n_groups < 0 || (n_groups > 0 && (!group_ptr || !group_states))) { jh_set_error("invalid states groups"); return JH_ERR_INVALID; }
  if (n_groups > 0 && general_model(m)) { jh_set_error("allowed states groups with silent states or transition order 0 are not on the native path"); return JH_ERR_UNSUPPORTED; }
  m->n_groups = 0;
  if (n_groups == 0) return JH_OK;
  std::vector<unsigned char> allowed((size_t)n_groups * m->Cmax, 0), in_group(m->S);
  const int lc = m->m, c0 = m->lc_ctx_ptr[lc], nc = m->lc_ctx_ptr[lc + 1] - c0;
  for (int g = 0; g < n_groups; g++) {
    std::fill(in_group.begin(), in_group.end(), 0);
    for (int k = group_ptr[g]; k < group_ptr[g + 1]; k++) {
      if (group_states[k] < 0 || group_states[k] >= m->S) { jh_set_error("states group %d names an unknown state", g); return JH_ERR_INVALID; }
      in_group[group_states[k]] = 1;
    }
    for (int c = 0; c < nc; c++) {
      const int e = m->ctx_elem[c0 + c];
      for (int p = m->elem_child_ptr[e]; p < m->elem_child_ptr[e + 1]; p++)
        if (in_group[m->child_state[p]]) allowed[(size_t)g * m->Cmax + m->child_next[(size_t)lc * m->n_child + p]] = 1;
    }
  }
  int r = m->d_allowed.upload(allowed, m->st());
  if (r) return r;
  JH_CUDA(cudaStreamSynchronize(m->st()));
  m->n_groups = n_groups;
  return JH_OK;
}

extern "C" int jh_dataset_info(const jh_dataset *d, int64_t *n_seq, int64_t *n_res, int64_t *max_len) {
  if (!d) { jh_set_error("null data set"); return JH_ERR_INVALID; }
  if (n_seq) *n_seq = d->n_seq;
  if (n_res) *n_res = d->n_res;
  if (max_len) *max_len = d->max_len;
  return JH_OK;
}

// ------------------------------------------------------------------------------------------------
static int check_pair(jh_model *m, jh_dataset *d) {
  if (!m || !d) { jh_set_error("null handle"); return JH_ERR_INVALID; }
  if (d->alphabet != m->a) { jh_set_error("The AlphabetContainer of the data set and the model do not match."); return JH_ERR_ALPHABET; }
  if (d->device != m->device) { jh_set_error("model (device %d) and data set (device %d) live on different devices", m->device, d->device); return JH_ERR_INVALID; }
  return JH_OK;
}
struct UseMark {  // the data set remembers the last work that reads it (jh_dataset_destroy / set_weights wait for it)
  jh_dataset *d;
  cudaStream_t s;
  ~UseMark() {
    if (d && d->ev_last) cudaEventRecord(d->ev_last, s);
  }
};
#define JH_ENTER2(m, d)                       \
  JH_ENTER(m);                                \
  {                                           \
    int r__ = check_pair(m, d);               \
    if (r__) return r__;                      \
  }                                           \
  UseMark mark__{d, (m)->st()}
// Models that run on the general context-graph kernels (kernels_general.cuh, one sequence per thread): silent states, and
// transition order 0 (a single context for every layer: every context is final, HigherOrderHMM.java:498, 1046-1058).
static bool general_model(const jh_model *m) { return m->has_silent || m->m == 0; }
// allowed states groups in force: the model declares groups and the data set carries the per-position constraints; only
// the reference-order kernels (kernels_exact.cuh) evaluate them
static bool groups_active(const jh_model *m, const jh_dataset *d) { return m->n_groups > 0 && d->has_groups; }

static int group_size(const jh_model *m) {
  int g = 2;
  while (g < 32 && g < m->Cmax) g *= 2;
  return g;
}

struct LaunchPlan {
  int G, threads, grid, groups_per_cta;
  size_t smem;
  int64_t slots;
};

// grid = resident CTAs (multiple of the SM count) bounded by the work and by the scratch budget
template <typename K>
static int plan_launch(const jh_model *m, K kernel, int G, size_t extra_smem_doubles, int64_t work, size_t scratch_bytes_per_slot,
                       LaunchPlan *pl, int64_t budget_bytes = 0) {
  pl->G = G; pl->threads = 256; pl->groups_per_cta = 256 / G;
  pl->smem = ((size_t)pl->groups_per_cta * (2 * m->Cmax + m->S) + extra_smem_doubles) * sizeof(double);
  if (pl->smem > 200 * 1024) { jh_set_error("model too large for the shared-memory layout (%zu bytes)", pl->smem); return JH_ERR_UNSUPPORTED; }
  if (pl->smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  int per_sm = 0;
  JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, pl->threads, pl->smem));
  if (per_sm < 1) { jh_set_error("kernel does not fit on an SM"); return JH_ERR_UNSUPPORTED; }
  if (m->opt.blocks_per_sm > 0) per_sm = std::min<int>(per_sm, (int)m->opt.blocks_per_sm);
  int64_t grid = (int64_t)m->sm_count * per_sm;
  int64_t need = (work + pl->groups_per_cta - 1) / pl->groups_per_cta;
  grid = std::max<int64_t>(1, std::min(grid, need));
  if (scratch_bytes_per_slot) {
    int64_t budget = budget_bytes > 0 ? budget_bytes : m->opt.scratch_mb * 1024 * 1024;
    int64_t max_slots = budget / (int64_t)scratch_bytes_per_slot;
    if (max_slots < pl->groups_per_cta) { jh_set_error("sequence too long for the DP scratch budget (%lld MB)", (long long)m->opt.scratch_mb); return JH_ERR_UNSUPPORTED; }
    grid = std::min<int64_t>(grid, max_slots / pl->groups_per_cta);
  }
  pl->grid = (int)grid;
  pl->slots = grid * pl->groups_per_cta;
  return JH_OK;
}

static void timing_begin(jh_model *m) { cudaEventRecord(m->ev0, m->st()); }
static void timing_end(jh_model *m) { cudaEventRecord(m->ev1, m->st()); m->timed = true; }

extern "C" int jh_last_kernel_ms(jh_model *m, double *ms) {
  if (!m || !ms || !m->timed) { jh_set_error("no timed call yet"); return JH_ERR_INVALID; }
  JH_CUDA(cudaEventSynchronize(m->ev1));
  float f = 0;
  JH_CUDA(cudaEventElapsedTime(&f, m->ev0, m->ev1));
  *ms = f;
  return JH_OK;
}

extern "C" int jh_last_main_kernel_ms(jh_model *m, double *ms) {
  if (!m || !ms || !m->ktimed) { jh_set_error("no timed E-step yet"); return JH_ERR_INVALID; }
  JH_CUDA(cudaEventSynchronize(m->evk1));
  float f = 0;
  JH_CUDA(cudaEventElapsedTime(&f, m->evk0, m->evk1));
  *ms = f;
  return JH_OK;
}

#define DISPATCH_G(G, ...)                                  \
  switch (G) {                                              \
    case 2: { constexpr int GG = 2; __VA_ARGS__; } break;   \
    case 4: { constexpr int GG = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int GG = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int GG = 16; __VA_ARGS__; } break; \
    default: { constexpr int GG = 32; __VA_ARGS__; } break; \
  }

// ---- log-likelihood --------------------------------------------------------------------------
static int run_loglik_exact(jh_model *m, jh_dataset *d, double *d_out, const int32_t *d_list, int64_t n_list) {
  int G = group_size(m), r = JH_OK;
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  DISPATCH_G(G, {
    LaunchPlan pl;
    r = plan_launch(m, jhx::k_exact_loglik<GG>, GG, 0, d_list ? n_list : d->n_seq, 0, &pl);
    if (!r) {
      jhx::k_exact_loglik<GG><<<pl.grid, pl.threads, pl.smem, s>>>(m->dev(), d->dev(), m->d_counter.p, d_out, d_list, n_list);
      JH_LAUNCHED();
    }
  });
  if (r) return r;
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

// ---- thread-per-sequence kernels for the general context graph (kernels_general.cuh) ----------------------------
// threads of the grid: bounded by the work and by the DP scratch each thread needs
static int plan_general(jh_model *m, int64_t work, size_t bytes_per_thread, int *grid, int64_t *nthreads) {
  const int NT = 128;
  int64_t n = std::min<int64_t>((int64_t)m->sm_count * 1024, (work + NT - 1) / NT * NT);
  int64_t cap = m->opt.scratch_mb * 1024 * 1024 / (int64_t)std::max<size_t>(bytes_per_thread, 1);
  cap = cap / 32 * 32;
  if (cap < 32) { jh_set_error("sequence too long for the DP scratch budget (%lld MB)", (long long)m->opt.scratch_mb); return JH_ERR_UNSUPPORTED; }
  n = std::max<int64_t>(32, std::min(n, cap));
  *grid = (int)((n + NT - 1) / NT);
  *nthreads = (int64_t)*grid * NT;
  if (*nthreads > cap) {  // a partial last block is not possible: shrink to whole blocks or run one small block
    *grid = (int)std::max<int64_t>(1, cap / NT);
    *nthreads = (int64_t)*grid * NT;
    if (*nthreads > cap) { jh_set_error("sequence too long for the DP scratch budget (%lld MB)", (long long)m->opt.scratch_mb); return JH_ERR_UNSUPPORTED; }
  }
  return JH_OK;
}

static int run_general_loglik(jh_model *m, jh_dataset *d, const jhg::WorkList &W, double *d_out) {
  int grid;
  int64_t nthreads;
  int r = plan_general(m, W.n, (size_t)2 * m->Cmax * sizeof(double), &grid, &nthreads);
  if (r) return r;
  if ((r = m->d_rows.reserve((size_t)2 * m->Cmax * nthreads))) return r;
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  jhg::k_general_loglik<<<grid, 128, 0, s>>>(m->dev(), m->gen(), d->dev(), W, m->d_counter.p, m->d_rows.p, d_out);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

static int run_general_viterbi(jh_model *m, jh_dataset *d, int32_t *d_paths, const int64_t *d_path_ptr, int64_t *d_path_len, double *d_scores,
                               double *d_stats, const jhg::WorkList *windows = nullptr) {
  if (m->max_children > 254) { jh_set_error("more than 254 children per context"); return JH_ERR_UNSUPPORTED; }
  int grid;
  int64_t nthreads;
  const int64_t layers = d->max_len + 1;
  const jhg::WorkList W = windows ? *windows : jhg::WorkList{nullptr, nullptr, nullptr, d->n_seq};
  int r = plan_general(m, W.n, (size_t)2 * m->Cmax * sizeof(double) + (size_t)layers * m->Cmax, &grid, &nthreads);
  if (r) return r;
  if ((r = m->d_rows.reserve((size_t)2 * m->Cmax * nthreads)) || (r = m->d_choice.reserve((size_t)layers * m->Cmax * nthreads))) return r;
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  jhg::k_general_viterbi<<<grid, 128, 0, s>>>(m->dev(), m->gen(), d->dev(), W, m->d_counter.p, m->d_rows.p, m->d_choice.p, layers, d_paths, d_path_ptr,
                                             d_path_len, d_scores, d_stats);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

static int run_general_estep(jh_model *m, jh_dataset *d) {
  int grid;
  int64_t nthreads;
  const int64_t layers = d->max_len + 1;
  int r = plan_general(m, d->n_seq, (size_t)(2 + layers) * m->Cmax * sizeof(double), &grid, &nthreads);
  if (r) return r;
  if ((r = m->d_rows.reserve((size_t)2 * m->Cmax * nthreads)) || (r = m->d_fwd.reserve((size_t)layers * m->Cmax * nthreads))) return r;
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  jhg::k_general_estep<<<grid, 128, 0, s>>>(m->dev(), m->gen(), d->dev(), m->d_counter.p, m->d_rows.p, m->d_fwd.p, layers, m->d_stats.p);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

// posterior matrix of one sequence (d_post) or posterior decoding of a work list with the thread-per-sequence kernel
static int run_general_posterior(jh_model *m, jh_dataset *d, const jhg::WorkList &W, int64_t max_len, double *d_post, PathOut d_paths, double *d_ll) {
  int grid;
  int64_t nthreads;
  const int64_t layers = max_len + 1;
  int r = plan_general(m, W.n, ((size_t)2 * layers * m->Cmax + m->S) * sizeof(double), &grid, &nthreads);
  if (r) return r;
  if ((r = m->d_fwd.reserve((size_t)layers * m->Cmax * nthreads)) || (r = m->d_rows.reserve((size_t)(layers * m->Cmax + m->S) * nthreads))) return r;
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  jhg::k_general_posterior<<<grid, 128, 0, s>>>(m->dev(), m->gen(), d->dev(), W, m->d_counter.p, m->d_fwd.p, m->d_rows.p,
                                               m->d_rows.p + (size_t)layers * m->Cmax * nthreads, d_post, d_paths, d_ll);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

// ---- sparse path (kernels_sparse.cuh): 33..256 states, chromosome-scale sequences ----------------------------------
// One warp per sequence/chunk.  Models with more than 32 states always take it; smaller ones when the data set is a few long
// sequences (the octet kernels need 8 sequences per warp to fill the GPU and (L+1) x 8 x G doubles of scratch per warp).
// fast_variant in force for a model: the lane-per-state comparison kernels (1..3) keep the emission table in shared memory,
// so models with a large table (e_global) always take the production mapping
static int variant_of(const jh_model *m) {
  return (m->fast.usable && m->fast.e_global && m->opt.fast_variant != 4) ? 0 : (int)m->opt.fast_variant;
}

// Tiny models (2..4 states) on a data set that is too small to fill the GPU with one sequence per lane (BASELINE configs[0]:
// 10 000 sequences = 313 warps, each a dependent chain of 2 x 1000 positions): the sequences are cut into short chunks and
// take the position-parallel checkpoint pass + the lane-per-chunk pass, which turns the chains into ~4x more, ~4x shorter
// ones (cfg1 E-step 0.81 -> 0.53 ms, the whole 10-iteration train() call 9.1 -> 6.2 ms).
static bool split_short_sequences(const jh_model *m, const jh_dataset *d) {
  return m->opt.split_short && m->opt.scan && m->opt.lane_chunks && m->opt.fast_variant == 0 && m->fast.usable && !m->fast.e_global && m->fast.G <= 4 &&
         d->n_seq < 32768 && d->n_res >= (1 << 20) && d->max_len >= 256 && d->max_len < 16384;
}

static bool prefer_sparse(const jh_model *m, const jh_dataset *d) {
  if (!m->opt.fast || !m->sparse.usable) return false;
  if (m->S > 32 || m->opt.fast_variant == 4) return true;
  if (m->opt.fast_variant != 0) return false;
  return (d->n_seq <= 1024 && d->max_len >= 16384) || split_short_sequences(m, d);
}

#define DISPATCH_SPARSE(P, ...)                                   \
  if ((P).e_global) {                                             \
    constexpr bool EG_ = true;                                    \
    DISPATCH_SPARSE_(P, __VA_ARGS__)                              \
  } else {                                                        \
    constexpr bool EG_ = false;                                   \
    DISPATCH_SPARSE_(P, __VA_ARGS__)                              \
  }
#define DISPATCH_SPARSE_(P, ...)                                                            \
  switch ((P).SPL * 100 + (P).IND) {                                                        \
    case 104: { constexpr int SPL_ = 1, IND_ = 4; __VA_ARGS__; } break;                     \
    case 108: { constexpr int SPL_ = 1, IND_ = 8; __VA_ARGS__; } break;                     \
    case 116: { constexpr int SPL_ = 1, IND_ = 16; __VA_ARGS__; } break;                    \
    case 132: { constexpr int SPL_ = 1, IND_ = 32; __VA_ARGS__; } break;                    \
    case 204: { constexpr int SPL_ = 2, IND_ = 4; __VA_ARGS__; } break;                     \
    case 208: { constexpr int SPL_ = 2, IND_ = 8; __VA_ARGS__; } break;                     \
    case 216: { constexpr int SPL_ = 2, IND_ = 16; __VA_ARGS__; } break;                    \
    case 404: { constexpr int SPL_ = 4, IND_ = 4; __VA_ARGS__; } break;                     \
    case 408: { constexpr int SPL_ = 4, IND_ = 8; __VA_ARGS__; } break;                     \
    default: { constexpr int SPL_ = 8, IND_ = 4; __VA_ARGS__; } break;                      \
  }

static int ensure_chunks(jh_model *m, jh_dataset *d, int64_t CK) {
  if (d->chunk_len == CK && d->d_chunk_ptr.p) return JH_OK;
  std::vector<int64_t> ptr(d->n_seq + 1, 0);
  for (int64_t i = 0; i < d->n_seq; i++) ptr[i + 1] = ptr[i] + (d->offsets[i + 1] - d->offsets[i] + CK - 1) / CK;
  const int64_t n = ptr[d->n_seq];
  if (n > 0x7fffffff) { jh_set_error("too many chunks"); return JH_ERR_UNSUPPORTED; }
  std::vector<int32_t> cs(std::max<int64_t>(n, 1)), ci(std::max<int64_t>(n, 1)), lg;
  std::vector<int64_t> seg;
  int64_t g = 0;
  for (int64_t k = 0; k < d->n_seq; k++) {  // longest sequences first
    const int32_t sid = d->order[k];
    if (ptr[sid + 1] - ptr[sid] > 1) {
      lg.push_back(sid);
      seg.push_back(g);
    }
    for (int64_t c = 0; c < ptr[sid + 1] - ptr[sid]; c++, g++) { cs[g] = sid; ci[g] = (int32_t)c; }
  }
  cudaStream_t s = m->st();
  int r;
  if ((r = d->d_chunk_ptr.upload(ptr, s)) || (r = d->d_chunk_seq.upload(cs, s)) || (r = d->d_chunk_idx.upload(ci, s)) || (r = d->d_long.upload(lg, s))) return r;
  {
    std::vector<int64_t> seg2 = seg;
    seg2.push_back(seg.empty() ? 0 : ptr[lg.back() + 1] - ptr[lg.back()] + seg.back());
    if ((r = d->d_long_seg.upload(seg2, s))) return r;
    JH_CUDA(cudaStreamSynchronize(s));  // seg2 is a local
  }
  JH_CUDA(cudaStreamSynchronize(s));
  d->chunk_len = CK;
  d->n_chunks = n;
  d->n_long = (int64_t)lg.size();
  seg.push_back(seg.empty() ? 0 : ptr[lg.back() + 1] - ptr[lg.back()] + seg.back());  // long sequences come first (decreasing length)
  d->long_seg = seg;
  return JH_OK;
}

// Band template of a model (kernels_band.cuh): the smallest instantiated offset range that contains the model's band.
// Returns false when the model is not banded or no instantiation covers it.
struct BandSel { int spl, lo, hi; };
static bool band_select(const jhs::SparseTables &P, BandSel *sel) {
  if (P.band_hi < P.band_lo || P.e_global || (size_t)P.H * P.SP * sizeof(double) > 160 * 1024) return false;
  static const BandSel cand[] = {{2, 0, 1}, {2, 0, 2}, {2, -1, 1}, {2, -2, 2}, {4, 0, 1}, {4, 0, 3}, {4, -1, 1}, {4, -4, 4},
                                 {8, 0, 3}, {8, -1, 1}, {8, -3, 3}};
  for (const BandSel &c : cand)
    if (c.spl == P.SPL && c.lo <= P.band_lo && P.band_hi <= c.hi) { *sel = c; return true; }
  return false;
}
#define JH_BAND_DISPATCH(sel, ...)                                                                        \
  switch ((sel).spl * 10000 + ((sel).lo + 50) * 100 + (sel).hi) {                                            \
    case 2 * 10000 + 50 * 100 + 1: { constexpr int SPL_ = 2, LO_ = 0, HI_ = 1; __VA_ARGS__; } break;       \
    case 2 * 10000 + 50 * 100 + 2: { constexpr int SPL_ = 2, LO_ = 0, HI_ = 2; __VA_ARGS__; } break;       \
    case 2 * 10000 + 49 * 100 + 1: { constexpr int SPL_ = 2, LO_ = -1, HI_ = 1; __VA_ARGS__; } break;      \
    case 2 * 10000 + 48 * 100 + 2: { constexpr int SPL_ = 2, LO_ = -2, HI_ = 2; __VA_ARGS__; } break;      \
    case 4 * 10000 + 50 * 100 + 1: { constexpr int SPL_ = 4, LO_ = 0, HI_ = 1; __VA_ARGS__; } break;       \
    case 4 * 10000 + 50 * 100 + 3: { constexpr int SPL_ = 4, LO_ = 0, HI_ = 3; __VA_ARGS__; } break;       \
    case 4 * 10000 + 49 * 100 + 1: { constexpr int SPL_ = 4, LO_ = -1, HI_ = 1; __VA_ARGS__; } break;      \
    case 4 * 10000 + 46 * 100 + 4: { constexpr int SPL_ = 4, LO_ = -4, HI_ = 4; __VA_ARGS__; } break;      \
    case 8 * 10000 + 50 * 100 + 3: { constexpr int SPL_ = 8, LO_ = 0, HI_ = 3; __VA_ARGS__; } break;       \
    case 8 * 10000 + 49 * 100 + 1: { constexpr int SPL_ = 8, LO_ = -1, HI_ = 1; __VA_ARGS__; } break;      \
    default: { constexpr int SPL_ = 8, LO_ = -3, HI_ = 3; __VA_ARGS__; } break;                             \
  }

// phase 1: n_dir = 1 scoring only, n_dir = 2 forward and backward warps with checkpoints
static int run_sparse_pass1(jh_model *m, jh_dataset *d, int n_dir, double *d_ll, const jhs::Pass1Extra *extra = nullptr, cudaStream_t on = nullptr,
                            bool reset_fallback = true) {
  jhs::SparseTables &P = m->sparse;
  cudaStream_t s = on ? on : m->st();
  int r = m->d_fallback.reserve(d->n_seq + 2);
  if (r) return r;
  jhs::Pass1Extra X{};
  if (extra) X = *extra;
  if (reset_fallback) JH_CUDA(cudaMemsetAsync(m->d_fallback.p, 0, sizeof(int32_t), s));
  if (!on) JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));  // only the warp-per-sequence kernel uses it
  const int64_t items = (X.list ? X.n_list : d->n_seq) * n_dir;
  if (items == 0) return JH_OK;
  BandSel bsel;
  if (m->opt.band && items <= 8 * (int64_t)m->sm_count && band_select(P, &bsel)) {  // banded model: warp-resident, no shared-memory exchange
    size_t smem_b = (size_t)P.H * P.SP * sizeof(double);
    if (on) {  // overlapped with chunk kernels of other sequences: keep the SM to the latency-bound warp (see below)
      int optin = 0;
      JH_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
      smem_b = std::max<size_t>(smem_b, (size_t)optin);
    }
    const bool ck = n_dir == 2 && !X.meet_vec;  // checkpoints: not when the two directions only meet in the middle (scoring)
    JH_BAND_DISPATCH(bsel, {
      auto kern = jhb::k_band_pass1<SPL_, LO_, HI_>;
      if (smem_b > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
      kern<<<(int)items, 32, smem_b, s>>>(P.dev(), P.IND, d->dev(), n_dir, ck ? (int)(d->chunk_len > 0 ? d->chunk_len : 1) : 0x3fffffff,
                                         ck ? d->d_chunk_ptr.p : nullptr, ck ? m->d_ck_alpha.p : nullptr, ck ? m->d_ck_beta.p : nullptr, d_ll,
                                         m->d_fallback.p, X);
      JH_LAUNCHED();
    });
    if (X.meet_vec) {
      const int64_t n = items / 2;
      jhb::k_band_meet<<<(int)((n + 3) / 4), 128, 0, s>>>(P.SP, d->dev(), X.list, n, X.meet_vec, X.meet_exp, X.meet_ok, d_ll, m->d_fallback.p);
      JH_LAUNCHED();
    }
    JH_CUDA(cudaGetLastError());
    return JH_OK;
  }
  if (X.meet_vec) { jh_set_error("internal: meet-in-the-middle scoring outside the band kernels"); return JH_ERR_INVALID; }
  if (m->opt.pass1_cta && items <= (int64_t)m->sm_count) {  // few long sequences: latency mode, one CTA per (sequence, direction)
    size_t smem_c = ((P.e_global ? 0 : (size_t)P.H * P.SP) + 2 * (size_t)P.SP + 16) * sizeof(double);
    if (on) {  // overlapped with chunk kernels of other sequences: claim the SM's whole shared memory, so that no chunk CTA
               // becomes co-resident and competes for the issue slots of this latency-bound CTA (measured: +25 % per step)
      int optin = 0;
      JH_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
      smem_c = std::max<size_t>(smem_c, (size_t)optin);
    }
#define JH_P1C(IND_)                                                                                                                     \
  if (P.e_global) JH_P1C_(IND_, true) else JH_P1C_(IND_, false)
#define JH_P1C_(IND_, EG_)                                                                                                               \
  {                                                                                                                                      \
    auto kern = jhs::k_sparse_pass1_cta<IND_, EG_>;                                                                                         \
    if (smem_c > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));               \
    kern<<<(int)items, P.SP, smem_c, s>>>(P.dev(), d->dev(), n_dir, (int)(d->chunk_len > 0 ? d->chunk_len : 1),                           \
                                          n_dir == 2 ? d->d_chunk_ptr.p : nullptr, n_dir == 2 ? m->d_ck_alpha.p : nullptr,               \
                                          n_dir == 2 ? m->d_ck_beta.p : nullptr, d_ll, m->d_fallback.p, X);                              \
    JH_LAUNCHED();                                                                                                                       \
  }
    switch (P.IND) {
      case 4: JH_P1C(4) break;
      case 8: JH_P1C(8) break;
      case 16: JH_P1C(16) break;
      default: JH_P1C(32) break;
    }
#undef JH_P1C
#undef JH_P1C_
    JH_CUDA(cudaGetLastError());
    return JH_OK;
  }
  const int NT = items <= 2 * (int64_t)m->sm_count ? 32 : 128;  // few long sequences: one warp per SM
  const size_t smem = ((P.e_global ? 0 : (size_t)P.H * P.SP) + (size_t)(NT / 32) * 2 * P.SP) * sizeof(double);
  DISPATCH_SPARSE(P, {
    auto kern = jhs::k_sparse_pass1<SPL_, IND_, EG_>;
    if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) { jh_set_error("sparse kernel does not fit on an SM"); return JH_ERR_UNSUPPORTED; }
    int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)m->sm_count * per_sm, (items + NT / 32 - 1) / (NT / 32)));
    kern<<<(int)grid, NT, smem, s>>>(P.dev(), d->dev(), n_dir, (int)(d->chunk_len > 0 ? d->chunk_len : 1), n_dir == 2 ? d->d_chunk_ptr.p : nullptr,
                                     n_dir == 2 ? m->d_ck_alpha.p : nullptr, n_dir == 2 ? m->d_ck_beta.p : nullptr, m->d_counter.p, d_ll, m->d_fallback.p, X);
    JH_LAUNCHED();
  });
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

// Checkpoint pass parallel in the position (kernels_scan.cuh) for the long sequences of the data set: per-chunk transfer rows,
// then one warp per (sequence, direction) combines them into checkpoints / P(x).  n_dir = 1: forward only (scoring).
// Positions per chunk of the long-sequence path.  Models of the parallel checkpoint pass (dense, <= 32 states: SP = 32) take
// 4x longer chunks: their combination step walks over the chunks one by one, and their chunk scratch per resident warp is
// a quarter of a 128-state model's at equal length.  (Tiny test chunk lengths are taken literally.)
static int64_t chunk_len_for(const jh_model *m, const jh_dataset *d) {
  if (split_short_sequences(m, d)) {  // ~49 000 chunks (lanes) for the whole data set, at least 64 positions each
    const int64_t ck = (d->n_res / 49152 + 15) / 16 * 16;
    return std::max<int64_t>(64, std::min<int64_t>(ck, m->opt.chunk));
  }
  const bool small_dense = m->opt.scan && m->fast.usable && !m->fast.e_global;
  return small_dense && m->opt.chunk >= 1024 ? m->opt.chunk * 4 : m->opt.chunk;
}

static bool scan_applies(const jh_model *m, const jh_dataset *d) {
  if (!(m->opt.scan && d->n_long > 0 && m->fast.usable && !m->fast.e_global)) return false;
  // 32 states: 32x the arithmetic of the sequential pass and 16 KB of transfer rows per chunk — only while those fit the scratch budget
  const int64_t n_lc = d->long_seg.empty() ? 0 : d->long_seg[d->n_long];
  return m->fast.G <= 16 || n_lc * 2 * 32 * 32 * 8 <= m->opt.scratch_mb * 1024 * 1024 / 2;
}
static int run_scan_checkpoints(jh_model *m, jh_dataset *d, int n_dir, double *d_ll, const jhs::Pass1Extra &X) {
  const jhf::FastTables &Ft = m->fast;
  jhs::SparseTables &P = m->sparse;
  cudaStream_t s = m->st();
  const int G = Ft.G;
  const int64_t CK = d->chunk_len;
  const int64_t n_lc = d->long_seg[d->n_long];  // the chunks of the long sequences are the head of the chunk list
  int r;
  if ((r = m->d_scanR.reserve((size_t)n_lc * 2 * G * G)) || (r = m->d_scanE.reserve((size_t)n_lc * 2 * G))) return r;
  const size_t smem_s = ((size_t)G * G + (size_t)Ft.H * G) * sizeof(double);
  const int64_t items = n_lc * n_dir * G;
  const int grid_a = (int)std::max<int64_t>(1, std::min<int64_t>((items + 127) / 128, (int64_t)m->sm_count * 16));
  const int grid_b = (int)((n_dir * d->n_long + 3) / 4);
#define JH_SCAN(GG)                                                                                                                       \
  {                                                                                                                                       \
    if (smem_s > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(jhp::k_scan_chunks<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s)); \
    jhp::k_scan_chunks<GG><<<grid_a, 128, smem_s, s>>>(Ft.dev(), d->dev(), (int)CK, n_lc, n_dir, d->d_chunk_seq.p, d->d_chunk_idx.p,         \
                                                       m->d_scanR.p, m->d_scanE.p);                                                      \
    JH_LAUNCHED();                                                                                                                        \
    jhp::k_scan_combine<GG><<<grid_b, 128, 0, s>>>(Ft.dev(), d->dev(), d->n_long, n_dir, d->d_long.p, d->d_long_seg.p, d->d_chunk_ptr.p,     \
                                                   m->d_scanR.p, m->d_scanE.p, P.SP, n_dir == 2 ? m->d_ck_alpha.p : nullptr,                \
                                                   n_dir == 2 ? m->d_ck_beta.p : nullptr, d_ll, m->d_fallback.p, X);                       \
    JH_LAUNCHED();                                                                                                                        \
  }
  switch (G) {
    case 2: JH_SCAN(2) break;
    case 4: JH_SCAN(4) break;
    case 8: JH_SCAN(8) break;
    case 16: JH_SCAN(16) break;
    default: JH_SCAN(32) break;
  }
#undef JH_SCAN
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

static int run_loglik_exact(jh_model *m, jh_dataset *d, double *d_out, const int32_t *d_list, int64_t n_list);

static int run_loglik_sparse(jh_model *m, jh_dataset *d, double *d_out) {
  int r;
  if (m->opt.scan && m->fast.usable && !m->fast.e_global && d->max_len > chunk_len_for(m, d)) {
    if ((r = ensure_chunks(m, d, chunk_len_for(m, d)))) return r;
  }
  if (d->chunk_len == chunk_len_for(m, d) && scan_applies(m, d)) {  // long sequences parallel in the position, the others one warp each
    if ((r = m->d_fallback.reserve(d->n_seq + 2))) return r;
    JH_CUDA(cudaMemsetAsync(m->d_fallback.p, 0, sizeof(int32_t), m->st()));
    jhs::Pass1Extra X{};
    if ((r = run_scan_checkpoints(m, d, 1, d_out, X))) return r;
    if (d->n_seq > d->n_long) {
      X.list = d->d_order.p + d->n_long; X.n_list = d->n_seq - d->n_long;
      if ((r = run_sparse_pass1(m, d, 1, d_out, &X, nullptr, false))) return r;
    }
  } else {
    // banded models, few long sequences: the forward chain over the first half and the backward chain over the second half
    // of a sequence run at the same time and meet in the middle (kernels_band.cuh: k_band_meet) — half the dependent chain
    BandSel bsel;
    int64_t n_meet = 0;
    if (m->opt.band && m->opt.band_meet && band_select(m->sparse, &bsel)) {
      const int64_t min_len = 4096;
      int64_t lo = 0, hi = d->n_seq;  // d->order: decreasing length
      while (lo < hi) {
        const int64_t mid = (lo + hi) / 2;
        if (d->offsets[d->order[mid] + 1] - d->offsets[d->order[mid]] >= min_len) lo = mid + 1;
        else hi = mid;
      }
      if (lo > 0 && 2 * lo <= (int64_t)m->sm_count) n_meet = lo;  // worth it (and possible) only while every chain has an SM of its own
    }
    if (n_meet > 0) {
      const size_t SP = (size_t)m->sparse.SP;
      if ((r = m->d_meet_vec.reserve((size_t)n_meet * 2 * SP)) || (r = m->d_meet_exp.reserve((size_t)n_meet * 2)) || (r = m->d_meet_ok.reserve((size_t)n_meet * 2)))
        return r;
      jhs::Pass1Extra Xm{};
      Xm.list = d->d_order.p; Xm.n_list = n_meet;
      Xm.meet_vec = m->d_meet_vec.p; Xm.meet_exp = m->d_meet_exp.p; Xm.meet_ok = m->d_meet_ok.p;
      if ((r = run_sparse_pass1(m, d, 2, d_out, &Xm))) return r;
      if (d->n_seq > n_meet) {
        jhs::Pass1Extra Xs{};
        Xs.list = d->d_order.p + n_meet; Xs.n_list = d->n_seq - n_meet;
        if ((r = run_sparse_pass1(m, d, 1, d_out, &Xs, nullptr, false))) return r;
      }
    } else if ((r = run_sparse_pass1(m, d, 1, d_out)))
      return r;
  }
  int32_t n_fb = 0;
  JH_CUDA(cudaMemcpyAsync(&n_fb, m->d_fallback.p, sizeof(int32_t), cudaMemcpyDeviceToHost, m->st()));
  JH_CUDA(cudaStreamSynchronize(m->st()));
  if (n_fb > 0) return run_loglik_exact(m, d, d_out, m->d_fallback.p + 1, n_fb);
  return JH_OK;
}

static int run_fb_exact(jh_model *m, jh_dataset *d, int mode, PathOut d_paths, double *d_loglik, double *d_post, const int32_t *d_list, int64_t n_list,
                        const jhx::Cond &cond = jhx::Cond{}, int64_t budget_bytes = 0);

// Long-sequence path (kernels_sparse.cuh), two exact phases: checkpoint pass over whole sequences (k_sparse_pass1 /
// k_sparse_pass1_cta), then every chunk of `chunk` positions as an independent work item (posterior decisions:
// k_sparse_decode_chunks; expected counts: k_sparse_estep_chunks).  Scratch per resident warp: one chunk of rows.
//
// With several long sequences of different lengths the checkpoint pass is as long as the LONGEST sequence while most SMs
// idle (one CTA per sequence and direction).  With `overlap` every long sequence gets its own stream: checkpoint pass of the
// sequence, then the chunk pass over ITS chunks — the chunk passes of the shorter sequences fill the SMs while the long
// ones are still being swept.  Plain stream ordering, no device-side waiting.
enum { SPARSE_DECODE = 0, SPARSE_ESTEP = 1 };

static int run_sparse_two_phase(jh_model *m, jh_dataset *d, int what, PathOut d_paths, double *d_ll) {
  jhs::SparseTables &P = m->sparse;
  cudaStream_t s = m->st();
  const int64_t CK = chunk_len_for(m, d);
  int r = ensure_chunks(m, d, CK);
  if (r) return r;
  if ((r = m->d_fallback.reserve(d->n_seq + 2))) return r;
  const int64_t n_stats = m->n_stats();
  if (d->n_long > 0 || what == SPARSE_DECODE) {
    if ((r = m->d_ck_alpha.reserve((size_t)(d->n_chunks + 1) * P.SP)) || (r = m->d_ck_beta.reserve((size_t)(d->n_chunks + 1) * P.SP))) return r;
  }
  if (what == SPARSE_ESTEP && d->n_long > 0) {
    if ((r = m->d_ck_exp.reserve((size_t)2 * (d->n_chunks + 1))) || (r = m->d_seq_sfin.reserve(d->n_seq)) || (r = m->d_seq_AL.reserve(d->n_seq))) return r;
  }
  jhs::Pass1Extra X{};
  if (what == SPARSE_ESTEP) {
    X.ck_alpha_exp = m->d_ck_exp.p; X.ck_beta_exp = m->d_ck_exp.p + d->n_chunks + 1; X.seq_sfin = m->d_seq_sfin.p; X.seq_AL = m->d_seq_AL.p;
    X.score = m->d_stats.p + n_stats;
  }
  // small dense models: the checkpoint pass of the long sequences runs parallel in the position (kernels_scan.cuh)
  const bool scan = scan_applies(m, d);
  // ---- launch plan of the chunk pass: one launch, or one per long sequence (own stream) + one for the single-chunk sequences
  const bool streamed = !scan && m->opt.overlap && m->opt.pass1_cta && d->n_long >= 2 && d->n_long <= 64;
  const int64_t n_launch = streamed ? d->n_long + 1 : 1;
  const int NT = 128, wpc = NT / 32;
  const size_t smem = ((P.e_global ? 0 : (size_t)P.H * P.SP) + (size_t)wpc * 2 * P.SP) * sizeof(double);
  const int64_t rows = std::min<int64_t>(CK, d->max_len);
  const int64_t per_slot = rows * (P.SP * (int64_t)sizeof(double) + (what == SPARSE_ESTEP ? (int64_t)sizeof(int) : 0));
  int per_sm = 0;
  DISPATCH_SPARSE(P, {
    if (what == SPARSE_ESTEP) {
      auto kern = jhs::k_sparse_estep_chunks<SPL_, IND_, EG_>;
      if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    } else {
      auto kern = jhs::k_sparse_decode_chunks<SPL_, IND_, EG_>;
      if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    }
  });
  if (per_sm < 1) { jh_set_error("sparse chunk kernel does not fit on an SM"); return JH_ERR_UNSUPPORTED; }
  if (m->opt.blocks_per_sm > 0) per_sm = std::min<int>(per_sm, (int)m->opt.blocks_per_sm);
  const int64_t budget_slots = std::max<int64_t>(wpc, m->opt.scratch_mb * 1024 * 1024 / per_slot);
  // streamed: every launch gets its own slice of the row scratch.  The longest sequences finish their checkpoint pass last
  // and then have the GPU to themselves: they get full grids, the others a share of an SM's worth each (many are resident together).
  const int64_t full_grid = (int64_t)m->sm_count * per_sm;
  int64_t n_big = streamed ? std::min<int64_t>(3, d->n_long) : 0;
  while (n_big > 0 && (budget_slots / wpc - n_big * full_grid) / (n_launch - n_big) < 16) n_big--;
  const int64_t grid_cap = streamed ? std::max<int64_t>(1, std::min<int64_t>(m->sm_count, (budget_slots / wpc - n_big * full_grid) / (n_launch - n_big)))
                                    : std::max<int64_t>(1, std::min<int64_t>(full_grid, budget_slots / wpc));
  std::vector<int64_t> first(n_launch), count(n_launch), grid(n_launch), slot0(n_launch);
  int64_t total_slots = 0;
  for (int64_t k = 0; k < n_launch; k++) {
    first[k] = streamed ? d->long_seg[std::min<int64_t>(k, d->n_long)] : 0;
    count[k] = streamed ? (k < d->n_long ? d->long_seg[k + 1] - d->long_seg[k] : d->n_chunks - d->long_seg[d->n_long]) : d->n_chunks;
    grid[k] = std::max<int64_t>(1, std::min<int64_t>(k < n_big ? full_grid : grid_cap, (count[k] + wpc - 1) / wpc));
    slot0[k] = total_slots;
    total_slots += grid[k] * wpc;
  }
  if ((r = m->d_fwd.reserve((size_t)total_slots * rows * P.SP)) || (r = m->d_counters.reserve(n_launch))) return r;
  if (what == SPARSE_ESTEP && (r = m->d_rexp.reserve((size_t)total_slots * rows))) return r;
  JH_CUDA(cudaMemsetAsync(m->d_fallback.p, 0, sizeof(int32_t), s));
  JH_CUDA(cudaMemsetAsync(m->d_counters.p, 0, sizeof(unsigned long long) * n_launch, s));
  jhs::EstepTables ET{P.d_out_stat, P.d_pistat, (int)n_stats};
  if (what == SPARSE_ESTEP) JH_CUDA(cudaMemsetAsync(P.d_ec, 0, P.ec_bytes(), s));
  auto launch_chunks = [&](int64_t k, cudaStream_t st) -> int {
    if (count[k] <= 0) return JH_OK;
    double *rows_k = m->d_fwd.p + (size_t)slot0[k] * rows * P.SP;
    DISPATCH_SPARSE(P, {
      if (what == SPARSE_ESTEP) {
        jhs::k_sparse_estep_chunks<SPL_, IND_, EG_><<<(int)grid[k], NT, smem, st>>>(
            P.dev(), ET, d->dev(), (int)CK, (int)rows, count[k], d->d_chunk_seq.p + first[k], d->d_chunk_idx.p + first[k], d->d_chunk_ptr.p,
            m->d_ck_alpha.p, m->d_ck_beta.p, m->d_ck_exp.p, m->d_ck_exp.p + d->n_chunks + 1, m->d_seq_sfin.p, m->d_seq_AL.p, m->d_counters.p + k,
            rows_k, m->d_rexp.p + (size_t)slot0[k] * rows, m->d_stats.p, P.d_ec, m->d_fallback.p);
      } else {
        jhs::k_sparse_decode_chunks<SPL_, IND_, EG_><<<(int)grid[k], NT, smem, st>>>(
            P.dev(), d->dev(), (int)CK, (int)rows, count[k], d->d_chunk_seq.p + first[k], d->d_chunk_idx.p + first[k], d->d_chunk_ptr.p,
            m->d_ck_alpha.p, m->d_ck_beta.p, m->d_counters.p + k, rows_k, d_paths);
      }
      JH_LAUNCHED();
    });
    JH_CUDA(cudaGetLastError());
    return JH_OK;
  };
  if (scan) {
    if ((r = run_scan_checkpoints(m, d, 2, d_ll, X))) return r;
    if (what == SPARSE_DECODE && d->n_seq > d->n_long) {  // log-likelihoods of the single-chunk sequences
      jhs::Pass1Extra Xs{};
      Xs.list = d->d_order.p + d->n_long; Xs.n_list = d->n_seq - d->n_long;
      if ((r = run_sparse_pass1(m, d, 2, d_ll, &Xs, nullptr, false))) return r;
    }
    if (m->opt.lane_chunks && m->fast.G <= 8) {  // small models: one chunk per lane instead of one state per lane
      const jhf::FastTables &Ft = m->fast;
      const int ec_smem = what == SPARSE_ESTEP && (size_t)Ft.H * Ft.G * 32 * wpc * sizeof(double) <= 96 * 1024;
      const size_t smem_l = ((size_t)Ft.H * Ft.G + (ec_smem ? (size_t)Ft.H * Ft.G * 32 * wpc : 0)) * sizeof(double);
      const int64_t per_warp = rows * 32 * (Ft.G * (int64_t)sizeof(double) + (what == SPARSE_ESTEP ? (int64_t)sizeof(int) : 0));
      const int64_t warps = std::max<int64_t>(1, std::min<int64_t>((d->n_chunks + 31) / 32, std::min<int64_t>((int64_t)m->sm_count * 16, m->opt.scratch_mb * 1024 * 1024 / per_warp)));
      const int grid_l = (int)((warps + wpc - 1) / wpc);
      if ((r = m->d_fwd.reserve((size_t)grid_l * wpc * rows * 32 * Ft.G))) return r;
      if (what == SPARSE_ESTEP && (r = m->d_rexp.reserve((size_t)grid_l * wpc * rows * 32))) return r;
      if (what == SPARSE_ESTEP) JH_CUDA(cudaMemsetAsync(Ft.d_ec, 0, Ft.ec_bytes(), s));
      jhc::LaneChunkIn In{(int)CK, (int)rows, P.SP, d->n_chunks, d->d_chunk_seq.p, d->d_chunk_idx.p, d->d_chunk_ptr.p, m->d_ck_alpha.p, m->d_ck_beta.p,
                          m->d_ck_exp.p, m->d_ck_exp.p + d->n_chunks + 1, m->d_seq_sfin.p, m->d_seq_AL.p};
#define JH_LC(GG)                                                                                                                        \
  {                                                                                                                                      \
    if (what == SPARSE_ESTEP) {                                                                                                          \
      if (smem_l > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(jhc::k_lane_chunks<GG, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l)); \
      jhc::k_lane_chunks<GG, 1><<<grid_l, NT, smem_l, s>>>(Ft.dev(), d->dev(), In, m->d_counters.p, m->d_fwd.p, m->d_rexp.p, PathOut{nullptr, 0}, m->d_stats.p, Ft.d_ec, m->d_fallback.p, ec_smem); \
    } else {                                                                                                                             \
      if (smem_l > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(jhc::k_lane_chunks<GG, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l)); \
      jhc::k_lane_chunks<GG, 0><<<grid_l, NT, smem_l, s>>>(Ft.dev(), d->dev(), In, m->d_counters.p, m->d_fwd.p, nullptr, d_paths, nullptr, Ft.d_ec, m->d_fallback.p, 0); \
    }                                                                                                                                    \
    JH_LAUNCHED();                                                                                                                       \
  }
      if (Ft.G == 2) JH_LC(2) else if (Ft.G == 4) JH_LC(4) else JH_LC(8)
#undef JH_LC
      JH_CUDA(cudaGetLastError());
      if (what == SPARSE_ESTEP) {
        const int nf = Ft.H * Ft.G;
        jhf::k_fast_fold_ec<<<(nf + 7) / 8, 256, 0, s>>>(Ft.dev(), Ft.d_ec, m->d_stats.p);
        JH_LAUNCHED();
        JH_CUDA(cudaGetLastError());
        return JH_OK;
      }
    } else if ((r = launch_chunks(0, s)))
      return r;
  } else if (!streamed) {
    if (what == SPARSE_DECODE) {
      if ((r = run_sparse_pass1(m, d, 2, d_ll, nullptr, nullptr, false))) return r;
    } else if (d->n_long > 0) {
      X.list = d->d_long.p; X.n_list = d->n_long;
      if ((r = run_sparse_pass1(m, d, 2, nullptr, &X, nullptr, false))) return r;
    }
    if ((r = launch_chunks(0, s))) return r;
  } else {
    while ((int64_t)m->seq_streams.size() < d->n_long) {
      cudaStream_t st;
      cudaEvent_t ev;
      JH_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      m->seq_streams.push_back(st);
      JH_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      m->seq_events.push_back(ev);
    }
    if (!m->ev_fork) JH_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    JH_CUDA(cudaEventRecord(m->ev_fork, s));
    // Issue order matters: the streams share a handful of hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default) and a
    // kernel that waits for its stream blocks the queue behind it.  So: all checkpoint passes first (longest first, they
    // start at once), then the chunk passes in the order in which their sequences finish (shortest first).
    for (int64_t k = 0; k < d->n_long; k++) {
      cudaStream_t st = m->seq_streams[k];
      JH_CUDA(cudaStreamWaitEvent(st, m->ev_fork, 0));
      jhs::Pass1Extra Xk = X;
      Xk.list = d->d_long.p + k; Xk.n_list = 1;
      if ((r = run_sparse_pass1(m, d, 2, d_ll, &Xk, st, false))) return r;
    }
    for (int64_t k = d->n_long - 1; k >= 0; k--) {
      if ((r = launch_chunks(k, m->seq_streams[k]))) return r;
      JH_CUDA(cudaEventRecord(m->seq_events[k], m->seq_streams[k]));
    }
    // sequences of a single chunk: scoring pass for their log-likelihoods (decoding only), then their chunks, on the main stream
    if (what == SPARSE_DECODE && d->n_seq > d->n_long) {
      jhs::Pass1Extra Xs{};
      Xs.list = d->d_order.p + d->n_long; Xs.n_list = d->n_seq - d->n_long;
      if ((r = run_sparse_pass1(m, d, 2, d_ll, &Xs, nullptr, false))) return r;
    }
    if ((r = launch_chunks(d->n_long, s))) return r;
    for (int64_t k = 0; k < d->n_long; k++) JH_CUDA(cudaStreamWaitEvent(s, m->seq_events[k], 0));
  }
  if (what == SPARSE_ESTEP) {
    const int n = P.H * P.SP;
    jhs::k_sparse_fold_ec<<<(n + 7) / 8, 256, 0, s>>>(P.dev(), P.d_etab, P.d_emod, m->n_child, P.d_ec, m->d_stats.p);
    JH_LAUNCHED();
    JH_CUDA(cudaGetLastError());
    return JH_OK;
  }
  int32_t n_fb = 0;
  JH_CUDA(cudaMemcpyAsync(&n_fb, m->d_fallback.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  if (n_fb > 0) return run_fb_exact(m, d, JHX_MODE_DECODE, d_paths, d_ll, nullptr, m->d_fallback.p + 1, n_fb);
  return JH_OK;
}

static int run_decode_sparse(jh_model *m, jh_dataset *d, PathOut d_paths, double *d_ll) { return run_sparse_two_phase(m, d, SPARSE_DECODE, d_paths, d_ll); }
static int run_estep_sparse(jh_model *m, jh_dataset *d) { return run_sparse_two_phase(m, d, SPARSE_ESTEP, PathOut{nullptr, 0}, nullptr); }

static int ensure_octet_hashes(jh_model *m, jh_dataset *d);

// Octet scoring (kernels_mma.cuh): ONE launch over all octets (longest first, fetched dynamically: no scratch to size);
// sequences the scaled recursion could not score are re-run by the log-space kernel.
template <int NT, bool EG>
static int launch_mma_loglik_eg(jh_model *m, jh_dataset *d, double *d_out) {
  jhf::FastTables &F = m->fast;
  cudaStream_t s = m->st();
  auto kern = jhm::k_mma_loglik<NT, EG>;
  const int64_t n_oct = (int64_t)d->oct_hbase.size() - 1;
  size_t smem = EG ? 64 : (size_t)F.H * F.G * sizeof(double);
  if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
  if (per_sm < 1) { jh_set_error("octet scoring kernel does not fit on an SM"); return JH_ERR_UNSUPPORTED; }
  if (m->opt.blocks_per_sm > 0) per_sm = std::min<int>(per_sm, (int)m->opt.blocks_per_sm);
  int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)m->sm_count * per_sm, (n_oct + 7) / 8));
  jhm::OctData D;
  D.oct_seq = d->d_oct_seq.p; D.oct_hbase = d->d_oct_hbase.p; D.hashes = d->d_hashes.p; D.offsets = d->d_offsets.p;
  D.weights = nullptr; D.n_oct = n_oct;
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  kern<<<(int)grid, 256, smem, s>>>(F.dev(), D, m->d_counter.p, d_out, m->d_fallback.p);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

template <int NT>
static int launch_mma_loglik(jh_model *m, jh_dataset *d, double *d_out) {
  return m->fast.e_global ? launch_mma_loglik_eg<NT, true>(m, d, d_out) : launch_mma_loglik_eg<NT, false>(m, d, d_out);
}

// Lane-per-sequence kernels for 2..4 states (kernels_lanes.cuh).  mode 1: E-step of bucket b; mode 0: scoring of all octets.
template <int S, int MODE, bool EG>
static int launch_lanes_eg(jh_model *m, jh_dataset *d, size_t b, double *d_out) {
  jhf::FastTables &F = m->fast;
  cudaStream_t s = m->st();
  auto kern = jhl::k_lanes<S, MODE, EG>;
  const int64_t oct0 = MODE == 1 ? d->bucket_oct_ptr[b] : 0;
  const int64_t n_oct = MODE == 1 ? d->bucket_oct_ptr[b + 1] - oct0 : (int64_t)d->oct_hbase.size() - 1;
  const int64_t rows = MODE == 1 ? d->bucket_len[b] + 1 : 0;
  if (n_oct <= 0) return JH_OK;
  const int NT = 128, wpc = NT / 32, HS = F.H * S;
  const int ec_smem = !EG && MODE == 1 && (size_t)HS * 32 * wpc * sizeof(double) <= 96 * 1024;
  size_t smem = EG ? 64 : ((size_t)HS + (ec_smem ? (size_t)HS * 32 * wpc : 0)) * sizeof(double);
  if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
  if (per_sm < 1) { jh_set_error("lane-per-sequence kernel does not fit on an SM (smem %zu)", smem); return JH_ERR_UNSUPPORTED; }
  if (m->opt.blocks_per_sm > 0) per_sm = std::min<int>(per_sm, (int)m->opt.blocks_per_sm);
  const int64_t n_quads = (n_oct + 3) / 4;
  int64_t grid = std::min<int64_t>((int64_t)m->sm_count * per_sm, (n_quads + wpc - 1) / wpc);
  if constexpr (MODE == 1) {
    const int64_t per_slot = std::max<int64_t>(1, rows * 32 * (S * (int64_t)sizeof(double) + (int64_t)sizeof(int)));
    const int64_t max_slots = m->opt.scratch_mb * 1024 * 1024 / per_slot;
    if (max_slots < wpc) { jh_set_error("sequence too long for the DP scratch budget (%lld MB)", (long long)m->opt.scratch_mb); return JH_ERR_UNSUPPORTED; }
    grid = std::min(grid, max_slots / wpc);
  }
  grid = std::max<int64_t>(1, grid);
  const int64_t slots = grid * wpc;
  if (MODE == 1) {
    int r = m->d_fwd.reserve((size_t)slots * rows * 32 * S);
    if (r) return r;
    if ((size_t)slots * rows * 32 > F.fexp_n) {
      if (F.d_fexp) cudaFree(F.d_fexp);
      F.d_fexp = nullptr;
      F.fexp_n = 0;
      if (cudaMalloc(&F.d_fexp, (size_t)slots * rows * 32 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); jh_set_error("out of device memory (exponent scratch)"); return JH_ERR_NOMEM; }
      F.fexp_n = (size_t)slots * rows * 32;
    }
  }
  jhm::OctData D;
  D.oct_seq = d->d_oct_seq.p + oct0 * 8; D.oct_hbase = d->d_oct_hbase.p + oct0; D.hashes = d->d_hashes.p; D.offsets = d->d_offsets.p;
  D.weights = (MODE == 1 && d->has_weights) ? d->d_weights.p : nullptr; D.n_oct = n_oct;
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  kern<<<(int)grid, NT, smem, s>>>(F.dev(), D, m->d_counter.p, m->d_fwd.p, F.d_fexp, rows, m->d_stats.p, F.d_ec, m->d_fallback.p, d_out, ec_smem);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

template <int S, int MODE>
static int launch_lanes(jh_model *m, jh_dataset *d, size_t b, double *d_out) {
  return m->fast.e_global ? launch_lanes_eg<S, MODE, true>(m, d, b, d_out) : launch_lanes_eg<S, MODE, false>(m, d, b, d_out);
}

static int run_loglik_lanes(jh_model *m, jh_dataset *d, double *d_out) {
  int r = ensure_octet_hashes(m, d);
  if (r) return r;
  if ((r = m->d_fallback.reserve(d->n_seq + 2))) return r;
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemsetAsync(m->d_fallback.p, 0, sizeof(int32_t), s));
  r = m->fast.G == 2 ? launch_lanes<2, 0>(m, d, 0, d_out) : launch_lanes<4, 0>(m, d, 0, d_out);
  if (r) return r;
  int32_t n_fb = 0;
  JH_CUDA(cudaMemcpyAsync(&n_fb, m->d_fallback.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  if (n_fb > 0) return run_loglik_exact(m, d, d_out, m->d_fallback.p + 1, n_fb);
  return JH_OK;
}

static int run_loglik_octets(jh_model *m, jh_dataset *d, double *d_out) {
  int r = ensure_octet_hashes(m, d);
  if (r) return r;
  if ((r = m->d_fallback.reserve(d->n_seq + 2))) return r;
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemsetAsync(m->d_fallback.p, 0, sizeof(int32_t), s));
  r = m->fast.G == 8 ? launch_mma_loglik<1>(m, d, d_out) : m->fast.G == 16 ? launch_mma_loglik<2>(m, d, d_out) : launch_mma_loglik<4>(m, d, d_out);
  if (r) return r;
  int32_t n_fb = 0;
  JH_CUDA(cudaMemcpyAsync(&n_fb, m->d_fallback.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  if (n_fb > 0) return run_loglik_exact(m, d, d_out, m->d_fallback.p + 1, n_fb);
  return JH_OK;
}

extern "C" int jh_loglik(jh_model *m, jh_dataset *d, double *out) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (d->n_seq == 0) return JH_OK;
  if (!out) { jh_set_error("null output"); return JH_ERR_INVALID; }
  DevBuf<double> d_out;
  if ((r = d_out.alloc(d->n_seq))) return r;
  timing_begin(m);
  if (general_model(m)) {
    jhg::WorkList W{nullptr, nullptr, nullptr, d->n_seq};
    r = run_general_loglik(m, d, W, d_out.p);
  } else if (groups_active(m, d)) r = run_loglik_exact(m, d, d_out.p, nullptr, 0);
  else if (prefer_sparse(m, d)) r = run_loglik_sparse(m, d, d_out.p);
  else if (m->opt.fast && m->fast.usable && m->fast.G >= 8 && variant_of(m) == 0) r = run_loglik_octets(m, d, d_out.p);
  else if (m->opt.fast && m->fast.usable && m->fast.G <= 4 && variant_of(m) == 0) r = run_loglik_lanes(m, d, d_out.p);
  else if (m->opt.fast && m->fast.usable) r = jhf::run_loglik(m->fast, m->dev(), d->dev(), d->n_seq, d->max_len, d_out.p, m->d_counter.p, m->sm_count, m->st());
  else r = run_loglik_exact(m, d, d_out.p, nullptr, 0);
  timing_end(m);
  if (r) return r;
  JH_CUDA(cudaMemcpyAsync(out, d_out.p, sizeof(double) * d->n_seq, cudaMemcpyDeviceToHost, m->st()));
  JH_CUDA(cudaStreamSynchronize(m->st()));
  return JH_OK;
}

extern "C" int jh_loglik_windows(jh_model *m, jh_dataset *d, int64_t n_win, const int64_t *seq_index, const int64_t *start, const int64_t *end,
                                 double *out) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (n_win == 0) return JH_OK;
  if (n_win < 0 || !seq_index || !start || !end || !out) { jh_set_error("null argument"); return JH_ERR_INVALID; }
  std::vector<int32_t> sid(n_win);
  for (int64_t i = 0; i < n_win; i++) {
    if (seq_index[i] < 0 || seq_index[i] >= d->n_seq) { jh_set_error("window %lld: sequence index out of range", (long long)i); return JH_ERR_INVALID; }
    const int64_t L = d->offsets[seq_index[i] + 1] - d->offsets[seq_index[i]];
    if (start[i] < 0 || end[i] >= L || end[i] < start[i]) {  // the reference raises on these too (AbstractHMM.java:985-998)
      jh_set_error("window %lld: [%lld,%lld] is not inside a sequence of length %lld", (long long)i, (long long)start[i], (long long)end[i], (long long)L);
      return JH_ERR_INVALID;
    }
    sid[i] = (int32_t)seq_index[i];
  }
  DevBuf<int32_t> d_sid;
  DevBuf<int64_t> d_se;
  DevBuf<double> d_out;
  if ((r = d_sid.alloc(n_win)) || (r = d_se.alloc(2 * n_win)) || (r = d_out.alloc(n_win))) return r;
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemcpyAsync(d_sid.p, sid.data(), sizeof(int32_t) * n_win, cudaMemcpyHostToDevice, s));
  JH_CUDA(cudaMemcpyAsync(d_se.p, start, sizeof(int64_t) * n_win, cudaMemcpyHostToDevice, s));
  JH_CUDA(cudaMemcpyAsync(d_se.p + n_win, end, sizeof(int64_t) * n_win, cudaMemcpyHostToDevice, s));
  jhg::WorkList W{d_sid.p, d_se.p, d_se.p + n_win, n_win};
  timing_begin(m);
  r = run_general_loglik(m, d, W, d_out.p);
  timing_end(m);
  if (r) return r;
  JH_CUDA(cudaMemcpyAsync(out, d_out.p, sizeof(double) * n_win, cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  return JH_OK;
}

// windows of sequences -> device work list (validated like the reference's range checks, AbstractHMM.java:985-998)
struct DevWindows {
  DevBuf<int32_t> sid;
  DevBuf<int64_t> se;
  jhg::WorkList W{nullptr, nullptr, nullptr, 0};
  int64_t max_len = 0;
};
static int upload_windows(jh_model *m, jh_dataset *d, int64_t n_win, const int64_t *seq_index, const int64_t *start, const int64_t *end, DevWindows &out) {
  if (n_win < 0 || !seq_index || !start || !end) { jh_set_error("null argument"); return JH_ERR_INVALID; }
  std::vector<int32_t> sid(std::max<int64_t>(n_win, 1));
  for (int64_t i = 0; i < n_win; i++) {
    if (seq_index[i] < 0 || seq_index[i] >= d->n_seq) { jh_set_error("window %lld: sequence index out of range", (long long)i); return JH_ERR_INVALID; }
    const int64_t L = d->offsets[seq_index[i] + 1] - d->offsets[seq_index[i]];
    if (start[i] < 0 || end[i] >= L || end[i] < start[i]) {
      jh_set_error("window %lld: [%lld,%lld] is not inside a sequence of length %lld", (long long)i, (long long)start[i], (long long)end[i], (long long)L);
      return JH_ERR_INVALID;
    }
    sid[i] = (int32_t)seq_index[i];
    out.max_len = std::max(out.max_len, end[i] - start[i] + 1);
  }
  int r;
  if ((r = out.sid.alloc(n_win)) || (r = out.se.alloc(2 * n_win))) return r;
  cudaStream_t s = m->st();
  if (n_win) {
    JH_CUDA(cudaMemcpyAsync(out.sid.p, sid.data(), sizeof(int32_t) * n_win, cudaMemcpyHostToDevice, s));
    JH_CUDA(cudaMemcpyAsync(out.se.p, start, sizeof(int64_t) * n_win, cudaMemcpyHostToDevice, s));
    JH_CUDA(cudaMemcpyAsync(out.se.p + n_win, end, sizeof(int64_t) * n_win, cudaMemcpyHostToDevice, s));
    JH_CUDA(cudaStreamSynchronize(s));  // `sid` is a local
  }
  out.W = jhg::WorkList{out.sid.p, out.se.p, out.se.p + n_win, n_win};
  return JH_OK;
}

extern "C" int jh_posterior_window(jh_model *m, jh_dataset *d, int64_t seq_index, int64_t start, int64_t end, double *out) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (!out) { jh_set_error("null output"); return JH_ERR_INVALID; }
  DevWindows dw;
  if ((r = upload_windows(m, d, 1, &seq_index, &start, &end, dw))) return r;
  const int64_t L = end - start + 1;
  DevBuf<double> d_post;
  if ((r = d_post.alloc((size_t)m->S * L))) return r;
  if ((r = run_general_posterior(m, d, dw.W, L, d_post.p, PathOut{nullptr, 0}, nullptr))) return r;
  JH_CUDA(cudaMemcpyAsync(out, d_post.p, sizeof(double) * m->S * L, cudaMemcpyDeviceToHost, m->st()));
  JH_CUDA(cudaStreamSynchronize(m->st()));
  return JH_OK;
}

// ---- Viterbi ------------------------------------------------------------------------------------
static int ensure_octet_hashes(jh_model *m, jh_dataset *d);

// Lane-per-sequence Viterbi (kernels_viterbi.cuh): plain first-order models with sorted child lists; one launch per bucket.
template <int S, bool EG, bool U8>
static int launch_viterbi_lanes_eg(jh_model *m, jh_dataset *d, size_t b, PathOut paths, const int64_t *d_path_ptr, double *d_scores) {
  using C = jhv::VitCfg<S>;
  jhf::FastTables &F = m->fast;
  cudaStream_t s = m->st();
  auto kern = jhv::k_viterbi_lanes<S, EG, U8>;
  const int64_t oct0 = d->bucket_oct_ptr[b], n_oct = d->bucket_oct_ptr[b + 1] - oct0, rows = d->bucket_len[b];
  if (n_oct <= 0) return JH_OK;
  const int NT = 128, wpc = NT / 32;
  size_t smem = ((size_t)S * S + S + (EG ? 0 : (size_t)F.H * S)) * sizeof(double);
  if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
  if (per_sm < 1) { jh_set_error("Viterbi kernel does not fit on an SM (smem %zu)", smem); return JH_ERR_UNSUPPORTED; }
  if (m->opt.blocks_per_sm > 0) per_sm = std::min<int>(per_sm, (int)m->opt.blocks_per_sm);
  const int64_t n_quads = (n_oct + 3) / 4;
  int64_t grid = std::min<int64_t>((int64_t)m->sm_count * per_sm, (n_quads + wpc - 1) / wpc);
  const int64_t per_slot = rows * 32 * C::CB;
  const int64_t max_slots = m->opt.scratch_mb * 1024 * 1024 / per_slot;
  if (max_slots < wpc) { jh_set_error("sequence too long for the DP scratch budget (%lld MB)", (long long)m->opt.scratch_mb); return JH_ERR_UNSUPPORTED; }
  grid = std::max<int64_t>(1, std::min(grid, max_slots / wpc));
  int r = m->d_choice.reserve((size_t)grid * wpc * per_slot);
  if (r) return r;
  jhv::VitTables V;
  V.T = F.d_vT; V.pi = F.d_vpi; V.fin = F.d_vfin; V.E = F.d_vE; V.korder = F.d_korder; V.H = F.H; V.Kmax = F.Kmax; V.log_uniform = -log((double)F.a);
  jhm::OctData D;
  D.oct_seq = d->d_oct_seq.p + oct0 * 8; D.oct_hbase = d->d_oct_hbase.p + oct0; D.hashes = d->d_hashes.p; D.offsets = d->d_offsets.p;
  D.weights = nullptr; D.n_oct = n_oct;
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  kern<<<(int)grid, NT, smem, s>>>(V, D, m->d_counter.p, m->d_choice.p, rows, paths.p, d_path_ptr, d_scores);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

template <int S>
static int launch_viterbi_lanes(jh_model *m, jh_dataset *d, size_t b, PathOut paths, const int64_t *d_path_ptr, double *d_scores) {
  if (m->fast.e_global)
    return paths.u8 ? launch_viterbi_lanes_eg<S, true, true>(m, d, b, paths, d_path_ptr, d_scores) : launch_viterbi_lanes_eg<S, true, false>(m, d, b, paths, d_path_ptr, d_scores);
  return paths.u8 ? launch_viterbi_lanes_eg<S, false, true>(m, d, b, paths, d_path_ptr, d_scores) : launch_viterbi_lanes_eg<S, false, false>(m, d, b, paths, d_path_ptr, d_scores);
}

static int ensure_chunks(jh_model *m, jh_dataset *d, int64_t CK);

// Chromosome-scale Viterbi (kernels_vitlong.cuh): taken when the choice bytes of the data set's longest sequence would not
// leave scratch for a full grid of the ordinary kernels (one byte per (position, context) of every resident sequence), or
// when forced (tests).  Exact: rows stored at chunk boundaries by a sequential sweep, decisions recomputed per chunk.
static int run_viterbi_long(jh_model *m, jh_dataset *d, PathOut paths, const int64_t *d_path_ptr, double *d_scores, bool *handled) {
  *handled = false;
  jhs::SparseTables &P = m->sparse;
  if (!P.usable || P.e_global || m->S > 255) return JH_OK;
  const int64_t CK = m->opt.vit_chunk;
  if (d->max_len <= 2 * CK) return JH_OK;
  if (!m->opt.vit_chunk_force) {
    // ordinary kernels: rows x Cmax bytes per resident group (k_exact_viterbi), rows x 32 x max(S,4) per warp (k_viterbi_lanes)
    const int64_t budget = m->opt.scratch_mb * 1024 * 1024;
    const bool lanes = m->opt.fast && m->fast.usable && m->fast.sorted_children;
    const int64_t per_slot = lanes ? d->max_len * 32 * std::max(m->fast.G, 4) : d->max_len * m->Cmax;
    const int64_t want = lanes ? (int64_t)m->sm_count * 4 : (int64_t)m->sm_count * (256 / group_size(m));
    // ... and they need a sequence per lane / group to fill the GPU: a few long sequences are one dependent chain each
    // there, while the chunk phases of this path run in parallel whatever the number of sequences
    const bool enough = (lanes ? (d->n_seq + 31) / 32 : d->n_seq) >= want;
    if (enough && per_slot * want <= budget) return JH_OK;
    if (!enough && d->n_res < ((int64_t)1 << 22) && per_slot * std::max<int64_t>(1, lanes ? (d->n_seq + 31) / 32 : d->n_seq) <= budget) return JH_OK;  // small data: either way is fast
  }
  *handled = true;
  cudaStream_t s = m->st();
  int r = ensure_chunks(m, d, CK);
  if (r) return r;
  const int SP = P.SP;
  if ((r = m->d_vit_ck.reserve((size_t)(d->n_chunks + 1) * SP)) || (r = m->d_vit_map.reserve((size_t)(d->n_chunks + 1) * SP)) ||
      (r = m->d_vit_entry.reserve((size_t)d->n_chunks + 1)))
    return r;
  jhw::VitLongDev V{P.d_out_lT, P.d_lpi, P.d_pi_rank, P.d_lfin, P.d_lE, 0.0 + -log((double)m->a)};
  const int NT = 128, wpc = NT / 32;
  const size_t smem_c = (size_t)wpc * 2 * SP * sizeof(double), smem_s = (size_t)2 * SP * sizeof(double);
  const int64_t rows = std::min<int64_t>(CK, d->max_len);
  const int64_t per_warp = rows * SP;
  int per_sm = 0;
  BandSel bsel;
  const bool banded = m->opt.band && band_select(P, &bsel);
  // With sequences of different lengths phase A lasts as long as the LONGEST sequence while most SMs idle, and phases B-D
  // of everything would only start after it.  Like the E-step (run_sparse_two_phase) every long sequence gets its own
  // stream: sweep of the sequence -> its chunks (B) -> its composition (C) -> its chunks again (D), so the chunk phases of
  // the shorter sequences run while the long ones are still being swept.  Same issue-order rules as there.
  const bool streamed = m->opt.overlap && d->n_long >= 2 && d->n_long <= 64;
  const int64_t n_launch = streamed ? d->n_long + 1 : 1;
  if ((r = m->d_counters.reserve(3 * n_launch + 4))) return r;
  JH_CUDA(cudaMemsetAsync(m->d_counters.p, 0, sizeof(unsigned long long) * (3 * n_launch + 4), s));
  if (streamed) {
    while ((int64_t)m->seq_streams.size() < d->n_long) {
      cudaStream_t st;
      cudaEvent_t ev;
      JH_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      m->seq_streams.push_back(st);
      JH_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      m->seq_events.push_back(ev);
    }
    if (!m->ev_fork) JH_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
  }
  std::vector<int64_t> first(n_launch), count(n_launch), grid_k(n_launch), slot0(n_launch);
  size_t smem_iso = 0;
  if (streamed) {
    int optin = 0;
    JH_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
    smem_iso = (size_t)optin;
  }
#define JH_VL(VS_, VI_)                                                                                                               \
  {                                                                                                                                     \
    JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jhw::k_vit_chunks<VS_, VI_, 0>, NT, smem_c));                       \
    if (per_sm < 1) { jh_set_error("Viterbi chunk kernel does not fit on an SM"); return JH_ERR_UNSUPPORTED; }                          \
    if (m->opt.blocks_per_sm > 0) per_sm = std::min<int>(per_sm, (int)m->opt.blocks_per_sm);                                              \
    const int64_t full_grid = (int64_t)m->sm_count * per_sm, budget_ctas = std::max<int64_t>(1, m->opt.scratch_mb * 1024 * 1024 / per_warp / wpc); \
    int64_t n_big = streamed ? std::min<int64_t>(3, d->n_long) : 0;                                                                      \
    while (n_big > 0 && (budget_ctas - n_big * full_grid) / (n_launch - n_big) < 16) n_big--;                                            \
    const int64_t grid_cap = streamed ? std::max<int64_t>(1, std::min<int64_t>(m->sm_count, (budget_ctas - n_big * full_grid) / (n_launch - n_big))) \
                                      : std::max<int64_t>(1, std::min<int64_t>(full_grid, budget_ctas));                                 \
    int64_t total_ctas = 0;                                                                                                             \
    for (int64_t k = 0; k < n_launch; k++) {                                                                                            \
      first[k] = streamed ? d->long_seg[std::min<int64_t>(k, d->n_long)] : 0;                                                           \
      count[k] = streamed ? (k < d->n_long ? d->long_seg[k + 1] - d->long_seg[k] : d->n_chunks - d->long_seg[d->n_long]) : d->n_chunks;  \
      grid_k[k] = std::max<int64_t>(1, std::min<int64_t>(k < n_big ? full_grid : grid_cap, (count[k] + wpc - 1) / wpc));                 \
      slot0[k] = total_ctas;                                                                                                            \
      total_ctas += grid_k[k];                                                                                                          \
    }                                                                                                                                   \
    if ((r = m->d_vit_nxt.reserve((size_t)total_ctas * wpc * per_warp + 16))) return r;                                                  \
    auto sweep = [&](int64_t k0, int64_t nk, cudaStream_t st, unsigned long long *ctr) -> int {  /* phase A of long sequences [k0, k0 + nk) */ \
      /* streamed: the latency-bound warp claims the SM's whole shared memory, so that no chunk CTA becomes co-resident */              \
      const size_t iso = streamed ? smem_iso : 0;                                                                                       \
      if (banded) {                                                                                                                     \
        JH_BAND_DISPATCH(bsel, {                                                                                                        \
          auto kern = jhb::k_band_vit_sweep<SPL_, LO_, HI_>;                                                                            \
          if (iso > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iso));              \
          kern<<<(int)nk, 32, iso, st>>>(P.dev(), V, P.IND, d->dev(), (int)CK, d->d_long.p + k0, nk, d->d_chunk_ptr.p, m->d_vit_ck.p, d_scores); \
        });                                                                                                                             \
      } else {                                                                                                                          \
        auto kern = jhw::k_vit_sweep<VS_, VI_>;                                                                                         \
        const size_t sm_ = std::max(smem_s, iso);                                                                                       \
        if (sm_ > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_));                \
        kern<<<(int)std::min<int64_t>(nk, (int64_t)m->sm_count * 8), 32, sm_, st>>>(P.dev(), V, d->dev(), (int)CK, d->d_long.p + k0, nk,  \
                                                                                    d->d_chunk_ptr.p, m->d_vit_ck.p, ctr, d_scores);     \
      }                                                                                                                                 \
      JH_LAUNCHED();                                                                                                                    \
      return JH_OK;                                                                                                                     \
    };                                                                                                                                  \
    auto chunks_b = [&](int64_t k, int64_t c0, int64_t nc, cudaStream_t st, unsigned long long *ctr) {  /* phase B */                     \
      jhw::k_vit_chunks<VS_, VI_, 0><<<(int)grid_k[k], NT, smem_c, st>>>(P.dev(), V, d->dev(), (int)CK, nc, d->d_chunk_seq.p + c0,        \
                                                                        d->d_chunk_idx.p + c0, d->d_chunk_ptr.p, m->d_vit_ck.p, ctr,    \
                                                                        m->d_vit_nxt.p + (size_t)slot0[k] * wpc * per_warp, m->d_vit_map.p, \
                                                                        nullptr, PathOut{nullptr, 0}, nullptr, nullptr);                  \
      JH_LAUNCHED();                                                                                                                    \
    };                                                                                                                                  \
    auto chunks_d = [&](int64_t k, int64_t c0, int64_t nc, cudaStream_t st, unsigned long long *ctr) {  /* phase D */                     \
      if (nc <= 0) return;                                                                                                              \
      jhw::k_vit_chunks<VS_, VI_, 1><<<(int)grid_k[k], NT, smem_c, st>>>(P.dev(), V, d->dev(), (int)CK, nc, d->d_chunk_seq.p + c0,        \
                                                                        d->d_chunk_idx.p + c0, d->d_chunk_ptr.p, m->d_vit_ck.p, ctr,    \
                                                                        m->d_vit_nxt.p + (size_t)slot0[k] * wpc * per_warp, nullptr,     \
                                                                        m->d_vit_entry.p, paths, d_path_ptr, d_scores);                  \
      JH_LAUNCHED();                                                                                                                    \
    };                                                                                                                                  \
    auto compose = [&](int64_t k0, int64_t nk, cudaStream_t st) {  /* phase C */                                                         \
      jhw::k_vit_compose<<<(int)((nk + 63) / 64), 64, 0, st>>>(SP, (int)CK, d->d_long.p + k0, nk, d->d_offsets.p, d->d_chunk_ptr.p,       \
                                                              m->d_vit_map.p, m->d_vit_entry.p);                                        \
      JH_LAUNCHED();                                                                                                                    \
    };                                                                                                                                  \
    if (!streamed) {                                                                                                                    \
      if (d->n_long > 0) {                                                                                                              \
        if ((r = sweep(0, d->n_long, s, m->d_counters.p))) return r;                                                                    \
        chunks_b(0, 0, d->long_seg[d->n_long], s, m->d_counters.p + 1);                                                                 \
        compose(0, d->n_long, s);                                                                                                       \
      }                                                                                                                                 \
      chunks_d(0, 0, d->n_chunks, s, m->d_counters.p + 2);                                                                              \
    } else {                                                                                                                            \
      JH_CUDA(cudaEventRecord(m->ev_fork, s));                                                                                          \
      for (int64_t k = 0; k < d->n_long; k++) {  /* all sweeps first, longest first */                                                   \
        JH_CUDA(cudaStreamWaitEvent(m->seq_streams[k], m->ev_fork, 0));                                                                 \
        if ((r = sweep(k, 1, m->seq_streams[k], m->d_counters.p + 3 * k))) return r;                                                    \
      }                                                                                                                                 \
      for (int64_t k = d->n_long - 1; k >= 0; k--) {  /* chunk phases in the order in which the sweeps finish */                         \
        cudaStream_t st = m->seq_streams[k];                                                                                            \
        chunks_b(k, first[k], count[k], st, m->d_counters.p + 3 * k + 1);                                                               \
        compose(k, 1, st);                                                                                                              \
        chunks_d(k, first[k], count[k], st, m->d_counters.p + 3 * k + 2);                                                               \
        JH_CUDA(cudaEventRecord(m->seq_events[k], st));                                                                                 \
      }                                                                                                                                 \
      chunks_d(d->n_long, first[d->n_long], count[d->n_long], s, m->d_counters.p + 3 * d->n_long);  /* single-chunk sequences */         \
      for (int64_t k = 0; k < d->n_long; k++) JH_CUDA(cudaStreamWaitEvent(s, m->seq_events[k], 0));                                      \
    }                                                                                                                                   \
  }
  switch (P.SPL * 100 + P.IND) {
    case 104: JH_VL(1, 4) break;
    case 108: JH_VL(1, 8) break;
    case 116: JH_VL(1, 16) break;
    case 132: JH_VL(1, 32) break;
    case 204: JH_VL(2, 4) break;
    case 208: JH_VL(2, 8) break;
    case 216: JH_VL(2, 16) break;
    case 404: JH_VL(4, 4) break;
    case 408: JH_VL(4, 8) break;
    default: JH_VL(8, 4) break;
  }
#undef JH_VL
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

// Models without silent states and with transition order >= 1.  One launch per length bucket so that the per-slot scratch
// follows the bucket's longest sequence.  windows: sub-sequences (exact kernel only).
static int run_viterbi(jh_model *m, jh_dataset *d, PathOut paths, const int64_t *d_path_ptr, double *d_scores, double *d_stats,
                       const WorkList *windows = nullptr, int64_t win_max_len = 0) {
  const bool grouped = groups_active(m, d);
  if (!d_stats && !windows && !grouped) {  // chromosome-scale sequences: checkpointed sweep, choice bytes recomputed per chunk
    bool handled = false;
    int r = run_viterbi_long(m, d, paths, d_path_ptr, d_scores, &handled);
    if (r || handled) return r;
  }
  if (!d_stats && !windows && !grouped && m->opt.fast && m->fast.usable && m->fast.sorted_children) {  // decoding: one sequence per lane
    int r = ensure_octet_hashes(m, d);
    for (size_t b = 0; b + 1 < d->bucket_ptr.size() && !r; b++) {
      switch (m->fast.G) {
        case 2: r = launch_viterbi_lanes<2>(m, d, b, paths, d_path_ptr, d_scores); break;
        case 4: r = launch_viterbi_lanes<4>(m, d, b, paths, d_path_ptr, d_scores); break;
        case 8: r = launch_viterbi_lanes<8>(m, d, b, paths, d_path_ptr, d_scores); break;
        case 16: r = launch_viterbi_lanes<16>(m, d, b, paths, d_path_ptr, d_scores); break;
        default: r = launch_viterbi_lanes<32>(m, d, b, paths, d_path_ptr, d_scores); break;
      }
    }
    return r;
  }
  if (m->max_children > 254) { jh_set_error("more than 254 children per context"); return JH_ERR_UNSUPPORTED; }
  int G = group_size(m), r = JH_OK;
  cudaStream_t s = m->st();
  const size_t nb = windows ? 1 : d->bucket_ptr.size() - 1;
  for (size_t b = 0; b < nb && !r; b++) {
    int64_t first = windows ? 0 : d->bucket_ptr[b], count = windows ? windows->n : d->bucket_ptr[b + 1] - first;
    int64_t rows = windows ? win_max_len : d->bucket_len[b];
    const WorkList W = windows ? *windows : WorkList{nullptr, nullptr, nullptr, count};
    JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
    DISPATCH_G(G, {
      LaunchPlan pl;
      r = plan_launch(m, jhx::k_exact_viterbi<GG>, GG, 0, count, (size_t)rows * m->Cmax, &pl);
      if (!r) r = m->d_choice.reserve((size_t)pl.slots * rows * m->Cmax);
      if (!r) {
        jhx::k_exact_viterbi<GG><<<pl.grid, pl.threads, pl.smem, s>>>(m->dev(), d->dev(first, count), W, m->d_counter.p, m->d_choice.p, rows,
                                                                    paths, d_path_ptr, d_scores, d_stats);
        JH_LAUNCHED();
      }
    });
  }
  if (r) return r;
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

// scores on the host: -inf = "Impossible path" (HigherOrderHMM.java:659) -> JH_ERR_IMPOSSIBLE (outputs stay filled)
static int check_possible(const double *scores, int64_t n) {
  if (!scores) return JH_OK;
  for (int64_t i = 0; i < n; i++)
    if (scores[i] == JH_NEG_INF) {
      jh_set_error("Impossible path (sequence %lld has no state path with a probability > 0)", (long long)i);
      return JH_ERR_IMPOSSIBLE;
    }
  return JH_OK;
}

extern "C" int jh_viterbi_paths(jh_model *m, jh_dataset *d, int32_t *paths, const int64_t *path_ptr, int64_t *path_len, double *scores) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (d->n_seq == 0) return JH_OK;
  if (paths && !path_ptr) { jh_set_error("paths without path_ptr"); return JH_ERR_INVALID; }
  DevBuf<int32_t> d_paths;
  DevBuf<int64_t> d_ptr, d_len;
  DevBuf<double> d_scores;
  cudaStream_t s = m->st();
  const int64_t total = paths ? path_ptr[d->n_seq] : 0;
  if ((r = d_paths.alloc(total)) || (r = d_ptr.alloc(d->n_seq + 1)) || (r = d_len.alloc(d->n_seq)) || (r = d_scores.alloc(d->n_seq))) return r;
  if (paths) JH_CUDA(cudaMemcpyAsync(d_ptr.p, path_ptr, sizeof(int64_t) * (d->n_seq + 1), cudaMemcpyHostToDevice, s));
  timing_begin(m);
  if (general_model(m)) {
    r = run_general_viterbi(m, d, paths ? d_paths.p : nullptr, d_ptr.p, d_len.p, d_scores.p, nullptr);
  } else {  // one state per residue: the data-set kernels (k_viterbi_lanes / k_exact_viterbi), then the variable-length bookkeeping
    if (paths)
      for (int64_t i = 0; i < d->n_seq; i++)
        if (path_ptr[i + 1] - path_ptr[i] < d->offsets[i + 1] - d->offsets[i]) { jh_set_error("path capacity of sequence %lld is smaller than its length", (long long)i); return JH_ERR_INVALID; }
    const bool same = paths && std::equal(path_ptr, path_ptr + d->n_seq + 1, d->offsets.begin());
    r = run_viterbi(m, d, PathOut{paths ? d_paths.p : nullptr, 0}, same || !paths ? nullptr : d_ptr.p, d_scores.p, nullptr);
    if (!r) {
      jha::k_finish_paths<<<(int)std::min<int64_t>(d->n_seq, 148 * 8), 128, 0, s>>>(d_scores.p, d->d_offsets.p, nullptr, d->n_seq, paths ? d_paths.p : nullptr,
                                                                                 same || !paths ? nullptr : d_ptr.p, d_len.p);
      JH_LAUNCHED();
      JH_CUDA(cudaGetLastError());
    }
  }
  timing_end(m);
  if (r) return r;
  if (paths && total && (r = stage_d2h(paths, d_paths.p, sizeof(int32_t) * total, s, m->device, m->opt.stage != 0))) return r;
  if (path_len) JH_CUDA(cudaMemcpyAsync(path_len, d_len.p, sizeof(int64_t) * d->n_seq, cudaMemcpyDeviceToHost, s));
  if (scores) JH_CUDA(cudaMemcpyAsync(scores, d_scores.p, sizeof(double) * d->n_seq, cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  return JH_OK;
}

extern "C" int jh_viterbi_windows(jh_model *m, jh_dataset *d, int64_t n_win, const int64_t *seq_index, const int64_t *start, const int64_t *end,
                                  int32_t *paths, const int64_t *path_ptr, int64_t *path_len, double *scores) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (n_win == 0) return JH_OK;
  if (paths && !path_ptr) { jh_set_error("paths without path_ptr"); return JH_ERR_INVALID; }
  DevWindows dw;
  if ((r = upload_windows(m, d, n_win, seq_index, start, end, dw))) return r;
  DevBuf<int32_t> d_paths;
  DevBuf<int64_t> d_ptr, d_len;
  DevBuf<double> d_scores;
  cudaStream_t s = m->st();
  const int64_t total = paths ? path_ptr[n_win] : 0;
  if ((r = d_paths.alloc(total)) || (r = d_ptr.alloc(n_win + 1)) || (r = d_len.alloc(n_win)) || (r = d_scores.alloc(n_win))) return r;
  if (paths) JH_CUDA(cudaMemcpyAsync(d_ptr.p, path_ptr, sizeof(int64_t) * (n_win + 1), cudaMemcpyHostToDevice, s));
  timing_begin(m);
  if (general_model(m)) {
    r = run_general_viterbi(m, d, paths ? d_paths.p : nullptr, d_ptr.p, d_len.p, d_scores.p, nullptr, &dw.W);
  } else {  // group-per-window kernel (k_exact_viterbi), bit-exact like the whole-sequence call
    if (paths)
      for (int64_t i = 0; i < n_win; i++)
        if (path_ptr[i + 1] - path_ptr[i] < end[i] - start[i] + 1) { jh_set_error("path capacity of window %lld is smaller than its length", (long long)i); return JH_ERR_INVALID; }
    if (!paths) {  // the kernel indexes its outputs through path_ptr
      std::vector<int64_t> zero(n_win + 1, 0);
      JH_CUDA(cudaMemcpyAsync(d_ptr.p, zero.data(), sizeof(int64_t) * (n_win + 1), cudaMemcpyHostToDevice, s));
      JH_CUDA(cudaStreamSynchronize(s));
    }
    r = run_viterbi(m, d, PathOut{paths ? d_paths.p : nullptr, 0}, d_ptr.p, d_scores.p, nullptr, &dw.W, dw.max_len);
    if (!r) {
      jha::k_finish_paths<<<(int)std::min<int64_t>(n_win, 148 * 8), 128, 0, s>>>(d_scores.p, dw.se.p, dw.se.p + n_win, n_win, paths ? d_paths.p : nullptr,
                                                                              paths ? d_ptr.p : nullptr, d_len.p);
      JH_LAUNCHED();
      JH_CUDA(cudaGetLastError());
    }
  }
  timing_end(m);
  if (r) return r;
  if (paths && total && (r = stage_d2h(paths, d_paths.p, sizeof(int32_t) * total, s, m->device, m->opt.stage != 0))) return r;
  if (path_len) JH_CUDA(cudaMemcpyAsync(path_len, d_len.p, sizeof(int64_t) * n_win, cudaMemcpyDeviceToHost, s));
  if (scores) JH_CUDA(cudaMemcpyAsync(scores, d_scores.p, sizeof(double) * n_win, cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  return JH_OK;
}

// paths as int32 (paths) or one byte per residue (paths8, 255 = no state), whichever is not null: the kernels emit the
// requested form directly
static int viterbi_impl(jh_model *m, jh_dataset *d, int32_t *paths, uint8_t *paths8, double *scores) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (m->has_silent) { jh_set_error("models with silent states emit paths of variable length: use jh_viterbi_paths"); return JH_ERR_UNSUPPORTED; }
  if (paths8 && m->S > 255) { jh_set_error("byte paths need at most 255 states"); return JH_ERR_UNSUPPORTED; }
  if (d->n_seq == 0) return JH_OK;
  const bool u8 = paths8 != nullptr, want = paths || paths8;
  DevBuf<int32_t> d_paths;
  DevBuf<uint8_t> d_paths8;
  DevBuf<double> d_scores;
  if ((r = d_scores.alloc(d->n_seq))) return r;
  timing_begin(m);
  if (general_model(m)) {  // order 0: one state per residue, the data set offsets are the path pointers
    DevBuf<int64_t> d_len;
    if ((r = d_len.alloc(d->n_seq)) || (r = d_paths.alloc(d->n_res))) return r;
    r = run_general_viterbi(m, d, d_paths.p, d->d_offsets.p, d_len.p, d_scores.p, nullptr);
    if (!r && u8) {
      if ((r = d_paths8.alloc((size_t)d->n_res + 16))) return r;
      jha::k_narrow_paths<<<(int)std::min<int64_t>(148 * 8, (d->n_res + 4095) / 4096), 256, 0, m->st()>>>(d_paths.p, d->n_res, d_paths8.p);
      JH_LAUNCHED();
      JH_CUDA(cudaGetLastError());
    }
    if (!r) JH_CUDA(cudaStreamSynchronize(m->st()));
  } else {
    if (want && (r = u8 ? d_paths8.alloc((size_t)d->n_res + 16) : d_paths.alloc(d->n_res))) return r;
    r = run_viterbi(m, d, PathOut{want ? (u8 ? (void *)d_paths8.p : (void *)d_paths.p) : nullptr, u8 ? 1 : 0}, nullptr, d_scores.p, nullptr);
  }
  timing_end(m);
  if (r) return r;
  if (scores) JH_CUDA(cudaMemcpyAsync(scores, d_scores.p, sizeof(double) * d->n_seq, cudaMemcpyDeviceToHost, m->st()));
  if (paths && (r = stage_d2h(paths, d_paths.p, sizeof(int32_t) * d->n_res, m->st(), m->device, m->opt.stage != 0))) return r;
  if (paths8 && (r = stage_d2h(paths8, d_paths8.p, (size_t)d->n_res, m->st(), m->device, m->opt.stage != 0))) return r;
  JH_CUDA(cudaStreamSynchronize(m->st()));
  return check_possible(scores, d->n_seq);
}

extern "C" int jh_viterbi(jh_model *m, jh_dataset *d, int32_t *paths, double *scores) { return viterbi_impl(m, d, paths, nullptr, scores); }
extern "C" int jh_viterbi_u8(jh_model *m, jh_dataset *d, uint8_t *paths, double *scores) { return viterbi_impl(m, d, nullptr, paths, scores); }

// ---- forward-backward family (exact) ----------------------------------------------------------
// d_list with n_list < 0: the list is the device-side fallback list [count | ids...] (count read by the kernel).
// cond: the launch does its work only if *cond.flag == cond.want (decided on the device, no host round trip).
static int run_fb_exact(jh_model *m, jh_dataset *d, int mode, PathOut d_paths, double *d_loglik, double *d_post,
                        const int32_t *d_list, int64_t n_list, const jhx::Cond &cond, int64_t budget_bytes) {
  int G = group_size(m), r = JH_OK;
  if (cond.n_list_dev) n_list = d->n_seq;  // upper bound for the grid; the kernel reads the real length
  cudaStream_t s = m->st();
  int64_t n_stats = m->n_stats();
  int stats_in_smem = mode == JHX_MODE_COUNTS && n_stats * 8 <= 96 * 1024;
  size_t nb = d_list ? 1 : d->bucket_ptr.size() - 1;
  for (size_t b = 0; b < nb && !r; b++) {
    int64_t first = d_list ? 0 : d->bucket_ptr[b], count = d_list ? n_list : d->bucket_ptr[b + 1] - first;
    int64_t rows = (d_list ? d->max_len : d->bucket_len[b]) + 1;
    JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
    DISPATCH_G(G, {
      LaunchPlan pl;
      r = plan_launch(m, jhx::k_exact_fb<GG>, GG, stats_in_smem ? n_stats : 0, count, (size_t)rows * m->Cmax * sizeof(double), &pl, budget_bytes);
      if (!r) r = m->d_fwd.reserve((size_t)pl.slots * rows * m->Cmax);
      if (!r) {
        jhx::k_exact_fb<GG><<<pl.grid, pl.threads, pl.smem, s>>>(m->dev(), d->dev(first, count), m->d_counter.p, m->d_fwd.p, rows, mode,
                                                               m->d_stats.p, stats_in_smem, d_paths, d_loglik, d_post, d_list, n_list, cond);
        JH_LAUNCHED();
      }
    });
  }
  if (r) return r;
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

extern "C" int jh_posterior(jh_model *m, jh_dataset *d, int64_t seq_index, double *out) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (seq_index < 0 || seq_index >= d->n_seq || !out) { jh_set_error("sequence index out of range"); return JH_ERR_INVALID; }
  int64_t L = d->offsets[seq_index + 1] - d->offsets[seq_index];
  DevBuf<double> d_post;
  DevBuf<int32_t> d_list;
  if ((r = d_post.alloc((size_t)m->S * L)) || (r = d_list.alloc(1))) return r;
  int32_t sid = (int32_t)seq_index;
  JH_CUDA(cudaMemcpyAsync(d_list.p, &sid, sizeof(int32_t), cudaMemcpyHostToDevice, m->st()));
  if (general_model(m)) {  // general context graph: silent states score -inf (silentZero, HigherOrderHMM.java:316); order 0: HOH:296-308
    jhg::WorkList W{d_list.p, nullptr, nullptr, 1};
    if ((r = run_general_posterior(m, d, W, L, d_post.p, PathOut{nullptr, 0}, nullptr))) return r;
    JH_CUDA(cudaMemcpyAsync(out, d_post.p, sizeof(double) * m->S * L, cudaMemcpyDeviceToHost, m->st()));
    JH_CUDA(cudaStreamSynchronize(m->st()));
    return JH_OK;
  }
  // scratch sized for this sequence only
  int64_t saved = d->max_len;
  d->max_len = L;
  r = run_fb_exact(m, d, JHX_MODE_POSTERIOR, PathOut{nullptr, 0}, nullptr, d_post.p, d_list.p, 1);
  d->max_len = saved;
  if (r) return r;
  JH_CUDA(cudaMemcpyAsync(out, d_post.p, sizeof(double) * m->S * L, cudaMemcpyDeviceToHost, m->st()));
  JH_CUDA(cudaStreamSynchronize(m->st()));
  return JH_OK;
}

// Launch groups of the octet kernels.  Octets are consumed longest first from ONE queue per launch, so a launch ends with a
// tail as long as its last octets: with a launch per length bucket (about as many octets as warp slots each) that tail was
// a second wave of half-length octets per bucket, 40 % of the E-step on ragged data (cfg4).  Consecutive buckets that agree
// on chunked / full rows are therefore merged: normally one launch for the long (chunked) sequences and one for the rest,
// whose queue ends with the shortest sequences of the data set.  Full-row groups are only merged while the rows of the
// group's longest bucket for all its slots stay inside the budget.
struct OctGroup {
  size_t b0, b1;  // buckets [b0, b1)
  bool chunked;
};
// Positions per chunk of the chunked octet kernels.  Automatic: the recomputed forward rows of ONE chunk per resident warp
// (what the backward sweep re-reads at once) should stay in the 126 MB L2 instead of making an HBM round trip: about 80 MB for
// all warp slots -> 64 positions for 16 states, 32 for 32 states (cfg4: E-step 135-141 -> 128 ms, decoding 131-144 -> 121 ms
// against chunks of 1024 positions; below 32 positions the per-chunk overhead takes the gain back).
template <int NT>
static int64_t oct_chunk_of(const jh_model *m, int64_t full_grid) {
  using C = jhm::MmaCfg<NT>;
  if (m->opt.oct_chunk >= 0) return m->opt.oct_chunk;
  const int64_t row_bytes = 8 * (C::G * (int64_t)sizeof(double) + (int64_t)sizeof(int));
  const int64_t ck = ((int64_t)80 << 20) / (std::max<int64_t>(full_grid, 1) * C::WARPS * row_bytes);
  return std::min<int64_t>(1024, std::max<int64_t>(16, ck & ~(int64_t)7));
}
template <int NT>
static std::vector<OctGroup> oct_groups(const jh_model *m, const jh_dataset *d, int64_t full_grid) {
  using C = jhm::MmaCfg<NT>;
  std::vector<OctGroup> out;
  const int64_t budget = std::min<int64_t>(m->opt.scratch_mb, m->opt.oct_full_rows_mb) * 1024 * 1024;
  auto slots_of = [&](int64_t n_oct) { return std::min<int64_t>(full_grid, (n_oct + C::WARPS - 1) / C::WARPS) * C::WARPS; };
  auto per_slot_of = [&](size_t b) { return (d->bucket_len[b] + 1) * 8 * (C::G * (int64_t)sizeof(double) + (int64_t)sizeof(int)); };
  const int64_t CK = oct_chunk_of<NT>(m, full_grid);
  for (size_t b = 0; b + 1 < d->bucket_ptr.size(); b++) {
    const int64_t n_oct = d->bucket_oct_ptr[b + 1] - d->bucket_oct_ptr[b];
    if (n_oct <= 0) continue;
    const bool chunked = CK > 0 && d->bucket_len[b] + 1 > 2 * CK &&
                         (m->opt.oct_chunk_force || slots_of(n_oct) * per_slot_of(b) > budget);
    if (!out.empty() && out.back().chunked == chunked && out.back().b1 == b) {
      const int64_t n_all = d->bucket_oct_ptr[b + 1] - d->bucket_oct_ptr[out.back().b0];
      if (chunked || slots_of(n_all) * per_slot_of(out.back().b0) <= budget) {
        out.back().b1 = b + 1;
        continue;
      }
    }
    out.push_back(OctGroup{b, b + 1, chunked});
  }
  return out;
}

// Octet posterior decoding (kernels_mma.cuh), one launch per group of length buckets; long buckets chunked like the E-step.
template <int NT, bool CHUNKED, bool EG>
static int launch_mma_decode_impl(jh_model *m, jh_dataset *d, OctGroup g, int64_t grid, size_t smem, PathOut d_paths, double *d_ll) {
  using C = jhm::MmaCfg<NT>;
  jhf::FastTables &F = m->fast;
  cudaStream_t s = m->st();
  auto kern = jhm::k_mma_decode<NT, CHUNKED, EG>;
  const int64_t oct0 = d->bucket_oct_ptr[g.b0], n_oct = d->bucket_oct_ptr[g.b1] - oct0, L = d->bucket_len[g.b0];
  const int64_t CK = oct_chunk_of<NT>(m, grid);
  grid = std::min<int64_t>(grid, (n_oct + C::WARPS - 1) / C::WARPS);
  const int64_t rows = CHUNKED ? CK + 1 : L + 1, ck_rows = CHUNKED ? (L + CK - 1) / CK + 1 : 0;
  if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t per_slot = (rows + ck_rows) * 8 * (C::G * (int64_t)sizeof(double) + (int64_t)sizeof(int));
  const int64_t max_slots = m->opt.scratch_mb * 1024 * 1024 / per_slot;
  if (max_slots < C::WARPS) { jh_set_error("sequence too long for the DP scratch budget (%lld MB)", (long long)m->opt.scratch_mb); return JH_ERR_UNSUPPORTED; }
  grid = std::max<int64_t>(1, std::min(grid, max_slots / C::WARPS));
  const int64_t slots = grid * C::WARPS;
  int r = m->d_fwd.reserve((size_t)slots * (rows + ck_rows) * 8 * C::G);
  if (r) return r;
  if ((size_t)slots * (rows + ck_rows) * 8 > F.fexp_n) {
    if (F.d_fexp) cudaFree(F.d_fexp);
    F.d_fexp = nullptr;
    F.fexp_n = 0;
    if (cudaMalloc(&F.d_fexp, (size_t)slots * (rows + ck_rows) * 8 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); jh_set_error("out of device memory (exponent scratch)"); return JH_ERR_NOMEM; }
    F.fexp_n = (size_t)slots * (rows + ck_rows) * 8;
  }
  jhm::OctData D;
  D.oct_seq = d->d_oct_seq.p + oct0 * 8; D.oct_hbase = d->d_oct_hbase.p + oct0; D.hashes = d->d_hashes.p; D.offsets = d->d_offsets.p;
  D.weights = nullptr; D.n_oct = n_oct;
  jhm::OctChunks X{(int)CK, ck_rows, m->d_fwd.p + (size_t)slots * rows * 8 * C::G, F.d_fexp + (size_t)slots * rows * 8};
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  kern<<<(int)grid, C::NTHREADS, smem, s>>>(F.dev(), D, m->d_counter.p, m->d_fwd.p, F.d_fexp, rows, d_paths, d_ll, m->d_fallback.p, X);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

template <int NT>
static int launch_mma_decode(jh_model *m, jh_dataset *d, PathOut d_paths, double *d_ll) {
  using C = jhm::MmaCfg<NT>;
  jhf::FastTables &F = m->fast;
  const bool eg = F.e_global;
  const size_t smem = (eg ? 0 : (size_t)F.H * C::G * sizeof(double)) + (size_t)C::WARPS * C::RING * C::SLOTB;
  if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(jhm::k_mma_decode<NT, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jhm::k_mma_decode<NT, false, false>, C::NTHREADS, smem));
  if (per_sm < 1) { jh_set_error("octet decoding kernel does not fit on an SM (smem %zu)", smem); return JH_ERR_UNSUPPORTED; }
  if (m->opt.blocks_per_sm > 0) per_sm = std::min<int>(per_sm, (int)m->opt.blocks_per_sm);
  const int64_t grid = (int64_t)m->sm_count * per_sm;
  int r = JH_OK;
  for (const OctGroup &g : oct_groups<NT>(m, d, grid)) {
    if (eg) r = g.chunked ? launch_mma_decode_impl<NT, true, true>(m, d, g, grid, smem, d_paths, d_ll) : launch_mma_decode_impl<NT, false, true>(m, d, g, grid, smem, d_paths, d_ll);
    else r = g.chunked ? launch_mma_decode_impl<NT, true, false>(m, d, g, grid, smem, d_paths, d_ll) : launch_mma_decode_impl<NT, false, false>(m, d, g, grid, smem, d_paths, d_ll);
    if (r) return r;
  }
  return JH_OK;
}

static int run_decode_octets(jh_model *m, jh_dataset *d, PathOut d_paths, double *d_ll) {
  int r = ensure_octet_hashes(m, d);
  if (r) return r;
  if ((r = m->d_fallback.reserve(d->n_seq + 2))) return r;
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemsetAsync(m->d_fallback.p, 0, sizeof(int32_t), s));
  r = m->fast.G == 8 ? launch_mma_decode<1>(m, d, d_paths, d_ll) : m->fast.G == 16 ? launch_mma_decode<2>(m, d, d_paths, d_ll)
                                                                                   : launch_mma_decode<4>(m, d, d_paths, d_ll);
  if (r) return r;
  int32_t n_fb = 0;
  JH_CUDA(cudaMemcpyAsync(&n_fb, m->d_fallback.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  if (n_fb > 0) return run_fb_exact(m, d, JHX_MODE_DECODE, d_paths, d_ll, nullptr, m->d_fallback.p + 1, n_fb);
  return JH_OK;
}

static int posterior_decode_impl(jh_model *m, jh_dataset *d, int32_t *paths, uint8_t *paths8, double *loglik) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (d->n_seq == 0) return JH_OK;
  if (paths8 && m->S > 255) { jh_set_error("byte paths need at most 255 states"); return JH_ERR_UNSUPPORTED; }
  const bool u8 = paths8 != nullptr;
  DevBuf<int32_t> d_paths;
  DevBuf<uint8_t> d_paths8;
  DevBuf<double> d_ll;
  if ((r = u8 ? d_paths8.alloc((size_t)d->n_res + 16) : d_paths.alloc(d->n_res)) || (r = d_ll.alloc(d->n_seq))) return r;
  const PathOut po{u8 ? (void *)d_paths8.p : (void *)d_paths.p, u8 ? 1 : 0};
  timing_begin(m);
  if (general_model(m)) {
    jhg::WorkList W{nullptr, nullptr, nullptr, d->n_seq};
    r = run_general_posterior(m, d, W, d->max_len, nullptr, po, d_ll.p);
  } else
  if (prefer_sparse(m, d)) r = run_decode_sparse(m, d, po, d_ll.p);
  else if (m->opt.fast && m->fast.usable && m->fast.G >= 8 && variant_of(m) == 0) r = run_decode_octets(m, d, po, d_ll.p);
  else r = run_fb_exact(m, d, JHX_MODE_DECODE, po, d_ll.p, nullptr, nullptr, 0);
  timing_end(m);
  if (r) return r;
  if (loglik) JH_CUDA(cudaMemcpyAsync(loglik, d_ll.p, sizeof(double) * d->n_seq, cudaMemcpyDeviceToHost, m->st()));
  if (paths && (r = stage_d2h(paths, d_paths.p, sizeof(int32_t) * d->n_res, m->st(), m->device, m->opt.stage != 0))) return r;
  if (paths8 && (r = stage_d2h(paths8, d_paths8.p, (size_t)d->n_res, m->st(), m->device, m->opt.stage != 0))) return r;
  JH_CUDA(cudaStreamSynchronize(m->st()));
  return JH_OK;
}

extern "C" int jh_posterior_decode(jh_model *m, jh_dataset *d, int32_t *paths, double *loglik) { return posterior_decode_impl(m, d, paths, nullptr, loglik); }
extern "C" int jh_posterior_decode_u8(jh_model *m, jh_dataset *d, uint8_t *paths, double *loglik) { return posterior_decode_impl(m, d, nullptr, paths, loglik); }

// ---- E-step / M-step / EM ------------------------------------------------------------------------
// Scaled linear-space E-step (kernels_fast.cuh), one launch per length bucket.
template <int G, int NT, bool POW2, int NS>
static int launch_fast_estep(jh_model *m, jh_dataset *d, int64_t first, int64_t count, int64_t rows) {
  jhf::FastTables &F = m->fast;
  cudaStream_t s = m->st();
  auto kern = jhf::k_fast_estep<G, NT, POW2, NS>;
  const int gpc = NT / G;
  // tables T, Tt, O-reduction [3 G^2], E [H G], pi [G], broadcast vectors [gpc][2][NS][G], hash tiles, raw tiles, cp.async ring
  size_t smem = ((size_t)3 * G * G + (size_t)F.H * G + G + (size_t)gpc * 2 * NS * G) * sizeof(double) +
                (size_t)gpc * NS * (16 * G) * 2 + (size_t)gpc * NS * (16 * G + 16) + (size_t)gpc * NS * 8 * (G * 8 + 16);
  if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
  if (per_sm < 1) { jh_set_error("fast E-step kernel does not fit on an SM (smem %zu)", smem); return JH_ERR_UNSUPPORTED; }
  if (m->opt.blocks_per_sm > 0) per_sm = std::min<int>(per_sm, (int)m->opt.blocks_per_sm);
  const int spc = gpc * NS;  // sequences in flight per CTA
  int64_t grid = std::min<int64_t>((int64_t)m->sm_count * per_sm, (count + spc - 1) / spc);
  int64_t per_slot = rows * G * (int64_t)sizeof(double) + rows * (int64_t)sizeof(int);
  int64_t max_slots = m->opt.scratch_mb * 1024 * 1024 / per_slot;
  if (max_slots < spc) { jh_set_error("sequence too long for the DP scratch budget (%lld MB)", (long long)m->opt.scratch_mb); return JH_ERR_UNSUPPORTED; }
  grid = std::max<int64_t>(1, std::min(grid, max_slots / spc));
  int64_t slots = grid * spc;
  int r = m->d_fwd.reserve((size_t)slots * rows * G);
  if (r) return r;
  if ((size_t)slots * rows > F.fexp_n) {
    if (F.d_fexp) cudaFree(F.d_fexp);
    F.d_fexp = nullptr;
    F.fexp_n = 0;
    if (cudaMalloc(&F.d_fexp, (size_t)slots * rows * sizeof(int)) != cudaSuccess) { cudaGetLastError(); jh_set_error("out of device memory (exponent scratch)"); return JH_ERR_NOMEM; }
    F.fexp_n = (size_t)slots * rows;
  }
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  kern<<<(int)grid, NT, smem, s>>>(F.dev(), d->dev(first, count), m->d_counter.p, m->d_fwd.p, F.d_fexp, rows, m->d_stats.p, F.d_ec, m->d_fallback.p);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

// Context-hash records of the octets (kernels_mma.cuh), built once per data set and (alphabet, emission order).
static int ensure_octet_hashes(jh_model *m, jh_dataset *d) {
  if (d->hash_a == m->a && d->hash_K == m->Kmax && d->d_hashes.p) return JH_OK;
  cudaStream_t s = m->st();
  int r;
  const int64_t n_oct = (int64_t)d->oct_hbase.size() - 1, n_rec = d->oct_hbase[n_oct];
  if (!d->d_oct_seq.p) {
    if ((r = d->d_oct_seq.upload(d->oct_seq, s)) || (r = d->d_oct_hbase.upload(d->oct_hbase, s))) return r;
  }
  if ((r = d->d_hashes.reserve((size_t)(n_rec + 4) * 8))) return r;  // one zero record in front, three behind (prefetch overrun)
  JH_CUDA(cudaMemsetAsync(d->d_hashes.p, 0, (size_t)(n_rec + 4) * 8 * sizeof(uint16_t), s));
  if (n_oct > 0) {
    jhm::k_hash_octets<<<(int)std::min<int64_t>(n_oct, 148 * 16), 128, 0, s>>>(d->d_codes.p, d->d_offsets.p, d->d_oct_seq.p, d->d_oct_hbase.p, n_oct,
                                                                           m->a, m->Kmax, d->d_hashes.p);
    JH_LAUNCHED();
    JH_CUDA(cudaGetLastError());
  }
  JH_CUDA(cudaStreamSynchronize(s));  // the host vectors are borrowed by the async uploads
  d->hash_a = m->a;
  d->hash_K = m->Kmax;
  return JH_OK;
}

// Octet E-step (kernels_mma.cuh): one launch per length bucket, 8 warps per CTA, one CTA per SM, octets fetched dynamically.
// Buckets whose full forward rows would not leave scratch for the whole grid run the chunked variant (checkpoints every
// `oct_chunk` positions, rows of one chunk per resident warp).
// fixed-point scale of the deterministic emission counts: the largest power of two that keeps sum_n w_n L_n below 2^62
static double det_scale_of(const jh_dataset *d) {
  const double total = std::max(1.0, d->has_weights ? d->weighted_total : (double)d->n_res);
  int e = 0;
  frexp(total, &e);  // total < 2^e
  return ldexp(1.0, 62 - e);
}

template <int NT, bool CHUNKED, bool EG, bool DET = false>
static int launch_mma_estep_impl(jh_model *m, jh_dataset *d, OctGroup g, int64_t grid, size_t smem) {
  using C = jhm::MmaCfg<NT>;
  jhf::FastTables &F = m->fast;
  cudaStream_t s = m->st();
  auto kern = jhm::k_mma_estep<NT, CHUNKED, EG, DET>;
  const int64_t oct0 = d->bucket_oct_ptr[g.b0], n_oct = d->bucket_oct_ptr[g.b1] - oct0, L = d->bucket_len[g.b0];
  const int64_t CK = oct_chunk_of<NT>(m, grid);
  grid = std::min<int64_t>(grid, (n_oct + C::WARPS - 1) / C::WARPS);
  const int64_t rows = CHUNKED ? CK + 1 : L + 1, ck_rows = CHUNKED ? (L + CK - 1) / CK + 1 : 0;
  if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t per_slot = (rows + ck_rows) * 8 * (C::G * (int64_t)sizeof(double) + (int64_t)sizeof(int));
  const int64_t max_slots = m->opt.scratch_mb * 1024 * 1024 / per_slot;
  if (max_slots < C::WARPS) { jh_set_error("sequence too long for the DP scratch budget (%lld MB)", (long long)m->opt.scratch_mb); return JH_ERR_UNSUPPORTED; }
  grid = std::max<int64_t>(1, std::min(grid, max_slots / C::WARPS));
  const int64_t slots = grid * C::WARPS;
  int r = m->d_fwd.reserve((size_t)slots * (rows + ck_rows) * 8 * C::G);
  if (r) return r;
  if ((size_t)slots * (rows + ck_rows) * 8 > F.fexp_n) {
    if (F.d_fexp) cudaFree(F.d_fexp);
    F.d_fexp = nullptr;
    F.fexp_n = 0;
    if (cudaMalloc(&F.d_fexp, (size_t)slots * (rows + ck_rows) * 8 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); jh_set_error("out of device memory (exponent scratch)"); return JH_ERR_NOMEM; }
    F.fexp_n = (size_t)slots * (rows + ck_rows) * 8;
  }
  jhm::OctData D;
  D.oct_seq = d->d_oct_seq.p + oct0 * 8; D.oct_hbase = d->d_oct_hbase.p + oct0; D.hashes = d->d_hashes.p; D.offsets = d->d_offsets.p;
  D.weights = d->has_weights ? d->d_weights.p : nullptr; D.n_oct = n_oct;
  jhm::OctChunks X{(int)CK, ck_rows, m->d_fwd.p + (size_t)slots * rows * 8 * C::G, F.d_fexp + (size_t)slots * rows * 8};
  JH_CUDA(cudaMemsetAsync(m->d_counter.p, 0, sizeof(unsigned long long), s));
  jhm::DetOut det{nullptr, 0.0};
  const int64_t NC = (int64_t)C::G * C::G + C::G + 1;
  if (DET) {
    if ((r = m->d_det.reserve((size_t)(slots + 1) * NC + (size_t)C::G * F.H))) return r;
    det.part = m->d_det.p;
    det.scale = det_scale_of(d);
  }
  kern<<<(int)grid, C::NTHREADS, smem, s>>>(F.dev(), D, m->d_counter.p, m->d_fwd.p, F.d_fexp, rows, m->d_stats.p, F.d_ec, m->d_fallback.p, X, det);
  JH_LAUNCHED();
  if (DET) {  // the launch's partial sums, added in slot order
    jhm::k_det_fold_parts<<<1, 256, 0, s>>>(F.dev(), m->d_det.p, slots, m->d_det.p + (size_t)slots * NC, m->d_stats.p);
    JH_LAUNCHED();
  }
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

template <int NT>
static int launch_mma_estep(jh_model *m, jh_dataset *d) {
  using C = jhm::MmaCfg<NT>;
  jhf::FastTables &F = m->fast;
  const bool eg = F.e_global;
  const size_t smem = C::smem_bytes(eg ? 0 : F.H);
  int per_sm = 0;
  if (smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(jhm::k_mma_estep<NT, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  JH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jhm::k_mma_estep<NT, false, false>, C::NTHREADS, smem));
  if (per_sm < 1) { jh_set_error("octet E-step kernel does not fit on an SM (smem %zu)", smem); return JH_ERR_UNSUPPORTED; }
  if (m->opt.blocks_per_sm > 0) per_sm = std::min<int>(per_sm, (int)m->opt.blocks_per_sm);
  const int64_t grid = (int64_t)m->sm_count * per_sm;
  int r = JH_OK;
  if (m->opt.deterministic && eg) { jh_set_error("deterministic mode: emission tables larger than shared memory are not covered"); return JH_ERR_UNSUPPORTED; }
  for (const OctGroup &g : oct_groups<NT>(m, d, grid)) {
    if (m->opt.deterministic) r = g.chunked ? launch_mma_estep_impl<NT, true, false, true>(m, d, g, grid, smem) : launch_mma_estep_impl<NT, false, false, true>(m, d, g, grid, smem);
    else if (eg) r = g.chunked ? launch_mma_estep_impl<NT, true, true>(m, d, g, grid, smem) : launch_mma_estep_impl<NT, false, true>(m, d, g, grid, smem);
    else r = g.chunked ? launch_mma_estep_impl<NT, true, false>(m, d, g, grid, smem) : launch_mma_estep_impl<NT, false, false>(m, d, g, grid, smem);
    if (r) return r;
  }
  return JH_OK;
}

static int run_estep_fast(jh_model *m, jh_dataset *d) {
  int r = JH_OK;
  JH_CUDA(cudaMemsetAsync(m->fast.d_ec, 0, m->fast.ec_bytes(), m->st()));
  const bool octets = variant_of(m) == 0 && m->fast.G >= 8;  // 8 sequences per warp on the FP64 MMA path
  const bool lanes = variant_of(m) == 0 && m->fast.G <= 4;  // one sequence per lane
  if (m->opt.deterministic && !octets) { jh_set_error("deterministic mode covers the octet kernels (5..32 states, fast_variant 0)"); return JH_ERR_UNSUPPORTED; }
  for (size_t b = 0; lanes && b + 1 < d->bucket_ptr.size() && !r; b++) r = m->fast.G == 2 ? launch_lanes<2, 1>(m, d, b, nullptr) : launch_lanes<4, 1>(m, d, b, nullptr);
  if (octets) r = m->fast.G == 8 ? launch_mma_estep<1>(m, d) : m->fast.G == 16 ? launch_mma_estep<2>(m, d) : launch_mma_estep<4>(m, d);
  for (size_t b = 0; !octets && !lanes && b + 1 < d->bucket_ptr.size() && !r; b++) {
    int64_t first = d->bucket_ptr[b], count = d->bucket_ptr[b + 1] - first, rows = d->bucket_len[b] + 1;
    const bool p2 = (m->a & (m->a - 1)) == 0;
#define JH_FE(GG, NTT, NSS) (p2 ? launch_fast_estep<GG, NTT, true, NSS>(m, d, first, count, rows) : launch_fast_estep<GG, NTT, false, NSS>(m, d, first, count, rows))
    switch (m->fast.G) {
      case 2: r = JH_FE(2, 512, 2); break;
      case 4: r = JH_FE(4, 512, 2); break;
      case 8: r = JH_FE(8, 512, 2); break;
      case 16: r = JH_FE(16, 256, 2); break;
      default:  // fast_variant: 0 = octets on the FP64 MMA path (above); 3 = two sequences per warp in lock step, 1 = one sequence per warp, 2 = one sequence, 12 warps
        r = m->opt.fast_variant == 1 ? JH_FE(32, 256, 1) : m->opt.fast_variant == 2 ? JH_FE(32, 384, 1) : JH_FE(32, 256, 2);
        break;
    }
#undef JH_FE
  }
  if (r) return r;
  int n = m->fast.H * m->fast.G;
  if (m->opt.deterministic) {
    jhm::k_det_fold_ec<<<1, 256, 0, m->st()>>>(m->fast.dev(), reinterpret_cast<const unsigned long long *>(m->fast.d_ec), 1.0 / det_scale_of(d),
                                               m->d_det.p, m->d_stats.p);
    JH_LAUNCHED();
    JH_CUDA(cudaGetLastError());
    return JH_OK;
  }
  jhf::k_fast_fold_ec<<<(n + 7) / 8, 256, 0, m->st()>>>(m->fast.dev(), m->fast.d_ec, m->d_stats.p);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

static int estep_device(jh_model *m, jh_dataset *d, int mode) {
  cudaStream_t s = m->st();
  JH_CUDA(cudaMemsetAsync(m->d_stats.p, 0, sizeof(double) * (m->n_stats() + 8), s));  // resetStatistics (HOH:890-895)
  if (d->n_seq == 0) return JH_OK;
  if (general_model(m)) {  // general context graph: one sequence per thread, reference order
    if (!m->capturing) cudaEventRecord(m->evk0, s);
    int r = mode == JH_MODE_VITERBI ? run_general_viterbi(m, d, nullptr, nullptr, nullptr, nullptr, m->d_stats.p) : run_general_estep(m, d);
    if (!m->capturing) cudaEventRecord(m->evk1, s);
    m->ktimed = true;
    return r;
  }
  if (mode == JH_MODE_VITERBI) {
    // Viterbi training: the fast decoders (lane-per-sequence / checkpointed long path) write byte paths, k_viterbi_path_counts
    // adds every sequence's weight along its path; models outside their scope keep the reference-order kernel with counts
    if (m->opt.vit_train_fast && m->opt.fast && m->fast.usable && m->S <= 255 && !groups_active(m, d)) {
      int r;
      if ((r = m->d_vt_paths.reserve((size_t)d->n_res + 16)) || (r = m->d_vt_scores.reserve((size_t)d->n_seq))) return r;
      if (!m->capturing) cudaEventRecord(m->evk0, s);
      if ((r = run_viterbi(m, d, PathOut{m->d_vt_paths.p, 1}, nullptr, m->d_vt_scores.p, nullptr))) return r;
      const int n_stats = (int)m->n_stats();
      const size_t smem = (size_t)n_stats * sizeof(double);
      const int in_smem = smem <= 160 * 1024;
      auto kern = jhv::k_viterbi_path_counts;
      if (in_smem && smem > 48 * 1024) JH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // latency bound (shared-memory atomics, path bytes): as many warps per SM as the statistics copies leave room for
      const int64_t per_sm = in_smem ? std::max<int64_t>(1, std::min<int64_t>(6, (int64_t)(200 * 1024) / std::max<size_t>(smem, 1))) : 6;
      const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((d->n_seq + 7) / 8, (int64_t)m->sm_count * per_sm));
      kern<<<(int)grid, 256, in_smem ? smem : 0, s>>>(m->fast.dev(), d->dev(), m->d_vt_paths.p, m->d_vt_scores.p, m->n_cond * m->a, in_smem, m->d_stats.p);
      JH_LAUNCHED();
      if (!m->capturing) cudaEventRecord(m->evk1, s);
      m->ktimed = true;
      JH_CUDA(cudaGetLastError());
      return JH_OK;
    }
    if (!m->capturing) cudaEventRecord(m->evk0, s);
    int r = run_viterbi(m, d, PathOut{nullptr, 0}, nullptr, nullptr, m->d_stats.p);
    if (!m->capturing) cudaEventRecord(m->evk1, s);
    m->ktimed = true;
    return r;
  }
  const bool sparse = prefer_sparse(m, d);
  if (m->opt.deterministic && (sparse || groups_active(m, d) || !(m->opt.fast && m->fast.usable))) {
    jh_set_error("deterministic mode covers the octet kernels only (plain first-order models with 5..32 states, many sequences)");
    return JH_ERR_UNSUPPORTED;
  }
  if (!groups_active(m, d) && (sparse || (m->opt.fast && m->fast.usable))) {
    int r = m->d_fallback.reserve(d->n_seq + 2);
    if (r) return r;
    int32_t *flag = m->d_fallback.p + d->n_seq + 1;
    JH_CUDA(cudaMemsetAsync(m->d_fallback.p, 0, sizeof(int32_t), s));
    JH_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), s));
    if (!sparse && variant_of(m) == 0 && (r = ensure_octet_hashes(m, d))) return r;
    if (!m->capturing) cudaEventRecord(m->evk0, s);
    r = sparse ? run_estep_sparse(m, d) : run_estep_fast(m, d);
    if (!m->capturing) cudaEventRecord(m->evk1, s);
    m->ktimed = true;
    if (r) return r;
    int n = (int)m->n_stats() + 1;
    jhf::k_check_finite<<<(n + 255) / 256, 256, 0, s>>>(m->d_stats.p, n, flag);
    JH_LAUNCHED();
    // Sequences whose scaled recursion lost all mass are re-run by the reference-order kernels, and an E-step whose
    // statistics are not finite (dynamic range exceeded) is redone entirely in log space — both decided ON THE DEVICE
    // (kernels_exact.cuh: Cond), so E-step -> all-reduce -> M-step stays stream ordered.  The log-space kernels need
    // (L+1) x Cmax forward rows per resident group; if those do not fit 1 GB of scratch (chromosome-scale sequences) the
    // decision is taken on the host as before.
    const int64_t fb_budget = (int64_t)1 << 30;
    const int gpc = 256 / group_size(m);
    if ((d->max_len + 1) * (int64_t)m->Cmax * 8 * gpc <= fb_budget) {
      jha::k_zero_if<<<(n + 7 + 255) / 256, 256, 0, s>>>(flag, m->d_stats.p, n + 7);
      JH_LAUNCHED();
      jhx::Cond redo{flag, 1, nullptr}, listed{flag, 0, m->d_fallback.p};
      if ((r = run_fb_exact(m, d, JHX_MODE_COUNTS, PathOut{nullptr, 0}, nullptr, nullptr, nullptr, 0, redo, fb_budget))) return r;
      return run_fb_exact(m, d, JHX_MODE_COUNTS, PathOut{nullptr, 0}, nullptr, nullptr, m->d_fallback.p + 1, 0, listed, fb_budget);
    }
    int32_t n_fb = 0, bad = 0;
    JH_CUDA(cudaMemcpyAsync(&n_fb, m->d_fallback.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    JH_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    JH_CUDA(cudaStreamSynchronize(s));
    if (bad) {  // dynamic range exceeded: redo the whole E-step in log space
      JH_CUDA(cudaMemsetAsync(m->d_stats.p, 0, sizeof(double) * (m->n_stats() + 8), s));
      return run_fb_exact(m, d, JHX_MODE_COUNTS, PathOut{nullptr, 0}, nullptr, nullptr, nullptr, 0);
    }
    if (n_fb > 0) return run_fb_exact(m, d, JHX_MODE_COUNTS, PathOut{nullptr, 0}, nullptr, nullptr, m->d_fallback.p + 1, n_fb);
    return JH_OK;
  }
  if (!m->capturing) cudaEventRecord(m->evk0, s);
  int r = run_fb_exact(m, d, JHX_MODE_COUNTS, PathOut{nullptr, 0}, nullptr, nullptr, nullptr, 0);
  if (!m->capturing) cudaEventRecord(m->evk1, s);
  m->ktimed = true;
  return r;
}

static int allreduce_stats(jh_model *m) {
  if (!m->comm || m->n_ranks <= 1) return JH_OK;
  int rc = g_nccl.AllReduce(m->d_stats.p, m->d_stats.p, (size_t)m->n_stats() + 1, /*ncclDouble*/ 8, /*ncclSum*/ 0, m->comm, m->st());
  if (rc != 0) {
    jh_set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return JH_ERR_COMM;
  }
  return JH_OK;
}

// Every rank has to start an EM run from the same parameters (the M-step is replicated, not broadcast): rank 0's
// parameters go to all ranks — e.g. after a random initialisation drawn independently per process.
static int broadcast_params(jh_model *m) {
  if (!m->comm || m->n_ranks <= 1) return JH_OK;
  if (!g_nccl.Broadcast) { jh_set_error("libnccl lacks ncclBroadcast"); return JH_ERR_COMM; }
  cudaStream_t s = m->st();
  struct { double *p; size_t n; } bufs[4] = {{m->d_trans_param.p, (size_t)m->n_child}, {m->d_trans_lognorm.p, (size_t)m->n_elem},
                                             {m->d_emis_param.p, (size_t)m->n_cond * m->a}, {m->d_emis_lognorm.p, (size_t)m->n_cond}};
  for (auto &b : bufs) {
    if (b.n == 0) continue;
    int rc = g_nccl.Broadcast(b.p, b.p, b.n, /*ncclDouble*/ 8, 0, m->comm, s);
    if (rc != 0) { jh_set_error("ncclBroadcast failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return JH_ERR_COMM; }
  }
  return prep_scores(m);
}

static int mstep_device(jh_model *m) {
  jha::k_mstep<<<1, 256, 0, m->st()>>>(m->params(), m->d_stats.p, m->d_stats.p + m->n_child);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return prep_scores(m);
}

static int estep_collective(jh_model *m, jh_dataset *d, int mode);

extern "C" int jh_estep(jh_model *m, jh_dataset *d, int mode, double *trans_stat, double *emis_stat, double *score) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (mode != JH_MODE_VITERBI && mode != JH_MODE_BAUM_WELCH) { jh_set_error("Training mode not available."); return JH_ERR_INVALID; }
  timing_begin(m);
  r = estep_collective(m, d, mode);
  timing_end(m);
  if (!r && m->failed) {
    r = m->failed;
    m->failed = 0;
    jh_set_error("%s", m->failed_msg);
  }
  if (r) return r;
  cudaStream_t s = m->st();
  if (trans_stat) JH_CUDA(cudaMemcpyAsync(trans_stat, m->d_stats.p, sizeof(double) * m->n_child, cudaMemcpyDeviceToHost, s));
  if (emis_stat) JH_CUDA(cudaMemcpyAsync(emis_stat, m->d_stats.p + m->n_child, sizeof(double) * m->n_cond * m->a, cudaMemcpyDeviceToHost, s));
  if (score) JH_CUDA(cudaMemcpyAsync(score, m->d_stats.p + m->n_stats(), sizeof(double), cudaMemcpyDeviceToHost, s));
  if (trans_stat || emis_stat || score) JH_CUDA(cudaStreamSynchronize(s));
  return JH_OK;
}

extern "C" int jh_mstep(jh_model *m) {
  JH_ENTER(m);
  return mstep_device(m);
}

extern "C" int jh_stats_device(jh_model *m, void **dev_ptr, int64_t *n_doubles) {
  if (!m || !dev_ptr || !n_doubles) { jh_set_error("null argument"); return JH_ERR_INVALID; }
  *dev_ptr = m->d_stats.p;
  *n_doubles = m->n_stats() + 1;
  return JH_OK;
}

// E-step + all-reduce.  With a communicator attached a rank whose local E-step fails must not leave its peers waiting in
// the collective: it remembers the error (m->failed), poisons the score slot (NaN: every rank's objective becomes NaN, so the
// peers notice) and keeps taking part in the collectives of the remaining iterations; its error is returned at the end.
static int estep_collective(jh_model *m, jh_dataset *d, int mode) {
  int r = m->failed ? m->failed : estep_device(m, d, mode);
  if (r) {
    if (!m->comm || m->n_ranks <= 1) return r;
    if (!m->failed) snprintf(m->failed_msg, sizeof(m->failed_msg), "%s", g_err);
    m->failed = r;
    cudaGetLastError();
    cudaMemsetAsync(m->d_stats.p + m->n_stats(), 0xff, sizeof(double), m->st());
  }
  return allreduce_stats(m);
}

// E-part of one EM iteration: E-step, count all-reduce, objective = getLogPriorTerm() + oneIteration() (HigherOrderHMM.java:762)
// stored at objective[iteration % cap] by the device-side iteration counter (no host value enters: graph replayable).
static int em_e_part(jh_model *m, jh_dataset *d, int mode, int obj_cap) {
  cudaStream_t s = m->st();
  int r;
  if ((r = estep_collective(m, d, mode))) return r;
  jha::k_log_prior<<<1, jha::LOG_PRIOR_THREADS, 0, s>>>(m->params(), m->d_stats.p + m->n_stats(), m->d_objective.p + 4095);
  JH_LAUNCHED();
  jha::k_store_objective<<<1, 1, 0, s>>>(m->d_objective.p + 4095, m->d_objective.p, m->d_iter.p, obj_cap);
  JH_LAUNCHED();
  JH_CUDA(cudaGetLastError());
  return JH_OK;
}

extern "C" int jh_baum_welch(jh_model *m, jh_dataset *d, const jh_train_opts *o, double *objective, int32_t cap, int32_t *n_iter) {
  JH_ENTER2(m, d);
  int r = JH_OK;
  if (!o || !n_iter) { jh_set_error("null argument"); return JH_ERR_INVALID; }
  if (o->term_kind != JH_TERM_ITERATIONS && o->term_kind != JH_TERM_SMALL_DIFFERENCE) { jh_set_error("unknown termination condition"); return JH_ERR_INVALID; }
  cudaStream_t s = m->st();
  const int obj_cap = 4000;
  double old_value, new_value = JH_NEG_INF;
  int it = 0;
  JH_CUDA(cudaMemsetAsync(m->d_iter.p, 0, sizeof(int32_t) * 4, s));
  if ((r = broadcast_params(m))) return r;
  // history of the iterations [base, it) -> host (ring of obj_cap values on the device)
  auto flush = [&](int upto) -> int {
    const int base = ((upto - 1) / obj_cap) * obj_cap, cnt = upto - base;
    if (objective && base < cap)
      JH_CUDA(cudaMemcpyAsync(objective + base, m->d_objective.p, sizeof(double) * std::min(cnt, (int)cap - base), cudaMemcpyDeviceToHost, s));
    JH_CUDA(cudaStreamSynchronize(s));
    return JH_OK;
  };
  timing_begin(m);
  if (o->term_kind == JH_TERM_ITERATIONS) {
    // IterationCondition.java:83-86: N E-steps, N-1 M-steps, no host round trip.  From the second iteration on the
    // (E-part + M-step) pair is replayed as a CUDA graph: cfg1-sized problems are bound by the ~15 launches per iteration.
    const int N = std::max(1, o->max_iter);
    if (m->opt.graph && N >= 6) {
      if ((r = em_e_part(m, d, o->mode, obj_cap))) return r;
      it = 1;
      if ((r = mstep_device(m))) return r;
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        m->capturing = true;
        const int r1 = em_e_part(m, d, o->mode, obj_cap), r2 = r1 ? r1 : mstep_device(m);
        m->capturing = false;
        ok = cudaStreamEndCapture(s, &graph) == cudaSuccess && !r1 && !r2 && graph;
        if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
      }
      cudaGetLastError();
      if (ok) {
        for (; it < N - 1; it++) {
          if (cudaGraphLaunch(exec, s) != cudaSuccess) { ok = false; break; }
          if ((it + 1) % obj_cap == 0 && (r = flush(it + 1))) break;
        }
      }
      if (exec) cudaGraphExecDestroy(exec);
      if (graph) cudaGraphDestroy(graph);
      if (r) return r;
      if (!ok && it > 1) { jh_set_error("CUDA graph replay of the EM iteration failed: %s", cudaGetErrorString(cudaGetLastError())); return JH_ERR_CUDA; }
    }
    for (;;) {  // eager iterations (all of them without the graph; the last E-part with it)
      if ((r = em_e_part(m, d, o->mode, obj_cap))) return r;
      it++;
      const bool go_on = it < N;
      if (it % obj_cap == 0 || !go_on)
        if ((r = flush(it))) return r;
      if (!go_on) break;
      if ((r = mstep_device(m))) return r;
    }
  } else {
    for (;;) {
      old_value = new_value;
      if ((r = em_e_part(m, d, o->mode, obj_cap))) return r;
      JH_CUDA(cudaMemcpyAsync(&new_value, m->d_objective.p + (it % obj_cap), sizeof(double), cudaMemcpyDeviceToHost, s));
      JH_CUDA(cudaStreamSynchronize(s));
      if (objective && it < cap) objective[it] = new_value;
      it++;
      const bool go_on = o->eps <= fabs(old_value - new_value) && (o->max_iter <= 0 || it < o->max_iter);  // SmallDifference...java:86-89
      if (!go_on) break;
      if ((r = mstep_device(m))) return r;
    }
  }
  timing_end(m);
  JH_CUDA(cudaMemcpyAsync(&m->last_objective, m->d_objective.p + ((it - 1) % obj_cap), sizeof(double), cudaMemcpyDeviceToHost, s));
  JH_CUDA(cudaStreamSynchronize(s));
  *n_iter = it;
  if (m->failed) {  // this rank's E-step failed; it only kept its peers' collectives company
    r = m->failed;
    m->failed = 0;
    jh_set_error("%s", m->failed_msg);
    return r;
  }
  if (m->comm && m->n_ranks > 1 && m->last_objective != m->last_objective) {
    jh_set_error("the E-step of a peer rank failed (the all-reduced objective is NaN)");
    return JH_ERR_COMM;
  }
  return JH_OK;
}

// Objective of the LAST iteration of the most recent jh_baum_welch on this model (the history buffer of that call may be
// shorter than the run).
extern "C" int jh_last_objective(jh_model *m, double *out) {
  if (!m || !out) { jh_set_error("null argument"); return JH_ERR_INVALID; }
  *out = m->last_objective;
  return JH_OK;
}

// ---- communicator ---------------------------------------------------------------------------------
extern "C" int jh_comm_unique_id(uint8_t id[JH_COMM_ID_BYTES]) {
  int r = nccl_load();
  if (r) return r;
  Id128 u;
  memset(&u, 0, sizeof(u));
  int rc = g_nccl.GetUniqueId(&u);
  if (rc != 0) { jh_set_error("ncclGetUniqueId failed (%d)", rc); return JH_ERR_COMM; }
  memcpy(id, &u, JH_COMM_ID_BYTES);
  return JH_OK;
}

extern "C" int jh_comm_init(jh_model *m, const uint8_t id[JH_COMM_ID_BYTES], int32_t n_ranks, int32_t rank) {
  if (!m || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) { jh_set_error("invalid communicator arguments"); return JH_ERR_INVALID; }
  int r = nccl_load();
  if (r) return r;
  JH_CUDA(cudaSetDevice(m->device));
  Id128 u;
  memcpy(&u, id, JH_COMM_ID_BYTES);
  void *comm = nullptr;
  int rc = g_nccl.CommInitRank(&comm, n_ranks, u, rank);
  if (rc != 0) { jh_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return JH_ERR_COMM; }
  m->comm = comm; m->n_ranks = n_ranks; m->rank = rank;
  return JH_OK;
}

extern "C" int jh_comm_broadcast_params(jh_model *m) {
  JH_ENTER(m);
  int r = broadcast_params(m);
  if (r) return r;
  JH_CUDA(cudaStreamSynchronize(m->st()));
  return JH_OK;
}

extern "C" int jh_comm_destroy(jh_model *m) {
  if (!m) { jh_set_error("null model"); return JH_ERR_INVALID; }
  if (m->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
  m->comm = nullptr; m->n_ranks = 1; m->rank = 0;
  return JH_OK;
}

// ---- measurement ---------------------------------------------------------------------------------
extern "C" int jh_measure_fp64_peak(double *tflops) {
  const int device = g_default_device;
  int r = ensure_device(device);
  if (r) return r;
  cudaDeviceProp prop;
  JH_CUDA(cudaGetDeviceProperties(&prop, device));
  double *d_out;
  JH_CUDA(cudaMalloc(&d_out, 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 16, threads = 512, blocks = prop.multiProcessorCount * 4;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    jha::k_fp64_peak<<<blocks, threads>>>(d_out, iters, 1.0 + rep);
    JH_LAUNCHED();
    cudaEventRecord(e1);
    JH_CUDA(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops = 2.0 * 8.0 * iters * (double)threads * blocks / (best * 1e-3) / 1e12;
  return JH_OK;
}

// Probe (profiles/): DFMA rate in TFLOP/s with `chains` independent FMA chains per thread, `threads` per block and
// `blocks_per_sm` resident blocks — used to size the ILP the E-step kernel needs at 8 warps per SM.
extern "C" int jh_probe_fp64(int chains, int threads, int blocks_per_sm, double *tflops) {
  const int device = g_default_device;
  int r = ensure_device(device);
  if (r) return r;
  cudaDeviceProp prop;
  JH_CUDA(cudaGetDeviceProperties(&prop, device));
  double *d_out;
  JH_CUDA(cudaMalloc(&d_out, 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 15, blocks = prop.multiProcessorCount * blocks_per_sm;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    switch (chains) {
      case 1: jha::k_fp64_chains<1><<<blocks, threads>>>(d_out, iters, 1.0 + rep); break;
      case 2: jha::k_fp64_chains<2><<<blocks, threads>>>(d_out, iters, 1.0 + rep); break;
      case 4: jha::k_fp64_chains<4><<<blocks, threads>>>(d_out, iters, 1.0 + rep); break;
      case 8: jha::k_fp64_chains<8><<<blocks, threads>>>(d_out, iters, 1.0 + rep); break;
      default: jha::k_fp64_chains<16><<<blocks, threads>>>(d_out, iters, 1.0 + rep); chains = 16; break;
    }
    JH_LAUNCHED();
    cudaEventRecord(e1);
    JH_CUDA(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops = 2.0 * chains * iters * (double)threads * blocks / (best * 1e-3) / 1e12;
  return JH_OK;
}

// Probe (profiles/): FP64 tensor-core (DMMA m8n8k4) rate in TFLOP/s — a measurement of the hardware ceiling only.
extern "C" int jh_probe_dmma(int chains, int threads, int blocks_per_sm, double *tflops) {
  const int device = g_default_device;
  int r = ensure_device(device);
  if (r) return r;
  cudaDeviceProp prop;
  JH_CUDA(cudaGetDeviceProperties(&prop, device));
  double *d_out;
  JH_CUDA(cudaMalloc(&d_out, 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 14, blocks = prop.multiProcessorCount * blocks_per_sm;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    switch (chains) {
      case 1: jha::k_dmma_chains<1><<<blocks, threads>>>(d_out, iters, 1.0 + rep); break;
      case 2: jha::k_dmma_chains<2><<<blocks, threads>>>(d_out, iters, 1.0 + rep); break;
      case 4: jha::k_dmma_chains<4><<<blocks, threads>>>(d_out, iters, 1.0 + rep); break;
      case 8: jha::k_dmma_chains<8><<<blocks, threads>>>(d_out, iters, 1.0 + rep); break;
      default: jha::k_dmma_chains<16><<<blocks, threads>>>(d_out, iters, 1.0 + rep); chains = 16; break;
    }
    JH_LAUNCHED();
    cudaEventRecord(e1);
    JH_CUDA(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops = 2.0 * 256.0 * chains * iters * (double)(threads / 32) * blocks / (best * 1e-3) / 1e12;
  return JH_OK;
}
