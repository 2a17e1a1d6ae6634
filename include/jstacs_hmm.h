/*
 * jstacs_hmm.h — C ABI of the B200-native HMM engine (libjstacs_hmm_b200.so).
 *
 * This is the drop-in boundary for the Jstacs HMM hot path.  Every entry point
 * names the Jstacs (reference) method it replaces; paths are relative to the
 * reference root and abbreviated
 *   HMM/ = de/jstacs/sequenceScores/statisticalModels/trainable/hmm/
 *
 * Conventions
 *   - plain C, no C++/torch types; all buffers are caller-owned HOST memory
 *     unless the name says `_device`;
 *   - every function returns 0 on success, a negative jh_status otherwise;
 *     jh_last_error() returns a thread-local message for the last failure
 *     (the Java shim turns it into the exception type the reference throws);
 *   - a handle is used by one caller thread at a time (the reference model
 *     instance is not thread-safe either, HMM/models/HigherOrderHMM.java:1189);
 *   - there is NO CPU fallback: without a usable CUDA device every compute
 *     entry point fails with JH_ERR_CUDA.
 */
#ifndef JSTACS_HMM_H
#define JSTACS_HMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JH_ABI_VERSION 2

typedef enum jh_status {
  JH_OK = 0,
  JH_ERR_INVALID = -1,      /* IllegalArgumentException in the reference            */
  JH_ERR_UNSUPPORTED = -2,  /* feature outside the native path (filters, reverse
                               strand states, position dependent elements, ...)     */
  JH_ERR_CUDA = -3,         /* no device / CUDA runtime failure                      */
  JH_ERR_ALPHABET = -4,     /* WrongAlphabetException (AbstractHMM.java:991)         */
  JH_ERR_NOMEM = -5,
  JH_ERR_IMPOSSIBLE = -6,   /* "Impossible path": no state path has prob > 0         */
  JH_ERR_COMM = -7          /* NCCL failure                                          */
} jh_status;

/*
 * Flattened model ("context graph" IR).
 *
 * It is exactly what the public Transition interface exposes
 * (HMM/transitions/Transition.java:63-199) plus the emission tables
 * (HMM/states/emissions/discrete/AbstractConditionalDiscreteEmission.java:100-135):
 *
 *   layer class lc = min(layer, max_order)         BasicHigherOrderTransition.java:558-560
 *   contexts of lc, in lookup[0][lc] order          BasicHigherOrderTransition.java:259-266
 *   element of a context  = lookup[0][lc][ctx]      :571-573
 *   child c of an element = {state, descendant}     :547-555 (fillTransitionInformation)
 *   next context          = lookup[1][min(layer+adv,m)][descendant], adv = silent?0:1
 *   log transition score  = parameters[c] - logNorm :1116-1118
 *   log emission score    = (0 - logNorm[cond]) + params[cond][sym], or -log(a)
 *                           if the condition is not available (pos < order)
 *                           AbstractConditionalDiscreteEmission.java:431-457
 *   cond                  = sum_i a^i * x[pos-1-i]  HigherOrderDiscreteEmission.java:138-146
 */
typedef struct jh_model_desc {
  int32_t abi_version;          /* JH_ABI_VERSION */
  int32_t n_states;             /* S */
  int32_t alphabet;             /* a (4 for DNA, DNAAlphabet.java:72) */
  int32_t max_order;            /* m = Transition.getMaximalMarkovOrder() */

  /* contexts, grouped by layer class 0..m */
  const int32_t *lc_ctx_ptr;    /* [m+2]  contexts of class lc are lc_ctx_ptr[lc] .. lc_ctx_ptr[lc+1]-1 */
  const int32_t *ctx_elem;      /* [n_ctx_total] transition element of each context */
  const int32_t *ctx_last_state;/* [n_ctx_total] Transition.getLastContextState, -1 for the start context */

  /* transition elements */
  int32_t n_elem;
  const int32_t *elem_child_ptr;/* [n_elem+1] */
  const int32_t *child_state;   /* [n_child_total] */
  const int32_t *child_desc;    /* [n_child_total] descendant element index */
  const int32_t *elem_ctx;      /* [(m+1)*n_elem] lookup[1][lc][elem]: local context index in class lc, or -1 */
  const int32_t *trans_index;   /* [n_elem] parameter tying, transIndex[t] <= t (BasicHigherOrderTransition.java:81-101) */
  const double *trans_param;    /* [n_child_total] AbstractTransitionElement.parameters */
  const double *trans_lognorm;  /* [n_elem]        AbstractTransitionElement.logNorm    */
  const double *trans_hyper;    /* [n_child_total] AbstractTransitionElement.hyperParameters */

  /* emissions */
  int32_t n_emis;
  const int32_t *state_emis;    /* [S] emissionIdx (AbstractHMM.java:273-283) */
  const int32_t *emis_order;    /* [n_emis] Markov order of the emission, -1 = SilentEmission */
  const int32_t *emis_cond_ptr; /* [n_emis+1] first condition row of each emission */
  const double *emis_param;     /* [n_cond_total*a] params[cond][sym]  */
  const double *emis_lognorm;   /* [n_cond_total]   logNorm[cond]      */
  const double *emis_hyper;     /* [n_cond_total*a] hyperParams        */

  const uint8_t *state_silent;  /* [S] State.isSilent (SimpleState.java:71-82) */
  const uint8_t *state_final;   /* [S] AbstractHMM.finalState (AbstractHMM.java:1101-1113) */
} jh_model_desc;

typedef struct jh_model jh_model;
typedef struct jh_dataset jh_dataset;

/* ---- library ----------------------------------------------------------- */
int jh_abi_version(void);
/* Thread-local message of the last failing call (never NULL). */
const char *jh_last_error(void);
/* Number of visible CUDA devices; <0 on failure. */
int jh_device_count(void);
/* Select the device used by handles created afterwards (default 0). */
int jh_set_device(int device);
/* Counts kernels launched by this library since load (bench.py: gpu_launches). */
int64_t jh_kernel_launches(void);
/* Default CUDA stream (cudaStream_t as void*) of the handles on the current default device that have no stream of their
 * own (jh_model_set_stream); NULL = every model works on its private stream.  The legacy default stream is addressed as
 * cudaStreamLegacy ((void*)0x1). */
int jh_set_stream(void *cuda_stream);
int jh_synchronize(void);
/* Tuning / test switches: "fast" (1: scaled linear-space kernels where the model allows, 0: reference-order
 * log-space kernels only), "scratch_mb" (DP scratch budget), "blocks_per_sm", "fast_variant", "chunk" (positions per
 * checkpoint of the long-sequence path), "oct_chunk" / "oct_chunk_force" (checkpointed forward rows of the octet kernels),
 * "pass1_cta", "overlap", "scan", "lane_chunks", "split_short" (mappings of the long-sequence path: CTA per sequence,
 * per-sequence streams, checkpoint pass parallel in the position, one chunk per lane, short sequences of tiny models cut
 * into chunks; all on by default, switched off by the tests to reach the other kernels), "trim_pool" (return unused
 * device-pool memory to the driver, keeping `value` bytes: device buffers are pooled so that per-call data sets do not
 * pay cudaMalloc/cudaFree). */
int jh_set_option(const char *name, int64_t value);
/* jh_set_option changes the DEFAULTS, which every handle follows unless it has its own value: device, stream and options
 * live in the handle, so models on different devices or threads with different options do not interfere.
 * Further switches: "graph" (replay the EM iteration as a CUDA graph, default 1), "stage" (pageable host memory is moved
 * through the library's pinned double buffers, default 1), "vit_chunk" / "vit_chunk_force" (checkpointed Viterbi),
 * "band" (warp-resident passes of banded models, default 1), "band_meet" (scoring of long sequences of banded models with
 * a forward and a backward chain that meet in the middle, default 1), "oct_chunk" < 0 = automatic chunk length (default),
 * "oct_full_rows_mb", "deterministic" (default 0; 1: Baum-Welch E-step in a fixed order — static work assignment,
 * fixed-point emission counts, partial sums folded in slot order — so that statistics, objectives and trained parameters
 * are bit-identical from run to run; covers the octet kernels, i.e. plain first-order models with 5..32 states on many
 * sequences, anything else returns JH_ERR_UNSUPPORTED while the option is set). */

/* ---- model ------------------------------------------------------------- */
/* Replaces the object graph built by HigherOrderHMM(...) / HMMFactory.createErgodicHMM
 * (HMM/models/HigherOrderHMM.java:156-187, HMM/HMMFactory.java:156-181): the shim reads the
 * model through the public Transition/Emission interfaces and hands the flattened copy over. */
int jh_model_create(const jh_model_desc *desc, jh_model **out);
void jh_model_destroy(jh_model *m);
/* DifferentiableHigherOrderHMM.setParameters(double[],int) (:339-344) — all arrays in jh_model_desc layout.
 * NULL lognorm pointers mean "recompute like precompute()" (TransitionElement.java:119-126,
 * AbstractConditionalDiscreteEmission.java:466-480). */
int jh_model_set_params(jh_model *m, const double *trans_param, const double *trans_lognorm,
                        const double *emis_param, const double *emis_lognorm);
/* getCurrentParameterValues (:316-329) + the logNorm fields. Any pointer may be NULL. */
int jh_model_get_params(jh_model *m, double *trans_param, double *trans_lognorm,
                        double *emis_param, double *emis_lognorm);
/* Per-handle overrides of jh_set_option / jh_set_stream (a handle is bound to the device that was the default when it
 * was created, jh_set_device). */
int jh_model_set_option(jh_model *m, const char *name, int64_t value);
int jh_model_set_stream(jh_model *m, void *cuda_stream);
/* sizes: n_child_total, n_elem, n_cond_total*a, n_cond_total */
int jh_model_sizes(const jh_model *m, int64_t sizes[4]);
/* HigherOrderHMM.getLogPriorTerm (:253-259). */
int jh_model_log_prior(jh_model *m, double *out);

/* ---- data -------------------------------------------------------------- */
/* Replaces per-symbol DataSet.getElementAt(n).discreteVal(p) (DataSet.java:1001-1012,
 * Sequence.java:102): sequences are packed ONCE into uint8 codes with int64 offsets
 * (offsets[n]..offsets[n+1]) and copied to the device; weights may be NULL (all 1,
 * HigherOrderHMM.java:808-810). Sequences are length-bucketed internally; all outputs
 * are in caller order. */
int jh_dataset_create(const uint8_t *codes, const int64_t *offsets, int64_t n_seq,
                      const double *weights, int32_t alphabet, jh_dataset **out);
/* The same from 2-bit packed residues (alphabets of at most 4 symbols): four residues per byte, lowest bits first, every
 * sequence starting on a byte boundary at packed[byte_ptr[n]] — the on-disk / in-memory form of genome-scale DNA, cf. the
 * 32-symbols-per-long storage of de/jstacs/data/sequences/SparseSequence.java:59-73.  A quarter of the bytes cross PCIe;
 * the device widens them to the uint8 codes the kernels read. */
int jh_dataset_create_packed2(const uint8_t *packed, const int64_t *byte_ptr, const int64_t *offsets, int64_t n_seq,
                              const double *weights, int32_t alphabet, jh_dataset **out);
/* Streamed construction for genome-scale inputs: the residues arrive shard by shard (e.g. memory-mapped files written by
 * a streaming FASTA reader) and go straight to the device through the pinned staging buffers — the host never holds the
 * whole data set.  create announces the totals; every append delivers the next `n_seq` whole sequences (lengths[n_seq])
 * as uint8 codes (packed2 = 0, sum(lengths) bytes) or 2-bit packed, every sequence on a byte boundary (packed2 = 1,
 * sum((lengths+3)/4) bytes); finish (which consumes the builder) yields the same handle jh_dataset_create would. */
typedef struct jh_dataset_builder jh_dataset_builder;
int jh_dataset_builder_create(int64_t n_seq, int64_t n_residues, int32_t alphabet, jh_dataset_builder **out);
int jh_dataset_builder_append(jh_dataset_builder *b, const uint8_t *codes, const int64_t *lengths, int64_t n_seq, int32_t packed2);
int jh_dataset_builder_finish(jh_dataset_builder *b, const double *weights, jh_dataset **out);
void jh_dataset_builder_destroy(jh_dataset_builder *b);
void jh_dataset_destroy(jh_dataset *d);
/* Allowed states groups — semi-supervised training and constrained scoring / decoding (HigherOrderHMM.java:331-340, 795-835;
 * AbstractHMM.java:146-228, AbstractHMM.AllowedStatesGroups :1198-1246).  The model declares groups of states
 * (jh_model_set_states_groups = the constructor argument int[][] statesGroups); a data set carries, per position, the index
 * of the group the state emitting that position has to belong to, or -1 (= the int[] of the sequence annotation
 * AllowedStatesGroups.groups).  While both are set, jh_loglik (getLogProbFor(seq, start, end, statesGroups)), jh_viterbi*
 * (getViterbiPathFor(..., allowedStatesGroup)), jh_estep and jh_baum_welch (doOneStep, :812-835) evaluate the constraint,
 * on the reference-order kernels.  Posterior matrices ignore it, as in the reference.  Models without silent states and with
 * transition order >= 1; groups == NULL detaches. */
int jh_model_set_states_groups(jh_model *m, int32_t n_groups, const int32_t *group_ptr, const int32_t *group_states);
int jh_dataset_set_groups(jh_dataset *d, const int32_t *groups);
/* Sequence weights of train(DataSet,double[]) (TrainableStatisticalModel.java:61-88); NULL = all 1. */
int jh_dataset_set_weights(jh_dataset *d, const double *weights);
int jh_dataset_info(const jh_dataset *d, int64_t *n_seq, int64_t *n_residues, int64_t *max_len);

/* ---- compute entry points ---------------------------------------------- */
/* HigherOrderHMM.getLogScoreFor(DataSet,double[]) (:927-941) == AbstractHMM.logProb per sequence
 * (AbstractHMM.java:1049-1070): backward recursion, result bwd[0][0]. out[n_seq]. */
int jh_loglik(jh_model *m, jh_dataset *d, double *out);

/* AbstractHMM.getLogProbFor(seq, start, end) (AbstractHMM.java:985-998) for MANY sub-sequences at once, e.g. the sliding
 * windows of projects/xanthogenomes/NHMMer.java:225-244: window i scores positions start[i]..end[i] (inclusive) of
 * sequence seq_index[i]; layer 0 is the window start, but higher-order emissions see the residues before it, exactly as
 * HigherOrderDiscreteEmission.getConditionIndex does with absolute positions (:138-146). Works for every model the
 * library accepts (silent states, transition order > 1). out[n_win]. */
int jh_loglik_windows(jh_model *m, jh_dataset *d, int64_t n_win, const int64_t *seq_index, const int64_t *start,
                      const int64_t *end, double *out);

/* AbstractHMM.getViterbiPathsFor(DataSet) (AbstractHMM.java:894-900) ==
 * HigherOrderHMM.viterbi (:617-708) per sequence.  paths: int32, CSR by the dataset offsets
 * (models without silent states emit exactly one state per residue); scores[n_seq].
 * A sequence without any state path of probability > 0 makes the reference throw "Impossible path" (:659): the call then
 * returns JH_ERR_IMPOSSIBLE after filling the outputs (score -inf, states -1 / 255 for that sequence). */
int jh_viterbi(jh_model *m, jh_dataset *d, int32_t *paths, double *scores);
/* The same with ONE BYTE per residue (models with at most 255 states; 255 = "no state"): a quarter of the device->host
 * traffic, which dominates the call (cfg2: 1e9 residues = 4 GB of int32).  The shim widens to IntList on the host. */
int jh_viterbi_u8(jh_model *m, jh_dataset *d, uint8_t *paths, double *scores);
/* The same for ANY model, in particular with silent states (profile HMMs: HMMFactory.createProfileHMM,
 * parseProfileHMMFromHMMer), whose paths contain the silent states and therefore have variable length
 * (HigherOrderHMM.java:617-708 including the trailing silent states :667-705): sequence n writes at most
 * path_ptr[n+1]-path_ptr[n] states at paths[path_ptr[n]] (rest -1); path_len[n] = states of the path, -1 if no
 * child could be chosen (the reference fails there). A safe capacity is L*(1+n_silent)+n_silent. */
int jh_viterbi_paths(jh_model *m, jh_dataset *d, int32_t *paths, const int64_t *path_ptr, int64_t *path_len, double *scores);

/* AbstractHMM.getViterbiPathFor(startPos, endPos, seq) (HigherOrderHMM.java:590-598) for MANY sub-sequences at once: window i
 * decodes positions start[i]..end[i] (inclusive) of sequence seq_index[i]; like jh_loglik_windows the emissions keep their
 * absolute positions.  paths / path_ptr / path_len / scores as in jh_viterbi_paths, indexed by the window. */
int jh_viterbi_windows(jh_model *m, jh_dataset *d, int64_t n_win, const int64_t *seq_index, const int64_t *start, const int64_t *end,
                       int32_t *paths, const int64_t *path_ptr, int64_t *path_len, double *scores);

/* AbstractHMM.getLogStatePosteriorMatrixFor(start,end,seq) (AbstractHMM.java:776-797):
 * out[S][L] row-major for sequence seq_index. */
int jh_posterior(jh_model *m, jh_dataset *d, int64_t seq_index, double *out);
/* The same for positions start..end (inclusive) of the sequence: out[S][end-start+1]. */
int jh_posterior_window(jh_model *m, jh_dataset *d, int64_t seq_index, int64_t start, int64_t end, double *out);
/* AbstractHMM.decodeStatePosterior (:1125-1139) fused with the posterior: paths CSR like
 * jh_viterbi, loglik[n_seq] (may be NULL). */
int jh_posterior_decode(jh_model *m, jh_dataset *d, int32_t *paths, double *loglik);
/* One byte per residue, as jh_viterbi_u8. */
int jh_posterior_decode_u8(jh_model *m, jh_dataset *d, uint8_t *paths, double *loglik);

/* One E-step == Compute.oneIteration (HigherOrderHMM.java:1250-1261): resets the statistics,
 * runs baumWelch (mode 1, :723-726) or viterbi(null,...) (mode 0, :617-708) over every sequence,
 * returns sum_n w_n*score_n in *score. Statistics stay on the device; the optional host
 * pointers receive copies (trans_stat[n_child_total] per element/child with tying applied,
 * emis_stat[n_cond_total*a]). */
#define JH_MODE_VITERBI 0
#define JH_MODE_BAUM_WELCH 1
int jh_estep(jh_model *m, jh_dataset *d, int mode, double *trans_stat, double *emis_stat, double *score);
/* Compute.estimateFromStatistics (:1263-1279) == M-step on the device from the statistics
 * of the last jh_estep (BasicHigherOrderTransition.java:1242-1262,
 * AbstractConditionalDiscreteEmission.java:682-699). */
int jh_mstep(jh_model *m);

/* Termination conditions (de/jstacs/algorithms/optimization/termination/):
 * IterationCondition.java:83-86, SmallDifferenceOfFunctionEvaluationsCondition.java:86-89. */
#define JH_TERM_ITERATIONS 0
#define JH_TERM_SMALL_DIFFERENCE 1
typedef struct jh_train_opts {
  int32_t mode;        /* JH_MODE_* */
  int32_t term_kind;   /* JH_TERM_* */
  int32_t max_iter;    /* IterationCondition(maxIter); also a hard cap for SMALL_DIFFERENCE (<=0: none) */
  double eps;          /* SmallDifferenceOfFunctionEvaluationsCondition(eps) */
} jh_train_opts;
/* HigherOrderHMM.train inner do-while (:760-770) for one start with skipInit: N E-steps and
 * N-1 M-steps for IterationCondition(N). objective[cap] receives prior + sum_n w_n*logP_n per
 * iteration; *n_iter the number of E-steps run. When a communicator is attached
 * (jh_comm_*) the statistics and the score are all-reduced over the ranks each iteration. */
int jh_baum_welch(jh_model *m, jh_dataset *d, const jh_train_opts *opts,
                  double *objective, int32_t cap, int32_t *n_iter);
/* Objective of the LAST iteration of the most recent jh_baum_welch of this model (what HigherOrderHMM.train compares
 * between starts, :773-779), independent of the capacity of the history buffer. */
int jh_last_objective(jh_model *m, double *out);

/* ---- multi-GPU: one process per GPU, shards of the DataSet per rank ------ */
/* The reference's joinStatistics + setParameters (HigherOrderHMM.java:1263-1288) is a sum over
 * workers followed by a broadcast; here it is ONE ncclAllReduce(sum, fp64) of
 * [transition statistics | emission statistics | score] per EM iteration, the M-step is then
 * run redundantly on every rank. id is the 128 byte ncclUniqueId made by rank 0. */
#define JH_COMM_ID_BYTES 128
int jh_comm_unique_id(uint8_t id[JH_COMM_ID_BYTES]);
int jh_comm_init(jh_model *m, const uint8_t id[JH_COMM_ID_BYTES], int32_t n_ranks, int32_t rank);
/* Rank 0's parameters -> every rank (ncclBroadcast).  jh_baum_welch does this itself at its start when a communicator is
 * attached, so that ranks whose host side initialised randomly (HigherOrderHMM.initializeRandomly, :880-885, draws from an
 * unseeded generator per process) still run the replicated M-step on identical parameters. */
int jh_comm_broadcast_params(jh_model *m);
int jh_comm_destroy(jh_model *m);
/* Device pointer and length (doubles) of the contiguous [trans|emis|score] statistics block,
 * for callers that run their own collective (e.g. torch.distributed) between jh_estep and jh_mstep. */
int jh_stats_device(jh_model *m, void **dev_ptr, int64_t *n_doubles);

/* ---- measurement helpers (bench.py) ------------------------------------- */
/* Sustained FP64 FMA rate of this device in TFLOP/s (2 flops per DFMA), a register-resident DFMA chain. */
int jh_measure_fp64_peak(double *tflops);
/* DFMA rate with `chains` independent FMA chains per thread at a given residency (latency/ILP probe, profiles/). */
int jh_probe_fp64(int chains, int threads, int blocks_per_sm, double *tflops);
/* FP64 tensor-core (DMMA m8n8k4) rate, same arguments; a hardware-ceiling measurement, not used by the product kernels. */
int jh_probe_dmma(int chains, int threads, int blocks_per_sm, double *tflops);
/* Device time in ms of the kernels of the last compute call (CUDA events on the library stream). */
int jh_last_kernel_ms(jh_model *m, double *ms);
/* Device time in ms of the E-step kernel launches alone (the dominant kernel) of the last E-step. */
int jh_last_main_kernel_ms(jh_model *m, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* JSTACS_HMM_H */
