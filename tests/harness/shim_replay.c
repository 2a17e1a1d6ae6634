/*
 * shim_replay.c — replays, in plain C against the C ABI, EXACTLY the call sequences the Java drop-in issues
 * (java/de/jstacs/.../models/GpuHmmEngine.java, class comment): what a Jstacs user gets through GpuHigherOrderHMM /
 * GpuBasicHigherOrderHMM is therefore the route the GPU parity tests exercise (tests/test_shim_replay.py compares the
 * outputs with the oracle).  The image has no JDK; this program is the executable stand-in of the shim.
 *
 *   first use          : jh_model_create(desc)
 *   new DataSet        : jh_dataset_destroy(old) ; jh_dataset_create(codes, offsets, n, weights, alphabet)
 *   same DataSet       : jh_dataset_set_weights(weights | NULL)
 *   every method       : jh_model_set_params(tp, tl, ep, el)          <- the "Java-side" parameters kept by this program
 *   train              : per start { [initialise ; set_params] ; jh_baum_welch ; jh_last_objective ; jh_model_get_params ; write back }
 *   score              : jh_loglik
 *   viterbi            : jh_viterbi_u8 (no silent states, <= 255 states) | jh_viterbi_paths (capacity L*(1+nSilent)+nSilent)
 *   decode             : jh_posterior_decode_u8 | jh_posterior_decode
 *   posterior          : jh_posterior per sequence
 *   windows            : jh_loglik_windows
 *   end                : jh_dataset_destroy ; jh_model_destroy
 *
 * usage: shim_replay <libjstacs_hmm_b200.so> <in.bin> <out.bin> op [op ...]
 *   ops: score | viterbi | decode | posterior | windows | train:<mode>:<term_kind>:<max_iter>:<eps>:<starts> | newdata
 * in.bin / out.bin: records  [u32 name_len][name][u8 dtype 0=i32 1=i64 2=f64 3=u8][u64 count][payload]
 * Build: gcc -O2 -o shim_replay shim_replay.c -ldl   (the library is loaded with dlopen, like SymbolLookup.libraryLookup)
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/jstacs_hmm.h"

typedef struct { char name[64]; int dtype; uint64_t count; void *data; } rec_t;
static rec_t recs[256];
static int n_recs = 0;
static const size_t DSIZE[4] = {4, 8, 8, 1};

static void load(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  for (;;) {
    uint32_t nl;
    if (fread(&nl, 4, 1, f) != 1) break;
    rec_t *r = &recs[n_recs++];
    memset(r->name, 0, sizeof(r->name));
    if (nl >= sizeof(r->name) || fread(r->name, 1, nl, f) != nl) { fprintf(stderr, "bad record\n"); exit(2); }
    uint8_t dt;
    if (fread(&dt, 1, 1, f) != 1 || fread(&r->count, 8, 1, f) != 1) { fprintf(stderr, "bad record\n"); exit(2); }
    r->dtype = dt;
    r->data = malloc(r->count * DSIZE[dt] + 16);
    if (r->count && fread(r->data, DSIZE[dt], r->count, f) != r->count) { fprintf(stderr, "short record %s\n", r->name); exit(2); }
  }
  fclose(f);
}
static rec_t *find(const char *name) {
  for (int i = 0; i < n_recs; i++)
    if (!strcmp(recs[i].name, name)) return &recs[i];
  return NULL;
}
static void *need(const char *name) {
  rec_t *r = find(name);
  if (!r) { fprintf(stderr, "missing input record %s\n", name); exit(2); }
  return r->data;
}
static int32_t scalar(const char *name) { return *(int32_t *)need(name); }

static FILE *out;
static void emit(const char *name, int dtype, uint64_t count, const void *data) {
  uint32_t nl = (uint32_t)strlen(name);
  uint8_t dt = (uint8_t)dtype;
  fwrite(&nl, 4, 1, out);
  fwrite(name, 1, nl, out);
  fwrite(&dt, 1, 1, out);
  fwrite(&count, 8, 1, out);
  if (count) fwrite(data, DSIZE[dtype], count, out);
}

/* the bound entry points (JstacsHmmNative.java) */
static const char *(*p_last_error)(void);
static int (*p_model_create)(const jh_model_desc *, jh_model **);
static void (*p_model_destroy)(jh_model *);
static int (*p_model_set_params)(jh_model *, const double *, const double *, const double *, const double *);
static int (*p_model_get_params)(jh_model *, double *, double *, double *, double *);
static int (*p_model_sizes)(const jh_model *, int64_t[4]);
static int (*p_dataset_create)(const uint8_t *, const int64_t *, int64_t, const double *, int32_t, jh_dataset **);
static void (*p_dataset_destroy)(jh_dataset *);
static int (*p_dataset_set_weights)(jh_dataset *, const double *);
static int (*p_loglik)(jh_model *, jh_dataset *, double *);
static int (*p_loglik_windows)(jh_model *, jh_dataset *, int64_t, const int64_t *, const int64_t *, const int64_t *, double *);
static int (*p_viterbi_u8)(jh_model *, jh_dataset *, uint8_t *, double *);
static int (*p_viterbi_paths)(jh_model *, jh_dataset *, int32_t *, const int64_t *, int64_t *, double *);
static int (*p_posterior)(jh_model *, jh_dataset *, int64_t, double *);
static int (*p_posterior_decode)(jh_model *, jh_dataset *, int32_t *, double *);
static int (*p_posterior_decode_u8)(jh_model *, jh_dataset *, uint8_t *, double *);
static int (*p_baum_welch)(jh_model *, jh_dataset *, const jh_train_opts *, double *, int32_t, int32_t *);
static int (*p_last_objective)(jh_model *, double *);
static int (*p_model_set_states_groups)(jh_model *, int32_t, const int32_t *, const int32_t *);
static int (*p_dataset_set_groups)(jh_dataset *, const int32_t *);

#define BIND(var, sym)                                                 \
  do {                                                                 \
    *(void **)(&var) = dlsym(lib, sym);                                \
    if (!var) { fprintf(stderr, "missing symbol %s\n", sym); exit(2); } \
  } while (0)
#define CHECK(call)                                                                        \
  do {                                                                                     \
    int st__ = (call);                                                                     \
    if (st__ != 0) {                                                                       \
      fprintf(stderr, "%s -> %d: %s\n", #call, st__, p_last_error());                      \
      emit("error_status", 0, 1, &(int32_t){st__});                                        \
      fclose(out);                                                                         \
      return 3;                                                                            \
    }                                                                                      \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s lib in.bin out.bin op...\n", argv[0]); return 2; }
  void *lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  BIND(p_last_error, "jh_last_error"); BIND(p_model_create, "jh_model_create"); BIND(p_model_destroy, "jh_model_destroy");
  BIND(p_model_set_params, "jh_model_set_params"); BIND(p_model_get_params, "jh_model_get_params"); BIND(p_model_sizes, "jh_model_sizes");
  BIND(p_dataset_create, "jh_dataset_create"); BIND(p_dataset_destroy, "jh_dataset_destroy"); BIND(p_dataset_set_weights, "jh_dataset_set_weights");
  BIND(p_loglik, "jh_loglik"); BIND(p_loglik_windows, "jh_loglik_windows"); BIND(p_viterbi_u8, "jh_viterbi_u8");
  BIND(p_viterbi_paths, "jh_viterbi_paths"); BIND(p_posterior, "jh_posterior"); BIND(p_posterior_decode, "jh_posterior_decode");
  BIND(p_posterior_decode_u8, "jh_posterior_decode_u8"); BIND(p_baum_welch, "jh_baum_welch"); BIND(p_last_objective, "jh_last_objective");
  BIND(p_model_set_states_groups, "jh_model_set_states_groups"); BIND(p_dataset_set_groups, "jh_dataset_set_groups");
  load(argv[2]);
  out = fopen(argv[3], "wb");
  if (!out) { perror(argv[3]); return 2; }

  /* ---- model(): flatten once, jh_model_create ---- */
  jh_model_desc d;
  memset(&d, 0, sizeof(d));
  d.abi_version = JH_ABI_VERSION;
  d.n_states = scalar("n_states"); d.alphabet = scalar("alphabet"); d.max_order = scalar("max_order");
  d.n_elem = scalar("n_elem"); d.n_emis = scalar("n_emis");
  d.lc_ctx_ptr = need("lc_ctx_ptr"); d.ctx_elem = need("ctx_elem"); d.ctx_last_state = need("ctx_last_state");
  d.elem_child_ptr = need("elem_child_ptr"); d.child_state = need("child_state"); d.child_desc = need("child_desc");
  d.elem_ctx = need("elem_ctx"); d.trans_index = need("trans_index"); d.trans_param = need("trans_param");
  d.trans_lognorm = need("trans_lognorm"); d.trans_hyper = need("trans_hyper"); d.state_emis = need("state_emis");
  d.emis_order = need("emis_order"); d.emis_cond_ptr = need("emis_cond_ptr"); d.emis_param = need("emis_param");
  d.emis_lognorm = need("emis_lognorm"); d.emis_hyper = need("emis_hyper"); d.state_silent = need("state_silent");
  d.state_final = need("state_final");
  jh_model *model = NULL;
  CHECK(p_model_create(&d, &model));
  if (find("group_ptr")) /* model(): the constructor's statesGroups, when there are any */
    CHECK(p_model_set_states_groups(model, (int32_t)find("group_ptr")->count - 1, need("group_ptr"), need("group_states")));
  int64_t sz[4];
  CHECK(p_model_sizes(model, sz));
  /* the "Java objects": parameters as the host side currently holds them */
  double *tp = malloc(8 * (sz[0] + 1)), *tl = malloc(8 * (sz[1] + 1)), *ep = malloc(8 * (sz[2] + 1)), *el = malloc(8 * (sz[3] + 1));
  memcpy(tp, d.trans_param, 8 * sz[0]); memcpy(tl, d.trans_lognorm, 8 * sz[1]);
  memcpy(ep, d.emis_param, 8 * sz[2]); memcpy(el, d.emis_lognorm, 8 * sz[3]);
  int n_silent = 0;
  for (int s = 0; s < d.n_states; s++) n_silent += d.state_silent[s] ? 1 : 0;

  const uint8_t *codes = need("codes");
  const int64_t *offsets = need("offsets");
  int64_t n = (int64_t)find("offsets")->count - 1;
  const double *weights = find("weights") ? find("weights")->data : NULL;
  jh_dataset *ds = NULL;
  int have_ds = 0;
  char key[96];

  for (int a = 4; a < argc; a++) {
    const char *op = argv[a];
    if (!strcmp(op, "newdata")) {  /* the DataSet object changed: the cached handle is dropped */
      if (have_ds) p_dataset_destroy(ds);
      have_ds = 0;
      continue;
    }
    const int is_train = !strncmp(op, "train:", 6);
    /* ---- dataset(data, weights): cached per DataSet ---- */
    const double *w = is_train ? weights : NULL;
    if (!have_ds) {
      CHECK(p_dataset_create(codes, offsets, n, w, d.alphabet, &ds));
      have_ds = 1;
    } else {
      CHECK(p_dataset_set_weights(ds, w));
    }
    /* per-position groups: attached by train() when the model has a sequence annotation type (record "seq_groups"),
       detached for every other call (GpuHmmEngine.dataset / attachGroups) */
    CHECK(p_dataset_set_groups(ds, is_train && find("seq_groups") ? (const int32_t *)find("seq_groups")->data : NULL));
    if (is_train) {
      int mode, kind, max_iter, starts;
      double eps;
      if (sscanf(op + 6, "%d:%d:%d:%lf:%d", &mode, &kind, &max_iter, &eps, &starts) != 5) { fprintf(stderr, "bad op %s\n", op); return 2; }
      jh_train_opts o;
      memset(&o, 0, sizeof(o));
      o.mode = mode; o.term_kind = kind; o.max_iter = max_iter; o.eps = eps;
      int cap = kind == JH_TERM_ITERATIONS ? (max_iter > 1 ? max_iter : 1) : 100000;
      double *obj = malloc(8 * (size_t)cap), best = -1.0 / 0.0;
      double *btp = NULL, *btl = NULL, *bep = NULL, *bel = NULL;
      for (int start = 0; start < starts; start++) {
        /* initialize(data, weights): the harness takes the "random" initial parameters of start k from the input */
        snprintf(key, sizeof(key), "start%d_trans_param", start);
        if (find(key)) {
          memcpy(tp, find(key)->data, 8 * sz[0]);
          snprintf(key, sizeof(key), "start%d_trans_lognorm", start); memcpy(tl, need(key), 8 * sz[1]);
          snprintf(key, sizeof(key), "start%d_emis_param", start); memcpy(ep, need(key), 8 * sz[2]);
          snprintf(key, sizeof(key), "start%d_emis_lognorm", start); memcpy(el, need(key), 8 * sz[3]);
        }
        CHECK(p_model_set_params(model, tp, tl, ep, el));
        int32_t n_it = 0;
        CHECK(p_baum_welch(model, ds, &o, obj, cap, &n_it));
        double last;
        CHECK(p_last_objective(model, &last));
        CHECK(p_model_get_params(model, tp, tl, ep, el));  /* write back into the "Java objects" */
        snprintf(key, sizeof(key), "train_objective_start%d", start);
        emit(key, 2, (uint64_t)(n_it < cap ? n_it : cap), obj);
        snprintf(key, sizeof(key), "train_last_start%d", start);
        emit(key, 2, 1, &last);
        if (last > best) {
          best = last;
          if (starts > 1) {
            if (!btp) { btp = malloc(8 * (sz[0] + 1)); btl = malloc(8 * (sz[1] + 1)); bep = malloc(8 * (sz[2] + 1)); bel = malloc(8 * (sz[3] + 1)); }
            memcpy(btp, tp, 8 * sz[0]); memcpy(btl, tl, 8 * sz[1]); memcpy(bep, ep, 8 * sz[2]); memcpy(bel, el, 8 * sz[3]);
          }
        }
      }
      if (btp) { memcpy(tp, btp, 8 * sz[0]); memcpy(tl, btl, 8 * sz[1]); memcpy(ep, bep, 8 * sz[2]); memcpy(el, bel, 8 * sz[3]); }
      emit("train_best", 2, 1, &best);
      emit("trained_trans_param", 2, (uint64_t)sz[0], tp); emit("trained_trans_lognorm", 2, (uint64_t)sz[1], tl);
      emit("trained_emis_param", 2, (uint64_t)sz[2], ep); emit("trained_emis_lognorm", 2, (uint64_t)sz[3], el);
      free(obj);
      continue;
    }
    CHECK(p_model_set_params(model, tp, tl, ep, el));  /* pushParameters */
    if (!strcmp(op, "score")) {
      double *res = malloc(8 * (size_t)(n + 1));
      CHECK(p_loglik(model, ds, res));
      emit("loglik", 2, (uint64_t)n, res);
      free(res);
    } else if (!strcmp(op, "viterbi")) {
      double *scores = malloc(8 * (size_t)(n + 1));
      if (n_silent == 0 && d.n_states <= 255) {
        uint8_t *paths = malloc((size_t)offsets[n] + 16);
        CHECK(p_viterbi_u8(model, ds, paths, scores));
        int32_t *wide = malloc(4 * ((size_t)offsets[n] + 1));  /* IntList.add(byte & 0xFF) */
        for (int64_t i = 0; i < offsets[n]; i++) wide[i] = paths[i];
        emit("viterbi_paths", 0, (uint64_t)offsets[n], wide);
        emit("viterbi_ptr", 1, (uint64_t)(n + 1), offsets);
        free(paths); free(wide);
      } else {
        int64_t *ptr = malloc(8 * (size_t)(n + 1)), *len = malloc(8 * (size_t)(n + 1)), total = 0;
        for (int64_t i = 0; i < n; i++) { ptr[i] = total; total += (offsets[i + 1] - offsets[i]) * (1 + n_silent) + n_silent; }
        ptr[n] = total;
        int32_t *paths = malloc(4 * ((size_t)total + 1));
        CHECK(p_viterbi_paths(model, ds, paths, ptr, len, scores));
        /* compact to the path lengths, as the shim builds the IntLists */
        int64_t *cptr = malloc(8 * (size_t)(n + 1)), q = 0;
        int32_t *cp = malloc(4 * ((size_t)total + 1));
        for (int64_t i = 0; i < n; i++) {
          cptr[i] = q;
          if (len[i] < 0) { fprintf(stderr, "Impossible path\n"); emit("error_status", 0, 1, &(int32_t){JH_ERR_IMPOSSIBLE}); fclose(out); return 3; }
          for (int64_t l = 0; l < len[i]; l++) cp[q++] = paths[ptr[i] + l];
        }
        cptr[n] = q;
        emit("viterbi_paths", 0, (uint64_t)q, cp);
        emit("viterbi_ptr", 1, (uint64_t)(n + 1), cptr);
        free(ptr); free(len); free(paths); free(cptr); free(cp);
      }
      emit("viterbi_scores", 2, (uint64_t)n, scores);
      free(scores);
    } else if (!strcmp(op, "decode")) {
      int32_t *wide = malloc(4 * ((size_t)offsets[n] + 1));
      if (d.n_states <= 255) {
        uint8_t *paths = malloc((size_t)offsets[n] + 16);
        CHECK(p_posterior_decode_u8(model, ds, paths, NULL));
        for (int64_t i = 0; i < offsets[n]; i++) wide[i] = paths[i];
        free(paths);
      } else {
        CHECK(p_posterior_decode(model, ds, wide, NULL));
      }
      emit("decoded_paths", 0, (uint64_t)offsets[n], wide);
      free(wide);
    } else if (!strcmp(op, "posterior")) {
      for (int64_t i = 0; i < n; i++) {
        int64_t L = offsets[i + 1] - offsets[i];
        double *m = malloc(8 * (size_t)(d.n_states * L + 1));
        CHECK(p_posterior(model, ds, i, m));
        snprintf(key, sizeof(key), "posterior_%lld", (long long)i);
        emit(key, 2, (uint64_t)(d.n_states * L), m);
        free(m);
      }
    } else if (!strcmp(op, "windows")) {
      int64_t nw = (int64_t)find("win_seq")->count;
      double *res = malloc(8 * (size_t)(nw + 1));
      CHECK(p_loglik_windows(model, ds, nw, need("win_seq"), need("win_start"), need("win_end"), res));
      emit("window_loglik", 2, (uint64_t)nw, res);
      free(res);
    } else {
      fprintf(stderr, "unknown op %s\n", op);
      return 2;
    }
  }
  if (have_ds) p_dataset_destroy(ds);
  p_model_destroy(model);
  emit("done", 0, 1, &(int32_t){1});
  fclose(out);
  return 0;
}
