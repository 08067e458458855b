/*
 * okb200.h — C ABI of libokb200.so, the B200-native (sm_100a) replacement for the
 * knowledge-graph-embedding hot path of luigiba/OpenKEonSpark.
 *
 * Two layers are exported.
 *
 *  (1) Reference-compatible layer: the same symbols, argument order, widths (INT = int64,
 *      REAL = float) and process-global state as the reference's release/Base.so, so the
 *      reference's own ctypes binding (/root/reference/Config.py:30-51) can load this library
 *      instead of Base.so.  Pointers are HOST buffers; each call runs the CUDA kernels on the
 *      current device and copies the result back.
 *
 *  (2) Native layer (okb_*): an explicit context instead of globals, DEVICE pointers and a
 *      cudaStream_t (passed as void*), int status codes (0 = ok; okb_last_error() explains).
 *      It also carries what the reference delegates to TensorFlow: the model's train step
 *      (loss_def + optimizer) and scoring (predict_def).
 *
 * No torch / C++ types cross this boundary.  The library never falls back to the CPU for the
 * hot path: without a CUDA device every compute entry point returns OKB_ERR_CUDA.
 *
 * Error behaviour of layer (1): like the reference (base/Reader.h:36-39 prints and returns), a failed call prints
 * "libokb200: <symbol> failed: <why>" to stderr and returns without filling its outputs; the text stays available
 * through okb_last_error(okb_default_ctx()).  Set OKB200_ABORT_ON_ERROR=1 to abort() instead.
 */
#ifndef OKB200_H
#define OKB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t INT;
typedef float REAL;

/* ------------------------------------------------------------------------------------------
 * (1) Reference-compatible layer (process-global default context)
 * ---------------------------------------------------------------------------------------- */
void setInPath(char *path);                 /* base/Setting.h:12-19 */
void setOutPath(char *path);                /* base/Setting.h:21-28 */
void setWorkThreads(INT threads);           /* base/Setting.h:36-39 */
INT getWorkThreads(void);                   /* base/Setting.h:41-44 */
void setBern(INT con);                      /* base/Setting.h:110-113 */
INT getEntityTotal(void);                   /* base/Setting.h:63-66 */
INT getRelationTotal(void);                 /* base/Setting.h:68-71 */
INT getTripleTotal(void);                   /* base/Setting.h:73-76 */
INT getTrainTotal(void);                    /* base/Setting.h:78-81  (deduplicated) */
INT getTrainTotal_(void);                   /* base/Setting.h:84-87  (file rows) */
INT getBatchTotal(void);                    /* base/Setting.h:90-93 */
INT getTestTotal(void);                     /* base/Setting.h:95-98 */
INT getValidTotal(void);                    /* base/Setting.h:100-103 */
void randReset(void);                       /* base/Random.h:8-13   (seeds from libc rand()) */
void importTrainFiles(void);                /* base/Reader.h:26-179 */
void importTestFiles(void);                 /* base/Reader.h:185-292 */
void importTypeFiles(void);                 /* base/Reader.h:301-365 */
void importOntologyFiles(void);             /* base/Reader.h:375-449 */
/* base/Base.cpp:149-172.  batch_h/t/r: INT[batchSize*(1+negRate+negRelRate)], plane-major. */
void sampling(INT *batch_h, INT *batch_t, INT *batch_r, REAL *batch_y, INT batchSize, INT negRate,
              INT negRelRate);
void getHeadBatch(INT index, INT *ph, INT *pt, INT *pr);   /* base/Test.h:10-17 */
void getTailBatch(INT index, INT *ph, INT *pt, INT *pr);   /* base/Test.h:19-26 */
/* base/Test.h:30-136 / 140-249.  con = REAL[entityTotal] scores.  Returns INT[8]; unlike the
 * reference (which leaks a new INT[8] per call) the buffer is owned by the library and reused. */
INT *testHead(INT index, REAL *con);
INT *testTail(INT index, REAL *con);
void getNegTest(void);                      /* base/Test.h:257-264 */
void getNegValid(void);                     /* base/Test.h:267-274 */
void getTestBatch(INT *ph, INT *pt, INT *pr, INT *nh, INT *nt, INT *nr);    /* base/Test.h:276-287 */
void getValidBatch(INT *ph, INT *pt, INT *pr, INT *nh, INT *nt, INT *nr);   /* base/Test.h:289-300 */
void getBestThreshold(REAL *relThresh, REAL *score_pos, REAL *score_neg);   /* base/Test.h:303-341 */
void test_triple_classification(REAL *relThresh, REAL *score_pos, REAL *score_neg,
                                REAL *acc_addr);                            /* base/Test.h:345-387 */
INT get_n_interval(INT r, REAL *score_pos, REAL *score_neg);                /* base/Test.h:390-407 */
INT *get_TPFP(INT r, REAL *score_pos, REAL *score_neg, REAL *score_pos_test,
              REAL *score_neg_test);                                        /* base/Test.h:410-444 */

/* ------------------------------------------------------------------------------------------
 * (2) Native layer
 * ---------------------------------------------------------------------------------------- */
typedef struct okb_ctx okb_ctx;

enum { OKB_OK = 0, OKB_ERR_ARG = 1, OKB_ERR_IO = 2, OKB_ERR_CUDA = 3, OKB_ERR_STATE = 4 };
enum { OKB_TRANSE = 0, OKB_TRANSH = 1, OKB_TRANSR = 2, OKB_TRANSD = 3 };
enum { OKB_SGD = 0, OKB_ADAM = 1 };

int okb_version(void);
okb_ctx *okb_default_ctx(void);                    /* the context behind layer (1) */
int okb_create(okb_ctx **out);
int okb_destroy(okb_ctx *c);
const char *okb_last_error(okb_ctx *c);
int okb_set_device(okb_ctx *c, int device);       /* cudaSetDevice for this thread; one process per GPU */

/* ---- dataset: replaces Setting.h / Reader.h.  Host-side parse + index build (C++), then the
 *      index is uploaded once and stays resident in HBM as int32 SoA. */
int okb_set_in_path(okb_ctx *c, const char *path);
int okb_set_bern(okb_ctx *c, INT flag);
int okb_set_work_threads(okb_ctx *c, INT w);
int okb_import_train_files(okb_ctx *c);
int okb_import_test_files(okb_ctx *c);
int okb_import_type_files(okb_ctx *c);
int okb_import_ontology_files(okb_ctx *c);
/* Same as the file importers but from host arrays (rows in file order; columns h, t, r).
 * Used for large synthetic graphs where writing and re-parsing text is pointless. */
int okb_import_train_arrays(okb_ctx *c, INT n_ent, INT n_rel, const INT *h, const INT *t, const INT *r,
                            INT n, INT new_batch_total);
int okb_import_test_arrays(okb_ctx *c, const INT *th, const INT *tt, const INT *tr, INT n_test,
                           const INT *vh, const INT *vt, const INT *vr, INT n_valid);
/* Build the type constraints from the loaded train+valid+test triples, as the reference's n_n()
 * generator does before every run (main_spark.py:209-290), without going through type_constrain.txt. */
int okb_build_type_constraints(okb_ctx *c);
/* what: 0 entities, 1 relations, 2 train rows (file), 3 train dedup, 4 test, 5 valid, 6 all triples,
 *       7 new-batch total */
INT okb_total(okb_ctx *c, int what);

/* ---- RNG streams (Random.h).  okb_rand_reset draws the seeds from libc rand() like the
 *      reference; okb_set_streams / okb_get_streams inject / read the W 64-bit LCG states. */
int okb_rand_reset(okb_ctx *c);
int okb_set_streams(okb_ctx *c, const uint64_t *state, INT w);
int okb_get_streams(okb_ctx *c, uint64_t *state_out, INT w);

/* ---- sampler (Base.cpp:74-172, Corrupt.h:7-101) on the GPU.
 * Fills the context's device-resident batch (int32, plane-major) for `steps` consecutive
 * sampling() calls in ONE launch (LCG jump-ahead makes every slot independent) and advances
 * the stream states exactly as `steps` reference calls would.
 * stream_lo/stream_hi: only slots owned by streams [stream_lo, stream_hi) are produced
 * (data-parallel ranks); pass 0, W for all. */
int okb_sample(okb_ctx *c, INT batch_size, INT neg_ent, INT neg_rel, INT steps, INT stream_lo,
               INT stream_hi, void *cuda_stream);
/* Device pointers to step `step` of the last okb_sample: int32[batch_size*(1+neg_ent+neg_rel)]. */
int okb_batch_ptrs(okb_ctx *c, INT step, const int32_t **h, const int32_t **t, const int32_t **r);
/* Copy step `step` to host in the reference's types (int64 ids, float labels +1/-1). */
int okb_batch_to_host(okb_ctx *c, INT step, INT *h, INT *t, INT *r, REAL *y, void *cuda_stream);
/* One reference sampling() call (Base.cpp:153-177) returning the batch in the caller's int64 arrays: okb_sample of ONE
 * step followed by okb_batch_to_host, as one launch when h/t/r are one page-locked block [3][S] (the sample kernel then
 * stores the int64 copy into it over PCIe itself and the call returns as soon as that kernel has finished).  The batch
 * also stays resident on the device as step 0.  Labels are the fixed +1/-1 pattern and are not written. */
int okb_sample_to_host(okb_ctx *c, INT batch_size, INT neg_ent, INT neg_rel, INT stream_lo, INT stream_hi, INT *h, INT *t,
                       INT *r, void *cuda_stream);
/* Replace the device batch of step 0 with caller-provided host ids (Config.train_step path). */
int okb_batch_from_host(okb_ctx *c, INT batch_size, INT neg_ent, INT neg_rel, const INT *h, const INT *t,
                        const INT *r, void *cuda_stream);

/* The ids okb_batch_from_host receives are validated on the device: an id outside [0, E) / [0, R) is replaced by 0 (so
 * nothing indexes out of bounds), the step that trains on such a batch leaves every table untouched and reports a NaN
 * loss, and this call (synchronises the stream) returns OKB_ERR_ARG once and clears the flag.  okb_train_step_host
 * calls it itself when the loss comes back NaN.  (TF's embedding_lookup raises InvalidArgument in the reference.) */
int okb_batch_check(okb_ctx *c, void *cuda_stream);

/* ---- model parameters: device pointers owned by the caller (row-major fp32). */
typedef struct {
    int32_t model;        /* OKB_TRANSE .. OKB_TRANSD */
    int32_t ent_dim;      /* De (hidden_size / ent_size) */
    int32_t rel_dim;      /* Dr (== De except TransR) */
    int32_t optimizer;    /* OKB_SGD / OKB_ADAM */
    float *ent;           /* ent_embeddings  [E, De] */
    float *rel;           /* rel_embeddings  [R, Dr] */
    float *ent_aux;       /* TransD ent_transfer [E, De]; else NULL */
    float *rel_aux;       /* TransH normal_vectors [R, D]; TransD rel_transfer [R, D];
                             TransR transfer_matrix [R, De*Dr]; TransE NULL */
    /* Adam slots, same shapes as the four tables above (NULL for SGD) */
    float *m_ent, *v_ent, *m_rel, *v_rel, *m_ent_aux, *v_ent_aux, *m_rel_aux, *v_rel_aux;
} okb_model;

typedef struct {
    float margin;         /* Config.margin */
    float lr;             /* SGD: alpha.  Adam: lr_t = alpha*sqrt(1-b2^t)/(1-b1^t), computed by the host */
    float beta1, beta2, eps;
} okb_hyper;

/* ---- train step = TransX.loss_def + optimizer.minimize (TransE.py:26-51 ...,
 *      distribute_training.py:94-101) on step `step` of the sampled batch.
 * Phases (all asynchronous on cuda_stream):
 *   okb_plan   : gradient-row keys of the batch + stable radix sort (integer work only)
 *   okb_grad   : fused gather / project / normalise / L1 / hinge / backward for positives
 *                [b_lo, b_hi); writes one gradient row per distinct (positive-group, role)
 *                into grad_ent / grad_rel and per-positive loss terms into loss_terms
 *   okb_update : segmented sum of gradient rows in sorted (deterministic) order fused with the
 *                SGD or TF1-Adam update
 * okb_train_step runs the three phases on context-owned buffers and writes the mean loss to
 * loss_out (device float).  The phases are public so that data-parallel ranks can all-gather
 * grad_ent / grad_rel / loss_terms between okb_grad and okb_update. */
/* TransR: grad_rel holds one [d rel_embeddings | d transfer_matrix] row per RELATION (rel_rows = R; the train kernel emits the
 * gradients already reduced per relation) when neg_rel == 0, and one such row per (positive, relation slot) — like the other
 * models — when neg_rel > 0 (TransR.py:61-65: every negative is projected by its own relation's matrix). */
int okb_grad_sizes(okb_ctx *c, const okb_model *m, INT batch_size, INT neg_ent, INT neg_rel,
                   INT *ent_rows, INT *ent_cols, INT *rel_rows, INT *rel_cols);
int okb_plan(okb_ctx *c, INT step, void *cuda_stream);
/* Plan steps [step_lo, step_hi) of the last okb_sample with ONE radix sort (sampling and planning do not
 * depend on the parameters, so whole chunks of steps are prepared ahead of the train kernels). */
int okb_plan_steps(okb_ctx *c, INT step_lo, INT step_hi, void *cuda_stream);
int okb_grad(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, INT b_lo, INT b_hi,
             float *grad_ent, float *grad_rel, float *loss_terms, void *cuda_stream);
int okb_update(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, const float *grad_ent,
               const float *grad_rel, const float *loss_terms, float *loss_out, void *cuda_stream);
int okb_train_step(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, float *loss_out,
                   void *cuda_stream);
/* n consecutive steps in one call (plans them with one sort if needed); hp[n] host array (Adam's lr_t differs
 * per step), loss_out: device float[n] or NULL. */
int okb_train_steps(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step_lo, INT n, float *loss_out,
                    void *cuda_stream);
/* loss_out may also be PAGE-LOCKED HOST memory (it is written with one plain store followed by a system fence).  The
 * host-batch path (Config.train_step, the reference's feed_dict call that returns the loss, Config.py:464-475) uses
 * that instead of a device->host copy: the caller stores `sentinel` (a bit pattern no loss can have) in *word, launches
 * the step, and okb_wait_word spins until the word changes — the Adam kernel's loss blocks are the FIRST blocks of its
 * grid, so the loss arrives while the table update is still running and the next batch is prepared under it.  Falls
 * back to a stream synchronize if the stream goes idle first; returns OKB_ERR_CUDA if the stream failed. */
int okb_wait_word(okb_ctx *c, const void *host_word, unsigned sentinel, void *cuda_stream);
/* The whole host-batch step as one call: okb_batch_from_host + okb_train_step(step 0, loss -> loss_word) + okb_wait_word.
 * loss_word: one page-locked host float (cudaHostAlloc / cudaHostRegister / torch pin_memory). */
int okb_train_step_host(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT batch_size, INT neg_ent, INT neg_rel,
                        const INT *h, const INT *t, const INT *r, float *loss_word, void *cuda_stream);

/* ---- synchronous data-parallel training inside one box, one process per GPU (replaces the asynchronous
 *      parameter-server path of distribute_training.py:161-364).  Owner-sharded: every rank keeps the full tables in a
 *      PEER ARENA (device memory the other ranks map over NVLink with CUDA IPC), computes the gradient rows of its own
 *      positives, pushes per-row partial sums into the row owner's staging slab with peer stores, and the owner applies
 *      SGD / TF1-Adam to its rows and stores the new rows into every rank's table.  No NCCL call on the data path.
 *   okb_peer_alloc / okb_peer_open : cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle (handle = 64 bytes, exchanged
 *                                    by the host, e.g. torch.distributed.all_gather_object)
 *   okb_dp_layout  : arena size and part offsets for a model at a world size
 *   okb_dp_attach  : this rank's geometry + every rank's arena as mapped in this process
 *   okb_dp_train_steps : n steps of the last okb_sample (which must cover this rank's streams): plan + grad of positives
 *                        [b_lo, b_hi) + fused reduce/push + owner update; loss_out[i] = mean hinge of the GLOBAL batch. */
#define OKB_DP_MAX 16
typedef struct {
    int32_t rank, world;
    INT b_lo, b_hi;                 /* this rank's positives of the global batch (its sampler streams' slots) */
    void *arena[OKB_DP_MAX];        /* arena base of every rank as mapped HERE (arena[rank] is the local allocation) */
    INT off_ent, off_ent_aux, off_rel, off_rel_aux;   /* tables inside an arena (bytes; -1 = absent) */
    INT off_stage_ent, off_stage_rel;                 /* staging slabs [world][ceil(rows/world)][cols] */
    INT off_flags;                                    /* 512-byte flag block */
    INT arena_bytes;
    /* "pull" form of the update (optional; set plan_steps > 0 before okb_dp_layout): the rank's plan (row map, slot
     * permutation), gradient rows and loss terms also live in the arena, so the row OWNER sums the partial rows of all
     * ranks straight out of their gradient buffers (peer loads) and the separate reduce+push kernel disappears. */
    INT plan_steps;                 /* steps planned per chunk at most (Config.plan_ahead) */
    INT max_local;                  /* positives per rank at most */
    INT neg_ent, neg_rel;
    INT off_rowhead, off_perm, off_gent, off_grel, off_loss, off_partial;   /* -1 = not reserved */
    /* "scatter" form (set scatter = 1, global_batch, neg_ent, neg_rel before okb_dp_layout): every rank samples and plans the
     * GLOBAL batch; the grad kernel stores each gradient row of this rank's positives straight into the arena of the rank
     * that owns its table row (off_gent / off_grel: receive buffers with one slot per gradient row of the global batch;
     * off_loss: hinge terms, stored to every rank); the owner runs the single-GPU update over its rows and stores the new
     * rows into every rank's tables.  Two kernels and two flag exchanges per step, no staging slabs, and tables + loss
     * bit-identical to ONE GPU training the global batch.
     * scatter = 2 is the "gather" form: the grad kernel stores every gradient row into EVERY rank's arena (two receive
     * buffers used in turn) and every rank runs the full single-GPU update on its own copy — ONE exchange per step instead
     * of two; (world - 1) x the gradient rows leave every rank, so it pays for two ranks only.  Same bit-identity. */
    INT scatter, global_batch;
} okb_dp;
int okb_peer_alloc(okb_ctx *c, INT bytes, void **dev_ptr, unsigned char *handle64);
int okb_peer_open(okb_ctx *c, const unsigned char *handle64, void **dev_ptr);
int okb_peer_close(okb_ctx *c, void *dev_ptr);
int okb_peer_free(okb_ctx *c, void *dev_ptr);
int okb_dp_layout(okb_ctx *c, const okb_model *m, INT world, okb_dp *out);
int okb_dp_attach(okb_ctx *c, const okb_dp *cfg);
int okb_dp_detach(okb_ctx *c);
int okb_dp_train_steps(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step_lo, INT n, float *loss_out,
                       void *cuda_stream);
/* stream-ordered wait until every rank's row updates of all steps issued so far have landed in THIS rank's tables (a
 * rank's last update kernel stores into its peers' arenas).  Not a collective. */
int okb_dp_quiesce(okb_ctx *c, void *cuda_stream);

/* ---- TransR under data parallelism: RELATION-sharded.  TransR's gradients come out reduced per relation (one CTA owns one
 *      relation's 40 KB matrix), so ranks split the relations instead of the positives: after okb_transr_set_shard(r_lo,
 *      r_hi) okb_grad (called with the full positive range) computes only the positives whose relation lies in [r_lo, r_hi)
 *      — their entity gradient rows, loss terms and the relations' [d rel | d M_r] rows — and okb_update moves only those
 *      relations' rel_embeddings / transfer_matrix rows (entity tables: all rows, from the gradient rows the ranks have
 *      summed).  M_r never crosses NVLink during training; the owners' rows are gathered when the tables are read.
 *      (0, 0) restores "all relations". */
int okb_transr_set_shard(okb_ctx *c, INT r_lo, INT r_hi);

/* ---- chunk pipeline.  Sampling and planning depend only on the RNG streams, so the next chunk of steps can be produced on
 *      an internal side stream while the current chunk trains.  okb_chunk_begin makes `steps` sampled + planned steps
 *      current (swapping in a matching prefetched chunk, else producing them on cuda_stream); okb_chunk_prefetch starts
 *      the following chunk.  Any call that observes or changes the RNG streams discards a prefetched chunk and restores
 *      the streams first, so results never depend on prefetching. */
int okb_chunk_begin(okb_ctx *c, INT batch_size, INT neg_ent, INT neg_rel, INT steps, void *cuda_stream);
int okb_chunk_prefetch(okb_ctx *c, INT batch_size, INT neg_ent, INT neg_rel, INT steps, void *cuda_stream);
/* the same pipeline under owner-sharded data parallelism (after okb_dp_attach): the chunk is sampled from the sampler streams
 * [stream_lo, stream_hi) and planned over the positives this rank plans (its own; the global batch in the scatter form) */
int okb_dp_chunk_begin(okb_ctx *c, INT batch_size, INT neg_ent, INT neg_rel, INT steps, INT stream_lo, INT stream_hi,
                       void *cuda_stream);
int okb_dp_chunk_prefetch(okb_ctx *c, INT batch_size, INT neg_ent, INT neg_rel, INT steps, INT stream_lo, INT stream_hi,
                          void *cuda_stream);

/* ---- scoring = TransX.predict_def (TransE.py:53-58 ...).  h,t,r: device int64[n]; out: device
 *      float[n] (TransE: mean over d; others: sum).  Canonical evaluation order, see DESIGN.md. */
int okb_predict(okb_ctx *c, const okb_model *m, const int64_t *h, const int64_t *t, const int64_t *r, INT n,
                float *out, void *cuda_stream);

/* ---- link prediction = getHead/TailBatch + predict + testHead/testTail for test triples
 *      [q_lo, q_hi) of the (r,h,t)-sorted test list, both sides if heads != 0, against candidate
 *      entities [cand_lo, cand_hi).  counts: device int64[(q_hi-q_lo)*2*4] better-than counts
 *      (raw, filter, raw_constrain, filter_constrain; side 0 = head, 1 = tail) ADDED to the
 *      buffer; best: device uint64[(q_hi-q_lo)*2*4] packed (score_bits<<32 | id) minima, combined
 *      with atomicMin.  Candidate-sharded ranks all-reduce counts (sum) and best (min). */
int okb_rank(okb_ctx *c, const okb_model *m, INT q_lo, INT q_hi, int heads, INT cand_lo, INT cand_hi,
             int64_t *counts, uint64_t *best, void *cuda_stream);
/* counts/best -> the reference's 8-int records (Test.h:99-134): out device int64[(q_hi-q_lo)*2*8] */
int okb_rank_finalize(okb_ctx *c, INT q_lo, INT q_hi, const int64_t *counts, const uint64_t *best,
                      int64_t *out, void *cuda_stream);
/* Reference-shaped single query with caller-provided scores (device float[E]); out device int64[8]. */
int okb_rank_scores(okb_ctx *c, INT index, int side, const float *scores, int64_t *out, void *cuda_stream);

/* ---- triple classification (Test.h:253-387).  Negatives are drawn on the host (libc rand(), inherently sequential);
 *      scores, thresholds and counts are computed on the device. */
int okb_tc_batch(okb_ctx *c, int which /*0 test, 1 valid*/, INT *ph, INT *pt, INT *pr, INT *nh, INT *nt, INT *nr);
int okb_best_threshold(okb_ctx *c, REAL *rel_thresh, const REAL *score_pos, const REAL *score_neg);
int okb_tc_eval(okb_ctx *c, const REAL *rel_thresh, const REAL *score_pos, const REAL *score_neg,
                INT *tp_tn_fp_fn, REAL *acc);
/* accuracy of rel_thresh on the VALID triples (the early-stopping check of distribute_training.py:299-316; see the
 * note in csrc/loader.cpp about the reference's use of the test ranges there) */
int okb_tc_eval_valid(okb_ctx *c, const REAL *rel_thresh, const REAL *score_pos, const REAL *score_neg,
                      INT *tp_tn_fp_fn, REAL *acc);
/* The same two steps with DEVICE score arrays (index-aligned with the (r,h,t)-sorted valid / test lists, e.g. straight
 * out of okb_predict) — csrc/tc.cu: one CTA per relation scans the threshold grid min + i * 0.01 (Test.h:303-341), a
 * second kernel counts TP / TN / FP / FN over the test (on_valid = 0) or valid (on_valid = 1) ranges (Test.h:345-387).
 * thresh: device float[R], in/out (relations without valid triples keep their entry).  The host-pointer entry points above
 * stage their arrays and call these; results are bit-identical to the reference library. */
int okb_tc_thresholds_dev(okb_ctx *c, const float *score_pos, const float *score_neg, float *thresh, void *cuda_stream);
int okb_tc_counts_dev(okb_ctx *c, const float *thresh, const float *score_pos, const float *score_neg, int on_valid,
                      INT *tp_tn_fp_fn /*host[4]*/, REAL *acc /*host*/, void *cuda_stream);
int okb_test_list(okb_ctx *c, int which, INT *h, INT *t, INT *r);
INT okb_n_interval(okb_ctx *c, INT r, const REAL *score_pos, const REAL *score_neg);            /* Test.h:390-407 */
INT *okb_tpfp(okb_ctx *c, INT r, const REAL *score_pos, const REAL *score_neg, const REAL *score_pos_test,
              const REAL *score_neg_test);                                                        /* Test.h:410-444 */

/* Flags.  OKB_FLAG_TRANSR_TC = 1: okb_rank projects TransR candidates (Ent . M_r per relation group) on the tcgen05
 * tensor cores with 3xTF32 split operands instead of the canonical sequential-fp32 kernel.  Scores then agree with
 * the canonical order to ~1e-6 relative instead of bit-for-bit, so the flag is off by default. */
enum { OKB_FLAG_TRANSR_TC = 1,
       /* OKB_FLAG_PDL = 2 (default on): launch the grad and update kernels with programmatic stream serialization so
        * that each one's parameter-independent prologue overlaps the tail of its predecessor. */
       OKB_FLAG_PDL = 2,
       /* OKB_FLAG_ADAM_TMA = 3 (default off): run the dense Adam pass as one wave of CTAs whose x / m / v tiles are staged
        * by bulk-async (TMA) copies instead of the register-only kernel; bit-identical results, kept for A/B runs. */
       OKB_FLAG_ADAM_TMA = 3,
       /* OKB_FLAG_L2_PREFETCH = 4 (default off): with Adam, the (latency-bound) grad kernel issues bulk L2 prefetches of
        * the tables and their m / v slots, so the dense update pass that follows streams from L2.  Measured on the bench
        * workload: grad +2.1 us, update -1.5 us. */
       OKB_FLAG_L2_PREFETCH = 4,
       /* OKB_FLAG_ADAM_LEGACY = 5 (default off): the first, grid-stride form of the dense Adam pass (A/B runs). */
       OKB_FLAG_ADAM_LEGACY = 5,
       /* OKB_FLAG_GRAD_GENERIC = 6 (default off): use the generic grad kernel even for the k = 1, kr = 0 batch (A/B runs). */
       OKB_FLAG_GRAD_GENERIC = 6,
       /* OKB_FLAG_DP_PULL = 7 (default off): owner-sharded data parallelism lets the row owner PULL the partial rows from its
        * peers' gradient buffers (peer loads) instead of running the reduce+push kernel (peer stores).  Bit-identical
        * results; measured 65 vs 40 us per step at 2 GPUs — dependent loads over NVLink cost several us each — so it is
        * kept for A/B runs only.  Needs the arena's optional plan / gradient slices (okb_dp.plan_steps > 0). */
       OKB_FLAG_DP_PULL = 7,
       /* OKB_FLAG_GRAD_SINGLE_WARP = 8 (default off): the generic grad kernel never splits a positive's negatives over several
        * warps (A/B runs; the split changes the association of the shared rows' gradient sums). */
       OKB_FLAG_GRAD_SINGLE_WARP = 8,
       /* OKB_FLAG_PLAN_MULTI = 9 (default off): plan a single step with the general multi-kernel segmented sort instead of the
          one-CTA single-kernel plan (A/B and test switch; results are identical). */
       OKB_FLAG_PLAN_MULTI = 9,
       /* OKB_FLAG_CHUNK_KERNEL = 10 (default off): okb_train_steps runs a chunk of planned steps as ONE persistent cooperative
          kernel (grad -> grid barrier -> update per step; csrc/chunk.cu) where the configuration is covered (TransE/H/D,
          k = 1 batch, single GPU, D in the vectorised layouts).  Bit-identical results; measured SLOWER than the per-phase
          kernels chained with programmatic dependent launch (18.0 vs 25.2 us per step on the bench workload), so it is
          kept for A/B runs only. */
       OKB_FLAG_CHUNK_KERNEL = 10,
       /* OKB_FLAG_ADAM_VPT = 11: vectors per thread (1..4) of the dense Adam pass: a CTA owns 256 * value consecutive vectors
          of one table and every thread has all its row-map entries and x / m / v vectors in flight before the first
          gradient load.  Bit-identical results for every value. */
       OKB_FLAG_ADAM_VPT = 11,
       /* OKB_FLAG_TRANSR_FUSED = 12 (default off): TransR's okb_grad runs a persistent kernel (one CTA per SM walks its
          relations, M_r of the next relation in flight by bulk-async copy while the current one is computed) and applies
          the RELATION-side update — rel_embeddings[r], transfer_matrix[r], their Adam slots — in place with the step's
          hyper-parameters; okb_update then only moves the entity rows and grad_rel is not written.  Measured on B200
          (FB15K shape, B = 4,831): 161 vs 163 us per step with SGD, 240 vs 195 with Adam — the step is bound by the
          per-relation chain of small dependent phases, not by M_r traffic — so the two-kernel form stays the default. */
       OKB_FLAG_TRANSR_FUSED = 12,
       /* OKB_FLAG_DP_TRACE = 13 (default off, debugging aid): the data-parallel grad / owner-update kernels stamp
          %globaltimer at their phase boundaries (first block in, grid dependency resolved, flags seen, hub blocks done,
          last block out) into a 64-step ring read by okb_debug_dp_trace. */
       OKB_FLAG_DP_TRACE = 13,
       /* OKB_FLAG_DP_HANDSHAKE = 14: how the data-parallel kernels exchange their epoch flags.  bit 0: waiters poll with
          relaxed loads and issue ONE acquire fence when the value has arrived (instead of an acquire per poll); bit 1: the
          announcing store is relaxed — the data it publishes was written by the PREVIOUS grid, whose completion
          (griddepcontrol.wait) has already made it visible. */
       OKB_FLAG_DP_HANDSHAKE = 14 };
int okb_set_flag(okb_ctx *c, int flag, INT value);

/* Optional per-kernel timing with CUDA events recorded on the launching stream, around:
 * id 0 sampler, 1 plan (keys + radix sort), 2 grad kernel, 3 update kernel (SGD / Adam), 4 rank kernel,
 * 5 rank preparation, 6 data-parallel reduce+push kernel, 7 data-parallel owner-update kernel.  okb_prof_read synchronises, returns the summed milliseconds and the number of
 * launches since the last read, and clears the list. */
int okb_prof_enable(okb_ctx *c, int on);
int okb_prof_read(okb_ctx *c, int id, double *total_ms, INT *count);
const char *okb_debug_cuda_error(void);
/* OKB_FLAG_DP_TRACE stamps: out = unsigned long long[64 * 16], row = epoch % 64 (all-ones = not stamped; slots 3, 4, 8, 9
 * hold the complement of the stamp); clears the ring. */
int okb_debug_dp_trace(okb_ctx *c, unsigned long long *out);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
INT okb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
