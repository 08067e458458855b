// Context lifecycle and the reference-compatible layer: the symbols of the reference's
// release/Base.so (Base.cpp:10-47 + Setting.h/Random.h/Reader.h/Test.h extern "C" blocks) on a
// process-global default context, with HOST pointers, so /root/reference/Config.py:30-51 binds to
// this library unchanged.  Every compute call lands on the CUDA kernels; nothing here falls back
// to a CPU implementation of the hot path.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <algorithm>

#include "okb_internal.h"


static okb_ctx *g_ctx = nullptr;

// The reference ABI has void returns and reports problems by printing and returning (Reader.h:36-39: a missing file
// prints a message, the totals stay 0).  Same here: the message goes to stderr AND stays readable through
// okb_last_error(okb_default_ctx()); nothing is computed on the CPU instead.  OKB200_ABORT_ON_ERROR=1 turns every
// failure into abort() for callers that would rather die than continue with unfilled output buffers.
static void die(okb_ctx *c, const char *where) {
    fprintf(stderr, "libokb200: %s failed: %s\n", where, c->err.c_str());
    const char *e = getenv("OKB200_ABORT_ON_ERROR");
    if (e && e[0] == '1') abort();
}
#define MUST(call, where) do { if ((call) != 0) { die(g_ctx_get(), where); return; } } while (0)
#define MUSTV(call, where, val) do { if ((call) != 0) { die(g_ctx_get(), where); return (val); } } while (0)

static okb_ctx *g_ctx_get() {
    if (!g_ctx) g_ctx = new okb_ctx();
    return g_ctx;
}

extern "C" {

int okb_version(void) { return 100; }
okb_ctx *okb_default_ctx(void) { return g_ctx_get(); }
int okb_create(okb_ctx **out) {
    if (!out) return OKB_ERR_ARG;
    *out = new okb_ctx();
    return 0;
}
int okb_destroy(okb_ctx *c) {
    if (!c) return 0;
    void *ptrs[] = {c->d_raw, c->d_run, c->d_run_ht, c->d_byh_t, c->d_byt_h, c->d_byht_r, c->d_prob, c->d_test_h, c->d_test_t,
                    c->d_test_r, c->d_known_t, c->d_known_h, c->d_test_run, c->d_state};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (Lists *L : {&c->head_type, &c->tail_type, &c->sup, &c->sub}) {
        if (L->d_lef) cudaFree(L->d_lef);
        if (L->d_rig) cudaFree(L->d_rig);
        if (L->d_ids) cudaFree(L->d_ids);
    }
    for (DevBuf *b : {&c->batch, &c->keys_ent, &c->keys_rel, &c->perm_ent, &c->perm_rel, &c->sort_tmp, &c->hist, &c->gent, &c->grel,
                      &c->flags, &c->lossterms, &c->rowseg_e, &c->rowseg_r, &c->rank_ws, &c->host_io, &c->partial, &c->legacy_scores,
                      &c->legacy_out})
        b->release();
    for (DevBuf *b : {&c->alt.batch, &c->alt.keys_ent, &c->alt.perm_ent, &c->alt.rowseg_e, &c->alt.sort_tmp, &c->alt.hist}) b->release();
    if (c->d_state_saved) cudaFree(c->d_state_saved);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_side) cudaEventDestroy(c->ev_side);
    if (c->ev_sampled) cudaEventDestroy(c->ev_sampled);
    if (c->host_flag) cudaFreeHost(c->host_flag);
    if (c == g_ctx) g_ctx = nullptr;
    delete c;
    return 0;
}
const char *okb_last_error(okb_ctx *c) { return c ? c->err.c_str() : "null context"; }
int okb_set_device(okb_ctx *c, int device) { OKB_CUDA(c, cudaSetDevice(device)); return 0; }
int okb_set_flag(okb_ctx *c, int flag, INT value) {
    if (flag == OKB_FLAG_TRANSR_TC) { c->transr_tc = value != 0; return 0; }
    if (flag == OKB_FLAG_PDL) { c->pdl = value != 0; return 0; }
    if (flag == OKB_FLAG_ADAM_TMA) { c->adam_tma = value != 0; return 0; }
    if (flag == OKB_FLAG_L2_PREFETCH) { c->l2_prefetch = value != 0; return 0; }
    if (flag == OKB_FLAG_ADAM_LEGACY) { c->adam_legacy = value != 0; return 0; }
    if (flag == OKB_FLAG_GRAD_GENERIC) { c->grad_generic = value != 0; return 0; }
    if (flag == OKB_FLAG_GRAD_SINGLE_WARP) { c->grad_single_warp = value != 0; return 0; }
    if (flag == OKB_FLAG_PLAN_MULTI) { c->plan_multi = value != 0; return 0; }
    if (flag == OKB_FLAG_DP_PULL) { c->dp_pull = value != 0; return 0; }
    if (flag == OKB_FLAG_CHUNK_KERNEL) { c->chunk_kernel = value != 0; return 0; }
    if (flag == OKB_FLAG_TRANSR_FUSED) { c->transr_fused = value != 0; return 0; }
    if (flag == OKB_FLAG_DP_HANDSHAKE) { c->dp_hs_mode = (int)value & 3; return 0; }
    if (flag == OKB_FLAG_DP_TRACE) {
        c->dp_trace_on = value != 0;
        if (c->dp_trace_on) {
            if (c->dp_trace.ensure(64 * 16 * 8)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory");
            OKB_CUDA(c, cudaMemset(c->dp_trace.p, 0xff, 64 * 16 * 8));
        }
        return 0;
    }
    if (flag == OKB_FLAG_ADAM_VPT) { if (value < 1 || value > 4) OKB_FAIL(c, OKB_ERR_ARG, "vectors per thread: 1..4"); c->adam_vpt = (int)value; return 0; }
    OKB_FAIL(c, OKB_ERR_ARG, "unknown flag");
}
int okb_prof_enable(okb_ctx *c, int on) { c->prof_on = on != 0; return 0; }
int okb_prof_read(okb_ctx *c, int id, double *total_ms, INT *count) {
    if (id < 0 || id >= 8) OKB_FAIL(c, OKB_ERR_ARG, "bad kernel id");
    std::vector<cudaEvent_t> &v = c->prof_ev[id];
    double tot = 0;
    INT n = 0;
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
        OKB_CUDA(c, cudaEventSynchronize(v[i + 1]));
        float ms = 0;
        OKB_CUDA(c, cudaEventElapsedTime(&ms, v[i], v[i + 1]));
        tot += ms; n++;
    }
    for (cudaEvent_t e : v) cudaEventDestroy(e);
    v.clear();
    if (total_ms) *total_ms = tot;
    if (count) *count = n;
    return 0;
}
// debugging aid: the stamps of OKB_FLAG_DP_TRACE (64 steps x 16 u64, indexed by epoch % 64), then cleared
int okb_debug_dp_trace(okb_ctx *c, unsigned long long *out) {
    if (!c->dp_trace.p) OKB_FAIL(c, OKB_ERR_STATE, "set OKB_FLAG_DP_TRACE first");
    OKB_CUDA(c, cudaDeviceSynchronize());
    OKB_CUDA(c, cudaMemcpy(out, c->dp_trace.p, 64 * 16 * 8, cudaMemcpyDeviceToHost));
    OKB_CUDA(c, cudaMemset(c->dp_trace.p, 0xff, 64 * 16 * 8));
    return 0;
}
// debugging aid: returns (and clears) the CUDA runtime's last error of this library's runtime instance
const char *okb_debug_cuda_error(void) { return cudaGetErrorString(cudaGetLastError()); }

// ------------------------------------------------------------------ reference-compatible layer
void setInPath(char *path) {
    okb_set_in_path(g_ctx_get(), path);
    // Setting.h:13-19 copies the string verbatim (no trailing slash added): keep that spelling
    g_ctx_get()->in_path = path;
    printf("Input Files Path : %s\n", path);
}
void setOutPath(char *path) { g_ctx_get()->out_path = path; printf("Output Files Path : %s\n", path); }
void setWorkThreads(INT threads) { MUST(okb_set_work_threads(g_ctx_get(), threads), "setWorkThreads"); }
INT getWorkThreads(void) { return g_ctx_get()->W; }
void setBern(INT con) { okb_set_bern(g_ctx_get(), con); }
INT getEntityTotal(void) { return g_ctx_get()->E; }
INT getRelationTotal(void) { return g_ctx_get()->R; }
INT getTripleTotal(void) { return g_ctx_get()->n_all; }
INT getTrainTotal(void) { okb_ctx *c = g_ctx_get(); return c->legacy_train_total >= 0 ? c->legacy_train_total : c->n; }
INT getTrainTotal_(void) { return g_ctx_get()->n_raw; }
INT getBatchTotal(void) { return g_ctx_get()->new_batch; }
INT getTestTotal(void) { return g_ctx_get()->n_test; }
INT getValidTotal(void) { return g_ctx_get()->n_valid; }
void randReset(void) { MUST(okb_rand_reset(g_ctx_get()), "randReset"); }
// Reader.h:36-39: a missing file prints a message and returns, leaving the totals at 0.
void importTrainFiles(void) { int rc = okb_import_train_files(g_ctx_get()); g_ctx_get()->legacy_train_total = -1; if (rc && rc != OKB_ERR_IO) die(g_ctx_get(), "importTrainFiles"); }
void importTestFiles(void) { int rc = okb_import_test_files(g_ctx_get()); if (!rc) g_ctx_get()->legacy_train_total = g_ctx_get()->n_raw; if (rc && rc != OKB_ERR_IO) die(g_ctx_get(), "importTestFiles"); }
void importTypeFiles(void) { int rc = okb_import_type_files(g_ctx_get()); if (rc && rc != OKB_ERR_IO) die(g_ctx_get(), "importTypeFiles"); }
void importOntologyFiles(void) { int rc = okb_import_ontology_files(g_ctx_get()); if (rc && rc != OKB_ERR_IO) die(g_ctx_get(), "importOntologyFiles"); }

void sampling(INT *bh, INT *bt, INT *br, REAL *by, INT B, INT k, INT kr) {
    okb_ctx *c = g_ctx_get();
    MUST(okb_sample(c, B, k, kr, 1, 0, c->W, nullptr), "sampling");
    MUST(okb_batch_to_host(c, 0, bh, bt, br, by, nullptr), "sampling");
}
static bool legacy_index_ok(okb_ctx *c, INT index, const char *where) {
    if (index >= 0 && index < c->n_test) return true;
    c->err = "test index out of range (import the test files first)";
    die(c, where);
    return false;
}

void getHeadBatch(INT index, INT *ph, INT *pt, INT *pr) {              // Test.h:11-17 (candidate fill; pure host)
    okb_ctx *c = g_ctx_get();
    if (!legacy_index_ok(c, index, "getHeadBatch")) return;
    for (INT i = 0; i < c->E; i++) { ph[i] = i; pt[i] = c->test_t[index]; pr[i] = c->test_r[index]; }
}
void getTailBatch(INT index, INT *ph, INT *pt, INT *pr) {              // Test.h:20-26
    okb_ctx *c = g_ctx_get();
    if (!legacy_index_ok(c, index, "getTailBatch")) return;
    for (INT i = 0; i < c->E; i++) { ph[i] = c->test_h[index]; pt[i] = i; pr[i] = c->test_r[index]; }
}
static INT *rank_host(INT index, REAL *con, int side) {
    okb_ctx *c = g_ctx_get();
    DevBuf &dscores = c->legacy_scores, &dout = c->legacy_out;       // per context, not per process
    for (int i = 0; i < 8; i++) c->res8[i] = 0;
    if (dscores.ensure(sizeof(float) * std::max<i64>(c->E, 1)) || dout.ensure(sizeof(i64) * 8)) { c->err = "out of device memory"; die(c, "testHead/testTail"); return c->res8; }
    if (cudaMemcpy(dscores.p, con, sizeof(float) * c->E, cudaMemcpyHostToDevice) != cudaSuccess) { c->err = "H2D copy failed"; die(c, "testHead/testTail"); return c->res8; }
    MUSTV(okb_rank_scores(c, index, side, dscores.as<float>(), dout.as<i64>(), nullptr), "testHead/testTail", c->res8);
    if (cudaMemcpy(c->res8, dout.p, sizeof(i64) * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { c->err = "D2H copy failed"; die(c, "testHead/testTail"); }
    return c->res8;
}
INT *testHead(INT index, REAL *con) { return rank_host(index, con, 0); }
INT *testTail(INT index, REAL *con) { return rank_host(index, con, 1); }

void getNegTest(void) { MUST(okb_tc_batch(g_ctx_get(), 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), "getNegTest"); }
void getNegValid(void) { MUST(okb_tc_batch(g_ctx_get(), 1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), "getNegValid"); }
void getTestBatch(INT *ph, INT *pt, INT *pr, INT *nh, INT *nt, INT *nr) { MUST(okb_tc_batch(g_ctx_get(), 0, ph, pt, pr, nh, nt, nr), "getTestBatch"); }
void getValidBatch(INT *ph, INT *pt, INT *pr, INT *nh, INT *nt, INT *nr) { MUST(okb_tc_batch(g_ctx_get(), 1, ph, pt, pr, nh, nt, nr), "getValidBatch"); }
void getBestThreshold(REAL *relThresh, REAL *score_pos, REAL *score_neg) { MUST(okb_best_threshold(g_ctx_get(), relThresh, score_pos, score_neg), "getBestThreshold"); }
void test_triple_classification(REAL *relThresh, REAL *score_pos, REAL *score_neg, REAL *acc_addr) {
    INT cnt[4];
    REAL acc;
    MUST(okb_tc_eval(g_ctx_get(), relThresh, score_pos, score_neg, cnt, &acc), "test_triple_classification");
    const double TP = cnt[0], TN = cnt[1], FP = cnt[2], FN = cnt[3];
    const double precision = TP / (TP + FP), recall = TP / (TP + FN);
    printf("triple classification accuracy is %lf\n", (TP + TN) / (TP + TN + FP + FN));      // Test.h:381-384
    printf("triple classification precision is %lf\n", precision);
    printf("triple classification recall is %lf\n", recall);
    printf("triple classification f-measure is %lf\n", (2 * precision * recall) / (precision + recall));
    if (acc_addr) acc_addr[0] = acc;
}
INT get_n_interval(INT r, REAL *score_pos, REAL *score_neg) { return okb_n_interval(g_ctx_get(), r, score_pos, score_neg); }
INT *get_TPFP(INT r, REAL *score_pos, REAL *score_neg, REAL *score_pos_test, REAL *score_neg_test) {
    return okb_tpfp(g_ctx_get(), r, score_pos, score_neg, score_pos_test, score_neg_test);
}

}  // extern "C"
