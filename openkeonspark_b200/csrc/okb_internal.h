// Internal declarations shared by the translation units of libokb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "../../include/okb200.h"

typedef int32_t i32;
typedef int64_t i64;
typedef uint64_t u64;

struct DevBuf {                       // grow-only device allocation
    void *p = nullptr;
    size_t cap = 0;
    bool external = false;            // memory owned by someone else (a slice of the peer arena): never freed, never grown
    int ensure(size_t bytes);
    void release();
    void adopt(void *ptr, size_t bytes) { release(); p = ptr; cap = bytes; external = true; }
    void disown() { if (external) { p = nullptr; cap = 0; external = false; } }
    template <class T> T *as() const { return (T *)p; }
};

// Sorted id lists keyed by relation or entity (type constraints, ontology): CSR, int32.
struct Lists {
    std::vector<i32> lef, rig, ids;   // ids[lef[k]..rig[k]) sorted
    i32 *d_lef = nullptr, *d_rig = nullptr, *d_ids = nullptr;
};

// One set of sampled batches + their plan.  The context works on the set held in its own fields; `alt` is a second set the
// next chunk is sampled and planned into on a side stream while the train kernels of the current chunk run.
struct PlanSlot {
    DevBuf batch, keys_ent, perm_ent, rowseg_e, sort_tmp, hist;
    i64 B = 0, K = 0, KR = 0, steps = 0, plan_ne = 0, plan_nr = 0, plan_lo = 0, plan_hi = 0, plan_b_lo = 0, plan_b_hi = 0;
    bool rowhead_ready = false;
};

struct okb_ctx {
    std::string in_path = "../data/FB15K/", out_path = "../data/FB15K/", err;
    i64 W = 1, bern = 0;
    // ---------------- totals
    i64 E = 0, R = 0, n_raw = 0, n = 0, new_batch = 0, n_test = 0, n_valid = 0, n_all = 0;
    i64 legacy_train_total = -1;      // Reader.h:227 re-reads trainTotal from the file header in importTestFiles
    // ---------------- host index (kept for the tiny host-side paths: TC negatives, thresholds)
    std::vector<i32> raw_h, raw_t, raw_r;              // train2id.txt order
    std::vector<i32> byh_r, byh_t, byt_r, byt_h, byht_t, byht_r;   // secondary key / value of the 3 sorted copies
    std::vector<i32> lef_h, rig_h, lef_t, rig_t, lef_ht, rig_ht;
    std::vector<float> tph, hpt;                       // left_mean / right_mean
    double max_rel_share = 0, max_ent_share = 0;       // largest fraction of train rows touching one relation / entity
    std::vector<i32> test_h, test_t, test_r, valid_h, valid_t, valid_r;   // sorted (r,h,t)
    std::vector<i32> test_lef, test_rig, valid_lef, valid_rig;
    std::vector<u64> all_hrt;                          // packed keys of train+valid+test sorted (h,r,t), dups kept
    std::vector<i32> neg_test_t, neg_valid_t;
    Lists head_type, tail_type, sup, sub;
    bool have_types = false, have_onto = false;
    // ---------------- device index
    int4 *d_raw = nullptr;            // {h, t, r, 0}
    int4 *d_run = nullptr;            // {llH, rrH, llT, rrT}: runs of (h,r) in by-head order and (t,r) in by-tail order
    int2 *d_run_ht = nullptr;         // {ll, rr}: run of (h,t) in by-(h,t) order
    i32 *d_byh_t = nullptr, *d_byt_h = nullptr, *d_byht_r = nullptr;
    float *d_prob = nullptr;          // per relation 1000*hpt/(hpt+tph) (bern) — Base.cpp:116-117
    // test side
    i32 *d_test_h = nullptr, *d_test_t = nullptr, *d_test_r = nullptr;
    i32 *d_known_t = nullptr, *d_known_h = nullptr;    // unique known tails per (h,r) run / heads per (t,r) run
    int4 *d_test_run = nullptr;       // {tail-list lo, hi, head-list lo, hi} into d_known_t / d_known_h
    std::vector<i32> grp_rel, grp_lo, grp_hi;          // relation groups of the test list
    // ---------------- sampler state
    std::vector<u64> state;           // host mirror (valid when !state_dirty)
    u64 *d_state = nullptr;
    bool state_dirty = false;         // device copy is newer than host mirror
    // ---------------- sampled batches
    i64 B = 0, K = 0, KR = 0, steps = 0;
    DevBuf batch;                     // int32 [steps][3][S]
    // ---------------- plan / workspace
    DevBuf keys_ent, keys_rel, perm_ent, perm_rel, sort_tmp, hist, gent, grel, flags, lossterms, rowseg_e, rowseg_r;
    DevBuf rank_ws, host_io, partial;
    i64 plan_ne = 0, plan_nr = 0, plan_lo = 0, plan_hi = 0;   // steps [plan_lo, plan_hi) of the sampled batches are planned
    i64 plan_b_lo = 0, plan_b_hi = 0;                          // ... for positives [plan_b_lo, plan_b_hi) of each step
    bool transr_tc = false;           // OKB_FLAG_TRANSR_TC: tensor-core candidate projection for TransR ranking
    bool loss_ctr_ready = false;
    const void *verify_h = nullptr;   // pending comparison of the caller's block (device alias) with the resident batch: rides in the next grad launch
    i64 verify_S = 0;
    // okb_sample_to_host, one-kernel plan: the int64 copy of the batch for the caller's page-locked block is written by extra
    // CTAs of the plan launch (plan_steps consumes mirror_dst), which then raise *host_flag (page-locked) for the waiting host
    long long *mirror_dst = nullptr;
    unsigned *host_flag = nullptr, *host_flag_dev = nullptr;          // page-locked word and its device alias
    bool defer_advance = false;       // sample_impl leaves the stream advance to its caller (okb_sample_to_host issues it after the plan)
    const void *spec_mirror = nullptr;// host block the last okb_sample_to_host filled while its batch is still resident and planned
    bool batch_from_host = false;     // the current batch came through okb_batch_from_host: update kernels honour the "bad id" flag
    bool chunk_kernel = false;        // OKB_FLAG_CHUNK_KERNEL: okb_train_steps runs a chunk as one persistent kernel where covered (measured slower: off)
    bool plan_multi = false;          // OKB_FLAG_PLAN_MULTI: one-step plans use the multi-kernel sort too
    bool plan_small_attr = false;     // dynamic shared memory limit of plan_small_kernel raised on this device
    bool grad_single_warp = false;    // OKB_FLAG_GRAD_SINGLE_WARP: never split a positive's negatives over several warps
    bool grad_generic = false;        // OKB_FLAG_GRAD_GENERIC: never use the k = 1 specialisation of the grad kernel
    int adam_vpt = 1;                 // vectors per thread of adam_tile_kernel (1..4; OKB200_ADAM_VPT for A/B runs)
    bool adam_legacy = false;         // OKB_FLAG_ADAM_LEGACY: grid-stride register kernel instead of the tile kernel
    bool adam_tma = false;            // OKB_FLAG_ADAM_TMA: TMA-staged single-wave Adam pass instead of the register-only one
    bool l2_prefetch = false;         // OKB_FLAG_L2_PREFETCH: grad kernel prefetches the Adam state into L2
    bool transr_fused = false;        // OKB_FLAG_TRANSR_FUSED: persistent TransR kernel with the relation update applied in place (measured: no faster; off)
    bool transr_rel_done = false;     // the last okb_grad (TransR) already applied the relation-side update
    i64 tr_lo = 0, tr_hi = 0;         // TransR relation shard [tr_lo, tr_hi) (okb_transr_set_shard); empty = all relations
    okb_dp dp = {};                   // owner-sharded data parallelism (okb_dp_attach)
    bool dp_on = false;
    unsigned long long dp_epoch = 0;
    DevBuf dp_trace;                  // OKB_FLAG_DP_TRACE: [64 steps][16] globaltimer stamps of the data-parallel kernels (debugging aid)
    bool dp_trace_on = false;
    int dp_hs_mode = 2;               // OKB_FLAG_DP_HANDSHAKE: see GradArgs::hs_mode (2 measured fastest: 42.4 -> 39.2 us per step at 2 GPUs)
    int sc_launch = 0;                // set around launch_grad by the scatter (1) / gather (2) form of the data-parallel step
    long long sc_buf_off[3] = {0, 0, 0};   // byte offsets of the receive buffers in use (gather form: alternating halves)
    unsigned long long hub_done = 0;  // value OKB_FLAGS_HUBCTR will have reached when every launch issued so far has finished
    bool dp_pull = false;             // OKB_FLAG_DP_PULL: row owners pull partial rows from their peers instead of the reduce+push kernel
    PlanSlot alt;                     // prefetched chunk (okb_chunk_prefetch)
    bool alt_ready = false, in_prefetch = false;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_main = nullptr, ev_side = nullptr, ev_sampled = nullptr;
    size_t saved_n = 0;
    u64 *d_state_saved = nullptr;     // RNG streams as they were before the prefetched chunk was sampled
    bool pdl = true;                  // programmatic dependent launch between the grad and update kernels
    bool rowhead_ready = false;       // Adam: per-step row -> first sorted position map built for the planned chunk
    int ent_bits = 0, rel_bits = 0;
    // ---------------- optional per-kernel timing (CUDA events on the launching stream; bench.py)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev[8];   // begin/end pairs per kernel id
    // legacy result buffers
    i64 res8[8];
    std::vector<i64> tpfp;
    // ---------------- per-context device facts / function attributes (never process-global: one context per device)
    int sm_count = 0;                 // multiProcessorCount of the context's device (okb_sms)
    std::map<const void *, size_t> smem_attr;   // largest dynamic shared memory opted into per kernel on this device
    DevBuf legacy_scores, legacy_out; // testHead / testTail staging of the reference-compatible layer
    DevBuf tc_ranges, tc_io;          // triple classification: per-relation valid / test ranges + counters; score staging of the host entry points
    bool tc_ranges_ready = false;
};

// number of SMs of the current device (grids are sized in multiples of it)
static inline int okb_sms(okb_ctx *c) {
    if (c->sm_count <= 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            c->sm_count = n;
        else { cudaGetLastError(); return 148; }
    }
    return c->sm_count;
}
// opt a kernel into `bytes` of dynamic shared memory once per context (cudaFuncSetAttribute is a per-device setting)
template <class F> static inline cudaError_t okb_smem_optin(okb_ctx *c, F *fn, size_t bytes) {
    size_t &have = c->smem_attr[(const void *)fn];
    if (bytes <= have) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

#define OKB_FAIL(c, code, msg) do { (c)->err = (msg); return (code); } while (0)
#define OKB_CUDA(c, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
    (c)->err = std::string(#expr) + ": " + cudaGetErrorString(e_); return OKB_ERR_CUDA; } } while (0)

#ifdef __CUDACC__
// One element of the TF1 Adam rule, spelled with explicitly rounded operations so that EVERY kernel that applies it
// (tile / legacy / TMA-staged / persistent-chunk / data-parallel owner / TransR relation rows) produces the same bits:
//   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; x <- x - lr_t m / (sqrt(v) + eps)
__device__ __forceinline__ void adam_elem(float &x, float &m, float &v, float g, float b1, float b2, float c1, float c2, float lr, float eps) {
    const float mq = __fmaf_rn(m, b1, __fmul_rn(g, c1));
    const float vq = __fmaf_rn(v, b2, __fmul_rn(__fmul_rn(g, g), c2));
    m = mq; v = vq;
    x = __fsub_rn(x, __fdiv_rn(__fmul_rn(lr, mq), __fadd_rn(__fsqrt_rn(vq), eps)));
}
#endif

// kernel ids for okb_prof_*
enum { PROF_SAMPLE = 0, PROF_PLAN = 1, PROF_GRAD = 2, PROF_UPDATE = 3, PROF_RANK = 4, PROF_RANK_PREP = 5, PROF_DP_PUSH = 6, PROF_DP_OWNER = 7 };
static inline void prof_mark(okb_ctx *c, int id, cudaStream_t s) {
    if (!c->prof_on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    c->prof_ev[id].push_back(e);
}
struct ProfScope {                     // records an event pair around the launches in its scope
    okb_ctx *c; int id; cudaStream_t s;
    ProfScope(okb_ctx *c_, int id_, cudaStream_t s_) : c(c_), id(id_), s(s_) { prof_mark(c, id, s); }
    ~ProfScope() { prof_mark(c, id, s); }
};

extern std::atomic<long long> g_launches;   // kernels launched by this library (all contexts, all threads)
#define OKB_LAUNCHED(n) (g_launches.fetch_add((n), std::memory_order_relaxed))

// loader.cpp
int okb_upload_train(okb_ctx *c);
int okb_upload_test(okb_ctx *c);
int okb_upload_lists(okb_ctx *c);
bool okb_host_find(const okb_ctx *c, i64 h, i64 t, i64 r);
i64 okb_host_new_tail(okb_ctx *c, i64 h, i64 r);      // Corrupt.h corrupt_head(0, h, r) on stream 0

// okb_ctx::flags, in 4-byte words: [0,64) partial loss sums | [64] loss ticket | [68] "bad id" flag of the host-batch path
//   | [72] arrival counter of the persistent chunk kernel's grid barrier
#define OKB_FLAGS_BYTES (sizeof(float) * 64 + 128)
#define OKB_FLAGS_BAD 68
#define OKB_FLAGS_GRIDBAR 72
#define OKB_FLAGS_MIRRORCTR 78  // mirror CTAs of the one-step plan launch that have finished
#define OKB_FLAGS_HUBCTR 76    // u64 (byte 304): hub blocks finished since the context was created (scatter-form owner update)
int okb_ensure_flags(okb_ctx *c, cudaStream_t s);     // train.cu: allocate + zero once

extern "C" int okb_verify_flush(okb_ctx *c, void *stream);       // sampler.cu: launch a pending comparison stand-alone
bool okb_plan_small_ok(const okb_ctx *c, INT B, INT k, INT kr);                     // train.cu: a one-step plan of this batch is ONE kernel
extern "C" int okb_batch_verify_host(okb_ctx *c, INT B, INT k, INT kr, const INT *h, const INT *t, const INT *r, void *stream);   // sampler.cu

// train.cu: forget a prefetched chunk (restores the RNG streams); every entry point that touches the streams calls it
int okb_discard_prefetch(okb_ctx *c);

// radix.cu
int okb_sort_pairs(okb_ctx *c, const i32 *keys, i32 *keys_out, i32 *perm_out, i64 n, int bits, cudaStream_t s);
int okb_sort_pairs_seg(okb_ctx *c, const i32 *keys, i32 *keys_out, i32 *perm_out, i64 seg, i64 nseg, int bits, cudaStream_t s);

static inline int bits_for(i64 n) { int b = 1; while ((1ll << b) < n) b++; return b; }
