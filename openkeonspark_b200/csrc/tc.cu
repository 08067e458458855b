// Triple classification on the device: per-relation threshold search and TP / TN / FP / FN counts
// (base/Test.h:303-341 getBestThreshold, :345-387 test_triple_classification), fed by the predict kernels (score.cu).
//
// The reference scans, for every relation with valid triples, the grid  min + i * 0.01  (i = 0 .. (max - min) / 0.01) and
// keeps the FIRST threshold with the best accuracy (#pos <= th + #neg > th) / total.  Here one CTA owns one relation:
//   * min / max of its valid scores by a comparison-only tree (exact in any order),
//   * every thread evaluates thresholds i = tid, tid + 128, ... against the relation's scores (staged in shared memory
//     when they fit), with the reference's arithmetic: th = fl(min + fl(float(i) * 0.01f)), acc = float(double(correct) /
//     double(total)),
//   * the winner is the maximum of the packed key (accuracy bits << 32 | ~i): best accuracy, smallest i — "first best".
// Results are bit-identical to the reference library (tests/test_gpu_api.py compares with the library built from
// /root/reference/base/Base.cpp).
#include <algorithm>

#include "okb_internal.h"

#define TC_THREADS 128
#define TC_STAGE 4096          // valid triples of one relation staged in shared memory (pos + neg: 32 KB)

struct TcArgs {
    const float *pos, *neg;
    const i32 *v_lef, *v_rig, *t_lef, *t_rig;
    float *thresh;
    unsigned long long *counts;      // [4] TP, TN, FP, FN
    i32 R, on_valid;
};

__device__ __forceinline__ float tc_sel_min(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float tc_sel_max(float a, float b) { return b > a ? b : a; }

__global__ void __launch_bounds__(TC_THREADS) tc_threshold_kernel(TcArgs a) {
    __shared__ float s_pos[TC_STAGE], s_neg[TC_STAGE];
    __shared__ float s_mn[TC_THREADS / 32], s_mx[TC_THREADS / 32];
    __shared__ unsigned long long s_key[TC_THREADS / 32];
    const i32 r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const i32 lo = a.v_lef[r];
    if (lo == -1) return;                                  // Test.h:310: no valid triples, threshold untouched
    const i32 n = a.v_rig[r] - lo + 1;
    const bool staged = n <= TC_STAGE;
    float mn = a.pos[lo], mx = mn;                         // Test.h:312-322
    for (i32 j = tid; j < n; j += TC_THREADS) {
        const float p = a.pos[lo + j], q = a.neg[lo + j];
        if (staged) { s_pos[j] = p; s_neg[j] = q; }
        mn = tc_sel_min(tc_sel_min(mn, p), q);
        mx = tc_sel_max(tc_sel_max(mx, p), q);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = tc_sel_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = tc_sel_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_mn[w] = mn; s_mx[w] = mx; }
    __syncthreads();
    mn = s_mn[0]; mx = s_mx[0];
#pragma unroll
    for (int q = 1; q < TC_THREADS / 32; q++) { mn = tc_sel_min(mn, s_mn[q]); mx = tc_sel_max(mx, s_mx[q]); }
    const long long n_int = (long long)__fdiv_rn(__fsub_rn(mx, mn), 0.01f);      // INT((max - min) / interval), Setting.h:118
    const double total = (double)(2 * (long long)n);
    const float *P = staged ? s_pos : a.pos + lo, *Q = staged ? s_neg : a.neg + lo;
    unsigned long long best = 0ull;
    for (long long i = tid; i <= n_int; i += TC_THREADS) {
        const float th = __fadd_rn(mn, __fmul_rn((float)i, 0.01f));              // Test.h:326
        long long ok = 0;
        for (i32 j = 0; j < n; j++) ok += (P[j] <= th) + (Q[j] > th);
        const float acc = (float)((double)ok / total);                           // Test.h:332: 1.0 * correct / total -> REAL
        const unsigned long long key = ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
        best = key > best ? key : best;                    // accuracy >= 0: its bit pattern orders like the value
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o); best = y > best ? y : best; }
    if (lane == 0) s_key[w] = best;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int q = 1; q < TC_THREADS / 32; q++) best = s_key[q] > best ? s_key[q] : best;
        const unsigned i = 0xffffffffu - (unsigned)(best & 0xffffffffull);
        a.thresh[r] = __fadd_rn(mn, __fmul_rn((float)(long long)i, 0.01f));
    }
}

// TP / FN over the positives, TN / FP over the negatives of every relation's test (or valid) range (Test.h:352-370)
__global__ void __launch_bounds__(TC_THREADS) tc_count_kernel(TcArgs a) {
    __shared__ unsigned s_cnt[4];
    const i32 r = blockIdx.x, tid = threadIdx.x;
    const i32 *lef = a.on_valid ? a.v_lef : a.t_lef, *rig = a.on_valid ? a.v_rig : a.t_rig;
    if (a.v_lef[r] == -1 || lef[r] == -1) return;          // Test.h:353
    if (tid < 4) s_cnt[tid] = 0u;
    __syncthreads();
    const float th = a.thresh[r];
    unsigned tp = 0, tn = 0, fp = 0, fn = 0;
    for (i32 i = lef[r] + tid; i <= rig[r]; i += TC_THREADS) {
        if (a.pos[i] <= th) tp++; else fn++;
        if (a.neg[i] > th) tn++; else fp++;
    }
    if (tp) atomicAdd(s_cnt + 0, tp);
    if (tn) atomicAdd(s_cnt + 1, tn);
    if (fp) atomicAdd(s_cnt + 2, fp);
    if (fn) atomicAdd(s_cnt + 3, fn);
    __syncthreads();
    if (tid < 4 && s_cnt[tid]) atomicAdd(a.counts + tid, (unsigned long long)s_cnt[tid]);
}

static int tc_ranges(okb_ctx *c, TcArgs &a, cudaStream_t s) {
    if (c->valid_lef.empty()) OKB_FAIL(c, OKB_ERR_STATE, "import test files first");
    const size_t R = (size_t)c->R;
    if (!c->tc_ranges_ready) {
        if (c->tc_ranges.ensure(sizeof(i32) * 4 * R + 64)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory");
        i32 *d = c->tc_ranges.as<i32>();
        OKB_CUDA(c, cudaMemcpyAsync(d, c->valid_lef.data(), sizeof(i32) * R, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaMemcpyAsync(d + R, c->valid_rig.data(), sizeof(i32) * R, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaMemcpyAsync(d + 2 * R, c->test_lef.data(), sizeof(i32) * R, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaMemcpyAsync(d + 3 * R, c->test_rig.data(), sizeof(i32) * R, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaStreamSynchronize(s));             // pageable host vectors: staged before they can change
        c->tc_ranges_ready = true;
    }
    const i32 *d = c->tc_ranges.as<i32>();
    a.v_lef = d; a.v_rig = d + R; a.t_lef = d + 2 * R; a.t_rig = d + 3 * R;
    a.counts = (unsigned long long *)(c->tc_ranges.as<char>() + sizeof(i32) * 4 * R + ((8 - (sizeof(i32) * 4 * R) % 8) % 8));
    a.R = (i32)c->R;
    return 0;
}

extern "C" {

// getBestThreshold (Test.h:303-341) with DEVICE score arrays (index-aligned with the (r,h,t)-sorted valid list) and a
// device threshold array [R] (entries of relations without valid triples are left untouched, like the reference).
int okb_tc_thresholds_dev(okb_ctx *c, const float *pos, const float *neg, float *thresh, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    TcArgs a = {};
    int rc = tc_ranges(c, a, s);
    if (rc) return rc;
    if (c->n_valid == 0) return 0;
    a.pos = pos; a.neg = neg; a.thresh = thresh;
    tc_threshold_kernel<<<(unsigned)c->R, TC_THREADS, 0, s>>>(a);
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

// test_triple_classification's counts (Test.h:345-387): TP, TN, FP, FN over the TEST ranges (on_valid = 0) or over the
// valid ranges (on_valid = 1: the early-stop check, see okb_tc_eval_valid) -> host INT[4]; acc = (TP + TN) / all.
int okb_tc_counts_dev(okb_ctx *c, const float *thresh, const float *pos, const float *neg, int on_valid, INT *tp_tn_fp_fn,
                      REAL *acc, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    TcArgs a = {};
    int rc = tc_ranges(c, a, s);
    if (rc) return rc;
    a.pos = pos; a.neg = neg; a.thresh = const_cast<float *>(thresh); a.on_valid = on_valid ? 1 : 0;
    OKB_CUDA(c, cudaMemsetAsync(a.counts, 0, 4 * sizeof(unsigned long long), s));
    tc_count_kernel<<<(unsigned)c->R, TC_THREADS, 0, s>>>(a);
    OKB_LAUNCHED(1);
    unsigned long long h[4];
    OKB_CUDA(c, cudaMemcpyAsync(h, a.counts, sizeof(h), cudaMemcpyDeviceToHost, s));
    OKB_CUDA(c, cudaStreamSynchronize(s));
    if (tp_tn_fp_fn) for (int i = 0; i < 4; i++) tp_tn_fp_fn[i] = (INT)h[i];
    if (acc) acc[0] = (REAL)(1.0 * (double)(h[0] + h[1]) / (double)(h[0] + h[1] + h[2] + h[3]));
    return 0;
}

}  // extern "C"
