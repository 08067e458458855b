// GPU negative sampler: the reference's getBatch / corrupt_head / corrupt_tail / corrupt_rel
// (base/Base.cpp:74-143, base/Corrupt.h:7-101) with the per-thread LCG streams of base/Random.h.
//
// The reference walks each stream sequentially.  Every batch slot consumes a FIXED number of
// draws (1 to pick the row, 2 per entity negative, 1 per relation negative — no rejection loops),
// so slot `b` of stream `id` in call `c` starts at draw offset
//     (c * slice_len(id) + (b - slice_lo(id))) * (1 + 2k + kr)
// and an O(log n) LCG jump-ahead puts an independent GPU thread exactly there.  One launch
// therefore produces `steps` consecutive sampling() calls, bit-identical to the reference.
#include <algorithm>
#include <cstdlib>

#include "okb_internal.h"

std::atomic<long long> g_launches{0};

#define LCG_A 25214903917ULL
#define LCG_C 11ULL

__host__ __device__ static inline void lcg_jump(u64 n, u64 &mul, u64 &add) {
    // state_{i+n} = mul * state_i + add  (mod 2^64)
    u64 a = LCG_A, c = LCG_C;
    mul = 1; add = 0;
    while (n) {
        if (n & 1) { mul = mul * a; add = add * a + c; }
        c = (a + 1) * c;
        a = a * a;
        n >>= 1;
    }
}
__host__ __device__ static inline u64 lcg_next(u64 &s) { s = s * LCG_A + LCG_C; return s; }

struct SampleArgs {
    const int4 *raw, *run;
    const int2 *run_ht;
    const i32 *byh_t, *byt_h, *byht_r;
    const float *prob;
    const u64 *state;
    i32 *out;            // [steps][3][S]
    i64 *mh, *mt, *mr;   // optional mirror of step 0 in page-locked HOST memory (int64, the reference's batch_h/t/r): the batch
                         // crosses PCIe as posted stores while the kernel is still sampling (okb_sample_to_host)
    i64 n_raw, new_batch;
    i32 E, R, B, k, kr, W, per, bern, steps, stream_lo, stream_hi;
};

// k-th id (0-based, k = tmp) that is NOT among the sorted values vals[ll..rr]  (Corrupt.h:25-36)
__device__ __forceinline__ i32 kth_absent(const i32 *__restrict__ vals, i32 ll, i32 rr, i32 tmp) {
    if (tmp < __ldg(vals + ll)) return tmp;
    if (tmp > __ldg(vals + rr) - rr + ll - 1) return tmp + rr - ll + 1;
    i32 lef = ll, rig = rr + 1;
    while (lef + 1 < rig) {
        i32 mid = (lef + rig) >> 1;
        if (__ldg(vals + mid) - mid + ll - 1 < tmp) lef = mid; else rig = mid;
    }
    return tmp + lef - ll + 1;
}

__global__ void __launch_bounds__(128) sample_kernel(SampleArgs a) {
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (i64)a.steps * a.B) return;
    const i32 step = (i32)(tid / a.B), b = (i32)(tid % a.B);
    const i32 id = b / a.per;                               // Base.cpp:85-92 slice geometry
    if (id < a.stream_lo || id >= a.stream_hi) return;
    const i32 lo = id * a.per;
    i32 hi = lo + a.per; if (hi > a.B) hi = a.B;
    const u64 draws = 1 + 2 * (u64)a.k + (u64)a.kr;
    u64 mul, add;
    lcg_jump(((u64)step * (u64)(hi - lo) + (u64)(b - lo)) * draws, mul, add);
    u64 s = mul * a.state[id] + add;

    const i32 S = a.B * (1 + a.k + a.kr);
    i32 *oh = a.out + (i64)step * 3 * S, *ot = oh + S, *orl = ot + S;

    i64 row;
    if (a.new_batch > 0) row = (i64)(lcg_next(s) % (u64)a.new_batch) + (a.n_raw - a.new_batch);   // Base.cpp:101-103
    else row = (i64)(lcg_next(s) % (u64)a.n_raw);
    const int4 p = __ldg(a.raw + row);                      // {h, t, r, -}
    const int4 rn = __ldg(a.run + row);                     // {llH, rrH, llT, rrT}
    const bool mir = a.mh != nullptr;
    auto put = [&](i32 at, i32 h, i32 t, i32 r) {
        oh[at] = h; ot[at] = t; orl[at] = r;
        if (mir) { a.mh[at] = h; a.mt[at] = t; a.mr[at] = r; }
    };
    put(b, p.x, p.y, p.z);
    const float prob = a.bern ? __ldg(a.prob + p.z) : 500.0f;
    i32 at = b + a.B;
    for (i32 m = 0; m < a.k; m++, at += a.B) {
        const u64 coin = lcg_next(s) % 1000ULL;
        const u64 d = lcg_next(s);
        if ((float)coin < prob) {                           // keep (h,r): new tail  (Base.cpp:118-121)
            const i32 tmp = (i32)(d % (u64)(a.E - (rn.y - rn.x + 1)));
            put(at, p.x, kth_absent(a.byh_t, rn.x, rn.y, tmp), p.z);
        } else {                                            // keep (t,r): new head  (Base.cpp:122-126)
            const i32 tmp = (i32)(d % (u64)(a.E - (rn.w - rn.z + 1)));
            put(at, kth_absent(a.byt_h, rn.z, rn.w, tmp), p.y, p.z);
        }
    }
    if (a.kr > 0) {
        const int2 rh = __ldg(a.run_ht + row);
        for (i32 m = 0; m < a.kr; m++, at += a.B) {          // Base.cpp:133-139
            const i32 tmp = (i32)(lcg_next(s) % (u64)(a.R - (rh.y - rh.x + 1)));
            put(at, p.x, p.y, kth_absent(a.byht_r, rh.x, rh.y, tmp));
        }
    }
}

// Advance every stream past the draws its slice consumed in `steps` calls.
__global__ void advance_kernel(u64 *state, i32 W, i32 B, i32 per, u64 draws, i32 steps, i32 lo_id, i32 hi_id) {
    const i32 id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= W || id < lo_id || id >= hi_id) return;
    i64 lo = (i64)id * per, hi = lo + per;
    if (hi > B) hi = B;
    const i64 len = hi > lo ? hi - lo : 0;
    u64 mul, add;
    lcg_jump((u64)len * draws * (u64)steps, mul, add);
    state[id] = mul * state[id] + add;
}

__global__ void widen_kernel(const i32 *__restrict__ src, i64 *__restrict__ h, i64 *__restrict__ t, i64 *__restrict__ r,
                             float *__restrict__ y, i32 S, i32 B) {
    const i32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    h[i] = src[i]; t[i] = src[S + i]; r[i] = src[2 * S + i];
    y[i] = i < B ? 1.0f : -1.0f;                            // Base.cpp:111,127,137
}
// Caller-supplied ids are VALIDATED here (TF's embedding_lookup raises InvalidArgument for an id outside the table): a bad id
// is replaced by 0 — nothing downstream can index out of bounds — and raises the context's "bad id" flag, which makes
// the update kernels of the step leave the tables alone and report a NaN loss; the host then returns OKB_ERR_ARG.
__global__ void narrow_kernel(const i64 *__restrict__ h, const i64 *__restrict__ t, const i64 *__restrict__ r,
                              i32 *__restrict__ dst, i32 S, i64 E, i64 R, unsigned *__restrict__ bad) {
    const i32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const i64 vh = h[i], vt = t[i], vr = r[i];
    const bool ok = vh >= 0 && vh < E && vt >= 0 && vt < E && vr >= 0 && vr < R;
    dst[i] = ok ? (i32)vh : 0; dst[S + i] = ok ? (i32)vt : 0; dst[2 * S + i] = ok ? (i32)vr : 0;
    if (!ok) atomicOr(bad, 1u);
}

static int push_state(okb_ctx *c) {
    okb_discard_prefetch(c);
    if (c->d_state) { cudaFree(c->d_state); c->d_state = nullptr; }
    OKB_CUDA(c, cudaMalloc((void **)&c->d_state, sizeof(u64) * c->state.size()));
    OKB_CUDA(c, cudaMemcpy(c->d_state, c->state.data(), sizeof(u64) * c->state.size(), cudaMemcpyHostToDevice));
    c->state_dirty = false;
    return 0;
}
static int pull_state(okb_ctx *c) {
    okb_discard_prefetch(c);
    if (c->state_dirty) {
        OKB_CUDA(c, cudaDeviceSynchronize());
        OKB_CUDA(c, cudaMemcpy(c->state.data(), c->d_state, sizeof(u64) * c->state.size(), cudaMemcpyDeviceToHost));
        c->state_dirty = false;
    }
    return 0;
}

// corrupt_head(0, h, r) on the host for the triple-classification fallback (Corrupt.h:132).
i64 okb_host_new_tail(okb_ctx *c, i64 h, i64 r) {
    if (c->state.empty()) return 0;
    pull_state(c);
    const i32 *b = c->byh_r.data();
    i64 lo = c->lef_h[h], hi = c->rig_h[h];
    i64 ll = lo, rr = lo - 1;
    if (hi >= lo) {
        ll = std::lower_bound(b + lo, b + hi + 1, (i32)r) - b;
        rr = (std::upper_bound(b + lo, b + hi + 1, (i32)r) - b) - 1;
    }
    i64 tmp = (i64)(lcg_next(c->state[0]) % (u64)(c->E - (rr - ll + 1)));
    push_state(c);
    if (rr < ll) return tmp;
    const i32 *v = c->byh_t.data();
    if (tmp < v[ll]) return tmp;
    if (tmp > v[rr] - rr + ll - 1) return tmp + rr - ll + 1;
    i64 lef = ll, rig = rr + 1;
    while (lef + 1 < rig) {
        i64 mid = (lef + rig) >> 1;
        if (v[mid] - mid + ll - 1 < tmp) lef = mid; else rig = mid;
    }
    return tmp + lef - ll + 1;
}

extern "C" {

int okb_rand_reset(okb_ctx *c) {                           // Random.h:9-13
    c->state.resize(c->W);
    for (i64 i = 0; i < c->W; i++) c->state[i] = (u64)rand();
    return push_state(c);
}
int okb_set_streams(okb_ctx *c, const uint64_t *st, INT w) {
    if (w < 1) OKB_FAIL(c, OKB_ERR_ARG, "need at least one stream");
    c->W = w;
    c->state.assign(st, st + w);
    return push_state(c);
}
int okb_get_streams(okb_ctx *c, uint64_t *out, INT w) {
    if (w != (INT)c->state.size()) OKB_FAIL(c, OKB_ERR_ARG, "stream count mismatch");
    int rc = pull_state(c);
    if (rc) return rc;
    for (i64 i = 0; i < w; i++) out[i] = c->state[i];
    return 0;
}

static int sample_impl(okb_ctx *c, INT B, INT k, INT kr, INT steps, INT stream_lo, INT stream_hi, i64 *mirror, cudaEvent_t sampled,
                       void *stream) {
    if (!c->d_raw) OKB_FAIL(c, OKB_ERR_STATE, "import the training files first");
    if ((i64)c->state.size() != c->W) OKB_FAIL(c, OKB_ERR_STATE, "call randReset / okb_set_streams after setWorkThreads");
    if (B < 1 || k < 0 || kr < 0 || steps < 1) OKB_FAIL(c, OKB_ERR_ARG, "bad batch geometry");
    if (B * (1 + k + kr) * steps > 0x7fffffffLL / 4) OKB_FAIL(c, OKB_ERR_ARG, "batch too large for int32 indexing");
    if (!c->in_prefetch) okb_discard_prefetch(c);          // a chunk sampled ahead is no longer the continuation of the streams
    cudaStream_t s = (cudaStream_t)stream;
    const i64 S = B * (1 + k + kr);
    if (c->batch.ensure(sizeof(i32) * 3 * S * steps)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (batch)");
    c->B = B; c->K = k; c->KR = kr; c->steps = steps;
    c->batch_from_host = false;
    c->spec_mirror = nullptr;
    c->plan_lo = c->plan_hi = 0;                            // new batches: any previous plan is stale
    SampleArgs a;
    a.raw = c->d_raw; a.run = c->d_run; a.run_ht = c->d_run_ht;
    a.byh_t = c->d_byh_t; a.byt_h = c->d_byt_h; a.byht_r = c->d_byht_r;
    a.prob = c->d_prob; a.state = c->d_state; a.out = c->batch.as<i32>();
    a.mh = mirror; a.mt = mirror ? mirror + S : nullptr; a.mr = mirror ? mirror + 2 * S : nullptr;
    a.n_raw = c->n_raw; a.new_batch = c->new_batch;
    a.E = (i32)c->E; a.R = (i32)c->R; a.B = (i32)B; a.k = (i32)k; a.kr = (i32)kr; a.W = (i32)c->W;
    a.per = (i32)(B / c->W + (B % c->W ? 1 : 0));
    a.bern = (i32)c->bern; a.steps = (i32)steps; a.stream_lo = (i32)stream_lo; a.stream_hi = (i32)stream_hi;
    const i64 total = B * steps;
    { ProfScope ps(c, PROF_SAMPLE, s);
    sample_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(a); }
    if (sampled) OKB_CUDA(c, cudaEventRecord(sampled, s));   // the batch is complete here; the stream advance below is not waited for
    if (!c->defer_advance)
        advance_kernel<<<(unsigned)((c->W + 63) / 64), 64, 0, s>>>(c->d_state, a.W, a.B, a.per, 1 + 2 * (u64)k + (u64)kr,
                                                                  a.steps, 0, a.W);   // every rank advances ALL streams
    OKB_LAUNCHED(c->defer_advance ? 1 : 2);
    c->state_dirty = true;
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}
int okb_sample(okb_ctx *c, INT B, INT k, INT kr, INT steps, INT stream_lo, INT stream_hi, void *stream) {
    return sample_impl(c, B, k, kr, steps, stream_lo, stream_hi, nullptr, nullptr, stream);
}

int okb_batch_ptrs(okb_ctx *c, INT step, const int32_t **h, const int32_t **t, const int32_t **r) {
    if (step < 0 || step >= c->steps) OKB_FAIL(c, OKB_ERR_ARG, "step out of range");
    const i64 S = c->B * (1 + c->K + c->KR);
    const i32 *base = c->batch.as<i32>() + step * 3 * S;
    *h = base; *t = base + S; *r = base + 2 * S;
    return 0;
}

// Device-visible alias of a page-locked host pointer (nullptr for pageable memory).  Under unified addressing every
// cudaHostAlloc / cudaHostRegister allocation is mapped, so a kernel can store the batch straight into the caller's
// buffer (or read it from there): one launch instead of a staging kernel plus a DMA per array.
static void *pinned_alias(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost) return at.devicePointer;
    cudaGetLastError();
    return nullptr;
}

int okb_batch_to_host(okb_ctx *c, INT step, INT *h, INT *t, INT *r, REAL *y, void *stream) {
    if (step < 0 || step >= c->steps) OKB_FAIL(c, OKB_ERR_ARG, "step out of range");
    cudaStream_t s = (cudaStream_t)stream;
    const i64 S = c->B * (1 + c->K + c->KR);
    if (t == h + S && r == t + S && !y) {                  // one pinned block (Config's batch buffers): zero-copy stores
        if (i64 *dh = (i64 *)pinned_alias(h)) {
            if (c->host_io.ensure(sizeof(float) * S)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory");
            widen_kernel<<<(unsigned)((S + 255) / 256), 256, 0, s>>>(c->batch.as<i32>() + step * 3 * S, dh, dh + S, dh + 2 * S,
                                                                       c->host_io.as<float>(), (i32)S, (i32)c->B);
            OKB_LAUNCHED(1);
            OKB_CUDA(c, cudaStreamSynchronize(s));
            return 0;
        }
    }
    if (c->host_io.ensure((sizeof(i64) * 3 + sizeof(float)) * S)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory");
    i64 *dh = c->host_io.as<i64>(), *dt = dh + S, *dr = dt + S;
    float *dy = (float *)(dr + S);
    widen_kernel<<<(unsigned)((S + 255) / 256), 256, 0, s>>>(c->batch.as<i32>() + step * 3 * S, dh, dt, dr, dy, (i32)S, (i32)c->B);
    OKB_LAUNCHED(1);
    if (t == h + S && r == t + S) {                        // caller's three arrays are one block: one copy
        OKB_CUDA(c, cudaMemcpyAsync(h, dh, sizeof(i64) * 3 * S, cudaMemcpyDeviceToHost, s));
    } else {
        OKB_CUDA(c, cudaMemcpyAsync(h, dh, sizeof(i64) * S, cudaMemcpyDeviceToHost, s));
        OKB_CUDA(c, cudaMemcpyAsync(t, dt, sizeof(i64) * S, cudaMemcpyDeviceToHost, s));
        OKB_CUDA(c, cudaMemcpyAsync(r, dr, sizeof(i64) * S, cudaMemcpyDeviceToHost, s));
    }
    if (y) OKB_CUDA(c, cudaMemcpyAsync(y, dy, sizeof(float) * S, cudaMemcpyDeviceToHost, s));
    OKB_CUDA(c, cudaStreamSynchronize(s));
    return 0;
}

// sampling() of the reference in one launch: sample one batch and return it in the caller's int64 arrays.  With one
// page-locked block (Config's batch buffers) the sample kernel itself stores the int64 mirror over PCIe and the call
// returns when that kernel is done (the stream advance runs behind it); otherwise okb_sample + okb_batch_to_host.
int okb_sample_to_host(okb_ctx *c, INT B, INT k, INT kr, INT stream_lo, INT stream_hi, INT *h, INT *t, INT *r, void *stream) {
    const i64 S = B * (1 + k + kr);
    i64 *dh = (t == h + S && r == t + S) ? (i64 *)pinned_alias(h) : nullptr;
    if (!dh) {
        int rc = sample_impl(c, B, k, kr, 1, stream_lo, stream_hi, nullptr, nullptr, stream);
        return rc ? rc : okb_batch_to_host(c, 0, h, t, r, nullptr, stream);
    }
    if (stream_lo == 0 && stream_hi >= c->W && !c->dp_on && okb_plan_small_ok(c, B, k, kr) && (i64)c->state.size() == c->W && c->d_raw) {
        // One-kernel plan: the sample kernel only fills the resident batch; the int64 copy for the caller's block is
        // written by a second cluster of the PLAN launch (posted PCIe stores next to the 8-SM sort instead of in front of
        // it), whose last CTA raises a page-locked word this call polls.  The stream advance is issued after the plan, so the
        // order on the stream is sample -> plan (+ copy) -> advance and nothing but the sort stands between the batch and
        // the step that okb_train_step_host starts on it.
        if (!c->host_flag) {
            OKB_CUDA(c, cudaHostAlloc((void **)&c->host_flag, 64, cudaHostAllocMapped));
            OKB_CUDA(c, cudaHostGetDevicePointer((void **)&c->host_flag_dev, c->host_flag, 0));
        }
        c->defer_advance = true;
        int rc = sample_impl(c, B, k, kr, 1, stream_lo, stream_hi, nullptr, nullptr, stream);
        c->defer_advance = false;
        if (rc) return rc;
        *(volatile unsigned *)c->host_flag = 0u;
        c->mirror_dst = (long long *)dh;
        c->spec_mirror = nullptr;
        rc = okb_plan_steps(c, 0, 1, stream);
        const bool copied = rc == 0 && c->mirror_dst == nullptr;          // consumed by the one-kernel plan
        c->mirror_dst = nullptr;
        const i32 per = (i32)(B / c->W + (B % c->W ? 1 : 0));
        advance_kernel<<<(unsigned)((c->W + 63) / 64), 64, 0, (cudaStream_t)stream>>>(c->d_state, (i32)c->W, (i32)B, per, 1 + 2 * (u64)k + (u64)kr, 1, 0, (i32)c->W);
        OKB_LAUNCHED(1);
        OKB_CUDA(c, cudaGetLastError());
        if (rc) return rc;
        if (!copied) return okb_batch_to_host(c, 0, h, t, r, nullptr, stream);
        c->spec_mirror = h;
        return okb_wait_word(c, c->host_flag, 0u, stream);
    }
    if (!c->ev_sampled) OKB_CUDA(c, cudaEventCreateWithFlags(&c->ev_sampled, cudaEventDisableTiming));
    int rc = sample_impl(c, B, k, kr, 1, stream_lo, stream_hi, dh, c->ev_sampled, stream);
    if (rc) return rc;
    // The batch also stays resident on the device, and the reference's loop hands exactly these arrays straight back to
    // train_step (distribute_training.py:274-282).  Plan that step NOW, behind the event the call waits on: the one-step
    // plan then runs while the host is between the two calls, and okb_train_step_host only has to VERIFY that the
    // caller's arrays still equal the resident batch (okb_batch_verify_host) instead of narrowing and planning them.
    c->spec_mirror = nullptr;
    if (stream_lo == 0 && stream_hi >= c->W && !c->dp_on && okb_plan_steps(c, 0, 1, stream) == 0) c->spec_mirror = h;
    OKB_CUDA(c, cudaEventSynchronize(c->ev_sampled));
    return 0;
}

// If (h, t, r) is the page-locked block the last okb_sample_to_host filled and that batch is still resident and planned:
// launch the comparison of the caller's arrays with the resident batch (a difference raises bit 1 of the "bad id" word,
// which makes the step's update kernels leave the tables alone and report NaN) and return 0; else return -1.
__global__ void verify_kernel(const i64 *__restrict__ h, const i64 *__restrict__ t, const i64 *__restrict__ r,
                              const i32 *__restrict__ dev, i32 S, unsigned *__restrict__ flag) {
    const i32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    if (h[i] != (i64)dev[i] || t[i] != (i64)dev[S + i] || r[i] != (i64)dev[2 * S + i]) atomicOr(flag, 2u);
}
int okb_batch_verify_host(okb_ctx *c, INT B, INT k, INT kr, const INT *h, const INT *t, const INT *r, void *stream) {
    const i64 S = B * (1 + k + kr);
    if (!c->spec_mirror || c->spec_mirror != h || t != h + S || r != t + S) return -1;
    if (c->B != B || c->K != k || c->KR != kr || c->steps < 1 || c->dp_on) return -1;
    if (!(c->plan_lo <= 0 && c->plan_hi >= 1 && c->plan_b_lo == 0 && c->plan_b_hi == B)) return -1;
    const i64 *ph = (const i64 *)pinned_alias(h);
    if (!ph) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (okb_ensure_flags(c, s)) return -1;
    // the comparison itself rides in the grad launch (extra blocks at the end of its grid read the caller's block over PCIe
    // while the positives are processed): launch_grad picks it up; okb_verify_flush launches it stand-alone otherwise
    c->verify_h = ph; c->verify_S = S;
    c->batch_from_host = true;                             // the update kernels honour the flag word
    return 0;
}
int okb_verify_flush(okb_ctx *c, void *stream) {
    if (!c->verify_h) return 0;
    const i64 *ph = (const i64 *)c->verify_h;
    const i64 S = c->verify_S;
    c->verify_h = nullptr;
    verify_kernel<<<(unsigned)((S + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ph, ph + S, ph + 2 * S, c->batch.as<i32>(), (i32)S, c->flags.as<unsigned>() + OKB_FLAGS_BAD);
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

int okb_batch_from_host(okb_ctx *c, INT B, INT k, INT kr, const INT *h, const INT *t, const INT *r, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const i64 S = B * (1 + k + kr);
    if (c->batch.ensure(sizeof(i32) * 3 * S)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (batch)");
    if (c->host_io.ensure(sizeof(i64) * 3 * S)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory");
    if (B < 1 || k < 0 || kr < 0 || !h || !t || !r) OKB_FAIL(c, OKB_ERR_ARG, "bad batch geometry");
    if (c->E < 1 || c->R < 1) OKB_FAIL(c, OKB_ERR_STATE, "import the training files first");
    int rcf = okb_ensure_flags(c, s);
    if (rcf) return rcf;
    unsigned *bad = c->flags.as<unsigned>() + OKB_FLAGS_BAD;
    c->B = B; c->K = k; c->KR = kr; c->steps = 1;
    c->batch_from_host = true;
    c->spec_mirror = nullptr;
    c->plan_lo = c->plan_hi = 0;
    if (t == h + S && r == t + S) {                        // one pinned block: the narrowing kernel reads it over PCIe itself
        if (const i64 *ph = (const i64 *)pinned_alias(h)) {
            narrow_kernel<<<(unsigned)((S + 255) / 256), 256, 0, s>>>(ph, ph + S, ph + 2 * S, c->batch.as<i32>(), (i32)S, c->E, c->R, bad);
            OKB_LAUNCHED(1);
            OKB_CUDA(c, cudaGetLastError());
            return 0;
        }
    }
    i64 *dh = c->host_io.as<i64>(), *dt = dh + S, *dr = dt + S;
    if (t == h + S && r == t + S) {
        OKB_CUDA(c, cudaMemcpyAsync(dh, h, sizeof(i64) * 3 * S, cudaMemcpyHostToDevice, s));
    } else {
        OKB_CUDA(c, cudaMemcpyAsync(dh, h, sizeof(i64) * S, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaMemcpyAsync(dt, t, sizeof(i64) * S, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaMemcpyAsync(dr, r, sizeof(i64) * S, cudaMemcpyHostToDevice, s));
    }
    narrow_kernel<<<(unsigned)((S + 255) / 256), 256, 0, s>>>(dh, dt, dr, c->batch.as<i32>(), (i32)S, c->E, c->R, bad);
    OKB_LAUNCHED(1);
    // no synchronisation needed: the copies are stream-ordered before the kernel; pageable sources are
    // staged by the driver before the call returns, pinned sources must stay untouched until the step ran
    // (Config.train_step reads the loss back, which synchronises)
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

// Stream-ordered check of the ids the last okb_batch_from_host received: synchronises, returns OKB_ERR_ARG if any id was
// outside [0, E) / [0, R) (the flag is cleared, the batch of such a call must not be trained on).
int okb_batch_check(okb_ctx *c, void *stream) {
    if (!c->flags.p) return 0;
    unsigned bad = 0;
    OKB_CUDA(c, cudaStreamSynchronize((cudaStream_t)stream));
    OKB_CUDA(c, cudaMemcpy(&bad, c->flags.as<unsigned>() + OKB_FLAGS_BAD, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (!bad) return 0;
    OKB_CUDA(c, cudaMemset(c->flags.as<unsigned>() + OKB_FLAGS_BAD, 0, sizeof(unsigned)));
    if (bad == 2u) { c->err = "caller's arrays differ from the resident batch"; return OKB_ERR_STATE; }   // okb_train_step_host falls back
    OKB_FAIL(c, OKB_ERR_ARG, "batch contains an entity or relation id outside the tables (InvalidArgument in the reference's embedding_lookup)");
}

INT okb_launch_count(void) { return (INT)g_launches.load(); }

}  // extern "C"
