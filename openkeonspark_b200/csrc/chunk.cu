// The train loop of a whole chunk of planned steps as ONE persistent kernel.
//
// Sampling and planning never depend on the parameters, so a chunk of up to 64 steps is sampled and planned ahead
// (train.cu).  What is left per step is grad -> update, two kernels of 10-17 us whose run time at the reference's batch
// size (B = 4,831: 8-62 MB per step) is launch, ramp, drain and a 2-3 deep chain of dependent loads — not bandwidth.
// Here one cooperative grid (one CTA of 17 warps per SM: 148 x 17 = 2,516 warps, 4,831 positives in two rounds) stays
// resident for the whole chunk and walks
//
//     for step in chunk:   grad phase  ->  grid barrier  ->  [hub pre-reduction -> grid barrier]  ->  update phase  ->  grid barrier
//
// with exactly the per-phase kernels' bodies (train_dev.cuh: grad_k1_body / grad_body, sgd_body, adam_gsum + adam_elem,
// loss_block), so results are bit-identical to the per-phase path; the tables (21-42 MB with Adam slots) stay in the
// 126 MB L2 from step to step.  Data another SM wrote earlier in the SAME launch (table rows, gradient rows, loss terms)
// is read with ld.global.cg (L2, never a stale L1 line); the plan and the batch ids are read-only for the whole launch.
// The grid barrier is a monotonically increasing arrival counter in global memory (release fence + atomic, acquire
// spin by one thread per CTA); cooperative launch guarantees that all CTAs are resident, the spin is bounded.
//
// MEASURED (B200, profiles/README.md r02c): slower than the per-phase kernels chained with programmatic dependent launch —
// TransH D=100 Adam 25.2-28.0 vs 18.0 us/step, TransE D=50 SGD 19.6-20.8 vs 19.2 — because one resident CTA of 16-20
// warps per SM (the grad body needs 94-128 registers) halves the occupancy of the Adam pass and nothing overlaps across
// a grid barrier, while PDL already hides most of the launch/ramp cost this kernel was built to remove.  Kept behind
// OKB_FLAG_CHUNK_KERNEL (default OFF) as the A/B baseline of that experiment.
//
// Replaces the loop body of /root/reference/distribute_training.py:267-283 (sess.run([train_op, loss]) per batch) for a
// chunk at a time; TransE.py:26-51 / TransH.py:33-69 / TransD.py:46-84 + distribute_training.py:94-101.
#include <algorithm>
#include <cstdlib>

#include "train_dev.cuh"

// Threads per CTA (one CTA per SM).  A scheduler partition holds 16 K registers, so 4 warps per partition (NT = 512) may
// use 128 registers each and 5 (NT = 640) 96: 512 keeps the k = 1 grad body (94 registers + loop state) spill-free but
// needs three rounds for 4,831 positives (148 x 16 = 2,368 warps), 640 needs two (2,960 warps) with a few spills.
// Both are compiled; okb_chunk_kernel_steps picks by measurement (OKB200_CHUNK_THREADS overrides for A/B runs).
#define CK_MAX_STEPS 64

struct ChunkArgs {
    GradArgs g;                // step 0's batch pointers; step s adds s * batch_stride
    UpdArgs u;                 // step 0's plan pointers; step s adds s * n (keys, permutation) / s * rows_all (row map)
    okb_hyper hp[CK_MAX_STEPS];
    float *loss_out;           // [n_steps] or null
    float *partial;            // hub pre-reduction buffer
    unsigned *bar;             // grid barrier arrival counter, zero at launch
    i64 batch_stride;          // 3 * S
    i32 n_steps, k1, adam, rows_all, loss_T, n_wtiles, n_pre;
};

__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned &gen) {
    __syncthreads();
    gen++;
    if (threadIdx.x == 0) {
        __threadfence();                                   // this CTA's writes are visible device-wide before it arrives
        atomicAdd(ctr, 1u);
        const unsigned target = gen * gridDim.x;
        unsigned v, polls = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            if (++polls > (1u << 27)) __trap();            // ~tens of seconds: a CTA died; surface it as a CUDA error
        }
        __threadfence();
    }
    __syncthreads();
}

// the dense TF1-Adam pass over warp tiles of 32 vectors (8 per 256-vector tile of adam_tile_kernel's table list), two
// tiles per iteration so that every lane has 2 x (row map + x + m + v) in flight before the first gradient sum
template <int VW>
__device__ __forceinline__ void adam_phase(const UpdArgs &a, i32 n_wtiles, i32 gw, i32 nw, int lane) {
    typedef typename VecT<VW>::T V;
    const float b1 = a.hp.beta1, b2 = a.hp.beta2, lr = a.hp.lr, eps = a.hp.eps, c1 = 1.f - b1, c2 = 1.f - b2;
    struct Item { const DenseTab *T; float *px, *pm, *pv; int4 seg; V xv, mv, vv; unsigned col; bool live; };
    auto fetch = [&](i32 wt, Item &it) {
        it.live = false;
        if (wt >= n_wtiles) return;
        const i32 bid = wt >> 3;
        int t = 0;
        while (bid >= a.tab[t].blk_end) t++;               // warp-uniform
        const DenseTab &T = a.tab[t];
        const unsigned nvec = (unsigned)(T.vec_end - (t ? a.tab[t - 1].vec_end : 0));
        const unsigned lv = ((unsigned)bid - (unsigned)(t ? a.tab[t - 1].blk_end : 0)) * 256u + (unsigned)(wt & 7) * 32u + lane;
        if (lv >= nvec) return;
        const unsigned vpr = (unsigned)T.D / VW;
        const unsigned row = T.magic ? (__umulhi(lv, T.magic) >> T.shift) : (lv >> T.shift);
        it.col = (lv - row * vpr) * VW;
        const size_t e = (size_t)lv * VW;
        it.T = &T; it.px = T.x + e; it.pm = T.m + e; it.pv = T.v + e;
        it.seg = __ldg(a.rowhead + T.key_off + row);
        it.xv = __ldcg(reinterpret_cast<const V *>(it.px)); it.mv = __ldcg(reinterpret_cast<const V *>(it.pm));
        it.vv = __ldcg(reinterpret_cast<const V *>(it.pv));
        it.live = true;
    };
    auto finish = [&](Item &it) {
        if (!it.live) return;
        float g[VW];
#pragma unroll
        for (int q = 0; q < VW; q++) g[q] = 0.f;
        adam_gsum<VW, true>(a, *it.T, it.seg, it.col, g);
        float *xs = reinterpret_cast<float *>(&it.xv), *ms = reinterpret_cast<float *>(&it.mv), *vs = reinterpret_cast<float *>(&it.vv);
#pragma unroll
        for (int q = 0; q < VW; q++) adam_elem(xs[q], ms[q], vs[q], g[q], b1, b2, c1, c2, lr, eps);
        *reinterpret_cast<V *>(it.px) = it.xv; *reinterpret_cast<V *>(it.pm) = it.mv; *reinterpret_cast<V *>(it.pv) = it.vv;
    };
    for (i32 wt = gw; wt < n_wtiles; wt += 2 * nw) {
        Item i0, i1;
        fetch(wt, i0);
        fetch(wt + nw, i1);
        finish(i0);
        finish(i1);
    }
}

template <int MODEL, int VW, int NV, int CK_THREADS, bool K1>
__global__ void __launch_bounds__(CK_THREADS, 1) chunk_kernel(const __grid_constant__ ChunkArgs c) {
    constexpr int CK_WARPS = CK_THREADS / 32;
    __shared__ GradArgs sg;                                // this step's argument blocks (the bodies take references)
    __shared__ UpdArgs su;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i32 gw = (i32)blockIdx.x * CK_WARPS + warp, nw = (i32)gridDim.x * CK_WARPS;
    {   // word-wise copy of the two blocks out of the kernel parameters
        const unsigned *s0 = reinterpret_cast<const unsigned *>(&c.g), *s1 = reinterpret_cast<const unsigned *>(&c.u);
        unsigned *d0 = reinterpret_cast<unsigned *>(&sg), *d1 = reinterpret_cast<unsigned *>(&su);
        for (int i = threadIdx.x; i < (int)(sizeof(GradArgs) / 4); i += CK_THREADS) d0[i] = s0[i];
        for (int i = threadIdx.x; i < (int)(sizeof(UpdArgs) / 4); i += CK_THREADS) d1[i] = s1[i];
    }
    __syncthreads();
    unsigned gen = 0;
    const i32 B = c.g.B, b_lo = c.g.b_lo, b_hi = c.g.b_hi;
    for (i32 s = 0; s < c.n_steps; s++) {
        if (threadIdx.x == 0) {                            // per-step pointers (every thread is past the previous step's barrier)
            sg.bh = c.g.bh + (i64)s * c.batch_stride; sg.bt = c.g.bt + (i64)s * c.batch_stride; sg.br = c.g.br + (i64)s * c.batch_stride;
            su.skeys = c.u.skeys + (i64)s * c.u.n; su.perm = c.u.perm + (i64)s * c.u.n;
            su.rowhead = c.u.rowhead + (i64)s * c.rows_all;
            su.hp = c.hp[s];
            su.loss_out = c.loss_out ? c.loss_out + s : nullptr;
        }
        __syncthreads();
        // ---------------- grad phase: one warp per positive, two rounds
        {
            const i32 *bh = sg.bh, *bt = sg.bt, *br = sg.br;
            if (K1) {
                for (i32 b = b_lo + gw; b < b_hi; b += nw) {
                    const i32 ph = bh[b], pt = bt[b], pr = br[b], nh = bh[b + B], nt = bt[b + B];
                    grad_k1_body<MODEL, VW, NV, true>(sg, b, lane, ph, pt, pr, nh, nt);
                }
            } else {
                for (i32 b = b_lo + gw; b < b_hi; b += nw) {
                    const i32 ph = bh[b], pt = bt[b], pr = br[b];
                    i32 nh = 0, nt = 0;
                    if (sg.k > 0) { nh = bh[b + B]; nt = bt[b + B]; }
                    grad_body<MODEL, VW, NV, 1, true>(sg, b, lane, 0, 1, nullptr, ph, pt, pr, nh, nt);
                }
            }
        }
        grid_barrier(c.bar, gen);
        // ---------------- hub rows: fixed-range pre-reduction of long segments
        if (su.hub) {
            for (i32 w = gw; w < c.n_pre; w += nw) prereduce_body<VW, NV, true>(su, c.partial, w, lane);
            grid_barrier(c.bar, gen);
        }
        // ---------------- update phase (+ the step's loss, by the first CTAs)
        if ((i32)blockIdx.x < su.loss_blocks) loss_block(su, (i32)blockIdx.x, c.loss_T);
        if (c.adam) {
            adam_phase<VW>(su, c.n_wtiles, gw, nw, lane);
        } else {
            const i32 limit = su.by_row ? su.key_limit : su.n;
            for (i32 w = gw; w < limit; w += nw) sgd_body<VW, NV, true>(su, w, lane);
        }
        if (s + 1 < c.n_steps) grid_barrier(c.bar, gen);
    }
}

template <int MODEL, int VW, int NV, int NT, bool K1>
static int launch_chunk_nt(okb_ctx *c, const ChunkArgs &a, cudaStream_t s) {
    int per_sm = 0;
    OKB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chunk_kernel<MODEL, VW, NV, NT, K1>, NT, 0));
    if (per_sm < 1) return -1;                             // cannot stay resident: the caller uses the per-phase kernels
    void *args[] = {(void *)&a};
    OKB_CUDA(c, cudaLaunchCooperativeKernel((const void *)chunk_kernel<MODEL, VW, NV, NT, K1>, dim3((unsigned)okb_sms(c)), dim3(NT), args, 0, s));
    return 0;
}
template <int MODEL, int VW, int NV>
static int launch_chunk(okb_ctx *c, const ChunkArgs &a, int nt, cudaStream_t s) {
    return nt == 512 ? launch_chunk_nt<MODEL, VW, NV, 512, true>(c, a, s) : launch_chunk_nt<MODEL, VW, NV, 640, true>(c, a, s);
}

int okb_chunk_kernel_steps(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step_lo, INT n, float *loss_out, cudaStream_t s) {
    if (!c->chunk_kernel || n < 2 || n > CK_MAX_STEPS) return -1;
    if (m->model != OKB_TRANSE && m->model != OKB_TRANSH && m->model != OKB_TRANSD) return -1;
    if (c->dp_on || c->batch_from_host || c->prof_on || c->l2_prefetch || c->adam_legacy || c->adam_tma) return -1;
    // the k = 1, kr = 0 batch only (the reference's default): the generic grad body inlined into this kernel is contracted
    // to FMAs differently than in grad_kernel, so its sums differ in the last bit — not acceptable for a drop-in A/B path
    if (!(c->K == 1 && c->KR == 0 && !c->grad_generic)) return -1;
    int vw, nv;
    if (!okb_pick_layout(m->ent_dim, vw, nv) || m->ent_dim != m->rel_dim) return -1;
    if (!((vw == 4 && (nv == 1 || nv == 2)) || (vw == 2 && nv == 1))) return -1;      // D in (64, 256] with D % 4 == 0; even D in (32, 64]
    int coop = 0, dev = 0;
    OKB_CUDA(c, cudaGetDevice(&dev));
    OKB_CUDA(c, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return -1;
    ChunkArgs a;
    okb_fill_grad(c, m, hp, step_lo, 0, c->B, 0, c->gent.as<float>(), c->grel.as<float>(), c->lossterms.as<float>(), a.g);
    i32 blk = 0;
    bool lean = true;
    int rc = okb_fill_update(c, m, hp, step_lo, c->gent.as<float>(), c->grel.as<float>(), c->lossterms.as<float>(), nullptr, vw, s, a.u, blk, lean);
    if (rc) return rc;
    if (m->optimizer == OKB_ADAM && !lean) return -1;
    for (INT i = 0; i < n; i++) a.hp[i] = hp[i];
    a.loss_out = loss_out;
    a.partial = c->partial.as<float>();
    a.bar = c->flags.as<unsigned>() + OKB_FLAGS_GRIDBAR;
    a.batch_stride = 3 * c->B * (1 + c->K + c->KR);
    a.n_steps = (i32)n;
    a.k1 = (c->K == 1 && c->KR == 0 && !c->grad_generic) ? 1 : 0;
    a.adam = m->optimizer == OKB_ADAM ? 1 : 0;
    a.rows_all = (i32)(c->E + c->R);
    a.loss_T = a.adam ? 256 : WARPS_PER_BLOCK * 32;       // block sizes of adam_tile_kernel / sgd_kernel (loss_block's order)
    a.n_wtiles = blk * 8;
    a.n_pre = (i32)(a.u.n / PCH + 1);
    if (!a.adam) a.u.by_row = a.u.key_limit < a.u.n;      // as okb_update chooses for sgd_kernel
    OKB_CUDA(c, cudaMemsetAsync(a.bar, 0, sizeof(unsigned), s));
    // CTA size: the fewest grad rounds, ties to the spill-free 512 (see the note at the top)
    const i64 sms = okb_sms(c), Bn = c->B;
    int nt = ((Bn + sms * 16 - 1) / (sms * 16) <= (Bn + sms * 20 - 1) / (sms * 20)) ? 512 : 640;
    if (const char *e = getenv("OKB200_CHUNK_THREADS")) { const int v = atoi(e); if (v == 512 || v == 640) nt = v; }
    rc = -1;
#define CK_CALL(VW, NV)                                                                     \
    if (m->model == OKB_TRANSE) rc = launch_chunk<OKB_TRANSE, VW, NV>(c, a, nt, s);          \
    else if (m->model == OKB_TRANSH) rc = launch_chunk<OKB_TRANSH, VW, NV>(c, a, nt, s);     \
    else rc = launch_chunk<OKB_TRANSD, VW, NV>(c, a, nt, s)
    if (vw == 4 && nv == 1) { CK_CALL(4, 1); } else if (vw == 4 && nv == 2) { CK_CALL(4, 2); } else { CK_CALL(2, 1); }
    if (rc) return rc;
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}
