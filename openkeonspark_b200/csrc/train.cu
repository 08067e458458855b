// Fused train step for TransE / TransH / TransD: the reference's loss_def graphs
// (TransE.py:26-51, TransH.py:33-69, TransD.py:46-84 on the batch layout of Model.py:55-74)
// plus optimizer.minimize (distribute_training.py:94-101), as three phases:
//
//   plan   (integer)  gradient-row keys of the batch -> segmented stable radix sort (radix.cu) -> per step
//                     (sorted keys, slot permutation, row map {first, end, slot0, slot1} per table row);
//                     whole chunks of steps are planned at once and, through okb_chunk_prefetch, one chunk
//                     ahead on a side stream (sampling and planning never depend on the parameters); a plan of
//                     ONE step (the host-batch API) is a single cluster kernel (plan_small_kernel)
//   grad   (fused)    one warp per positive: 128-bit gathers of the h/t/r rows (+ the model's auxiliary
//                     rows) issued before the first reduction, projection, l2-normalise, L1, margin hinge
//                     against each of its negatives, and the backward pass; rows shared between a positive
//                     and its negatives are gathered ONCE and their gradients accumulate in registers, so a
//                     positive group emits exactly (2 + k) entity and (1 + kr) relation gradient rows.
//                     grad_k1_kernel: the k = 1, kr = 0 batch with interleaved reduction chains;
//                     grad_kernel<..., 4>: 2-4 warps per positive for small batches with many negatives
//   update (fused)    segmented sum of the gradient rows in sorted (slot) order — a fixed fp32 order, so runs
//                     and replicas are bit-identical — and SGD (sgd_kernel: one warp per touched row) or the
//                     TF1 dense-decay Adam rule over every row (adam_tile_kernel: one 256-vector tile of one
//                     table per CTA) in the same pass; the mean hinge loss is reduced by a few extra blocks of
//                     the same launch (the FIRST blocks of the Adam grid, so that a host waiting for the loss in
//                     a page-locked word — okb_wait_word / okb_train_step_host — gets it while the update runs)
//
// The grad and update kernels are chained with programmatic dependent launch (griddepcontrol): each one's
// parameter-independent prologue overlaps the tail of its predecessor.
// The last section is the owner-sharded data-parallel form of the step over NVLink peer memory: the scatter form (the grad
// kernels' SC variant + dp_scatter_update_kernel: two kernels per step, bit-identical to one GPU), the push form
// (dp_reduce_push_kernel / dp_owner_kernel) and the measured-slower gather / pull alternatives.
//
// Roofline: latency-bound gathers and an HBM/L2 stream of fp32 rows; no dense contraction, so no tensor cores
// here (TransR's projection lives in transr.cu / transr_tc.cu).
#include <cooperative_groups.h>
#include <algorithm>
#include <cstring>

#include "okb_internal.h"

#include "train_dev.cuh"

// ------------------------------------------------------------------------------------------ plan
struct PlanArgs {
    const i32 *batch;          // [steps][3][S] plane-major batches
    i32 *keys;                 // [C][B*NE + B*NR] composite keys
    i32 B, k, kr, NE, NR, E, R, S, step_lo, C;
    i32 b_lo, Bl;              // positives [b_lo, b_lo + Bl) of every step are planned (a data-parallel rank plans its own)
};
// Combined key space of one step: entity row e -> e, relation row r -> E + r, unused slot -> E + R.
// Several steps are planned by ONE segmented sort (one segment per step, radix.cu).
__global__ void plan_keys_kernel(PlanArgs a) {
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (i64)a.C * a.Bl) return;
    const i32 c = (i32)(tid / a.Bl), bl = (i32)(tid % a.Bl), b = a.b_lo + bl;
    const i32 *bh = a.batch + (i64)(a.step_lo + c) * 3 * a.S, *bt = bh + a.S, *br = bt + a.S;
    const i32 ph = bh[b], pt = bt[b], pr = br[b];
    const i32 n = a.Bl * (a.NE + a.NR), off = 0;
    i32 *ke = a.keys + (i64)c * n + (i64)bl * a.NE;
    i32 *kr_ = a.keys + (i64)c * n + (i64)a.Bl * a.NE + (i64)bl * a.NR;
    const i32 none = off + a.E + a.R;
    ke[0] = off + ph; ke[1] = off + pt; kr_[0] = off + a.E + pr;
    for (i32 m = 0; m < a.k; m++) {
        const i32 at = b + (m + 1) * a.B;
        const i32 nh = bh[at], nt = bt[at];
        ke[2 + m] = nh != ph ? off + nh : (nt != pt ? off + nt : none);
    }
    for (i32 m = 0; m < a.kr; m++) {
        const i32 nr = br[b + (1 + a.k + m) * a.B];
        kr_[1 + m] = nr != pr ? off + a.E + nr : none;
    }
}

// ---- one-step plan in ONE kernel (the host-buffer path plans a single step per call: keys + 8 sort launches + memset +
// head marking were ~11 launches of 2-8 us kernels; launch-bound).  One CLUSTER of 8 CTAs x 1024 threads: every thread
// keeps its <= 4 packed items (key << ib | index) in registers; warp g of the cluster owns the contiguous index range
// [g*chunk, (g+1)*chunk), so warp-private counter columns + a ballot-built peer mask give a STABLE two-pass LSD sort
// (digit = half the key bits).  Per pass: local histogram -> per-CTA digit totals exchanged through distributed shared
// memory -> every CTA derives its own bases -> scatter.  Both passes scatter into the owning CTA's shared-memory slice
// by remote stores (pass 1 from registers, pass 2 from the local slice); the sorted slice is then written out coalesced
// (keys, permutation) and the row -> {first, end, slot0, slot1} map is marked straight from shared memory.
// Output is identical to plan_keys + okb_sort_pairs_seg + mark_heads.
// (MATCH.ANY issues about once per 64 cycles per SM -- measured: a one-CTA version built on it took 90 us -- hence ballots.)
struct PlanSmallArgs {
    PlanArgs p;
    i32 *skeys, *perm;         // [n] sorted keys, original index of every sorted entry
    int4 *rowhead;             // [rows]
    i32 n, rows, ib, db;
    unsigned magic_e, magic_r; // ceil(2^32 / NE), ceil(2^32 / NR): exact division of an index < 2^16 by one multiply
    i32 chunk;                 // entries per warp: 32, 64 or 128 (a power of two, so that position -> owning CTA is a shift)
    // optional second cluster of the launch: int64 copy of the step's batch [3][S] into the caller's page-locked block
    // (okb_sample_to_host), then *mirror_flag = mirror_token for the host that waits on it
    long long *mirror;
    unsigned *mirror_flag, *mirror_ctr;
    unsigned mirror_token;
    i32 nsteps;                // clusters 0 .. nsteps-1 plan steps step_lo .. step_lo+nsteps-1 (outputs one after the other)
};
#define PS_CTAS 8
#define PS_THREADS 1024
#define PS_WARPS (PS_THREADS / 32)
#define PS_ITEMS 4
#define PS_MAX_N (PS_CTAS * PS_THREADS * PS_ITEMS)
// key of sort entry i of the step (plan_keys_kernel's layout), branch-free so that a thread's entries load together
__device__ __forceinline__ i32 plan_key(const PlanArgs &a, unsigned magic_e, unsigned magic_r, const i32 *bh, const i32 *bt,
                                        const i32 *br, i32 i) {
    const i32 ne_tot = a.Bl * a.NE, none = a.E + a.R;
    const bool ent = i < ne_tot;
    const i32 i2 = ent ? i : i - ne_tot, per = ent ? a.NE : a.NR;
    const i32 bl = per == 1 ? i2 : (i32)__umulhi((unsigned)i2, ent ? magic_e : magic_r), j = i2 - bl * per, b = a.b_lo + bl;
    const i32 plane = ent ? (j >= 2 ? j - 1 : 0) : (j >= 1 ? a.k + j : 0), at = b + plane * a.B;
    const i32 *p0 = ent ? bh : br;
    const i32 v0 = p0[b], v1 = bt[b], w0 = p0[at], w1 = bt[at];
    if (ent) return j == 0 ? v0 : (j == 1 ? v1 : (w0 != v0 ? w0 : (w1 != v1 ? w1 : none)));
    return j == 0 ? a.E + v0 : (w0 != v0 ? a.E + w0 : none);
}
// lanes of the warp that hold a valid item with the same digit as this lane (meaningful for valid lanes only)
__device__ __forceinline__ unsigned plan_peers(unsigned dg, bool ok, int db) {
    unsigned m = __ballot_sync(FULL, ok);
#pragma unroll
    for (int b = 0; b < 8; b++)
        if (b < db) {
            const bool bit = (dg >> b) & 1u;
            const unsigned v = __ballot_sync(FULL, bit);
            m &= bit ? v : ~v;
        }
    return m;
}
struct PlanSmem { unsigned *items, *items2, *cnt, *tot, *base, *wtot; };
// cnt[w][d] (this CTA's warp histograms) -> cnt[w][d] = entries of digit d in lower warps of this CTA, and
// base[d] = first output position of this CTA's entries of digit d (digit-major, then CTA, then warp order)
__device__ __forceinline__ void plan_bases(cooperative_groups::cluster_group &cl, const PlanSmem &S, int ndig, int t) {
    const int lane = t & 31, w = t >> 5, stride = ndig + 1;
    if (t < ndig) {
        unsigned run = 0;
#pragma unroll
        for (int q = 0; q < PS_WARPS; q++) { const unsigned v = S.cnt[q * stride + t]; S.cnt[q * stride + t] = run; run += v; }
        S.tot[t] = run;
    }
    cl.sync();
    unsigned total = 0, lower = 0;
    if (t < ndig) {
        const unsigned me = cl.block_rank();
#pragma unroll
        for (unsigned q = 0; q < PS_CTAS; q++) {
            const unsigned v = cl.map_shared_rank(S.tot, q)[t];
            total += v;
            if (q < me) lower += v;
        }
    }
    unsigned inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) S.wtot[w] = inc;
    __syncthreads();
    if (w == 0) {
        const unsigned x = lane < PS_WARPS ? S.wtot[lane] : 0u;
        unsigned i2 = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(FULL, i2, o); if (lane >= o) i2 += y; }
        if (lane < PS_WARPS) S.wtot[lane] = i2 - x;
    }
    __syncthreads();
    if (t < ndig) S.base[t] = S.wtot[w] + inc - total + lower;
    __syncthreads();
}
// stable scatter of one 32-entry row of a warp: entry v (digit dg) of every valid lane goes to position
// base[dg] + (entries of dg before it in this CTA), into the slice of the cluster-distributed array `dst` that owns it
__device__ __forceinline__ void plan_scatter(cooperative_groups::cluster_group &cl, const PlanSmem &S, unsigned *col, unsigned *dst,
                                             unsigned v, unsigned dg, bool ok, int db, int lane, int ss) {
    const unsigned m = plan_peers(dg, ok, db);
    const int leader = ok ? __ffs(m) - 1 : 0;
    unsigned b0 = 0;
    if (ok && lane == leader) { const unsigned cur = col[dg]; col[dg] = cur + __popc(m); b0 = S.base[dg] + cur; }
    b0 = __shfl_sync(FULL, b0, leader);
    if (ok) {
        const unsigned pos = b0 + __popc(m & ((1u << lane) - 1));
        cl.map_shared_rank(dst, pos >> ss)[pos & ((1u << ss) - 1)] = v;
    }
    __syncwarp();
}
__global__ void __cluster_dims__(PS_CTAS, 1, 1) __launch_bounds__(PS_THREADS, 1) plan_small_kernel(PlanSmallArgs a) {
    namespace cg = cooperative_groups;
#ifdef PS_TIMING                                            // per-phase globaltimer stamps of CTA 0 (tools/build_variant.sh)
    unsigned long long ts[12]; int nts = 0;
#define PS_STAMP() do { if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts[nts])); nts++; } while (0)
#else
#define PS_STAMP() do { } while (0)
#endif
    PS_STAMP();
    if ((i32)blockIdx.x >= a.nsteps * PS_CTAS) {            // the mirror cluster: posted PCIe stores while the first cluster sorts
        pdl_wait();
        pdl_launch_dependents();
        const i32 *src = a.p.batch + (i64)a.p.step_lo * 3 * a.p.S;
        const i32 first = a.nsteps * PS_CTAS, tot = 3 * a.p.S, nth = ((i32)gridDim.x - first) * PS_THREADS;
        for (i32 i = ((i32)blockIdx.x - first) * PS_THREADS + (i32)threadIdx.x; i < tot; i += nth) a.mirror[i] = (long long)src[i];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned done = atomicAdd(a.mirror_ctr, 1u);
            if (done == gridDim.x - (unsigned)(a.nsteps * PS_CTAS) - 1) {     // last mirror CTA: every CTA's stores are ordered before the flag
                *a.mirror_ctr = 0u;
                __threadfence_system();
                *(volatile unsigned *)a.mirror_flag = a.mirror_token;
                __threadfence_system();
            }
        }
        return;
    }
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ __align__(16) unsigned psm[];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, cta = (int)cl.block_rank();
    const i32 n = a.n, ndig = 1 << a.db, stride = ndig + 1;
    const i32 chunk = a.chunk, slice = chunk * PS_WARPS, ss = 31 - __clz(slice);     // slice is a power of two
    PlanSmem S;
    S.items = psm; S.items2 = psm + slice; S.cnt = S.items2 + slice; S.tot = S.cnt + PS_WARPS * stride; S.base = S.tot + ndig;
    S.wtot = S.base + ndig;
    unsigned *col = S.cnt + w * stride;
    const unsigned dmask = ndig - 1, imask = (1u << a.ib) - 1;
    const i32 s_lo = cta * slice, s_hi = min(n, s_lo + slice);               // this CTA's slice of the index / position space
    const i32 lo = s_lo + w * chunk, hi = min(n, lo + chunk);                // this warp's entries
    const int ib = a.ib, db = a.db, gt = cta * PS_THREADS + t;
    for (i32 i = t; i < PS_WARPS * stride; i += PS_THREADS) S.cnt[i] = 0;
    // a chunk of steps: cluster c plans step step_lo + c into the c-th slice of the outputs
    const i32 cstep = (i32)blockIdx.x / PS_CTAS;
    a.skeys += (i64)cstep * n; a.perm += (i64)cstep * n; a.rowhead += (i64)cstep * a.rows;
    pdl_wait();                   // the batch comes from the previous kernel; the previous step's update still reads the row map
    pdl_launch_dependents();      // the grad kernel may become resident now: it waits for this grid before it touches the plan
    for (i32 i = gt; i < a.rows; i += PS_CTAS * PS_THREADS) a.rowhead[i] = make_int4(-1, -1, -1, -1);
    const i32 *bh = a.p.batch + (i64)(a.p.step_lo + cstep) * 3 * a.p.S, *bt = bh + a.p.S, *br = bt + a.p.S;
    unsigned x[PS_ITEMS];
#pragma unroll
    for (int it = 0; it < PS_ITEMS; it++) {
        const i32 i = lo + it * 32 + lane;
        x[it] = i < hi ? ((unsigned)plan_key(a.p, a.magic_e, a.magic_r, bh, bt, br, i) << ib) | (unsigned)i : 0xffffffffu;
    }
    __syncthreads();
    PS_STAMP();
    // ---- pass 1 (low digit): registers -> the cluster's distributed array `items`
#pragma unroll
    for (int it = 0; it < PS_ITEMS; it++)
        if (x[it] != 0xffffffffu) atomicAdd(&col[(x[it] >> ib) & dmask], 1u);
    __syncthreads();
    PS_STAMP();
    plan_bases(cl, S, ndig, t);
    PS_STAMP();
#pragma unroll
    for (int it = 0; it < PS_ITEMS; it++)
        if (lo + it * 32 < hi) {
            const bool ok = x[it] != 0xffffffffu;
            plan_scatter(cl, S, col, S.items, x[it], ok ? (x[it] >> ib) & dmask : 0u, ok, db, lane, ss);
        }
    cl.sync();
    PS_STAMP();
    // ---- pass 2 (high digit): local slice of `items` -> the distributed array `items2`
    for (i32 i = t; i < PS_WARPS * stride; i += PS_THREADS) S.cnt[i] = 0;
    __syncthreads();
    for (i32 i = lo + lane; i < hi; i += 32) atomicAdd(&col[(S.items[i - s_lo] >> (ib + db)) & dmask], 1u);
    __syncthreads();
    PS_STAMP();
    plan_bases(cl, S, ndig, t);
    PS_STAMP();
    for (i32 base = lo; base < hi; base += 32) {
        const i32 i = base + lane;
        const bool ok = i < hi;
        const unsigned v = ok ? S.items[i - s_lo] : 0u;
        plan_scatter(cl, S, col, S.items2, v, ok ? (v >> (ib + db)) & dmask : 0u, ok, db, lane, ss);
    }
    cl.sync();
    PS_STAMP();
    // ---- sorted keys / permutation (coalesced) and the row map (mark_heads_kernel) out of shared memory; the three
    // neighbours of an entry at a slice boundary are read from the adjacent CTA's slice
    auto sorted_key = [&](i32 j) -> i32 {
        if (j < 0 || j >= n) return -1;
        const unsigned v = (j >= s_lo && j < s_hi) ? S.items2[j - s_lo] : cl.map_shared_rank(S.items2, j >> ss)[j & (slice - 1)];
        return (i32)(v >> ib);
    };
    for (i32 i = s_lo + t; i < s_hi; i += PS_THREADS) {
        const unsigned v = S.items2[i - s_lo];
        const i32 key = (i32)(v >> ib), idx = (i32)(v & imask);
        a.skeys[i] = key;
        a.perm[i] = idx;
        if (key >= a.rows) continue;
        int4 *o = a.rowhead + key;
        if (sorted_key(i - 1) != key) { o->x = i; o->z = idx; }
        else if (sorted_key(i - 2) != key) o->w = idx;               // second entry of its segment
        if (sorted_key(i + 1) != key) o->y = i + 1;
    }
    cl.sync();                                                       // no CTA exits while a neighbour may still read its slice
#ifdef PS_TIMING
    PS_STAMP();
    if (cta == 0 && t == 0)
        printf("plan_small ns: keys %llu hist1 %llu bases1 %llu scat1 %llu hist2 %llu bases2 %llu scat2 %llu heads %llu total %llu\n",
               ts[1] - ts[0], ts[2] - ts[1], ts[3] - ts[2], ts[4] - ts[3], ts[5] - ts[4], ts[6] - ts[5], ts[7] - ts[6], ts[8] - ts[7],
               ts[8] - ts[0]);
#endif
}

// ---- Pipelined TMA form of the same pass (128-bit rows, D % 4 == 0).
// The register kernel above runs as two lock-stepped waves: every warp of a wave loads, then every warp computes
// (IEEE sqrt + divide per element: ~330 instructions per warp), then every warp stores — the memory pipe and the
// issue slots take turns.  Here a PERSISTENT grid of 2 CTAs per SM walks the tiles (256 vectors of one table) through
// a ring of AP_STAGES shared-memory stages: thread 0 keeps the x / m / v slices of the next AP_STAGES tiles in
// flight with bulk-async (TMA) copies on per-stage mbarriers — 144 KB per SM without holding a register — while all
// threads update the tile that has landed and hand it back to the copy engine with bulk stores.  The row-map entry
// of tile j+2 and the first two gradient rows of tile j+1 are requested while tile j is computed, so the dependent
// gather chain is off the critical path too.  Everything but the gradient rows is independent of this step's grad
// kernel: under programmatic dependent launch the prologue overlaps that kernel's tail.
#define AP_TILE 256
#define AP_STAGES 6
#define ADAM_TILE_V AP_TILE
struct __align__(128) ApStage { float4 x[AP_TILE], m[AP_TILE], v[AP_TILE]; };
__device__ __forceinline__ unsigned a_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

struct ApTile { int t; i64 vec0; i32 nvec; };
__device__ __forceinline__ ApTile ap_tile(const UpdArgs &a, int tile) {
    ApTile r;
    r.t = 0;
    while (tile >= a.tab[r.t].blk_end) r.t++;
    const i64 tab_vecs = a.tab[r.t].vec_end - (r.t ? a.tab[r.t - 1].vec_end : 0);
    r.vec0 = (i64)(tile - (r.t ? a.tab[r.t - 1].blk_end : 0)) * AP_TILE;
    r.nvec = (i32)min((i64)AP_TILE, tab_vecs - r.vec0);
    return r;
}

__global__ void __launch_bounds__(AP_TILE, 2) adam_pipe_kernel(UpdArgs a, int ntiles) {
    typedef float4 V;
    extern __shared__ __align__(128) unsigned char adam_sm[];
    ApStage *st = reinterpret_cast<ApStage *>(adam_sm);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(st + AP_STAGES);
    const int G = a.work_blocks, tid = threadIdx.x;
    if ((int)blockIdx.x >= G) { pdl_launch_dependents(); pdl_wait(); loss_block(a, (i32)blockIdx.x - G, (i32)blockDim.x); return; }
    if (a.bad) { pdl_launch_dependents(); pdl_wait(); if (upd_bad(a)) return; }
    const int my_n = (ntiles - (int)blockIdx.x + G - 1) / G;         // this CTA's tiles: blockIdx.x + j * G
    if (tid == 0) {
        for (int s = 0; s < AP_STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a_smem(full + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int j) {                              // thread 0: request tile j's slices into stage j % AP_STAGES
        const ApTile ti = ap_tile(a, (int)blockIdx.x + j * G);
        const DenseTab &T = a.tab[ti.t];
        ApStage *sg = st + (j % AP_STAGES);
        const unsigned bar = a_smem(full + (j % AP_STAGES)), bytes = (unsigned)ti.nvec * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(3u * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(a_smem(sg->x)), "l"(T.x + ti.vec0 * 4), "r"(bytes), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(a_smem(sg->m)), "l"(T.m + ti.vec0 * 4), "r"(bytes), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(a_smem(sg->v)), "l"(T.v + ti.vec0 * 4), "r"(bytes), "r"(bar) : "memory");
    };
    if (tid == 0) for (int j = 0; j < min(AP_STAGES, my_n); j++) issue(j);
    pdl_launch_dependents();                               // next step's grad kernel may prefetch its batch ids

    // per-thread view of a tile: its vector's row-map entry, gradient base and column
    struct View { int4 seg; const float *gb; i32 cols, soff, pcol; bool live; };
    auto view = [&](int j) {
        View w;
        w.seg = make_int4(-1, 0, 0, 0); w.gb = nullptr; w.cols = 0; w.soff = 0; w.pcol = 0; w.live = false;
        if (j < my_n) {
            const ApTile ti = ap_tile(a, (int)blockIdx.x + j * G);
            if (tid < ti.nvec) {
                const DenseTab &T = a.tab[ti.t];
                const unsigned lv = (unsigned)(ti.vec0 + tid), vpr = (unsigned)T.D / 4u, row = lv / vpr;
                w.pcol = T.part * T.D + (i32)((lv - row * vpr) * 4u);
                w.gb = T.grad + w.pcol; w.cols = T.cols; w.soff = T.slot_off; w.live = true;
                w.seg = __ldg(a.rowhead + T.key_off + row);
            }
        }
        return w;
    };
    auto first_two = [&](const View &w, V &g0, V &g1) {    // the first two contributions come straight from the row map
        g0 = make_float4(0.f, 0.f, 0.f, 0.f);
        g1 = g0;
        if (w.seg.x >= 0) {
            g0 = __ldg(reinterpret_cast<const V *>(w.gb + (i64)(w.seg.z - w.soff) * w.cols));
            if (w.seg.y - w.seg.x > 1) g1 = __ldg(reinterpret_cast<const V *>(w.gb + (i64)(w.seg.w - w.soff) * w.cols));
        }
    };
    View cur = view(0), nxt = view(1);
    pdl_wait();                                            // gradient rows of this step are complete from here on
    V g0, g1;
    first_two(cur, g0, g1);
    const float b1 = a.hp.beta1, b2 = a.hp.beta2, lr = a.hp.lr, eps = a.hp.eps;
    for (int j = 0; j < my_n; j++) {
        View nn = view(j + 2);
        V h0, h1;
        first_two(nxt, h0, h1);
        if (tid == 0 && j >= 1 && j + AP_STAGES - 1 < my_n) {
            // stage (j-1) % AP_STAGES was handed to the copy engine at the end of the previous iteration: once the engine
            // has read it, tile j + AP_STAGES - 1 can land there
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue(j + AP_STAGES - 1);
        }
        {
            const unsigned bar = a_smem(full + (j % AP_STAGES)), parity = (unsigned)(j / AP_STAGES) & 1u;
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        ApStage *sg = st + (j % AP_STAGES);
        if (cur.live) {
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            if (cur.seg.x >= 0) {
                const i32 sx_ = cur.seg.x, sy = cur.seg.y, cnt = sy - sx_;
                const bool long_seg = a.hub && cnt > PCH && (sx_ + PCH - 1) / PCH < sy / PCH;
                auto add = [&](const V &w) { g[0] += w.x; g[1] += w.y; g[2] += w.z; g[3] += w.w; };
                auto add_raw = [&](i32 lo, i32 hi) {
                    for (i32 q = lo; q < hi; q++) add(__ldg(reinterpret_cast<const V *>(cur.gb + (i64)(__ldg(a.perm + q) - cur.soff) * cur.cols)));
                };
                if (!long_seg) {
                    add(g0);
                    if (cnt > 1) add(g1);
                    add_raw(sx_ + 2, sy);
                } else {                                   // hub row: fringe rows + pre-reduced interior blocks, ascending order
                    const i32 bb0 = (sx_ + PCH - 1) / PCH, bb1 = sy / PCH;
                    const float *pb = a.partial + cur.pcol;
                    add_raw(sx_, bb0 * PCH);
                    for (i32 bb = bb0; bb < bb1; bb++) add(__ldg(reinterpret_cast<const V *>(pb + (i64)bb * a.pcols)));
                    add_raw(bb1 * PCH, sy);
                }
            }
            V xv = sg->x[tid], mv = sg->m[tid], vv = sg->v[tid];
            float *xs = reinterpret_cast<float *>(&xv), *ms = reinterpret_cast<float *>(&mv), *vs = reinterpret_cast<float *>(&vv);
#pragma unroll
            for (int q = 0; q < 4; q++) adam_elem(xs[q], ms[q], vs[q], g[q], b1, b2, 1.f - b1, 1.f - b2, lr, eps);
            sg->x[tid] = xv; sg->m[tid] = mv; sg->v[tid] = vv;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            const ApTile ti = ap_tile(a, (int)blockIdx.x + j * G);
            const DenseTab &T = a.tab[ti.t];
            const unsigned bytes = (unsigned)ti.nvec * 16u;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(T.x + ti.vec0 * 4), "r"(a_smem(sg->x)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(T.m + ti.vec0 * 4), "r"(a_smem(sg->m)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(T.v + ti.vec0 * 4), "r"(a_smem(sg->v)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        cur = nxt; nxt = nn; g0 = h0; g1 = h1;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// rowhead[c][key] = {first, end, slot0, slot1}: range (within step c) of the sorted entries of table row
// `key` and the gradient slots of its first two entries
__global__ void mark_heads_kernel(const i32 *__restrict__ skeys, const i32 *__restrict__ perm, int4 *__restrict__ rowhead,
                                  i32 n, i32 rows, i64 total) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const i32 c = (i32)(i / n), pos = (i32)(i - (i64)c * n);
    const i32 key = skeys[i];
    if (key >= rows) return;
    int4 *o = rowhead + (i64)c * rows + key;
    const bool first = pos == 0 || skeys[i - 1] != key;
    if (first) { o->x = pos; o->z = perm[i]; }
    else if (pos == 1 || skeys[i - 2] != key) o->w = perm[i];      // second entry of its segment
    if (pos == n - 1 || skeys[i + 1] != key) o->y = pos + 1;
}

// ------------------------------------------------------------------------------------------ dispatch
// Lane layout for a row of D floats: the widest vector (<= 128 bit) dividing D that still keeps
// most lanes busy, and 1, 2 or 4 vectors per lane (D <= 512 for D % 4 == 0).
bool okb_pick_layout(int D, int &vw, int &nv) {
    vw = (D % 4 == 0 && D > 64) ? 4 : ((D % 2 == 0 && D > 32) ? 2 : 1);
    for (;;) {
        int n = (D + 32 * vw - 1) / (32 * vw);
        nv = n <= 1 ? 1 : (n <= 2 ? 2 : 4);
        if (n <= 4) return true;
        if (vw < 4 && D % (vw * 2) == 0) vw *= 2; else return false;
    }
}

#define DISPATCH_LAYOUT(vw, nv, CALL)                                                    \
    do {                                                                                 \
        if (vw == 4 && nv == 1) { CALL(4, 1); } else if (vw == 4 && nv == 2) { CALL(4, 2); } \
        else if (vw == 4 && nv == 4) { CALL(4, 4); } else if (vw == 2 && nv == 1) { CALL(2, 1); } \
        else if (vw == 2 && nv == 2) { CALL(2, 2); } else if (vw == 2 && nv == 4) { CALL(2, 4); } \
        else if (vw == 1 && nv == 1) { CALL(1, 1); } else if (vw == 1 && nv == 2) { CALL(1, 2); } \
        else { CALL(1, 4); }                                                             \
    } while (0)

int okb_transr_check(okb_ctx *c, const okb_model *m);
int okb_transr_launch_grad(okb_ctx *c, const okb_model *m, const okb_hyper *hp, const i32 *batch, const i32 *skeys, const i32 *perm,
                           const int4 *rowhead, i64 n, INT b_lo, INT b_hi, float *gent, float *grel, float *loss_terms, cudaStream_t s);
int okb_transr_launch_rel_update(okb_ctx *c, const okb_model *m, const okb_hyper *hp, const int4 *rowhead, const float *grel, cudaStream_t s,
                                 const i32 *perm);

static int check_model(okb_ctx *c, const okb_model *m, int &vw, int &nv) {
    if (!m || !m->ent || !m->rel) OKB_FAIL(c, OKB_ERR_ARG, "model tables missing");
    if (m->model == OKB_TRANSR) {                          // entity rows go through the generic kernels, the rest through transr.cu
        int rc = okb_transr_check(c, m);
        if (rc) return rc;
        if (!okb_pick_layout(m->ent_dim, vw, nv)) OKB_FAIL(c, OKB_ERR_ARG, "embedding dimension not supported");
        return 0;
    }
    if (m->model != OKB_TRANSE && m->model != OKB_TRANSH && m->model != OKB_TRANSD) OKB_FAIL(c, OKB_ERR_ARG, "unknown model");
    if (m->ent_dim != m->rel_dim) OKB_FAIL(c, OKB_ERR_ARG, "TransE/H/D need ent_dim == rel_dim");
    if (m->model != OKB_TRANSE && !m->rel_aux) OKB_FAIL(c, OKB_ERR_ARG, "rel_aux table missing");
    if (m->model == OKB_TRANSD && !m->ent_aux) OKB_FAIL(c, OKB_ERR_ARG, "ent_aux table missing");
    if (!okb_pick_layout(m->ent_dim, vw, nv)) OKB_FAIL(c, OKB_ERR_ARG, "embedding dimension not supported (need D <= 512 with D % 4 == 0, D <= 256 even, or D <= 128)");
    return 0;
}
static void group_cols(const okb_model *m, i32 &ce, i32 &cr) {
    ce = m->model == OKB_TRANSD ? 2 * m->ent_dim : m->ent_dim;
    cr = m->model == OKB_TRANSE ? m->rel_dim : 2 * m->rel_dim;
    if (m->model == OKB_TRANSR) cr = m->rel_dim + m->ent_dim * m->rel_dim;     // [d rel | d M_r], one row per RELATION
}

int okb_ensure_flags(okb_ctx *c, cudaStream_t s) {
    if (c->flags.ensure(OKB_FLAGS_BYTES)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory");
    if (!c->loss_ctr_ready) { OKB_CUDA(c, cudaMemsetAsync(c->flags.p, 0, OKB_FLAGS_BYTES, s)); c->loss_ctr_ready = true; }
    return 0;
}

// per-step map table row -> {first, end, slot0, slot1} in the sorted plan, built once per planned chunk
static int ensure_rowhead(okb_ctx *c, cudaStream_t s) {
    if (c->rowhead_ready) return 0;
    const i64 n = c->plan_ne + c->plan_nr, C = c->plan_hi - c->plan_lo, total = C * n, rows = c->E + c->R;
    if (c->rowseg_e.ensure(sizeof(int4) * rows * C)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (rowhead)");
    OKB_CUDA(c, cudaMemsetAsync(c->rowseg_e.p, 0xff, sizeof(int4) * rows * C, s));
    mark_heads_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(c->keys_ent.as<i32>() + total, c->perm_ent.as<i32>(), c->rowseg_e.as<int4>(), (i32)n, (i32)rows, total);
    OKB_LAUNCHED(1);
    c->rowhead_ready = true;
    return 0;
}

extern "C" {

int okb_grad_sizes(okb_ctx *c, const okb_model *m, INT B, INT k, INT kr, INT *er, INT *ec, INT *rr, INT *rc) {
    i32 ce, cr;
    group_cols(m, ce, cr);
    // TransR: one [d rel | d M_r] row per RELATION (gradients come out reduced) — or, with relation negatives, per
    // (positive, relation slot) like the other models (transr_general_kernel)
    *er = B * (2 + k); *ec = ce; *rr = (m->model == OKB_TRANSR && kr == 0) ? c->R : B * (1 + kr); *rc = cr;
    return 0;
}

// Plan steps [step_lo, step_hi) of the sampled batches with ONE radix sort.
static int plan_steps(okb_ctx *c, INT step_lo, INT step_hi, INT b_lo, INT b_hi, void *stream);
int okb_plan_steps(okb_ctx *c, INT step_lo, INT step_hi, void *stream) { return plan_steps(c, step_lo, step_hi, 0, c->B, stream); }
}  // extern "C"
static bool planned(const okb_ctx *c, INT lo, INT hi, INT b_lo, INT b_hi) {
    return lo >= c->plan_lo && hi <= c->plan_hi && c->plan_b_lo == b_lo && c->plan_b_hi == b_hi;
}
bool okb_plan_small_ok(const okb_ctx *c, INT B, INT k, INT kr) {
    const i64 n = B * ((2 + k) + (1 + kr));
    return n <= PS_MAX_N && bits_for(c->E + c->R + 1) <= 16 && !c->plan_multi;
}
static int plan_steps(okb_ctx *c, INT step_lo, INT step_hi, INT b_lo, INT b_hi, void *stream) {
    if (step_lo < 0 || step_hi > c->steps || step_lo >= step_hi) OKB_FAIL(c, OKB_ERR_ARG, "step range out of range (sample first)");
    if (b_lo < 0 || b_hi > c->B || b_lo >= b_hi) OKB_FAIL(c, OKB_ERR_ARG, "bad positive range");
    cudaStream_t s = (cudaStream_t)stream;
    const i64 B = b_hi - b_lo, NE = 2 + c->K, NR = 1 + c->KR, n = B * (NE + NR), S = c->B * (1 + c->K + c->KR);
    const i64 C = step_hi - step_lo, ks = c->E + c->R + 1, total = C * n;
    if (total > 0x3fffffffLL || C > 65535) OKB_FAIL(c, OKB_ERR_ARG, "too many steps planned at once");
    if (c->keys_ent.ensure(sizeof(i32) * total * 2) || c->perm_ent.ensure(sizeof(i32) * total))
        OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (plan)");
    ProfScope ps(c, PROF_PLAN, s);
    PlanArgs a;
    a.batch = c->batch.as<i32>();
    a.keys = c->keys_ent.as<i32>();
    a.B = (i32)c->B; a.k = (i32)c->K; a.kr = (i32)c->KR; a.NE = (i32)NE; a.NR = (i32)NR; a.E = (i32)c->E; a.R = (i32)c->R;
    a.S = (i32)S; a.step_lo = (i32)step_lo; a.C = (i32)C; a.b_lo = (i32)b_lo; a.Bl = (i32)B;
    c->plan_ne = B * NE; c->plan_nr = B * NR; c->plan_lo = step_lo; c->plan_hi = step_hi;
    c->plan_b_lo = b_lo; c->plan_b_hi = b_hi;
    c->rowhead_ready = false;
    const int kb = bits_for(ks);
    // The single-kernel plan: one cluster of 8 CTAs per step.  One step (the host-batch path) — or a whole chunk in ONE launch
    // when its clusters fit the GPU in about two waves (18 clusters at a time on 148 SMs): 20 steps in ~35 us instead of the
    // 11 launches (50 us) of the segmented sort below; longer chunks are planned ahead on the side stream by the sort.
    if ((C == 1 || (C <= 2 * (okb_sms(c) / PS_CTAS) && !c->rowseg_e.external)) && n <= PS_MAX_N && kb <= 16 && !c->plan_multi) {
        const i64 rows = c->E + c->R;
        if (c->rowseg_e.ensure(sizeof(int4) * rows * C)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (rowhead)");
        PlanSmallArgs q;
        q.p = a; q.skeys = a.keys + total; q.perm = c->perm_ent.as<i32>(); q.rowhead = c->rowseg_e.as<int4>();
        q.n = (i32)n; q.rows = (i32)rows; q.ib = bits_for(n); q.db = (kb + 1) / 2 < 5 ? 5 : (kb + 1) / 2;
        i64 chunk = 32;
        while (chunk * PS_CTAS * PS_WARPS < n) chunk *= 2;
        const i64 ndig = 1 << q.db;
        q.chunk = (i32)chunk;
        q.magic_e = (unsigned)(((1ull << 32) + NE - 1) / NE); q.magic_r = (unsigned)(((1ull << 32) + NR - 1) / NR);
        const size_t smem = (size_t)(2 * chunk * PS_WARPS + PS_WARPS * (ndig + 1) + 2 * ndig + PS_WARPS) * 4;
        OKB_CUDA(c, okb_smem_optin(c, plan_small_kernel, 80 * 1024));     // up to 2 x 16 KB of items + 33 KB of counters
        q.mirror = nullptr; q.mirror_flag = nullptr; q.mirror_ctr = nullptr; q.mirror_token = 0;
        q.nsteps = (i32)C;
        if (c->mirror_dst && C == 1 && b_lo == 0 && b_hi == c->B) {                 // okb_sample_to_host: second cluster copies the batch out
            int rc2 = okb_ensure_flags(c, s);
            if (rc2) return rc2;
            q.mirror = c->mirror_dst; q.mirror_flag = (unsigned *)c->host_flag_dev; q.mirror_ctr = c->flags.as<unsigned>() + OKB_FLAGS_MIRRORCTR;
            q.mirror_token = 1u;
        }
        c->mirror_dst = nullptr;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((q.mirror ? C + 1 : C) * PS_CTAS)); cfg.blockDim = dim3(PS_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = c->pdl ? 1 : 0;
        OKB_CUDA(c, cudaLaunchKernelEx(&cfg, plan_small_kernel, q));
        OKB_LAUNCHED(1);
        OKB_CUDA(c, cudaGetLastError());
        c->rowhead_ready = true;
        return 0;
    }
    plan_keys_kernel<<<(unsigned)((C * B + 127) / 128), 128, 0, s>>>(a);
    OKB_LAUNCHED(1);
    int rc = okb_sort_pairs_seg(c, a.keys, a.keys + total, c->perm_ent.as<i32>(), n, C, bits_for(ks), s);
    if (rc) return rc;
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}
extern "C" {
int okb_wait_word(okb_ctx *c, const void *host_word, unsigned sentinel, void *stream) {
    const volatile unsigned *w = (const volatile unsigned *)host_word;
    for (unsigned spins = 1;; spins++) {
        if (*w != sentinel) return 0;
        if ((spins & 0x3ffu) == 0) {                       // every ~1k polls: has the stream finished or failed meanwhile?
            cudaError_t e = cudaStreamQuery((cudaStream_t)stream);
            if (e == cudaSuccess) {
                if (*w != sentinel) return 0;
                OKB_FAIL(c, OKB_ERR_STATE, "stream is idle but the awaited word was never written");
            }
            if (e != cudaErrorNotReady) { OKB_CUDA(c, e); }
        }
    }
}
int okb_train_step(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, float *loss_out, void *stream);
int okb_batch_check(okb_ctx *c, void *stream);
// Config.train_step(batch_h, batch_t, batch_r, batch_y) of the reference (Config.py:464-475) as ONE library call: feed the
// host batch, run the step, return as soon as the loss has arrived in the caller's page-locked word.
int okb_train_step_host(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT B, INT k, INT kr, const INT *h, const INT *t,
                        const INT *r, float *loss_word, void *stream) {
    const unsigned sentinel = 0xFFC0DEADu;                 // a NaN payload no computed loss has
    if (!loss_word) OKB_FAIL(c, OKB_ERR_ARG, "loss_word must be a page-locked host float");
    int rc;
    // Fast path: the caller hands back the block the last okb_sample_to_host filled (the reference's loop does exactly
    // that).  The batch is still resident and already planned, so the step starts at once and a small kernel only VERIFIES
    // that the arrays were not changed in between; a difference makes the update kernels skip the step (NaN loss) and
    // the call falls through to the general path below with the tables untouched.
    if (okb_batch_verify_host(c, B, k, kr, h, t, r, stream) == 0) {
        *(volatile unsigned *)loss_word = sentinel;
        if ((rc = okb_train_step(c, m, hp, 0, loss_word, stream))) return rc;
        if ((rc = okb_wait_word(c, loss_word, sentinel, stream))) return rc;
        if (*(volatile float *)loss_word == *(volatile float *)loss_word) return 0;
        rc = okb_batch_check(c, stream);
        if (rc == 0) return 0;                             // a genuine NaN loss (NaN parameters)
        if (rc != OKB_ERR_STATE) return rc;
    }
    if ((rc = okb_batch_from_host(c, B, k, kr, h, t, r, stream))) return rc;
    *(volatile unsigned *)loss_word = sentinel;
    if ((rc = okb_train_step(c, m, hp, 0, loss_word, stream))) return rc;
    if ((rc = okb_wait_word(c, loss_word, sentinel, stream))) return rc;
    if (*(volatile float *)loss_word != *(volatile float *)loss_word) return okb_batch_check(c, stream);   // NaN: an id was out of range?
    return 0;
}
int okb_plan(okb_ctx *c, INT step, void *stream) {
    if (planned(c, step, step + 1, 0, c->B)) return 0;          // already planned as part of a chunk
    return okb_plan_steps(c, step, step + 1, stream);
}

}  // extern "C"
// The grad kernels' argument block for positives [b_lo, b_hi) of one sampled step (no launch)
void okb_fill_grad(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, INT b_lo, INT b_hi, INT slot_base, float *gent,
                   float *grel, float *loss_terms, GradArgs &a) {
    const i64 S = c->B * (1 + c->K + c->KR);
    a.m = *m;
    const i32 *base = c->batch.as<i32>() + step * 3 * S;
    a.bh = base; a.bt = base + S; a.br = base + 2 * S;
    a.gent = gent; a.grel = grel; a.loss_terms = loss_terms;
    a.margin = hp->margin; a.w = 1.0f / (float)(c->B * (c->K + c->KR));
    a.B = (i32)c->B; a.k = (i32)c->K; a.kr = (i32)c->KR; a.NE = (i32)(2 + c->K); a.NR = (i32)(1 + c->KR);
    a.b_lo = (i32)b_lo; a.b_hi = (i32)b_hi; a.slot_base = (i32)slot_base;
    a.wait_flags = nullptr; a.wait_epoch = 0; a.wait_n = 0; a.npf = 0;
    a.vh = nullptr; a.vflag = nullptr; a.vS = 0; a.vblocks = 0;
    a.sc_world = 0; a.sc_gather = 0; a.trace = nullptr; a.hs_mode = c->dp_hs_mode;
}
// warps per positive the generic grad kernel uses for this batch (1 = one warp per positive); a function of the GLOBAL
// batch, so that a data-parallel rank's slice is computed exactly like the same positives on one GPU
int okb_grad_wpp(okb_ctx *c) {
    const bool k1 = c->K == 1 && c->KR == 0 && !c->grad_generic;
    if (k1 || c->K < 2 || c->grad_single_warp) return 1;
    return (int)std::max<i64>(1, std::min<i64>(std::min<i64>(4, c->K), (i64)okb_sms(c) * 20 / std::max<i64>(1, c->B)));
}
static int launch_grad(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, INT b_lo, INT b_hi, INT slot_base, float *gent,
                       float *grel, float *loss_terms, const unsigned long long *wait_flags, unsigned long long wait_epoch, int wait_n,
                       void *stream) {
    int vw, nv;
    int rc = check_model(c, m, vw, nv);
    if (rc) return rc;
    if (step < 0 || step >= c->steps) OKB_FAIL(c, OKB_ERR_ARG, "step out of range (sample first)");
    if (b_lo < 0 || b_hi > c->B || b_lo > b_hi) OKB_FAIL(c, OKB_ERR_ARG, "bad positive range");
    if (b_lo == b_hi) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const i64 S = c->B * (1 + c->K + c->KR);
    if (m->model == OKB_TRANSR) {                          // relation-bucketed kernel; needs the plan's relation segments
        if ((rc = okb_verify_flush(c, stream))) return rc;
        if (slot_base || wait_flags) OKB_FAIL(c, OKB_ERR_ARG, "TransR is not supported by the owner-sharded data-parallel path");
        if (!planned(c, step, step + 1, 0, c->B)) OKB_FAIL(c, OKB_ERR_STATE, "step has not been planned (okb_plan / okb_plan_steps)");
        if ((rc = ensure_rowhead(c, s))) return rc;
        const i64 n = c->plan_ne + c->plan_nr, total = (c->plan_hi - c->plan_lo) * n, rel = (step - c->plan_lo) * n;
        return okb_transr_launch_grad(c, m, hp, c->batch.as<i32>() + step * 3 * S, c->keys_ent.as<i32>() + total + rel,
                                      c->perm_ent.as<i32>() + rel, c->rowseg_e.as<int4>() + (step - c->plan_lo) * (c->E + c->R), n,
                                      b_lo, b_hi, gent, grel, loss_terms, s);
    }
    GradArgs a;
    okb_fill_grad(c, m, hp, step, b_lo, b_hi, slot_base, gent, grel, loss_terms, a);
    a.wait_flags = wait_flags; a.wait_epoch = wait_epoch; a.wait_n = wait_n;
    if (wait_flags)
        for (int q = 0; q < wait_n; q++)
            a.announce[q] = (unsigned long long *)((char *)c->dp.arena[q] + c->dp.off_flags) + DP_FLAG_X + c->dp.rank;
    if (c->sc_launch) {                                    // scatter form: gradient rows go to their table rows' owners (global
        const okb_dp &P = c->dp;                           // slots); gather form: to every rank
        const i64 own_e = (c->E + P.world - 1) / P.world, own_r = (c->R + P.world - 1) / P.world;
        a.sc_world = P.world; a.sc_gather = c->sc_launch == 2;
        for (int q = 0; q < P.world; q++) { a.sc_arena[q] = (char *)P.arena[q]; a.sc_delta[q] = (i64)((char *)P.arena[q] - (char *)P.arena[0]); }
        for (int q = 0; q <= P.world; q++) { a.sc_ent_lo[q] = (i32)std::min<i64>(c->E, q * own_e); a.sc_rel_lo[q] = (i32)std::min<i64>(c->R, q * own_r); }
        a.sc_gent = P.off_gent + c->sc_buf_off[0]; a.sc_grel = P.off_grel + c->sc_buf_off[1]; a.sc_loss = P.off_loss + c->sc_buf_off[2];
    }
    if (c->dp_on && c->dp_trace_on) a.trace = c->dp_trace.as<unsigned long long>() + ((c->dp_epoch + 1) % 64) * 16;
    const unsigned grid = (unsigned)((b_hi - b_lo + GRAD_WARPS - 1) / GRAD_WARPS);
    a.npf = 0;
    if (m->optimizer == OKB_ADAM && c->l2_prefetch && m->m_ent) {
        auto reg = [&](const float *p, i64 elems) {
            if (!p || a.npf >= 12) return;
            const i64 bytes = elems * 4;
            if (bytes >= 0xffffffffLL) return;             // >4 GB regions do not fit the L2 anyway
            a.pf_ptr[a.npf] = (const char *)p; a.pf_bytes[a.npf] = (unsigned)bytes;
            a.pf_slice[a.npf] = (unsigned)((((bytes + (b_hi - b_lo) - 1) / (b_hi - b_lo)) + 127) & ~(i64)127);
            a.npf++;
        };
        const i64 ne = c->E * m->ent_dim, nr = c->R * m->rel_dim;
        reg(m->m_ent, ne); reg(m->v_ent, ne); reg(m->ent, ne);
        if (m->model == OKB_TRANSD) { reg(m->m_ent_aux, ne); reg(m->v_ent_aux, ne); reg(m->ent_aux, ne); }
        reg(m->m_rel, nr); reg(m->v_rel, nr); reg(m->rel, nr);
        if (m->model != OKB_TRANSE) { reg(m->m_rel_aux, nr); reg(m->v_rel_aux, nr); reg(m->rel_aux, nr); }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GRAD_WARPS * 32); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = c->pdl ? 1 : 0;
    // sc: the variant with the scatter form's owner search and the phase stamps compiled in.  The k = 1 kernel (the default
    // batch, the bench path) exists in both variants — without that code it is 200 instructions (1.3 us) shorter, and the two
    // compile to the same floating-point instructions (tests/dist_gpu_check.py compares them bit for bit); the generic
    // kernels have ONE variant, because there the compiler's FMA contraction was seen to differ between the two.
    const bool sc = a.sc_world > 0 || a.trace != nullptr;
#define CALL_GRAD_(VW, NV, W, SC)                                                                          \
    if (m->model == OKB_TRANSE) cudaLaunchKernelEx(&cfg, grad_kernel<OKB_TRANSE, VW, NV, W, SC>, a);       \
    else if (m->model == OKB_TRANSH) cudaLaunchKernelEx(&cfg, grad_kernel<OKB_TRANSH, VW, NV, W, SC>, a);  \
    else cudaLaunchKernelEx(&cfg, grad_kernel<OKB_TRANSD, VW, NV, W, SC>, a)
#define CALL_GRAD(VW, NV) CALL_GRAD_(VW, NV, 1, true)
#define CALL_GRADW(VW, NV) CALL_GRAD_(VW, NV, 4, true)
#define CALL_GRAD1_(VW, NV, SC)                                                                            \
    if (m->model == OKB_TRANSE) cudaLaunchKernelEx(&cfg, grad_k1_kernel<OKB_TRANSE, VW, NV, SC>, a);       \
    else if (m->model == OKB_TRANSH) cudaLaunchKernelEx(&cfg, grad_k1_kernel<OKB_TRANSH, VW, NV, SC>, a);  \
    else cudaLaunchKernelEx(&cfg, grad_k1_kernel<OKB_TRANSD, VW, NV, SC>, a)
#define CALL_GRAD1(VW, NV) if (sc) { CALL_GRAD1_(VW, NV, true); } else { CALL_GRAD1_(VW, NV, false); }
    const bool k1 = c->K == 1 && c->KR == 0 && !c->grad_generic && a.npf == 0;
    if (k1) { cfg.gridDim = dim3((unsigned)(b_hi - b_lo)); cfg.blockDim = dim3(32); }
    // small batches with several negatives: 2..4 warps per positive, as many as fill ~20 warp slots per SM
    const int wpp = k1 ? 1 : okb_grad_wpp(c);
    if (wpp > 1) {
        const int N = vw * nv, F = 2 * (m->model == OKB_TRANSD ? 2 : 1) + (m->model == OKB_TRANSE ? 1 : 2);
        cfg.gridDim = dim3((unsigned)(b_hi - b_lo)); cfg.blockDim = dim3(32 * wpp);
        cfg.dynamicSmemBytes = sizeof(float) * (size_t)(wpp - 1) * (F * 32 * N + 64);
    }
    if (c->verify_h && step == 0 && b_lo == 0 && b_hi == c->B && !wait_flags) {
        // host-batch fast path: compare the caller's block with the resident batch in extra blocks of this launch
        a.vh = (const long long *)c->verify_h; a.vS = (i32)c->verify_S; a.vflag = c->flags.as<unsigned>() + OKB_FLAGS_BAD;
        a.vblocks = (i32)((c->verify_S + cfg.blockDim.x - 1) / cfg.blockDim.x);      // one batch row per thread
        cfg.gridDim = dim3(cfg.gridDim.x + (unsigned)a.vblocks);
        c->verify_h = nullptr;
    } else if ((rc = okb_verify_flush(c, stream))) return rc;
    { ProfScope ps(c, PROF_GRAD, s);
      if (k1) { DISPATCH_LAYOUT(vw, nv, CALL_GRAD1); } else if (wpp > 1) { DISPATCH_LAYOUT(vw, nv, CALL_GRADW); } else { DISPATCH_LAYOUT(vw, nv, CALL_GRAD); } }
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}
extern "C" {
int okb_grad(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, INT b_lo, INT b_hi, float *gent, float *grel,
             float *loss_terms, void *stream) {
    return launch_grad(c, m, hp, step, b_lo, b_hi, 0, gent, grel, loss_terms, nullptr, 0, 0, stream);
}

}  // extern "C"
// Everything the update kernels of one step need (no launch): plan pointers, gradient buffers, hub decision + its
// partial-sum buffer, loss reduction scratch and — for Adam — the dense table list.  blk: 256-vector tiles of the Adam pass.
int okb_fill_update(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, const float *gent, const float *grel,
                    const float *loss_terms, float *loss_out, int vw, cudaStream_t s, UpdArgs &a, i32 &blk, bool &lean) {
    int rc;
    const bool is_tr = m->model == OKB_TRANSR;
    const i64 n = c->plan_ne + c->plan_nr;
    if (n == 0 || !planned(c, step, step + 1, 0, c->B)) OKB_FAIL(c, OKB_ERR_STATE, "step has not been planned (okb_plan / okb_plan_steps)");
    const i64 total = (c->plan_hi - c->plan_lo) * n, rel = (step - c->plan_lo) * n;
    a.m = *m; a.hp = *hp;
    a.skeys = c->keys_ent.as<i32>() + total + rel; a.perm = c->perm_ent.as<i32>() + rel;
    a.gent = gent; a.grel = grel; a.loss_terms = loss_terms; a.loss_out = loss_out;
    a.n = (i32)n; a.n_ent_slots = (i32)c->plan_ne; a.E = (i32)c->E; a.R = (i32)c->R; a.B = (i32)c->B;
    a.w = 1.0f / (float)(c->B * (c->K + c->KR));
    group_cols(m, a.ce, a.cr);
    a.ntab = 0;
    a.key_limit = is_tr ? (i32)c->E : (i32)(c->E + c->R);
    if ((rc = ensure_rowhead(c, s))) return rc;
    a.rowhead = c->rowseg_e.as<int4>() + (step - c->plan_lo) * (c->E + c->R);
    // hub heuristic from load-time statistics: expected longest segment of this batch
    a.pcols = std::max(a.ce, is_tr ? 0 : a.cr);
    a.hub = (double)c->plan_ne * c->max_ent_share > PCH || (!is_tr && (double)c->plan_nr * c->max_rel_share > PCH);
    a.partial = nullptr; a.by_row = 0;
    a.loss_blocks = (i32)std::max<i64>(1, std::min<i64>(64, c->B / 8192));
    if ((rc = okb_ensure_flags(c, s))) return rc;
    a.loss_part = c->flags.as<float>(); a.loss_ctr = (unsigned *)(c->flags.as<float>() + 64);
    a.bad = c->batch_from_host ? c->flags.as<unsigned>() + OKB_FLAGS_BAD : nullptr;
    if (a.hub) {
        const i64 nblocks = n / PCH;
        if (c->partial.ensure(sizeof(float) * (size_t)(nblocks + 1) * a.pcols)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (partial sums)");
        a.partial = c->partial.as<float>();
    }
    a.work_blocks = (i32)((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    blk = 0;
    lean = !c->adam_legacy;
    if (m->optimizer == OKB_ADAM) {
        if (!m->m_ent || !m->v_ent || !m->m_rel || !m->v_rel) OKB_FAIL(c, OKB_ERR_ARG, "Adam slots missing");
        i64 acc = 0;
        i32 sblk = 0;
        auto add = [&](float *x, float *mm, float *vv, i64 nrows, int D, bool is_ent, int part) {
            if (!x) return;
            acc += nrows * D / vw;
            blk += (i32)((nrows * D / vw + ADAM_TILE_V - 1) / ADAM_TILE_V);
            sblk += (i32)((nrows * D / vw + ADAM_TILE_V * c->adam_vpt - 1) / (ADAM_TILE_V * c->adam_vpt));
            DenseTab T;
            T.x = x; T.m = mm; T.v = vv; T.grad = is_ent ? gent : grel; T.vec_end = acc; T.D = D;
            T.key_off = is_ent ? 0 : (i32)c->E; T.cols = is_ent ? a.ce : a.cr; T.part = part;
            T.slot_off = is_ent ? 0 : (i32)c->plan_ne; T.blk_end = blk; T.sblk_end = sblk;
            {   // row = vector / vpr for vector < 2^31: multiply-high by ceil(2^(32+s) / vpr), s = floor(log2 vpr)
                const unsigned vpr = (unsigned)(D / vw);
                unsigned sh = 0;
                while ((2u << sh) <= vpr) sh++;
                if ((1u << sh) == vpr) { T.magic = 0; T.shift = sh; }
                else { T.magic = (unsigned)((((unsigned long long)1 << (32 + sh)) + vpr - 1) / vpr); T.shift = sh; }
                if (nrows * D / vw >= 0x7fffffffLL) lean = false;
            }
            a.tab[a.ntab++] = T;
        };
        add(m->ent, m->m_ent, m->v_ent, c->E, m->ent_dim, true, 0);
        if (m->model == OKB_TRANSD) add(m->ent_aux, m->m_ent_aux, m->v_ent_aux, c->E, m->ent_dim, true, 1);
        if (!is_tr) add(m->rel, m->m_rel, m->v_rel, c->R, m->rel_dim, false, 0);
        if (!is_tr && m->model != OKB_TRANSE) add(m->rel_aux, m->m_rel_aux, m->v_rel_aux, c->R, m->rel_dim, false, 1);
#ifdef EXP_ADAM_BLOCKS
        a.work_blocks = (i32)std::min<i64>((acc + 255) / 256, (i64)EXP_ADAM_BLOCKS);
#else
        a.work_blocks = (i32)std::min<i64>((acc + 255) / 256, (i64)okb_sms(c) * 16);
#endif
    }
    return 0;
}
extern "C" {
int okb_update(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, const float *gent, const float *grel,
               const float *loss_terms, float *loss_out, void *stream) {
    int vw, nv;
    int rc = check_model(c, m, vw, nv);
    if (rc) return rc;
    const bool is_tr = m->model == OKB_TRANSR;
    cudaStream_t s = (cudaStream_t)stream;
    const i64 n = c->plan_ne + c->plan_nr;
    UpdArgs a;
    i32 blk = 0;
    bool lean = true;
    if ((rc = okb_fill_update(c, m, hp, step, gent, grel, loss_terms, loss_out, vw, s, a, blk, lean))) return rc;
    if (a.hub) {
        const i64 nblocks = n / PCH;
        const unsigned pg = (unsigned)((nblocks + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK + 1);
#define CALL_PRE(VW, NV) prereduce_kernel<VW, NV><<<pg, WARPS_PER_BLOCK * 32, 0, s>>>(a, c->partial.as<float>())
        DISPATCH_LAYOUT(vw, nv, CALL_PRE);
        OKB_LAUNCHED(1);
    }
    if (m->optimizer == OKB_ADAM) {
        ProfScope ps(c, PROF_UPDATE, s);
        // programmatic dependent launch: the kernel may start while the grad kernel drains (see adam_kernel)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(a.work_blocks + a.loss_blocks)); cfg.blockDim = dim3(256); cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = c->pdl ? 1 : 0;
        if (vw == 4 && c->adam_tma) {                      // pipelined TMA ring (see adam_pipe_kernel)
            const size_t smem = sizeof(ApStage) * AP_STAGES + 8 * AP_STAGES;
            OKB_CUDA(c, okb_smem_optin(c, adam_pipe_kernel, smem));
            a.work_blocks = std::min<i32>(blk, okb_sms(c) * 2);
            cfg.gridDim = dim3((unsigned)(a.work_blocks + a.loss_blocks)); cfg.dynamicSmemBytes = smem;
            OKB_CUDA(c, cudaLaunchKernelEx(&cfg, adam_pipe_kernel, a, (int)blk));
        } else if (lean) {                                 // one super tile per CTA, block-uniform table (adam_tile_kernel)
            a.work_blocks = a.tab[a.ntab - 1].sblk_end;
            cfg.gridDim = dim3((unsigned)(a.work_blocks + a.loss_blocks));
#define CALL_TILE(VW)                                                                                   \
            if (c->adam_vpt == 1) OKB_CUDA(c, cudaLaunchKernelEx(&cfg, adam_tile_kernel<VW, 1>, a));         \
            else if (c->adam_vpt == 2) OKB_CUDA(c, cudaLaunchKernelEx(&cfg, adam_tile_kernel<VW, 2>, a));    \
            else if (c->adam_vpt == 3) OKB_CUDA(c, cudaLaunchKernelEx(&cfg, adam_tile_kernel<VW, 3>, a));    \
            else OKB_CUDA(c, cudaLaunchKernelEx(&cfg, adam_tile_kernel<VW, 4>, a))
            if (vw == 4) { CALL_TILE(4); } else if (vw == 2) { CALL_TILE(2); } else { CALL_TILE(1); }
        } else if (vw == 4) OKB_CUDA(c, cudaLaunchKernelEx(&cfg, adam_kernel<4>, a));
        else if (vw == 2) OKB_CUDA(c, cudaLaunchKernelEx(&cfg, adam_kernel<2>, a));
        else OKB_CUDA(c, cudaLaunchKernelEx(&cfg, adam_kernel<1>, a));
        OKB_LAUNCHED(1);
    } else {
        a.by_row = a.key_limit < n;                       // fewer table rows than gradient rows: one warp per table row
        a.work_blocks = (i32)(((a.by_row ? (i64)a.key_limit : n) + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
        ProfScope ps(c, PROF_UPDATE, s);
        // programmatic dependent launch, like the Adam pass: row map + table rows are requested while the grad kernel drains
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(a.work_blocks + a.loss_blocks)); cfg.blockDim = dim3(WARPS_PER_BLOCK * 32); cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = (c->pdl && !a.hub) ? 1 : 0;     // after the hub pre-reduction kernel: plain stream order
#define CALL_SGD(VW, NV) OKB_CUDA(c, cudaLaunchKernelEx(&cfg, sgd_kernel<VW, NV>, a))
        DISPATCH_LAYOUT(vw, nv, CALL_SGD);
        OKB_LAUNCHED(1);
    }
    if (is_tr) {                                          // rel_embeddings + transfer_matrix rows (gradients already per relation)
        if ((rc = ensure_rowhead(c, s))) return rc;
        if ((rc = okb_transr_launch_rel_update(c, m, hp, c->rowseg_e.as<int4>() + (step - c->plan_lo) * (c->E + c->R), grel, s,
                                               c->perm_ent.as<i32>() + (step - c->plan_lo) * n))) return rc;
    }
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

int okb_train_step(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, float *loss_out, void *stream) {
    INT er, ec, rr, rcn;
    int rc = okb_grad_sizes(c, m, c->B, c->K, c->KR, &er, &ec, &rr, &rcn);
    if (rc) return rc;
    if (c->gent.ensure(sizeof(float) * er * ec) || c->grel.ensure(sizeof(float) * rr * rcn) ||
        c->lossterms.ensure(sizeof(float) * c->B))
        OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (gradient rows)");
    if (!planned(c, step, step + 1, 0, c->B)) { if ((rc = okb_plan(c, step, stream))) return rc; }
    if ((rc = okb_grad(c, m, hp, step, 0, c->B, c->gent.as<float>(), c->grel.as<float>(), c->lossterms.as<float>(), stream))) return rc;
    return okb_update(c, m, hp, step, c->gent.as<float>(), c->grel.as<float>(), c->lossterms.as<float>(), loss_out, stream);
}


// `n` consecutive train steps (steps step_lo .. step_lo+n-1 of the sampled batches) in one call: the
// host-side loop of distribute_training.py:267-283 without a Python round trip per step.
// hp[i] are the per-step hyper-parameters (Adam's lr_t changes every step); loss_out[i] (device) receives
// each step's loss, or pass NULL.
int okb_train_steps(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step_lo, INT n, float *loss_out, void *stream) {
    if (step_lo < 0 || n < 1 || step_lo + n > c->steps) OKB_FAIL(c, OKB_ERR_ARG, "step range out of range (sample first)");
    if (!planned(c, step_lo, step_lo + n, 0, c->B)) {
        int rc = okb_plan_steps(c, step_lo, step_lo + n, stream);
        if (rc) return rc;
    }
    {   // the whole chunk as one persistent kernel where covered (chunk.cu); -1 = not covered: per-phase kernels below
        INT er, ec, rr, rcn;
        int rc = okb_grad_sizes(c, m, c->B, c->K, c->KR, &er, &ec, &rr, &rcn);
        if (rc) return rc;
        if (c->gent.ensure(sizeof(float) * er * ec) || c->grel.ensure(sizeof(float) * rr * rcn) || c->lossterms.ensure(sizeof(float) * c->B))
            OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (gradient rows)");
        rc = okb_chunk_kernel_steps(c, m, hp, step_lo, n, loss_out, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    for (INT i = 0; i < n; i++) {
        int rc = okb_train_step(c, m, hp + i, step_lo + i, loss_out ? loss_out + i : nullptr, stream);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ data parallel, owner-sharded
// One process per GPU inside a box.  Every rank holds the full tables (the gathers of the grad kernel stay local) but
// OWNS the update of a contiguous range of rows: its Adam slots are only maintained for that range.  In the PUSH form
// (described first; the scatter form follows further down, see DpSc) a step is
//
//   grad         local positives only (this rank's sampler streams)            -> local gradient rows
//   reduce+push  one warp per table row: fixed-order segment sum of the local gradient rows of that row, stored
//                straight into the OWNER's staging slab over NVLink (peer stores), slot [this rank][row]
//   owner update waits for every rank's slab, sums the `world` partial rows in rank order, applies SGD / TF1-Adam to
//                its own rows and stores the new row into EVERY rank's table (peer stores)
//
// i.e. the reduce-scatter and all-gather of a synchronous data-parallel step are fused into the two update kernels
// as plain stores into peer memory; per rank and step (N-1)/N of one dense gradient leaves and (N-1)/N of one table
// arrives, independent of N, while the dense Adam pass shrinks to 1/N of the rows.  Cross-GPU ordering: each kernel
// ends with "last block publishes the epoch in every peer's flag word" (release, system scope) and the consumer
// kernel's blocks spin on their local flag words (acquire) before touching peer-written data.  Sums are taken in
// rank order, so all replicas stay bit-identical to each other (not to the single-GPU order: (a+b)+(c+d)).

struct DpPush {
    char *arena[OKB_DP_MAX];
    i64 off_stage_ent, off_stage_rel, off_flags;
    i32 ent_lo[OKB_DP_MAX + 1], rel_lo[OKB_DP_MAX + 1];
    i32 own_max_ent, own_max_rel, world, rank, Bl;
    unsigned long long epoch;
    unsigned long long *trace;
    i32 hs_mode;
};
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Signalling rides on kernel boundaries: all peer stores of a kernel are performed when the grid completes, so the
// NEXT kernel in the stream announces it — block 0 publishes "my previous kernel is done" (release, system scope) in
// every peer's flag word, then every block waits (acquire) until all ranks have announced the same epoch.  No fences
// or tickets inside the producing kernels.
__device__ __forceinline__ void dp_announce_and_wait(char *const *arena, i64 off_flags, int world, int rank, int which,
                                                     unsigned long long epoch, bool announce, int hs_mode = 0) {
    if ((int)threadIdx.x < world) {
        if (announce) st_flag_sys((unsigned long long *)(arena[threadIdx.x] + off_flags) + which + rank, epoch, hs_mode & 2);
        spin_until((const unsigned long long *)(arena[rank] + off_flags) + which + threadIdx.x, epoch, hs_mode & 1);
    }
    __syncthreads();
}

template <int VW, int NV>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) dp_reduce_push_kernel(UpdArgs a, DpPush d) {
    constexpr int N = VW * NV;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    if (threadIdx.x == 0) trace_min(d.trace, 10);
    if ((i32)blockIdx.x >= a.work_blocks) {                // the extra block: this rank's hinge sum, in a fixed order
        __shared__ float sh[WARPS_PER_BLOCK];
        pdl_wait();
        float x = 0.f;
        for (i32 i = threadIdx.x; i < d.Bl; i += WARPS_PER_BLOCK * 32) x += a.loss_terms[i];
        x = wsum(x);
        if (lane == 0) sh[threadIdx.x >> 5] = x;
        __syncthreads();
        if ((int)threadIdx.x < d.world) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < WARPS_PER_BLOCK; w++) t += sh[w];
            ((float *)((unsigned long long *)(d.arena[threadIdx.x] + d.off_flags) + DP_FLAG_LOSS))[d.rank] = t;
        }
    } else {
        const i32 key = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
        if (key < a.key_limit) {
            const bool is_ent = key < a.E;
            const int D = is_ent ? a.m.ent_dim : a.m.rel_dim;
            const i32 row = is_ent ? key : key - a.E;
            const i32 *lo = is_ent ? d.ent_lo : d.rel_lo;
            int o = 0;
            while (row >= lo[o + 1]) o++;
            const i32 cols = is_ent ? a.ce : a.cr, own_max = is_ent ? d.own_max_ent : d.own_max_rel;
            float *dst = (float *)(d.arena[o] + (is_ent ? d.off_stage_ent : d.off_stage_rel)) +
                         ((i64)d.rank * own_max + (row - lo[o])) * cols;
            const int4 seg = __ldg(a.rowhead + key);
            const int parts = cols / D;
            pdl_wait();                                    // the grad kernel's rows are complete from here on
            if (threadIdx.x == 0) trace_min(d.trace, 11);
            for (int p = 0; p < parts; p++) {
                Frag<VW, NV> f;
                f.zero();
                if (seg.x >= 0) seg_sum_map<VW, NV>(a, seg, is_ent, D, p, lane, f.v);
                f.store(dst + p * D, D, lane);
            }
        }
        if (threadIdx.x == 0) trace_max(d.trace, 12);
    }
}

struct DpTab {
    float *m, *v;              // local Adam slots (full-size arrays; only the owned rows are maintained)
    i64 x_off, stage_off;      // byte offsets in an arena: the table / this table's staging slab
    i64 vec_end;               // cumulative vector count over the OWNED rows of the table list
    i32 D, cols, part, row_lo, own_max;
    i32 blk_end;               // cumulative count of 256-vector tiles (one CTA each)
    unsigned magic, shift;     // owned-row index = vector / (D / VW) as a multiply-high (see DenseTab)
};
struct DpOwn {
    okb_hyper hp;
    char *arena[OKB_DP_MAX];
    DpTab tab[4];
    i64 off_flags;
    i32 ntab, world, rank, adam;
    unsigned long long epoch;
    float *loss_out;
    float w;
    unsigned long long *trace;
    i32 hs_mode;
};
// One 256-vector tile of ONE table's owned rows per CTA (block-uniform table, multiply-high row split — the lean form of
// adam_tile_kernel): x / m / v of the tile are requested BEFORE the wait for this rank's push kernel and for the peers'
// "stage ready" flags — none of them is written by anyone but this owner — so their HBM latency runs under the flag wait;
// the world's partial rows are then loaded four at a time and added in rank order.
template <int VW>
__global__ void __launch_bounds__(256) dp_owner_kernel(DpOwn d) {
    typedef typename VecT<VW>::T V;
    pdl_launch_dependents();
    const i32 bid = (i32)blockIdx.x;
    int t = 0;
    while (bid >= d.tab[t].blk_end) t++;                   // block-uniform
    const DpTab &T = d.tab[t];
    const unsigned nvec = (unsigned)(T.vec_end - (t ? d.tab[t - 1].vec_end : 0));
    const unsigned lv = ((unsigned)bid - (unsigned)(t ? d.tab[t - 1].blk_end : 0)) * 256u + threadIdx.x;
    const bool live = lv < nvec;
    const unsigned vpr = (unsigned)T.D / VW;
    const unsigned rl = T.magic ? (__umulhi(lv, T.magic) >> T.shift) : (lv >> T.shift), col = (lv - rl * vpr) * VW;
    const i64 e = ((i64)T.row_lo + rl) * T.D + col;         // element index inside the full table
    const char *own = d.arena[d.rank];
    V xv, mv, vv;
    if (live) {
        xv = *reinterpret_cast<const V *>(own + T.x_off + e * 4);
        if (d.adam) { mv = *reinterpret_cast<const V *>(T.m + e); vv = *reinterpret_cast<const V *>(T.v + e); }
    }
    if (threadIdx.x == 0) trace_min(d.trace, 5);
    pdl_wait();                                            // this rank's reduce+push kernel is complete: announce it
    if (threadIdx.x == 0) trace_min(d.trace, 6);
    dp_announce_and_wait(d.arena, d.off_flags, d.world, d.rank, DP_FLAG_STAGE, d.epoch, blockIdx.x == 0, d.hs_mode);
    if (threadIdx.x == 0) trace_min(d.trace, 7);
    if (blockIdx.x == 0 && threadIdx.x == 0 && d.loss_out) {   // mean hinge over the GLOBAL batch, rank order
        const volatile float *lp = (const volatile float *)((unsigned long long *)(d.arena[d.rank] + d.off_flags) + DP_FLAG_LOSS);
        float tot = 0.f;
        for (int q = 0; q < d.world; q++) tot += lp[q];
        d.loss_out[0] = tot * d.w;
    }
    if (!live) return;
    const float b1 = d.hp.beta1, b2 = d.hp.beta2, lr = d.hp.lr, eps = d.hp.eps;
    const float *st = (const float *)(own + T.stage_off) + (i64)rl * T.cols + T.part * T.D + col;
    float g[VW];
#pragma unroll
    for (int q = 0; q < VW; q++) g[q] = 0.f;
    bool any = false;
    for (int p0 = 0; p0 < d.world; p0 += 4) {              // partial rows in rank order; four loads in flight at a time
        V gp[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (p0 + u < d.world) gp[u] = __ldcg(reinterpret_cast<const V *>(st + (i64)(p0 + u) * T.own_max * T.cols));
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (p0 + u >= d.world) break;
            const float *pg = reinterpret_cast<const float *>(&gp[u]);
#pragma unroll
            for (int q = 0; q < VW; q++) { g[q] += pg[q]; any |= pg[q] != 0.f; }
        }
    }
    float *xs = reinterpret_cast<float *>(&xv);
    if (d.adam) {
        float *ms = reinterpret_cast<float *>(&mv), *vs = reinterpret_cast<float *>(&vv);
#pragma unroll
        for (int q = 0; q < VW; q++) adam_elem(xs[q], ms[q], vs[q], g[q], b1, b2, 1.f - b1, 1.f - b2, lr, eps);
        *reinterpret_cast<V *>(T.m + e) = mv; *reinterpret_cast<V *>(T.v + e) = vv;
    } else {
        if (!any) return;                                  // SGD leaves rows without gradient untouched: nothing to publish
#pragma unroll
        for (int q = 0; q < VW; q++) xs[q] -= lr * g[q];
    }
    for (int p = 0; p < d.world; p++) *reinterpret_cast<V *>(d.arena[p] + T.x_off + e * 4) = xv;
    if (threadIdx.x == 0) trace_max(d.trace, 9);
}

// End of a library call: publish "my owner-update kernels up to `epoch` are complete" without waiting for the next step's
// grad kernel to do it, so that a peer can find out ON ITS OWN (okb_dp_quiesce) when its tables are complete.
__global__ void dp_announce_kernel(DpPush d) {
    if ((int)threadIdx.x < d.world)
        st_release_sys((unsigned long long *)(d.arena[threadIdx.x] + d.off_flags) + DP_FLAG_X + d.rank, d.epoch);
}
__global__ void dp_wait_kernel(DpPush d) {
    if ((int)threadIdx.x < d.world)
        spin_until((const unsigned long long *)(d.arena[d.rank] + d.off_flags) + DP_FLAG_X + threadIdx.x, d.epoch);
}

// ---- "pull" form of the owner update: no reduce+push kernel.  Every rank's plan of the step (row map + slot permutation),
// its gradient rows and its loss terms live in its peer arena; the OWNER of a table row walks each rank's segment of that
// row with peer loads (ld.global.cv over NVLink), adds the per-rank sums in rank order — bit-identical to the push form —
// applies SGD / TF1-Adam and publishes the row.  A step is then two kernels (grad, this one) and two flag barriers.
// Hub rows (segments longer than PCH) keep the push form, whose pre-reduced block sums are local.
struct PullTab {
    float *m, *v;              // local Adam slots (full-size arrays; only the owned rows are maintained)
    i64 x_off;                 // byte offset of the table in an arena
    i64 vec_end;               // cumulative vector count over the OWNED rows of the table list
    i32 D, cols, part, row_lo, key_off, is_ent;
};
struct DpPull {
    okb_hyper hp;
    char *arena[OKB_DP_MAX];
    PullTab tab[4];
    i64 off_flags, off_rowhead, off_perm, off_gent, off_grel, off_loss;
    i32 nloc[OKB_DP_MAX], nes[OKB_DP_MAX], bl[OKB_DP_MAX];   // per rank: plan entries per step, entity slots, positives
    i32 ntab, world, rank, adam, rows_all, step_rel, work_blocks;
    unsigned long long epoch;
    float *loss_out;
    float w;
};
template <int VW>
__global__ void __launch_bounds__(256) dp_pull_kernel(DpPull d) {
    typedef typename VecT<VW>::T V;
    pdl_launch_dependents();
    pdl_wait();                                            // this rank's grad kernel is complete: announce it
    dp_announce_and_wait(d.arena, d.off_flags, d.world, d.rank, DP_FLAG_STAGE, d.epoch, blockIdx.x == 0);
    if ((i32)blockIdx.x >= d.work_blocks) {                // mean hinge over the GLOBAL batch: per-rank sums in a fixed order
        if (!d.loss_out) return;
        __shared__ float sh[8];
        float tot = 0.f;
        for (int p = 0; p < d.world; p++) {
            const float *lt = (const float *)(d.arena[p] + d.off_loss);
            float x = 0.f;
            for (i32 i = threadIdx.x; i < d.bl[p]; i += 256) x += __ldcv(lt + i);
            x = wsum(x);
            if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
            __syncthreads();
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) t += sh[w];
            tot += t;
            __syncthreads();
        }
        if (threadIdx.x == 0) d.loss_out[0] = tot * d.w;
        return;
    }
    const i64 total = d.tab[d.ntab - 1].vec_end;
    const float b1 = d.hp.beta1, b2 = d.hp.beta2, lr = d.hp.lr, eps = d.hp.eps;
    const char *own = d.arena[d.rank];
    for (i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (i64)d.work_blocks * blockDim.x) {
        int t = 0;
        while (v >= d.tab[t].vec_end) t++;
        const PullTab &T = d.tab[t];
        const unsigned lv = (unsigned)(v - (t ? d.tab[t - 1].vec_end : 0));
        const unsigned vpr = (unsigned)T.D / VW;
        const unsigned rl = lv / vpr, col = (lv - rl * vpr) * VW;
        const i32 row = T.row_lo + (i32)rl;
        const i64 e = (i64)row * T.D + col;                 // element index inside the full table
        const i64 key = (i64)d.step_rel * d.rows_all + T.key_off + row;
        V xv = *reinterpret_cast<const V *>(own + T.x_off + e * 4);
        V mv, vv;
        if (d.adam) { mv = *reinterpret_cast<const V *>(T.m + e); vv = *reinterpret_cast<const V *>(T.v + e); }
        float g[VW];
#pragma unroll
        for (int q = 0; q < VW; q++) g[q] = 0.f;
        bool any = false;
        for (int p0 = 0; p0 < d.world; p0 += 4) {           // four ranks at a time: their loads are in flight together
            int4 seg[4];
            V g0[4], g1[4];
            const float *gb[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                seg[u] = make_int4(-1, 0, 0, 0);
                if (p0 + u < d.world) seg[u] = __ldg(reinterpret_cast<const int4 *>(d.arena[p0 + u] + d.off_rowhead) + key);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (seg[u].x < 0) continue;
                const int p = p0 + u;
                gb[u] = (const float *)(d.arena[p] + (T.is_ent ? d.off_gent : d.off_grel)) + T.part * T.D + col -
                        (T.is_ent ? (i64)0 : (i64)d.nes[p] * T.cols);
                g0[u] = __ldg(reinterpret_cast<const V *>(gb[u] + (i64)seg[u].z * T.cols));
                if (seg[u].y - seg[u].x > 1) g1[u] = __ldg(reinterpret_cast<const V *>(gb[u] + (i64)seg[u].w * T.cols));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (seg[u].x < 0) continue;
                const int p = p0 + u;
                float sp[VW];                              // this rank's segment sum, in slot order
                const float *q0 = reinterpret_cast<const float *>(&g0[u]), *q1 = reinterpret_cast<const float *>(&g1[u]);
                const i32 cnt = seg[u].y - seg[u].x;
#pragma unroll
                for (int q = 0; q < VW; q++) { sp[q] = 0.f + q0[q]; if (cnt > 1) sp[q] += q1[q]; }
                const i32 *perm = reinterpret_cast<const i32 *>(d.arena[p] + d.off_perm) + (i64)d.step_rel * d.nloc[p];
                for (i32 j = seg[u].x + 2; j < seg[u].y; j++) {
                    const V gj = __ldg(reinterpret_cast<const V *>(gb[u] + (i64)__ldg(perm + j) * T.cols));
                    const float *pj = reinterpret_cast<const float *>(&gj);
#pragma unroll
                    for (int q = 0; q < VW; q++) sp[q] += pj[q];
                }
#pragma unroll
                for (int q = 0; q < VW; q++) { g[q] += sp[q]; any |= sp[q] != 0.f; }
            }
        }
        float *xs = reinterpret_cast<float *>(&xv);
        if (d.adam) {
            float *ms = reinterpret_cast<float *>(&mv), *vs = reinterpret_cast<float *>(&vv);
#pragma unroll
            for (int q = 0; q < VW; q++) adam_elem(xs[q], ms[q], vs[q], g[q], b1, b2, 1.f - b1, 1.f - b2, lr, eps);
            *reinterpret_cast<V *>(T.m + e) = mv; *reinterpret_cast<V *>(T.v + e) = vv;
        } else {
            if (!any) continue;                            // SGD leaves rows without gradient untouched: nothing to publish
#pragma unroll
            for (int q = 0; q < VW; q++) xs[q] -= lr * g[q];
        }
        for (int p = 0; p < d.world; p++) *reinterpret_cast<V *>(d.arena[p] + T.x_off + e * 4) = xv;
    }
}

// ---- "scatter" form: the step is the single-GPU pair of kernels.  Every rank samples and plans the GLOBAL batch (integer
// work, identical everywhere), computes the positives of its own streams and its grad kernel stores each gradient row
// straight into the arena of the rank that owns the row's table row, at the row's slot of the global batch (grad_dst_ent /
// grad_dst_rel in train_dev.cuh) — the reduce-scatter is those peer stores, there is no reduce+push kernel and no staging
// slab.  The owner then runs the single-GPU update over ITS rows: same row map, same slots, same ascending-slot fp32 order,
// so the tables are bit-identical to one GPU training the global batch; the new row goes to every rank's table (peer
// stores: the all-gather).  Hinge terms are stored into every rank's loss buffer, so each rank reduces the global loss in
// the single-GPU order too.
struct DpSc {
    char *arena[OKB_DP_MAX];
    i64 delta[OKB_DP_MAX];     // arena[p] - arena[rank]: rank p's copy of a local table element
    i64 off_flags;
    i32 world, rank, adam;
    i32 ent_lo, ent_hi, rel_lo, rel_hi;                    // owned rows
    i32 nbcast;                                            // copies of a new row: world (scatter form) or 1 (gather form: this rank's own)
    unsigned long long epoch;
    // hub rows: `hub_blocks` CTAs in front of the tiles pre-reduce the PCH-blocks of this rank's long segments (what
    // prereduce_kernel does on one GPU) and count themselves off in *hub_ctr; a tile thread whose row has a long segment
    // waits until the counter has reached hub_target before it reads the block sums
    unsigned long long *hub_ctr, hub_target;
    i32 hub_blocks;
    unsigned long long *trace;
    i32 hs_mode;
};
__device__ __forceinline__ bool dp_owns_key(const UpdArgs &a, const DpSc &d, i32 key) {
    return key < a.E ? (key >= d.ent_lo && key < d.ent_hi) : (key - a.E >= d.rel_lo && key - a.E < d.rel_hi);
}
// One 256-vector tile of ONE table's owned rows per CTA (a.tab[] describes the owned ranges); the loss blocks come first.
// Row map, x / m / v are requested before the waits (nobody but this owner writes them).
template <int VW>
__global__ void __launch_bounds__(256, 4) dp_scatter_update_kernel(UpdArgs a, DpSc d) {
    typedef typename VecT<VW>::T V;
    pdl_launch_dependents();
    if (threadIdx.x == 0) trace_min(d.trace, 5);
    const i32 hid = (i32)blockIdx.x - a.loss_blocks;       // loss blocks, then hub blocks, then the tiles
    if (hid >= 0 && hid < d.hub_blocks) {
        // one warp per PCH-block of sorted positions; same additions in the same order as prereduce_body.  Which blocks lie
        // inside ONE segment of a row this rank owns is plan data: a CTA without such a block leaves before the waits
        const i32 w = hid * 8 + (i32)(threadIdx.x >> 5), lo = w * PCH, lane = threadIdx.x & 31;
        bool mine = false;
        i32 key = 0;
        if (lo + PCH <= a.n) {
            key = a.skeys[lo];
            mine = key < a.key_limit && dp_owns_key(a, d, key) && a.skeys[lo + PCH - 1] == key;
        }
        if (!__syncthreads_or(mine)) {
            if (threadIdx.x == 0) { atomicAdd(d.hub_ctr, 1ull); trace_max(d.trace, 8); }
            return;
        }
        pdl_wait();
        dp_announce_and_wait(d.arena, d.off_flags, d.world, d.rank, DP_FLAG_STAGE, d.epoch, false, d.hs_mode);
        {
            if (mine) {
                const bool is_ent = key < a.E;
                const i32 cols = is_ent ? a.ce : a.cr;
                const float *gb = is_ent ? a.gent : a.grel - (i64)a.n_ent_slots * a.cr;
                i32 slot[PCH];
#pragma unroll
                for (int j = 0; j < PCH; j++) slot[j] = __ldg(a.perm + lo + j);
                for (i32 v = lane * VW; v < cols; v += 32 * VW) {
                    V r[PCH];
#pragma unroll
                    for (int j = 0; j < PCH; j++) r[j] = __ldcg(reinterpret_cast<const V *>(gb + (i64)slot[j] * cols + v));
                    V acc;
                    float *pa = reinterpret_cast<float *>(&acc);
#pragma unroll
                    for (int q = 0; q < VW; q++) pa[q] = 0.f;
#pragma unroll
                    for (int j = 0; j < PCH; j++)
#pragma unroll
                        for (int q = 0; q < VW; q++) pa[q] += reinterpret_cast<const float *>(&r[j])[q];
                    *reinterpret_cast<V *>(const_cast<float *>(a.partial) + (i64)w * a.pcols + v) = acc;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); atomicAdd(d.hub_ctr, 1ull); trace_max(d.trace, 8); }
        return;
    }
    const i32 bid = hid < 0 ? -1 : hid - d.hub_blocks;
    int t = 0;
    if (bid >= 0) while (bid >= a.tab[t].blk_end) t++;     // block-uniform
    const DenseTab &T = a.tab[t];
    bool live = false;
    unsigned lv = 0, col = 0;
    int4 seg = make_int4(-1, 0, 0, 0);
    V xv, mv, vv;
    if (bid >= 0) {
        const unsigned nvec = (unsigned)(T.vec_end - (t ? a.tab[t - 1].vec_end : 0));
        lv = ((unsigned)bid - (unsigned)(t ? a.tab[t - 1].blk_end : 0)) * 256u + threadIdx.x;
        live = lv < nvec;
        if (live) {
            const unsigned vpr = (unsigned)T.D / VW;
            const unsigned row = T.magic ? (__umulhi(lv, T.magic) >> T.shift) : (lv >> T.shift);
            col = (lv - row * vpr) * VW;
            seg = __ldg(a.rowhead + T.key_off + row);
            const size_t e = (size_t)lv * VW;
            if (d.adam || seg.x >= 0) xv = *reinterpret_cast<const V *>(T.x + e);
            if (d.adam) { mv = *reinterpret_cast<const V *>(T.m + e); vv = *reinterpret_cast<const V *>(T.v + e); }
        }
    }
    pdl_wait();                                            // this rank's grad kernel is complete: announce it
    if (threadIdx.x == 0) trace_min(d.trace, 6);
    dp_announce_and_wait(d.arena, d.off_flags, d.world, d.rank, DP_FLAG_STAGE, d.epoch, blockIdx.x == 0, d.hs_mode);
    if (threadIdx.x == 0) trace_min(d.trace, 7);
    // the loss is summed by as many threads as the single-GPU kernel of this optimizer has (sgd_kernel: 128): same fp32 order
    if (bid < 0) { loss_block(a, (i32)blockIdx.x, d.adam ? (i32)blockDim.x : WARPS_PER_BLOCK * 32); return; }
    if (!live || (!d.adam && seg.x < 0)) return;           // SGD leaves rows without gradient untouched
    float g[VW];
#pragma unroll
    for (int q = 0; q < VW; q++) g[q] = 0.f;
    if (a.hub && seg.x >= 0 && seg.y - seg.x > PCH && (seg.x + PCH - 1) / PCH < seg.y / PCH) {     // long segment: block sums ready?
        unsigned long long seen;
        do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(d.hub_ctr) : "memory"); } while (seen < d.hub_target);
    }
    adam_gsum<VW, true>(a, T, seg, col, g);                // peers wrote these rows: coherent loads
    const size_t e = (size_t)lv * VW;
    float *xs = reinterpret_cast<float *>(&xv);
    if (d.adam) {
        const float b1 = a.hp.beta1, b2 = a.hp.beta2;
        float *ms = reinterpret_cast<float *>(&mv), *vs = reinterpret_cast<float *>(&vv);
#pragma unroll
        for (int q = 0; q < VW; q++) adam_elem(xs[q], ms[q], vs[q], g[q], b1, b2, 1.f - b1, 1.f - b2, a.hp.lr, a.hp.eps);
        *reinterpret_cast<V *>(T.m + e) = mv; *reinterpret_cast<V *>(T.v + e) = vv;
    } else {
#pragma unroll
        for (int q = 0; q < VW; q++) xs[q] -= a.hp.lr * g[q];
    }
    char *xp = reinterpret_cast<char *>(T.x + e);
    for (int p = 0; p < d.nbcast; p++) *reinterpret_cast<V *>(xp + d.delta[p]) = xv;
    if (threadIdx.x == 0) trace_max(d.trace, 9);
}

static void fill_upd_args(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, const float *gent, const float *grel,
                          const float *loss_terms, UpdArgs &a) {
    const i64 n = c->plan_ne + c->plan_nr, total = (c->plan_hi - c->plan_lo) * n, rel = (step - c->plan_lo) * n;
    a.m = *m; a.hp = *hp;
    a.skeys = c->keys_ent.as<i32>() + total + rel; a.perm = c->perm_ent.as<i32>() + rel;
    a.gent = gent; a.grel = grel; a.loss_terms = loss_terms; a.loss_out = nullptr;
    a.n = (i32)n; a.n_ent_slots = (i32)c->plan_ne; a.E = (i32)c->E; a.R = (i32)c->R; a.B = (i32)c->B;
    a.w = 1.0f / (float)(c->B * (c->K + c->KR));
    group_cols(m, a.ce, a.cr);
    a.ntab = 0;
    a.key_limit = (i32)(c->E + c->R);
    a.rowhead = c->rowseg_e.as<int4>() + (step - c->plan_lo) * (c->E + c->R);
    a.pcols = std::max(a.ce, a.cr);
    a.hub = (double)c->plan_ne * c->max_ent_share > PCH || (double)c->plan_nr * c->max_rel_share > PCH;
    a.partial = nullptr; a.by_row = 1; a.loss_blocks = 1;
    a.loss_part = nullptr; a.loss_ctr = nullptr; a.bad = nullptr;
}

// n steps of the scatter form (see DpSc): per step  grad (peer stores of the gradient rows)  ->  [hub pre-reduction]  ->
// owner update (peer stores of the new rows); two flag exchanges per step, riding on the kernel boundaries
static int dp_scatter_steps(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step_lo, INT n, float *loss_out, int vw, int nv,
                            cudaStream_t s) {
    const okb_dp &P = c->dp;
    int rc;
    if (P.off_gent < 0 || P.off_grel < 0 || P.off_loss < 0) OKB_FAIL(c, OKB_ERR_ARG, "arena has no receive buffers (okb_dp_layout with scatter = 1)");
    if (P.global_batch != c->B || P.neg_ent != c->K || P.neg_rel != c->KR) OKB_FAIL(c, OKB_ERR_ARG, "sampled batch does not match the arena's receive buffers");
    if (!planned(c, step_lo, step_lo + n, 0, c->B)) {      // the GLOBAL batch, like one GPU would plan it
        if ((rc = plan_steps(c, step_lo, c->steps, 0, c->B, s))) return rc;
    }
    char *own = (char *)P.arena[P.rank];
    const bool gather = P.scatter == 2;
    i32 ce0, cr0;
    group_cols(m, ce0, cr0);
    const i64 buf_bytes[3] = {((c->B * (2 + c->K) * ce0 * 4) + 255) & ~(i64)255, ((c->B * (1 + c->KR) * cr0 * 4) + 255) & ~(i64)255,
                              ((c->B * 4) + 255) & ~(i64)255};
    cudaLaunchAttribute pat[1];
    pat[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pat[0].val.programmaticStreamSerializationAllowed = 1;
    const i64 own_e = (c->E + P.world - 1) / P.world, own_r = (c->R + P.world - 1) / P.world;
    DpSc d;
    for (int q = 0; q < OKB_DP_MAX; q++) {
        d.arena[q] = q < P.world ? (char *)P.arena[q] : nullptr;
        d.delta[q] = q < P.world ? (i64)((char *)P.arena[q] - own) : 0;
    }
    d.off_flags = P.off_flags; d.world = P.world; d.rank = P.rank; d.adam = m->optimizer == OKB_ADAM;
    d.ent_lo = (i32)std::min<i64>(c->E, P.rank * own_e); d.ent_hi = (i32)std::min<i64>(c->E, (P.rank + 1) * own_e);
    d.rel_lo = (i32)std::min<i64>(c->R, P.rank * own_r); d.rel_hi = (i32)std::min<i64>(c->R, (P.rank + 1) * own_r);
    d.nbcast = P.world;
    if (gather) {                                          // every rank updates every row of its own copy
        d.ent_lo = 0; d.ent_hi = (i32)c->E; d.rel_lo = 0; d.rel_hi = (i32)c->R;
        d.nbcast = 1; d.delta[0] = 0;
    }
    const unsigned long long *xflags = (const unsigned long long *)(own + P.off_flags) + DP_FLAG_X;
    for (INT i = 0; i < n; i++) {
        const INT step = step_lo + i;
        d.epoch = c->dp_epoch + 1;
        // gather form: the receive buffers are used in turn — a peer's grad kernel of step n + 1 can only run after it has
        // seen this rank's "gradient rows landed" announcement of step n, i.e. after this rank's update of step n - 1
        const i64 par = gather ? (i64)(d.epoch & 1) : 0;
        for (int z = 0; z < 3; z++) c->sc_buf_off[z] = par * buf_bytes[z];
        float *gent = (float *)(own + P.off_gent + c->sc_buf_off[0]), *grel = (float *)(own + P.off_grel + c->sc_buf_off[1]),
              *lterms = (float *)(own + P.off_loss + c->sc_buf_off[2]);
        c->sc_launch = gather ? 2 : 1;
        rc = launch_grad(c, m, hp + i, step, P.b_lo, P.b_hi, 0, gent, grel, lterms, gather ? nullptr : xflags, c->dp_epoch, gather ? 0 : P.world, s);
        c->sc_launch = 0;
        if (rc) return rc;
        UpdArgs a;
        i32 blk = 0;
        bool lean = true;
        if ((rc = okb_fill_update(c, m, hp + i, step, gent, grel, lterms, loss_out ? loss_out + i : nullptr, vw, s, a, blk, lean))) return rc;
        a.ntab = 0;                                        // the table list: this rank's rows only
        i64 acc = 0;
        i32 oblk = 0;
        auto add = [&](float *x, float *mm, float *vv, bool is_ent, int part) {
            if (!x) return;
            const int D = is_ent ? m->ent_dim : m->rel_dim;
            const i64 lo = is_ent ? d.ent_lo : d.rel_lo, hi = is_ent ? d.ent_hi : d.rel_hi;
            acc += (hi - lo) * D / vw;
            oblk += (i32)(((hi - lo) * D / vw + 255) / 256);
            DenseTab T = {};
            T.x = x + lo * D; T.m = mm ? mm + lo * D : nullptr; T.v = vv ? vv + lo * D : nullptr;
            T.grad = is_ent ? gent : grel; T.vec_end = acc; T.D = D;
            T.key_off = (i32)((is_ent ? 0 : c->E) + lo); T.cols = is_ent ? a.ce : a.cr; T.part = part;
            T.slot_off = is_ent ? 0 : (i32)c->plan_ne; T.blk_end = oblk; T.sblk_end = oblk;
            const unsigned vpr = (unsigned)(D / vw);
            unsigned sh = 0;
            while ((2u << sh) <= vpr) sh++;
            if ((1u << sh) == vpr) { T.magic = 0; T.shift = sh; }
            else { T.magic = (unsigned)((((unsigned long long)1 << (32 + sh)) + vpr - 1) / vpr); T.shift = sh; }
            a.tab[a.ntab++] = T;
        };
        add(m->ent, m->m_ent, m->v_ent, true, 0);
        if (m->model == OKB_TRANSD) add(m->ent_aux, m->m_ent_aux, m->v_ent_aux, true, 1);
        add(m->rel, m->m_rel, m->v_rel, false, 0);
        if (m->model != OKB_TRANSE) add(m->rel_aux, m->m_rel_aux, m->v_rel_aux, false, 1);
        if (oblk == 0) a.tab[0].blk_end = 1;               // owns no rows: the loss blocks still take part in the flag exchange
        d.hub_blocks = a.hub ? (i32)((a.n / PCH + 7) / 8) : 0;
        d.hub_ctr = (unsigned long long *)(c->flags.as<unsigned>() + OKB_FLAGS_HUBCTR);
        c->hub_done += (unsigned long long)d.hub_blocks;
        d.hub_target = c->hub_done;
        d.hs_mode = c->dp_hs_mode;
        d.trace = c->dp_trace_on ? c->dp_trace.as<unsigned long long>() + (d.epoch % 64) * 16 : nullptr;
        {
            ProfScope po(c, PROF_DP_OWNER, s);
            cudaLaunchConfig_t oc = {};
            oc.gridDim = dim3((unsigned)(oblk + a.loss_blocks + d.hub_blocks)); oc.blockDim = dim3(256); oc.stream = s; oc.attrs = pat; oc.numAttrs = c->pdl ? 1 : 0;
            if (vw == 4) cudaLaunchKernelEx(&oc, dp_scatter_update_kernel<4>, a, d);
            else if (vw == 2) cudaLaunchKernelEx(&oc, dp_scatter_update_kernel<2>, a, d);
            else cudaLaunchKernelEx(&oc, dp_scatter_update_kernel<1>, a, d);
        }
        OKB_LAUNCHED(1);
        c->dp_epoch = d.epoch;
    }
    {
        DpPush e;
        for (int q = 0; q < OKB_DP_MAX; q++) e.arena[q] = q < P.world ? (char *)P.arena[q] : nullptr;
        e.off_flags = P.off_flags; e.world = P.world; e.rank = P.rank; e.epoch = c->dp_epoch;
        dp_announce_kernel<<<1, 32, 0, s>>>(e);
        OKB_LAUNCHED(1);
    }
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

extern "C" {

int okb_dp_attach(okb_ctx *c, const okb_dp *cfg) {
    if (!cfg || cfg->world < 1 || cfg->world > OKB_DP_MAX || cfg->rank < 0 || cfg->rank >= cfg->world)
        OKB_FAIL(c, OKB_ERR_ARG, "bad data-parallel geometry");
    if (cfg->b_lo < 0 || cfg->b_hi <= cfg->b_lo) OKB_FAIL(c, OKB_ERR_ARG, "bad positive range");
    for (int q = 0; q < cfg->world; q++) if (!cfg->arena[q]) OKB_FAIL(c, OKB_ERR_ARG, "peer arena missing");
    c->dp = *cfg;
    c->dp_on = true;
    c->dp_epoch = 0;
    if (cfg->off_rowhead >= 0 && cfg->plan_steps > 0) {    // pull form: the plan / gradient buffers are slices of the arena
        char *own = (char *)cfg->arena[cfg->rank];
        const i64 NE = 2 + cfg->neg_ent, NR = 1 + cfg->neg_rel, Bl = cfg->max_local;
        c->rowseg_e.adopt(own + cfg->off_rowhead, (size_t)(cfg->plan_steps * (c->E + c->R) * (i64)sizeof(int4)));
        c->perm_ent.adopt(own + cfg->off_perm, (size_t)(cfg->plan_steps * Bl * (NE + NR) * 4));
        c->gent.adopt(own + cfg->off_gent, (size_t)(cfg->off_grel - cfg->off_gent));
        c->grel.adopt(own + cfg->off_grel, (size_t)(cfg->off_loss - cfg->off_grel));
        c->lossterms.adopt(own + cfg->off_loss, (size_t)(Bl * 4));
        c->plan_lo = c->plan_hi = 0;
        c->rowhead_ready = false;
    }
    return 0;
}
int okb_dp_detach(okb_ctx *c) {
    c->dp_on = false;
    for (DevBuf *b : {&c->rowseg_e, &c->perm_ent, &c->gent, &c->grel, &c->lossterms}) b->disown();
    c->plan_lo = c->plan_hi = 0;
    c->rowhead_ready = false;
    return 0;
}

/* bytes a rank's arena needs, and the offsets of its parts (all 256-byte aligned) */
int okb_dp_layout(okb_ctx *c, const okb_model *m, INT world, okb_dp *out) {
    if (world < 1 || world > OKB_DP_MAX) OKB_FAIL(c, OKB_ERR_ARG, "world size not supported");
    if (m->model == OKB_TRANSR) OKB_FAIL(c, OKB_ERR_ARG, "TransR is not supported by the owner-sharded data-parallel path");
    i32 ce, cr;
    group_cols(m, ce, cr);
    i64 off = 0;
    auto take = [&](i64 bytes) { i64 o = off; off += (bytes + 255) & ~(i64)255; return o; };
    const i64 tb_e = c->E * m->ent_dim * 4, tb_r = c->R * m->rel_dim * 4;
    out->off_ent = take(tb_e);
    out->off_ent_aux = m->model == OKB_TRANSD ? take(tb_e) : -1;
    out->off_rel = take(tb_r);
    out->off_rel_aux = m->model != OKB_TRANSE ? take(tb_r) : -1;
    const i64 own_e = (c->E + world - 1) / world, own_r = (c->R + world - 1) / world;
    out->off_stage_ent = out->off_stage_rel = -1;
    if (!out->scatter) {
        out->off_stage_ent = take(world * own_e * ce * 4);
        out->off_stage_rel = take(world * own_r * cr * 4);
    }
    out->off_flags = take(DP_FLAG_BYTES);
    out->off_rowhead = out->off_perm = out->off_gent = out->off_grel = out->off_loss = out->off_partial = -1;
    if (out->scatter) {                                    // scatter form: receive buffers for the gradient rows of the GLOBAL batch
        if (out->global_batch < 1 || out->neg_ent < 0 || out->neg_rel < 0) OKB_FAIL(c, OKB_ERR_ARG, "scatter form needs global_batch, neg_ent, neg_rel");
        const i64 nb = out->scatter == 2 ? 2 : 1;          // gather form: two buffers, used in turn (no "buffer free" exchange)
        out->off_gent = take(nb * (((out->global_batch * (2 + out->neg_ent) * ce * 4) + 255) & ~(i64)255));
        out->off_grel = take(nb * (((out->global_batch * (1 + out->neg_rel) * cr * 4) + 255) & ~(i64)255));
        out->off_loss = take(nb * (((out->global_batch * 4) + 255) & ~(i64)255));
        out->plan_steps = 0;
    }
    if (out->plan_steps > 0 && out->max_local > 0) {       // pull form: plan, gradient rows and loss terms in the arena too
        const i64 NE = 2 + out->neg_ent, NR = 1 + out->neg_rel, Bl = out->max_local;
        out->off_rowhead = take(out->plan_steps * (c->E + c->R) * (i64)sizeof(int4));
        out->off_perm = take(out->plan_steps * Bl * (NE + NR) * 4);
        out->off_gent = take(Bl * NE * ce * 4);
        out->off_grel = take(Bl * NR * cr * 4);
        out->off_loss = take(Bl * 4);
    }
    out->arena_bytes = off;
    out->world = (int32_t)world;
    return 0;
}

int okb_dp_train_steps(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step_lo, INT n, float *loss_out, void *stream) {
    if (!c->dp_on) OKB_FAIL(c, OKB_ERR_STATE, "okb_dp_attach first");
    int vw, nv;
    int rc = check_model(c, m, vw, nv);
    if (rc) return rc;
    if (m->model == OKB_TRANSR) OKB_FAIL(c, OKB_ERR_ARG, "TransR is not supported by the owner-sharded data-parallel path");
    const okb_dp &P = c->dp;
    if (step_lo < 0 || n < 1 || step_lo + n > c->steps) OKB_FAIL(c, OKB_ERR_ARG, "step range out of range (sample first)");
    if (P.b_hi > c->B) OKB_FAIL(c, OKB_ERR_ARG, "rank's positive range exceeds the sampled batch");
    char *own = (char *)P.arena[P.rank];
    if ((char *)m->ent != own + P.off_ent || (char *)m->rel != own + P.off_rel ||
        (m->ent_aux && (char *)m->ent_aux != own + P.off_ent_aux) || (m->rel_aux && (char *)m->rel_aux != own + P.off_rel_aux))
        OKB_FAIL(c, OKB_ERR_ARG, "model tables must live in this rank's peer arena at the okb_dp_layout offsets");
    cudaStream_t s = (cudaStream_t)stream;
    if (P.scatter) return dp_scatter_steps(c, m, hp, step_lo, n, loss_out, vw, nv, s);
    const i64 Bl = P.b_hi - P.b_lo;
    INT er, ec, rr, rcn;
    if ((rc = okb_grad_sizes(c, m, Bl, c->K, c->KR, &er, &ec, &rr, &rcn))) return rc;
    if (c->gent.ensure(sizeof(float) * er * ec) || c->grel.ensure(sizeof(float) * rr * rcn) || c->lossterms.ensure(sizeof(float) * Bl))
        OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (gradient rows)");
    const bool pull_ok = c->dp_pull && P.off_rowhead >= 0 && c->rowseg_e.external;
    if (!planned(c, step_lo, step_lo + n, P.b_lo, P.b_hi)) {      // plan everything that is sampled from here on with one sort
        if (pull_ok && c->dp_epoch > 0) {
            // peers read this rank's plan in their update kernels: do not overwrite it before they have all finished the last
            // step (they announce that at the end of their own call)
            DpPush w;
            for (int q = 0; q < OKB_DP_MAX; q++) w.arena[q] = q < P.world ? (char *)P.arena[q] : nullptr;
            w.off_flags = P.off_flags; w.world = P.world; w.rank = P.rank; w.epoch = c->dp_epoch;
            dp_wait_kernel<<<1, 32, 0, s>>>(w);
            OKB_LAUNCHED(1);
        }
        if (c->steps - step_lo > P.plan_steps && pull_ok) OKB_FAIL(c, OKB_ERR_ARG, "more steps sampled than the arena's plan_steps");
        if ((rc = plan_steps(c, step_lo, c->steps, P.b_lo, P.b_hi, stream))) return rc;
    }
    if ((rc = ensure_rowhead(c, s))) return rc;
    cudaLaunchAttribute pat[1];
    pat[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pat[0].val.programmaticStreamSerializationAllowed = 1;
    const i64 own_e = (c->E + P.world - 1) / P.world, own_r = (c->R + P.world - 1) / P.world;
    for (INT i = 0; i < n; i++) {
        const INT step = step_lo + i;
        const unsigned long long epoch = c->dp_epoch + 1;
        const unsigned long long *xflags = (const unsigned long long *)(own + P.off_flags) + DP_FLAG_X;
        if ((rc = launch_grad(c, m, hp + i, step, P.b_lo, P.b_hi, P.b_lo, c->gent.as<float>(), c->grel.as<float>(), c->lossterms.as<float>(),
                              xflags, c->dp_epoch, P.world, stream))) return rc;
        UpdArgs a;
        fill_upd_args(c, m, hp + i, step, c->gent.as<float>(), c->grel.as<float>(), c->lossterms.as<float>(), a);
        // the choice must be the same on every rank: judge hub rows by the LARGEST rank's share of the batch
        const bool hub_any = (double)(P.max_local * (2 + c->K)) * c->max_ent_share > PCH || (double)(P.max_local * (1 + c->KR)) * c->max_rel_share > PCH;
        if (pull_ok && !hub_any) {                         // two-kernel step: the row owners pull the partial rows
            DpPull o;
            o.hp = hp[i];
            for (int q = 0; q < OKB_DP_MAX; q++) o.arena[q] = q < P.world ? (char *)P.arena[q] : nullptr;
            o.off_flags = P.off_flags; o.off_rowhead = P.off_rowhead; o.off_perm = P.off_perm; o.off_gent = P.off_gent;
            o.off_grel = P.off_grel; o.off_loss = P.off_loss;
            const i64 NE = 2 + c->K, NR = 1 + c->KR, per = c->B / c->W + (c->B % c->W ? 1 : 0), chunk = per * (c->W / P.world);
            for (int q = 0; q < P.world; q++) {            // the slice geometry of parallel.partition (Base.cpp:85-92)
                const i64 lo = std::min<i64>(c->B, q * chunk), hi = std::min<i64>(c->B, (q + 1) * chunk);
                o.bl[q] = (i32)(hi - lo); o.nloc[q] = (i32)((hi - lo) * (NE + NR)); o.nes[q] = (i32)((hi - lo) * NE);
            }
            if (o.bl[P.rank] != (i32)Bl) { OKB_FAIL(c, OKB_ERR_ARG, "rank's positive range does not follow the stream-slice geometry"); }
            o.world = P.world; o.rank = P.rank; o.adam = m->optimizer == OKB_ADAM; o.epoch = epoch;
            o.rows_all = (i32)(c->E + c->R); o.step_rel = (i32)(step - c->plan_lo);
            o.loss_out = loss_out ? loss_out + i : nullptr;
            o.w = 1.0f / (float)(c->B * (c->K + c->KR));
            o.ntab = 0;
            i64 acc = 0;
            auto addt = [&](i64 x_off, float *mm, float *vv, bool is_ent, int part) {
                if (x_off < 0) return;
                const int D = is_ent ? m->ent_dim : m->rel_dim;
                const i64 rows = is_ent ? c->E : c->R, own = (rows + P.world - 1) / P.world;
                const i64 lo = std::min<i64>(rows, P.rank * own), hi = std::min<i64>(rows, (P.rank + 1) * own);
                acc += (hi - lo) * D / vw;
                PullTab T;
                T.m = mm; T.v = vv; T.x_off = x_off; T.vec_end = acc; T.D = D; T.cols = is_ent ? a.ce : a.cr; T.part = part;
                T.row_lo = (i32)lo; T.key_off = is_ent ? 0 : (i32)c->E; T.is_ent = is_ent ? 1 : 0;
                o.tab[o.ntab++] = T;
            };
            addt(P.off_ent, m->m_ent, m->v_ent, true, 0);
            if (m->model == OKB_TRANSD) addt(P.off_ent_aux, m->m_ent_aux, m->v_ent_aux, true, 1);
            addt(P.off_rel, m->m_rel, m->v_rel, false, 0);
            if (m->model != OKB_TRANSE) addt(P.off_rel_aux, m->m_rel_aux, m->v_rel_aux, false, 1);
            o.work_blocks = (i32)std::max<i64>(1, std::min<i64>((acc + 255) / 256, (i64)okb_sms(c) * 8));
            cudaLaunchConfig_t oc = {};
            oc.gridDim = dim3((unsigned)o.work_blocks + 1); oc.blockDim = dim3(256); oc.stream = s; oc.attrs = pat; oc.numAttrs = c->pdl ? 1 : 0;
            {
                ProfScope po(c, PROF_DP_OWNER, s);
                if (vw == 4) cudaLaunchKernelEx(&oc, dp_pull_kernel<4>, o);
                else if (vw == 2) cudaLaunchKernelEx(&oc, dp_pull_kernel<2>, o);
                else cudaLaunchKernelEx(&oc, dp_pull_kernel<1>, o);
            }
            OKB_LAUNCHED(1);
            c->dp_epoch = epoch;
            continue;
        }
        if (a.hub) {
            const i64 nblocks = a.n / PCH;
            if (c->partial.ensure(sizeof(float) * (size_t)(nblocks + 1) * a.pcols)) { OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (partial sums)"); }
            a.partial = c->partial.as<float>();
            const unsigned pg = (unsigned)((nblocks + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK + 1);
            DISPATCH_LAYOUT(vw, nv, CALL_PRE);
            OKB_LAUNCHED(1);
        }
        DpPush d;
        for (int q = 0; q < OKB_DP_MAX; q++) d.arena[q] = q < P.world ? (char *)P.arena[q] : nullptr;
        d.off_stage_ent = P.off_stage_ent; d.off_stage_rel = P.off_stage_rel; d.off_flags = P.off_flags;
        for (int q = 0; q <= P.world; q++) { d.ent_lo[q] = (i32)std::min<i64>(c->E, q * own_e); d.rel_lo[q] = (i32)std::min<i64>(c->R, q * own_r); }
        d.own_max_ent = (i32)own_e; d.own_max_rel = (i32)own_r; d.world = P.world; d.rank = P.rank; d.Bl = (i32)Bl; d.epoch = epoch;
        d.trace = c->dp_trace_on ? c->dp_trace.as<unsigned long long>() + (epoch % 64) * 16 : nullptr; d.hs_mode = c->dp_hs_mode;
        a.work_blocks = (i32)((a.key_limit + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
        {
            ProfScope ps(c, PROF_UPDATE, s);
            cudaLaunchConfig_t pc = {};
            pc.gridDim = dim3((unsigned)(a.work_blocks + 1)); pc.blockDim = dim3(WARPS_PER_BLOCK * 32); pc.stream = s;
            pc.attrs = pat; pc.numAttrs = c->pdl ? 1 : 0;
#define CALL_PUSH(VW, NV) cudaLaunchKernelEx(&pc, dp_reduce_push_kernel<VW, NV>, a, d)
            { ProfScope pp(c, PROF_DP_PUSH, s); DISPATCH_LAYOUT(vw, nv, CALL_PUSH); }
            ProfScope po(c, PROF_DP_OWNER, s);
            DpOwn o;
            o.hp = hp[i];
            for (int q = 0; q < OKB_DP_MAX; q++) o.arena[q] = d.arena[q];
            o.off_flags = P.off_flags; o.world = P.world; o.rank = P.rank; o.adam = m->optimizer == OKB_ADAM; o.epoch = epoch;
            o.loss_out = loss_out ? loss_out + i : nullptr;
            o.w = 1.0f / (float)(c->B * (c->K + c->KR));
            o.trace = d.trace; o.hs_mode = c->dp_hs_mode;
            o.ntab = 0;
            i64 acc = 0;
            i32 oblk = 0;
            auto add = [&](i64 x_off, float *mm, float *vv, bool is_ent, int part) {
                if (x_off < 0) return;
                const int D = is_ent ? m->ent_dim : m->rel_dim;
                const i64 lo = is_ent ? d.ent_lo[P.rank] : d.rel_lo[P.rank], hi = is_ent ? d.ent_lo[P.rank + 1] : d.rel_lo[P.rank + 1];
                acc += (hi - lo) * D / vw;
                oblk += (i32)(((hi - lo) * D / vw + 255) / 256);
                DpTab T;
                T.m = mm; T.v = vv; T.x_off = x_off; T.stage_off = is_ent ? P.off_stage_ent : P.off_stage_rel; T.vec_end = acc;
                T.D = D; T.cols = is_ent ? a.ce : a.cr; T.part = part; T.row_lo = (i32)lo; T.own_max = (i32)(is_ent ? own_e : own_r);
                T.blk_end = oblk;
                const unsigned vpr = (unsigned)(D / vw);
                unsigned sh = 0;
                while ((2u << sh) <= vpr) sh++;
                if ((1u << sh) == vpr) { T.magic = 0; T.shift = sh; }
                else { T.magic = (unsigned)((((unsigned long long)1 << (32 + sh)) + vpr - 1) / vpr); T.shift = sh; }
                o.tab[o.ntab++] = T;
            };
            add(P.off_ent, m->m_ent, m->v_ent, true, 0);
            if (m->model == OKB_TRANSD) add(P.off_ent_aux, m->m_ent_aux, m->v_ent_aux, true, 1);
            add(P.off_rel, m->m_rel, m->v_rel, false, 0);
            if (m->model != OKB_TRANSE) add(P.off_rel_aux, m->m_rel_aux, m->v_rel_aux, false, 1);
            const unsigned og = (unsigned)std::max<i32>(1, oblk);
            if (oblk == 0) { o.tab[0].blk_end = 1; }          // a rank that owns no rows still takes part in the flag exchange
            cudaLaunchConfig_t oc = {};
            oc.gridDim = dim3(og); oc.blockDim = dim3(256); oc.stream = s; oc.attrs = pat; oc.numAttrs = c->pdl ? 1 : 0;
            if (vw == 4) cudaLaunchKernelEx(&oc, dp_owner_kernel<4>, o);
            else if (vw == 2) cudaLaunchKernelEx(&oc, dp_owner_kernel<2>, o);
            else cudaLaunchKernelEx(&oc, dp_owner_kernel<1>, o);
        }
        OKB_LAUNCHED(2);
        c->dp_epoch = epoch;
    }
    {
        DpPush d;
        for (int q = 0; q < OKB_DP_MAX; q++) d.arena[q] = q < P.world ? (char *)P.arena[q] : nullptr;
        d.off_flags = P.off_flags; d.world = P.world; d.rank = P.rank; d.epoch = c->dp_epoch;
        dp_announce_kernel<<<1, 32, 0, s>>>(d);
        OKB_LAUNCHED(1);
    }
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

/* Returns (stream-ordered) once every rank's row updates of all steps issued so far have landed in THIS rank's tables.
 * Not a collective: it only waits for flags the peers publish at the end of their own okb_dp_train_steps calls. */
int okb_dp_quiesce(okb_ctx *c, void *stream) {
    if (!c->dp_on) return 0;
    const okb_dp &P = c->dp;
    DpPush d;
    for (int q = 0; q < OKB_DP_MAX; q++) d.arena[q] = q < P.world ? (char *)P.arena[q] : nullptr;
    d.off_flags = P.off_flags; d.world = P.world; d.rank = P.rank; d.epoch = c->dp_epoch;
    dp_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d);
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

/* ---- chunk pipeline: sampling and planning depend only on the RNG streams, never on the parameters, so the NEXT chunk of
 * steps is sampled and planned on a side stream while the train kernels of the current chunk run.
 *   okb_chunk_begin    makes `steps` sampled + planned steps current: a matching prefetched chunk is swapped in (the caller's
 *                      stream waits for the side stream's event), otherwise they are produced on the caller's stream
 *   okb_chunk_prefetch starts producing the following chunk on the side stream
 * The RNG streams are saved before a prefetch; any call that observes or changes them (okb_sample, okb_get_streams,
 * randReset ...) first discards the prefetched chunk and restores them, so results never depend on prefetching. */
static void swap_slot(okb_ctx *c) {
    PlanSlot &a = c->alt;
    std::swap(c->batch, a.batch); std::swap(c->keys_ent, a.keys_ent); std::swap(c->perm_ent, a.perm_ent);
    std::swap(c->rowseg_e, a.rowseg_e); std::swap(c->sort_tmp, a.sort_tmp); std::swap(c->hist, a.hist);
    std::swap(c->B, a.B); std::swap(c->K, a.K); std::swap(c->KR, a.KR); std::swap(c->steps, a.steps);
    std::swap(c->plan_ne, a.plan_ne); std::swap(c->plan_nr, a.plan_nr); std::swap(c->plan_lo, a.plan_lo); std::swap(c->plan_hi, a.plan_hi);
    std::swap(c->plan_b_lo, a.plan_b_lo); std::swap(c->plan_b_hi, a.plan_b_hi); std::swap(c->rowhead_ready, a.rowhead_ready);
}
}  // extern "C"
int okb_discard_prefetch(okb_ctx *c) {
    if (!c->alt_ready) return 0;
    c->alt_ready = false;
    // the side stream sampled (and advanced the streams) after everything the caller's stream had issued before the
    // prefetch; restoring on the same side stream orders the copy behind its advance kernel without touching the
    // legacy default stream (which does not synchronise with non-blocking streams)
    OKB_CUDA(c, cudaMemcpyAsync(c->d_state, c->d_state_saved, sizeof(u64) * c->state.size(), cudaMemcpyDeviceToDevice, c->side));
    OKB_CUDA(c, cudaStreamSynchronize(c->side));
    c->state_dirty = true;
    return 0;
}
extern "C" {
int okb_chunk_begin(okb_ctx *c, INT B, INT k, INT kr, INT steps, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (c->alt_ready && c->alt.B == B && c->alt.K == k && c->alt.KR == kr && c->alt.steps == steps && !c->dp_on) {
        OKB_CUDA(c, cudaStreamWaitEvent(s, c->ev_side, 0));
        swap_slot(c);
        c->alt_ready = false;
        return 0;
    }
    int rc = okb_discard_prefetch(c);
    if (rc) return rc;
    if ((rc = okb_sample(c, B, k, kr, steps, 0, c->W, stream))) return rc;
    if ((rc = okb_plan_steps(c, 0, steps, stream))) return rc;
    return ensure_rowhead(c, s);
}
int okb_chunk_prefetch(okb_ctx *c, INT B, INT k, INT kr, INT steps, void *stream) {
    if (c->dp_on || steps < 1) return 0;
    int rc = okb_discard_prefetch(c);
    if (rc) return rc;
    if (!c->side) {
        OKB_CUDA(c, cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
        OKB_CUDA(c, cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
        OKB_CUDA(c, cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
    }
    if (!c->d_state_saved || c->saved_n != c->state.size()) {
        if (c->d_state_saved) cudaFree(c->d_state_saved);
        c->saved_n = c->state.size();
        OKB_CUDA(c, cudaMalloc((void **)&c->d_state_saved, sizeof(u64) * std::max<size_t>(c->saved_n, 1)));
    }
    // the side stream starts after everything issued so far: the slot it overwrites was read by the previous chunk's steps
    OKB_CUDA(c, cudaEventRecord(c->ev_main, (cudaStream_t)stream));
    OKB_CUDA(c, cudaStreamWaitEvent(c->side, c->ev_main, 0));
    OKB_CUDA(c, cudaMemcpyAsync(c->d_state_saved, c->d_state, sizeof(u64) * c->state.size(), cudaMemcpyDeviceToDevice, c->side));
    swap_slot(c);
    c->in_prefetch = true;
    rc = okb_sample(c, B, k, kr, steps, 0, c->W, c->side);
    if (!rc) rc = okb_plan_steps(c, 0, steps, c->side);
    if (!rc) rc = ensure_rowhead(c, c->side);
    c->in_prefetch = false;
    swap_slot(c);
    if (rc) return rc;
    OKB_CUDA(c, cudaEventRecord(c->ev_side, c->side));
    c->alt_ready = true;
    return 0;
}

/* The same pipeline under owner-sharded data parallelism: the chunk is sampled from the sampler streams [stream_lo, stream_hi)
 * (this rank's — or all of them in the scatter form) and planned over the positives the rank plans (its own; the global batch
 * in the scatter form).  Not for the pull form, whose plan lives in the peer arena. */
static void dp_plan_range(const okb_ctx *c, INT B, INT &lo, INT &hi) {
    if (c->dp.scatter) { lo = 0; hi = B; } else { lo = c->dp.b_lo; hi = c->dp.b_hi; }
}
int okb_dp_chunk_begin(okb_ctx *c, INT B, INT k, INT kr, INT steps, INT stream_lo, INT stream_hi, void *stream) {
    if (!c->dp_on) OKB_FAIL(c, OKB_ERR_STATE, "okb_dp_attach first");
    cudaStream_t s = (cudaStream_t)stream;
    INT lo, hi;
    dp_plan_range(c, B, lo, hi);
    if (c->alt_ready && c->alt.B == B && c->alt.K == k && c->alt.KR == kr && c->alt.steps == steps && c->alt.plan_b_lo == lo &&
        c->alt.plan_b_hi == hi && !c->rowseg_e.external) {
        OKB_CUDA(c, cudaStreamWaitEvent(s, c->ev_side, 0));
        swap_slot(c);
        c->alt_ready = false;
        return 0;
    }
    int rc = okb_discard_prefetch(c);
    if (rc) return rc;
    if ((rc = okb_sample(c, B, k, kr, steps, stream_lo, stream_hi, stream))) return rc;
    if (c->rowseg_e.external) return 0;                    // pull form: okb_dp_train_steps plans into the arena
    if ((rc = plan_steps(c, 0, steps, lo, hi, stream))) return rc;
    return ensure_rowhead(c, s);
}
int okb_dp_chunk_prefetch(okb_ctx *c, INT B, INT k, INT kr, INT steps, INT stream_lo, INT stream_hi, void *stream) {
    if (!c->dp_on || steps < 1 || c->rowseg_e.external) return 0;
    int rc = okb_discard_prefetch(c);
    if (rc) return rc;
    if (!c->side) {
        OKB_CUDA(c, cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
        OKB_CUDA(c, cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
        OKB_CUDA(c, cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
    }
    if (!c->d_state_saved || c->saved_n != c->state.size()) {
        if (c->d_state_saved) cudaFree(c->d_state_saved);
        c->saved_n = c->state.size();
        OKB_CUDA(c, cudaMalloc((void **)&c->d_state_saved, sizeof(u64) * std::max<size_t>(c->saved_n, 1)));
    }
    INT lo, hi;
    dp_plan_range(c, B, lo, hi);
    OKB_CUDA(c, cudaEventRecord(c->ev_main, (cudaStream_t)stream));
    OKB_CUDA(c, cudaStreamWaitEvent(c->side, c->ev_main, 0));
    OKB_CUDA(c, cudaMemcpyAsync(c->d_state_saved, c->d_state, sizeof(u64) * c->state.size(), cudaMemcpyDeviceToDevice, c->side));
    swap_slot(c);
    c->in_prefetch = true;
    rc = okb_sample(c, B, k, kr, steps, stream_lo, stream_hi, c->side);
    if (!rc) rc = plan_steps(c, 0, steps, lo, hi, c->side);
    if (!rc) rc = ensure_rowhead(c, c->side);
    c->in_prefetch = false;
    swap_slot(c);
    if (rc) return rc;
    OKB_CUDA(c, cudaEventRecord(c->ev_side, c->side));
    c->alt_ready = true;
    return 0;
}

/* ---- peer memory: allocations other processes of the box map over NVLink (CUDA IPC) */
int okb_peer_alloc(okb_ctx *c, INT bytes, void **ptr, unsigned char *handle64) {
    if (bytes <= 0 || !ptr || !handle64) OKB_FAIL(c, OKB_ERR_ARG, "bad peer allocation request");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void *p = nullptr;
    OKB_CUDA(c, cudaMalloc(&p, (size_t)bytes));
    OKB_CUDA(c, cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    OKB_CUDA(c, cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *ptr = p;
    return 0;
}
int okb_peer_open(okb_ctx *c, const unsigned char *handle64, void **ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    OKB_CUDA(c, cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int okb_peer_close(okb_ctx *c, void *ptr) { OKB_CUDA(c, cudaIpcCloseMemHandle(ptr)); return 0; }
int okb_peer_free(okb_ctx *c, void *ptr) { OKB_CUDA(c, cudaFree(ptr)); return 0; }

}  // extern "C"
