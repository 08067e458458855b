// TransR train step (TransR.py:36-75): entities are mapped into relation space by the positive's
// matrix, e' = e . M_r with M_r = transfer_matrix[r] viewed as [ent_size, rel_size]; with
// negative_rel == 0 the negatives use the POSITIVE's matrix (TransR.py:57-60).  (negative_rel > 0, TransR.py:61-65:
// transr_general_kernel further down.)
//
// The reference gathers one 40 KB matrix per batch row (B x De x Dr floats materialised per step) and
// runs a batched [1,De]x[De,Dr] matmul.  Here the batch is bucketed by relation — the plan's sorted
// relation-key segments already list each relation's positives — and ONE CTA owns one relation: M_r is
// staged in shared memory once, the CTA walks its positives in chunks, and per chunk runs the three small
// dense contractions on register-tiled fp32 FMAs out of shared memory:
//     P  = A . M_r        (projection, forward)
//     dA = G . M_r^T      (gradient of the gathered entity rows)
//     dM += A^T . G       (gradient of the matrix, accumulated in registers across chunks)
// so the matrix gradient is produced already reduced per relation (one [De,Dr] row per relation per
// step instead of one per positive) in a fixed order: deterministic, no atomics.
// Data parallelism shards RELATIONS (okb_transr_set_shard): a rank computes and updates only the relations
// [r_lo, r_hi) it owns — M_r, the 40 KB-per-relation operand, never crosses NVLink during training; only the
// entity gradient rows are summed across ranks (parallel.RelationSharded).
// The roof that binds at FB15K batch sizes is M_r traffic, not flops.
#include <algorithm>

#include "okb_internal.h"

#define FULL 0xffffffffu
#define TR_THREADS 256
#define TR_ROWS 48            // gathered entity rows per chunk (positives per chunk = TR_ROWS / (2 + k))
#define TR_MAXT 4             // dM tiles (4x4) per thread: De*Dr <= 16 * 256 * 4
#define EPS_NORM 1e-12f

struct TrArgs {
    okb_model m;
    const i32 *bh, *bt, *br;  // plane-major batch
    const i32 *skeys, *perm;  // this step's plan
    const int4 *rowhead;      // [E + R]: {first, end, slot0, slot1}
    float *gent, *grel, *loss_terms;
    float margin, w;
    i32 B, k, NE, E, R, n, nes, b_lo, b_hi, CH;
    i32 r_lo;                 // first relation of this context's shard (relation-sharded data parallelism; else 0)
};

__device__ __forceinline__ float wsum_t(float x) {
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__global__ void __launch_bounds__(TR_THREADS) transr_bucket_kernel(TrArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int De = a.m.ent_dim, Dr = a.m.rel_dim, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int De4 = De >> 2, Dr4 = Dr >> 2, per = 2 + a.k;
    const i32 r = a.r_lo + (i32)blockIdx.x;
    const int4 seg = a.rowhead[a.E + r];
    if (seg.x < 0) return;
    float *Msh = sm;                                   // [De][Dr]
    float *A = Msh + De * Dr;                          // [TR_ROWS][De]
    float *P = A + TR_ROWS * De;                       // [TR_ROWS][Dr]  projected, then normalised
    float *G = P + TR_ROWS * Dr;                       // [TR_ROWS][Dr]  gradient w.r.t. the projected rows
    float *Rg = G + TR_ROWS * Dr;                      // [CH][Dr]       per-positive gradient w.r.t. r_hat
    float *rhat = Rg + a.CH * Dr;                      // [Dr]
    float *rsum = rhat + Dr;                           // [Dr]
    float *inv = rsum + Dr;                            // [TR_ROWS]
    i32 *proj = (i32 *)(inv + TR_ROWS);                // [TR_ROWS]
    i32 *rowid = proj + TR_ROWS;                       // [TR_ROWS] entity id of each gathered row
    i32 *posb = rowid + TR_ROWS;                       // [CH] batch index of each positive in the chunk
    i32 *side = posb + a.CH;                           // [CH][k]  0: head replaced, 1: tail replaced, 2: same triple
    __shared__ float s_invr;
    __shared__ int s_projr;

    const float4 *Mg = reinterpret_cast<const float4 *>(a.m.rel_aux + (i64)r * De * Dr);
    for (int i = tid; i < De * Dr4; i += TR_THREADS) reinterpret_cast<float4 *>(Msh)[i] = __ldg(Mg + i);
    if (warp == 0) {                                   // r_hat = l2n(rel_embeddings[r])
        float ss = 0.f;
        for (int k = lane; k < Dr; k += 32) { const float v = a.m.rel[(i64)r * Dr + k]; ss += v * v; }
        ss = wsum_t(ss);
        const float iv = rsqrtf(fmaxf(ss, EPS_NORM));
        for (int k = lane; k < Dr; k += 32) rhat[k] = a.m.rel[(i64)r * Dr + k] * iv;
        if (lane == 0) { s_invr = iv; s_projr = ss > EPS_NORM; }
    }
    float accM[TR_MAXT][16];
#pragma unroll
    for (int q = 0; q < TR_MAXT; q++)
#pragma unroll
        for (int e = 0; e < 16; e++) accM[q][e] = 0.f;
    float racc = 0.f;                                   // thread k < Dr: sum over positives of d loss / d r_hat[k]
    const int ntM = De4 * Dr4;

    for (i32 base = seg.x; base < seg.y; base += a.CH) {
        const int np = min(a.CH, seg.y - base);         // positives in this chunk
        const int rows = np * per, rows4 = (rows + 3) & ~3;
        __syncthreads();
        if (tid < np) {
            const i32 b = a.perm[base + tid] - a.nes;   // relation slots are numbered nes + b  (NR == 1)
            posb[tid] = b;
            const i32 ph = a.bh[b], pt = a.bt[b];
            rowid[tid * per] = ph; rowid[tid * per + 1] = pt;
            for (int m = 0; m < a.k; m++) {
                const i32 at = b + (m + 1) * a.B;
                const i32 nh = a.bh[at], nt = a.bt[at];
                const int sd = nh != ph ? 0 : (nt != pt ? 1 : 2);
                side[tid * a.k + m] = sd;
                rowid[tid * per + 2 + m] = sd == 0 ? nh : (sd == 1 ? nt : ph);
            }
        }
        __syncthreads();
        for (int i = tid; i < rows4 * De4; i += TR_THREADS) {            // gather the entity rows (128-bit)
            const int row = i / De4, q = i - row * De4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows) v = __ldg(reinterpret_cast<const float4 *>(a.m.ent + (i64)rowid[row] * De) + q);
            reinterpret_cast<float4 *>(A)[row * De4 + q] = v;
        }
        __syncthreads();
        // ---- P = A . M   (4x4 register tiles)
        for (int t = tid; t < (rows4 >> 2) * Dr4; t += TR_THREADS) {
            const int tr = t / Dr4, tc = t - tr * Dr4;
            float acc[4][4] = {};
            for (int i = 0; i < De; i += 4) {
                float4 av[4], mv[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    av[q] = reinterpret_cast<const float4 *>(A)[(tr * 4 + q) * De4 + (i >> 2)];
                    mv[q] = reinterpret_cast<const float4 *>(Msh)[(i + q) * Dr4 + tc];
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    acc[q][0] += av[q].x * mv[0].x + av[q].y * mv[1].x + av[q].z * mv[2].x + av[q].w * mv[3].x;
                    acc[q][1] += av[q].x * mv[0].y + av[q].y * mv[1].y + av[q].z * mv[2].y + av[q].w * mv[3].y;
                    acc[q][2] += av[q].x * mv[0].z + av[q].y * mv[1].z + av[q].z * mv[2].z + av[q].w * mv[3].z;
                    acc[q][3] += av[q].x * mv[0].w + av[q].y * mv[1].w + av[q].z * mv[2].w + av[q].w * mv[3].w;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                reinterpret_cast<float4 *>(P)[(tr * 4 + q) * Dr4 + tc] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        }
        __syncthreads();
        // ---- normalise the projected rows (tf.nn.l2_normalize, TransR.py:19-23)
        for (int row = warp; row < rows; row += TR_THREADS / 32) {
            float ss = 0.f;
            for (int k = lane; k < Dr; k += 32) { const float v = P[row * Dr + k]; ss += v * v; }
            ss = wsum_t(ss);
            const float iv = rsqrtf(fmaxf(ss, EPS_NORM));
            for (int k = lane; k < Dr; k += 32) P[row * Dr + k] *= iv;
            if (lane == 0) { inv[row] = iv; proj[row] = ss > EPS_NORM; }
        }
        __syncthreads();
        // ---- scores, hinge and the gradient w.r.t. the projected rows: one warp per positive
        for (int c = warp; c < np; c += TR_THREADS / 32) {
            const int r0 = c * per;
            const float *Ph = P + r0 * Dr, *Pt = P + (r0 + 1) * Dr;
            float gh[4] = {0.f, 0.f, 0.f, 0.f}, gt[4] = {0.f, 0.f, 0.f, 0.f}, gr[4] = {0.f, 0.f, 0.f, 0.f}, gp[4];
            float sp = 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = lane + 32 * i;
                const float u = k < Dr ? (Ph[k] + rhat[k]) - Pt[k] : 0.f;
                sp += fabsf(u);
                gp[i] = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
            }
            sp = wsum_t(sp);
            // adds coef * (d score / d projected row) for the three roles of a triple with sign vector g
            auto backward = [&](const float *Hrow, const float *Trow, int hrow_i, int trow_i, const float *g, float coef,
                                float *dh, float *dt) {
                float d1 = 0.f, d2 = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = lane + 32 * i;
                    if (k < Dr) { d1 += g[i] * Hrow[k]; d2 += g[i] * Trow[k]; }
                }
                d1 = wsum_t(d1); d2 = wsum_t(d2);
                if (!proj[hrow_i]) d1 = 0.f;
                if (!proj[trow_i]) d2 = 0.f;
                const float ih = inv[hrow_i] * coef, it = inv[trow_i] * coef;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = lane + 32 * i;
                    if (k < Dr) {
                        dh[i] += ih * (g[i] - Hrow[k] * d1);
                        dt[i] -= it * (g[i] - Trow[k] * d2);
                        gr[i] += coef * g[i];
                    }
                }
            };
            float hinge = 0.f;
            int active = 0;
            for (int m = 0; m < a.k; m++) {
                const int sd = side[c * a.k + m], nrow = r0 + 2 + m;
                const float *Pn = P + nrow * Dr;
                const float *Hrow = sd == 0 ? Pn : Ph, *Trow = sd == 1 ? Pn : Pt;
                float gn[4], gnew[4] = {0.f, 0.f, 0.f, 0.f};
                float sn = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = lane + 32 * i;
                    const float u = k < Dr ? (Hrow[k] + rhat[k]) - Trow[k] : 0.f;
                    sn += fabsf(u);
                    gn[i] = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
                }
                sn = wsum_t(sn);
                const float x = sp - sn + a.margin;
                if (x >= 0.f) {
                    hinge += x; active++;
                    if (sd == 0) backward(Hrow, Trow, nrow, r0 + 1, gn, -a.w, gnew, gt);
                    else if (sd == 1) backward(Hrow, Trow, r0, nrow, gn, -a.w, gh, gnew);
                    else backward(Hrow, Trow, r0, r0 + 1, gn, -a.w, gh, gt);
                }
#pragma unroll
                for (int i = 0; i < 4; i++) { const int k = lane + 32 * i; if (k < Dr) G[nrow * Dr + k] = gnew[i]; }
            }
            if (active) backward(Ph, Pt, r0, r0 + 1, gp, a.w * (float)active, gh, gt);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = lane + 32 * i;
                if (k < Dr) { G[r0 * Dr + k] = gh[i]; G[(r0 + 1) * Dr + k] = gt[i]; Rg[c * Dr + k] = gr[i]; }
            }
            if (lane == 0) a.loss_terms[posb[c]] = hinge;
        }
        for (int i = tid + rows * Dr; i < rows4 * Dr; i += TR_THREADS) G[i] = 0.f;      // padding rows
        __syncthreads();
        // ---- dA = G . M^T  -> entity gradient rows of this chunk
        for (int t = tid; t < (rows4 >> 2) * De4; t += TR_THREADS) {
            const int tr = t / De4, ti = t - tr * De4;       // this thread: rows 4tr..4tr+3, columns ti + q*De4
            float acc[4][4] = {};
            for (int k = 0; k < Dr4; k++) {
                float4 gv[4], mv[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    gv[q] = reinterpret_cast<const float4 *>(G)[(tr * 4 + q) * Dr4 + k];
                    mv[q] = reinterpret_cast<const float4 *>(Msh)[(ti + q * De4) * Dr4 + k];
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[q][e] += dot4(gv[q], mv[e]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int row = tr * 4 + q;
                if (row < rows) {
                    const int c = row / per, j = row - c * per;
                    float *dst = a.gent + ((i64)posb[c] * a.NE + j) * De;
#pragma unroll
                    for (int e = 0; e < 4; e++) dst[ti + e * De4] = acc[q][e];
                }
            }
        }
        // ---- dM += A^T . G
#pragma unroll
        for (int q = 0; q < TR_MAXT; q++) {
            const int t = tid + q * TR_THREADS;
            if (t < ntM) {
                const int ti = t / Dr4, tk = t - ti * Dr4;
                for (int row = 0; row < rows; row++) {
                    const float4 av = reinterpret_cast<const float4 *>(A)[row * De4 + ti];
                    const float4 gv = reinterpret_cast<const float4 *>(G)[row * Dr4 + tk];
                    const float as[4] = {av.x, av.y, av.z, av.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                    for (int e = 0; e < 4; e++)
#pragma unroll
                        for (int f = 0; f < 4; f++) accM[q][e * 4 + f] += as[e] * gs[f];
                }
            }
        }
        if (tid < Dr) for (int c = 0; c < np; c++) racc += Rg[c * Dr + tid];       // fixed order over positives
    }
    // ---- gradient rows of this relation: [d rel_embeddings (Dr) | d M_r (De*Dr)]
    float *out = a.grel + (i64)r * (Dr + De * Dr);
    if (tid < Dr) rsum[tid] = racc;
    __syncthreads();
    if (tid < Dr) {
        float d = 0.f;
        for (int k = 0; k < Dr; k++) d += rsum[k] * rhat[k];
        out[tid] = s_invr * (rsum[tid] - (s_projr ? rhat[tid] * d : 0.f));
    }
#pragma unroll
    for (int q = 0; q < TR_MAXT; q++) {
        const int t = tid + q * TR_THREADS;
        if (t < ntM) {
            const int ti = t / Dr4, tk = t - ti * Dr4;
#pragma unroll
            for (int e = 0; e < 4; e++)
                reinterpret_cast<float4 *>(out + Dr + (ti * 4 + e) * Dr)[tk] =
                    make_float4(accM[q][e * 4 + 0], accM[q][e * 4 + 1], accM[q][e * 4 + 2], accM[q][e * 4 + 3]);
        }
    }
}

// ------------------------------------------------------------------------------------------ persistent, fused form
// An experiment, kept behind OKB_FLAG_TRANSR_FUSED (default off).  Hypothesis: the per-relation kernel above (119 us at the
// FB15K shape, B = 4,831) is bound by staging 40 KB of M_r per CTA, writing 40 KB of dM_r back and a second kernel that
// re-reads both (260 MB of HBM traffic for 104 MB of algorithmic bytes).  So: ONE CTA per SM stays resident and walks its
// share of the relations,
//   * M_r of the NEXT touched relation is in flight (one 40 KB bulk-async / TMA copy on an mbarrier into the other half of a
//     double buffer) while the current relation is computed;
//   * 512 threads per relation (the matrix-gradient tiles stay in registers: 2 x 16 per thread);
//   * the relation's update is applied on the spot — M_r - lr dM_r (or TF1 Adam) straight from the staged copy and the
//     register tiles, rel_embeddings[r] likewise — no dM round trip, no second kernel.  (Only this CTA ever touches
//     relation r during the step: negatives share the positive's relation, TransR.py:57-60.)
// MEASURED (profiles/README.md, r02h): 161.2 vs 163.4 us per step with SGD, 240 vs 195 with Adam.  The traffic was never
// the limiter: a relation owns ~11 gathered rows, and its pass is a chain of seven small dependent phases (gather, P,
// normalise, scores, dA, dM, update) of 1-3 us each in which most of the 512 threads idle — ~17 us per relation whether or
// not M_r is already in shared memory.  TF1 Adam's dense decay of untouched relations is streamed by the owning CTA.
#define TRF_THREADS 512
#define TRF_MAXT 2             // dM tiles (4x4) per thread: De*Dr <= 16 * 512 * 2

struct TrFuse {
    okb_hyper hp;
    const unsigned *bad;      // "bad id" flag of the host-batch path: set -> leave everything alone
    i32 adam, r_lo, r_hi;
};

__device__ __forceinline__ unsigned trf_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(TRF_THREADS, 1) transr_fused_kernel(TrArgs a, TrFuse f) {
    extern __shared__ __align__(128) float sm[];
    const int De = a.m.ent_dim, Dr = a.m.rel_dim, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int De4 = De >> 2, Dr4 = Dr >> 2, per = 2 + a.k, MS = De * Dr;
    if (f.bad && *(const volatile unsigned *)f.bad) return;
    float *Mb0 = sm, *Mb1 = sm + MS;                   // double-buffered M_r
    float *A = Mb1 + MS;                               // [TR_ROWS][De]
    float *P = A + TR_ROWS * De;                       // [TR_ROWS][Dr]  projected, then normalised
    float *G = P + TR_ROWS * Dr;                       // [TR_ROWS][Dr]  gradient w.r.t. the projected rows
    float *Rg = G + TR_ROWS * Dr;                      // [CH][Dr]       per-positive gradient w.r.t. r_hat
    float *rhat = Rg + a.CH * Dr;                      // [Dr]
    float *rsum = rhat + Dr;                           // [Dr]
    float *inv = rsum + Dr;                            // [TR_ROWS]
    i32 *proj = (i32 *)(inv + TR_ROWS);                // [TR_ROWS]
    i32 *rowid = proj + TR_ROWS;                       // [TR_ROWS] entity id of each gathered row
    i32 *posb = rowid + TR_ROWS;                       // [CH] batch index of each positive in the chunk
    i32 *side = posb + a.CH;                           // [CH][k]  0: head replaced, 1: tail replaced, 2: same triple
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ float s_invr;
    __shared__ int s_projr;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(trf_smem(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(trf_smem(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const i32 stride = (i32)gridDim.x;
    auto next_touched = [&](i32 r) { while (r < f.r_hi && __ldg(&a.rowhead[a.E + r].x) < 0) r += stride; return r; };
    auto issue = [&](i32 r, int b) {                   // thread 0: 40 KB of M_r into half b of the double buffer
        const unsigned bytes = (unsigned)MS * 4u, br = trf_smem(bar + b);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(br), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(trf_smem(b ? Mb1 : Mb0)),
                     "l"(a.m.rel_aux + (i64)r * MS), "r"(bytes), "r"(br) : "memory");
    };
    const float b1 = f.hp.beta1, b2 = f.hp.beta2, lr = f.hp.lr, eps = f.hp.eps, c1 = 1.f - b1, c2 = 1.f - b2;
    const int ntM = De4 * Dr4;
    i32 cur = next_touched(f.r_lo + (i32)blockIdx.x);
    int buf = 0;
    unsigned phase[2] = {0u, 0u};
    if (cur < f.r_hi && tid == 0) issue(cur, 0);
    while (cur < f.r_hi) {
        const i32 r = cur, nxt = next_touched(cur + stride);
        if (tid == 0) {
            if (nxt < f.r_hi) issue(nxt, buf ^ 1);     // the other half was released by the barrier that ended the previous pass
            if (f.adam) {                              // this relation's Adam slots are needed at the end of the pass: start them towards L2
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.m.m_rel_aux + (i64)r * MS), "r"((unsigned)MS * 4u) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.m.v_rel_aux + (i64)r * MS), "r"((unsigned)MS * 4u) : "memory");
            }
        }
        const int4 seg = a.rowhead[a.E + r];
        float *Msh = buf ? Mb1 : Mb0;
        if (warp == 0) {                                   // r_hat = l2n(rel_embeddings[r])
            float ss = 0.f;
            for (int k = lane; k < Dr; k += 32) { const float v = a.m.rel[(i64)r * Dr + k]; ss += v * v; }
            ss = wsum_t(ss);
            const float iv = rsqrtf(fmaxf(ss, EPS_NORM));
            for (int k = lane; k < Dr; k += 32) rhat[k] = a.m.rel[(i64)r * Dr + k] * iv;
            if (lane == 0) { s_invr = iv; s_projr = ss > EPS_NORM; }
        }
        float accM[TRF_MAXT][16];
#pragma unroll
        for (int q = 0; q < TRF_MAXT; q++)
#pragma unroll
            for (int e = 0; e < 16; e++) accM[q][e] = 0.f;
        float racc = 0.f;                                  // thread k < Dr: sum over positives of d loss / d r_hat[k]
        {   // M_r has landed?
            unsigned done = 0;
            const unsigned br = trf_smem(bar + buf);
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(br), "r"(phase[buf]) : "memory");
            phase[buf] ^= 1u;
        }
        for (i32 base = seg.x; base < seg.y; base += a.CH) {
            const int np = min(a.CH, seg.y - base);         // positives in this chunk
            const int rows = np * per, rows4 = (rows + 3) & ~3;
            __syncthreads();
            if (tid < np) {
                const i32 b = a.perm[base + tid] - a.nes;   // relation slots are numbered nes + b  (NR == 1)
                posb[tid] = b;
                const i32 ph = a.bh[b], pt = a.bt[b];
                rowid[tid * per] = ph; rowid[tid * per + 1] = pt;
                for (int m = 0; m < a.k; m++) {
                    const i32 at = b + (m + 1) * a.B;
                    const i32 nh = a.bh[at], nt = a.bt[at];
                    const int sd = nh != ph ? 0 : (nt != pt ? 1 : 2);
                    side[tid * a.k + m] = sd;
                    rowid[tid * per + 2 + m] = sd == 0 ? nh : (sd == 1 ? nt : ph);
                }
            }
            __syncthreads();
            for (int i = tid; i < rows4 * De4; i += TRF_THREADS) {            // gather the entity rows (128-bit)
                const int row = i / De4, q = i - row * De4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < rows) v = __ldg(reinterpret_cast<const float4 *>(a.m.ent + (i64)rowid[row] * De) + q);
                reinterpret_cast<float4 *>(A)[row * De4 + q] = v;
            }
            __syncthreads();
            // ---- P = A . M   (4x4 register tiles)
            for (int t = tid; t < (rows4 >> 2) * Dr4; t += TRF_THREADS) {
                const int tr = t / Dr4, tc = t - tr * Dr4;
                float acc[4][4] = {};
                for (int i = 0; i < De; i += 4) {
                    float4 av[4], mv[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        av[q] = reinterpret_cast<const float4 *>(A)[(tr * 4 + q) * De4 + (i >> 2)];
                        mv[q] = reinterpret_cast<const float4 *>(Msh)[(i + q) * Dr4 + tc];
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        acc[q][0] += av[q].x * mv[0].x + av[q].y * mv[1].x + av[q].z * mv[2].x + av[q].w * mv[3].x;
                        acc[q][1] += av[q].x * mv[0].y + av[q].y * mv[1].y + av[q].z * mv[2].y + av[q].w * mv[3].y;
                        acc[q][2] += av[q].x * mv[0].z + av[q].y * mv[1].z + av[q].z * mv[2].z + av[q].w * mv[3].z;
                        acc[q][3] += av[q].x * mv[0].w + av[q].y * mv[1].w + av[q].z * mv[2].w + av[q].w * mv[3].w;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    reinterpret_cast<float4 *>(P)[(tr * 4 + q) * Dr4 + tc] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
            }
            __syncthreads();
            // ---- normalise the projected rows (tf.nn.l2_normalize, TransR.py:19-23)
            for (int row = warp; row < rows; row += TRF_THREADS / 32) {
                float ss = 0.f;
                for (int k = lane; k < Dr; k += 32) { const float v = P[row * Dr + k]; ss += v * v; }
                ss = wsum_t(ss);
                const float iv = rsqrtf(fmaxf(ss, EPS_NORM));
                for (int k = lane; k < Dr; k += 32) P[row * Dr + k] *= iv;
                if (lane == 0) { inv[row] = iv; proj[row] = ss > EPS_NORM; }
            }
            __syncthreads();
            // ---- scores, hinge and the gradient w.r.t. the projected rows: one warp per positive
            for (int c = warp; c < np; c += TRF_THREADS / 32) {
                const int r0 = c * per;
                const float *Ph = P + r0 * Dr, *Pt = P + (r0 + 1) * Dr;
                float gh[4] = {0.f, 0.f, 0.f, 0.f}, gt[4] = {0.f, 0.f, 0.f, 0.f}, gr[4] = {0.f, 0.f, 0.f, 0.f}, gp[4];
                float sp = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = lane + 32 * i;
                    const float u = k < Dr ? (Ph[k] + rhat[k]) - Pt[k] : 0.f;
                    sp += fabsf(u);
                    gp[i] = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
                }
                sp = wsum_t(sp);
                auto backward = [&](const float *Hrow, const float *Trow, int hrow_i, int trow_i, const float *g, float coef,
                                    float *dh, float *dt) {
                    float d1 = 0.f, d2 = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int k = lane + 32 * i;
                        if (k < Dr) { d1 += g[i] * Hrow[k]; d2 += g[i] * Trow[k]; }
                    }
                    d1 = wsum_t(d1); d2 = wsum_t(d2);
                    if (!proj[hrow_i]) d1 = 0.f;
                    if (!proj[trow_i]) d2 = 0.f;
                    const float ih = inv[hrow_i] * coef, it = inv[trow_i] * coef;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int k = lane + 32 * i;
                        if (k < Dr) {
                            dh[i] += ih * (g[i] - Hrow[k] * d1);
                            dt[i] -= it * (g[i] - Trow[k] * d2);
                            gr[i] += coef * g[i];
                        }
                    }
                };
                float hinge = 0.f;
                int active = 0;
                for (int m = 0; m < a.k; m++) {
                    const int sd = side[c * a.k + m], nrow = r0 + 2 + m;
                    const float *Pn = P + nrow * Dr;
                    const float *Hrow = sd == 0 ? Pn : Ph, *Trow = sd == 1 ? Pn : Pt;
                    float gn[4], gnew[4] = {0.f, 0.f, 0.f, 0.f};
                    float sn = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int k = lane + 32 * i;
                        const float u = k < Dr ? (Hrow[k] + rhat[k]) - Trow[k] : 0.f;
                        sn += fabsf(u);
                        gn[i] = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
                    }
                    sn = wsum_t(sn);
                    const float x = sp - sn + a.margin;
                    if (x >= 0.f) {
                        hinge += x; active++;
                        if (sd == 0) backward(Hrow, Trow, nrow, r0 + 1, gn, -a.w, gnew, gt);
                        else if (sd == 1) backward(Hrow, Trow, r0, nrow, gn, -a.w, gh, gnew);
                        else backward(Hrow, Trow, r0, r0 + 1, gn, -a.w, gh, gt);
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) { const int k = lane + 32 * i; if (k < Dr) G[nrow * Dr + k] = gnew[i]; }
                }
                if (active) backward(Ph, Pt, r0, r0 + 1, gp, a.w * (float)active, gh, gt);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = lane + 32 * i;
                    if (k < Dr) { G[r0 * Dr + k] = gh[i]; G[(r0 + 1) * Dr + k] = gt[i]; Rg[c * Dr + k] = gr[i]; }
                }
                if (lane == 0) a.loss_terms[posb[c]] = hinge;
            }
            for (int i = tid + rows * Dr; i < rows4 * Dr; i += TRF_THREADS) G[i] = 0.f;      // padding rows
            __syncthreads();
            // ---- dA = G . M^T  -> entity gradient rows of this chunk
            for (int t = tid; t < (rows4 >> 2) * De4; t += TRF_THREADS) {
                const int tr = t / De4, ti = t - tr * De4;       // this thread: rows 4tr..4tr+3, columns ti + q*De4
                float acc[4][4] = {};
                for (int k = 0; k < Dr4; k++) {
                    float4 gv[4], mv[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        gv[q] = reinterpret_cast<const float4 *>(G)[(tr * 4 + q) * Dr4 + k];
                        mv[q] = reinterpret_cast<const float4 *>(Msh)[(ti + q * De4) * Dr4 + k];
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++)
#pragma unroll
                        for (int e = 0; e < 4; e++) acc[q][e] += dot4(gv[q], mv[e]);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int row = tr * 4 + q;
                    if (row < rows) {
                        const int c = row / per, j = row - c * per;
                        float *dst = a.gent + ((i64)posb[c] * a.NE + j) * De;
#pragma unroll
                        for (int e = 0; e < 4; e++) dst[ti + e * De4] = acc[q][e];
                    }
                }
            }
            // ---- dM += A^T . G
#pragma unroll
            for (int q = 0; q < TRF_MAXT; q++) {
                const int t = tid + q * TRF_THREADS;
                if (t < ntM) {
                    const int ti = t / Dr4, tk = t - ti * Dr4;
                    for (int row = 0; row < rows; row++) {
                        const float4 av = reinterpret_cast<const float4 *>(A)[row * De4 + ti];
                        const float4 gv = reinterpret_cast<const float4 *>(G)[row * Dr4 + tk];
                        const float as[4] = {av.x, av.y, av.z, av.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                        for (int e = 0; e < 4; e++)
#pragma unroll
                            for (int f2 = 0; f2 < 4; f2++) accM[q][e * 4 + f2] += as[e] * gs[f2];
                    }
                }
            }
            if (tid < Dr) for (int c = 0; c < np; c++) racc += Rg[c * Dr + tid];       // fixed order over positives
        }
        // ---- this relation's update, in place: rel_embeddings[r] and M_r
        if (tid < Dr) rsum[tid] = racc;
        __syncthreads();
        if (tid < Dr) {
            float d = 0.f;
            for (int k = 0; k < Dr; k++) d += rsum[k] * rhat[k];
            const float g = s_invr * (rsum[tid] - (s_projr ? rhat[tid] * d : 0.f));
            const i64 off = (i64)r * Dr + tid;
            float x = a.m.rel[off];
            if (f.adam) {
                float mm = a.m.m_rel[off], vv = a.m.v_rel[off];
                adam_elem(x, mm, vv, g, b1, b2, c1, c2, lr, eps);
                a.m.m_rel[off] = mm; a.m.v_rel[off] = vv;
            } else x -= lr * g;
            a.m.rel[off] = x;
        }
        float *Mg = a.m.rel_aux + (i64)r * MS;
#pragma unroll
        for (int q = 0; q < TRF_MAXT; q++) {
            const int t = tid + q * TRF_THREADS;
            if (t < ntM) {
                const int ti = t / Dr4, tk = t - ti * Dr4;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int idx = (ti * 4 + e) * Dr + tk * 4;
                    float4 xv = *reinterpret_cast<const float4 *>(Msh + idx);
                    float *xs = reinterpret_cast<float *>(&xv);
                    if (f.adam) {
                        float4 mv = *reinterpret_cast<const float4 *>(a.m.m_rel_aux + (i64)r * MS + idx), vv = *reinterpret_cast<const float4 *>(a.m.v_rel_aux + (i64)r * MS + idx);
                        float *ms = reinterpret_cast<float *>(&mv), *vs = reinterpret_cast<float *>(&vv);
#pragma unroll
                        for (int f2 = 0; f2 < 4; f2++) adam_elem(xs[f2], ms[f2], vs[f2], accM[q][e * 4 + f2], b1, b2, c1, c2, lr, eps);
                        *reinterpret_cast<float4 *>(a.m.m_rel_aux + (i64)r * MS + idx) = mv;
                        *reinterpret_cast<float4 *>(a.m.v_rel_aux + (i64)r * MS + idx) = vv;
                    } else {
#pragma unroll
                        for (int f2 = 0; f2 < 4; f2++) xs[f2] -= lr * accM[q][e * 4 + f2];
                    }
                    *reinterpret_cast<float4 *>(Mg + idx) = xv;
                }
            }
        }
        __syncthreads();                                   // every read of this half of the double buffer is done
        cur = nxt; buf ^= 1;
    }
    // ---- TF1 Adam: the dense decay also moves the relations this batch did not touch (g = 0)
    if (f.adam) {
        const int c4 = (Dr + MS) >> 2;
        for (i32 r = f.r_lo + (i32)blockIdx.x; r < f.r_hi; r += stride) {
            if (a.rowhead[a.E + r].x >= 0) continue;
            for (int v = tid; v < c4; v += TRF_THREADS) {
                const int e = v * 4;
                const bool is_rel = e < Dr;
                const i64 off = is_rel ? (i64)r * Dr + e : (i64)r * MS + (e - Dr);
                float *x = (is_rel ? a.m.rel : a.m.rel_aux) + off;
                float *mp = (is_rel ? a.m.m_rel : a.m.m_rel_aux) + off, *vp = (is_rel ? a.m.v_rel : a.m.v_rel_aux) + off;
                float4 xv = *reinterpret_cast<float4 *>(x), mv = *reinterpret_cast<float4 *>(mp), vv = *reinterpret_cast<float4 *>(vp);
                float *xs = reinterpret_cast<float *>(&xv), *ms = reinterpret_cast<float *>(&mv), *vs = reinterpret_cast<float *>(&vv);
#pragma unroll
                for (int q = 0; q < 4; q++) adam_elem(xs[q], ms[q], vs[q], 0.f, b1, b2, c1, c2, lr, eps);
                *reinterpret_cast<float4 *>(mp) = mv; *reinterpret_cast<float4 *>(vp) = vv; *reinterpret_cast<float4 *>(x) = xv;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ general form (rel_neg_rate > 0)
// TransR.py:61-65: with relation negatives EVERY negative is projected by ITS OWN relation's matrix (for an entity negative
// that is the positive's), so a relation negative (h, t, r') needs h.M_r' and t.M_r' — a second 40 KB matrix per negative,
// and its gradients go to M_r' and rel_embeddings[r'].  One CTA per POSITIVE does everything for its group (matrices read
// through L1 / L2, no bucketing): entity rows -> the positive's (2 + k) entity gradient rows (the relation negatives' share
// of d h, d t is added into the positive's rows: same entities), relation side -> one [d rel | d M] gradient row PER
// (positive, relation slot), summed per relation in slot order by transr_rel_update_kernel (general mode).  A slow path
// (770 MB of matrix reads and 390 MB of gradient rows per step at the FB15K shape) for a non-default option; deterministic.
#define TG_THREADS 128
__global__ void __launch_bounds__(TG_THREADS) transr_general_kernel(TrArgs a, i32 kr) {
    extern __shared__ __align__(16) float sm[];
    const int De = a.m.ent_dim, Dr = a.m.rel_dim, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k, per = 2 + k, NR = 1 + kr;
    const i32 b = (i32)blockIdx.x;
    float *A = sm;                                      // [per][De]     gathered entity rows: h, t, entity negatives
    float *P = A + per * De;                            // [per][Dr]     projected by M_r, normalised
    float *G = P + per * Dr;                            // [per][Dr]     gradient w.r.t. the normalised projected rows
    float *P2 = G + per * Dr;                           // [kr][2][Dr]   h, t projected by the negative relation's matrix
    float *G2 = P2 + kr * 2 * Dr;                       // [kr][2][Dr]
    float *rh = G2 + kr * 2 * Dr;                       // [NR][Dr]      r_hat of the positive's and the negatives' relations
    float *gr = rh + NR * Dr;                           // [NR][Dr]      gradient w.r.t. each r_hat
    float *dHT = gr + NR * Dr;                          // [2][De]       d h, d t
    float *inv = dHT + 2 * De;                          // [per + 2 kr]
    i32 *proj = (i32 *)(inv + per + 2 * kr);            // [per + 2 kr]
    i32 *rowid = proj + per + 2 * kr;                   // [per]
    i32 *side = rowid + per;                            // [k]
    i32 *relid = side + k;                              // [NR]          pr, then the negative relations (-1: same triple)
    float *rinv = (float *)(relid + NR);                // [NR]
    i32 *rproj = (i32 *)(rinv + NR);                    // [NR]
    const i32 ph = a.bh[b], pt = a.bt[b], pr = a.br[b];
    if (tid == 0) {
        rowid[0] = ph; rowid[1] = pt; relid[0] = pr;
        for (int m = 0; m < k; m++) {
            const i32 at = b + (m + 1) * a.B, nh = a.bh[at], nt = a.bt[at];
            const int sd = nh != ph ? 0 : (nt != pt ? 1 : 2);
            side[m] = sd; rowid[2 + m] = sd == 0 ? nh : (sd == 1 ? nt : ph);
        }
        for (int m = 0; m < kr; m++) { const i32 nr = a.br[b + (1 + k + m) * a.B]; relid[1 + m] = nr != pr ? nr : -1; }
    }
    __syncthreads();
    for (int i = tid; i < per * De; i += TG_THREADS) A[i] = a.m.ent[(i64)rowid[i / De] * De + i % De];
    for (int i = tid; i < per * Dr; i += TG_THREADS) G[i] = 0.f;
    for (int i = tid; i < kr * 2 * Dr; i += TG_THREADS) { P2[i] = 0.f; G2[i] = 0.f; }
    for (int i = tid; i < NR * Dr; i += TG_THREADS) gr[i] = 0.f;
    for (int i = tid; i < 2 * De; i += TG_THREADS) dHT[i] = 0.f;
    __syncthreads();
    // ---- projections: P[row] = A[row] . M_pr ; P2[m][0/1] = h / t . M_r'   (sequential over the input dimension)
    for (int o = tid; o < (per + 2 * kr) * Dr; o += TG_THREADS) {
        const int row = o / Dr, kk = o - row * Dr;
        const float *src; const float *M;
        if (row < per) { src = A + row * De; M = a.m.rel_aux + (i64)pr * De * Dr; }
        else { const int m = (row - per) >> 1; if (relid[1 + m] < 0) continue; src = A + ((row - per) & 1) * De; M = a.m.rel_aux + (i64)relid[1 + m] * De * Dr; }
        float acc = 0.f;
        for (int i = 0; i < De; i++) acc += src[i] * __ldg(M + (i64)i * Dr + kk);
        if (row < per) P[row * Dr + kk] = acc; else P2[(row - per) * Dr + kk] = acc;
    }
    __syncthreads();
    // ---- l2-normalise every projected row and every r_hat (one warp per vector)
    for (int v = warp; v < per + 2 * kr + NR; v += TG_THREADS / 32) {
        const bool is_rel = v >= per + 2 * kr;
        const int ri = v - (per + 2 * kr);
        if (is_rel ? relid[ri] < 0 : (v >= per && relid[1 + ((v - per) >> 1)] < 0)) continue;
        float *vec = is_rel ? rh + ri * Dr : (v < per ? P + v * Dr : P2 + (v - per) * Dr);
        float ss = 0.f;
        for (int kk = lane; kk < Dr; kk += 32) { const float x = is_rel ? a.m.rel[(i64)relid[ri] * Dr + kk] : vec[kk]; ss += x * x; }
        ss = wsum_t(ss);
        const float iv = rsqrtf(fmaxf(ss, EPS_NORM));
        for (int kk = lane; kk < Dr; kk += 32) vec[kk] = (is_rel ? a.m.rel[(i64)relid[ri] * Dr + kk] : vec[kk]) * iv;
        if (lane == 0) { if (is_rel) { rinv[ri] = iv; rproj[ri] = ss > EPS_NORM; } else { inv[v] = iv; proj[v] = ss > EPS_NORM; } }
    }
    __syncthreads();
    // ---- scores, hinge, gradients w.r.t. the normalised rows and the r_hats: warp 0
    if (warp == 0) {
        const float *Ph = P, *Pt = P + Dr;
        float sp = 0.f;
        for (int kk = lane; kk < Dr; kk += 32) sp += fabsf((Ph[kk] + rh[kk]) - Pt[kk]);
        sp = wsum_t(sp);
        // adds coef * d score(H, T, r_hat) to the gradient rows gH, gT, gR (score = sum |H + r_hat - T|, rows normalised)
        auto backward = [&](const float *H, const float *T, const float *R, int hi, int ti, float coef, float *gH, float *gT, float *gR) {
            float d1 = 0.f, d2 = 0.f;
            for (int kk = lane; kk < Dr; kk += 32) {
                const float u = (H[kk] + R[kk]) - T[kk], g = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
                d1 += g * H[kk]; d2 += g * T[kk];
            }
            d1 = wsum_t(d1); d2 = wsum_t(d2);
            if (!proj[hi]) d1 = 0.f;
            if (!proj[ti]) d2 = 0.f;
            const float ih = inv[hi] * coef, it = inv[ti] * coef;
            for (int kk = lane; kk < Dr; kk += 32) {
                const float u = (H[kk] + R[kk]) - T[kk], g = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
                gH[kk] += ih * (g - H[kk] * d1);
                gT[kk] -= it * (g - T[kk] * d2);
                gR[kk] += coef * g;
            }
        };
        float hinge = 0.f;
        int active = 0;
        for (int m = 0; m < k; m++) {                      // entity negatives: the positive's matrix and r_hat
            const int sd = side[m], nrow = 2 + m;
            if (sd == 2) { hinge += a.margin; continue; }  // same triple: the term is the margin, its gradient is zero
            const float *H = sd == 0 ? P + nrow * Dr : Ph, *T = sd == 1 ? P + nrow * Dr : Pt;
            float sn = 0.f;
            for (int kk = lane; kk < Dr; kk += 32) sn += fabsf((H[kk] + rh[kk]) - T[kk]);
            sn = wsum_t(sn);
            const float x = sp - sn + a.margin;
            if (x >= 0.f) {
                hinge += x; active++;
                backward(H, T, rh, sd == 0 ? nrow : 0, sd == 1 ? nrow : 1, -a.w, sd == 0 ? G + nrow * Dr : G, sd == 1 ? G + nrow * Dr : G + Dr, gr);
            }
        }
        for (int m = 0; m < kr; m++) {                     // relation negatives: their own matrix and r_hat
            if (relid[1 + m] < 0) { hinge += a.margin; continue; }
            const float *H = P2 + (2 * m) * Dr, *T = P2 + (2 * m + 1) * Dr, *R = rh + (1 + m) * Dr;
            float sn = 0.f;
            for (int kk = lane; kk < Dr; kk += 32) sn += fabsf((H[kk] + R[kk]) - T[kk]);
            sn = wsum_t(sn);
            const float x = sp - sn + a.margin;
            if (x >= 0.f) {
                hinge += x; active++;
                backward(H, T, R, per + 2 * m, per + 2 * m + 1, -a.w, G2 + (2 * m) * Dr, G2 + (2 * m + 1) * Dr, gr + (1 + m) * Dr);
            }
        }
        if (active) backward(Ph, Pt, rh, 0, 1, a.w * (float)active, G, G + Dr, gr);
        if (lane == 0) a.loss_terms[b] = hinge;
    }
    __syncthreads();
    // ---- entity gradient rows: dA = G . M^T; the relation negatives' share of d h, d t goes into the positive's rows
    for (int o = tid; o < per * De; o += TG_THREADS) {
        const int row = o / De, i = o - row * De;
        const float *M = a.m.rel_aux + ((i64)pr * De + i) * Dr;
        float acc = 0.f;
        for (int kk = 0; kk < Dr; kk++) acc += G[row * Dr + kk] * __ldg(M + kk);
        if (row < 2)
            for (int m = 0; m < kr; m++) {
                if (relid[1 + m] < 0) continue;
                const float *M2 = a.m.rel_aux + ((i64)relid[1 + m] * De + i) * Dr;
                float a2 = 0.f;
                for (int kk = 0; kk < Dr; kk++) a2 += G2[(2 * m + row) * Dr + kk] * __ldg(M2 + kk);
                acc += a2;
            }
        a.gent[((i64)b * a.NE + row) * De + i] = acc;
    }
    // ---- relation gradient rows [d rel (Dr) | d M (De x Dr)], one per (positive, relation slot); unused slots are not read
    const i64 cols = (i64)Dr + (i64)De * Dr;
    for (int j = 0; j < NR; j++) {
        if (relid[j] < 0) continue;
        float *out = a.grel + ((i64)b * NR + j) * cols;
        const float *R = rh + j * Dr, *gR = gr + j * Dr;
        float d = 0.f;                                     // every thread: dot(gR, R) in index order (same bits everywhere)
        for (int kk = 0; kk < Dr; kk++) d += gR[kk] * R[kk];
        for (int kk = tid; kk < Dr; kk += TG_THREADS) out[kk] = rinv[j] * (gR[kk] - (rproj[j] ? R[kk] * d : 0.f));
        for (int o = tid; o < De * Dr; o += TG_THREADS) {
            const int i = o / Dr, kk = o - i * Dr;
            float acc = 0.f;
            if (j == 0) for (int row = 0; row < per; row++) acc += A[row * De + i] * G[row * Dr + kk];
            else acc = A[i] * G2[(2 * (j - 1)) * Dr + kk] + A[De + i] * G2[(2 * (j - 1) + 1) * Dr + kk];
            out[Dr + o] = acc;
        }
    }
}

// rel_embeddings and transfer_matrix rows: gradients arrive already reduced per relation.
//   SGD : touched relations only, x -= lr g.      Adam: every relation (TF1 dense decay), g = 0 if untouched.
struct TrUpdArgs {
    okb_model m;
    okb_hyper hp;
    const int4 *rowhead;
    const float *grel;
    const unsigned *bad;       // "bad id" flag of the host-batch path: set -> leave the tables alone
    i32 E, R, adam, r_lo;
    // general form (rel_neg_rate > 0): grel holds one row per (positive, relation slot); a relation's gradient is the sum of
    // its segment's rows in slot order (perm: slot of each sorted position, nes: first relation slot)
    const i32 *perm;
    i32 general, nes;
};
__global__ void __launch_bounds__(256) transr_rel_update_kernel(TrUpdArgs a) {
    const i32 r = a.r_lo + (i32)blockIdx.x;
    const int De = a.m.ent_dim, Dr = a.m.rel_dim;
    const bool touched = a.rowhead[a.E + r].x >= 0;
    if (!touched && !a.adam) return;
    if (a.bad && *(const volatile unsigned *)a.bad) return;
    const int cols = Dr + De * Dr, c4 = cols >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(a.grel + (i64)r * cols);
    for (int v = blockIdx.y * blockDim.x + threadIdx.x; v < c4; v += gridDim.y * blockDim.x) {
        const int e = v * 4;
        const bool is_rel = e < Dr;
        const i64 off = is_rel ? (i64)r * Dr + e : (i64)r * De * Dr + (e - Dr);
        float *x = (is_rel ? a.m.rel : a.m.rel_aux) + off;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (touched && !a.general) g = __ldg(g4 + v);
        else if (touched) {
            const int4 seg = a.rowhead[a.E + r];
            for (i32 j = seg.x; j < seg.y; j++) {
                const float4 y = __ldg(reinterpret_cast<const float4 *>(a.grel + (i64)(__ldg(a.perm + j) - a.nes) * cols) + v);
                g.x += y.x; g.y += y.y; g.z += y.z; g.w += y.w;
            }
        }
        float4 xv = *reinterpret_cast<float4 *>(x);
        const float gs[4] = {g.x, g.y, g.z, g.w};
        float *xs = reinterpret_cast<float *>(&xv);
        if (a.adam) {
            float *mp = (is_rel ? a.m.m_rel : a.m.m_rel_aux) + off, *vp = (is_rel ? a.m.v_rel : a.m.v_rel_aux) + off;
            float4 mv = *reinterpret_cast<float4 *>(mp), vv = *reinterpret_cast<float4 *>(vp);
            float *ms = reinterpret_cast<float *>(&mv), *vs = reinterpret_cast<float *>(&vv);
#pragma unroll
            for (int q = 0; q < 4; q++)
                adam_elem(xs[q], ms[q], vs[q], gs[q], a.hp.beta1, a.hp.beta2, 1.f - a.hp.beta1, 1.f - a.hp.beta2, a.hp.lr, a.hp.eps);
            *reinterpret_cast<float4 *>(mp) = mv; *reinterpret_cast<float4 *>(vp) = vv;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) xs[q] -= a.hp.lr * gs[q];
        }
        *reinterpret_cast<float4 *>(x) = xv;
    }
}

int okb_transr_check(okb_ctx *c, const okb_model *m) {
    if (!m->rel_aux) OKB_FAIL(c, OKB_ERR_ARG, "transfer_matrix table missing");
    if (m->ent_dim % 4 || m->rel_dim % 4 || m->ent_dim > 128 || m->rel_dim > 128)
        OKB_FAIL(c, OKB_ERR_ARG, "TransR training needs ent_size, rel_size multiples of 4 and <= 128");
    if ((m->ent_dim / 4) * (m->rel_dim / 4) > TR_MAXT * TR_THREADS) OKB_FAIL(c, OKB_ERR_ARG, "TransR matrix too large");
    if (c->KR == 0 && 2 + c->K > TR_ROWS) OKB_FAIL(c, OKB_ERR_ARG, "TransR training: ent_neg_rate too large for the chunk size");
    if (c->KR != 0 && (c->tr_hi > c->tr_lo)) OKB_FAIL(c, OKB_ERR_ARG, "TransR with rel_neg_rate > 0 cannot be relation-sharded (a relation negative belongs to two shards)");
    return 0;
}

int okb_transr_launch_grad(okb_ctx *c, const okb_model *m, const okb_hyper *hp, const i32 *batch, const i32 *skeys,
                           const i32 *perm, const int4 *rowhead, i64 n, INT b_lo, INT b_hi, float *gent, float *grel,
                           float *loss_terms, cudaStream_t s) {
    int rc = okb_transr_check(c, m);
    if (rc) return rc;
    if (b_lo != 0 || b_hi != c->B) OKB_FAIL(c, OKB_ERR_ARG, "TransR gradients are reduced per relation: shard the RELATIONS (okb_transr_set_shard), not the positives");
    const i64 S = c->B * (1 + c->K + c->KR);
    TrArgs a;
    a.m = *m; a.bh = batch; a.bt = batch + S; a.br = batch + 2 * S;
    a.skeys = skeys; a.perm = perm; a.rowhead = rowhead;
    a.gent = gent; a.grel = grel; a.loss_terms = loss_terms;
    a.margin = hp->margin; a.w = 1.0f / (float)(c->B * (c->K + c->KR));
    a.B = (i32)c->B; a.k = (i32)c->K; a.NE = (i32)(2 + c->K); a.E = (i32)c->E; a.R = (i32)c->R; a.n = (i32)n;
    a.nes = (i32)c->plan_ne; a.b_lo = (i32)b_lo; a.b_hi = (i32)b_hi;
    a.CH = (i32)(TR_ROWS / (2 + c->K));
    const i64 r_lo = c->tr_hi > c->tr_lo ? c->tr_lo : 0, r_hi = c->tr_hi > c->tr_lo ? c->tr_hi : c->R;
    a.r_lo = (i32)r_lo;
    const int De = m->ent_dim, Dr = m->rel_dim;
    if (c->KR > 0) {                                       // relation negatives: the general form, one CTA per positive
        const i64 per = 2 + c->K, kr = c->KR, NR = 1 + kr;
        const size_t gsmem = sizeof(float) * (size_t)(per * De + 2 * per * Dr + 4 * kr * Dr + 2 * NR * Dr + 2 * De + (per + 2 * kr) + NR) +
                             sizeof(i32) * (size_t)((per + 2 * kr) + per + c->K + 2 * NR) + 64;
        if (gsmem > 200 * 1024) OKB_FAIL(c, OKB_ERR_ARG, "TransR with relation negatives: too many negatives for shared memory");
        OKB_CUDA(c, okb_smem_optin(c, transr_general_kernel, gsmem));
        c->transr_rel_done = false;
        {
            ProfScope ps(c, PROF_GRAD, s);
            transr_general_kernel<<<(unsigned)c->B, TG_THREADS, gsmem, s>>>(a, (i32)kr);
        }
        OKB_LAUNCHED(1);
        OKB_CUDA(c, cudaGetLastError());
        return 0;
    }
    if (c->transr_fused && loss_terms) {
        // persistent form with the relation-side update applied in place (transr_fused_kernel); grel is not used
        const size_t fsmem = sizeof(float) * (2 * (size_t)De * Dr + (size_t)TR_ROWS * De + 2 * (size_t)TR_ROWS * Dr + (size_t)a.CH * Dr + 2 * Dr + TR_ROWS) +
                             sizeof(i32) * (2 * TR_ROWS + a.CH + (size_t)a.CH * a.k) + 128;
        const bool adam = m->optimizer == OKB_ADAM;
        if (fsmem <= 224 * 1024 && (De / 4) * (Dr / 4) <= TRF_MAXT * TRF_THREADS && (!adam || (m->m_rel && m->v_rel && m->m_rel_aux && m->v_rel_aux))) {
            TrFuse f;
            f.hp = *hp; f.adam = adam ? 1 : 0; f.r_lo = (i32)r_lo; f.r_hi = (i32)r_hi;
            f.bad = c->batch_from_host && c->flags.p ? c->flags.as<unsigned>() + OKB_FLAGS_BAD : nullptr;
            OKB_CUDA(c, okb_smem_optin(c, transr_fused_kernel, fsmem));
            const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>(okb_sms(c), r_hi - r_lo));
            {
                ProfScope ps(c, PROF_GRAD, s);
                transr_fused_kernel<<<grid, TRF_THREADS, fsmem, s>>>(a, f);
            }
            OKB_LAUNCHED(1);
            OKB_CUDA(c, cudaGetLastError());
            c->transr_rel_done = true;                     // okb_update: the relation rows of this step are already updated
            return 0;
        }
    }
    c->transr_rel_done = false;
    const size_t smem = sizeof(float) * ((size_t)De * Dr + (size_t)TR_ROWS * De + 2 * (size_t)TR_ROWS * Dr + (size_t)a.CH * Dr + 2 * Dr + TR_ROWS) +
                        sizeof(i32) * (2 * TR_ROWS + a.CH + (size_t)a.CH * a.k) + 64;
    if (smem > 226 * 1024) OKB_FAIL(c, OKB_ERR_ARG, "TransR dimensions too large for shared memory");
    OKB_CUDA(c, okb_smem_optin(c, transr_bucket_kernel, smem));      // static + dynamic must stay within 227 KB
    {
        ProfScope ps(c, PROF_GRAD, s);
        transr_bucket_kernel<<<(unsigned)(r_hi - r_lo), TR_THREADS, smem, s>>>(a);
    }
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

int okb_transr_launch_rel_update(okb_ctx *c, const okb_model *m, const okb_hyper *hp, const int4 *rowhead, const float *grel,
                                 cudaStream_t s, const i32 *perm) {
    if (c->transr_rel_done) { c->transr_rel_done = false; return 0; }      // applied in place by transr_fused_kernel
    TrUpdArgs a;
    a.m = *m; a.hp = *hp; a.rowhead = rowhead; a.grel = grel; a.E = (i32)c->E; a.R = (i32)c->R;
    a.general = c->KR > 0 ? 1 : 0; a.perm = perm; a.nes = (i32)c->plan_ne;
    if (a.general && !perm) OKB_FAIL(c, OKB_ERR_STATE, "TransR with relation negatives: the step's plan is missing");
    a.adam = m->optimizer == OKB_ADAM;
    a.bad = c->batch_from_host && c->flags.p ? c->flags.as<unsigned>() + OKB_FLAGS_BAD : nullptr;
    if (a.adam && (!m->m_rel_aux || !m->v_rel_aux)) OKB_FAIL(c, OKB_ERR_ARG, "Adam slots missing");
    const int cols = m->rel_dim + m->ent_dim * m->rel_dim;
    const unsigned gy = (unsigned)std::max(1, std::min(8, (cols / 4 + 255) / 256));
    const i64 r_lo = c->tr_hi > c->tr_lo ? c->tr_lo : 0, r_hi = c->tr_hi > c->tr_lo ? c->tr_hi : c->R;
    a.r_lo = (i32)r_lo;
    transr_rel_update_kernel<<<dim3((unsigned)(r_hi - r_lo), gy), 256, 0, s>>>(a);
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

extern "C" int okb_transr_set_shard(okb_ctx *c, INT r_lo, INT r_hi) {
    if (r_lo == 0 && r_hi == 0) { c->tr_lo = c->tr_hi = 0; return 0; }      // back to "all relations"
    if (r_lo < 0 || r_hi > c->R || r_lo >= r_hi) OKB_FAIL(c, OKB_ERR_ARG, "bad relation shard");
    c->tr_lo = r_lo; c->tr_hi = r_hi;
    return 0;
}
