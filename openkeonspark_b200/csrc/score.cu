// Scoring and link-prediction ranking.
//
//   okb_predict        TransX.predict_def on arbitrary (h,t,r) lists (Config.test_step)
//   okb_rank           getHeadBatch/getTailBatch + predict + testHead/testTail (Test.h:11-249,
//                      distribute_training.py:465-590) for a range of test triples: every candidate
//                      entity is scored against every query and the raw / filtered /
//                      type-constrained better-than counts and argmins are accumulated on the GPU
//   okb_rank_finalize  counts + argmins -> the reference's 8-int records (incl. ontology classes)
//   okb_rank_scores    the reference-shaped call: one query, caller-provided scores[E]
//
// Canonical fp32 order (shared with oracle/kge_oracle.c so scores, hence ranks, are bit-identical):
// reductions over the embedding dimension run sequentially d = 0..D-1, every multiply and add is
// individually rounded (__fmul_rn/__fadd_rn: never contracted to FMA), 1/sqrt = IEEE sqrt + divide.
//
// Ranking layout: the test list is sorted by relation (Reader.h:260), so queries are processed in
// relation groups.  Per group the (projected, normalised) candidate table is materialised once,
// TRANSPOSED ([D][E_pad]: thread j reads column j, coalesced), tiles of 128 candidates x D are
// staged into shared memory with bulk-async (TMA) copies, and each thread keeps QB query
// accumulators per side in registers: 2 FADD per (query, candidate, dim) — FP32-ALU bound; L1
// distance has no tensor-core form.
#include <algorithm>
#include <cstdlib>

#include "okb_internal.h"

#define FULL 0xffffffffu
#define EPS_NORM 1e-12f
#define CT 128                 // candidates per tile (= threads per CTA)
#define QB 8                   // queries per register block

// ------------------------------------------------------------------------------------------ canonical helpers
__device__ __forceinline__ float c_inv_norm(float ss) { return __fdiv_rn(1.0f, __fsqrt_rn(ss > EPS_NORM ? ss : EPS_NORM)); }

// sequential dot over a strided vector pair
__device__ __forceinline__ float c_dot(const float *a, int sa, const float *b, int sb, int D) {
    float s = 0.f;
    for (int d = 0; d < D; d++) s = __fadd_rn(s, __fmul_rn(a[d * sa], b[d * sb]));
    return s;
}
__device__ __forceinline__ float c_finish(int model, float s, int D) { return model == OKB_TRANSE ? __fdiv_rn(s, (float)D) : s; }

// ------------------------------------------------------------------------------------------ predict (any triples)
// One thread per triple; rows are re-read from global/L2 in several sequential passes instead of
// being held in local arrays.  Not a hot path (triple classification, predict_*).
struct PredArgs {
    okb_model m;
    const i64 *h, *t, *r;
    float *out;
    i64 n;
    i32 E, R;
};

#define TRANSR_MAXD 256
// TransR.py:77-87: every row is projected by the matrix of predict_r[0] (r0), its own r only picks rel_embeddings
__device__ float pred_one_transr(const okb_model &m, i64 h, i64 t, i64 r, i64 r0) {
    const int De = m.ent_dim, Dr = m.rel_dim;
    const float *eh = m.ent + h * De, *et = m.ent + t * De, *er = m.rel + r * Dr, *M = m.rel_aux + r0 * (i64)De * Dr;
    float hp[TRANSR_MAXD], tp[TRANSR_MAXD];
    float sh = 0.f, st = 0.f;
    for (int k = 0; k < Dr; k++) {
        float a = 0.f, b = 0.f;
        for (int d = 0; d < De; d++) {
            const float mk = M[(i64)d * Dr + k];
            a = __fadd_rn(a, __fmul_rn(eh[d], mk));
            b = __fadd_rn(b, __fmul_rn(et[d], mk));
        }
        hp[k] = a; tp[k] = b;
    }
    for (int k = 0; k < Dr; k++) { sh = __fadd_rn(sh, __fmul_rn(hp[k], hp[k])); st = __fadd_rn(st, __fmul_rn(tp[k], tp[k])); }
    const float ih = c_inv_norm(sh), it = c_inv_norm(st), ir = c_inv_norm(c_dot(er, 1, er, 1, Dr));
    float s = 0.f;
    for (int k = 0; k < Dr; k++) {
        const float a = __fadd_rn(__fmul_rn(hp[k], ih), __fmul_rn(er[k], ir));
        s = __fadd_rn(s, fabsf(__fsub_rn(a, __fmul_rn(tp[k], it))));
    }
    return s;
}

// canonical score of one triple from its rows (all unit stride): eh / et entity rows, er relation row,
// aux = normal vector (TransH) or rel_transfer (TransD), eth / ett = ent_transfer rows (TransD)
__device__ __forceinline__ float pred_rows(int model, int D, const float *eh, const float *et, const float *er, const float *aux,
                                           const float *eth, const float *ett) {
    const float inv_r = c_inv_norm(c_dot(er, 1, er, 1, D));
    float s = 0.f;
    if (model == OKB_TRANSE) {
        const float ih = c_inv_norm(c_dot(eh, 1, eh, 1, D)), it = c_inv_norm(c_dot(et, 1, et, 1, D));
        for (int d = 0; d < D; d++) {
            const float a = __fadd_rn(__fmul_rn(eh[d], ih), __fmul_rn(er[d], inv_r));
            s = __fadd_rn(s, fabsf(__fsub_rn(a, __fmul_rn(et[d], it))));
        }
    } else if (model == OKB_TRANSH) {
        const float *n = aux;
        const float in = c_inv_norm(c_dot(n, 1, n, 1, D));
        float dh = 0.f, dt = 0.f;
        for (int d = 0; d < D; d++) {
            const float nh = __fmul_rn(n[d], in);
            dh = __fadd_rn(dh, __fmul_rn(eh[d], nh));
            dt = __fadd_rn(dt, __fmul_rn(et[d], nh));
        }
        float sh = 0.f, st = 0.f;
        for (int d = 0; d < D; d++) {
            const float nh = __fmul_rn(n[d], in);
            const float ph = __fsub_rn(eh[d], __fmul_rn(dh, nh)), pt = __fsub_rn(et[d], __fmul_rn(dt, nh));
            sh = __fadd_rn(sh, __fmul_rn(ph, ph));
            st = __fadd_rn(st, __fmul_rn(pt, pt));
        }
        const float ih = c_inv_norm(sh), it = c_inv_norm(st);
        for (int d = 0; d < D; d++) {
            const float nh = __fmul_rn(n[d], in);
            const float ph = __fsub_rn(eh[d], __fmul_rn(dh, nh)), pt = __fsub_rn(et[d], __fmul_rn(dt, nh));
            const float a = __fadd_rn(__fmul_rn(ph, ih), __fmul_rn(er[d], inv_r));
            s = __fadd_rn(s, fabsf(__fsub_rn(a, __fmul_rn(pt, it))));
        }
    } else {  // TransD
        const float *rt = aux;
        const float ch = c_dot(eh, 1, eth, 1, D), ct = c_dot(et, 1, ett, 1, D);
        float sh = 0.f, st = 0.f;
        for (int d = 0; d < D; d++) {
            const float ph = __fadd_rn(eh[d], __fmul_rn(ch, rt[d])), pt = __fadd_rn(et[d], __fmul_rn(ct, rt[d]));
            sh = __fadd_rn(sh, __fmul_rn(ph, ph));
            st = __fadd_rn(st, __fmul_rn(pt, pt));
        }
        const float ih = c_inv_norm(sh), it = c_inv_norm(st);
        for (int d = 0; d < D; d++) {
            const float ph = __fadd_rn(eh[d], __fmul_rn(ch, rt[d])), pt = __fadd_rn(et[d], __fmul_rn(ct, rt[d]));
            const float a = __fadd_rn(__fmul_rn(ph, ih), __fmul_rn(er[d], inv_r));
            s = __fadd_rn(s, fabsf(__fsub_rn(a, __fmul_rn(pt, it))));
        }
    }
    return c_finish(model, s, D);
}
__device__ float pred_one(const okb_model &m, i64 h, i64 t, i64 r) {
    const int D = m.ent_dim;
    return pred_rows(m.model, D, m.ent + h * D, m.ent + t * D, m.rel + r * D, m.rel_aux ? m.rel_aux + r * D : nullptr,
                     m.ent_aux ? m.ent_aux + h * D : nullptr, m.ent_aux ? m.ent_aux + t * D : nullptr);
}

// Staged form for TransE / TransH / TransD: a warp owns 32 triples, copies their rows into shared memory with coalesced
// loads (lane = column) and then every lane walks ITS triple's rows in the canonical sequential order (row stride D + 1:
// conflict-free).  Same arithmetic as pred_one — same bits — without 32 strided global streams per warp.
__global__ void __launch_bounds__(64) predict_stage_kernel(PredArgs a, int narr) {
    extern __shared__ float psm[];
    const int D = a.m.ent_dim, lane = threadIdx.x & 31, w = threadIdx.x >> 5, sd = D + 1;
    float *base = psm + (size_t)w * narr * 32 * sd;
    const i64 i0 = ((i64)blockIdx.x * (blockDim.x >> 5) + w) * 32;
    for (int rr = 0; rr < 32; rr++) {
        const i64 i = i0 + rr;
        if (i >= a.n) break;
        const i64 h = a.h[i], t = a.t[i], r = a.r[i];
        if (h < 0 || h >= a.E || t < 0 || t >= a.E || r < 0 || r >= a.R) continue;
        const float *src[6] = {a.m.ent + h * D, a.m.ent + t * D, a.m.rel + r * D, a.m.rel_aux ? a.m.rel_aux + r * D : nullptr,
                               a.m.ent_aux ? a.m.ent_aux + h * D : nullptr, a.m.ent_aux ? a.m.ent_aux + t * D : nullptr};
        for (int q = 0; q < narr; q++) {
            float *dst = base + ((size_t)q * 32 + rr) * sd;
            for (int d = lane; d < D; d += 32) dst[d] = __ldg(src[q] + d);
        }
    }
    __syncwarp();
    const i64 i = i0 + lane;
    if (i >= a.n) return;
    const i64 h = a.h[i], t = a.t[i], r = a.r[i];
    if (h < 0 || h >= a.E || t < 0 || t >= a.E || r < 0 || r >= a.R) { a.out[i] = __int_as_float(0x7fc00000); return; }
    auto row = [&](int q) { return base + ((size_t)q * 32 + lane) * sd; };
    a.out[i] = pred_rows(a.m.model, D, row(0), row(1), row(2), narr > 3 ? row(3) : nullptr, narr > 4 ? row(4) : nullptr, narr > 5 ? row(5) : nullptr);
}

__global__ void __launch_bounds__(128) predict_kernel(PredArgs a) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const i64 h = a.h[i], t = a.t[i], r = a.r[i];
    if (h < 0 || h >= a.E || t < 0 || t >= a.E || r < 0 || r >= a.R) { a.out[i] = __int_as_float(0x7fc00000); return; }
    a.out[i] = a.m.model == OKB_TRANSR ? pred_one_transr(a.m, h, t, r, a.r[0]) : pred_one(a.m, h, t, r);
}

// ------------------------------------------------------------------------------------------ rank: preparation
// relation vectors of a group: rv[g][0][D] = unit(rel[r]); rv[g][1][D] = unit(normal[r]) (TransH) or rel_transfer[r] (TransD)
__global__ void relvec_kernel(okb_model m, const i32 *__restrict__ grp_rel, float *__restrict__ rv, i32 G) {
    const i32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int D = m.rel_dim;
    const i64 r = grp_rel[g];
    const float *er = m.rel + r * D;
    float *o = rv + (i64)g * 2 * D;
    const float inv = c_inv_norm(c_dot(er, 1, er, 1, D));
    for (int d = 0; d < D; d++) o[d] = __fmul_rn(er[d], inv);
    if (m.model == OKB_TRANSH) {
        const float *n = m.rel_aux + r * D;
        const float in = c_inv_norm(c_dot(n, 1, n, 1, D));
        for (int d = 0; d < D; d++) o[D + d] = __fmul_rn(n[d], in);
    } else if (m.model == OKB_TRANSD) {
        const float *rt = m.rel_aux + r * D;
        for (int d = 0; d < D; d++) o[D + d] = rt[d];
    }
}

// Canonical transfer + l2-normalise of one entity row held in shared memory (lane-private, unit
// stride).  `in` has Din floats; the result (D floats) is left in `out` (== in except for TransR, whose
// projection e . M_r changes the dimension).  aux: TransH n_hat / TransD rel_transfer; et_row: TransD
// ent_transfer[e]; Msm: TransR matrix of the group, row-major [Din][D] in shared memory.
__device__ __forceinline__ void canon_row(int model, const float *in, float *out, const float *aux, const float *et_row,
                                          const float *Msm, int Din, int D, const float *pre_dot = nullptr) {
    if (model == OKB_TRANSH) {
        const float dh = c_dot(in, 1, aux, 1, D);
        for (int d = 0; d < D; d++) out[d] = __fsub_rn(in[d], __fmul_rn(dh, aux[d]));
    } else if (model == OKB_TRANSD) {
        const float ch = pre_dot ? *pre_dot : c_dot(in, 1, et_row, 1, D);      // e . e_transfer does not depend on the relation
        for (int d = 0; d < D; d++) out[d] = __fadd_rn(in[d], __fmul_rn(ch, aux[d]));
    } else if (model == OKB_TRANSR) {
        if ((D & 3) == 0 && (((size_t)Msm) & 15) == 0) {
            // four outputs per pass over the input row: one broadcast 128-bit load of M_r per 4 multiply-adds instead of one
            // 32-bit load each; every output still accumulates d = 0..Din-1 in order with individually rounded operations
            int k = 0;
            for (; k + 8 <= D; k += 8) {                       // eight outputs per pass where they fit
                float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int d = 0; d < Din; d++) {
                    const float x = in[d];
                    const float4 m0 = *reinterpret_cast<const float4 *>(Msm + d * D + k), m1 = *reinterpret_cast<const float4 *>(Msm + d * D + k + 4);
                    a[0] = __fadd_rn(a[0], __fmul_rn(x, m0.x)); a[1] = __fadd_rn(a[1], __fmul_rn(x, m0.y));
                    a[2] = __fadd_rn(a[2], __fmul_rn(x, m0.z)); a[3] = __fadd_rn(a[3], __fmul_rn(x, m0.w));
                    a[4] = __fadd_rn(a[4], __fmul_rn(x, m1.x)); a[5] = __fadd_rn(a[5], __fmul_rn(x, m1.y));
                    a[6] = __fadd_rn(a[6], __fmul_rn(x, m1.z)); a[7] = __fadd_rn(a[7], __fmul_rn(x, m1.w));
                }
#pragma unroll
                for (int q = 0; q < 8; q++) out[k + q] = a[q];
            }
            for (; k < D; k += 4) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                for (int d = 0; d < Din; d++) {
                    const float x = in[d];
                    const float4 m4 = *reinterpret_cast<const float4 *>(Msm + d * D + k);
                    a0 = __fadd_rn(a0, __fmul_rn(x, m4.x)); a1 = __fadd_rn(a1, __fmul_rn(x, m4.y));
                    a2 = __fadd_rn(a2, __fmul_rn(x, m4.z)); a3 = __fadd_rn(a3, __fmul_rn(x, m4.w));
                }
                out[k] = a0; out[k + 1] = a1; out[k + 2] = a2; out[k + 3] = a3;
            }
        } else {
            for (int k = 0; k < D; k++) {
                float a = 0.f;
                for (int d = 0; d < Din; d++) a = __fadd_rn(a, __fmul_rn(in[d], Msm[d * D + k]));
                out[k] = a;
            }
        }
    }
    const float inv = c_inv_norm(c_dot(out, 1, out, 1, D));
    for (int d = 0; d < D; d++) out[d] = __fmul_rn(out[d], inv);
}

// Candidate tables: out[tab][d][j - j0] for entities j in [j0, j0 + ncol) (ncol multiple of CT, zero padded).
// A warp stages 32 entity rows in shared memory (coalesced), each lane canonicalises its row,
// and the result is written transposed (coalesced over j for every d).
struct CandArgs {
    okb_model m;
    const float *rv;           // [G][2][D]
    const i32 *grp_rel;        // relation of each group (TransR: selects M_r)
    float *out;                // [ntab][D][ncol]
    const float *entdot;       // TransD: canonical e . ent_transfer[e] per candidate column, computed once per call
    i32 j0, ncol, E, ntab;
};
// TransD: the dot e . e_transfer of every candidate (sequential canonical order), shared by all relation groups
__global__ void entdot_kernel(okb_model m, float *out, i32 j0, i32 ncol, i32 E) {
    const i32 col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const i32 j = j0 + col;
    float s = 0.f;
    if (j < E) {
        const float *e = m.ent + (i64)j * m.ent_dim, *t = m.ent_aux + (i64)j * m.ent_dim;
        for (int d = 0; d < m.ent_dim; d++) s = __fadd_rn(s, __fmul_rn(__ldg(e + d), __ldg(t + d)));
    }
    out[col] = s;
}
// dynamic shared memory: per warp 32 input rows [Din+1] (+32 rows of a second array: TransD ent_transfer /
// TransR projected output [D+1]); per block the group's aux vector [D] (+ TransR: M_r [Din*D])
__global__ void __launch_bounds__(128) cand_kernel(CandArgs a) {
    extern __shared__ float sm[];
    const int wpb = blockDim.x >> 5;
    const int Din = a.m.ent_dim, D = a.m.rel_dim, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sin = Din + 1, s2 = (a.m.model == OKB_TRANSR ? D : Din) + 1;
    const bool second = (a.m.model == OKB_TRANSD && !a.entdot) || a.m.model == OKB_TRANSR;
    const size_t per_warp = (size_t)32 * sin + (second ? (size_t)32 * s2 : 0);
    float *rows = sm + (size_t)w * per_warp;
    float *rows2 = rows + 32 * sin;
    float *aux = sm + (size_t)wpb * per_warp;              // [D]
    float *Msm = aux + ((D + 3) & ~3);                     // [Din*D] (TransR)
    const i32 tab = blockIdx.y;
    const i32 jbase = a.j0 + (blockIdx.x * wpb + w) * 32;
    if (a.m.model == OKB_TRANSH || a.m.model == OKB_TRANSD)
        for (int d = threadIdx.x; d < D; d += blockDim.x) aux[d] = a.rv[((i64)tab * 2 + 1) * D + d];
    if (a.m.model == OKB_TRANSR) {
        const float *M = a.m.rel_aux + (i64)a.grp_rel[tab] * Din * D;
        for (int d = threadIdx.x; d < Din * D; d += blockDim.x) Msm[d] = M[d];
    }
    for (int idx = lane; idx < 32 * Din; idx += 32) {
        const int rr = idx / Din, d = idx - rr * Din;
        const i32 j = jbase + rr;
        rows[rr * sin + d] = j < a.E ? a.m.ent[(i64)j * Din + d] : 0.f;
        if (a.m.model == OKB_TRANSD && !a.entdot) rows2[rr * s2 + d] = j < a.E ? a.m.ent_aux[(i64)j * Din + d] : 0.f;
    }
    __syncthreads();
    const i32 j = jbase + lane;
    float *res = a.m.model == OKB_TRANSR ? rows2 + lane * s2 : rows + lane * sin;
    if (j < a.E) canon_row(a.m.model, rows + lane * sin, res, aux, rows2 + lane * s2, Msm, Din, D,
                           a.entdot ? a.entdot + (jbase - a.j0 + lane) : nullptr);
    __syncwarp();
    const i32 col = jbase - a.j0 + lane;
    if (col < a.ncol)
        for (int d = 0; d < D; d++) a.out[((i64)tab * D + d) * a.ncol + col] = j < a.E ? res[d] : 0.f;
}

// Query vectors for test triples [q_lo, q_hi): qa[q][d] = fl(h_hat[d] + r_hat[d]) (tail side),
// qt[q][d] = t_hat[d] (head side), ref[q] = score of the true triple (shared by both sides).
struct QArgs {
    okb_model m;
    const i32 *th, *tt, *tr;
    const i32 *q_group;        // group index of each query (relative to chunk)
    const i32 *grp_rel;
    const i32 *g_qlo, *g_blk;  // first query / first 8-query block of each group
    const float *rv;
    float *qa, *qt;            // blocked: [block][d][8], blocks of a group contiguous (zero padded)
    float *ref, *thr;          // per query: finished score of the true triple; raw-sum threshold (see below)
    i32 q_lo, nq;
};
// dynamic shared memory per warp: H, T input rows [32][Din+1]; a second pair (TransD: ent_transfer rows,
// TransR: projected outputs [32][D+1]).  TransR reads M_r from global memory (queries of a warp may span groups).
__global__ void __launch_bounds__(128) qvec_kernel(QArgs a) {
    extern __shared__ float sm[];
    const int Din = a.m.ent_dim, D = a.m.rel_dim, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sin = Din + 1, s2 = (a.m.model == OKB_TRANSR ? D : Din) + 1;
    const bool second = a.m.model == OKB_TRANSD || a.m.model == OKB_TRANSR;
    const int wpb = blockDim.x >> 5;
    const size_t per_warp = (size_t)2 * 32 * sin + (second ? (size_t)2 * 32 * s2 : 0);
    float *H = sm + (size_t)w * per_warp, *T = H + 32 * sin, *H2 = T + 32 * sin, *T2 = H2 + 32 * s2;
    const i32 qbase = (blockIdx.x * wpb + w) * 32;
    for (int rr = 0; rr < 32; rr++) {
        const i32 q = qbase + rr;
        if (q >= a.nq) break;
        const i64 h = a.th[a.q_lo + q], t = a.tt[a.q_lo + q];
        for (int d = lane; d < Din; d += 32) {
            H[rr * sin + d] = a.m.ent[h * Din + d];
            T[rr * sin + d] = a.m.ent[t * Din + d];
            if (a.m.model == OKB_TRANSD) { H2[rr * s2 + d] = a.m.ent_aux[h * Din + d]; T2[rr * s2 + d] = a.m.ent_aux[t * Din + d]; }
        }
    }
    __syncwarp();
    const i32 q = qbase + lane;
    if (q >= a.nq) return;
    const i32 g = a.q_group[q];
    const float *rv = a.rv + (i64)g * 2 * D;
    const float *M = a.m.model == OKB_TRANSR ? a.m.rel_aux + (i64)a.grp_rel[g] * Din * D : nullptr;
    float *h = a.m.model == OKB_TRANSR ? H2 + lane * s2 : H + lane * sin;
    float *t = a.m.model == OKB_TRANSR ? T2 + lane * s2 : T + lane * sin;
    canon_row(a.m.model, H + lane * sin, h, rv + D, H2 + lane * s2, M, Din, D);
    canon_row(a.m.model, T + lane * sin, t, rv + D, T2 + lane * s2, M, Din, D);
    const i32 ql = q - a.g_qlo[g];
    float *oa = a.qa + ((i64)(a.g_blk[g] + (ql >> 3)) * D) * 8 + (ql & 7), *ot = a.qt + (oa - a.qa);
    float s = 0.f;
    for (int d = 0; d < D; d++) {
        const float x = __fadd_rn(h[d], rv[d]);
        oa[d * 8] = x;
        ot[d * 8] = t[d];
        s = __fadd_rn(s, fabsf(__fsub_rn(x, t[d])));
    }
    const float ref = c_finish(a.m.model, s, D);
    a.ref[q] = ref;
    // The reference compares FINISHED scores (TransE: sum / D).  fl(x / D) is monotone in x, so
    // "fl(x / D) < ref" is "x < T" with T = min{x : fl(x / D) >= ref}: the ranking kernel compares raw sums
    // against T and divides only for the candidates that count.
    float thr = ref;
    if (a.m.model == OKB_TRANSE && ref > 0.f) {
        thr = __fmul_rn(ref, (float)D);
        while (thr > 0.f && __fdiv_rn(thr, (float)D) >= ref) thr = __uint_as_float(__float_as_uint(thr) - 1u);
        while (__fdiv_rn(thr, (float)D) < ref) thr = __uint_as_float(__float_as_uint(thr) + 1u);
    }
    a.thr[q] = thr;
}

// type flags per (group, candidate): bit 0 = in head_type[r], bit 1 = in tail_type[r]
__global__ void typeflag_kernel(const i32 *__restrict__ grp_rel, const i32 *__restrict__ hl, const i32 *__restrict__ hr,
                                const i32 *__restrict__ hid, const i32 *__restrict__ tl, const i32 *__restrict__ tr_,
                                const i32 *__restrict__ tid, unsigned char *__restrict__ flags, i32 j0, i32 ncol) {
    const i32 g = blockIdx.y, r = grp_rel[g];
    unsigned char *f = flags + (i64)g * ncol;
    for (i32 i = hl[r] + blockIdx.x * blockDim.x + threadIdx.x; i < hr[r]; i += gridDim.x * blockDim.x) {
        const i32 c = hid[i] - j0;
        if (c >= 0 && c < ncol) atomicOr((unsigned int *)(f + (c & ~3)), 1u << (8 * (c & 3)));
    }
    for (i32 i = tl[r] + blockIdx.x * blockDim.x + threadIdx.x; i < tr_[r]; i += gridDim.x * blockDim.x) {
        const i32 c = tid[i] - j0;
        if (c >= 0 && c < ncol) atomicOr((unsigned int *)(f + (c & ~3)), 2u << (8 * (c & 3)));
    }
}

// ------------------------------------------------------------------------------------------ rank: main kernel
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned phase) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}
// 1-D bulk async copy global -> shared (TMA engine; SASS UBLKCP), completion on an mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Packed fp32 pairs (sm_100 add/sub.rn.f32x2 -> SASS FADD2): two individually rounded adds per issue slot.  ptxas folds
// the |.| of the distance and the scalar broadcast of the candidate element into the instruction's source modifiers
// (FADD2 R, R.F32x2.HI_LO, |R|.F32x2.HI_LO / -R.F32), so the L1 distance loop costs ONE issue slot per (query, candidate,
// dim) instead of two — the loop is issue-bound (ncu: issue-active 69 %, FMA pipe 38 % with scalar FADDs).  Each half
// is rounded to nearest like __fadd_rn / __fsub_rn: scores stay bit-identical to the canonical order.
__device__ __forceinline__ float2 sub2_rn(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 abs2(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }

struct RankArgs {
    const float *cand;         // [ntab][D][ncol]
    const float *rv;           // [G][2][D]
    const float *qa, *qt;      // blocked query vectors [block][D][8]
    const float *thr;          // raw-sum thresholds per query of the chunk
    const i32 *g_blk;          // first block of each group
    const unsigned char *tflag;   // [G][ncol]
    const i32 *th, *tt;        // test heads / tails (global test index)
    const int4 *trun;          // known-list runs per test triple
    const i32 *known_t, *known_h;
    const i32 *g_qlo, *g_qhi;  // query range of each group, relative to the chunk
    unsigned long long *counts, *best;   // [(nq)*2*4], chunk-relative
    i32 D, ncol, j0, cand_lo, cand_hi, q_lo, model, tab_per_group;
};

// QG thread groups of CT threads share ONE staged candidate tile; group g owns queries g*QB .. g*QB+QB-1
// of every pass, so a CTA ranks QG*QB queries per pass against 128 candidates.  (One candidate column
// costs D*4 bytes of shared memory: sharing it between QG threads is what lifts residency from 12 to
// 24-32 warps per SM.)
template <bool HEADS, int QG>
__global__ void __launch_bounds__(CT * QG, 1024 / (CT * QG)) rank_kernel(RankArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int NQ = QG * QB;                            // queries per pass
    const int D = a.D, tid = threadIdx.x, jl = tid % CT, qg = tid / CT;
    float *tile = (float *)smraw;                          // [D][CT]
    float *rhat = tile + (size_t)D * CT;                   // [D]
    float *qbuf = rhat + ((D + 3) & ~3);                   // [2 buffers][qa | qt][QG][D][8]  (the global blocked layout)
    float *refs = qbuf + 4 * (size_t)D * NQ;               // [NQ] raw-sum thresholds
    i32 *tgt = (i32 *)(refs + NQ);                         // [NQ][2]
    int4 *runs = (int4 *)(tgt + 2 * NQ);                   // [NQ]
    int2 *krange = (int2 *)(runs + NQ);                    // [NQ][2] known-true ids that fall inside this CTA's candidate tile
    unsigned *cnt = (unsigned *)(krange + 2 * NQ);         // [NQ][2][4]
    unsigned long long *bst = (unsigned long long *)(cnt + NQ * 8);   // [NQ][2][4]
    unsigned long long *bar = bst + NQ * 8;

    const i32 g = blockIdx.y;
    const i32 col0 = blockIdx.x * CT;
    const i32 tab = a.tab_per_group ? g : 0;
    const i32 qlo = a.g_qlo[g], qhi = a.g_qhi[g];
    const i32 nblk = (qhi - qlo + QB - 1) / QB, blk0 = a.g_blk[g];
    // The query vectors of a pass are double-buffered: pass p+1's vectors are requested while pass p is computed, and
    // pass 0's go out together with the candidate tile, so a CTA exposes ONE bulk-copy latency instead of one per pass.
    auto load_queries = [&](i32 pb, int buf) {             // thread 0
        const i32 nb = min(QG, nblk - pb);
        const unsigned bytes = (unsigned)(nb * D * QB * sizeof(float));
        float *qa = qbuf + (size_t)buf * 2 * D * NQ, *qt = qa + (size_t)D * NQ;
        mbar_expect_tx(bar + 1 + buf, HEADS ? 2 * bytes : bytes);
        tma_load_1d(qa, a.qa + (i64)(blk0 + pb) * D * QB, bytes, bar + 1 + buf);
        if (HEADS) tma_load_1d(qt, a.qt + (i64)(blk0 + pb) * D * QB, bytes, bar + 1 + buf);
    };
    if (tid < 32) {                                        // warp 0: the D row copies of the tile go out from all 32 lanes at once
        if (tid == 0) {                                    // (one thread issuing them back to back costs ~1 us per CTA)
            mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1);
            mbar_expect_tx(bar, (unsigned)(D * CT * sizeof(float)));
        }
        __syncwarp();
        const float *src = a.cand + (i64)tab * D * a.ncol + col0;
        for (int d = tid; d < D; d += 32) tma_load_1d(tile + (size_t)d * CT, src + (i64)d * a.ncol, CT * sizeof(float), bar);
        if (tid == 0 && nblk > 0) load_queries(0, 0);
    }
    for (int d = tid; d < D; d += CT * QG) rhat[d] = a.rv[(i64)g * 2 * D + d];
    const i32 j = a.j0 + col0 + jl;
    const bool valid = j >= a.cand_lo && j < a.cand_hi;
    const unsigned tf = valid ? a.tflag[(i64)g * a.ncol + col0 + jl] : 0u;
    __syncthreads();                                       // barrier init visible to all waiters
    // (the tile is awaited right before the first distance loop: its copy latency runs under the first pass's set-up)

    int pass = 0;
    for (i32 pb = 0; pb < nblk; pb += QG, pass++) {        // a pass = QG consecutive 8-query blocks of the group
        const i32 nb = min(QG, nblk - pb), qb = qlo + pb * QB;
        const int buf = pass & 1;
        const float *qa = qbuf + (size_t)buf * 2 * D * NQ, *qt = qa + (size_t)D * NQ;
        __syncthreads();                                   // previous pass fully consumed (its buffer is free again)
        if (tid == 0 && pb + QG < nblk) load_queries(pb + QG, buf ^ 1);
        if (tid < NQ) {
            const bool ok = qb + tid < qhi;
            const i32 ti = a.q_lo + qb + tid;
            refs[tid] = ok ? a.thr[qb + tid] : 0.f;
            tgt[2 * tid] = ok ? a.th[ti] : -1;
            tgt[2 * tid + 1] = ok ? a.tt[ti] : -1;
            runs[tid] = ok ? a.trun[ti] : make_int4(0, 0, 0, 0);
        }
        if (tid < 2 * NQ) {
            // known-true filter (Corrupt.h:104-115), hoisted: the sorted known list of a (query, side) is cut down ONCE
            // per pass to the ids inside this tile [jlo, jlo + CT) — almost always 0 or 1 of them — so the per-candidate
            // test below is a warp-uniform scan of that sub-range instead of a divergent binary search per candidate
            const int ql = tid >> 1, side = tid & 1;
            int2 kr = make_int2(0, 0);
            if (qb + ql < qhi) {
                const int4 rn = a.trun[a.q_lo + qb + ql];
                const i32 *lst = side ? a.known_t : a.known_h;
                const i32 beg = side ? rn.x : rn.z, end = side ? rn.y : rn.w, jlo = a.j0 + col0;
                i32 lo = beg, hi = end;
                while (lo < hi) { const i32 mid = (lo + hi) >> 1; if (__ldg(lst + mid) < jlo) lo = mid + 1; else hi = mid; }
                kr.x = lo; hi = end;
                while (lo < hi) { const i32 mid = (lo + hi) >> 1; if (__ldg(lst + mid) < jlo + CT) lo = mid + 1; else hi = mid; }
                kr.y = lo;
            }
            krange[tid] = kr;
        }
        // argmins start from the best packed (score, id) any CTA has published so far: candidates that cannot beat it
        // never enter the warp reductions below
        for (int idx = tid; idx < NQ * 8; idx += CT * QG) {
            cnt[idx] = 0u;
            bst[idx] = (qb + (idx >> 3) < qhi) ? a.best[(i64)(qb + (idx >> 3)) * 8 + (idx & 7)] : ~0ull;
        }
        __syncthreads();
        if (pass == 0) mbar_wait(bar, 0);
        mbar_wait(bar + 1 + buf, (unsigned)(pass >> 1) & 1u);

        if (qg < nb) {
            const i32 q0 = qg * QB;                        // this thread group's queries within the pass
            const float *qas = qa + (size_t)qg * D * QB, *qts = qt + (size_t)qg * D * QB;
            float2 accT2[QB / 2], accH2[QB / 2];
#pragma unroll
            for (int q = 0; q < QB / 2; q++) { accT2[q] = make_float2(0.f, 0.f); accH2[q] = make_float2(0.f, 0.f); }
#pragma unroll 4
            for (int d = 0; d < D; d++) {
                const float c = tile[d * CT + jl];
                const float cr = __fadd_rn(c, rhat[d]);
                const float2 c2 = make_float2(c, c), cr2 = make_float2(cr, cr);
                const float4 *pa = (const float4 *)(qas + d * QB), *pt = (const float4 *)(qts + d * QB);
#pragma unroll
                for (int v = 0; v < QB / 4; v++) {
                    const float4 x = pa[v];
                    accT2[2 * v + 0] = add2_rn(accT2[2 * v + 0], abs2(sub2_rn(make_float2(x.x, x.y), c2)));
                    accT2[2 * v + 1] = add2_rn(accT2[2 * v + 1], abs2(sub2_rn(make_float2(x.z, x.w), c2)));
                    if (HEADS) {
                        const float4 y = pt[v];
                        accH2[2 * v + 0] = add2_rn(accH2[2 * v + 0], abs2(sub2_rn(cr2, make_float2(y.x, y.y))));
                        accH2[2 * v + 1] = add2_rn(accH2[2 * v + 1], abs2(sub2_rn(cr2, make_float2(y.z, y.w))));
                    }
                }
            }
            float accT[QB], accH[QB];
#pragma unroll
            for (int q = 0; q < QB / 2; q++) {
                accT[2 * q] = accT2[q].x; accT[2 * q + 1] = accT2[q].y;
                accH[2 * q] = accH2[q].x; accH[2 * q + 1] = accH2[q].y;
            }
            const int lane = tid & 31;
#pragma unroll
            for (int q = 0; q < QB; q++) {
                const int ql = q0 + q;
                const bool qok = valid && qb + ql < qhi;
                const float thr = refs[ql];
#pragma unroll
                for (int side = HEADS ? 0 : 1; side < 2; side++) {
                    const float raw = side ? accT[q] : accH[q];
                    const bool better = qok && j != tgt[2 * ql + side] && raw < thr;       // Test.h:59,62 / 168,171
                    const unsigned mb = __ballot_sync(FULL, better);
                    if (mb == 0u) continue;                                              // warp-uniform
                    const int2 kr = krange[ql * 2 + side];
                    const i32 *lst = side ? a.known_t : a.known_h;
                    bool known = false;
                    for (i32 i = kr.x; i < kr.y; i++) known |= __ldg(lst + i) == j;        // (j, t, r) / (h, j, r) in train+valid+test?
                    const bool typed = (tf >> side) & 1u;
                    const unsigned m1 = __ballot_sync(FULL, better && !known), m2 = __ballot_sync(FULL, better && typed),
                                   m3 = __ballot_sync(FULL, better && typed && !known);
                    unsigned *cq = cnt + (ql * 2 + side) * 4;
                    unsigned long long *bq = bst + (ql * 2 + side) * 4;
                    if (lane < 4) {
                        const unsigned mm = lane == 0 ? mb : (lane == 1 ? m1 : (lane == 2 ? m2 : m3));
                        if (mm) atomicAdd(cq + lane, (unsigned)__popc(mm));
                    }
                    // argmins: packed (finished score bits, id); scores are non-negative, so their bit patterns order like uints
                    const float sfin = c_finish(a.model, raw, D);
                    const unsigned long long pk = ((unsigned long long)__float_as_uint(sfin) << 32) | (unsigned)j;
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        const bool in = v == 0 ? better : (v == 1 ? (better && !known) : (v == 2 ? (better && typed) : (better && typed && !known)));
                        const bool cand = in && pk < bq[v];
                        if (__ballot_sync(FULL, cand) == 0u) continue;
                        const unsigned sb = cand ? __float_as_uint(sfin) : 0xffffffffu;
                        const unsigned smin = __reduce_min_sync(FULL, sb);
                        const unsigned idmin = __reduce_min_sync(FULL, (cand && sb == smin) ? (unsigned)j : 0xffffffffu);
                        if (lane == 0) atomicMin(bq + v, ((unsigned long long)smin << 32) | idmin);
                    }
                }
            }
        }
        __syncthreads();
        for (int idx = tid; idx < NQ * 8; idx += CT * QG) {
            const int q = idx >> 3;
            if (qb + q < qhi) {
                const i64 o = (i64)(qb + q) * 8 + (idx & 7);
                if (cnt[idx]) atomicAdd(a.counts + o, (unsigned long long)cnt[idx]);
                if (bst[idx] != ~0ull) atomicMin(a.best + o, bst[idx]);       // seeded from a.best: a no-op unless improved
            }
        }
    }
    if (pass == 0) mbar_wait(bar, 0);                       // never leave with the tile copy still in flight
}

// ------------------------------------------------------------------------------------------ finalize
struct FinArgs {
    const unsigned long long *counts, *best;
    const i32 *th, *tt;
    const i32 *sup_l, *sup_r, *sup_id, *sub_l, *sub_r, *sub_id;
    i64 *out;
    i32 q_lo, nq, has_onto;
};
__device__ __forceinline__ void onto_classes(const FinArgs &a, i32 target, const i64 *arg, i64 *out4) {
    // Test.h:109-132: the two cursors are shared by the four lookups and never rewound
    i32 p = a.has_onto ? a.sup_l[target] : 0, pe = a.has_onto ? a.sup_r[target] : 0;
    i32 s = a.has_onto ? a.sub_l[target] : 0, se = a.has_onto ? a.sub_r[target] : 0;
    for (int k = 0; k < 4; k++) {
        const i64 id = arg[k];
        if (id == target) { out4[k] = 0; continue; }
        while (p < pe && a.sup_id[p] < id) p++;
        if (p < pe && a.sup_id[p] == id) { out4[k] = 1; continue; }
        while (s < se && a.sub_id[s] < id) s++;
        if (s < se && a.sub_id[s] == id) { out4[k] = 2; continue; }
        out4[k] = 3;
    }
}
__global__ void finalize_kernel(FinArgs a) {
    const i32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nq * 2) return;
    const i32 q = i >> 1, side = i & 1;
    const i32 target = side ? a.tt[a.q_lo + q] : a.th[a.q_lo + q];
    i64 arg[4];
    i64 *o = a.out + (i64)i * 8;
    for (int k = 0; k < 4; k++) {
        o[k] = (i64)a.counts[(i64)i * 4 + k];
        const unsigned long long b = a.best[(i64)i * 4 + k];
        arg[k] = b == ~0ull ? target : (i64)(b & 0xffffffffull);
    }
    if (side == 0) arg[1] = arg[0];     // Test.h:69-74: head-side filter argmin is not guarded by the filter
    onto_classes(a, target, arg, o + 4);
}

// ------------------------------------------------------------------------------------------ one query, given scores
struct RankScoresArgs {
    const float *scores;
    const i32 *known, *types;
    i32 E, target, klo, khi, tlo, thi;
    unsigned long long *counts, *best;   // [4] each, pre-initialised
};
__global__ void __launch_bounds__(256) rank_scores_kernel(RankScoresArgs a) {
    __shared__ unsigned cnt[4];
    __shared__ unsigned long long bst[4];
    if (threadIdx.x < 4) { cnt[threadIdx.x] = 0; bst[threadIdx.x] = ~0ull; }
    __syncthreads();
    const float ref = a.scores[a.target];
    for (i32 j = blockIdx.x * blockDim.x + threadIdx.x; j < a.E; j += gridDim.x * blockDim.x) {
        const float s = a.scores[j];
        if (j == a.target || !(s < ref)) continue;
        i32 lo = a.klo, hi = a.khi;
        while (lo < hi) { const i32 mid = (lo + hi) >> 1; if (a.known[mid] < j) lo = mid + 1; else hi = mid; }
        const bool known = lo < a.khi && a.known[lo] == j;
        lo = a.tlo; hi = a.thi;
        while (lo < hi) { const i32 mid = (lo + hi) >> 1; if (a.types[mid] < j) lo = mid + 1; else hi = mid; }
        const bool typed = lo < a.thi && a.types[lo] == j;
        // scores may be negative here (caller-provided): order-preserving float -> uint key
        unsigned u = __float_as_uint(s);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        const unsigned long long pk = ((unsigned long long)u << 32) | (unsigned)j;
        atomicAdd(cnt + 0, 1u); atomicMin(bst + 0, pk);
        if (!known) { atomicAdd(cnt + 1, 1u); atomicMin(bst + 1, pk); }
        if (typed) {
            atomicAdd(cnt + 2, 1u); atomicMin(bst + 2, pk);
            if (!known) { atomicAdd(cnt + 3, 1u); atomicMin(bst + 3, pk); }
        }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        if (cnt[threadIdx.x]) atomicAdd(a.counts + threadIdx.x, (unsigned long long)cnt[threadIdx.x]);
        if (bst[threadIdx.x] != ~0ull) atomicMin(a.best + threadIdx.x, bst[threadIdx.x]);
    }
}

__global__ void fill_u64_kernel(unsigned long long *p, unsigned long long v, i64 n) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------ host side
extern bool okb_pick_layout(int D, int &vw, int &nv);
int okb_transr_tc_supported(const okb_model *m);
int okb_transr_project_tc(okb_ctx *c, const okb_model *m, const i32 *d_grp_rel, i64 G, float *out, i64 j0, i64 ncol, cudaStream_t s);

static int check_score_model(okb_ctx *c, const okb_model *m) {
    if (!m || !m->ent || !m->rel) OKB_FAIL(c, OKB_ERR_ARG, "model tables missing");
    if (m->model != OKB_TRANSR && m->ent_dim != m->rel_dim) OKB_FAIL(c, OKB_ERR_ARG, "TransE/H/D need ent_dim == rel_dim");
    if (m->model == OKB_TRANSR && m->rel_dim > TRANSR_MAXD) OKB_FAIL(c, OKB_ERR_ARG, "TransR rel_dim > 256 not supported");
    if (m->model != OKB_TRANSE && !m->rel_aux) OKB_FAIL(c, OKB_ERR_ARG, "rel_aux table missing");
    if (m->model == OKB_TRANSD && !m->ent_aux) OKB_FAIL(c, OKB_ERR_ARG, "ent_aux table missing");
    return 0;
}

extern "C" {

int okb_predict(okb_ctx *c, const okb_model *m, const int64_t *h, const int64_t *t, const int64_t *r, INT n, float *out,
                void *stream) {
    int rc = check_score_model(c, m);
    if (rc) return rc;
    if (n <= 0) return 0;
    PredArgs a;
    a.m = *m; a.h = h; a.t = t; a.r = r; a.out = out; a.n = n; a.E = (i32)c->E; a.R = (i32)c->R;
    // TransE / TransH / TransD with rows that fit the staging buffers: coalesced staged form; else one thread per triple
    const int narr = m->model == OKB_TRANSE ? 3 : (m->model == OKB_TRANSH ? 4 : 6);
    const size_t smem = sizeof(float) * 2 * (size_t)narr * 32 * (m->ent_dim + 1);
    if (m->model != OKB_TRANSR && smem <= 200 * 1024 && n >= 64) {
        OKB_CUDA(c, okb_smem_optin(c, predict_stage_kernel, smem));
        predict_stage_kernel<<<(unsigned)((n + 63) / 64), 64, smem, (cudaStream_t)stream>>>(a, narr);
    } else
        predict_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

int okb_rank(okb_ctx *c, const okb_model *m, INT q_lo, INT q_hi, int heads, INT cand_lo, INT cand_hi, int64_t *counts,
             uint64_t *best, void *stream) {
    int rc = check_score_model(c, m);
    if (rc) return rc;
    if (!c->d_test_h || !c->have_types) OKB_FAIL(c, OKB_ERR_STATE, "import test and type files first");
    if (q_lo < 0 || q_hi > c->n_test || q_lo > q_hi) OKB_FAIL(c, OKB_ERR_ARG, "bad query range");
    if (cand_lo < 0 || cand_hi > c->E || cand_lo >= cand_hi) OKB_FAIL(c, OKB_ERR_ARG, "bad candidate range");
    if (q_lo == q_hi) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int D = m->rel_dim, Din = m->ent_dim;              // D: dimension the L1 distance runs over
    const i64 j0 = (cand_lo / CT) * CT;
    const i64 ncol = ((cand_hi - j0 + CT - 1) / CT) * CT;
    const bool per_group = m->model != OKB_TRANSE;

    // relation groups intersecting [q_lo, q_hi)
    std::vector<i32> grel, gqlo, gqhi;
    for (size_t g = 0; g < c->grp_rel.size(); g++) {
        const i64 lo = std::max<i64>(c->grp_lo[g], q_lo), hi = std::min<i64>(c->grp_hi[g], q_hi);
        if (lo < hi) { grel.push_back(c->grp_rel[g]); gqlo.push_back((i32)lo); gqhi.push_back((i32)hi); }
    }
    // chunk the groups so that candidate tables stay within a memory budget
    const size_t tab_bytes = sizeof(float) * D * ncol;
    const size_t budget = (size_t)4 << 30;
    i64 gmax = per_group ? std::max<i64>(1, (i64)(budget / tab_bytes)) : (i64)grel.size();
    gmax = std::min<i64>(gmax, 8192);
    // thread groups per CTA: more groups = more warps per staged tile; 4 when groups are large enough to fill 32 queries
    const i64 avg_q = grel.empty() ? 1 : (q_hi - q_lo) / (i64)grel.size();
    auto rank_smem = [&](int qg) {
        const size_t nq = (size_t)qg * QB;
        return sizeof(float) * ((size_t)D * CT + ((D + 3) & ~3) + 4 * (size_t)D * nq + nq) + sizeof(i32) * 2 * nq +
               sizeof(int4) * nq + sizeof(int2) * 2 * nq + sizeof(unsigned) * nq * 8 + sizeof(unsigned long long) * (nq * 8 + 3) + 128;
    };
    // Thread groups per CTA: 4 when groups are large enough to fill 32 query slots per pass.  (Picking the QG in {2,3,4}
    // that wastes the fewest 8-query block slots was measured and is within +-10 % of this rule either way: fewer idle
    // thread groups but fewer warps sharing a staged tile.)
    int QGsel = avg_q >= 20 ? 4 : 2;
    if (QGsel == 4 && rank_smem(4) > 227 * 1024) QGsel = 2;          // wide embeddings: fewer query slots per pass
    if (const char *e = getenv("OKB200_RANK_QG")) { const int v = atoi(e); if (v >= 2 && v <= 4 && rank_smem(v) <= 227 * 1024) QGsel = v; }   // A/B runs
    const size_t NQ = (size_t)QGsel * QB;
    const size_t smem_rank = rank_smem(QGsel);
    if (smem_rank > 227 * 1024) OKB_FAIL(c, OKB_ERR_ARG, "embedding dimension too large for the ranking tile");
    {   // per context (= per device), not per process
        const size_t lim = 227 * 1024;
        OKB_CUDA(c, okb_smem_optin(c, rank_kernel<true, 2>, lim)); OKB_CUDA(c, okb_smem_optin(c, rank_kernel<false, 2>, lim));
        OKB_CUDA(c, okb_smem_optin(c, rank_kernel<true, 3>, lim)); OKB_CUDA(c, okb_smem_optin(c, rank_kernel<false, 3>, lim));
        OKB_CUDA(c, okb_smem_optin(c, rank_kernel<true, 4>, lim)); OKB_CUDA(c, okb_smem_optin(c, rank_kernel<false, 4>, lim));
        OKB_CUDA(c, okb_smem_optin(c, cand_kernel, lim)); OKB_CUDA(c, okb_smem_optin(c, qvec_kernel, lim));
    }
    const bool second = m->model == OKB_TRANSD || m->model == OKB_TRANSR;
    const size_t s2 = (m->model == OKB_TRANSR ? D : Din) + 1;
    const size_t warp_q = 2 * ((size_t)32 * (Din + 1) + (second ? 32 * s2 : 0));
    // candidate kernel: TransD's second staged array (ent_transfer rows) is replaced by one precomputed dot per candidate
    const size_t warp_c = (size_t)32 * (Din + 1) + (m->model == OKB_TRANSR ? 32 * s2 : 0);
    const size_t blk_c = ((D + 3) & ~3) + (m->model == OKB_TRANSR ? (size_t)Din * D : 0);
    int wpb_c = 4, wpb_q = 4;
    while (wpb_c > 1 && sizeof(float) * (wpb_c * warp_c + blk_c) > 160 * 1024) wpb_c >>= 1;
    while (wpb_q > 1 && sizeof(float) * wpb_q * warp_q > 100 * 1024) wpb_q >>= 1;
    const size_t smem_cand = sizeof(float) * (wpb_c * warp_c + blk_c);
    const size_t smem_q = sizeof(float) * wpb_q * warp_q;
    if (smem_q > 227 * 1024 || smem_cand > 227 * 1024) OKB_FAIL(c, OKB_ERR_ARG, "embedding dimension too large for the ranking prep kernels");

    for (size_t g0 = 0; g0 < grel.size(); g0 += gmax) {
        const i64 G = std::min<i64>(gmax, (i64)grel.size() - g0);
        const i64 cq_lo = gqlo[g0], cq_hi = gqhi[g0 + G - 1], nq = cq_hi - cq_lo;
        const i64 ntab = per_group ? G : 1;
        // workspace layout
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        std::vector<i32> gblk(G);
        i64 NBq = 0;                                       // 8-query blocks, per group padded
        for (i64 g = 0; g < G; g++) { gblk[g] = (i32)NBq; NBq += (gqhi[g0 + g] - gqlo[g0 + g] + QB - 1) / QB; }
        const size_t o_cand = take(tab_bytes * ntab), o_rv = take(sizeof(float) * G * 2 * D), o_qa = take(sizeof(float) * NBq * QB * D),
                     o_qt = take(sizeof(float) * NBq * QB * D), o_ref = take(sizeof(float) * nq), o_thr = take(sizeof(float) * nq),
                     o_ed = take(sizeof(float) * ncol), o_tf = take((size_t)G * ncol), o_grel = take(sizeof(i32) * G), o_gqlo = take(sizeof(i32) * G),
                     o_gqhi = take(sizeof(i32) * G), o_gblk = take(sizeof(i32) * G), o_qg = take(sizeof(i32) * nq);
        if (c->rank_ws.ensure(off)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (ranking workspace)");
        char *ws = c->rank_ws.as<char>();
        std::vector<i32> rel_lo(G), rel_hi(G), qg(nq);
        for (i64 g = 0; g < G; g++) {
            rel_lo[g] = (i32)(gqlo[g0 + g] - cq_lo); rel_hi[g] = (i32)(gqhi[g0 + g] - cq_lo);
            for (i32 q = rel_lo[g]; q < rel_hi[g]; q++) qg[q] = (i32)g;
        }
        OKB_CUDA(c, cudaMemcpyAsync(ws + o_grel, grel.data() + g0, sizeof(i32) * G, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaMemcpyAsync(ws + o_gqlo, rel_lo.data(), sizeof(i32) * G, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaMemcpyAsync(ws + o_gqhi, rel_hi.data(), sizeof(i32) * G, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaMemcpyAsync(ws + o_qg, qg.data(), sizeof(i32) * nq, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaMemcpyAsync(ws + o_gblk, gblk.data(), sizeof(i32) * G, cudaMemcpyHostToDevice, s));
        OKB_CUDA(c, cudaStreamSynchronize(s));            // host vectors above go out of scope
        OKB_CUDA(c, cudaMemsetAsync(ws + o_tf, 0, (size_t)G * ncol, s));
        OKB_CUDA(c, cudaMemsetAsync(ws + o_qa, 0, sizeof(float) * NBq * QB * D, s));      // padded query slots stay finite
        OKB_CUDA(c, cudaMemsetAsync(ws + o_qt, 0, sizeof(float) * NBq * QB * D, s));

        prof_mark(c, PROF_RANK_PREP, s);
        relvec_kernel<<<(unsigned)((G + 63) / 64), 64, 0, s>>>(*m, (const i32 *)(ws + o_grel), (float *)(ws + o_rv), (i32)G);
        CandArgs ca;
        ca.m = *m; ca.rv = (const float *)(ws + o_rv); ca.out = (float *)(ws + o_cand); ca.grp_rel = (const i32 *)(ws + o_grel);
        ca.j0 = (i32)j0; ca.ncol = (i32)ncol; ca.E = (i32)c->E; ca.ntab = (i32)ntab;
        ca.entdot = nullptr;
        if (m->model == OKB_TRANSD) {
            entdot_kernel<<<(unsigned)((ncol + 127) / 128), 128, 0, s>>>(*m, (float *)(ws + o_ed), (i32)j0, (i32)ncol, (i32)c->E);
            OKB_LAUNCHED(1);
            ca.entdot = (const float *)(ws + o_ed);
        }
        if (m->model == OKB_TRANSR && c->transr_tc && okb_transr_tc_supported(m)) {
            int rc2 = okb_transr_project_tc(c, m, (const i32 *)(ws + o_grel), G, (float *)(ws + o_cand), j0, ncol, s);
            if (rc2) return rc2;
        } else
            cand_kernel<<<dim3((unsigned)(ncol / (32 * wpb_c)), (unsigned)ntab), 32 * wpb_c, smem_cand, s>>>(ca);
        QArgs qa;
        qa.m = *m; qa.th = c->d_test_h; qa.tt = c->d_test_t; qa.tr = c->d_test_r; qa.q_group = (const i32 *)(ws + o_qg);
        qa.grp_rel = (const i32 *)(ws + o_grel);
        qa.rv = (const float *)(ws + o_rv); qa.qa = (float *)(ws + o_qa); qa.qt = (float *)(ws + o_qt); qa.ref = (float *)(ws + o_ref);
        qa.thr = (float *)(ws + o_thr); qa.g_qlo = (const i32 *)(ws + o_gqlo); qa.g_blk = (const i32 *)(ws + o_gblk);
        qa.q_lo = (i32)cq_lo; qa.nq = (i32)nq;
        qvec_kernel<<<(unsigned)((nq + 32 * wpb_q - 1) / (32 * wpb_q)), 32 * wpb_q, smem_q, s>>>(qa);
        typeflag_kernel<<<dim3(8, (unsigned)G), 128, 0, s>>>((const i32 *)(ws + o_grel), c->head_type.d_lef, c->head_type.d_rig,
                                                            c->head_type.d_ids, c->tail_type.d_lef, c->tail_type.d_rig,
                                                            c->tail_type.d_ids, (unsigned char *)(ws + o_tf), (i32)j0, (i32)ncol);
        prof_mark(c, PROF_RANK_PREP, s);
        RankArgs ra;
        ra.cand = (const float *)(ws + o_cand); ra.rv = (const float *)(ws + o_rv);
        ra.qa = (const float *)(ws + o_qa); ra.qt = (const float *)(ws + o_qt); ra.thr = (const float *)(ws + o_thr);
        ra.g_blk = (const i32 *)(ws + o_gblk);
        ra.tflag = (const unsigned char *)(ws + o_tf);
        ra.th = c->d_test_h; ra.tt = c->d_test_t; ra.trun = c->d_test_run; ra.known_t = c->d_known_t; ra.known_h = c->d_known_h;
        ra.g_qlo = (const i32 *)(ws + o_gqlo); ra.g_qhi = (const i32 *)(ws + o_gqhi);
        ra.counts = (unsigned long long *)counts + (cq_lo - q_lo) * 8;
        ra.best = (unsigned long long *)best + (cq_lo - q_lo) * 8;
        ra.D = D; ra.ncol = (i32)ncol; ra.j0 = (i32)j0; ra.cand_lo = (i32)cand_lo; ra.cand_hi = (i32)cand_hi;
        ra.q_lo = (i32)cq_lo; ra.model = m->model; ra.tab_per_group = per_group ? 1 : 0;
        const dim3 grid((unsigned)(ncol / CT), (unsigned)G);
        { ProfScope ps(c, PROF_RANK, s);
        if (QGsel == 4) { if (heads) rank_kernel<true, 4><<<grid, CT * 4, smem_rank, s>>>(ra); else rank_kernel<false, 4><<<grid, CT * 4, smem_rank, s>>>(ra); }
        else if (QGsel == 3) { if (heads) rank_kernel<true, 3><<<grid, CT * 3, smem_rank, s>>>(ra); else rank_kernel<false, 3><<<grid, CT * 3, smem_rank, s>>>(ra); }
        else { if (heads) rank_kernel<true, 2><<<grid, CT * 2, smem_rank, s>>>(ra); else rank_kernel<false, 2><<<grid, CT * 2, smem_rank, s>>>(ra); } }
        OKB_LAUNCHED(5);
        OKB_CUDA(c, cudaGetLastError());
    }
    return 0;
}

int okb_rank_finalize(okb_ctx *c, INT q_lo, INT q_hi, const int64_t *counts, const uint64_t *best, int64_t *out, void *stream) {
    if (q_lo < 0 || q_hi > c->n_test || q_lo > q_hi) OKB_FAIL(c, OKB_ERR_ARG, "bad query range");
    if (q_lo == q_hi) return 0;
    FinArgs a;
    a.counts = (const unsigned long long *)counts; a.best = (const unsigned long long *)best;
    a.th = c->d_test_h; a.tt = c->d_test_t;
    a.has_onto = c->have_onto ? 1 : 0;
    a.sup_l = c->sup.d_lef; a.sup_r = c->sup.d_rig; a.sup_id = c->sup.d_ids;
    a.sub_l = c->sub.d_lef; a.sub_r = c->sub.d_rig; a.sub_id = c->sub.d_ids;
    a.out = out; a.q_lo = (i32)q_lo; a.nq = (i32)(q_hi - q_lo);
    finalize_kernel<<<(unsigned)((a.nq * 2 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}

int okb_rank_scores(okb_ctx *c, INT index, int side, const float *scores, int64_t *out, void *stream) {
    if (!c->d_test_h || !c->have_types) OKB_FAIL(c, OKB_ERR_STATE, "import test and type files first");
    if (index < 0 || index >= c->n_test) OKB_FAIL(c, OKB_ERR_ARG, "test index out of range");
    cudaStream_t s = (cudaStream_t)stream;
    if (c->rank_ws.ensure(1024)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory");
    unsigned long long *cnt = c->rank_ws.as<unsigned long long>(), *bst = cnt + 8;
    OKB_CUDA(c, cudaMemsetAsync(cnt, 0, 64, s));
    OKB_CUDA(c, cudaMemsetAsync(bst, 0xff, 64, s));
    // this entry point serves ONE (index, side); lay its 4 counters where finalize expects side `side`
    RankScoresArgs a;
    a.scores = scores; a.E = (i32)c->E;
    const i32 h = c->test_h[index], t = c->test_t[index], r = c->test_r[index];
    a.target = side ? t : h;
    Lists &L = side ? c->tail_type : c->head_type;
    a.types = L.d_ids; a.tlo = L.lef[r]; a.thi = L.rig[r];
    int4 run;
    OKB_CUDA(c, cudaMemcpy(&run, c->d_test_run + index, sizeof(int4), cudaMemcpyDeviceToHost));
    a.known = side ? c->d_known_t : c->d_known_h;
    a.klo = side ? run.x : run.z; a.khi = side ? run.y : run.w;
    a.counts = cnt + (side ? 4 : 0); a.best = bst + (side ? 4 : 0);
    rank_scores_kernel<<<okb_sms(c), 256, 0, s>>>(a);
    OKB_LAUNCHED(1);
    if (c->host_io.ensure(sizeof(i64) * 16)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory");
    int rc = okb_rank_finalize(c, index, index + 1, (const int64_t *)cnt, (const uint64_t *)bst, c->host_io.as<i64>(), stream);
    if (rc) return rc;
    OKB_CUDA(c, cudaMemcpyAsync(out, c->host_io.as<i64>() + (side ? 8 : 0), sizeof(i64) * 8, cudaMemcpyDeviceToDevice, s));
    return 0;
}

}  // extern "C"
