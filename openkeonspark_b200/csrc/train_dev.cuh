// Device-side building blocks of the train step, shared by train.cu (one kernel per phase) and chunk.cu (the persistent
// kernel that runs grad -> grid barrier -> update for a whole chunk of planned steps): row fragments, the four models'
// forward / backward pieces, the grad kernels' bodies, the segmented gradient sum and the SGD / TF1-Adam update bodies.
// Everything here is a template or an inline device function; see train.cu for the description of the phases.
#pragma once
#include "okb_internal.h"

#define FULL 0xffffffffu
// flag block of a data-parallel peer arena (see "data parallel, owner-sharded" below), in u64 units
#define DP_FLAG_STAGE 0        // stage_ready[16]: rank q has finished pushing its partial rows of epoch e
#define DP_FLAG_X 16           // x_ready[16]: rank q has finished publishing its updated rows of epoch e
#define DP_FLAG_LOSS 32        // float loss_part[16] (byte 256)
#define DP_FLAG_BYTES 512
// Programmatic dependent launch (PTX griddepcontrol): a kernel launched with the stream-serialization attribute may
// start while its predecessor drains; `wait` blocks until the predecessor grid has completed and flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Spin until a peer's flag word reaches `epoch` (acquire, system scope).  A peer that died would otherwise hang this GPU
// for good: after ~20 s without progress the kernel traps, which surfaces as a CUDA error in the host process.
// mode 1: relaxed polls and ONE acquire fence once the value has arrived (instead of an acquire per poll)
__device__ __forceinline__ void spin_until(const unsigned long long *flag, unsigned long long epoch, int mode = 0) {
    unsigned long long v, t0 = 0;
    unsigned polls = 0;
    for (;;) {
        if (mode) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        else asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= epoch) { if (mode) asm volatile("fence.acq_rel.sys;" ::: "memory"); return; }
        if ((++polls & 0x3ffu) == 0u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) __trap();
        }
    }
}
// OKB_FLAG_DP_TRACE: phase stamps of the data-parallel kernels (tr = nullptr: off).  "max" slots keep the complement.
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void trace_min(unsigned long long *tr, int slot) { if (tr) atomicMin(tr + slot, gtimer()); }
__device__ __forceinline__ void trace_max(unsigned long long *tr, int slot) { if (tr) atomicMin(tr + slot, ~gtimer()); }
#define WARPS_PER_BLOCK 4
#define PCH 8                  // long segments (> PCH rows) are pre-reduced in fixed chunks of PCH sorted positions
// grad kernel: one warp per block so that residency is quantised in single warps: <= 120 registers
// -> 17 warps per SM -> B = 4831 positives fit in 2 waves instead of 2.04 (= 3)
#ifndef GRAD_WARPS
#define GRAD_WARPS 1
#endif
#ifndef GRAD_MIN_BLOCKS
#define GRAD_MIN_BLOCKS 17
#endif


// ------------------------------------------------------------------------------------------ row fragments
// A row of D floats is spread over a warp: lane l holds NV vectors of VW floats, vector i covering
// elements [(i*32 + l)*VW, +VW).  VW = 4 gives the 128-bit gathers; D % VW == 0 is required.
template <int VW> struct VecT;
template <> struct VecT<4> { typedef float4 T; };
template <> struct VecT<2> { typedef float2 T; };
template <> struct VecT<1> { typedef float T; };

template <int VW, int NV> struct Frag {
    float v[VW * NV];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < VW * NV; i++) v[i] = 0.f;
    }
    // COH = true: ld.global.cg (L2 only) instead of the non-coherent read-only path — for data that OTHER SMs write
    // during the same kernel (the persistent chunk kernel: tables, gradient rows); never a stale L1 line
    template <bool COH = false> __device__ __forceinline__ void load(const float *__restrict__ row, int D, int lane) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const int e = (i * 32 + lane) * VW;
            if (e < D) {
                typename VecT<VW>::T x = COH ? __ldcg(reinterpret_cast<const typename VecT<VW>::T *>(row + e))
                                             : __ldg(reinterpret_cast<const typename VecT<VW>::T *>(row + e));
                const float *xs = reinterpret_cast<const float *>(&x);
#pragma unroll
                for (int j = 0; j < VW; j++) v[i * VW + j] = xs[j];
            } else {
#pragma unroll
                for (int j = 0; j < VW; j++) v[i * VW + j] = 0.f;
            }
        }
    }
    __device__ __forceinline__ void store(float *__restrict__ row, int D, int lane) const {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const int e = (i * 32 + lane) * VW;
            if (e < D) {
                typename VecT<VW>::T x;
                float *xs = reinterpret_cast<float *>(&x);
#pragma unroll
                for (int j = 0; j < VW; j++) xs[j] = v[i * VW + j];
                *reinterpret_cast<typename VecT<VW>::T *>(row + e) = x;
            }
        }
    }
};

__device__ __forceinline__ float wsum(float x) {
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ void wsum2(float &a, float &b) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(FULL, a, o); b += __shfl_xor_sync(FULL, b, o); }
}
__device__ __forceinline__ void wsum3(float &a, float &b, float &c) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(FULL, a, o); b += __shfl_xor_sync(FULL, b, o); c += __shfl_xor_sync(FULL, c, o);
    }
}
#define FOR_N for (int i = 0; i < N; i++)

template <int N> __device__ __forceinline__ float dot(const float *a, const float *b) {
    float s = 0.f;
#pragma unroll
    FOR_N s = fmaf(a[i], b[i], s);
    return s;
}

// ------------------------------------------------------------------------------------------ model pieces
// Relation-side state of one triple.
template <int MODEL, int N> struct RelS {
    float rhat[N];      // l2n(rel_embeddings[r])
    float inv;          // rsqrt(max(|r|^2, 1e-12))
    bool proj;          // |r|^2 > 1e-12 (normalisation differentiates through the norm)
    float aux[N];       // TransH: n_hat = l2n(normal_vectors[r]);  TransD: rel_transfer[r]
    float inv_n;        // TransH only
    bool proj_n;
};
// Entity-side state of one (entity, relation) pair.
template <int MODEL, int N> struct EntS {
    float raw[N];       // ent_embeddings[e]
    float aux[N];       // TransD: ent_transfer[e]
    float hat[N];       // l2n(transfer(e))
    float inv, a;       // a = e.n_hat (TransH) / e.e_t (TransD)
    bool proj;
};

#define EPS_NORM 1e-12f

// Row gathers are split from the arithmetic so that a warp can put EVERY row of its positive group in flight
// (relation rows, head, tail, first negative) before the first shuffle reduction consumes one.
template <int MODEL, int VW, int NV, bool COH = false>
__device__ __forceinline__ void rel_load(RelS<MODEL, VW * NV> &R, const okb_model &m, i32 r, int lane) {
    constexpr int N = VW * NV;
    const int D = m.rel_dim;
    Frag<VW, NV> f;
    f.template load<COH>(m.rel + (i64)r * D, D, lane);
#pragma unroll
    FOR_N R.rhat[i] = f.v[i];
    if (MODEL != OKB_TRANSE) {
        f.template load<COH>(m.rel_aux + (i64)r * D, D, lane);
#pragma unroll
        FOR_N R.aux[i] = f.v[i];
    }
}
template <int MODEL, int N>
__device__ __forceinline__ void rel_math(RelS<MODEL, N> &R) {
    float ss = wsum(dot<N>(R.rhat, R.rhat));
    R.proj = ss > EPS_NORM;
    R.inv = rsqrtf(fmaxf(ss, EPS_NORM));
#pragma unroll
    FOR_N R.rhat[i] = R.rhat[i] * R.inv;
    if (MODEL == OKB_TRANSH) {
        float sn = wsum(dot<N>(R.aux, R.aux));
        R.proj_n = sn > EPS_NORM;
        R.inv_n = rsqrtf(fmaxf(sn, EPS_NORM));
#pragma unroll
        FOR_N R.aux[i] = R.aux[i] * R.inv_n;
    }
}
template <int MODEL, int VW, int NV, bool COH = false>
__device__ __forceinline__ void rel_forward(RelS<MODEL, VW * NV> &R, const okb_model &m, i32 r, int lane) {
    rel_load<MODEL, VW, NV, COH>(R, m, r, lane);
    rel_math<MODEL, VW * NV>(R);
}

template <int MODEL, int VW, int NV, bool COH = false>
__device__ __forceinline__ void ent_load(EntS<MODEL, VW * NV> &S, const okb_model &m, i32 e, int lane) {
    constexpr int N = VW * NV;
    const int D = m.ent_dim;
    Frag<VW, NV> f;
    f.template load<COH>(m.ent + (i64)e * D, D, lane);
#pragma unroll
    FOR_N S.raw[i] = f.v[i];
    if (MODEL == OKB_TRANSD) {
        f.template load<COH>(m.ent_aux + (i64)e * D, D, lane);
#pragma unroll
        FOR_N S.aux[i] = f.v[i];
    }
}
template <int MODEL, int N>
__device__ __forceinline__ void ent_math(EntS<MODEL, N> &S, const RelS<MODEL, N> &R) {
    float p[N];
    if (MODEL == OKB_TRANSE) {
#pragma unroll
        FOR_N p[i] = S.raw[i];
        S.a = 0.f;
    } else if (MODEL == OKB_TRANSH) {                      // TransH.py:12-14: e - (e.n_hat) n_hat
        S.a = wsum(dot<N>(S.raw, R.aux));
#pragma unroll
        FOR_N p[i] = S.raw[i] - S.a * R.aux[i];
    } else {                                               // TransD.py:23-25: e + (e.e_t) r_t
        S.a = wsum(dot<N>(S.raw, S.aux));
#pragma unroll
        FOR_N p[i] = S.raw[i] + S.a * R.aux[i];
    }
    float ss = wsum(dot<N>(p, p));
    S.proj = ss > EPS_NORM;
    S.inv = rsqrtf(fmaxf(ss, EPS_NORM));
#pragma unroll
    FOR_N S.hat[i] = p[i] * S.inv;
}
template <int MODEL, int VW, int NV, bool COH = false>
__device__ __forceinline__ void ent_forward(EntS<MODEL, VW * NV> &S, const RelS<MODEL, VW * NV> &R, const okb_model &m,
                                            i32 e, int lane) {
    ent_load<MODEL, VW, NV, COH>(S, m, e, lane);
    ent_math<MODEL, VW * NV>(S, R);
}

// score = sum_d |h_hat + r_hat - t_hat| (association as in TransE.py:15); g = sign of the summand
template <int MODEL, int N>
__device__ __forceinline__ float score_fw(const EntS<MODEL, N> &H, const EntS<MODEL, N> &T, const RelS<MODEL, N> &R, float *g) {
    float s = 0.f;
#pragma unroll
    FOR_N {
        const float u = (H.hat[i] + R.rhat[i]) - T.hat[i];
        s += fabsf(u);
        g[i] = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);       // tf.abs gradient: sign(u), 0 at 0
    }
    return wsum(s);
}

// Gradient accumulators of one table-row group: part 0 = main table, part 1 = auxiliary table.
template <int MODEL, int N> struct EntG {
    float d[N];
    float da[MODEL == OKB_TRANSD ? N : 1];
    __device__ __forceinline__ void zero() {
#pragma unroll
        FOR_N d[i] = 0.f;
        if (MODEL == OKB_TRANSD) {
#pragma unroll
            FOR_N da[i] = 0.f;
        }
    }
};
template <int MODEL, int N> struct RelG {
    float d[N];
    float da[MODEL == OKB_TRANSE ? 1 : N];
    __device__ __forceinline__ void zero() {
#pragma unroll
        FOR_N d[i] = 0.f;
        if (MODEL != OKB_TRANSE) {
#pragma unroll
            FOR_N da[i] = 0.f;
        }
    }
};

// Backward of c * score(H, T, R) into the three accumulators.
template <int MODEL, int N>
__device__ __forceinline__ void score_bw(const EntS<MODEL, N> &H, const EntS<MODEL, N> &T, const RelS<MODEL, N> &R,
                                         const float *g, float c, EntG<MODEL, N> &gh, EntG<MODEL, N> &gt, RelG<MODEL, N> &gr) {
    // through l2_normalize: gx = inv * (gy - y_hat (gy . y_hat))   [no projection term when clamped]
    float d1 = dot<N>(g, H.hat), d2 = dot<N>(g, T.hat), d3 = dot<N>(g, R.rhat);
    wsum3(d1, d2, d3);
    if (!H.proj) d1 = 0.f;
    if (!T.proj) d2 = 0.f;
    if (!R.proj) d3 = 0.f;
    float GH[N], GT[N];
#pragma unroll
    FOR_N {
        GH[i] = H.inv * (g[i] - H.hat[i] * d1);
        GT[i] = -T.inv * (g[i] - T.hat[i] * d2);
        gr.d[i] += c * (R.inv * (g[i] - R.rhat[i] * d3));
    }
    if (MODEL == OKB_TRANSE) {
#pragma unroll
        FOR_N { gh.d[i] += c * GH[i]; gt.d[i] += c * GT[i]; }
    } else if (MODEL == OKB_TRANSH) {
        // e' = e - (e.n) n  =>  de = G - n (n.G);  dn_hat = -[(e.n) G + (G.n) e]
        float bh = dot<N>(GH, R.aux), bt = dot<N>(GT, R.aux);
        wsum2(bh, bt);
        float dn[N];
#pragma unroll
        FOR_N {
            gh.d[i] += c * (GH[i] - R.aux[i] * bh);
            gt.d[i] += c * (GT[i] - R.aux[i] * bt);
            dn[i] = -(H.a * GH[i] + bh * H.raw[i] + T.a * GT[i] + bt * T.raw[i]);
        }
        float d4 = wsum(dot<N>(dn, R.aux));
        if (!R.proj_n) d4 = 0.f;
#pragma unroll
        FOR_N gr.da[i] += c * (R.inv_n * (dn[i] - R.aux[i] * d4));
    } else {
        // e' = e + (e.e_t) r_t  =>  de = G + (G.r_t) e_t;  de_t = (G.r_t) e;  dr_t = (e.e_t) G
        float bh = dot<N>(GH, R.aux), bt = dot<N>(GT, R.aux);
        wsum2(bh, bt);
#pragma unroll
        FOR_N {
            gh.d[i] += c * (GH[i] + bh * H.aux[i]);
            gh.da[i] += c * (bh * H.raw[i]);
            gt.d[i] += c * (GT[i] + bt * T.aux[i]);
            gt.da[i] += c * (bt * T.raw[i]);
            gr.da[i] += c * (H.a * GH[i] + T.a * GT[i]);
        }
    }
}

template <int VW, int NV> __device__ __forceinline__ void put(float *dst, const float *src, int D, int lane) {
    Frag<VW, NV> f;
#pragma unroll
    for (int i = 0; i < VW * NV; i++) f.v[i] = src[i];
    f.store(dst, D, lane);
}
template <int MODEL, int VW, int NV>
__device__ __forceinline__ void put_ent(float *row, const EntG<MODEL, VW * NV> &g, int D, int lane) {
    put<VW, NV>(row, g.d, D, lane);
    if (MODEL == OKB_TRANSD) put<VW, NV>(row + D, g.da, D, lane);
}
template <int MODEL, int VW, int NV>
__device__ __forceinline__ void put_rel(float *row, const RelG<MODEL, VW * NV> &g, int D, int lane) {
    put<VW, NV>(row, g.d, D, lane);
    if (MODEL != OKB_TRANSE) put<VW, NV>(row + D, g.da, D, lane);
}

struct GradArgs {
    okb_model m;
    const i32 *bh, *bt, *br;
    float *gent, *grel, *loss_terms;
    float margin, w;           // w = 1 / (B * (k + kr))   (reduce_mean, TransE.py:51)
    i32 B, k, kr, NE, NR, b_lo, b_hi;
    // L2 prefetch list (Adam): the update kernel streams every table with its m and v slots; this gather kernel is
    // latency-bound and leaves HBM idle, so each warp asks the L2 for one slice of every region on its way in.
    const char *pf_ptr[12];
    unsigned pf_bytes[12], pf_slice[12];
    i32 npf;
    i32 slot_base;             // gradient rows of positive b go to slot b - slot_base (a data-parallel rank keeps only its own)
    // owner-sharded data parallelism: the tables are complete once every rank has published `wait_epoch`
    const unsigned long long *wait_flags;
    unsigned long long *announce[OKB_DP_MAX];              // this rank's x_ready word in every rank's flag block
    unsigned long long wait_epoch;
    i32 wait_n;
    // host-batch fast path: the caller's int64 block [3][S] (page-locked, device alias) is compared with the resident batch
    // by `vblocks` extra blocks at the FRONT of the grid; a difference raises bit 1 of *vflag (the update kernels then skip)
    const long long *vh;
    unsigned *vflag;
    i32 vS, vblocks;
    // scatter form of the owner-sharded data-parallel step (sc_world > 0): every gradient row is stored straight into the
    // arena of the rank that OWNS its table row, at the row's slot of the GLOBAL batch, and the hinge term into every
    // rank's loss buffer — the reduce-scatter of the step is these peer stores (train.cu, "data parallel")
    char *sc_arena[OKB_DP_MAX];
    i64 sc_gent, sc_grel, sc_loss;                         // byte offsets of the receive buffers inside an arena
    i32 sc_ent_lo[OKB_DP_MAX + 1], sc_rel_lo[OKB_DP_MAX + 1];
    i32 sc_world;
    // gather form (sc_gather): every row goes to EVERY rank (arena[0] + sc_delta[p]) and every rank runs the full update
    i32 sc_gather;
    i64 sc_delta[OKB_DP_MAX];
    unsigned long long *trace;
    i32 hs_mode;               // flag handshake: 0 = release store / acquire polls, bit 0 = relaxed polls + one fence, bit 1 = relaxed store
};
__device__ __forceinline__ void st_flag_sys(unsigned long long *p, unsigned long long v, int relaxed) {
    if (relaxed) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// destination of the gradient row in global slot `slot` whose table row is entity / relation `id` (id < 0: the slot is not
// part of any segment — a negative equal to its positive — and nothing is stored)
// SC = false (single GPU, push form): plain local rows, no owner search compiled in (the k = 1 kernel is 20 % shorter)
template <bool SC> __device__ __forceinline__ float *grad_dst_ent(const GradArgs &a, i32 id, i64 slot, int ce) {
    if (!SC || !a.sc_world) return a.gent + slot * ce;
    if (id < 0) return nullptr;
    int o = 0;
    if (!a.sc_gather) while (id >= a.sc_ent_lo[o + 1]) o++;
    return reinterpret_cast<float *>(a.sc_arena[o] + a.sc_gent) + slot * ce;
}
template <bool SC> __device__ __forceinline__ float *grad_dst_rel(const GradArgs &a, i32 id, i64 slot, int cr) {
    if (!SC || !a.sc_world) return a.grel + slot * cr;
    if (id < 0) return nullptr;
    int o = 0;
    if (!a.sc_gather) while (id >= a.sc_rel_lo[o + 1]) o++;
    return reinterpret_cast<float *>(a.sc_arena[o] + a.sc_grel) + slot * cr;
}
template <bool SC> __device__ __forceinline__ void grad_put_loss(const GradArgs &a, i32 b, float term, int lane) {
    if (!SC || !a.sc_world) { if (lane == 0) a.loss_terms[b - a.slot_base] = term; return; }
    term = __shfl_sync(FULL, term, 0);
    if (lane < a.sc_world) reinterpret_cast<float *>(a.sc_arena[lane] + a.sc_loss)[b - a.slot_base] = term;
}
// store a gradient row at its destination (and, in the gather form, at the same place in every other rank's arena)
template <int MODEL, int VW, int NV, bool SC>
__device__ __forceinline__ void put_ent_sc(const GradArgs &a, float *dst, const EntG<MODEL, VW * NV> &g, int D, int lane) {
    put_ent<MODEL, VW, NV>(dst, g, D, lane);
    if (SC && a.sc_gather)
        for (int p = 1; p < a.sc_world; p++) put_ent<MODEL, VW, NV>(reinterpret_cast<float *>(reinterpret_cast<char *>(dst) + a.sc_delta[p]), g, D, lane);
}
template <int MODEL, int VW, int NV, bool SC>
__device__ __forceinline__ void put_rel_sc(const GradArgs &a, float *dst, const RelG<MODEL, VW * NV> &g, int D, int lane) {
    put_rel<MODEL, VW, NV>(dst, g, D, lane);
    if (SC && a.sc_gather)
        for (int p = 1; p < a.sc_world; p++) put_rel<MODEL, VW, NV>(reinterpret_cast<float *>(reinterpret_cast<char *>(dst) + a.sc_delta[p]), g, D, lane);
}
__device__ __forceinline__ void grad_verify_block(const GradArgs &a, i32 vb) {
    // one element per thread: every PCIe read of the launch is in flight at once (a loop per thread would serialise
    // round trips of ~2 us each)
    const i32 S = a.vS, i = vb * (i32)blockDim.x + (i32)threadIdx.x;
    if (i >= S) return;
    const i32 *dev = a.bh;                                  // [3][S]: bh, bt = bh + S, br = bh + 2S
    const long long vh = a.vh[i], vt = a.vh[S + i], vr = a.vh[2 * (i64)S + i];
    if (vh != (long long)dev[i] || vt != (long long)dev[S + i] || vr != (long long)dev[2 * (i64)S + i]) atomicOr(a.vflag, 2u);
}

// Everything of one positive group after its batch ids are known (ph, pt, pr and the first negative's nh, nt): gathers,
// forward, hinge, backward, gradient rows.  COH: coherent (L2) gathers, for the persistent chunk kernel.
template <int MODEL, int VW, int NV, int WPPMAX, bool COH, bool SC = false>
__device__ __forceinline__ void grad_body(const GradArgs &a, i32 b, int lane, int wid, int wpp, float *grad_sm, i32 ph, i32 pt, i32 pr,
                                          i32 nh, i32 nt) {
    constexpr int N = VW * NV;
    const int D = a.m.ent_dim;
    const int ce = MODEL == OKB_TRANSD ? 2 * D : D, cr = MODEL == OKB_TRANSE ? D : 2 * D;
    RelS<MODEL, N> Rp;
    EntS<MODEL, N> Hp, Tp, Nx;                             // Nx: the current negative's replacement entity
    rel_load<MODEL, VW, NV, COH>(Rp, a.m, pr, lane);            // all gathers of the group go out before the first reduction
    ent_load<MODEL, VW, NV, COH>(Hp, a.m, ph, lane);
    ent_load<MODEL, VW, NV, COH>(Tp, a.m, pt, lane);
    if (wid < a.k) ent_load<MODEL, VW, NV, COH>(Nx, a.m, nh != ph ? nh : nt, lane);
    rel_math<MODEL, N>(Rp);
    ent_math<MODEL, N>(Hp, Rp);
    ent_math<MODEL, N>(Tp, Rp);
    float gp[N];
    const float sp = score_fw<MODEL, N>(Hp, Tp, Rp, gp);

    EntG<MODEL, N> accH, accT;
    RelG<MODEL, N> accR;
    accH.zero(); accT.zero(); accR.zero();
    const i64 es = (i64)(b - a.slot_base) * a.NE, rs = (i64)(b - a.slot_base) * a.NR;      // first entity / relation slot
    float hinge_sum = 0.f;
    i32 active = 0;

    for (i32 m = wid; m < a.k; m += wpp) {                 // entity negatives (Base.cpp:113-131)
        const i32 cnh = nh, cnt_ = nt;
        EntS<MODEL, N> Nn = Nx;
        if (m + wpp < a.k) {                               // next negative's row is in flight while this one is scored
            const i32 at = b + (m + wpp + 1) * a.B;
            nh = a.bh[at]; nt = a.bt[at];
            ent_load<MODEL, VW, NV, COH>(Nx, a.m, nh != ph ? nh : nt, lane);
        }
        EntG<MODEL, N> gnew;
        gnew.zero();
        float gn[N];
        if (cnh != ph) {                                   // head replaced; (t, r) rows shared
            ent_math<MODEL, N>(Nn, Rp);
            const float sn = score_fw<MODEL, N>(Nn, Tp, Rp, gn);
            const float x = sp - sn + a.margin;
            if (x >= 0.f) { hinge_sum += x; active++; score_bw<MODEL, N>(Nn, Tp, Rp, gn, -a.w, gnew, accT, accR); }
        } else if (cnt_ != pt) {                           // tail replaced; (h, r) rows shared
            ent_math<MODEL, N>(Nn, Rp);
            const float sn = score_fw<MODEL, N>(Hp, Nn, Rp, gn);
            const float x = sp - sn + a.margin;
            if (x >= 0.f) { hinge_sum += x; active++; score_bw<MODEL, N>(Hp, Nn, Rp, gn, -a.w, accH, gnew, accR); }
        } else {                                           // degenerate: negative == positive
            const float x = a.margin;
            if (x >= 0.f) { hinge_sum += x; active++; score_bw<MODEL, N>(Hp, Tp, Rp, gp, -a.w, accH, accT, accR); }
        }
        if (float *dst = grad_dst_ent<SC>(a, cnh != ph ? cnh : (cnt_ != pt ? cnt_ : -1), es + 2 + m, ce)) put_ent_sc<MODEL, VW, NV, SC>(a, dst, gnew, D, lane);
    }
    for (i32 m = 0; m < (wid == 0 ? a.kr : 0); m++) {      // relation negatives (Base.cpp:133-139): warp 0
        const i32 nr = a.br[b + (1 + a.k + m) * a.B];
        RelG<MODEL, N> gnew;
        gnew.zero();
        float gn[N];
        if (nr != pr) {
            RelS<MODEL, N> Rn;
            rel_forward<MODEL, VW, NV, COH>(Rn, a.m, nr, lane);
            if (MODEL == OKB_TRANSE) {
                const float sn = score_fw<MODEL, N>(Hp, Tp, Rn, gn);
                const float x = sp - sn + a.margin;
                if (x >= 0.f) { hinge_sum += x; active++; score_bw<MODEL, N>(Hp, Tp, Rn, gn, -a.w, accH, accT, gnew); }
            } else {                                       // projections depend on the relation: redo both sides
                EntS<MODEL, N> Hn, Tn;
                ent_forward<MODEL, VW, NV, COH>(Hn, Rn, a.m, ph, lane);
                ent_forward<MODEL, VW, NV, COH>(Tn, Rn, a.m, pt, lane);
                const float sn = score_fw<MODEL, N>(Hn, Tn, Rn, gn);
                const float x = sp - sn + a.margin;
                if (x >= 0.f) { hinge_sum += x; active++; score_bw<MODEL, N>(Hn, Tn, Rn, gn, -a.w, accH, accT, gnew); }
            }
        } else {
            const float x = a.margin;
            if (x >= 0.f) { hinge_sum += x; active++; score_bw<MODEL, N>(Hp, Tp, Rp, gp, -a.w, accH, accT, accR); }
        }
        if (float *dst = grad_dst_rel<SC>(a, nr != pr ? nr : -1, rs + 1 + m, cr)) put_rel_sc<MODEL, VW, NV, SC>(a, dst, gnew, D, lane);
    }
    if (WPPMAX > 1 && wpp > 1) {                           // warps 1.. hand their accumulators to warp 0, which adds them in warp order
        constexpr int F = 2 * (MODEL == OKB_TRANSD ? 2 : 1) + (MODEL == OKB_TRANSE ? 1 : 2);      // fragments per warp
        float *slab = grad_sm + (size_t)(wid > 0 ? wid - 1 : 0) * (F * 32 * N + 64);
        if (wid > 0) {
            int f = 0;
            auto out = [&](const float *v) {
#pragma unroll
                for (int i = 0; i < N; i++) slab[(f * 32 + lane) * N + i] = v[i];
                f++;
            };
            out(accH.d); out(accT.d); out(accR.d);
            if (MODEL == OKB_TRANSD) { out(accH.da); out(accT.da); }
            if (MODEL != OKB_TRANSE) out(accR.da);
            if (lane == 0) { slab[F * 32 * N] = hinge_sum; slab[F * 32 * N + 1] = (float)active; }
        }
        __syncthreads();
        if (wid > 0) return;
        for (int w = 1; w < wpp; w++) {
            const float *sl = grad_sm + (size_t)(w - 1) * (F * 32 * N + 64);
            int f = 0;
            auto in = [&](float *v) {
#pragma unroll
                for (int i = 0; i < N; i++) v[i] += sl[(f * 32 + lane) * N + i];
                f++;
            };
            in(accH.d); in(accT.d); in(accR.d);
            if (MODEL == OKB_TRANSD) { in(accH.da); in(accT.da); }
            if (MODEL != OKB_TRANSE) in(accR.da);
            hinge_sum += sl[F * 32 * N];
            active += (i32)sl[F * 32 * N + 1];
        }
    }
    if (active) score_bw<MODEL, N>(Hp, Tp, Rp, gp, a.w * (float)active, accH, accT, accR);
    put_ent_sc<MODEL, VW, NV, SC>(a, grad_dst_ent<SC>(a, ph, es, ce), accH, D, lane);
    put_ent_sc<MODEL, VW, NV, SC>(a, grad_dst_ent<SC>(a, pt, es + 1, ce), accT, D, lane);
    put_rel_sc<MODEL, VW, NV, SC>(a, grad_dst_rel<SC>(a, pr, rs, cr), accR, D, lane);
    grad_put_loss<SC>(a, b, hinge_sum, lane);
}

// WPPMAX = 1: one warp per positive (GRAD_WARPS positives per block).  WPPMAX = 4: ONE positive per block and blockDim / 32
// (2..4) warps share its entity negatives — small batches with many negatives (WN18-shaped: B = 1,414, k = 10) otherwise
// leave most warp slots empty while each warp walks its negatives one after the other.  Every warp recomputes the
// positive's forward pass, takes the negatives m = w, w + wpp, ..., and warp 0 adds the other warps' accumulators in warp
// order (through shared memory) before the positive's own backward pass.
template <int MODEL, int VW, int NV, int WPPMAX, bool SC = false>
__global__ void __launch_bounds__(WPPMAX == 1 ? GRAD_WARPS * 32 : WPPMAX * 32, WPPMAX == 1 ? GRAD_MIN_BLOCKS : 5) grad_kernel(GradArgs a) {
    extern __shared__ float grad_sm[];
    const int lane = threadIdx.x & 31;
    const int wpp = WPPMAX == 1 ? 1 : (int)(blockDim.x >> 5), wid = WPPMAX == 1 ? 0 : (int)(threadIdx.x >> 5);
    // the verification blocks come FIRST in the grid: their PCIe reads are in flight for the whole launch
    if ((i32)blockIdx.x < a.vblocks) { grad_verify_block(a, (i32)blockIdx.x); return; }
    const i32 bx = (i32)blockIdx.x - a.vblocks;
    const i32 b = WPPMAX == 1 ? a.b_lo + bx * GRAD_WARPS + (i32)(threadIdx.x >> 5) : a.b_lo + bx;
    if (b >= a.b_hi) return;
    const i32 ph = a.bh[b], pt = a.bt[b], pr = a.br[b];
    // entity that replaces a side in negative m (Base.cpp:118-126): the new head if the head changed, else the tail
    i32 nh = 0, nt = 0;
    if (wid < a.k) { nh = a.bh[b + (wid + 1) * a.B]; nt = a.bt[b + (wid + 1) * a.B]; }
    // The batch ids do not depend on the previous update kernel; the table rows do.  Dependents (this step's update
    // kernel) are released only after the wait, so they can never start before the previous update has finished.
    if (threadIdx.x == 0) trace_min(SC ? a.trace : nullptr, 0);
    pdl_wait();
    if (threadIdx.x == 0) trace_min(SC ? a.trace : nullptr, 1);
    pdl_launch_dependents();
    if (a.wait_flags) {                                    // owner-sharded data parallelism: peers still publishing rows?
        if (lane < a.wait_n) {
            // this rank's previous owner-update kernel has completed (griddepcontrol.wait above): tell every peer, then
            // wait until every peer has said the same
            if (blockIdx.x == 0 && threadIdx.x < 32)
                st_flag_sys(a.announce[lane], a.wait_epoch, a.hs_mode & 2);
            spin_until(a.wait_flags + lane, a.wait_epoch, a.hs_mode & 1);
        }
        __syncwarp();
    }
    if (threadIdx.x == 0) { trace_min(SC ? a.trace : nullptr, 2); trace_max(SC ? a.trace : nullptr, 3); }
    if (lane < a.npf) {
        const unsigned off = (unsigned)(b - a.b_lo) * a.pf_slice[lane];
        if (off < a.pf_bytes[lane]) {
            const unsigned sz = min(a.pf_slice[lane], a.pf_bytes[lane] - off) & ~15u;
            if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.pf_ptr[lane] + off), "r"(sz) : "memory");
        }
    }

    grad_body<MODEL, VW, NV, WPPMAX, false, SC>(a, b, lane, wid, wpp, grad_sm, ph, pt, pr, nh, nt);
    if (threadIdx.x == 0) trace_max(SC ? a.trace : nullptr, 4);
}

// ------------------------------------------------------------------------------------------ grad, k = 1 specialisation
// The generic kernel walks 16 dependent shuffle-reduction chains per positive (norms, projections, scores, the
// backward dots of the negative and then of the positive): with 5 warps per scheduler the issue slots idle while
// every warp sits in a chain (ncu: issue-active 50 %, short-scoreboard + wait stalls dominate).  With one entity
// negative and no relation negative — the reference's default batch (Config.py:58-59) — the positive and its negative
// are independent until the hinge, so their reductions are interleaved: 7 chains of 2–6 values each.  Same
// operations, same per-value reduction tree, same accumulation order (negative first, then the positive) as the
// generic kernel.
#ifndef GRAD1_MIN_BLOCKS
#define GRAD1_MIN_BLOCKS 16
#endif
template <int K> __device__ __forceinline__ void wsumk(float (&v)[K]) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int i = 0; i < K; i++) v[i] += __shfl_xor_sync(FULL, v[i], o);
    }
}
template <int MODEL, int N>
__device__ __forceinline__ void ent_select(EntS<MODEL, N> &D, bool take, const EntS<MODEL, N> &A, const EntS<MODEL, N> &B) {
#pragma unroll
    FOR_N { D.raw[i] = take ? A.raw[i] : B.raw[i]; D.hat[i] = take ? A.hat[i] : B.hat[i]; }
    if (MODEL == OKB_TRANSD) {
#pragma unroll
        FOR_N D.aux[i] = take ? A.aux[i] : B.aux[i];
    }
    D.inv = take ? A.inv : B.inv; D.a = take ? A.a : B.a; D.proj = take ? A.proj : B.proj;
}

template <int MODEL, int VW, int NV, bool COH, bool SC = false>
__device__ __forceinline__ void grad_k1_body(const GradArgs &a, i32 b, int lane, i32 ph, i32 pt, i32 pr, i32 nh, i32 nt) {
    constexpr int N = VW * NV;
    const int D = a.m.ent_dim;
    const int ce = MODEL == OKB_TRANSD ? 2 * D : D, cr = MODEL == OKB_TRANSE ? D : 2 * D;
    const bool head_rep = nh != ph, tail_rep = !head_rep && nt != pt;

    RelS<MODEL, N> R;
    EntS<MODEL, N> H, T, X;                                // X: the entity that replaces a side in the negative
    rel_load<MODEL, VW, NV, COH>(R, a.m, pr, lane);
    ent_load<MODEL, VW, NV, COH>(H, a.m, ph, lane);
    ent_load<MODEL, VW, NV, COH>(T, a.m, pt, lane);
    ent_load<MODEL, VW, NV, COH>(X, a.m, head_rep ? nh : nt, lane);

    // ---- relation: both norms in one chain
    {
        float s[2] = {dot<N>(R.rhat, R.rhat), MODEL == OKB_TRANSH ? dot<N>(R.aux, R.aux) : 0.f};
        if (MODEL == OKB_TRANSH) wsumk<2>(s); else s[0] = wsum(s[0]);
        R.proj = s[0] > EPS_NORM;
        R.inv = rsqrtf(fmaxf(s[0], EPS_NORM));
#pragma unroll
        FOR_N R.rhat[i] = R.rhat[i] * R.inv;
        if (MODEL == OKB_TRANSH) {
            R.proj_n = s[1] > EPS_NORM;
            R.inv_n = rsqrtf(fmaxf(s[1], EPS_NORM));
#pragma unroll
            FOR_N R.aux[i] = R.aux[i] * R.inv_n;
        }
    }
    // ---- entities: the three projection dots in one chain, the three norms in the next
    float pH[N], pT[N], pX[N];
    if (MODEL == OKB_TRANSE) {
#pragma unroll
        FOR_N { pH[i] = H.raw[i]; pT[i] = T.raw[i]; pX[i] = X.raw[i]; }
        H.a = T.a = X.a = 0.f;
    } else {
        float av[3];
        if (MODEL == OKB_TRANSH) { av[0] = dot<N>(H.raw, R.aux); av[1] = dot<N>(T.raw, R.aux); av[2] = dot<N>(X.raw, R.aux); }
        else { av[0] = dot<N>(H.raw, H.aux); av[1] = dot<N>(T.raw, T.aux); av[2] = dot<N>(X.raw, X.aux); }
        wsumk<3>(av);
        H.a = av[0]; T.a = av[1]; X.a = av[2];
        if (MODEL == OKB_TRANSH) {
#pragma unroll
            FOR_N { pH[i] = H.raw[i] - H.a * R.aux[i]; pT[i] = T.raw[i] - T.a * R.aux[i]; pX[i] = X.raw[i] - X.a * R.aux[i]; }
        } else {
#pragma unroll
            FOR_N { pH[i] = H.raw[i] + H.a * R.aux[i]; pT[i] = T.raw[i] + T.a * R.aux[i]; pX[i] = X.raw[i] + X.a * R.aux[i]; }
        }
    }
    {
        float sv[3] = {dot<N>(pH, pH), dot<N>(pT, pT), dot<N>(pX, pX)};
        wsumk<3>(sv);
        H.proj = sv[0] > EPS_NORM; H.inv = rsqrtf(fmaxf(sv[0], EPS_NORM));
        T.proj = sv[1] > EPS_NORM; T.inv = rsqrtf(fmaxf(sv[1], EPS_NORM));
        X.proj = sv[2] > EPS_NORM; X.inv = rsqrtf(fmaxf(sv[2], EPS_NORM));
#pragma unroll
        FOR_N { H.hat[i] = pH[i] * H.inv; T.hat[i] = pT[i] * T.inv; X.hat[i] = pX[i] * X.inv; }
    }
    // ---- the negative triple (Hn, Tn, r): the replaced side comes from X, a degenerate negative is the positive itself
    EntS<MODEL, N> Hn, Tn;
    ent_select<MODEL, N>(Hn, head_rep, X, H);
    ent_select<MODEL, N>(Tn, tail_rep, X, T);
    float gp[N], gn[N];
    float sc[2] = {0.f, 0.f};
#pragma unroll
    FOR_N {
        const float up = (H.hat[i] + R.rhat[i]) - T.hat[i], un = (Hn.hat[i] + R.rhat[i]) - Tn.hat[i];
        sc[0] += fabsf(up); sc[1] += fabsf(un);
        gp[i] = up > 0.f ? 1.f : (up < 0.f ? -1.f : 0.f);
        gn[i] = un > 0.f ? 1.f : (un < 0.f ? -1.f : 0.f);
    }
    wsumk<2>(sc);
    const float x = sc[0] - sc[1] + a.margin;
    const bool active = x >= 0.f;

    EntG<MODEL, N> accH, accT, gnew;
    RelG<MODEL, N> accR;
    accH.zero(); accT.zero(); gnew.zero(); accR.zero();
    if (active) {
        const float w = a.w;
        // through l2_normalize, negative and positive together: six dots in one chain
        float d[6] = {dot<N>(gn, Hn.hat), dot<N>(gn, Tn.hat), dot<N>(gn, R.rhat), dot<N>(gp, H.hat), dot<N>(gp, T.hat), dot<N>(gp, R.rhat)};
        wsumk<6>(d);
        if (!Hn.proj) d[0] = 0.f;
        if (!Tn.proj) d[1] = 0.f;
        if (!H.proj) d[3] = 0.f;
        if (!T.proj) d[4] = 0.f;
        if (!R.proj) { d[2] = 0.f; d[5] = 0.f; }
        float GHn[N], GTn[N], GHp[N], GTp[N];
#pragma unroll
        FOR_N {
            GHn[i] = Hn.inv * (gn[i] - Hn.hat[i] * d[0]);
            GTn[i] = -Tn.inv * (gn[i] - Tn.hat[i] * d[1]);
            accR.d[i] += -w * (R.inv * (gn[i] - R.rhat[i] * d[2]));
            GHp[i] = H.inv * (gp[i] - H.hat[i] * d[3]);
            GTp[i] = -T.inv * (gp[i] - T.hat[i] * d[4]);
            accR.d[i] += w * (R.inv * (gp[i] - R.rhat[i] * d[5]));
        }
        EntG<MODEL, N> dHn, dTn;                           // the negative's entity gradients, routed below
        dHn.zero(); dTn.zero();
        if (MODEL == OKB_TRANSE) {
#pragma unroll
            FOR_N { dHn.d[i] = -w * GHn[i]; dTn.d[i] = -w * GTn[i]; accH.d[i] = 0.f; }
        } else {
            float bb[4] = {dot<N>(GHn, R.aux), dot<N>(GTn, R.aux), dot<N>(GHp, R.aux), dot<N>(GTp, R.aux)};
            wsumk<4>(bb);
            if (MODEL == OKB_TRANSH) {
                float dnn[N], dnp[N];
#pragma unroll
                FOR_N {
                    dHn.d[i] = -w * (GHn[i] - R.aux[i] * bb[0]);
                    dTn.d[i] = -w * (GTn[i] - R.aux[i] * bb[1]);
                    dnn[i] = -(Hn.a * GHn[i] + bb[0] * Hn.raw[i] + Tn.a * GTn[i] + bb[1] * Tn.raw[i]);
                    dnp[i] = -(H.a * GHp[i] + bb[2] * H.raw[i] + T.a * GTp[i] + bb[3] * T.raw[i]);
                }
                float d4[2] = {dot<N>(dnn, R.aux), dot<N>(dnp, R.aux)};
                wsumk<2>(d4);
                if (!R.proj_n) { d4[0] = 0.f; d4[1] = 0.f; }
#pragma unroll
                FOR_N {
                    accR.da[i] += -w * (R.inv_n * (dnn[i] - R.aux[i] * d4[0]));
                    accR.da[i] += w * (R.inv_n * (dnp[i] - R.aux[i] * d4[1]));
                }
            } else {
#pragma unroll
                FOR_N {
                    dHn.d[i] = -w * (GHn[i] + bb[0] * Hn.aux[i]);
                    dHn.da[i] = -w * (bb[0] * Hn.raw[i]);
                    dTn.d[i] = -w * (GTn[i] + bb[1] * Tn.aux[i]);
                    dTn.da[i] = -w * (bb[1] * Tn.raw[i]);
                    accR.da[i] += -w * (Hn.a * GHn[i] + Tn.a * GTn[i]);
                    accR.da[i] += w * (H.a * GHp[i] + T.a * GTp[i]);
                }
            }
            // the positive's entity gradients use bb[2], bb[3] below
            if (MODEL == OKB_TRANSH) {
#pragma unroll
                FOR_N { GHp[i] = GHp[i] - R.aux[i] * bb[2]; GTp[i] = GTp[i] - R.aux[i] * bb[3]; }
            }
            if (MODEL == OKB_TRANSD) {
                // route the negative first (accumulation order of the generic kernel), then add the positive
#pragma unroll
                FOR_N {
                    accH.d[i] = head_rep ? 0.f : dHn.d[i]; accH.da[i] = head_rep ? 0.f : dHn.da[i];
                    accT.d[i] = tail_rep ? 0.f : dTn.d[i]; accT.da[i] = tail_rep ? 0.f : dTn.da[i];
                    gnew.d[i] = head_rep ? dHn.d[i] : (tail_rep ? dTn.d[i] : 0.f);
                    gnew.da[i] = head_rep ? dHn.da[i] : (tail_rep ? dTn.da[i] : 0.f);
                    accH.d[i] += w * (GHp[i] + bb[2] * H.aux[i]); accH.da[i] += w * (bb[2] * H.raw[i]);
                    accT.d[i] += w * (GTp[i] + bb[3] * T.aux[i]); accT.da[i] += w * (bb[3] * T.raw[i]);
                }
            }
        }
        if (MODEL != OKB_TRANSD) {
#pragma unroll
            FOR_N {
                accH.d[i] = head_rep ? 0.f : dHn.d[i];
                accT.d[i] = tail_rep ? 0.f : dTn.d[i];
                gnew.d[i] = head_rep ? dHn.d[i] : (tail_rep ? dTn.d[i] : 0.f);
                accH.d[i] += w * GHp[i];
                accT.d[i] += w * GTp[i];
            }
        }
    }
    const i64 es = (i64)(b - a.slot_base) * a.NE, rs = (i64)(b - a.slot_base) * a.NR;
    put_ent_sc<MODEL, VW, NV, SC>(a, grad_dst_ent<SC>(a, ph, es, ce), accH, D, lane);
    put_ent_sc<MODEL, VW, NV, SC>(a, grad_dst_ent<SC>(a, pt, es + 1, ce), accT, D, lane);
    if (float *dst = grad_dst_ent<SC>(a, nh != ph ? nh : (nt != pt ? nt : -1), es + 2, ce)) put_ent_sc<MODEL, VW, NV, SC>(a, dst, gnew, D, lane);
    put_rel_sc<MODEL, VW, NV, SC>(a, grad_dst_rel<SC>(a, pr, rs, cr), accR, D, lane);
    grad_put_loss<SC>(a, b, active ? x : 0.f, lane);
}

template <int MODEL, int VW, int NV, bool SC = false>
__global__ void __launch_bounds__(32, GRAD1_MIN_BLOCKS) grad_k1_kernel(GradArgs a) {
    const int lane = threadIdx.x & 31;
    if ((i32)blockIdx.x < a.vblocks) { grad_verify_block(a, (i32)blockIdx.x); return; }      // first in the grid: see grad_kernel
    const i32 b = a.b_lo + (i32)blockIdx.x - a.vblocks;
    if (b >= a.b_hi) return;
    const i32 ph = a.bh[b], pt = a.bt[b], pr = a.br[b];
    const i32 nh = a.bh[b + a.B], nt = a.bt[b + a.B];
    if (lane == 0) trace_min(SC ? a.trace : nullptr, 0);
    pdl_wait();
    if (lane == 0) trace_min(SC ? a.trace : nullptr, 1);
    pdl_launch_dependents();
    if (a.wait_flags) {                                    // owner-sharded data parallelism: see grad_kernel
        if (lane < a.wait_n) {
            if (blockIdx.x == 0)
                st_flag_sys(a.announce[lane], a.wait_epoch, a.hs_mode & 2);
            spin_until(a.wait_flags + lane, a.wait_epoch, a.hs_mode & 1);
        }
        __syncwarp();
    }
    if (lane == 0) { trace_min(SC ? a.trace : nullptr, 2); trace_max(SC ? a.trace : nullptr, 3); }
    grad_k1_body<MODEL, VW, NV, false, SC>(a, b, lane, ph, pt, pr, nh, nt);
    if (lane == 0) trace_max(SC ? a.trace : nullptr, 4);
}

// ------------------------------------------------------------------------------------------ update
// one parameter table for the flat Adam pass; vec_end: cumulative vector count over the table list
struct DenseTab { float *x, *m, *v; const float *grad; i64 vec_end; i32 D, key_off, cols, part, slot_off, blk_end;
                  i32 sblk_end;                // cumulative count of (256 * vectors-per-thread)-vector super tiles (adam_tile_kernel)
                  unsigned magic, shift; };    // row = vector / (D / VW) as __umulhi(vector, magic) >> shift (magic 0: shift only)
struct UpdArgs {
    okb_model m;
    okb_hyper hp;
    const i32 *skeys, *perm;   // sorted keys / slot of each sorted position (this step)
    const float *gent, *grel, *loss_terms;
    float *loss_out;
    const int4 *rowhead;       // Adam: {first, end, slot0, slot1}: sorted-position range of each table row in this step's
                               // plan (first = -1 if untouched) with its first two gradient slots inlined
    DenseTab tab[4];
    const float *partial;      // hub path: chunk sums of long segments, row i = sum of sorted positions [i, i + PCH)
    i32 pcols, hub, by_row, loss_blocks;
    float *loss_part;
    unsigned *loss_ctr;
    const unsigned *bad;       // "bad id" flag of the host-batch path (narrow_kernel): set -> tables untouched, loss = NaN
    i32 n, n_ent_slots, E, R, ce, cr, B, key_limit, ntab, work_blocks;
    float w;
};
__device__ __forceinline__ bool upd_bad(const UpdArgs &a) { return a.bad && *(const volatile unsigned *)a.bad != 0u; }

// mean hinge over B*(k+kr) pairs in a fixed order, by `loss_blocks` extra blocks of the update launch: each
// sums a fixed contiguous range of the per-positive terms; the block that finishes last adds the partial
// sums in index order (so the result does not depend on which block that is).
// T: the first T threads of the block do the summing (T = blockDim.x in the per-phase kernels; the persistent chunk kernel
// passes the block size of the per-phase kernel it replaces, so that the fp32 order — hence the loss bits — is the same).
// Every thread of the block must call it (block barriers inside).
__device__ __forceinline__ void loss_block(const UpdArgs &a, i32 lb, i32 T) {
    __shared__ float sh[32];
    __shared__ unsigned ticket;
    if (!a.loss_out) return;
    const i32 per = (a.B + a.loss_blocks - 1) / a.loss_blocks;
    const i32 lo = lb * per, hi = min(a.B, lo + per);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if ((i32)threadIdx.x < T) {
        for (i32 i = lo + threadIdx.x; i < hi; i += 4 * T) {
            s0 += __ldcg(a.loss_terms + i);
            if (i + T < hi) s1 += __ldcg(a.loss_terms + i + T);
            if (i + 2 * T < hi) s2 += __ldcg(a.loss_terms + i + 2 * T);
            if (i + 3 * T < hi) s3 += __ldcg(a.loss_terms + i + 3 * T);
        }
    }
    float s = wsum((s0 + s1) + (s2 + s3));
    if ((threadIdx.x & 31) == 0 && (i32)threadIdx.x < T) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = wsum((i32)threadIdx.x < (T >> 5) ? sh[threadIdx.x] : 0.f);
        if (threadIdx.x == 0) {
            a.loss_part[lb] = s;
            __threadfence();
            ticket = atomicAdd(a.loss_ctr, 1u);
        }
    }
    __syncthreads();
    if (ticket == (unsigned)a.loss_blocks - 1 && threadIdx.x == 0) {
        __threadfence();
        float tot = 0.f;
        for (i32 q = 0; q < a.loss_blocks; q++) tot += ((volatile float *)a.loss_part)[q];
        a.loss_out[0] = upd_bad(a) ? __int_as_float(0x7fc00000) : tot * a.w;
        *a.loss_ctr = 0u;
        __threadfence_system();                            // loss_out may be page-locked host memory a waiting caller polls
    }
}

// Gradient row of sorted position j, part `part` (D floats)
__device__ __forceinline__ const float *grad_row(const UpdArgs &a, i32 j, bool is_ent, int D, int part) {
    const i32 slot = __ldg(a.perm + j);                   // the plan is read-only while the train kernels run
    return (is_ent ? a.gent + (i64)slot * a.ce : a.grel + (i64)(slot - a.n_ent_slots) * a.cr) + part * D;
}

// Sum `cnt` rows starting at index lo with stride `step` (raw gradient rows via perm[], or pre-reduced
// block sums); loads are issued four at a time, additions stay in ascending order.
template <int VW, int NV, bool COH = false>
__device__ __forceinline__ void sum_rows(const UpdArgs &a, i32 lo, i32 hi, bool is_ent, int D, int part, int lane, float *acc,
                                         bool from_partial) {
    constexpr int N = VW * NV;
    for (i32 j = lo; j < hi; j += 4) {
        Frag<VW, NV> f[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const i32 jj = j + u;
            if (jj < hi) f[u].template load<COH>(from_partial ? a.partial + (i64)jj * a.pcols + part * D : grad_row(a, jj, is_ent, D, part), D, lane);
            else f[u].zero();
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int q = 0; q < N; q++) acc[q] += f[u].v[q];
    }
}
// Whole segment [s, e).  Long segments (hub rows) use the block sums of prereduce_kernel for the PCH-aligned
// blocks that lie inside the segment and raw rows for the two fringes — always in ascending position order.
template <int VW, int NV, bool COH = false>
__device__ __forceinline__ void seg_sum(const UpdArgs &a, i32 s, i32 e, bool is_ent, int D, int part, int lane, float *acc) {
    const i32 b0 = (s + PCH - 1) / PCH, b1 = e / PCH;     // blocks [b0, b1) are interior
    if (a.hub && e - s > PCH && b0 < b1) {
        sum_rows<VW, NV, COH>(a, s, b0 * PCH, is_ent, D, part, lane, acc, false);
        sum_rows<VW, NV, COH>(a, b0, b1, is_ent, D, part, lane, acc, true);
        sum_rows<VW, NV, COH>(a, b1 * PCH, e, is_ent, D, part, lane, acc, false);
    } else {
        sum_rows<VW, NV, COH>(a, s, e, is_ent, D, part, lane, acc, false);
    }
}

// Same sum driven by the row-map entry {first, end, slot0, slot1}: the gradient slots of the first two entries are inlined
// there, so the common short segment needs no perm[] hop at all.  Same ascending order as seg_sum -> same bits.
template <int VW, int NV, bool COH = false>
__device__ __forceinline__ void seg_sum_map(const UpdArgs &a, const int4 seg, bool is_ent, int D, int part, int lane, float *acc) {
    constexpr int N = VW * NV;
    const i32 cnt = seg.y - seg.x;
    if (a.hub && cnt > PCH && (seg.x + PCH - 1) / PCH < seg.y / PCH) { seg_sum<VW, NV, COH>(a, seg.x, seg.y, is_ent, D, part, lane, acc); return; }
    auto row = [&](i32 slot) { return (is_ent ? a.gent + (i64)slot * a.ce : a.grel + (i64)(slot - a.n_ent_slots) * a.cr) + part * D; };
    Frag<VW, NV> f0, f1;
    f0.template load<COH>(row(seg.z), D, lane);
    if (cnt > 1) f1.template load<COH>(row(seg.w), D, lane); else f1.zero();
#pragma unroll
    for (int q = 0; q < N; q++) acc[q] += f0.v[q];
    if (cnt > 1) {
#pragma unroll
        for (int q = 0; q < N; q++) acc[q] += f1.v[q];
    }
    if (cnt > 2) sum_rows<VW, NV, COH>(a, seg.x + 2, seg.y, is_ent, D, part, lane, acc, false);
}

// Hub path, level 1: one warp per PCH-aligned block of sorted positions; a block that lies inside ONE
// segment is summed into partial[block].  Block boundaries depend only on positions: fixed order.
template <int VW, int NV, bool COH>
__device__ __forceinline__ void prereduce_body(const UpdArgs &a, float *partial, i32 w, int lane) {
    constexpr int N = VW * NV;
    const i32 lo = w * PCH;
    if (lo + PCH > a.n) return;
    const i32 key = a.skeys[lo];
    if (key >= a.key_limit || a.skeys[lo + PCH - 1] != key) return;
    const bool is_ent = key < a.E;
    const int D = is_ent ? a.m.ent_dim : a.m.rel_dim;
    const int parts = (is_ent ? a.ce : a.cr) / D;
    for (int p = 0; p < parts; p++) {
        float acc[N];
#pragma unroll
        for (int q = 0; q < N; q++) acc[q] = 0.f;
        sum_rows<VW, NV, COH>(a, lo, lo + PCH, is_ent, D, p, lane, acc, false);
        Frag<VW, NV> f;
#pragma unroll
        for (int q = 0; q < N; q++) f.v[q] = acc[q];
        f.store(partial + (i64)w * a.pcols + p * D, D, lane);
    }
}
template <int VW, int NV>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) prereduce_kernel(UpdArgs a, float *partial) {
    prereduce_body<VW, NV, false>(a, partial, blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5), threadIdx.x & 31);
}

// SGD: row -= lr * sum of its gradient rows (GradientDescentOptimizer's sparse apply: duplicates accumulate).
// One warp per sorted position (only segment heads work), or — when the batch has more gradient rows than the
// tables have rows — one warp per table row.
// PDL: the kernel was launched with programmatic stream serialization — the row map and the table row (neither written by
// the grad kernel) are requested first, griddepcontrol.wait then guarantees the gradient rows.
template <int VW, int NV, bool COH, bool PDL = false>
__device__ __forceinline__ void sgd_body(const UpdArgs &a, i32 w, int lane) {
    constexpr int N = VW * NV;
    i32 key;
    int4 seg;
    if (a.by_row) {
        key = w;
        if (key >= a.key_limit) return;
        seg = __ldg(a.rowhead + key);
        if (seg.x < 0) return;
    } else {
        if (w >= a.n) return;
        key = a.skeys[w];
        if (key >= a.key_limit || (w > 0 && a.skeys[w - 1] == key)) return;
        seg = __ldg(a.rowhead + key);
    }
    const bool is_ent = key < a.E;
    const int D = is_ent ? a.m.ent_dim : a.m.rel_dim;
    const i32 row = is_ent ? key : key - a.E;
    const int parts = (is_ent ? a.ce : a.cr) / D;
    for (int p = 0; p < parts; p++) {
        float *tab = is_ent ? (p ? a.m.ent_aux : a.m.ent) : (p ? a.m.rel_aux : a.m.rel);
        const i64 off = (i64)row * D;
        Frag<VW, NV> x;
        x.template load<COH>(tab + off, D, lane);         // independent of the gradient rows: in flight first
        if (PDL) pdl_wait();
        float g[N];
#pragma unroll
        for (int q = 0; q < N; q++) g[q] = 0.f;
        seg_sum_map<VW, NV, COH>(a, seg, is_ent, D, p, lane, g);
#pragma unroll
        for (int q = 0; q < N; q++) x.v[q] -= a.hp.lr * g[q];
        x.store(tab + off, D, lane);
    }
}
template <int VW, int NV>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) sgd_kernel(UpdArgs a) {
    pdl_launch_dependents();                               // next step's grad kernel may fetch its batch ids
    if ((i32)blockIdx.x >= a.work_blocks) { pdl_wait(); loss_block(a, (i32)blockIdx.x - a.work_blocks, (i32)blockDim.x); return; }
    if (a.bad) { pdl_wait(); if (upd_bad(a)) return; }     // host-batch steps: the flag is final once the grad launch has completed
    sgd_body<VW, NV, false, true>(a, blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5), threadIdx.x & 31);
}

// TF1 AdamOptimizer._apply_sparse_shared: m and v decay over the WHOLE variable and every row moves
// each step; only the (1-beta) g terms are sparse.  One flat, vectorised pass over all tables:
// thread -> one VW-wide vector of one row; rows touched this step (rowhead >= 0) first sum their
// gradient rows in sorted slot order (every thread of the row walks the same segment, reading its
// own columns: coalesced), then the same Adam arithmetic runs everywhere.
//   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; x <- x - lr_t m / (sqrt(v) + eps)
template <int VW>
__global__ void __launch_bounds__(256, 6) adam_kernel(UpdArgs a) {
    pdl_launch_dependents();
    if ((i32)blockIdx.x >= a.work_blocks) { pdl_wait(); loss_block(a, (i32)blockIdx.x - a.work_blocks, (i32)blockDim.x); return; }
    if (a.bad) { pdl_wait(); if (upd_bad(a)) return; }
    typedef typename VecT<VW>::T V;
    const i64 total = a.tab[a.ntab - 1].vec_end;
    const i64 stride = (i64)a.work_blocks * blockDim.x;
    const float b1 = a.hp.beta1, b2 = a.hp.beta2, lr = a.hp.lr, eps = a.hp.eps;
    for (i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
        int t = 0;
        while (v >= a.tab[t].vec_end) t++;
        const DenseTab &T = a.tab[t];
        const unsigned lv = (unsigned)(v - (t ? a.tab[t - 1].vec_end : 0));     // < 2^31 vectors per table
        const unsigned vpr = (unsigned)T.D / VW;                                // vectors per row
        const unsigned row = lv / vpr, col = (lv - row * vpr) * VW;
        const i64 e = (i64)lv * VW;
        // issue every independent load before the first dependent use
#ifdef EXP_ADAM_NOGRAD
        const int4 seg = make_int4(-1, 0, 0, 0);
#else
        const int4 seg = __ldg(a.rowhead + T.key_off + row);
#endif
        V xv = *reinterpret_cast<const V *>(T.x + e), mv = *reinterpret_cast<const V *>(T.m + e), vv = *reinterpret_cast<const V *>(T.v + e);
        pdl_wait();                                        // gradient rows of this step are complete from here on
        float g[VW];
#pragma unroll
        for (int q = 0; q < VW; q++) g[q] = 0.f;
        const bool long_seg = seg.x >= 0 && a.hub && seg.y - seg.x > PCH && (seg.x + PCH - 1) / PCH < seg.y / PCH;
        if (seg.x >= 0 && !long_seg) {
            const float *gbase = T.grad + T.part * T.D + col;
            const i32 cnt = seg.y - seg.x;
            // the first two contributions come straight from the row map: no perm[] hop for the common case
            const V g0 = __ldg(reinterpret_cast<const V *>(gbase + (i64)(seg.z - T.slot_off) * T.cols));
            V g1 = g0;
            if (cnt > 1) g1 = __ldg(reinterpret_cast<const V *>(gbase + (i64)(seg.w - T.slot_off) * T.cols));
            const float *p0 = reinterpret_cast<const float *>(&g0), *p1 = reinterpret_cast<const float *>(&g1);
#pragma unroll
            for (int q = 0; q < VW; q++) { g[q] += p0[q]; if (cnt > 1) g[q] += p1[q]; }
            for (i32 j = seg.x + 2; j < seg.y; j++) {
                const V gj = __ldg(reinterpret_cast<const V *>(gbase + (i64)(__ldg(a.perm + j) - T.slot_off) * T.cols));
                const float *pj = reinterpret_cast<const float *>(&gj);
#pragma unroll
                for (int q = 0; q < VW; q++) g[q] += pj[q];
            }
        }
        if (long_seg) {                                    // hub row: fringe rows + pre-reduced interior blocks, ascending order
            const float *gbase = T.grad + T.part * T.D + col, *pbase = a.partial + T.part * T.D + col;
            const i32 b0 = (seg.x + PCH - 1) / PCH, b1 = seg.y / PCH;
            auto add_raw = [&](i32 lo, i32 hi) {
                for (i32 j = lo; j < hi; j++) {
                    const V gj = __ldg(reinterpret_cast<const V *>(gbase + (i64)(__ldg(a.perm + j) - T.slot_off) * T.cols));
                    const float *pj = reinterpret_cast<const float *>(&gj);
#pragma unroll
                    for (int q = 0; q < VW; q++) g[q] += pj[q];
                }
            };
            add_raw(seg.x, b0 * PCH);
            for (i32 b = b0; b < b1; b++) {
                const V gj = __ldg(reinterpret_cast<const V *>(pbase + (i64)b * a.pcols));
                const float *pj = reinterpret_cast<const float *>(&gj);
#pragma unroll
                for (int q = 0; q < VW; q++) g[q] += pj[q];
            }
            add_raw(b1 * PCH, seg.y);
        }
        float *xs = reinterpret_cast<float *>(&xv), *ms = reinterpret_cast<float *>(&mv), *vs = reinterpret_cast<float *>(&vv);
#pragma unroll
        for (int q = 0; q < VW; q++) adam_elem(xs[q], ms[q], vs[q], g[q], b1, b2, 1.f - b1, 1.f - b2, lr, eps);
        *reinterpret_cast<V *>(T.x + e) = xv; *reinterpret_cast<V *>(T.m + e) = mv; *reinterpret_cast<V *>(T.v + e) = vv;
    }
}

// ---- Lean form of the register kernel: one 256-vector tile of ONE table per CTA, so the table descriptor is
// block-uniform (uniform-datapath loads instead of a per-thread search through the kernel parameters), all index math
// is 32-bit and the row / column split is a multiply-high by a host-computed reciprocal instead of an integer
// division.  ncu's source page of adam_kernel attributes 65 % of its 336 instructions per warp to exactly that
// integer / control / parameter-load overhead; the pass is issue-bound in disguise (each wave loads, then all warps
// compete for the issue slots), so fewer instructions is what shortens it.
#ifndef ADAM_TILE_MIN_BLOCKS
#define ADAM_TILE_MIN_BLOCKS 4     // 62 registers, no spills: 14.7 us vs 18.8 us at 6 CTAs/SM with spilled vectors (TransH D=100 FB15K)
#endif
// Sum of the gradient rows of one table row (row-map entry `seg`) for the VW columns starting at `col`, in sorted slot
// order.  COH: coherent (L2) loads, for the persistent chunk kernel.
template <int VW, bool COH>
__device__ __forceinline__ void adam_gsum(const UpdArgs &a, const DenseTab &T, const int4 seg, unsigned col, float *g) {
    typedef typename VecT<VW>::T V;
    if (seg.x < 0) return;
    const float *gbase = T.grad + (T.part * T.D + (i32)col);
    const i32 cols = T.cols, soff = T.slot_off, cnt = seg.y - seg.x;
    auto ld = [&](const float *p) { return COH ? __ldcg(reinterpret_cast<const V *>(p)) : __ldg(reinterpret_cast<const V *>(p)); };
    auto add = [&](const V &w) {
        const float *pw = reinterpret_cast<const float *>(&w);
#pragma unroll
        for (int q = 0; q < VW; q++) g[q] += pw[q];
    };
    auto add_raw = [&](i32 lo, i32 hi) {
        for (i32 j = lo; j < hi; j++) add(ld(gbase + (size_t)(__ldg(a.perm + j) - soff) * cols));
    };
    const bool long_seg = a.hub && cnt > PCH && (seg.x + PCH - 1) / PCH < seg.y / PCH;
    if (!long_seg) {
        // the first two contributions come straight from the row map: no perm[] hop for the common case
        const V g0 = ld(gbase + (size_t)(seg.z - soff) * cols);
        if (cnt > 1) {
            const V g1 = ld(gbase + (size_t)(seg.w - soff) * cols);
            add(g0); add(g1);
            add_raw(seg.x + 2, seg.y);
        } else add(g0);
    } else {                                           // hub row: fringe rows + pre-reduced interior blocks, ascending order
        const i32 b0 = (seg.x + PCH - 1) / PCH, b1 = seg.y / PCH;
        const float *pb = a.partial + (T.part * T.D + (i32)col);
        add_raw(seg.x, b0 * PCH);
        for (i32 b = b0; b < b1; b++) add(ld(pb + (size_t)b * a.pcols));
        add_raw(b1 * PCH, seg.y);
    }
}
// VPT vectors per thread: the CTA owns a super tile of 256 * VPT consecutive vectors of one table and every thread has the
// row-map entries and x / m / v of ALL its vectors in flight before the first dependent gradient load — fewer, fatter
// waves instead of three thin ones (the pass is latency-, not bandwidth-bound at these table sizes).
template <int VW, int VPT>
__global__ void __launch_bounds__(256, VPT == 1 ? ADAM_TILE_MIN_BLOCKS : (VPT == 2 ? 3 : 2)) adam_tile_kernel(UpdArgs a) {
    pdl_launch_dependents();                               // next step's grad kernel may prefetch its batch ids
    // the loss blocks come FIRST in the grid: the step's loss is out a few microseconds after the grad kernel has
    // finished, so a caller waiting for it (okb_wait_word) gets on with the next batch while the tables are updated
    const i32 bid = (i32)blockIdx.x - a.loss_blocks;
    if (bid < 0) { pdl_wait(); loss_block(a, (i32)blockIdx.x, (i32)blockDim.x); return; }
    typedef typename VecT<VW>::T V;
    int t = 0;
    while (bid >= a.tab[t].sblk_end) t++;                  // block-uniform
    const DenseTab &T = a.tab[t];
    const unsigned nvec = (unsigned)(T.vec_end - (t ? a.tab[t - 1].vec_end : 0));
    const unsigned lv0 = ((unsigned)bid - (unsigned)(t ? a.tab[t - 1].sblk_end : 0)) * (256u * VPT) + threadIdx.x;
    if (lv0 >= nvec) return;
    const unsigned vpr = (unsigned)T.D / VW;
    unsigned col[VPT];
    int4 seg[VPT];
    V xv[VPT], mv[VPT], vv[VPT];
#pragma unroll
    for (int u = 0; u < VPT; u++) {
        const unsigned lv = lv0 + 256u * u;
        seg[u] = make_int4(-1, 0, 0, 0);
        if (lv < nvec) {
            const unsigned row = T.magic ? (__umulhi(lv, T.magic) >> T.shift) : (lv >> T.shift);
            col[u] = (lv - row * vpr) * VW;
            const size_t e = (size_t)lv * VW;
            seg[u] = __ldg(a.rowhead + T.key_off + row);
            xv[u] = *reinterpret_cast<const V *>(T.x + e); mv[u] = *reinterpret_cast<const V *>(T.m + e); vv[u] = *reinterpret_cast<const V *>(T.v + e);
        }
    }
    pdl_wait();                                            // gradient rows of this step are complete from here on
    // host-batch steps only: the flag may be raised by the verification blocks of THIS step's grad launch, so it is read
    // after the wait (narrow_kernel / grad_verify_block)
    if (upd_bad(a)) return;
    const float b1 = a.hp.beta1, b2 = a.hp.beta2, lr = a.hp.lr, eps = a.hp.eps, c1 = 1.f - b1, c2 = 1.f - b2;
#pragma unroll
    for (int u = 0; u < VPT; u++) {
        const unsigned lv = lv0 + 256u * u;
        if (lv >= nvec) break;
        float g[VW];
#pragma unroll
        for (int q = 0; q < VW; q++) g[q] = 0.f;
        adam_gsum<VW, false>(a, T, seg[u], col[u], g);
        float *xs = reinterpret_cast<float *>(&xv[u]), *ms = reinterpret_cast<float *>(&mv[u]), *vs = reinterpret_cast<float *>(&vv[u]);
#pragma unroll
        for (int q = 0; q < VW; q++) adam_elem(xs[q], ms[q], vs[q], g[q], b1, b2, c1, c2, lr, eps);
        const size_t e = (size_t)lv * VW;
        *reinterpret_cast<V *>(T.x + e) = xv[u]; *reinterpret_cast<V *>(T.m + e) = mv[u]; *reinterpret_cast<V *>(T.v + e) = vv[u];
    }
}


// ------------------------------------------------------------------------------------------ host-side builders (train.cu)
int okb_fill_update(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, const float *gent, const float *grel,
                    const float *loss_terms, float *loss_out, int vw, cudaStream_t s, UpdArgs &a, i32 &blk, bool &lean);
void okb_fill_grad(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step, INT b_lo, INT b_hi, INT slot_base, float *gent,
                   float *grel, float *loss_terms, GradArgs &a);
int okb_grad_wpp(okb_ctx *c);
bool okb_pick_layout(int D, int &vw, int &nv);
// chunk.cu: n consecutive planned steps as ONE persistent cooperative kernel; returns -1 if this configuration is not
// covered (the caller then runs the per-phase kernels), 0 on success, else an OKB_ERR_* code
int okb_chunk_kernel_steps(okb_ctx *c, const okb_model *m, const okb_hyper *hp, INT step_lo, INT n, float *loss_out, cudaStream_t s);
