// TransR candidate projection on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Link-prediction evaluation projects EVERY entity by the relation's matrix: P = Ent[E,De] . M_r[De,Dr]
// per relation group (TransR.py:77-87 with one relation per query batch) — the one dense contraction of
// the hot path (2*E*De*Dr flop per relation; FB15K: 0.3 GFLOP x 1,345 relations).  One CTA computes a
// 128-entity tile:
//   * A = 128 entity rows, B = M_r^T, both staged by the CTA into shared memory in the canonical
//     no-swizzle K-major core-matrix layout (8 rows x 16 bytes per core matrix);
//   * fp32 parity needs more than TF32's 10 mantissa bits, so every operand is split x = hi + lo with hi, lo
//     TF32-representable, and three MMAs accumulate hi.hi + lo.hi + hi.lo in fp32 (3xTF32, ~2^-21 relative);
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=Dr rounded to 16, K=8 per
//     instruction), the accumulator lives in TMEM (128 lanes x N columns), completion arrives on an mbarrier
//     through tcgen05.commit;
//   * epilogue: each thread owns one TMEM lane = one entity; tcgen05.ld hands it its row, it normalises the
//     row (tf.nn.l2_normalize) and writes it TRANSPOSED into the candidate table [Dr][E_pad] the ranking
//     kernel consumes — lanes are consecutive entities, so every store is a coalesced 128-byte line.
// Scores computed from these candidates differ from the canonical sequential-fp32 order in the last bits,
// so this path is opt-in (okb_set_flag(OKB_FLAG_TRANSR_TC, 1)); the default TransR ranking stays bit-exact.
#include <algorithm>

#include "okb_internal.h"

#define TC_M 128
#define EPS_NORM 1e-12f

__device__ __forceinline__ unsigned tc_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor, no swizzle, K-major:
//   bits [0,14) start address >> 4 | [16,30) leading-dimension byte offset >> 4 (core matrix -> next along K)
//   [32,46) stride-dimension byte offset >> 4 (8-row group -> next along M/N) | [46,48) version = 1 | [61,64) swizzle = 0
__device__ __forceinline__ unsigned long long tc_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;
    return d;
}

struct TcArgs {
    const float *ent;          // [E][De]
    const float *mats;         // transfer_matrix [R][De*Dr]
    const i32 *grp_rel;        // relation of each group
    float *out;                // [G][Dr][ncol]
    i32 E, De, Dr, Kp, Np, j0, ncol;
};

#define TC_THREADS 512        // 16 warps: all of them stage operands and drain TMEM; one thread issues the MMAs

__device__ __forceinline__ float4 tf32_hi(const float4 v) {
    float4 h;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    return h;
}

__global__ void __launch_bounds__(TC_THREADS) transr_project_tc_kernel(TcArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    __shared__ float ss_part[4][TC_M];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int De = a.De, Dr = a.Dr, Kp = a.Kp, Np = a.Np, kc_n = Kp >> 2;      // kc_n: 16-byte chunks along K
    const unsigned LBO = 128, SBO = (unsigned)kc_n * 128;
    const size_t a_bytes = (size_t)(TC_M / 8) * SBO, b_bytes = (size_t)(Np / 8) * SBO;
    unsigned char *A_hi = smraw, *A_lo = A_hi + a_bytes, *B_hi = A_lo + a_bytes, *B_lo = B_hi + b_bytes;

    const i32 g = blockIdx.y;
    if (warp == 0) {                                       // TMEM: 128 lanes x 128 columns of fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_s)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- stage B = M_r^T once per CTA: B[n][k] = M_r[k][n], split hi / lo
    //      core-matrix layout: (row, k) -> (row/8)*SBO + (k/4)*128 + (row%8)*16 + (k%4)*4
    const float *M = a.mats + (i64)a.grp_rel[g] * De * Dr;
    for (int c = tid; c < Np * kc_n; c += TC_THREADS) {
        const int n = c % Np, kc = c / Np;
        float4 v;
        float *vs = reinterpret_cast<float *>(&v);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = kc * 4 + q;
            vs[q] = (n < Dr && k < De) ? __ldg(M + (i64)k * Dr + n) : 0.f;
        }
        const float4 hi = tf32_hi(v);
        const size_t off = (size_t)(n >> 3) * SBO + (size_t)kc * 128 + (n & 7) * 16;
        *reinterpret_cast<float4 *>(B_hi + off) = hi;
        *reinterpret_cast<float4 *>(B_lo + off) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
    }
    unsigned phase = 0;
    const int ntiles = a.ncol / TC_M;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i32 col0 = tile * TC_M;
        // ---- stage A: 128 entity rows of this tile, split hi / lo
        for (int c = tid; c < TC_M * kc_n; c += TC_THREADS) {
            const int m = c % TC_M, kc = c / TC_M;
            const i32 j = a.j0 + col0 + m;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < a.E && kc * 4 < De) v = __ldg(reinterpret_cast<const float4 *>(a.ent + (i64)j * De) + kc);
            const float4 hi = tf32_hi(v);
            const size_t off = (size_t)(m >> 3) * SBO + (size_t)kc * 128 + (m & 7) * 16;
            *reinterpret_cast<float4 *>(A_hi + off) = hi;
            *reinterpret_cast<float4 *>(A_lo + off) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        }
        // generic-proxy writes -> visible to the async proxy the tensor core reads shared memory through
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned tmem = tmem_base_s;

        if (tid == 0) {
            // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
            // N >> 3 at bits 17-22, M >> 4 at bits 24-28
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(Np >> 3) << 17) | ((unsigned)(TC_M >> 4) << 24);
            const unsigned a_hi = tc_smem_u32(A_hi), a_lo = tc_smem_u32(A_lo), b_hi = tc_smem_u32(B_hi), b_lo = tc_smem_u32(B_lo);
            unsigned accum = 0;
            for (int pass = 0; pass < 3; pass++) {         // hi.hi, lo.hi, hi.lo
                const unsigned ab = pass == 1 ? a_lo : a_hi, bb = pass == 2 ? b_lo : b_hi;
                for (int ks = 0; ks < Kp / 8; ks++) {      // one MMA covers K = 8 = two core matrices
                    const unsigned long long da = tc_desc(ab + ks * 256, LBO, SBO), db = tc_desc(bb + ks * 256, LBO, SBO);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                        "l"(da), "l"(db), "r"(idesc), "r"(accum)
                        : "memory");
                    accum = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&bar)) : "memory");
        }
        {   // wait for the accumulator
            unsigned done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(tc_smem_u32(&bar)), "r"(phase)
                    : "memory");
            }
            phase ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue: a warp reads TMEM lanes 32*(warp%4).. (its hardware window); the four warps sharing a
        //      window split the 16-column chunks.  thread <-> lane <-> entity.
        const int wq = warp & 3, wg = warp >> 2, row = wq * 32 + lane;
        const unsigned taddr = tmem + ((unsigned)(wq * 32) << 16);
        const i32 j = a.j0 + col0 + row, col = col0 + row;
        float ss = 0.f;
        for (int c0 = wg * 16; c0 < Np; c0 += 64) {
            unsigned r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; q++)
                if (c0 + q < Dr) { const float p = __uint_as_float(r[q]); ss = __fadd_rn(ss, __fmul_rn(p, p)); }
        }
        ss_part[wg][row] = ss;
        __syncthreads();
        ss = __fadd_rn(__fadd_rn(ss_part[0][row], ss_part[1][row]), __fadd_rn(ss_part[2][row], ss_part[3][row]));
        const float inv = __fdiv_rn(1.0f, __fsqrt_rn(ss > EPS_NORM ? ss : EPS_NORM));
        for (int c0 = wg * 16; c0 < Np; c0 += 64) {
            unsigned r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; q++)
                if (c0 + q < Dr && col < a.ncol)
                    a.out[((i64)g * Dr + c0 + q) * a.ncol + col] = j < a.E ? __fmul_rn(__uint_as_float(r[q]), inv) : 0.f;
        }
        // TMEM and the A buffers are rewritten by the next tile: every warp must be done reading them
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(128));
}

// Returns 1 if the shape is supported by the tensor-core path (else the caller uses the canonical SIMT kernel).
int okb_transr_tc_supported(const okb_model *m) {
    const int Kp = (m->ent_dim + 7) & ~7, Np = (m->rel_dim + 15) & ~15;
    if (m->ent_dim % 4 || Np > 128) return 0;
    const size_t smem = (size_t)(TC_M / 8 + Np / 8) * (Kp / 4) * 128 * 2;
    return smem <= 220 * 1024;
}

int okb_transr_project_tc(okb_ctx *c, const okb_model *m, const i32 *d_grp_rel, i64 G, float *out, i64 j0, i64 ncol, cudaStream_t s) {
    TcArgs a;
    a.ent = m->ent; a.mats = m->rel_aux; a.grp_rel = d_grp_rel; a.out = out;
    a.E = (i32)c->E; a.De = m->ent_dim; a.Dr = m->rel_dim;
    a.Kp = (m->ent_dim + 7) & ~7; a.Np = (m->rel_dim + 15) & ~15;
    a.j0 = (i32)j0; a.ncol = (i32)ncol;
    const size_t smem = (size_t)(TC_M / 8 + a.Np / 8) * (a.Kp / 4) * 128 * 2;
    OKB_CUDA(c, okb_smem_optin(c, transr_project_tc_kernel, smem));
    // each CTA keeps M_r staged and walks every gx-th entity tile of its group; enough CTAs to fill the chip twice over
    const unsigned ntiles = (unsigned)(ncol / TC_M);
    const unsigned gx = (unsigned)std::max<i64>(1, std::min<i64>(ntiles, (2 * okb_sms(c) + G - 1) / G));
    transr_project_tc_kernel<<<dim3(gx, (unsigned)G), TC_THREADS, smem, s>>>(a);
    OKB_LAUNCHED(1);
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}
