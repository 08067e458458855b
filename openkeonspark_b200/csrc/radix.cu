// Stable LSD radix sort of (key, slot) pairs, 8 bits per pass — the "sort" half of the
// deterministic sort-then-segmented-reduce gradient scatter.  Keys are table row ids, values the
// gradient-row slots; stability keeps equal keys in slot order, which fixes the fp32 summation
// order of duplicate rows.  Three kernels per pass (tile histogram, scan, stable scatter);
// integer work only, off the critical path when batches are planned ahead.
#include "okb_internal.h"

#define RADIX_BITS 8
#define RADIX 256
#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)

__global__ void __launch_bounds__(SORT_THREADS) radix_hist(const i32 *__restrict__ keys, i32 *__restrict__ ghist,
                                                           i32 n, i32 shift, i32 nblk) {
    __shared__ i32 h[RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    const i32 base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (i32 j = 0; j < SORT_ITEMS; j++) {
        const i32 i = base + j * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & (RADIX - 1)], 1);
    }
    __syncthreads();
    ghist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `len` ints by ONE block (len = 256 * number of tiles)
__global__ void __launch_bounds__(1024) radix_scan(i32 *__restrict__ a, i32 len) {
    __shared__ i32 warp_sum[32];
    __shared__ i32 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const i32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (i32 base = 0; base < len; base += 1024) {
        const i32 i = base + threadIdx.x;
        const i32 v = i < len ? a[i] : 0;
        i32 x = v;
#pragma unroll
        for (i32 o = 1; o < 32; o <<= 1) { i32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        if (w == 0) {
            i32 s = warp_sum[lane];
#pragma unroll
            for (i32 o = 1; o < 32; o <<= 1) { i32 y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            warp_sum[lane] = s;
        }
        __syncthreads();
        const i32 before = carry + (w ? warp_sum[w - 1] : 0) + x - v;
        if (i < len) a[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
}

// Stable scatter.  Item order inside a tile: warp-contiguous chunks of 256, striped across lanes
// (item = warp*256 + j*32 + lane), so ranking rounds j = 0..7 visit items in ascending index.
__global__ void __launch_bounds__(SORT_THREADS) radix_scatter(const i32 *__restrict__ keys, const i32 *__restrict__ vals,
                                                              i32 *__restrict__ keys_out, i32 *__restrict__ vals_out,
                                                              const i32 *__restrict__ gscan, i32 n, i32 shift, i32 nblk,
                                                              i32 iota_vals) {
    __shared__ i32 wh[SORT_THREADS / 32][RADIX];
    const i32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (i32 i = threadIdx.x; i < (SORT_THREADS / 32) * RADIX; i += SORT_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    const i32 base = blockIdx.x * SORT_TILE + w * (32 * SORT_ITEMS);
    i32 key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (i32 j = 0; j < SORT_ITEMS; j++) {
        const i32 i = base + j * 32 + lane;
        const bool ok = i < n;
        key[j] = ok ? keys[i] : 0x7fffffff;
        val[j] = ok ? (iota_vals ? i : vals[i]) : 0;
        const i32 d = ok ? ((key[j] >> shift) & (RADIX - 1)) : RADIX - 1;
        const unsigned peers = __match_any_sync(0xffffffffu, ok ? d : (RADIX + lane));
        const i32 leader = __ffs(peers) - 1;
        i32 start = 0;
        if (ok && lane == leader) { start = wh[w][d]; wh[w][d] = start + __popc(peers); }
        start = __shfl_sync(0xffffffffu, start, leader);
        rank[j] = start + __popc(peers & ((1u << lane) - 1));
        __syncwarp();
    }
    __syncthreads();
    {   // per digit: exclusive scan over the warps, then add the tile's global base
        const i32 d = threadIdx.x;
        i32 run = gscan[d * nblk + blockIdx.x];
#pragma unroll
        for (i32 ww = 0; ww < SORT_THREADS / 32; ww++) { const i32 cnt = wh[ww][d]; wh[ww][d] = run; run += cnt; }
    }
    __syncthreads();
#pragma unroll
    for (i32 j = 0; j < SORT_ITEMS; j++) {
        const i32 i = base + j * 32 + lane;
        if (i < n) {
            const i32 d = (key[j] >> shift) & (RADIX - 1);
            const i32 dst = wh[w][d] + rank[j];
            keys_out[dst] = key[j];
            vals_out[dst] = val[j];
        }
    }
}

// keys[n] -> keys_out[n] ascending, perm_out[n] = original positions (stable).  `keys` is preserved.
int okb_sort_pairs(okb_ctx *c, const i32 *keys, i32 *keys_out, i32 *perm_out, i64 n, int bits, cudaStream_t s) {
    if (n <= 0) return 0;
    const i32 nblk = (i32)((n + SORT_TILE - 1) / SORT_TILE);
    const int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    // ping-pong buffers: tmp holds (keys, vals)
    if (c->sort_tmp.ensure(sizeof(i32) * 2 * n)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (sort)");
    if (c->hist.ensure(sizeof(i32) * RADIX * nblk)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (sort)");
    i32 *tk = c->sort_tmp.as<i32>(), *tv = tk + n, *gh = c->hist.as<i32>();
    // choose the starting side so that the last pass lands in (keys_out, perm_out)
    const i32 *src_k = keys, *src_v = nullptr;
    i32 *dst_k = (passes & 1) ? keys_out : tk, *dst_v = (passes & 1) ? perm_out : tv;
    for (int p = 0; p < passes; p++) {
        const i32 shift = p * RADIX_BITS;
        radix_hist<<<nblk, SORT_THREADS, 0, s>>>(src_k, gh, (i32)n, shift, nblk);
        radix_scan<<<1, 1024, 0, s>>>(gh, RADIX * nblk);
        radix_scatter<<<nblk, SORT_THREADS, 0, s>>>(src_k, src_v, dst_k, dst_v, gh, (i32)n, shift, nblk, p == 0);
        OKB_LAUNCHED(3);
        src_k = dst_k; src_v = dst_v;
        if (dst_k == keys_out) { dst_k = tk; dst_v = tv; } else { dst_k = keys_out; dst_v = perm_out; }
    }
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}
