// Stable LSD radix sort of (key, slot) pairs, 8 bits per pass — the "sort" half of the
// deterministic sort-then-segmented-reduce gradient scatter.  Keys are table row ids, values the
// gradient-row slots; stability keeps equal keys in slot order, which fixes the fp32 summation
// order of duplicate rows.  Four kernels per pass (tile histogram, two-level scan, stable scatter);
// integer work only, off the critical path when batches are planned ahead.
#include "okb_internal.h"

#define RADIX_BITS 8
#define RADIX 256
#define SORT_THREADS 256
// items per thread: 8 for large sorts; 2 when a whole sort is only a few hundred tiles (a chunk of 20 planned steps of 24 k
// pairs = 240 tiles of 2,048 leaves the four kernels of a pass latency-bound on 1.6 CTAs per SM)

// Segmented form: the input is `nseg` independent segments of `seg` items each (one per planned train step),
// every segment is cut into T tiles of SORT_TILE items, and the tile histograms are laid out
// [segment][digit][tile].  An exclusive scan of that array in storage order, plus segment * seg, is then each
// (segment, digit, tile)'s output position — segments never mix, so one sort call plans a whole chunk of steps
// with the pass count of ONE step's key range.
template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS) radix_hist(const i32 *__restrict__ keys, i32 *__restrict__ ghist,
                                                           i32 seg, i32 T, i32 shift) {
    constexpr int SORT_ITEMS = ITEMS, SORT_TILE = SORT_THREADS * ITEMS;
    __shared__ i32 h[RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    const i32 c = blockIdx.x / T, t = blockIdx.x - c * T;
    const i32 base = t * SORT_TILE;
    const i32 *k = keys + (i64)c * seg;
#pragma unroll
    for (i32 j = 0; j < SORT_ITEMS; j++) {
        const i32 i = base + j * SORT_THREADS + threadIdx.x;
        if (i < seg) atomicAdd(&h[(k[i] >> shift) & (RADIX - 1)], 1);
    }
    __syncthreads();
    ghist[((i64)c * RADIX + threadIdx.x) * T + t] = h[threadIdx.x];
}

// Exclusive scan of every segment's RADIX*T counters, two launches, no cross-block waiting:
//   scan_reduce: block (j, c) sums its SCAN_TILE-wide slice of segment c's counters
//   scan_apply : block (j, c) adds up the sums of the slices before it (<= a few hundred), scans its slice,
//                and adds c * seg (segments have a fixed length, so their bases need no scan)
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
__device__ __forceinline__ i32 block_sum_256(i32 x, i32 *sh) {
#pragma unroll
    for (i32 o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    i32 t = 0;
#pragma unroll
    for (i32 w = 0; w < SCAN_THREADS / 32; w++) t += sh[w];
    __syncthreads();
    return t;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce(const i32 *__restrict__ a, i32 *__restrict__ bsum, i32 len) {
    __shared__ i32 sh[SCAN_THREADS / 32];
    const i32 *p = a + (i64)blockIdx.y * len;
    const i32 base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    i32 x = 0;
#pragma unroll
    for (i32 j = 0; j < SCAN_ITEMS; j++) if (base + j < len) x += p[base + j];
    const i32 t = block_sum_256(x, sh);
    if (threadIdx.x == 0) bsum[blockIdx.y * gridDim.x + blockIdx.x] = t;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(i32 *__restrict__ a, const i32 *__restrict__ bsum, i32 len, i32 seg) {
    __shared__ i32 sh[SCAN_THREADS / 32];
    __shared__ i32 wsum[SCAN_THREADS / 32];
    i32 *p = a + (i64)blockIdx.y * len;
    i32 pre = 0;
    for (i32 j = threadIdx.x; j < (i32)blockIdx.x; j += SCAN_THREADS) pre += bsum[blockIdx.y * gridDim.x + j];
    const i32 offset = block_sum_256(pre, sh) + (i32)blockIdx.y * seg;
    const i32 base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    i32 v[SCAN_ITEMS], x = 0;
#pragma unroll
    for (i32 j = 0; j < SCAN_ITEMS; j++) { v[j] = base + j < len ? p[base + j] : 0; x += v[j]; }
    const i32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    i32 inc = x;
#pragma unroll
    for (i32 o = 1; o < 32; o <<= 1) { const i32 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    i32 run = offset + inc - x;
    for (i32 ww = 0; ww < w; ww++) run += wsum[ww];
#pragma unroll
    for (i32 j = 0; j < SCAN_ITEMS; j++) { if (base + j < len) p[base + j] = run; run += v[j]; }
}

// Stable scatter.  Item order inside a tile: warp-contiguous chunks of 256, striped across lanes
// (item = warp*256 + j*32 + lane), so ranking rounds j = 0..7 visit items in ascending index.
template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS) radix_scatter(const i32 *__restrict__ keys, const i32 *__restrict__ vals,
                                                              i32 *__restrict__ keys_out, i32 *__restrict__ vals_out,
                                                              const i32 *__restrict__ gscan, i32 seg, i32 T, i32 shift,
                                                              i32 iota_vals) {
    constexpr int SORT_ITEMS = ITEMS, SORT_TILE = SORT_THREADS * ITEMS;
    __shared__ i32 wh[SORT_THREADS / 32][RADIX];
    const i32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (i32 i = threadIdx.x; i < (SORT_THREADS / 32) * RADIX; i += SORT_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    const i32 c = blockIdx.x / T, t = blockIdx.x - c * T;
    const i64 sb = (i64)c * seg;                           // segment base: inputs are read from it, outputs are absolute
    const i32 base = t * SORT_TILE + w * (32 * SORT_ITEMS);
    i32 key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (i32 j = 0; j < SORT_ITEMS; j++) {
        const i32 i = base + j * 32 + lane;
        const bool ok = i < seg;
        key[j] = ok ? keys[sb + i] : 0x7fffffff;
        val[j] = ok ? (iota_vals ? i : vals[sb + i]) : 0;
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
        i32 run = gscan[((i64)c * RADIX + d) * T + t];
#pragma unroll
        for (i32 ww = 0; ww < SORT_THREADS / 32; ww++) { const i32 cnt = wh[ww][d]; wh[ww][d] = run; run += cnt; }
    }
    __syncthreads();
#pragma unroll
    for (i32 j = 0; j < SORT_ITEMS; j++) {
        const i32 i = base + j * 32 + lane;
        if (i < seg) {
            const i32 d = (key[j] >> shift) & (RADIX - 1);
            const i32 dst = wh[w][d] + rank[j];
            keys_out[dst] = key[j];
            vals_out[dst] = val[j];
        }
    }
}

// `nseg` segments of `seg` keys each: keys_out = every segment sorted ascending (stable), perm_out = position of each
// sorted item inside ITS segment.  `keys` is preserved.  bits = width of one segment's key range.
int okb_sort_pairs_seg(okb_ctx *c, const i32 *keys, i32 *keys_out, i32 *perm_out, i64 seg, i64 nseg, int bits, cudaStream_t s) {
    const i64 n = seg * nseg;
    if (n <= 0) return 0;
    const bool small = n <= 1000000;      // measured: 20 steps x 24 k pairs plan in 50.4 us with 512-item tiles, 58.4 us with 2,048-item tiles
    const i64 SORT_TILE = SORT_THREADS * (small ? 2 : 8);
    const i32 T = (i32)((seg + SORT_TILE - 1) / SORT_TILE);
    const i64 nblk = (i64)T * nseg, len = (i64)RADIX * T;
    const i32 sblk = (i32)((len + SCAN_TILE - 1) / SCAN_TILE);
    if (n > 0x7fffffffLL || nblk > 0x7fffffffLL || nseg > 65535) OKB_FAIL(c, OKB_ERR_ARG, "sort too large");
    const int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    // ping-pong buffers: tmp holds (keys, vals)
    if (c->sort_tmp.ensure(sizeof(i32) * 2 * n)) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (sort)");
    if (c->hist.ensure(sizeof(i32) * (RADIX * nblk + (i64)sblk * nseg))) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory (sort)");
    i32 *tk = c->sort_tmp.as<i32>(), *tv = tk + n, *gh = c->hist.as<i32>(), *bsum = gh + RADIX * nblk;
    // choose the starting side so that the last pass lands in (keys_out, perm_out)
    const i32 *src_k = keys, *src_v = nullptr;
    i32 *dst_k = (passes & 1) ? keys_out : tk, *dst_v = (passes & 1) ? perm_out : tv;
    for (int p = 0; p < passes; p++) {
        const i32 shift = p * RADIX_BITS;
        if (small) radix_hist<2><<<(unsigned)nblk, SORT_THREADS, 0, s>>>(src_k, gh, (i32)seg, T, shift);
        else radix_hist<8><<<(unsigned)nblk, SORT_THREADS, 0, s>>>(src_k, gh, (i32)seg, T, shift);
        scan_reduce<<<dim3(sblk, (unsigned)nseg), SCAN_THREADS, 0, s>>>(gh, bsum, (i32)len);
        scan_apply<<<dim3(sblk, (unsigned)nseg), SCAN_THREADS, 0, s>>>(gh, bsum, (i32)len, (i32)seg);
        if (small) radix_scatter<2><<<(unsigned)nblk, SORT_THREADS, 0, s>>>(src_k, src_v, dst_k, dst_v, gh, (i32)seg, T, shift, p == 0);
        else radix_scatter<8><<<(unsigned)nblk, SORT_THREADS, 0, s>>>(src_k, src_v, dst_k, dst_v, gh, (i32)seg, T, shift, p == 0);
        OKB_LAUNCHED(4);
        src_k = dst_k; src_v = dst_v;
        if (dst_k == keys_out) { dst_k = tk; dst_v = tv; } else { dst_k = keys_out; dst_v = perm_out; }
    }
    OKB_CUDA(c, cudaGetLastError());
    return 0;
}
// keys[n] -> keys_out[n] ascending, perm_out[n] = original positions (stable).  `keys` is preserved.
int okb_sort_pairs(okb_ctx *c, const i32 *keys, i32 *keys_out, i32 *perm_out, i64 n, int bits, cudaStream_t s) {
    return okb_sort_pairs_seg(c, keys, keys_out, perm_out, n, 1, bits, s);
}
