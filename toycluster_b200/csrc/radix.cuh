// radix.cuh -- stable LSD radix sort of (64-bit key, 32-bit index), replacing the serial
// gsl_heapsort_index of sort.c:189-195 (the 240-line OpenMP quicksort below it is dead code).
//
// Sorting the upper 64 bits of the 128-bit Peano key (21 triplets + 1 bit: 2^-21 Boxsize
// per axis) orders all but pathologically close particles; k_fix_ties then orders every run
// of equal upper halves by the lower half, and equal 128-bit keys (bit-identical positions,
// which the reference's unstable heapsort leaves in heap order) by upload index.
//
// Each pass is three kernels over tiles of RS_TILE keys:
//   k_radix_hist    per-tile digit histogram -> hist[digit][tile]
//   k_radix_scan    exclusive scan of each digit's row (one block per digit) + digit totals
//   k_radix_scatter stable rank of every key inside its tile (warp match + per-warp counters),
//                   the tile staged in digit order in shared memory, then written out so that
//                   each digit's run lands at hist[digit][tile] as one segment
#pragma once
#include "common.cuh"

#define RS_BITS 8
#define RS_BINS 256
#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 8                       // keys per thread
#define RS_TILE (RS_THREADS * RS_ITEMS)  // 2048 keys per block
#define RS_WARP_SEG (32 * RS_ITEMS)      // consecutive keys owned by one warp

__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist(int n, const uint64_t *__restrict__ keys, int shift, int ntiles,
             unsigned *__restrict__ hist)
{
    __shared__ unsigned bins[RS_BINS];
    for (int b = threadIdx.x; b < RS_BINS; b += RS_THREADS) bins[b] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        const int k = base + it * RS_THREADS + threadIdx.x;
        if (k < n) atomicAdd(&bins[(keys[k] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < RS_BINS; b += RS_THREADS)
        hist[(size_t)b * ntiles + blockIdx.x] = bins[b];
}

// Exclusive scan of every digit's row hist[d][0..ntiles) in place, one block per digit, and
// the row total into bin_total[d].  (A single-block scan of the whole 256 x ntiles table was
// 0.57 ms per pass at 10 M keys -- 3.5 % of a step.)  The scatter kernel adds the exclusive
// scan of the 256 totals itself.
__global__ void __launch_bounds__(1024) k_radix_scan(int ntiles, unsigned *__restrict__ hist,
                                                     unsigned *__restrict__ bin_total)
{
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned carry;
    unsigned *data = hist + (size_t)blockIdx.x * ntiles;
    const size_t count = ntiles;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (size_t base = 0; base < count; base += 1024 * 4) {
        // each thread owns 4 consecutive values
        const size_t k0 = base + (size_t)threadIdx.x * 4;
        unsigned v[4], s = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { v[j] = (k0 + j < count) ? data[k0 + j] : 0u; s += v[j]; }
        unsigned incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        if (w == 0) {
            unsigned t = warp_tot[lane], i2 = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned u = __shfl_up_sync(FULL_MASK, i2, o);
                if (lane >= o) i2 += u;
            }
            warp_tot[lane] = i2 - t;   // exclusive over warps
        }
        __syncthreads();
        unsigned excl = carry + warp_tot[w] + incl - s;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (k0 + j < count) data[k0 + j] = excl;
            excl += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_total[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(RS_THREADS)
k_radix_scatter(int n, const uint64_t *__restrict__ keys_in, const int *__restrict__ idx_in,
                uint64_t *__restrict__ keys_out, int *__restrict__ idx_out, int shift,
                int ntiles, const unsigned *__restrict__ hist, const unsigned *__restrict__ bin_total)
{
    __shared__ unsigned cnt[RS_WARPS][RS_BINS];   // per-warp digit counters -> warp offsets
    __shared__ unsigned bin_base[RS_BINS];        // exclusive scan of the digit totals
    __shared__ unsigned local_excl[RS_BINS];      // start of each digit's run inside the tile
    __shared__ unsigned gbase[RS_BINS];           // ... and in the output
    __shared__ unsigned warp_tot[RS_WARPS];
    __shared__ uint64_t s_key[RS_TILE];
    __shared__ int s_val[RS_TILE];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < RS_WARPS * RS_BINS; b += RS_THREADS) (&cnt[0][0])[b] = 0;
    {   // RS_THREADS == RS_BINS: thread d scans bin_total[0..d)
        unsigned v = bin_total[threadIdx.x], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) bin_base[w] = incl;          // warp totals, 8 of them
        __syncthreads();
        unsigned before = 0;
        for (int ww = 0; ww < w; ww++) before += bin_base[ww];
        __syncthreads();
        bin_base[threadIdx.x] = before + incl - v;
    }
    __syncthreads();

    // Warp w owns keys [base + w*RS_WARP_SEG, +RS_WARP_SEG) in rounds of 32: tile order ==
    // (warp, round, lane) order, so ranks assigned in that order are stable.
    const int base = blockIdx.x * RS_TILE + w * RS_WARP_SEG;
    uint64_t key[RS_ITEMS];
    int val[RS_ITEMS];
    unsigned rank[RS_ITEMS];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        const int k = base + it * 32 + lane;
        const bool live = k < n;
        key[it] = live ? keys_in[k] : ~0ull;
        val[it] = live ? idx_in[k] : 0;
        const unsigned d = (unsigned)(key[it] >> shift) & (RS_BINS - 1);
        // lanes with the same digit (dead lanes form their own group and touch nothing)
        unsigned peers = __match_any_sync(FULL_MASK, live ? d : 0x10000u);
        const unsigned before = __popc(peers & ((1u << lane) - 1));
        const int leader = __ffs(peers) - 1;
        unsigned start = 0;
        if (live && lane == leader) {
            start = cnt[w][d];
            cnt[w][d] = start + __popc(peers);
        }
        start = __shfl_sync(FULL_MASK, start, leader);
        rank[it] = start + before;
        __syncwarp();
    }
    __syncthreads();

    // Per digit: offsets of the warps inside the digit's run, the run's start inside the tile
    // (exclusive scan over digits) and in the output.  RS_THREADS == RS_BINS: thread d owns digit d.
    {
        const int d = threadIdx.x;
        unsigned tot = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            const unsigned c = cnt[ww][d];
            cnt[ww][d] = tot;                      // warp offset inside the digit's run
            tot += c;
        }
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        unsigned before = 0;
        for (int ww = 0; ww < w; ww++) before += warp_tot[ww];
        local_excl[d] = before + incl - tot;
        gbase[d] = bin_base[d] + hist[(size_t)d * ntiles + blockIdx.x];
    }
    __syncthreads();

    // Stage the tile in digit order in shared memory, then write it out: consecutive staged
    // entries of one digit go to consecutive addresses, so every digit's run leaves as one
    // segment instead of 8 + 4 byte stores scattered over 256 destinations.
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        const int k = base + it * 32 + lane;
        if (k < n) {
            const unsigned d = (unsigned)(key[it] >> shift) & (RS_BINS - 1);
            const unsigned lpos = local_excl[d] + cnt[w][d] + rank[it];
            s_key[lpos] = key[it];
            s_val[lpos] = val[it];
        }
    }
    __syncthreads();
    const int tile_n = min(RS_TILE, n - blockIdx.x * RS_TILE);
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const int sidx = j * RS_THREADS + threadIdx.x;
        if (sidx < tile_n) {
            const uint64_t kk = s_key[sidx];
            const unsigned d = (unsigned)(kk >> shift) & (RS_BINS - 1);
            const unsigned dst = gbase[d] + ((unsigned)sidx - local_excl[d]);
            keys_out[dst] = kk;
            idx_out[dst] = s_val[sidx];
        }
    }
}

// After the radix passes over the top `64 - low_bits` bits of key_hi: order every run of keys
// that agree in those bits by (key_hi, key_lo, index).  One thread per run start; runs are a
// handful of particles at most (the sorted bits resolve cells of 2^-16 Boxsize or finer), so a
// serial insertion sort is fine.  hi_sorted is permuted along with idx.
// A run longer than RS_TIE_CAP (thousands of particles inside one 2^-16 Boxsize cell, or exact
// duplicates) would make that O(L^2) on one thread: it is recorded in `long_runs` and sorted by
// a whole block in k_fix_long_ties (bitonic network, O(L log^2 L)).  n_tied: [0] particles in
// runs, [1] set when more than RS_LONG_CAP long runs exist (an error), [2] long runs recorded.
#define RS_TIE_CAP 256
#define RS_LONG_CAP 1024
__global__ void k_fix_ties(int n, uint64_t *__restrict__ hi_sorted, int *__restrict__ idx,
                           const uint64_t *__restrict__ key_lo, int low_bits, int *__restrict__ n_tied,
                           int2 *__restrict__ long_runs)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t h = hi_sorted[k] >> low_bits;
    if (k > 0 && (hi_sorted[k - 1] >> low_bits) == h) return;      // not a run start
    if (k + 1 >= n || (hi_sorted[k + 1] >> low_bits) != h) return;  // run of one
    int end = k + 1;
    while (end < n && (hi_sorted[end] >> low_bits) == h) end++;
    atomicAdd(n_tied, end - k);
    if (end - k > RS_TIE_CAP) {
        const int slot = atomicAdd(n_tied + 2, 1);
        if (slot < RS_LONG_CAP) long_runs[slot] = make_int2(k, end);
        else atomicMax(n_tied + 1, end - k);
        return;
    }
    for (int a = k + 1; a < end; a++) {
        const int ia = idx[a];
        const uint64_t ha = hi_sorted[a], la = key_lo[ia];
        int b = a - 1;
        while (b >= k) {
            const int ib = idx[b];
            const uint64_t hb = hi_sorted[b], lb = key_lo[ib];
            if (hb < ha || (hb == ha && (lb < la || (lb == la && ib < ia)))) break;
            idx[b + 1] = ib;
            hi_sorted[b + 1] = hb;
            b--;
        }
        idx[b + 1] = ia;
        hi_sorted[b + 1] = ha;
    }
}

// The long runs k_fix_ties left: one block per run, bitonic sorting network on (key_hi, key_lo,
// index) in global memory.  Every merge is ascending (first step of a stage mirrors the upper
// half, partner = i ^ (size - 1)), so the virtual +inf pad of a run that is no power of two
// never moves and pairs that reach into it are skipped.  Launched every step; without long runs
// the blocks return at once.
__global__ void __launch_bounds__(1024) k_fix_long_ties(uint64_t *__restrict__ hi_sorted, int *__restrict__ idx,
                                                        const uint64_t *__restrict__ key_lo,
                                                        const int *__restrict__ n_tied,
                                                        const int2 *__restrict__ long_runs)
{
    const int nruns = min(n_tied[2], RS_LONG_CAP);
    for (int r = blockIdx.x; r < nruns; r += gridDim.x) {
        const int k = long_runs[r].x, L = long_runs[r].y - long_runs[r].x;
        uint64_t *h = hi_sorted + k;
        int *ix = idx + k;
        auto cmpswap = [&](int i, int j) {               // i < j < L: smaller element to i
            const uint64_t hi_i = h[i], hi_j = h[j];
            const int ii = ix[i], ij = ix[j];
            bool swap = hi_j < hi_i;
            if (hi_j == hi_i) {
                const uint64_t li = key_lo[ii], lj = key_lo[ij];
                swap = lj < li || (lj == li && ij < ii);
            }
            if (swap) { h[i] = hi_j; h[j] = hi_i; ix[i] = ij; ix[j] = ii; }
        };
        int P = 1;
        while (P < L) P <<= 1;
        for (int size = 2; size <= P; size <<= 1) {
            const int half = size >> 1;
            for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
                const int blk = t / half, off = t - blk * half;
                const int i = blk * size + off, j = blk * size + size - 1 - off;
                if (j < L) cmpswap(i, j);
            }
            __syncthreads();
            for (int stride = half >> 1; stride >= 1; stride >>= 1) {
                for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
                    const int i = (t / stride) * 2 * stride + (t % stride), j = i + stride;
                    if (j < L) cmpswap(i, j);
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
}
