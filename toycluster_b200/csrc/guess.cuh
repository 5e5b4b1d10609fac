// guess.cuh -- cold-start smoothing length: hsml = 2*Guess_hsml(i) (sph.c:25-26,
// tree.c:113-121), which reads Npart and Size of the particle's Tree_Parent node.
//
// Those two numbers are a property of the reference's sequentially built octree
// (tree.c:124-236), but they can be stated in closed form over the sorted key array.
// With cpl[b] = number of leading key triplets particles b-1 and b share (= level of their
// deepest common cell; the sort key and the tree levels use the same triplets, peano.c:183-198
// and peano.c:266-279):
//   * without any leaf collapse, Tree_Parent(p) is the deepest cell holding p and one more
//     particle: level max(cpl[p], cpl[p+1]) (tree.c:163-171 refines until they separate);
//   * inserting particle b collapses a finished branch (tree.c:201-226) iff b lies outside
//     the previous parent cell, i.e. cpl[b] < cpl[b-1].  Candidates, in this order: the child
//     C of the common cell that holds b-1 (level cpl[b]+1), else the previous parent L
//     (level cpl[b-1]); the first with <= 8 particles becomes a leaf covering [b-count, b);
//   * later collapses can only swallow earlier ones (cells nest, ranges end at increasing b),
//     so the final Tree_Parent of p is the collapse with the LARGEST b <= p+8 covering p.
// Validated against the unmodified tree.c in tests (tg_guess_hsml vs ref Guess_hsml).
#pragma once
#include "common.cuh"

// number of common leading triplets of two 128-bit keys (42 = identical)
static __device__ __forceinline__ int common_triplets(uint64_t ah, uint64_t al, uint64_t bh,
                                                      uint64_t bl)
{
    const uint64_t xh = ah ^ bh, xl = al ^ bl;
    int lz;
    if (xh) lz = __clzll((long long)xh);
    else if (xl) lz = 64 + __clzll((long long)xl);
    else return 42;
    const int c = lz / 3;
    return c > 42 ? 42 : c;
}

__global__ void k_cpl(int n, const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo,
                      signed char *__restrict__ cpl)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    cpl[b] = b == 0 ? (signed char)-1
                    : (signed char)common_triplets(hi[b - 1], lo[b - 1], hi[b], lo[b]);
}

// Collapse decided when particle b is inserted: start = first particle of the new leaf (or
// -1), its level and particle count.
static __device__ __forceinline__ void collapse_event(int b, const signed char *__restrict__ cpl,
                                                      int &start, int &level, int &count)
{
    start = -1; level = 0; count = 0;
    if (b >= 2 && cpl[b] < cpl[b - 1]) {
        // particles b-1, b-2, ... sharing at least `lvl` triplets with b-1
        auto run = [&](int lvl) {
            int c = 1;
            for (int k = b - 1; k >= 1 && c <= 8 && cpl[k] >= lvl; k--) c++;
            return c;
        };
        const int lc = cpl[b] + 1, ll = cpl[b - 1];
        const int cc = run(lc);
        if (cc <= 8) { start = b - cc; level = lc; count = cc; }
        else {
            const int cl = run(ll);
            if (cl <= 8) { start = b - cl; level = ll; count = cl; }
        }
    }
}

__global__ void k_collapse_events(int n, const signed char *__restrict__ cpl,
                                  int *__restrict__ ev_start, signed char *__restrict__ ev_level,
                                  signed char *__restrict__ ev_count)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    int start, level, count;
    collapse_event(b, cpl, start, level, count);
    ev_start[b] = start;
    ev_level[b] = (signed char)level;
    ev_count[b] = (signed char)count;
}

static __device__ __forceinline__ bool key_less(uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl)
{
    return ah < bh || (ah == bh && al < bl);
}

__global__ void k_guess_hsml(int n, const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo,
                             const signed char *__restrict__ cpl, const int *__restrict__ ev_start,
                             const signed char *__restrict__ ev_level,
                             const signed char *__restrict__ ev_count, double boxsize,
                             float *__restrict__ guess2, int *__restrict__ parent_level)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;

    int level = -1, npart = 0;
    for (int b = min(p + 8, n - 1); b > p; b--) {
        if (ev_start[b] >= 0 && ev_start[b] <= p) { level = ev_level[b]; npart = ev_count[b]; break; }
    }
    if (level < 0) {
        level = max((int)cpl[p], p + 1 < n ? (int)cpl[p + 1] : -1);
        if (level < 0) level = 0;
        // all particles sharing `level` triplets with p: [first key >= prefix, first key > prefix|ones)
        const int bits = 3 * level;
        uint64_t mh, ml;   // mask of the prefix bits
        if (bits == 0) { mh = 0; ml = 0; }
        else if (bits <= 64) { mh = bits == 64 ? ~0ull : ~(~0ull >> bits); ml = 0; }
        else { mh = ~0ull; ml = ~(~0ull >> (bits - 64)); }
        const uint64_t lh = hi[p] & mh, ll = lo[p] & ml;       // smallest key of the cell
        const uint64_t uh = lh | ~mh, ul = ll | ~ml;           // largest key of the cell
        int a = 0, z = n;
        while (a < z) {   // first index with key >= (lh, ll)
            const int m = (a + z) >> 1;
            if (key_less(hi[m], lo[m], lh, ll)) a = m + 1; else z = m;
        }
        const int first = a;
        z = n;
        while (a < z) {   // first index with key > (uh, ul)
            const int m = (a + z) >> 1;
            if (!key_less(uh, ul, hi[m], lo[m])) a = m + 1; else z = m;
        }
        npart = a - first;
    }
    // tree.c:304: float size = Boxsize / (1 << lvl); tree.c:117-120
    const float size = (float)(boxsize / (double)(1 << (level & 31)));
    const float numdens = __fdiv_rn((float)npart, __fmul_rn(__fmul_rn(size, size), size));
    const float s = (float)pow(K_FOURPITHIRD / (double)numdens, 1. / 3.);
    guess2[p] = __fmul_rn(2.f, __fmul_rn(2.f, s));     // caller doubles: sph.c:26
    if (parent_level) parent_level[p] = level;
}
