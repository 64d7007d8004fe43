// SIFT / integer-L2 matching on 5th-gen tensor cores: everything sift_tc4_kernel (sift_tc4.cuh, the product kernel) shares
// with its exact fix-up -- the norm-sorted map layout, the per-64-row tile records, the bracketed pre-filter and the
// candidate format.  (This file was the third kernel generation; its own main kernel, N = 64 tiles with double-buffered
// accumulators, and the two generations before it were removed once sift_tc4 superseded them -- DESIGN.md section 4.1
// keeps their measurements, git history keeps the code.)
//
//   replaces: Matcher.SIFTMatch (Matcher.py:262-291) looped by Matcher.run (Matcher.py:305-338) looped by
//             analyzer.createRawP (analyze.py:78-90): knnMatch(k=2) + ratio lambda (Matcher.py:280) + len().
//
// At map build the rows of every image are SORTED BY NORM, so the 32 rows of a group have nearly the same |m|^2 (real
// SIFT: 512^2 +- a few hundred over an image, a few units inside a group).  The epilogue takes the bare maximum of the
// accumulators A_ij = q_i.m_j over the group (3-input max tree, no per-element add, no shared-memory bias) and brackets
// the group's best key 2 q.m - |m|^2 by
//        lo_g = 2*max_j A - nmax_g   <=   best key in g   <=   2*max_j A - nmin_g = hi_g .
// Per image it tracks the group with the largest hi (G1), and over all other groups max(hi) and max(lo).  A (query row,
// image) pair is rejected for good when even the most favourable reading (best = hi_G1, second = max lo of the others)
// fails the ratio test; survivors go to sift_fixup3_kernel, which recomputes G1 exactly (dp4a), evaluates the ratio test
// at both ends of the bracket of the others and, only if the verdicts differ, rescans the image exactly.  Counts are
// bit-exact whatever the norms are; the sort only decides how often the exact rescan runs (practically never).
#pragma once
#include "common.cuh"

namespace mcl {
namespace sifttc3 {

struct Chunk {  // a contiguous range of whole images handled by one work item
    int32_t tile_begin, tile_end;  // tiles [begin,end)
    int32_t img_begin, img_end;    // images [begin,end) whose results this chunk owns
};

constexpr int TILE_N = 64;
constexpr int GROUP = 32;              // rows per group; images are padded to a multiple of GROUP rows
constexpr int GROUPS = TILE_N / GROUP;
constexpr int QTILE = 128;
constexpr int ATILES = 3;
constexpr int QBLK = ATILES * QTILE;   // 384 query rows per work item
constexpr int META_BYTES = 48;
constexpr int NONE = INT_MIN / 2;      // "no group yet": Q - NONE still fits in int32
constexpr int POS_WIDE = 1 << 24;      // flag in a group position / candidate row: G1 is a whole 128-row tile (sift_tc4)

struct TileMeta {
    int32_t img[GROUPS];    // image id of each group (-1: padding group)
    int32_t nval[GROUPS];   // number of real descriptors of that image
    int32_t nmin[GROUPS];   // smallest / largest |m|^2 among the real rows of the group
    int32_t nmax[GROUPS];
    int32_t end_mask;       // bit g: group g is the last group of its image
    int32_t tnmin, tnmax;   // bracket of the whole 64-row tile (meaningful when both groups belong to one image)
    int32_t merge;          // even tiles: 1 = this tile and the next lie inside one image, no image ends before the end of the
                            // pair, and the pair's |m|^2 bracket is at most MERGE_WIDTH wide -> sift_tc4 updates its brackets
                            // once per 128 rows.  A wide bracket (outlier norms) would only multiply the candidates.
};
constexpr int MERGE_WIDTH = 1024;
static_assert(sizeof(TileMeta) == META_BYTES, "meta size");

struct Params {
    const uint8_t* q_desc;    // SW128 atoms, n_qblocks*QBLK rows (batch base)
    const int32_t* q_norm;    // -1 for padding rows
    const uint8_t* m_desc;    // SW128 atoms, n_tiles*TILE_N rows, rows of an image sorted by norm
    const TileMeta* m_meta;   // n_tiles
    const Chunk* chunks;      // tile indices in units of TILE_N rows
    int n_qblocks, n_chunks;
    double ratio;
    int4* cand;               // {q row, first map row of G1, max hi of the other groups, max lo of the other groups}
    unsigned int* cand_count;
    unsigned int cand_cap;
    int* overflow;            // set when a candidate did not fit the list
    uint8_t* dirty;           // [n_frames][n_images]: pairs that lost a candidate (recomputed exactly after the fix-up)
    const int32_t* q_frame;   // frame of every query row (batch base)
    int n_images;
    int q_row_base;
    int meta_cap;             // chunks up to this many tiles stage their metadata in shared memory (<= META_CAP)
    long long* diag;          // optional (MCL_TC9_DIAG): per CTA, cycles the MMA warp waited for {query tiles, map tiles, accumulators} + total
};

// ------------------------------------------------------------------------------------------
// per-tile metadata (built once per map)
// ------------------------------------------------------------------------------------------
__global__ void build_meta_kernel(const int32_t* __restrict__ m_norm, const int32_t* __restrict__ img_of,
                                  const int64_t* __restrict__ img_src_off, int n_tiles, TileMeta* __restrict__ meta) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    TileMeta mt;
    int mask = 0;
    for (int g = 0; g < GROUPS; ++g) {
        const int64_t r0 = (int64_t)t * TILE_N + g * GROUP;
        const int im = img_of[r0];
        mt.img[g] = im;
        mt.nval[g] = im >= 0 ? (int)(img_src_off[im + 1] - img_src_off[im]) : 0;
        int lo = INT_MAX, hi = 0;
        for (int r = 0; r < GROUP; ++r) {
            const int n = m_norm[r0 + r];
            if (n >= 0) { lo = min(lo, n); hi = max(hi, n); }
        }
        mt.nmin[g] = lo == INT_MAX ? 0 : lo;
        mt.nmax[g] = hi;
        const int64_t rn = r0 + GROUP;
        const int nx = (rn < (int64_t)n_tiles * TILE_N) ? img_of[rn] : -1;
        if (im >= 0 && nx != im) mask |= 1 << g;
    }
    mt.end_mask = mask;
    mt.tnmin = min(mt.nmin[0], mt.nmin[1]);
    mt.tnmax = max(mt.nmax[0], mt.nmax[1]);
    mt.merge = 0;
    if ((t & 1) == 0 && t + 1 < n_tiles && mask == 0 && mt.img[0] >= 0 && mt.img[0] == mt.img[1]) {
        const int64_t r0 = (int64_t)(t + 1) * TILE_N;
        // the next tile: both groups in the same image, and the image does not end after its first group
        bool ok = img_of[r0] == mt.img[0] && img_of[r0 + GROUP] == mt.img[0];
        int lo = mt.tnmin, hi = mt.tnmax;
        for (int r = 0; r < TILE_N; ++r) {
            const int n = m_norm[r0 + r];
            if (n >= 0) { lo = min(lo, n); hi = max(hi, n); }
        }
        mt.merge = (ok && hi - lo <= MERGE_WIDTH) ? 1 : 0;
    }
    meta[t] = mt;
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void finalize_image(const Params& p, int lane, bool row_valid, int Q, int64_t qrow, int H1, int L1,
                                            int H2, int L2, int pos, int img, int nval, const Chunk& ch) {
    bool pass = false;
    if (row_valid && nval >= 2 && img >= ch.img_begin && img < ch.img_end) {
        const int d0lo = max(0, Q - H1);                       // smallest possible best distance^2
        // G1 holds a row with key >= L1 and some other tile a row with key >= L2, so the second-best key of the image is
        // at least min(L1, L2) (G1 is the tile with the largest UPPER bound; with wide brackets its own rows may be poor)
        const int d1hi = L2 <= NONE ? (1 << 30) : Q - min(L1, L2);   // largest possible second-best distance^2
        pass = ratio_pass(d0lo, d1hi, p.ratio);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (bal) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(p.cand_count, (unsigned)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) {
            const unsigned idx = base + __popc(bal & ((1u << lane) - 1u));
            if (idx < p.cand_cap)
                p.cand[idx] = make_int4(p.q_row_base + (int)qrow, (pos & (POS_WIDE - 1)) * GROUP | ((pos & POS_WIDE) << 6), H2, L2);
            else {   // list full: this (frame, image) pair is recomputed by the exact SIMT kernel at the end of the call
                p.dirty[(int64_t)p.q_frame[qrow] * p.n_images + img] = 1;
                *p.overflow = 1;
            }
        }
    }
}

// Inline gate in front of finalize_image (sift_tc4): most (query row, image) pairs fail the ratio test by a wide margin, and
// the exact test costs two correctly rounded square roots, float64 arithmetic and a function call per image end -- which is
// every second tile for 256-row images.  A pair whose most favourable reading has d0 > ratio^2 * d1 * (1 + 1e-4) + 2 cannot
// pass the exact float test (its roundings are ~1e-7 relative), so a warp in which every lane is that far off skips the call.
// c2 = (float)(ratio^2 * 1.0001).
__device__ __forceinline__ void finalize_gate(const Params& p, float c2, int lane, bool row_valid, int Q, int64_t qrow, int H1, int L1,
                                              int H2, int L2, int pos, int img, int nval, const Chunk& ch) {
    if (nval < 2 || img < ch.img_begin || img >= ch.img_end) return;   // warp-uniform
    const int d0lo = max(0, Q - H1);
    const int d1hi = L2 <= NONE ? (1 << 30) : Q - min(L1, L2);
    const bool maybe = row_valid && !((float)d0lo > fmaf((float)d1hi, c2, 2.0f));
    if (__any_sync(0xffffffffu, maybe)) finalize_image(p, lane, row_valid, Q, qrow, H1, L1, H2, L2, pos, img, nval, ch);
}

template <int MODE>
__device__ __forceinline__ int group_max(const uint32_t (&v)[32]) {
    if (MODE == 3) return (int)(v[0] | v[31]);  // diagnostic: no arithmetic (MODE 5 = product arithmetic + timeline dump)
    int l[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) l[i] = max3((int)v[3 * i], (int)v[3 * i + 1], (int)v[3 * i + 2]);
    l[10] = max((int)v[30], (int)v[31]);
    const int a0 = max3(l[0], l[1], l[2]), a1 = max3(l[3], l[4], l[5]), a2 = max3(l[6], l[7], l[8]);
    return max3(max3(a0, a1, a2), l[9], l[10]);
}

// ------------------------------------------------------------------------------------------
// exact resolution of candidates.  One warp per candidate; eight lanes share a row of the tile that holds G1 (dots64).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int exact_key(const uint4 (&q)[8], const uint8_t* __restrict__ m_desc, int64_t mrow, int nm) {
    unsigned dot = 0;
#pragma unroll
    for (int chunk = 0; chunk < 8; ++chunk) {
        const uint4 b = *reinterpret_cast<const uint4*>(m_desc + sw128_chunk_offset(mrow, chunk));
        dot = __dp4a(q[chunk].x, b.x, dot); dot = __dp4a(q[chunk].y, b.y, dot);
        dot = __dp4a(q[chunk].z, b.z, dot); dot = __dp4a(q[chunk].w, b.w, dot);
    }
    return (nm >= 0) ? (int)(2u * dot) - nm : INT_MIN;
}

// top-2 (with multiplicity) over the warp of per-lane (b1 >= b2)
__device__ __forceinline__ void warp_top2(int b1, int b2, int& best, int& second) {
    int m = b1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    const unsigned eq = __ballot_sync(0xffffffffu, b1 == m);
    int s = (b1 == m) ? b2 : b1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = max(s, __shfl_xor_sync(0xffffffffu, s, o));
    best = m;
    second = (__popc(eq) >= 2) ? m : s;
}

// Partial dot products of 64 consecutive map rows with one query row, 8 lanes per row: lane (sub = lane & 7, slot = lane >> 3)
// reads the 16 bytes at physical chunk position `sub` of rows base + 4 j + slot (j = 0..15), so a warp-wide load covers four
// whole 128-byte rows (coalesced), all 16 loads are independent, and a lane needs only the two query chunks those positions
// hold under the SW128 swizzle (logical chunk = sub ^ (row & 7); row & 7 = 4 (j & 1) + slot).  A transposing butterfly over
// the three sub bits then leaves lane `sub` with the complete dots of j = sub and j = 8 + sub.
__device__ __forceinline__ void dots64(const uint8_t* __restrict__ m_desc, int64_t base, int sub, int slot, const uint4& q_even,
                                       const uint4& q_odd, unsigned& d_lo, unsigned& d_hi) {
    unsigned p[16];
    uint4 b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
        b[j] = *reinterpret_cast<const uint4*>(m_desc + (size_t)(base + 4 * j + slot) * 128u + (size_t)(sub << 4));
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint4& q = (j & 1) ? q_odd : q_even;
        unsigned d = __dp4a(q.x, b[j].x, 0u);
        d = __dp4a(q.y, b[j].y, d); d = __dp4a(q.z, b[j].z, d); d = __dp4a(q.w, b[j].w, d);
        p[j] = d;
    }
    // butterfly: after the step on bit k, a lane keeps the j's whose bit k equals its own sub bit k
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned keep = (sub & 1) ? p[2 * i + 1] : p[2 * i], send = (sub & 1) ? p[2 * i] : p[2 * i + 1];
        p[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);          // j = 2 i + b0
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const unsigned keep = (sub & 2) ? p[2 * i + 1] : p[2 * i], send = (sub & 2) ? p[2 * i] : p[2 * i + 1];
        p[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);          // j = 4 i + 2 b1 + b0
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const unsigned keep = (sub & 4) ? p[2 * i + 1] : p[2 * i], send = (sub & 4) ? p[2 * i] : p[2 * i + 1];
        p[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);          // j = 8 i + sub
    }
    d_lo = p[0];
    d_hi = p[1];
}

__global__ void __launch_bounds__(256, 2)   // 120 registers, no spills; 3 blocks (80 registers) measured 3 % slower
sift_fixup3_kernel(const int4* __restrict__ cand, const unsigned int* __restrict__ cand_count,
                                   unsigned int cand_cap, const uint8_t* __restrict__ q_desc,
                                   const int32_t* __restrict__ q_norm, const int32_t* __restrict__ q_frame,
                                   const uint8_t* __restrict__ m_desc, const int32_t* __restrict__ m_norm,
                                   const int32_t* __restrict__ m_img_of, const int64_t* __restrict__ m_dst_off,
                                   double ratio, int n_images, int32_t* __restrict__ counts,
                                   unsigned int* __restrict__ n_rescans, unsigned long long* __restrict__ n_candidates) {
    const int lane = threadIdx.x & 31, sub = lane & 7, slot = lane >> 3;
    const unsigned n = min(*cand_count, cand_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(n_candidates, (unsigned long long)n);   // statistics
    const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int4 cd_next = c < n ? cand[c] : make_int4(0, 0, 0, 0);
    for (; c < n; c += n_warps) {
        const int4 cd = cd_next;
        if (c + n_warps < n) cd_next = cand[c + n_warps];     // the next candidate's record travels under this one's work
        const int64_t qrow = cd.x;
        // G1 is a 32-row group, a 64-row tile, or (flag, sift_tc4's merged brackets) a 128-row tile of one image:
        // recompute the tile's rows of this image exactly
        const bool wide = (cd.y & (POS_WIDE << 6)) != 0;
        const int64_t g1row = (int64_t)(cd.y & ((POS_WIDE << 6) - 1));
        const int64_t tile0 = g1row & ~(int64_t)((wide ? 2 * TILE_N : TILE_N) - 1);
        const uint4 q_even = *reinterpret_cast<const uint4*>(q_desc + sw128_chunk_offset(qrow, sub ^ slot));
        const uint4 q_odd = *reinterpret_cast<const uint4*>(q_desc + sw128_chunk_offset(qrow, sub ^ (4 + slot)));
        const int Q = q_norm[qrow];
        const int img = m_img_of[g1row];
        // this lane ends up with the rows tile0 + 4 (8 i + sub) + slot, i = 0..1 (0..3 when wide)
        int64_t row[4];
        int nm[4], io[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            row[i] = tile0 + 4 * (8 * i + sub) + slot;
            const bool on = i < 2 || wide;
            nm[i] = on ? m_norm[row[i]] : -1;
            io[i] = on ? m_img_of[row[i]] : -1;
        }
        unsigned d[4] = {0, 0, 0, 0};
        dots64(m_desc, tile0, sub, slot, q_even, q_odd, d[0], d[1]);
        if (wide) dots64(m_desc, tile0 + 64, sub, slot, q_even, q_odd, d[2], d[3]);   // warp-uniform
        int m1 = INT_MIN, s1 = INT_MIN;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = (nm[i] >= 0 && io[i] == img) ? (int)(2u * d[i]) - nm[i] : INT_MIN;
            s1 = max(s1, min(m1, k));
            m1 = max(m1, k);
        }
        int k1, k2;
        warp_top2(m1, s1, k1, k2);
        // the best key of the other groups lies in [lo, hi]; no key exceeds Q (d^2 >= 0), and an unclamped upper bound
        // would turn into a negative d^2 (NaN under sqrtf: the reading would silently count as a failed test)
        const int hi = min(cd.z, Q), lo = min(cd.w, Q);
        bool verdict, decided = true;
        if (hi <= NONE) {                 // the image has no other group
            verdict = k2 > INT_MIN && ratio_pass(Q - k1, Q - k2, ratio);
        } else if (k2 > INT_MIN && hi <= k2) {
            verdict = ratio_pass(Q - k1, Q - k2, ratio);
        } else {
            // every reading of (best, second) consistent with the bracket, at its two ends (the test is monotone)
            bool any = false, all = true;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int ko = r ? hi : lo;
                if (ko <= k1) {
                    const bool x = ratio_pass(Q - k1, Q - max(k2, ko), ratio);
                    any |= x; all &= x;
                } else {   // another group may hold the best row; the second best is then between k1 and ko
                    const bool x = ratio_pass(Q - ko, Q - k1, ratio);
                    any |= x; all = false;  // (best = second = ko) never passes
                }
            }
            verdict = all;
            decided = (any == all);
        }
        if (!decided) {  // exact scan of the whole image (warp-uniform branch)
            const int64_t r0 = m_dst_off[img], r1 = m_dst_off[img + 1];
            uint4 q[8];
#pragma unroll
            for (int chunk = 0; chunk < 8; ++chunk) q[chunk] = *reinterpret_cast<const uint4*>(q_desc + sw128_chunk_offset(qrow, chunk));
            int b1 = INT_MIN, b2 = INT_MIN;
            for (int64_t r = r0 + lane; r < r1; r += 32) {
                const int k = exact_key(q, m_desc, r, m_norm[r]);
                b2 = max(b2, min(b1, k));
                b1 = max(b1, k);
            }
            int best, second;
            warp_top2(b1, b2, best, second);
            verdict = second > INT_MIN && ratio_pass(Q - best, Q - second, ratio);
            if (lane == 0) atomicAdd(n_rescans, 1u);
        }
        if (lane == 0 && verdict) atomicAdd(&counts[(int64_t)q_frame[qrow] * n_images + img], 1);
    }
}

}  // namespace sifttc3
}  // namespace mcl
