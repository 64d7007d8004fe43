// SIFT / integer-L2 matching on 5th-gen tensor cores, third generation: the kernel the product path uses.
//
//   replaces: Matcher.SIFTMatch (Matcher.py:262-291) looped by Matcher.run (Matcher.py:305-338) looped by
//             analyzer.createRawP (analyze.py:78-90): knnMatch(k=2) + ratio lambda (Matcher.py:280) + len().
//
// sift_tc2 made the kernel tensor-bound (tensor pipe 94 % busy) but pays a 5th MMA per tile to carry the -|m|^2/2
// term inside the contraction: 20 % of the tensor time is not algorithmic work.  This generation gets the norm term
// for free instead: at map build the rows of every image are SORTED BY NORM, so the 32 rows of a group have nearly
// the same |m|^2 (real SIFT: 512^2 +- a few hundred over an image, a few units inside a group).  The epilogue takes
// the bare maximum of the accumulators A_ij = q_i.m_j over the group (3-input max tree, no per-element add, no
// shared-memory bias) and brackets the group's best key 2 q.m - |m|^2 by
//        lo_g = 2*max_j A - nmax_g   <=   best key in g   <=   2*max_j A - nmin_g = hi_g .
// Per image it tracks the group with the largest hi (G1), and over all other groups max(hi) and max(lo).  A (query row,
// image) pair is rejected for good when even the most favourable reading (best = hi_G1, second = max lo of the others)
// fails the ratio test; survivors go to sift_fixup3_kernel, which recomputes G1 exactly (dp4a), evaluates the ratio test
// at both ends of the bracket of the others and, only if the verdicts differ, rescans the image exactly.  Counts are
// bit-exact whatever the norms are; the sort only decides how often the exact rescan runs (practically never).
//
// Everything else is sift_tc2's pipeline: query block (3 x 128 rows) stationary in TENSOR MEMORY (A-from-TMEM MMA, 64 B/clk
// of shared-memory reads instead of 128), 64-row map tiles streamed by TMA bulk copies in the SWIZZLE_128B image, 2 x 3
// accumulators of 64 columns with their own full/empty barriers, 12 epilogue warps, one TMA warp, two MMA issuing warps,
// tcgen05.mma / commit / cp.async.bulk issued through elect.sync in warp-uniform code.
#pragma once
#include "common.cuh"
#include "sift_tc.cuh"  // Chunk

namespace mcl {
namespace sifttc3 {

using sifttc::Chunk;

constexpr int TILE_N = 64;
constexpr int GROUP = sifttc::GROUP;   // 32
constexpr int GROUPS = TILE_N / GROUP;
constexpr int QTILE = 128;
constexpr int ATILES = 3;
constexpr int QBLK = ATILES * QTILE;   // 384 query rows per work item
constexpr int A_COLS = 32;             // TMEM columns per A tile (128 bytes per row)
constexpr int ACC_COL0 = 128;
constexpr int ACC_STAGES = 2;
constexpr int B_TILE_BYTES = TILE_N * 128;
constexpr int META_BYTES = 48;
constexpr int STAGE_BYTES = B_TILE_BYTES;          // 8 KB: every stage stays 1024-byte aligned
constexpr int B_STAGES = 16;
constexpr int EPI_WARPS = ATILES * 4;
constexpr int MMA_WARPS = 2;
constexpr int THREADS = (EPI_WARPS + 1 + MMA_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr int META_CAP = 352;          // tiles of a chunk whose metadata is staged in shared memory (longer chunks read it from L2)
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

struct Smem {
    alignas(1024) uint8_t stage[B_STAGES][STAGE_BYTES];
    alignas(16) int4 meta[META_CAP * 3];   // the current chunk's TileMeta records
    alignas(8) uint64_t b_full[B_STAGES];
    uint64_t b_empty[B_STAGES];
    uint64_t acc_full[ACC_STAGES][ATILES], acc_empty[ACC_STAGES][ATILES];
    uint64_t a_full, a_free;
    uint32_t tmem_base;
};

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
    int* overflow;
    int q_row_base;
    int meta_cap;             // chunks up to this many tiles stage their metadata in shared memory (<= META_CAP)
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
            else
                *p.overflow = 1;
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

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) sift_tc3_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int warp = warp_id_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t leader = lane == 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < B_STAGES; ++i) { mbar_init(&s.b_full[i], 1); mbar_init(&s.b_empty[i], 1); }
        for (int i = 0; i < ACC_STAGES; ++i)
            for (int a = 0; a < ATILES; ++a) { mbar_init(&s.acc_full[i][a], 1); mbar_init(&s.acc_empty[i][a], 4); }
        mbar_init(&s.a_full, EPI_WARPS);
        mbar_init(&s.a_free, MMA_WARPS);
        mbar_fence_init();
    }
    if (warp == EPI_WARPS + 1) tmem_alloc<TMEM_COLS>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    const int n_items = p.n_qblocks * p.n_chunks;

    if (warp == EPI_WARPS) {
        // ================= TMA producer: one 64-row map tile (8 KB) per stage ==========================================
        uint32_t bn = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Chunk ch = p.chunks[item / p.n_qblocks];
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                mbar_wait(&s.b_empty[bs], bph ^ 1);
                mbar_arrive_expect_tx_if(leader, &s.b_full[bs], B_TILE_BYTES);
                bulk_g2s_if(leader, s.stage[bs], p.m_desc + (size_t)t * B_TILE_BYTES, B_TILE_BYTES, &s.b_full[bs]);
            }
        }
    } else if (warp > EPI_WARPS) {
        // ================= MMA issuers (one per accumulator stage): A from tensor memory, B from shared memory ========
        const uint32_t my_stage = warp - (EPI_WARPS + 1);
        constexpr uint32_t idesc = umma_idesc(/*c=S32*/ 2, /*a=u8*/ 0, /*b=u8*/ 0, QTILE, TILE_N);
        uint32_t bn = 0, item_n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
            const Chunk ch = p.chunks[item / p.n_qblocks];
            mbar_wait(&s.a_full, item_n & 1);
            tc_fence_after();
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                const uint32_t as = bn % ACC_STAGES, aph = (bn / ACC_STAGES) & 1;
                if (as != my_stage) continue;
                mbar_wait(&s.b_full[bs], bph);
                const uint64_t b_desc = umma_desc_sw128(smem_u32(s.stage[bs]));
                // The three accumulators of this stage are independent: issue each A tile's MMAs as soon as ITS accumulator
                // has been handed back, in whatever order that happens (in-order issue lets the slowest of the 12 epilogue
                // warps hold up the other two A tiles: a convoy).
                uint32_t pending = (1u << ATILES) - 1u;
                while (pending) {
#pragma unroll
                    for (int a = 0; a < ATILES; ++a) {
                        if (!(pending & (1u << a))) continue;
                        if (!mbar_test_wait(&s.acc_empty[as][a], aph ^ 1)) continue;
                        pending &= ~(1u << a);
                        tc_fence_after();
                        const uint32_t d = tmem_base + ACC_COL0 + (as * ATILES + a) * TILE_N;
                        const uint32_t ta = tmem_base + a * A_COLS;
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_i8_ts_if(leader, d, ta + 8 * k, b_desc + 2 * k, idesc, k > 0);
                        tc_commit_if(leader, &s.acc_full[as][a]);
                    }
                }
                tc_commit_if(leader, &s.b_empty[bs]);
                __syncwarp();
            }
            tc_commit_if(leader, &s.a_free);
            __syncwarp();
        }
    } else {
        // ================= epilogue warps ===========================================================================
        const int atile = warp >> 2;
        const int lane_base = (warp & 3) * 32;
        const int row_in_blk = atile * QTILE + lane_base + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)lane_base << 16);
        uint32_t bn = 0, item_n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
            const int qb = item % p.n_qblocks;
            const Chunk ch = p.chunks[item / p.n_qblocks];
            const int64_t qrow = (int64_t)qb * QBLK + row_in_blk;
            {   // this thread's query row -> tensor memory (row = TMEM lane, 4 K-bytes per column)
                uint32_t w[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 x = *reinterpret_cast<const uint4*>(p.q_desc + sw128_chunk_offset(qrow, c));
                    w[4 * c] = x.x; w[4 * c + 1] = x.y; w[4 * c + 2] = x.z; w[4 * c + 3] = x.w;
                }
                mbar_wait(&s.a_free, (item_n & 1) ^ 1);   // every MMA of the previous item has read its A tiles
                tc_fence_after();
                tmem_st32(t_lane + atile * A_COLS, w);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.a_full);
            }
            const int Q = p.q_norm[qrow];
            const bool row_valid = Q >= 0;
            int H1 = NONE, L1 = NONE, H2 = NONE, L2 = NONE, pos = 0;
            // one (hi, lo) bracket joins the running state: the loser of (bracket, current G1) joins "the others"
            auto join = [&](int gm, int nmin, int nmax, int id) {
                const int hi = 2 * gm - nmin, lo = 2 * gm - nmax;
                const bool better = hi > H1;
                H2 = max(H2, better ? H1 : hi);
                L2 = max(L2, better ? L1 : lo);
                pos = better ? id : pos;
                L1 = better ? lo : L1;
                H1 = max(H1, hi);
            };
            // the chunk's tile metadata (48 B per tile) is staged in shared memory once per work item by the 12 epilogue
            // warps (named barrier 1); per-tile reads from global memory miss L1 whenever an SM sees a chunk only once
            const int n_ct = ch.tile_end - ch.tile_begin;
            const int4* mp = reinterpret_cast<const int4*>(p.m_meta + ch.tile_begin);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");   // nobody still reads the previous chunk's records
            if (n_ct <= p.meta_cap) {
                for (int i = threadIdx.x; i < n_ct * 3; i += EPI_WARPS * 32) s.meta[i] = __ldg(mp + i);
                mp = s.meta;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            int4 m0 = mp[0], m1 = mp[1], m2 = mp[2];
#pragma unroll 2   // halves the loop-head cost per tile (measured +1.5 %); deferring the scalar tail under the next load did not pay
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const uint32_t as = bn % ACC_STAGES, aph = (bn / ACC_STAGES) & 1;
                // m0 = {img0, img1, nval0, nval1}, m1 = {nmin0, nmin1, nmax0, nmax1}, m2 = {end_mask, tnmin, tnmax, -}
                const int4 c0 = m0, c1 = m1, c2 = m2;
                if (t + 1 < ch.tile_end) {
                    mp += 3;
                    m0 = mp[0]; m1 = mp[1]; m2 = mp[2];
                }
                long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
                if (MODE == 5) tk0 = clock64();
                mbar_wait(&s.acc_full[as][atile], aph);
                tc_fence_after();
                if (MODE == 5) tk1 = clock64();
                const uint32_t tcol = t_lane + ACC_COL0 + (as * ATILES + atile) * TILE_N;
                uint32_t v[GROUPS][32];
                if (MODE == 6) {   // diagnostic: half of the TMEM read volume (results meaningless)
                    tmem_ld32(tcol, v[0]);
                    tmem_wait_all();
                    tmem_fence_regs(v[0]);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[1][i] = v[0][i];
                } else if (MODE != 4) {
                    tmem_ld64(tcol, reinterpret_cast<uint32_t(&)[64]>(v));
                    tmem_wait_all();
#pragma unroll
                    for (int g = 0; g < GROUPS; ++g) tmem_fence_regs(v[g]);
                } else {
#pragma unroll
                    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[g][i] = 0;
                }
                if (MODE == 5) tk2 = clock64();
                // tcgen05.wait::ld is warp-aligned and returns when every lane's columns are in registers: the accumulator
                // can go back to the MMA warp right away (no further tcgen05 fence is needed for a completed load)
                if (lane == 0) mbar_arrive(&s.acc_empty[as][atile]);
                if (MODE == 5) tk3 = clock64();
                const int end_mask = c2.x;
                const int g0 = group_max<MODE>(v[0]), g1 = group_max<MODE>(v[1]);
                if ((end_mask & 1) == 0) {
                    // both groups belong to one image (the common case): one 64-row bracket
                    join(max(g0, g1), c2.y, c2.z, t * GROUPS);
                    if (end_mask & 2) {
                        finalize_image(p, lane, row_valid, Q, qrow, H1, L1, H2, L2, pos, c0.y, c0.w, ch);
                        H1 = NONE; L1 = NONE; H2 = NONE; L2 = NONE;
                    }
                } else {
                    join(g0, c1.x, c1.z, t * GROUPS);
                    finalize_image(p, lane, row_valid, Q, qrow, H1, L1, H2, L2, pos, c0.x, c0.z, ch);
                    H1 = NONE; L1 = NONE; H2 = NONE; L2 = NONE;
                    if (c0.y >= 0) {
                        join(g1, c1.y, c1.w, t * GROUPS + 1);
                        if (end_mask & 2) {
                            finalize_image(p, lane, row_valid, Q, qrow, H1, L1, H2, L2, pos, c0.y, c0.w, ch);
                            H1 = NONE; L1 = NONE; H2 = NONE; L2 = NONE;
                        }
                    }
                }
                if (MODE == 5 && blockIdx.x == 0 && warp == 5 && lane == 0 && bn >= 100 && bn < 164) {
                    long long* tl = reinterpret_cast<long long*>(p.cand) + (size_t)(1 << 20) + (bn - 100) * 5;
                    tl[0] = tk0; tl[1] = tk1; tl[2] = tk2; tl[3] = tk3; tl[4] = clock64();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// exact resolution of candidates.  One warp per candidate; lane l owns rows l and l+32 of the 64-row tile that holds G1.
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

__global__ void sift_fixup3_kernel(const int4* __restrict__ cand, const unsigned int* __restrict__ cand_count,
                                   unsigned int cand_cap, const uint8_t* __restrict__ q_desc,
                                   const int32_t* __restrict__ q_norm, const int32_t* __restrict__ q_frame,
                                   const uint8_t* __restrict__ m_desc, const int32_t* __restrict__ m_norm,
                                   const int32_t* __restrict__ m_img_of, const int64_t* __restrict__ m_dst_off,
                                   double ratio, int n_images, int32_t* __restrict__ counts,
                                   unsigned int* __restrict__ n_rescans) {
    const int lane = threadIdx.x & 31;
    const unsigned n = min(*cand_count, cand_cap);
    const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n; c += n_warps) {
        const int4 cd = cand[c];
        const int64_t qrow = cd.x;
        uint4 q[8];
#pragma unroll
        for (int chunk = 0; chunk < 8; ++chunk) q[chunk] = *reinterpret_cast<const uint4*>(q_desc + sw128_chunk_offset(qrow, chunk));
        const int Q = q_norm[qrow];
        // G1 is a 32-row group, a 64-row tile, or (flag, sift_tc4's merged brackets) a 128-row tile of one image:
        // recompute the tile's rows of this image exactly (2 or 4 per lane)
        const bool wide = (cd.y & (POS_WIDE << 6)) != 0;
        const int64_t g1row = (int64_t)(cd.y & ((POS_WIDE << 6) - 1));
        const int img = m_img_of[g1row];
        const int64_t tile0 = g1row & ~(int64_t)((wide ? 2 * TILE_N : TILE_N) - 1);
        int va = INT_MIN, vb = INT_MIN;
        {
            const int64_t ra = tile0 + lane, rb = tile0 + 32 + lane;
            if (m_img_of[ra] == img) va = exact_key(q, m_desc, ra, m_norm[ra]);
            if (m_img_of[rb] == img) vb = exact_key(q, m_desc, rb, m_norm[rb]);
        }
        int m1 = max(va, vb), s1 = min(va, vb);
        if (wide) {   // warp-uniform
            const int64_t rc = tile0 + 64 + lane, rd = tile0 + 96 + lane;
            const int vc = exact_key(q, m_desc, rc, m_norm[rc]), vd = exact_key(q, m_desc, rd, m_norm[rd]);   // one image by construction
            const int m2 = max(vc, vd), s2 = min(vc, vd);
            s1 = max(min(m1, m2), max(s1, s2));
            m1 = max(m1, m2);
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
