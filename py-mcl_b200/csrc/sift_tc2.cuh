// SIFT / integer-L2 matching on 5th-gen tensor cores, second generation: the kernel the product path uses.
//
//   replaces: Matcher.SIFTMatch (Matcher.py:262-291) looped by Matcher.run (Matcher.py:305-338) looped by
//             analyzer.createRawP (analyze.py:78-90): knnMatch(k=2) + ratio lambda (Matcher.py:280) + len().
//
// What the first generation (sift_tc.cuh) taught on the B200 (profiles/): with both operands in shared memory a
// 128x128x32 kind::i8 MMA reads 8 KB of shared memory every 64 cycles -- the whole 128 B/clk of the SM -- so TMA
// writes, MMA reads and the epilogue's bias loads fight for the same port, and the epilogue spends an IMAD + half a
// 3-input max + a quarter LDS per s32 output when the MMA only spends 128 MACs on it.  Hence:
//
//   * the query block (3 x 128 rows, stationary for a whole chunk of map tiles) lives in TENSOR MEMORY (A-from-TMEM
//     MMA); only the streamed map tile is read from shared memory: 2 KB per 32-cycle MMA = 64 B/clk;
//   * the "- |m|^2 / 2" term rides in the contraction: K is extended from 128 to 160 bytes.  Query rows carry the
//     constants [1, 255 x 31] there, map row j carries the base-255 digits of E_j = (C - |m_j|^2 - parity)/2, so the
//     accumulator already holds  A_ij = q_i.m_j + E_j  and the epilogue is a bare 3-input-max tree (no adds, no
//     shared-memory bias reads).  Cost: a 5th MMA per tile (25% more tensor time); the tensor pipe is then the bound;
//   * 2*A_ij - C equals the true key 2 q.m - |m|^2 up to the dropped parity bit, so the tensor pass is a conservative
//     pre-filter: a (query row, image) pair is rejected for good when even the most favourable reading of
//     (best, best-of-other-groups) fails the ratio test; survivors go to sift_fixup2_kernel, which recomputes the
//     winning 32-row group exactly with dp4a and, in the rare case where the +-1 uncertainty of the other groups could
//     change the verdict, rescans the whole image exactly.  Counts are bit-exact.
//
// TMEM map (512 columns): [0,120) three A tiles x (32 data + 8 extension columns); [128,512) accumulators,
// 2 stages x 3 A tiles x 64 columns.  Warps: 0-11 epilogue (A tile = warp/4, lane quadrant = warp%4), 12 TMA
// producer, 13 MMA issuer.
#pragma once
#include "common.cuh"
#include "sift_tc.cuh"  // Chunk

namespace mcl {
namespace sifttc2 {

using sifttc::Chunk;

constexpr int TILE_N = 64;
constexpr int GROUP = sifttc::GROUP;   // 32: images are padded to a multiple of GROUP rows
constexpr int GROUPS = TILE_N / GROUP;
constexpr int QTILE = 128;
constexpr int ATILES = 3;
constexpr int QBLK = ATILES * QTILE;   // 384 query rows per work item
constexpr int A_COLS = 40;             // TMEM columns per A tile: 128 data bytes + 32 extension bytes
constexpr int ACC_COL0 = 128;
constexpr int ACC_STAGES = 2;
constexpr int B_TILE_BYTES = TILE_N * 128;     // 8 KB, SWIZZLE_128B atoms
constexpr int EXT_BYTES = TILE_N * 32;         // 2 KB, canonical no-swizzle K-major (LBO 128, SBO 256)
constexpr int META_BYTES = 32;
constexpr int EXT_META_BYTES = EXT_BYTES + META_BYTES;
constexpr int STAGE_BYTES = 11264;             // 8192 + 3072: keeps every stage 1024-byte aligned
constexpr int B_STAGES = 12;
constexpr int EPI_WARPS = ATILES * 4;
constexpr int MMA_WARPS = 2;            // one issuing warp per accumulator stage (the issue loop is latency bound)
constexpr int THREADS = (EPI_WARPS + 1 + MMA_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr int EXT_CAP = 254 + 255 * 31 * 255;  // largest E the 32 extension bytes can carry

struct TileMeta {
    int32_t img[GROUPS];    // image id of each group (-1: padding group)
    int32_t nval[GROUPS];   // number of real descriptors of that image
    int32_t end_mask;       // bit g: group g is the last group of its image
    int32_t pad[3];
};
static_assert(sizeof(TileMeta) == META_BYTES, "meta size");

struct Smem {
    alignas(1024) uint8_t stage[B_STAGES][STAGE_BYTES];
    alignas(8) uint64_t b_full[B_STAGES];
    uint64_t b_empty[B_STAGES];
    uint64_t acc_full[ACC_STAGES][ATILES], acc_empty[ACC_STAGES][ATILES];   // one accumulator = (stage, A tile)
    uint64_t a_full, a_free;
    uint32_t tmem_base;
};

struct Params {
    const uint8_t* q_desc;    // SW128 atoms, n_qblocks*QBLK rows (batch base)
    const int32_t* q_norm;    // -1 for padding rows
    const uint8_t* m_desc;    // SW128 atoms, n_tiles*TILE_N rows
    const uint8_t* m_ext;     // n_tiles x (EXT_BYTES + META_BYTES)
    const Chunk* chunks;      // tile indices in units of TILE_N rows
    int n_qblocks, n_chunks;
    int key_c;                // C: true key = 2*A - C (+ parity)
    double ratio;
    int4* cand;               // {q row, first map row of winning group, A of the best other group (-1: none), image}
    unsigned int* cand_count;
    unsigned int cand_cap;
    int* overflow;
    int q_row_base;
};

// ------------------------------------------------------------------------------------------
// per-tile extension bytes + metadata (built once per map)
// ------------------------------------------------------------------------------------------
__global__ void build_ext_kernel(const int32_t* __restrict__ m_norm, const int32_t* __restrict__ img_of,
                                 const int64_t* __restrict__ img_src_off, int n_tiles, int key_c,
                                 uint8_t* __restrict__ ext) {
    const int t = blockIdx.x;
    if (t >= n_tiles) return;
    const int c = threadIdx.x;  // 0..63: row inside the tile
    const int64_t row = (int64_t)t * TILE_N + c;
    uint8_t* tile = ext + (size_t)t * EXT_META_BYTES;
    const int nm = m_norm[row];
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (nm >= 0) {
        const int d = key_c - nm;
        const int e = (d - (d & 1)) >> 1;       // E_j >= 1 by the choice of C
        uint8_t b[32];
        b[0] = (uint8_t)(e % 255);
        int s = e / 255;
#pragma unroll
        for (int k = 1; k < 32; ++k) {
            const int x = s < 255 ? s : 255;
            b[k] = (uint8_t)x;
            s -= x;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            w[k] = (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) | ((uint32_t)b[4 * k + 3] << 24);
    }
    // canonical K-major no-swizzle layout: core matrix = 8 rows x 16 B; offset(row, kchunk) = (row/8)*256 + kchunk*128 + (row%8)*16
    uint4* dst0 = reinterpret_cast<uint4*>(tile + (c >> 3) * 256 + (c & 7) * 16);
    uint4* dst1 = reinterpret_cast<uint4*>(tile + (c >> 3) * 256 + 128 + (c & 7) * 16);
    *dst0 = make_uint4(w[0], w[1], w[2], w[3]);
    *dst1 = make_uint4(w[4], w[5], w[6], w[7]);
    if (c == 0) {
        TileMeta mt;
        int mask = 0;
        for (int g = 0; g < GROUPS; ++g) {
            const int64_t r0 = (int64_t)t * TILE_N + g * GROUP;
            const int im = img_of[r0];
            mt.img[g] = im;
            mt.nval[g] = im >= 0 ? (int)(img_src_off[im + 1] - img_src_off[im]) : 0;
            const int64_t rn = r0 + GROUP;
            const int nx = (rn < (int64_t)n_tiles * TILE_N) ? img_of[rn] : -1;
            if (im >= 0 && nx != im) mask |= 1 << g;
        }
        mt.end_mask = mask;
        mt.pad[0] = mt.pad[1] = mt.pad[2] = 0;
        *reinterpret_cast<TileMeta*>(tile + EXT_BYTES) = mt;
    }
}

// min / max of the real rows' norms (chooses C and checks that E fits the extension bytes)
__global__ void norm_minmax_kernel(const int32_t* __restrict__ norm, int64_t n, int* __restrict__ out /* [min, max] */) {
    int lo = INT_MAX, hi = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = norm[i];
        if (v >= 0) { lo = min(lo, v); hi = max(hi, v); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], lo);
        atomicMax(&out[1], hi);
    }
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
// End of an image: ratio test of the most favourable reading of (best, best of the other groups).
__device__ __noinline__ void finalize_image(const Params& p, int lane, bool row_valid, int Q, int64_t qrow, int M1, int M2,
                                            int pos, int img, int nval, const Chunk& ch) {
    bool pass = false;
    if (row_valid && nval >= 2 && img >= ch.img_begin && img < ch.img_end) {
        const int d0lo = max(0, Q - (2 * M1 - p.key_c + 1));          // smallest possible best distance^2
        const int d1hi = M2 < 0 ? (1 << 30) : Q - (2 * M2 - p.key_c);  // largest possible second-best distance^2
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
                p.cand[idx] = make_int4(p.q_row_base + (int)qrow, pos * GROUP, M2, img);
            else
                *p.overflow = 1;
        }
    }
}

template <int MODE>
__device__ __forceinline__ int group_max(const uint32_t (&v)[32]) {
    if (MODE == 3) return (int)(v[0] | v[31]);  // diagnostic: no arithmetic
    int l[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) l[i] = max3((int)v[3 * i], (int)v[3 * i + 1], (int)v[3 * i + 2]);
    l[10] = max((int)v[30], (int)v[31]);
    const int a0 = max3(l[0], l[1], l[2]), a1 = max3(l[3], l[4], l[5]), a2 = max3(l[6], l[7], l[8]);
    return max3(max3(a0, a1, a2), l[9], l[10]);
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) sift_tc2_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int warp = warp_id_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t leader = lane == 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < B_STAGES; ++i) { mbar_init(&s.b_full[i], 1); mbar_init(&s.b_empty[i], 1 + EPI_WARPS); }
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
        // ================= TMA producer: one map tile (8 KB) + its extension bytes and metadata per stage ==========
        uint32_t bn = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Chunk ch = p.chunks[item / p.n_qblocks];
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                mbar_wait(&s.b_empty[bs], bph ^ 1);
                mbar_arrive_expect_tx_if(leader, &s.b_full[bs], B_TILE_BYTES + EXT_META_BYTES);
                bulk_g2s_if(leader, s.stage[bs], p.m_desc + (size_t)t * B_TILE_BYTES, B_TILE_BYTES, &s.b_full[bs]);
                bulk_g2s_if(leader, s.stage[bs] + B_TILE_BYTES, p.m_ext + (size_t)t * EXT_META_BYTES, EXT_META_BYTES, &s.b_full[bs]);
            }
        }
    } else if (warp > EPI_WARPS) {
        // ================= MMA issuers: A from tensor memory, B from shared memory ==================================
        // Warp EPI_WARPS+1+s issues the tiles that land in accumulator stage s.  One warp alone spends ~580 cycles per tile
        // on its 4 mbarrier waits + 15 MMAs + 4 commits, more than the 500 cycles the tensor pipe needs for the tile.
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
                {
                    const uint64_t b_desc = umma_desc_sw128(smem_u32(s.stage[bs]));
                    const uint64_t e_desc = umma_desc_noswizzle(smem_u32(s.stage[bs] + B_TILE_BYTES), 128, 256);
#pragma unroll
                    for (int a = 0; a < ATILES; ++a) {
                        // each (stage, A tile) accumulator has its own full/empty pair: its epilogue warps start as soon
                        // as its 5 MMAs retire, and hand it back ~2/3 of a tile earlier than a per-stage barrier would
                        mbar_wait(&s.acc_empty[as][a], aph ^ 1);
                        tc_fence_after();
                        const uint32_t d = tmem_base + ACC_COL0 + (as * ATILES + a) * TILE_N;
                        const uint32_t ta = tmem_base + a * A_COLS;
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_i8_ts_if(leader, d, ta + 8 * k, b_desc + 2 * k, idesc, k > 0);
                        umma_i8_ts_if(leader, d, ta + 32, e_desc, idesc, 1);
                        tc_commit_if(leader, &s.acc_full[as][a]);
                    }
                    tc_commit_if(leader, &s.b_empty[bs]);
                }
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
            // ---- this thread's query row -> tensor memory (row = TMEM lane, 4 K-bytes per column) + extension constants
            {
                uint32_t w[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 x = *reinterpret_cast<const uint4*>(p.q_desc + sw128_chunk_offset(qrow, c));
                    w[4 * c] = x.x; w[4 * c + 1] = x.y; w[4 * c + 2] = x.z; w[4 * c + 3] = x.w;
                }
                const uint32_t e[8] = {0xFFFFFF01u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu,
                                       0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
                mbar_wait(&s.a_free, (item_n & 1) ^ 1);   // every MMA of the previous item has read its A tiles
                tc_fence_after();
                tmem_st32(t_lane + atile * A_COLS, w);
                tmem_st8(t_lane + atile * A_COLS + 32, e);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.a_full);
            }
            const int Q = p.q_norm[qrow];
            const bool row_valid = Q >= 0;
            int M1 = -1, M2 = -1, pos = 0;
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                const uint32_t as = bn % ACC_STAGES, aph = (bn / ACC_STAGES) & 1;
                mbar_wait(&s.acc_full[as][atile], aph);
                tc_fence_after();
                const uint32_t tcol = t_lane + ACC_COL0 + (as * ATILES + atile) * TILE_N;
                uint32_t v[GROUPS][32];
                if (MODE != 4) {
#pragma unroll
                    for (int g = 0; g < GROUPS; ++g) tmem_ld32(tcol + g * GROUP, v[g]);
                    tmem_wait_all();
#pragma unroll
                    for (int g = 0; g < GROUPS; ++g) tmem_fence_regs(v[g]);
                } else {
#pragma unroll
                    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[g][i] = 0;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.acc_empty[as][atile]);
                mbar_wait(&s.b_full[bs], bph);  // already complete (the MMAs consumed the stage); orders the metadata read
                const TileMeta& mt = *reinterpret_cast<const TileMeta*>(s.stage[bs] + B_TILE_BYTES + EXT_BYTES);
                const int end_mask = mt.end_mask;
#pragma unroll
                for (int g = 0; g < GROUPS; ++g) {
                    const int gm = group_max<MODE>(v[g]);
                    pos = gm > M1 ? t * GROUPS + g : pos;
                    M2 = max(M2, min(M1, gm));
                    M1 = max(M1, gm);
                    if (end_mask & (1 << g)) {
                        finalize_image(p, lane, row_valid, Q, qrow, M1, M2, pos, mt.img[g], mt.nval[g], ch);
                        M1 = -1; M2 = -1;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.b_empty[bs]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// exact resolution of candidates.  One warp per candidate; lane l owns map row l of the winning 32-row group.
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

__global__ void sift_fixup2_kernel(const int4* __restrict__ cand, const unsigned int* __restrict__ cand_count,
                                   unsigned int cand_cap, const uint8_t* __restrict__ q_desc,
                                   const int32_t* __restrict__ q_norm, const int32_t* __restrict__ q_frame,
                                   const uint8_t* __restrict__ m_desc, const int32_t* __restrict__ m_norm,
                                   const int64_t* __restrict__ m_dst_off, int key_c, double ratio, int n_images,
                                   int32_t* __restrict__ counts, unsigned int* __restrict__ n_rescans) {
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
        const int64_t mrow = (int64_t)cd.y + lane;
        const int v = exact_key(q, m_desc, mrow, m_norm[mrow]);
        int k1, k2;
        warp_top2(v, INT_MIN, k1, k2);
        bool verdict;
        bool decided = true;
        if (cd.z < 0) {  // the image has no other group
            verdict = k2 > INT_MIN && ratio_pass(Q - k1, Q - k2, ratio);
        } else {
            const int lo = min(2 * cd.z - key_c, Q), hi = min(lo + 1, Q);  // the best key of the other groups is lo or hi (no key exceeds Q)
            if (hi <= k2) {
                verdict = ratio_pass(Q - k1, Q - k2, ratio);
            } else {
                // every reading of (best, second) that is consistent with the bounds
                bool any = false, all = true;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int ko = lo + r;
                    if (ko <= k1) {
                        const bool x = ratio_pass(Q - k1, Q - max(k2, ko), ratio);
                        any |= x; all &= x;
                    } else {
                        const bool x = ratio_pass(Q - ko, Q - k1, ratio), y = ratio_pass(Q - ko, Q - ko, ratio);
                        any |= x | y; all &= x & y;
                    }
                }
                verdict = all;
                decided = (any == all);
            }
        }
        if (!decided) {  // rare: exact scan of the whole image (warp-uniform branch)
            const int64_t r0 = m_dst_off[cd.w], r1 = m_dst_off[cd.w + 1];
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
        if (lane == 0 && verdict) atomicAdd(&counts[(int64_t)q_frame[qrow] * n_images + cd.w], 1);
    }
}

}  // namespace sifttc2
}  // namespace mcl
