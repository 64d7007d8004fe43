// tc9: one tensor-core kernel family for the two paths whose rows are 256-288 bytes wide, built like sift_tc4.
//
//   KIND_ORB  replaces Matcher.ORBMatch's brute force (Matcher.py:169-194; north_star mode: knnMatch(k=2) + ratio lambda) --
//             the popc pipe caps the SIMT kernel at 2.2 M (500 x 500) pairs/s (98.5 % busy, profiles/r1_orb_knn_ncu.md).
//             A 256-bit descriptor becomes 256 signed bytes, bit b -> 2b - 1: then a.b = 256 - 2 Hamming(a, b) EXACTLY, no
//             norm term, and the distance matrix is a dense s8 x s8 -> s32 contraction with K = 256:
//             tcgen05.mma kind::i8, 8 k-steps of 32 (15x the popc roofline at the int8 tensor peak).
//   KIND_BOW  replaces scipy.cluster.vq.vq of Matcher.BOWMatch (Matcher.py:250): argmax_j (o.c_j - |c_j|^2 / 2) with o = SIFT
//             rows (integers <= 255: exact in fp16) and c = float32 centroids rounded to fp16 (relative error 2^-11 per
//             component; everything is non-negative, so the error of a key is at most 2^-11 o.c).  The norm term rides
//             in a K extension: query rows carry (256, 256) and code-book rows (v1, v2) with 256 (v1 + v2) = C - |c|^2 / 2
//             to 2^-22 relative, C chosen per index so that every key is positive -- positive float32 order like int32, so
//             both kinds share the 3-input integer max tree.  tcgen05.mma kind::f16, 8 + 1 k-steps of 16.  The approximate
//             keys only select a 32-word group; bow_resolve_kernel then evaluates that group in float32 exactly as before.
//
// Pipeline (per CTA, persistent): TWO query tiles of 128 rows stationary in TENSOR MEMORY (64 / 72 columns each: row = lane,
// 4 K-bytes per column), 128-column accumulators in the rest (BOW: two, one per query tile -- while the epilogue warps drain
// one, the tensor pipe fills the other, 8-9 MMAs of 64 cycles each, twice sift_tc4's time per tile, so two tiles suffice where
// sift_tc4 needs three; ORB: three, taken round robin -- two measured 78 % of the tensor pipe), 128-row map
// tiles (two SWIZZLE_128B atoms of 16 KB + a 4 KB no-swizzle extension slice) streamed by ONE linear bulk copy each into a ring
// of 6 (ORB) / 5 (BOW) stages; CTA pairs (clusters of 2) fetch every tile once and multicast it.
// Warps 0-7: epilogue (query tile = warp / 4, TMEM lane quadrant = warp % 4), warp 8: TMA producer, warp 9: MMA issuer.
// Epilogue per 128 x 128 tile and thread: two tcgen05.ld.x64, four 32-column max trees, top-2 over GROUP maxima with the
// best group's position.  Keys are exact (ORB) or approximate (BOW), there are no norm brackets.
//   ORB, end of an image: best key H1 is exact; the second best is >= H2 (the best other group); a row whose ratio test
//        fails with (H1, H2) fails with the true second best too and is dropped; the others go to a candidate list and
//        orb_fixup_kernel recomputes the 32 rows of the best group with xor + popc on the packed descriptors.
//   BOW, end of a code-book range: (the three best groups, their keys, the fourth-best key, and the keys of the best group's
//        four 8-word sub-blocks) per (range, location, row).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "sift_tc3.cuh"   // Chunk, group_max, warp_top2, NONE

namespace mcl {
namespace tc9 {

using sifttc3::Chunk;
using sifttc3::NONE;

constexpr int KIND_ORB = 0, KIND_BOW = 1;
constexpr int TILE_N = 128;                   // map rows per tile (= MMA N)
constexpr int GROUP = 32;                     // rows per group whose maximum the epilogue tracks
constexpr int QTILE = 128;
constexpr int ATILES = 2;
constexpr int QBLK = ATILES * QTILE;          // 256 query rows per work item
constexpr int ATOM_BYTES = TILE_N * 128;      // 16 KB: 128 rows x 128 K-bytes, SWIZZLE_128B
constexpr int EXT_BYTES = TILE_N * 32;        // 4 KB: 128 rows x 32 K-bytes, no swizzle (8 x 16 B core matrices, LBO 128, SBO 256)
constexpr int EPI_WARPS = ATILES * 4;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 512;
constexpr int META_CAP = 512;                 // tile records staged per chunk (16 KB); longer chunks read them from global memory

template <int KIND> struct Shape;
// A_COLS: TMEM columns per query tile (4 K-bytes per column).  ACCS accumulators of 128 columns fill the rest of the 512
// columns: ORB has room for three, used round robin by the two query tiles (one spare: the MMAs of a tile need not wait for
// the epilogue of the tile before last), BOW for two.
template <> struct Shape<KIND_ORB> { static constexpr int KSTEPS = 8, TILE_BYTES = 2 * ATOM_BYTES, ROW_BYTES = 256, A_COLS = 64, ACCS = 3, STAGES = 6; };
template <> struct Shape<KIND_BOW> { static constexpr int KSTEPS = 9, TILE_BYTES = 2 * ATOM_BYTES + EXT_BYTES, ROW_BYTES = 288, A_COLS = 72, ACCS = 2, STAGES = 5; };
constexpr int MAX_ACCS = 3;

// per 128-row map tile
struct TileRec {
    int32_t seg[4];        // segment (map image / location) of each 32-row group, -1 = padding group
    int32_t nvalid;        // real rows of each group, one byte per group (0..32)
    int32_t end_mask;      // bit g: group g is the last group of its segment
    int32_t plain;         // bit 0: all four groups full, one segment, no segment end before the tile's end;
                           // ORB, bit 8 + g: the image of group g has at least two descriptors (others count 0)
    int32_t first_group;   // index of group 0 within its segment (BOW: word = 32 * group index + column)
};
static_assert(sizeof(TileRec) == 32, "tile record size");

struct Params {
    const uint8_t* q_rows;    // query rows, row-major ROW_BYTES each, n_qblocks * QBLK rows (batch base); padding rows zero
    const int32_t* q_aux;     // ORB: popcount-independent validity (>= 0 real row, -1 padding).  BOW: same
    const int32_t* q_frame;   // ORB: frame of every query row
    const uint8_t* m_tiles;   // n_tiles x TILE_BYTES
    const TileRec* m_rec;     // n_tiles
    const Chunk* chunks;      // tile ranges; ORB: whole images [img_begin, img_end); BOW: img_begin = location, img_end = split index
    const int32_t* seg_rows;  // ORB: descriptors per image (images with < 2 count 0)
    int n_qblocks, n_chunks;
    double ratio;
    // ORB outputs
    int4* cand;               // {q row, first padded map row of the best group, best key of the other groups (NONE: none), image}
    unsigned int* cand_count;
    unsigned int cand_cap;
    int* overflow;
    uint8_t* dirty;           // [n_frames][n_images]
    int n_images;
    int q_row_base;
    // BOW outputs
    int4* out;                // [n_splits][n_loc][n_q][3]: {the three best groups, -}, {their key bits, fourth-best key bits},
                              // {key bits of the best group's four 8-word sub-blocks}
    int n_q, n_loc;
    int meta_cap;
    long long* diag;          // optional (MCL_TC9_DIAG): per CTA, cycles the MMA warp waited for {query tiles, map tiles, accumulators} + total
};

template <int KIND>
struct Smem {
    alignas(1024) uint8_t stage[Shape<KIND>::STAGES][Shape<KIND>::TILE_BYTES];
    alignas(16) TileRec rec[META_CAP];
    int16_t ratio_thr[260];   // ORB: largest best distance that passes the ratio test against a second-best distance d1
    alignas(8) uint64_t b_full[Shape<KIND>::STAGES];
    uint64_t b_empty[Shape<KIND>::STAGES];
    uint64_t acc_full[MAX_ACCS], acc_empty[MAX_ACCS];
    uint64_t a_full, a_free;
    uint32_t tmem_base;
};

__device__ __forceinline__ void umma_f16_ts_if(uint32_t pred, uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], %3, %4, p;\n\t}" ::"r"(pred),
        "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// columns c >= nv of a 32-column group never win (padding rows of the last group of a segment)
// (warp-uniform nv: full groups skip the 64 instructions -- the epilogue of a non-plain tile is one warp's dependent chain,
// and at 500-row images every fourth tile is one)
__device__ __forceinline__ void mask_group(uint32_t (&v)[32], int nv) {
    if (nv >= 32) return;
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = c < nv ? v[c] : (uint32_t)INT_MIN;
}

// maximum of a 32-column group through the maxima of its four 8-column sub-blocks (BOW keeps the best group's four)
__device__ __forceinline__ int group_max_sub(const uint32_t (&v)[32], int (&m8)[4]) {
#pragma unroll
    for (int b = 0; b < 4; ++b)
        m8[b] = max3(max3((int)v[8 * b], (int)v[8 * b + 1], (int)v[8 * b + 2]), max3((int)v[8 * b + 3], (int)v[8 * b + 4], (int)v[8 * b + 5]),
                     max((int)v[8 * b + 6], (int)v[8 * b + 7]));
    return max(max3(m8[0], m8[1], m8[2]), m8[3]);
}

template <int KIND, int CL>
__global__ void __launch_bounds__(THREADS, 1) tc9_kernel(const __grid_constant__ Params p) {
    using S = Shape<KIND>;
    constexpr int A_COLS = S::A_COLS, ACCS = S::ACCS, ACC_COL0 = TMEM_COLS - ACCS * TILE_N;
    static_assert(ATILES * A_COLS <= ACC_COL0, "query tiles and accumulators overlap in tensor memory");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem<KIND>& s = *reinterpret_cast<Smem<KIND>*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int warp = warp_id_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t leader = lane == 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S::STAGES; ++i) { mbar_init(&s.b_full[i], 1); mbar_init(&s.b_empty[i], CL); }
        for (int a = 0; a < MAX_ACCS; ++a) { mbar_init(&s.acc_full[a], 1); mbar_init(&s.acc_empty[a], 4); }
        mbar_init(&s.a_full, EPI_WARPS);
        mbar_init(&s.a_free, 1);
        mbar_fence_init();
    }
    if (warp == EPI_WARPS + 1) tmem_alloc<TMEM_COLS>(&s.tmem_base);
    if (KIND == KIND_ORB) {
        // Matcher.py:280 on Hamming distances reported as floats: (double)d0 < ratio * (double)d1.  As a table over d1: the
        // largest d0 that passes is ceil(ratio * d1) - 1 (the product is one IEEE multiplication, as in the fix-up kernels)
        for (int d1 = threadIdx.x; d1 <= 256; d1 += blockDim.x) {
            const double t = __dmul_rn(p.ratio, (double)d1);
            const double c = ceil(t) - 1.0;
            s.ratio_thr[d1] = (int16_t)(c < -1.0 ? -1.0 : (c > 1000.0 ? 1000.0 : c));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();   // the peer's barriers exist before any multicast copy or commit can reach them
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    // work items: (chunk, group of CL query blocks); CTA `crank` of the cluster takes query block group * CL + crank
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    // ORB: items are numbered query-group major (item = query group * n_chunks + chunk) and every cluster takes a CONTIGUOUS
    // range of them: consecutive items then share their query tiles, which stay in tensor memory -- the tiles are reloaded (and
    // the tensor pipe drained for it) only when the query group changes, once or twice per cluster instead of once per item
    // (1102 -> 1063 cycles per tile pair).  BOW keeps items dealt round robin, chunk major (item = chunk * n_qgroups + query
    // group): all clusters then stream the SAME code-book range at the same time; with contiguous ranges its wait for query
    // tiles fell from 5.1 to 1.0 % but the wait for map tiles rose from 4.8 to 7.8 % (93 MB of tiles read in 32 places at
    // once) -- 1294 -> 1309 cycles.
    constexpr bool CONTIG = KIND == KIND_ORB;
    const int n_qgroups = (p.n_qblocks + CL - 1) / CL;
    const int n_items = n_qgroups * p.n_chunks;
    const int n_workers = gridDim.x / CL, worker = blockIdx.x / CL;
    const int item_begin = CONTIG ? (int)((int64_t)n_items * worker / n_workers) : worker;
    const int item_end = CONTIG ? (int)((int64_t)n_items * (worker + 1) / n_workers) : n_items;
    const int item_step = CONTIG ? 1 : n_workers;
    auto qg_of = [&](int item) { return CONTIG ? item / p.n_chunks : item % n_qgroups; };
    auto ck_of = [&](int item) { return CONTIG ? item % p.n_chunks : item / n_qgroups; };
    constexpr uint16_t ALL_CTAS = (1u << CL) - 1u;

    if (warp == EPI_WARPS) {
        // ================= TMA producer: one 128-row map tile per stage ===============================================
        uint32_t bn = 0;
        for (int item = item_begin; item < item_end; item += item_step) {
            const Chunk ch = p.chunks[ck_of(item)];
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const uint32_t bs = bn % S::STAGES, bph = (bn / S::STAGES) & 1;
                mbar_wait(&s.b_empty[bs], bph ^ 1);
                mbar_arrive_expect_tx_if(leader, &s.b_full[bs], S::TILE_BYTES);
                const uint8_t* src = p.m_tiles + (size_t)t * S::TILE_BYTES;
                if (CL == 1) {
                    bulk_g2s_if(leader, s.stage[bs], src, S::TILE_BYTES, &s.b_full[bs]);
                } else {   // this CTA's share of the tile, delivered to every CTA of the cluster
                    constexpr uint32_t part = S::TILE_BYTES / CL;
                    bulk_g2s_multicast_if(leader, s.stage[bs] + crank * part, src + crank * part, part, &s.b_full[bs], ALL_CTAS);
                }
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ================= MMA issuer: A from tensor memory, B from shared memory, N = 128 ==============================
        constexpr uint32_t idesc = KIND == KIND_ORB ? umma_idesc(/*c=S32*/ 2, /*a=s8*/ 1, /*b=s8*/ 1, QTILE, TILE_N)
                                                    : umma_idesc(/*c=F32*/ 1, /*a=f16*/ 0, /*b=f16*/ 0, QTILE, TILE_N);
        uint32_t bn = 0, load_n = 0;
        int prev_qg = -1;
        long long w_a = 0, w_b = 0, w_acc = 0;
#ifdef MCL_TC_DIAG   // compiled out by default: even disabled at run time the probe's branches slow the issue loop (BOW: 4 %)
        const bool diag = p.diag != nullptr;
#else
        constexpr bool diag = false;
#endif
        const long long t_begin = diag ? clock64() : 0;
        for (int item = item_begin; item < item_end; item += item_step) {
            const int qg = qg_of(item);
            const Chunk ch = p.chunks[ck_of(item)];
            long long t0 = diag ? clock64() : 0;
            if (qg != prev_qg) {   // new query tiles
                mbar_wait(&s.a_full, load_n & 1);
                if (diag) w_a += clock64() - t0;
                tc_fence_after();
                prev_qg = qg;
            }
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const uint32_t bs = bn % S::STAGES, bph = (bn / S::STAGES) & 1;
                if (diag) t0 = clock64();
                mbar_wait(&s.b_full[bs], bph);
                if (diag) w_b += clock64() - t0;
                const uint32_t sb = smem_u32(s.stage[bs]);
                const uint64_t d0 = umma_desc_sw128(sb), d1 = umma_desc_sw128(sb + ATOM_BYTES);
                const uint64_t de = umma_desc_noswizzle(sb + 2 * ATOM_BYTES, 128, 256);
#pragma unroll
                for (int a = 0; a < ATILES; ++a) {
                    // the n-th (tile, query tile) product of this CTA goes to accumulator n mod ACCS
                    const uint32_t n = ATILES * bn + a, x = n % ACCS, use = n / ACCS;
                    if (diag) t0 = clock64();
                    mbar_wait(&s.acc_empty[x], (use & 1) ^ 1);
                    if (diag) w_acc += clock64() - t0;
                    tc_fence_after();
                    const uint32_t d = tmem_base + ACC_COL0 + x * TILE_N;
                    const uint32_t ta = tmem_base + a * A_COLS;
#pragma unroll
                    for (int k = 0; k < S::KSTEPS; ++k) {
                        const uint64_t bd = k < 4 ? d0 + 2 * k : (k < 8 ? d1 + 2 * (k - 4) : de);
                        if (KIND == KIND_ORB) umma_i8_ts_if(leader, d, ta + 8 * k, bd, idesc, k > 0);
                        else umma_f16_ts_if(leader, d, ta + 8 * k, bd, idesc, k > 0);
                    }
                    tc_commit_if(leader, &s.acc_full[x]);
                }
                if (CL == 1) tc_commit_if(leader, &s.b_empty[bs]);
                else tc_commit_multicast_if(leader, &s.b_empty[bs], ALL_CTAS);
                __syncwarp();
            }
            if (item + item_step >= item_end || qg_of(item + item_step) != qg) {   // the last item that reads these query tiles
                tc_commit_if(leader, &s.a_free);
                ++load_n;
            }
            __syncwarp();
        }
        if (diag && lane == 0) {
            long long* o = p.diag + 5 * blockIdx.x;
            o[0] = w_a; o[1] = w_b; o[2] = w_acc; o[3] = clock64() - t_begin; o[4] = bn;
        }
    } else {
        // ================= epilogue warps ===========================================================================
        const int atile = warp >> 2;
        const int lane_base = (warp & 3) * 32;
        const int row_in_blk = atile * QTILE + lane_base + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)lane_base << 16);
        uint32_t bn = 0, load_n = 0;
        int prev_qg = -1;
        // ORB: the query row of the NEXT query group is fetched into registers while the current one is being matched (the
        // tensor pipe idles from the last MMA on a set of query tiles until the next set is in tensor memory).  BOW measured
        // slower with the row held across the tile loop (72 more live registers in an epilogue that already holds the whole
        // accumulator): it fetches when the group changes.
        constexpr bool PREFETCH = KIND == KIND_ORB;
        constexpr int QW = S::ROW_BYTES / 4;   // 32-bit words per query row
        uint32_t wq[QW];
        int aux_next = -1;
        auto fetch = [&](int qg) {
            const int qb = min(qg * CL + (int)crank, p.n_qblocks - 1);
            const int64_t qrow = (int64_t)qb * QBLK + row_in_blk;
            const uint4* src = reinterpret_cast<const uint4*>(p.q_rows + (size_t)qrow * S::ROW_BYTES);
#pragma unroll
            for (int c = 0; c < QW / 4; ++c) {
                const uint4 x = __ldg(src + c);
                wq[4 * c] = x.x; wq[4 * c + 1] = x.y; wq[4 * c + 2] = x.z; wq[4 * c + 3] = x.w;
            }
            aux_next = p.q_aux[qrow];
        };
        static_assert(!PREFETCH || CONTIG, "the prefetch below assumes contiguous item ranges");
        if (PREFETCH && item_begin < item_end) fetch(qg_of(item_begin));
        bool row_valid = false;
        for (int item = item_begin; item < item_end; item += item_step) {
            const int qg = qg_of(item);
            const Chunk ch = p.chunks[ck_of(item)];
            const int qb_raw = qg * CL + (int)crank;
            const int qb = min(qb_raw, p.n_qblocks - 1);   // an odd tail: the second CTA repeats the last block and reports nothing
            const int64_t qrow = (int64_t)qb * QBLK + row_in_blk;
            if (qg != prev_qg) {   // this thread's query row -> tensor memory (row = TMEM lane, 4 K-bytes per column)
                if (!PREFETCH) fetch(qg);
                mbar_wait(&s.a_free, (load_n & 1) ^ 1);   // every MMA on the previous query tiles has retired
                tc_fence_after();
                tmem_st32(t_lane + atile * A_COLS, reinterpret_cast<uint32_t(&)[32]>(wq[0]));
                tmem_st32(t_lane + atile * A_COLS + 32, reinterpret_cast<uint32_t(&)[32]>(wq[32]));
                if (S::ROW_BYTES > 256) tmem_st8(t_lane + atile * A_COLS + 64, reinterpret_cast<uint32_t(&)[8]>(wq[QW - 8]));
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.a_full);
                ++load_n;
                prev_qg = qg;
                row_valid = aux_next >= 0 && qb_raw < p.n_qblocks;
                if (PREFETCH && (int64_t)(qg + 1) * p.n_chunks < item_end) fetch(qg + 1);   // the group this cluster takes next
            }
            // running top keys over the GROUP maxima of the current segment: ORB keeps the best group and the best key of the
            // others; BOW keeps the THREE best groups and the fourth-best key (a near-tie between up to three groups is then
            // resolved inside those groups instead of over the whole code book)
            int H1 = NONE, H2 = NONE, H3 = NONE, H4 = NONE, pos = 0, pos2 = 0, pos3 = 0;
            int S0 = NONE, S1 = NONE, S2 = NONE, S3 = NONE;   // BOW: maxima of the best group's four 8-word sub-blocks
            auto join = [&](int gm, int id, const int (&m8)[4]) {
                const bool g1 = gm > H1;
                if (KIND == KIND_BOW) {
                    const bool g2 = gm > H2, g3 = gm > H3;
                    H4 = max(H4, g3 ? H3 : gm);
                    H3 = g2 ? H2 : (g3 ? gm : H3);
                    pos3 = g2 ? pos2 : (g3 ? id : pos3);
                    H2 = g1 ? H1 : (g2 ? gm : H2);
                    pos2 = g1 ? pos : (g2 ? id : pos2);
                    S0 = g1 ? m8[0] : S0; S1 = g1 ? m8[1] : S1; S2 = g1 ? m8[2] : S2; S3 = g1 ? m8[3] : S3;
                } else {
                    H2 = max(H2, g1 ? H1 : gm);
                }
                pos = g1 ? id : pos;
                H1 = max(H1, gm);
            };
            // the chunk's tile records, staged once per work item (named barrier 1 among the epilogue warps)
            const int n_rec = ch.tile_end - ch.tile_begin;
            const int4* rp = reinterpret_cast<const int4*>(p.m_rec + ch.tile_begin);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            const bool staged = n_rec <= p.meta_cap;
            if (staged) {
                int4* dst = reinterpret_cast<int4*>(s.rec);
                for (int i = threadIdx.x; i < n_rec * 2; i += EPI_WARPS * 32) dst[i] = __ldg(rp + i);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            const uint32_t rec_s = smem_u32(s.rec);
            auto ldrec = [&](int i) {
                int4 r;
                if (staged) asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(rec_s + 16u * (uint32_t)i));
                else r = __ldg(rp + i);
                return r;
            };
            // end of a segment: ORB -> ratio gate + candidate; BOW -> the range's record for this row
            auto finalize = [&](int seg, int t, bool two_rows) {
                if (KIND == KIND_ORB) {
                    // warp-uniform: the image has at least two descriptors (tile record: no global load on this path), and this
                    // chunk owns the image (neighbouring chunks share the tile in which one's last image ends and the other's
                    // first begins)
                    const bool counted = two_rows && seg >= ch.img_begin && seg < ch.img_end;
                    bool pass = false;
                    if (row_valid && counted) {
                        // a.b = 256 - 2 d.  The best distance is exact; the second best is at most that of the best OTHER
                        // group (unknown when the image has a single group: then the fix-up decides)
                        const int d0 = (256 - H1) >> 1;
                        pass = H2 <= NONE || d0 <= (int)s.ratio_thr[(256 - H2) >> 1];
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, pass);
                    if (bal) {
                        unsigned base = 0;
                        if (lane == 0) base = atomicAdd(p.cand_count, (unsigned)__popc(bal));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (pass) {
                            const unsigned idx = base + __popc(bal & ((1u << lane) - 1u));
                            if (idx < p.cand_cap) p.cand[idx] = make_int4(p.q_row_base + (int)qrow, pos * GROUP, H2, seg);
                            else {   // list full: this (frame, image) pair is recomputed by the exact SIMT kernel
                                p.dirty[(int64_t)p.q_frame[qrow] * p.n_images + seg] = 1;
                                *p.overflow = 1;
                            }
                        }
                    }
                } else {
                    // ch.img_begin = location, ch.img_end = index of this code-book range among the location's splits
                    if (row_valid) {
                        int4* o = p.out + 3 * (((size_t)ch.img_end * p.n_loc + ch.img_begin) * p.n_q + (p.q_row_base + qrow));
                        o[0] = make_int4(pos, pos2, pos3, 0);
                        o[1] = make_int4(H1, H2, H3, H4);
                        o[2] = make_int4(S0, S1, S2, S3);
                    }
                }
                H1 = NONE; H2 = NONE; H3 = NONE; H4 = NONE;
                S0 = NONE; S1 = NONE; S2 = NONE; S3 = NONE;
            };
#pragma unroll 1
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const int ri = (t - ch.tile_begin) * 2;
                const int4 r0 = ldrec(ri), r1 = ldrec(ri + 1);   // r0 = seg[4]; r1 = {nvalid, end_mask, plain, first_group}
                const uint32_t n = ATILES * bn + atile, x = n % ACCS, use = n / ACCS;
                mbar_wait(&s.acc_full[x], use & 1);
                tc_fence_after();
                const uint32_t tcol = t_lane + ACC_COL0 + x * TILE_N;
                // BOW reads both halves of the accumulator before anything is computed: the accumulator goes back to the MMA warp
                // one tcgen05.ld latency after it was filled (one accumulator per query tile -- the time to this arrive is on
                // the tensor pipe's critical path).  ORB has a spare accumulator and 64 registers of prefetched query row:
                // it reads and reduces half by half.
                constexpr bool WHOLE = KIND == KIND_BOW;
                uint32_t v[2][32], u[WHOLE ? 2 : 1][32];
                const bool plain = (r1.z & 1) != 0;   // warp-uniform
                int sub[4][4];
                int g0, g1, g2, g3;
                if (WHOLE) {
                    tmem_ld64(tcol, reinterpret_cast<uint32_t(&)[64]>(v));
                    tmem_ld64(tcol + 64, reinterpret_cast<uint32_t(&)[64]>(u));
                    tmem_wait_all();
                    tmem_fence_regs(v[0]);
                    tmem_fence_regs(v[1]);
                    tmem_fence_regs(u[0]);
                    tmem_fence_regs(u[WHOLE ? 1 : 0]);
                    if (lane == 0) mbar_arrive(&s.acc_empty[x]);   // the accumulator is free
                    if (!plain) {
                        mask_group(v[0], r1.x & 255); mask_group(v[1], (r1.x >> 8) & 255);
                        mask_group(u[0], (r1.x >> 16) & 255); mask_group(u[WHOLE ? 1 : 0], (r1.x >> 24) & 255);
                    }
                    g0 = group_max_sub(v[0], sub[0]);
                    g1 = group_max_sub(v[1], sub[1]);
                    g2 = group_max_sub(u[0], sub[2]);
                    g3 = group_max_sub(u[WHOLE ? 1 : 0], sub[3]);
                } else {
                    tmem_ld64(tcol, reinterpret_cast<uint32_t(&)[64]>(v));
                    tmem_wait_all();
                    tmem_fence_regs(v[0]);
                    tmem_fence_regs(v[1]);
                    if (!plain) { mask_group(v[0], r1.x & 255); mask_group(v[1], (r1.x >> 8) & 255); }
                    g0 = KIND == KIND_BOW ? group_max_sub(v[0], sub[0]) : sifttc3::group_max<0>(v[0]);
                    g1 = KIND == KIND_BOW ? group_max_sub(v[1], sub[1]) : sifttc3::group_max<0>(v[1]);
                    tmem_ld64(tcol + 64, reinterpret_cast<uint32_t(&)[64]>(v));
                    tmem_wait_all();
                    tmem_fence_regs(v[0]);
                    tmem_fence_regs(v[1]);
                    if (lane == 0) mbar_arrive(&s.acc_empty[x]);   // both halves are out of tensor memory: the accumulator is free
                    if (!plain) { mask_group(v[0], (r1.x >> 16) & 255); mask_group(v[1], (r1.x >> 24) & 255); }
                    g2 = KIND == KIND_BOW ? group_max_sub(v[0], sub[2]) : sifttc3::group_max<0>(v[0]);
                    g3 = KIND == KIND_BOW ? group_max_sub(v[1], sub[3]) : sifttc3::group_max<0>(v[1]);
                }
                // group position: ORB = padded row / 32 over the whole map; BOW = group index within the location
                const int gbase = KIND == KIND_ORB ? 4 * t : r1.w;
                if (plain) {
                    join(g0, gbase, sub[0]); join(g1, gbase + 1, sub[1]); join(g2, gbase + 2, sub[2]); join(g3, gbase + 3, sub[3]);
                    if (r1.y & 8) finalize(r0.x, t, (r1.z & (1 << 11)) != 0);
                } else {
                    const int gm[4] = {g0, g1, g2, g3};
                    const int sg[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (sg[g] < 0) continue;   // padding group
                        join(gm[g], gbase + g, sub[g]);
                        if (r1.y & (1 << g)) finalize(sg[g], t, (r1.z & (256 << g)) != 0);
                    }
                }
            }
            if (KIND == KIND_BOW && H1 > NONE) finalize(ch.img_begin, ch.tile_end - 1, true);   // a range that ends inside its location
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();   // no CTA leaves while its peer may still multicast into it
    if (warp == EPI_WARPS + 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------------
// ORB: map / query preparation and the exact fix-up
// ------------------------------------------------------------------------------------------------------------------------
// bit b of a packed descriptor -> signed byte 2b - 1 (0x01 / 0xFF); 4 bits -> one 32-bit word
__device__ __forceinline__ uint32_t pm1_word(uint32_t nib) {
    // spread 4 bits to the low bit of 4 bytes, then byte = 2b - 1 = b ? 0x01 : 0xFF: (b ^ 1) * 0xFE + 1 ... per byte: 0xFF - 0xFE*b
    const uint32_t b = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
    return 0xFFFFFFFFu - 0xFEu * b;   // no carries: each byte is 0xFF - 0xFE*b
}

// Map rows: packed (rows, 32) u8 in source order -> tc9 tiles (two SW128 atoms of +-1 bytes), the packed rows again in
// padded order (for the fix-up), and the image of every padded row.  One warp per source row.
__global__ void orb_map_prep_kernel(const uint8_t* __restrict__ src, int64_t n_rows, const int64_t* __restrict__ src_off,
                                    const int64_t* __restrict__ dst_off, int n_img, uint8_t* __restrict__ tiles,
                                    uint8_t* __restrict__ packed, int32_t* __restrict__ img_of) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n_rows; row += n_warps) {
        // segment lookup
        int lo = 0, hi = n_img;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (src_off[mid] <= row) lo = mid; else hi = mid; }
        const int64_t drow = dst_off[lo] + (row - src_off[lo]);
        const uint8_t byte = src[row * 32 + lane];          // lane l: bits 8l .. 8l+7 = K-bytes 8l .. 8l+7
        packed[drow * 32 + lane] = byte;
        const uint32_t w0 = pm1_word(byte & 15u), w1 = pm1_word(byte >> 4);
        uint8_t* tile = tiles + (size_t)(drow >> 7) * (2 * ATOM_BYTES);
        const int r = (int)(drow & 127), kb = lane * 8;     // K-byte offset 0..255
        uint8_t* atom = tile + (kb >> 7) * ATOM_BYTES;
        const int kk = kb & 127;
        *reinterpret_cast<uint2*>(atom + sw128_chunk_offset(r, kk >> 4) + (kk & 15)) = make_uint2(w0, w1);
        if (lane == 0) img_of[drow] = lo;
    }
}

// Query rows: packed (rows, 32) -> row-major 256 signed bytes, validity flag and frame of every row.  One warp per row.
__global__ void orb_query_prep_kernel(const uint8_t* __restrict__ src, int64_t n_rows, const int64_t* __restrict__ q_off, int n_frames,
                                      uint8_t* __restrict__ rows, int32_t* __restrict__ aux, int32_t* __restrict__ frame_of) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n_rows; row += n_warps) {
        const uint8_t byte = src[row * 32 + lane];
        *reinterpret_cast<uint2*>(rows + row * 256 + lane * 8) = make_uint2(pm1_word(byte & 15u), pm1_word(byte >> 4));
        if (lane == 0) {
            int lo = 0, hi = n_frames;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (q_off[mid] <= row) lo = mid; else hi = mid; }
            aux[row] = 0;
            frame_of[row] = lo;
        }
    }
}

__global__ void orb_build_rec_kernel(const int32_t* __restrict__ img_of, const int64_t* __restrict__ src_off,
                                     const int64_t* __restrict__ dst_off, int n_tiles, TileRec* __restrict__ rec) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    TileRec r;
    int nvalid = 0, end_mask = 0;
    bool plain = true;
    int two_rows = 0;
    for (int g = 0; g < 4; ++g) {
        const int64_t r0 = (int64_t)t * TILE_N + g * GROUP;
        const int im = img_of[r0];
        r.seg[g] = im;
        int nv = 0;
        if (im >= 0) {
            const int64_t rows = src_off[im + 1] - src_off[im];
            nv = (int)min((int64_t)GROUP, rows - (r0 - dst_off[im]));
            if (rows >= 2) two_rows |= 256 << g;
        }
        nvalid |= nv << (8 * g);
        const int64_t rn = r0 + GROUP;
        const int nx = (rn < (int64_t)n_tiles * TILE_N) ? img_of[rn] : -1;
        const bool end = im >= 0 && nx != im;
        if (end) end_mask |= 1 << g;
        plain = plain && im >= 0 && im == r.seg[0] && nv == GROUP && (!end || g == 3);
    }
    r.nvalid = nvalid;
    r.end_mask = end_mask;
    r.plain = (plain ? 1 : 0) | two_rows;
    r.first_group = 0;
    rec[t] = r;
}

__device__ __forceinline__ int hamming32B(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// One warp per candidate: lane l recomputes the Hamming distance to row l of the best group exactly; the second best is the
// smaller of the group's own runner-up and the best distance of the other groups.  Matcher.py:280 on Hamming distances
// reported as floats: float(d0) < ratio * float(d1).
__global__ void orb_fixup_kernel(const int4* __restrict__ cand, const unsigned int* __restrict__ cand_count, unsigned int cand_cap,
                                 const uint4* __restrict__ q_packed, const int32_t* __restrict__ q_frame,
                                 const uint4* __restrict__ m_packed, const int64_t* __restrict__ m_src_off,
                                 const int64_t* __restrict__ m_dst_off, double ratio, int n_images,
                                 int32_t* __restrict__ counts, unsigned long long* __restrict__ n_candidates) {
    const int lane = threadIdx.x & 31;
    const unsigned n = min(*cand_count, cand_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(n_candidates, (unsigned long long)n);
    const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n; c += n_warps) {
        const int4 cd = cand[c];
        const int64_t qrow = cd.x, g1row = cd.y;
        const int img = cd.w;
        const uint4 a0 = q_packed[qrow * 2], a1 = q_packed[qrow * 2 + 1];
        const int64_t mrow = g1row + lane;
        const bool real = (mrow - m_dst_off[img]) < (m_src_off[img + 1] - m_src_off[img]);
        int key = INT_MIN;
        if (real) key = 256 - 2 * hamming32B(a0, a1, m_packed[mrow * 2], m_packed[mrow * 2 + 1]);
        int k1, k2;
        sifttc3::warp_top2(key, INT_MIN, k1, k2);
        const int second = max(k2, cd.z <= NONE ? INT_MIN : cd.z);
        const bool verdict = second > INT_MIN && (double)((256 - k1) >> 1) < __dmul_rn(ratio, (double)((256 - second) >> 1));
        if (lane == 0 && verdict) atomicAdd(&counts[(int64_t)q_frame[qrow] * n_images + img], 1);
    }
}

__global__ void zero_where_mask_kernel(int32_t* __restrict__ counts, const uint8_t* __restrict__ mask, int64_t n,
                                       const int* __restrict__ only_if) {
    if (only_if && *only_if == 0) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (mask[i]) counts[i] = 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// BOW: code book -> fp16 tiles with the norm extension, query rows -> fp16 rows
// ------------------------------------------------------------------------------------------------------------------------
// One warp per word row; lane l owns dims 4l .. 4l+3.  key_c[loc] = C of the location (>= max |c|^2 / 2 + 1).
__global__ void bow_build_tiles_kernel(const float* __restrict__ voc, const int64_t* __restrict__ voc_off, int n_loc,
                                       const int32_t* __restrict__ tile_off, const float* __restrict__ key_c,
                                       uint8_t* __restrict__ tiles, TileRec* __restrict__ rec, int* __restrict__ bad) {
    constexpr int TB = Shape<KIND_BOW>::TILE_BYTES;
    const int lane = threadIdx.x & 31;
    const int wtile = blockIdx.x;
    int loc = 0;
    while (loc + 1 < n_loc && tile_off[loc + 1] <= wtile) ++loc;
    const int64_t v0 = voc_off[loc];
    const int n_words = (int)(voc_off[loc + 1] - v0);
    const int w0 = (wtile - tile_off[loc]) * TILE_N;
    uint8_t* tile = tiles + (size_t)wtile * TB;
    const float C = key_c[loc];
    for (int r = threadIdx.x >> 5; r < TILE_N; r += blockDim.x >> 5) {
        const int w = w0 + r;
        uint2 packed = make_uint2(0u, 0u);
        double nn = 0.0;
        bool ok = true;
        if (w < n_words) {
            const float4 c = reinterpret_cast<const float4*>(voc + (v0 + w) * 128)[lane];
            const float cc[4] = {c.x, c.y, c.z, c.w};
            __half h[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ok = ok && (cc[k] >= 0.0f) && (cc[k] < 256.0f);
                h[k] = __float2half_rn(fminf(fmaxf(cc[k], 0.0f), 255.875f));
                nn += (double)cc[k] * (double)cc[k];
            }
            packed.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
            packed.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        if (!ok) *bad = 1;
        // dims 4l..4l+3 = K-bytes 8l .. 8l+7 of the row (2 bytes per dim)
        const int kb = lane * 8;
        uint8_t* atom = tile + (kb >> 7) * ATOM_BYTES;
        const int kk = kb & 127;
        *reinterpret_cast<uint2*>(atom + sw128_chunk_offset(r, kk >> 4) + (kk & 15)) = packed;
        if (lane < 2) {
            // extension slice, canonical no-swizzle K-major: offset(row, 16-byte K chunk q) = (row/8)*256 + q*128 + (row%8)*16
            uint4 e = make_uint4(0u, 0u, 0u, 0u);
            if (lane == 0 && w < n_words) {
                const double V = ((double)C - 0.5 * nn) / 256.0;      // >= 1/256 by the choice of C
                const __half v1 = __double2half(V);
                const __half v2 = __double2half(V - (double)__half2float(v1));
                e.x = (uint32_t)__half_as_ushort(v1) | ((uint32_t)__half_as_ushort(v2) << 16);
            }
            *reinterpret_cast<uint4*>(tile + 2 * ATOM_BYTES + (r >> 3) * 256 + lane * 128 + (r & 7) * 16) = e;
        }
    }
    if (threadIdx.x == 0) {
        TileRec t;
        const int left = n_words - w0;
        int nvalid = 0;
        for (int g = 0; g < 4; ++g) {
            const int nv = max(0, min(GROUP, left - g * GROUP));
            t.seg[g] = nv > 0 ? loc : -1;
            nvalid |= nv << (8 * g);
        }
        t.nvalid = nvalid;
        t.end_mask = 0;                 // code-book ranges end with their chunk (tc9_kernel finalises there)
        t.plain = left >= TILE_N ? 1 : 0;
        t.first_group = (wtile - tile_off[loc]) * 4;
        rec[wtile] = t;
    }
}

// per location: C = max_j |c_j|^2 / 2 + 1 (one CTA per location)
__global__ void bow_key_c_kernel(const float* __restrict__ voc, const int64_t* __restrict__ voc_off, float* __restrict__ key_c) {
    __shared__ float red[8];
    const int loc = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t v0 = voc_off[loc], v1 = voc_off[loc + 1];
    float best = 0.f;
    for (int64_t w = v0 + warp; w < v1; w += blockDim.x >> 5) {
        const float4 c = reinterpret_cast<const float4*>(voc + w * 128)[lane];
        float nn = c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        best = fmaxf(best, nn);
    }
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
        key_c[loc] = 0.5f * m * 1.0001f + 2.0f;
    }
}

// float32 (n, 128) integer-valued query rows -> fp16 rows of 288 bytes: 128 halves + (256, 256, 0 x 14); aux[row] = 0; bad is
// set when a value is not an integer in [0, 255] (then nothing from the tensor-core pass can be trusted)
__global__ void bow_query_prep_kernel(const float* __restrict__ src, int64_t n_rows, uint8_t* __restrict__ rows, int32_t* __restrict__ aux,
                                      int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n_rows; row += n_warps) {
        const float4 f = reinterpret_cast<const float4*>(src + row * 128)[lane];
        const float v[4] = {f.x, f.y, f.z, f.w};
        __half h[4];
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ok = ok && rintf(v[k]) == v[k] && v[k] >= 0.f && v[k] <= 255.f;
            h[k] = __float2half_rn(v[k]);
        }
        if (!ok) *bad = 1;
        uint8_t* dst = rows + row * 288;
        *reinterpret_cast<uint2*>(dst + lane * 8) = make_uint2((uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
                                                               (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
        if (lane < 8) {
            const uint32_t one = 0x5C005C00u;   // (256.0h, 256.0h)
            reinterpret_cast<uint32_t*>(dst + 256)[lane] = lane == 0 ? one : 0u;
        }
        if (lane == 0) aux[row] = 0;
    }
}

}  // namespace tc9
}  // namespace mcl
