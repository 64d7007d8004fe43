// SIFT / integer-L2 matching on 5th-gen tensor cores, fourth generation.
//
// sift_tc3 is bound by TENSOR-MEMORY READ BANDWIDTH, not by the tensor pipe or the ALUs: halving the tcgen05.ld volume
// (diagnostic mode 6) lifts it from 76 % to 90 % of the int8 peak.  Per 64 map rows and 384 query rows it reads 96 KB of
// accumulators (inherent: 4 bytes per 128 MACs) plus 48 KB of A operand, because with the query block in tensor memory
// every MMA re-reads its 128 x 32-byte A slice and N = 64 needs 24 MMAs per 128 map rows.
// This generation issues N = 128 MMAs (12 per 128 map rows): the A-operand reads halve (144 -> 120 KB per 64 rows).
// Tensor memory then has room for one accumulator per A tile only (96 + 3 x 128 columns), so accumulators are single
// buffered and the three A tiles form the pipeline: while the epilogue warps of A tile a drain its accumulator, the
// tensor pipe works on the other two (2 x 268 cycles), which covers the drain (two 64-column loads + hand-back).
// Everything else (norm-sorted rows, bracketed pre-filter, candidate list, exact fix-up, metadata records per 64 rows)
// is sift_tc3's; the candidate format and sift_fixup3_kernel are shared.
//
// What measurements added on top (DESIGN.md section 4.1 has the numbers):
//  * the epilogue was the limiter (3 warps per SMSP, ALU pipe at half rate): tile records are prefetched with ld.shared
//    before the accumulator wait, a 128-row tile that lies in one image with close norms gets ONE bracket update
//    (TileMeta::merge), and the exact image-end ratio test sits behind an inline conservative gate (finalize_gate);
//  * every work item starts with a pipeline refill (the query tiles in tensor memory are single buffered), so chunks are
//    long: up to 448 tiles, their records staged in 48 KB of shared memory, the chunk set picked by a makespan estimate;
//  * CTA pairs (clusters of 2, template parameter CL) fetch every map tile once and multicast it: half the L2 -> SM traffic;
//  * MODE 5 stamps clock64 around the waits of one epilogue warp and of the MMA warp (how the above was found);
//    SS > 0 reads A tiles from shared memory instead (measured: no gain, kept as an experiment).
#pragma once
#include "common.cuh"
#include "sift_tc3.cuh"

namespace mcl {
namespace sifttc4 {

using sifttc3::Chunk;
using sifttc3::Params;          // chunks / tile indices are in units of TILE_N = 128 rows here; m_meta stays per 64 rows
using sifttc3::NONE;

constexpr int TILE_N = 128;
constexpr int HALF = 64;
constexpr int GROUP = 32;
constexpr int QTILE = 128;
constexpr int ATILES = 3;
constexpr int QBLK = ATILES * QTILE;
constexpr int A_COLS = 32;
constexpr int ACC_COL0 = 128;
constexpr int B_TILE_BYTES = TILE_N * 128;   // 16 KB
constexpr int B_STAGES = 8;
constexpr int EPI_WARPS = ATILES * 4;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 512;
constexpr int META_CAP = 1024;               // 64-row records staged per chunk (48 KB): chunks of up to 512 tiles

constexpr int A_TILE_BYTES = QTILE * 128;    // 16 KB

// SS = number of A tiles (query tiles) the MMAs read from SHARED memory instead of tensor memory.  With all three in tensor
// memory the kernel sits on the tensor-memory read bandwidth (per 128 map rows: 3 x 64 KB of accumulators for the epilogue
// + 3 x 16 KB of A operand = 240 KB per 768 MMA cycles = 312 B/clk) while shared memory idles at ~83 B/clk of its 128
// (B operand 3 x 16 KB + the TMA writes).  Moving ONE A tile to shared memory trades 16 KB of tensor-memory reads for 16 KB
// of shared-memory reads per B tile (292 / 104 B/clk).  The tile is double buffered across work items and loaded by the
// TMA warp with one linear bulk copy (the query rows are stored as SW128 atoms, like the map).
template <int SS>
struct SmemT {
    alignas(1024) uint8_t stage[B_STAGES][B_TILE_BYTES];
    alignas(1024) uint8_t a_tile[2][SS > 0 ? SS : 1][SS > 0 ? A_TILE_BYTES : 16];
    alignas(16) int4 meta[META_CAP * 3];
    alignas(8) uint64_t b_full[B_STAGES];
    uint64_t b_empty[B_STAGES];
    uint64_t acc_full[ATILES], acc_empty[ATILES];
    uint64_t a_full, a_free;
    uint64_t as_full[2], as_empty[2];
    uint32_t tmem_base;
};
using Smem = SmemT<0>;

// CL = CTAs per thread-block cluster (1 or 2).  With CL = 2 the two CTAs of a cluster work on the same chunk of map tiles
// for two different query blocks, in lockstep: every map tile is fetched from L2 ONCE, each CTA issuing one half of it as a
// multicast bulk copy that lands in both CTAs' shared memory (halves the L2 -> SM traffic, 6 TB/s at full speed, which is
// what pushes the kernel into the power cap).  A stage may be refilled only when BOTH CTAs' MMAs have consumed it: the
// MMA warps commit to the stage's empty barrier in both CTAs (tcgen05.commit multicast).
template <int MODE, int SS = 0, int CL = 1>
__global__ void __launch_bounds__(THREADS, 1) sift_tc4_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemT<SS>& s = *reinterpret_cast<SmemT<SS>*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int warp = warp_id_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t leader = lane == 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < B_STAGES; ++i) { mbar_init(&s.b_full[i], 1); mbar_init(&s.b_empty[i], CL); }
        for (int a = 0; a < ATILES; ++a) { mbar_init(&s.acc_full[a], 1); mbar_init(&s.acc_empty[a], 4); }
        mbar_init(&s.a_full, EPI_WARPS - 4 * SS);   // the warps whose A tile lives in tensor memory
        mbar_init(&s.a_free, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&s.as_full[i], 1); mbar_init(&s.as_empty[i], 1); }
        mbar_fence_init();
    }
    if (warp == EPI_WARPS + 1) tmem_alloc<TMEM_COLS>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();   // the peer's barriers exist before any multicast copy or commit can reach them
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    // work items: (chunk, group of CL query blocks); CTA `crank` of the cluster takes query block group * CL + crank
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    const int n_qgroups = (p.n_qblocks + CL - 1) / CL;
    const int n_items = n_qgroups * p.n_chunks;
    const int item0 = blockIdx.x / CL, item_step = gridDim.x / CL;
    constexpr uint16_t ALL_CTAS = (1u << CL) - 1u;

    if (warp == EPI_WARPS) {
        // ================= TMA producer: one 128-row map tile (16 KB) per stage ========================================
        uint32_t bn = 0, item_n = 0;
        for (int item = item0; item < n_items; item += item_step, ++item_n) {
            const Chunk ch = p.chunks[item / n_qgroups];
            if (SS > 0) {   // this item's shared-memory A tiles (buffer item_n & 1; free once the MMAs of item_n - 2 retired)
                const uint32_t ab = item_n & 1;
                mbar_wait(&s.as_empty[ab], ((item_n >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx_if(leader, &s.as_full[ab], SS * A_TILE_BYTES);
                const int qb = min((item % n_qgroups) * CL + (int)crank, p.n_qblocks - 1);
#pragma unroll
                for (int a = 0; a < SS; ++a)
                    bulk_g2s_if(leader, s.a_tile[ab][a], p.q_desc + ((size_t)qb * QBLK + (size_t)a * QTILE) * 128, A_TILE_BYTES,
                                &s.as_full[ab]);
            }
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                mbar_wait(&s.b_empty[bs], bph ^ 1);
                mbar_arrive_expect_tx_if(leader, &s.b_full[bs], B_TILE_BYTES);
                if (CL == 1) {
                    bulk_g2s_if(leader, s.stage[bs], p.m_desc + (size_t)t * B_TILE_BYTES, B_TILE_BYTES, &s.b_full[bs]);
                } else {   // this CTA's share of the tile, delivered to every CTA of the cluster
                    constexpr uint32_t part = B_TILE_BYTES / CL;
                    bulk_g2s_multicast_if(leader, s.stage[bs] + crank * part, p.m_desc + (size_t)t * B_TILE_BYTES + crank * part, part,
                                          &s.b_full[bs], ALL_CTAS);
                }
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ================= MMA issuer: A from tensor memory, B from shared memory, N = 128 ==============================
        constexpr uint32_t idesc = umma_idesc(/*c=S32*/ 2, /*a=u8*/ 0, /*b=u8*/ 0, QTILE, TILE_N);
        uint32_t bn = 0, item_n = 0;
        long long w_a = 0, w_b = 0, w_acc = 0, t0 = 0;
        // the wait-time probe is compiled in only with -DMCL_TC_DIAG: even disabled at run time its branches in this loop cost the
        // kernel 3.7 % (one issuing warp per query tile instead of one for all three was measured too: no difference)
#ifdef MCL_TC_DIAG
        const bool diag = p.diag != nullptr;
#else
        constexpr bool diag = false;
#endif
        const long long t_begin = diag ? clock64() : 0;
        for (int item = item0; item < n_items; item += item_step, ++item_n) {
            const Chunk ch = p.chunks[item / n_qgroups];
            if (diag) t0 = clock64();
            if (SS < ATILES) mbar_wait(&s.a_full, item_n & 1);
            if (SS > 0) mbar_wait(&s.as_full[item_n & 1], (item_n >> 1) & 1);
            if (diag) w_a += clock64() - t0;
            tc_fence_after();
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                long long tm[8];
                if (MODE == 5) tm[0] = clock64();
                if (diag) t0 = clock64();
                mbar_wait(&s.b_full[bs], bph);
                if (diag) w_b += clock64() - t0;
                if (MODE == 5) tm[1] = clock64();
                const uint64_t b_desc = umma_desc_sw128(smem_u32(s.stage[bs]));
                // round robin over the A tiles: accumulator a was filled first last time, so it is drained first; a blocking
                // try_wait wakes ~60 cycles after the arrive, a polling loop over three barriers costs ~150 per probe
#pragma unroll
                for (int a = 0; a < ATILES; ++a) {
                    if (diag) t0 = clock64();
                    mbar_wait(&s.acc_empty[a], (bn & 1) ^ 1);
                    if (diag) w_acc += clock64() - t0;
                    tc_fence_after();
                    if (MODE == 5) tm[2 + a] = clock64();
                    const uint32_t d = tmem_base + ACC_COL0 + a * TILE_N;
                    if (a < SS) {
                        const uint64_t a_desc = umma_desc_sw128(smem_u32(s.a_tile[item_n & 1][a < SS ? a : 0]));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_i8_if(leader, d, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0);
                    } else {
                        const uint32_t ta = tmem_base + a * A_COLS;
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_i8_ts_if(leader, d, ta + 8 * k, b_desc + 2 * k, idesc, k > 0);
                    }
                    tc_commit_if(leader, &s.acc_full[a]);
                }
                if (CL == 1) tc_commit_if(leader, &s.b_empty[bs]);
                else tc_commit_multicast_if(leader, &s.b_empty[bs], ALL_CTAS);
                __syncwarp();
                if (MODE == 5 && blockIdx.x == 0 && lane == 0 && bn >= 100 && bn < 164) {
                    long long* tl = reinterpret_cast<long long*>(p.cand + (p.cand_cap - 1024)) + 1024 + (bn - 100) * 6;   // tail of the candidate buffer
                    tl[0] = tm[0]; tl[1] = tm[1]; tl[2] = tm[2]; tl[3] = tm[3]; tl[4] = tm[4]; tl[5] = clock64();
                }
            }
            tc_commit_if(leader, &s.a_free);
            if (SS > 0) tc_commit_if(leader, &s.as_empty[item_n & 1]);
            __syncwarp();
        }
        if (diag && lane == 0) {
            long long* o = p.diag + 5 * blockIdx.x;
            o[0] = w_a; o[1] = w_b; o[2] = w_acc; o[3] = clock64() - t_begin; o[4] = bn;
        }
    } else {
        // ================= epilogue warps ===========================================================================
        const int atile = warp >> 2;
        const float ratio_gate = (float)(p.ratio * p.ratio * 1.0001);
        const int lane_base = (warp & 3) * 32;
        const int row_in_blk = atile * QTILE + lane_base + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)lane_base << 16);
        uint32_t bn = 0, item_n = 0;
        for (int item = item0; item < n_items; item += item_step, ++item_n) {
            const int qb_raw = (item % n_qgroups) * CL + (int)crank;
            const int qb = min(qb_raw, p.n_qblocks - 1);   // an odd tail: the second CTA repeats the last block and reports nothing
            const Chunk ch = p.chunks[item / n_qgroups];
            const int64_t qrow = (int64_t)qb * QBLK + row_in_blk;
            if (atile >= SS) {   // this thread's query row -> tensor memory (row = TMEM lane, 4 K-bytes per column)
                uint32_t w[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 x = *reinterpret_cast<const uint4*>(p.q_desc + sw128_chunk_offset(qrow, c));
                    w[4 * c] = x.x; w[4 * c + 1] = x.y; w[4 * c + 2] = x.z; w[4 * c + 3] = x.w;
                }
                mbar_wait(&s.a_free, (item_n & 1) ^ 1);
                tc_fence_after();
                tmem_st32(t_lane + atile * A_COLS, w);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.a_full);
            }
            const int Q = p.q_norm[qrow];
            const bool row_valid = Q >= 0 && qb_raw < p.n_qblocks;
            int H1 = NONE, L1 = NONE, H2 = NONE, L2 = NONE, pos = 0;
            auto join = [&](int gm, int nmin, int nmax, int id) {
                const int hi = 2 * gm - nmin, lo = 2 * gm - nmax;
                const bool better = hi > H1;
                H2 = max(H2, better ? H1 : hi);
                L2 = max(L2, better ? L1 : lo);
                pos = better ? id : pos;
                L1 = better ? lo : L1;
                H1 = max(H1, hi);
            };
            // the chunk's 64-row metadata records, staged once per work item (named barrier 1 among the epilogue warps)
            const int n_rec = 2 * (ch.tile_end - ch.tile_begin);
            const int4* mp = reinterpret_cast<const int4*>(p.m_meta + 2 * (int64_t)ch.tile_begin);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            const bool staged = n_rec <= p.meta_cap;
            if (staged) {
                for (int i = threadIdx.x; i < n_rec * 3; i += EPI_WARPS * 32) s.meta[i] = __ldg(mp + i);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            const uint32_t meta_s = smem_u32(s.meta);
            // record i of the chunk (16 bytes): an explicit shared-memory load when staged (a generic pointer would compile
            // to generic LD.E), the read-only path otherwise.  `staged` is uniform over the CTA.
            auto ldrec = [&](int i) {
                int4 r;
                if (staged) asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(meta_s + 16u * (uint32_t)i));
                else r = __ldg(mp + i);
                return r;
            };
            // one 64-row half: maxima of its two 32-row groups, bracket update, end-of-image test
            auto half = [&](int g0, int g1, int rec, const int4 c2 /* {end_mask, tnmin, tnmax, -} */, int unit /* 64-row unit index */) {
                const int end_mask = c2.x;
                if ((end_mask & 1) == 0) {
                    join(max(g0, g1), c2.y, c2.z, unit * 2);
                    if (end_mask & 2) {
                        const int4 c0 = ldrec(rec);   // {img0, img1, nval0, nval1}
                        sifttc3::finalize_gate(p, ratio_gate, lane, row_valid, Q, qrow, H1, L1, H2, L2, pos, c0.y, c0.w, ch);
                        H1 = NONE; L1 = NONE; H2 = NONE; L2 = NONE;
                    }
                } else {
                    const int4 c0 = ldrec(rec), c1 = ldrec(rec + 1);   // c1 = {nmin0, nmin1, nmax0, nmax1}
                    join(g0, c1.x, c1.z, unit * 2);
                    sifttc3::finalize_gate(p, ratio_gate, lane, row_valid, Q, qrow, H1, L1, H2, L2, pos, c0.x, c0.z, ch);
                    H1 = NONE; L1 = NONE; H2 = NONE; L2 = NONE;
                    if (c0.y >= 0) {
                        join(g1, c1.y, c1.w, unit * 2 + 1);
                        if (end_mask & 2) {
                            sifttc3::finalize_gate(p, ratio_gate, lane, row_valid, Q, qrow, H1, L1, H2, L2, pos, c0.y, c0.w, ch);
                            H1 = NONE; L1 = NONE; H2 = NONE; L2 = NONE;
                        }
                    }
                }
            };
#pragma unroll 1
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn) {
                long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
                if (MODE == 5) tk0 = clock64();
                // the two 64-row records of this tile, fetched before the wait so that their latency is hidden
                const int rec = (t - ch.tile_begin) * 6;
                const int4 c2a = ldrec(rec + 2), c2b = ldrec(rec + 5);
                mbar_wait(&s.acc_full[atile], bn & 1);
                tc_fence_after();
                if (MODE == 5) tk1 = clock64();
                const uint32_t tcol = t_lane + ACC_COL0 + atile * TILE_N;
                uint32_t v[2][32];
                if (MODE != 4) {
                    tmem_ld64(tcol, reinterpret_cast<uint32_t(&)[64]>(v));
                    tmem_wait_all();
                    tmem_fence_regs(v[0]);
                    tmem_fence_regs(v[1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) { v[0][i] = 0; v[1][i] = 0; }
                }
                // first half: only the two max trees run before the second load is issued; the brackets follow the release
                if (MODE == 5) tk2 = clock64();
                const int e0 = sifttc3::group_max<MODE>(v[0]), e1 = sifttc3::group_max<MODE>(v[1]);
                if (MODE != 4) {
                    tmem_ld64(tcol + HALF, reinterpret_cast<uint32_t(&)[64]>(v));
                    tmem_wait_all();
                    tmem_fence_regs(v[0]);
                    tmem_fence_regs(v[1]);
                }
                if (lane == 0) mbar_arrive(&s.acc_empty[atile]);   // both halves are out of tensor memory: the accumulator is free
                if (MODE == 5) tk3 = clock64();
                const int f0 = sifttc3::group_max<MODE>(v[0]), f1 = sifttc3::group_max<MODE>(v[1]);
                if (c2a.w != 0) {
                    // the whole 128-row tile lies inside one image and its norms are close (TileMeta::merge): ONE bracket update
                    // instead of two (the bracket is the union of the two units' -- exact whatever the row order).
                    // Warp-uniform: the records are the same for every lane.
                    join(max(max3(e0, e1, f0), f1), min(c2a.y, c2b.y), max(c2a.z, c2b.z), (4 * t) | sifttc3::POS_WIDE);
                    if (c2b.x & 2) {
                        const int4 c0 = ldrec(rec + 3);
                        sifttc3::finalize_gate(p, ratio_gate, lane, row_valid, Q, qrow, H1, L1, H2, L2, pos, c0.y, c0.w, ch);
                        H1 = NONE; L1 = NONE; H2 = NONE; L2 = NONE;
                    }
                } else {
                    half(e0, e1, rec, c2a, 2 * t);
                    half(f0, f1, rec + 3, c2b, 2 * t + 1);
                }
                if (MODE == 5 && blockIdx.x == 0 && warp == 5 && lane == 0 && bn >= 100 && bn < 164) {
                    long long* tl = reinterpret_cast<long long*>(p.cand + (p.cand_cap - 1024)) + (bn - 100) * 5;
                    tl[0] = tk0; tl[1] = tk1; tl[2] = tk2; tl[3] = tk3; tl[4] = clock64();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();   // no CTA leaves while its peer may still multicast into it
    if (warp == EPI_WARPS + 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace sifttc4
}  // namespace mcl
