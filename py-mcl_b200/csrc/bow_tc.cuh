// BOW visual-word assignment on tensor cores.
//
//   replaces: words, distance = vq(query_descriptors, voc)   (Matcher.py:250, scipy.cluster.vq.vq)
//
// words[i] = argmin_j ||o_i - c_j||^2 = argmax_j (2 o_i.c_j - |c_j|^2), o = SIFT rows (integer valued 0..255), c = k-means
// centroids (float32, NOT integer valued).  The codebook is quantised to 16-bit fixed point g = rint(256 c)
// and split into two u8 planes, g = 256 hi + lo, so that o.g = 256 (o.hi) + (o.lo) is two u8 x u8 -> s32 tensor-core
// contractions (tcgen05.mma kind::i8) into two accumulators.  The epilogue forms T = 2 (256 P_hi + P_lo) - rint(256 |c|^2)
// (the key in units of 1/256) and keeps, per query row and branch-free, the largest key of any 32-word GROUP, that group,
// and the largest group maximum among the OTHER groups (5 ALU operations per group on top of the 3-input max tree).
// |256 key - T| <= sum_k o_k + 1/2, so when best - second > 2 (sum_k o_k + 1) the exact argmin lies inside the best group
// (and is far from any float32 tie with a word outside it): bow_accept_kernel then evaluates that group's 32 words in
// float32 direct-difference form with lowest-index tie-breaking, exactly like the SIMT kernel; otherwise (every exact tie
// across groups, and the near-ties) bow_rescan_kernel rescans the row over the whole codebook the same way.  Words are
// therefore identical to the SIMT path's.
//
// Work item = (word-range split, location, block of 128 query rows); one CTA per item, persistent grid.  Warps 0-7 epilogue
// (warp % 4 = TMEM lane quadrant, warp / 4 = which half of a tile's 128 words; the halves report like two more splits),
// warp 8 TMA producer, warp 9 MMA issuer.  TMEM: 2 stages x (hi, lo) x 128 columns.
#pragma once
#include "common.cuh"

namespace mcl {
namespace bowtc {

constexpr int TILE_N = 128;                 // words per tile
constexpr int QTILE = 128;                  // query rows per item
constexpr int PLANE_BYTES = TILE_N * 128;   // 16 KB
constexpr int TILE_BYTES = 2 * PLANE_BYTES + TILE_N * 4;   // hi plane | lo plane | nb[128]
constexpr int STAGE_BYTES = 33792;          // 32 KB + 512 B, rounded to 1 KB
constexpr int B_STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int EPI_WARPS = 8;
constexpr int GROUP_W = 16;                 // words per group whose maximum the epilogue tracks
constexpr int HALVES = EPI_WARPS / 4;       // epilogue warps per TMEM lane quadrant, each owning TILE_N / HALVES words of a tile
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 512;
constexpr int NB_PAD = 1 << 30;             // nb of a padding word: its key can never win

struct Smem {
    alignas(1024) uint8_t a[QTILE * 128];
    alignas(1024) uint8_t stage[B_STAGES][2 * PLANE_BYTES];   // hi plane | lo plane
    alignas(16) int32_t nb[B_STAGES][TILE_N];                  // rint(256 |c|^2) of the stage's words
    alignas(8) uint64_t a_full, a_empty;
    uint64_t b_full[B_STAGES], b_empty[B_STAGES], nb_empty[B_STAGES];
    uint64_t acc_full[ACC_STAGES], acc_empty[ACC_STAGES];
    uint32_t tmem_base;
};

struct Params {
    const uint8_t* q_desc;      // SW128 atoms, n_qtiles*128 rows
    const uint8_t* tiles;       // per tile: TILE_BYTES (hi | lo | nb), 16-byte aligned records of STAGE_BYTES stride
    const int32_t* tile_off;    // n_loc+1: first tile of each location
    int n_loc, n_qtiles, n_q;
    int n_splits;               // each location's tiles are cut into n_splits contiguous ranges (more work items)
    int4* out;                  // [n_splits * HALVES][n_loc][n_q]: {best GROUP_W-word group, its largest T, largest T of the other groups, -}
};

// codebook -> planes (built once per index).  One warp per word row; lane l owns dims 4l..4l+3.
__global__ void build_planes_kernel(const float* __restrict__ voc, const int64_t* __restrict__ voc_off, int n_loc,
                                    const int32_t* __restrict__ tile_off, uint8_t* __restrict__ tiles, int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int wtile = blockIdx.x;           // global tile index
    // find the location of this tile
    int loc = 0;
    while (loc + 1 < n_loc && tile_off[loc + 1] <= wtile) ++loc;
    const int64_t v0 = voc_off[loc];
    const int n_words = (int)(voc_off[loc + 1] - v0);
    const int w0 = (wtile - tile_off[loc]) * TILE_N;
    uint8_t* rec = tiles + (size_t)wtile * STAGE_BYTES;
    for (int r = threadIdx.x >> 5; r < TILE_N; r += blockDim.x >> 5) {
        const int w = w0 + r;
        uint32_t hi = 0, lo = 0;
        double nn = 0.0;
        bool ok = true;
        if (w < n_words) {
            const float4 c = reinterpret_cast<const float4*>(voc + (v0 + w) * 128)[lane];
            const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float g = rintf(cc[k] * 256.0f);
                ok = ok && (g >= 0.0f) && (g <= 65535.0f);
                const uint32_t gi = (uint32_t)(int)fminf(fmaxf(g, 0.0f), 65535.0f);
                hi |= (gi >> 8) << (8 * k);
                lo |= (gi & 255u) << (8 * k);
                nn += (double)cc[k] * (double)cc[k];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        if (!ok) *bad = 1;
        *reinterpret_cast<uint32_t*>(rec + sw128_chunk_offset(r, lane >> 2) + (lane & 3) * 4) = hi;
        *reinterpret_cast<uint32_t*>(rec + PLANE_BYTES + sw128_chunk_offset(r, lane >> 2) + (lane & 3) * 4) = lo;
        if (lane == 0) {
            int nb = NB_PAD;
            if (w < n_words) {
                const double x = rint(nn * 256.0);
                if (x >= (double)(1 << 28)) { *bad = 1; } else { nb = (int)x; }
            }
            reinterpret_cast<int32_t*>(rec + 2 * PLANE_BYTES)[r] = nb;
        }
    }
}

// keys of 32 columns: T = 512 P_hi + 2 P_lo - nb
__device__ __forceinline__ void keys32(const uint32_t (&h)[32], const uint32_t (&l)[32], const int32_t* __restrict__ nb, int (&k)[32]) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const int4 c = *reinterpret_cast<const int4*>(nb + i);
        // one IMAD (FMA pipe) + one shift-add (ALU pipe) per key: both pipes run at half rate, so the work is split
        k[i] = (int)((h[i] << 9) + (uint32_t)((int)l[i] * 2 - c.x));
        k[i + 1] = (int)((h[i + 1] << 9) + (uint32_t)((int)l[i + 1] * 2 - c.y));
        k[i + 2] = (int)((h[i + 2] << 9) + (uint32_t)((int)l[i + 2] * 2 - c.z));
        k[i + 3] = (int)((h[i + 3] << 9) + (uint32_t)((int)l[i + 3] * 2 - c.w));
    }
}

// CL = CTAs per thread-block cluster.  A work item uses every codebook tile for only one 128-row query tile, so at full
// speed the CTAs would pull 32 KB per 512 MMA cycles each out of L2 (17 TB/s over the chip): the kernel is L2-bound.  With
// CL = 2 the two CTAs of a cluster take two query tiles of the same (split, location) and fetch each codebook tile ONCE:
// CTA 0 issues the hi plane, CTA 1 the lo plane, both as multicast bulk copies that land in both CTAs' shared memory; a
// stage is refilled when the MMAs of BOTH CTAs have consumed it (tcgen05.commit multicast).  The 512-byte nb[] vector is
// loaded by each CTA for itself and released by its own epilogue warps.
template <int CL>
__global__ void __launch_bounds__(THREADS, 1) bow_vq_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int warp = warp_id_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t leader = lane == 0;

    if (threadIdx.x == 0) {
        mbar_init(&s.a_full, 1);
        mbar_init(&s.a_empty, 1);
        for (int i = 0; i < B_STAGES; ++i) {
            mbar_init(&s.b_full[i], 1);
            mbar_init(&s.b_empty[i], CL);
            mbar_init(&s.nb_empty[i], EPI_WARPS);
        }
        for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&s.acc_full[i], 1); mbar_init(&s.acc_empty[i], EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == EPI_WARPS + 1) tmem_alloc<TMEM_COLS>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();   // the peer's barriers exist before a multicast copy or commit can reach them
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t ALL_CTAS = (1u << CL) - 1u;
    const int n_qgroups = (p.n_qtiles + CL - 1) / CL;
    const int n_items = p.n_splits * p.n_loc * n_qgroups;
    const int item0 = blockIdx.x / CL, item_step = gridDim.x / CL;
    // item -> (split, location, group of CL query tiles) and the split's tile range; this CTA takes tile group * CL + crank
    // (an odd tail: the last CTA repeats the last query tile and reports nothing)
    auto decode = [&](int item, int& sp, int& loc, int& qt, int& tb, int& te, int& t0) {
        qt = (item % n_qgroups) * CL + (int)crank;
        loc = (item / n_qgroups) % p.n_loc;
        sp = item / (n_qgroups * p.n_loc);
        t0 = p.tile_off[loc];
        const int nt = p.tile_off[loc + 1] - t0;
        const int per = (nt + p.n_splits - 1) / p.n_splits;
        tb = t0 + min(nt, sp * per);
        te = t0 + min(nt, (sp + 1) * per);
    };

    if (warp == EPI_WARPS) {
        // ================= TMA producer =================
        uint32_t bn = 0, item_n = 0;
        for (int item = item0; item < n_items; item += item_step, ++item_n) {
            int sp, loc, qt, tb, te, t0;
            decode(item, sp, loc, qt, tb, te, t0);
            mbar_wait(&s.a_empty, (item_n & 1) ^ 1);
            mbar_arrive_expect_tx_if(leader, &s.a_full, QTILE * 128);
            bulk_g2s_if(leader, s.a, p.q_desc + (size_t)min(qt, p.n_qtiles - 1) * QTILE * 128, QTILE * 128, &s.a_full);
            for (int t = tb; t < te; ++t, ++bn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                mbar_wait(&s.b_empty[bs], bph ^ 1);
                mbar_wait(&s.nb_empty[bs], bph ^ 1);
                mbar_arrive_expect_tx_if(leader, &s.b_full[bs], TILE_BYTES);
                const uint8_t* rec = p.tiles + (size_t)t * STAGE_BYTES;
                if (CL == 1) {
                    bulk_g2s_if(leader, s.stage[bs], rec, 2 * PLANE_BYTES, &s.b_full[bs]);
                } else {   // CL == 2: one plane per CTA, delivered to both
                    bulk_g2s_multicast_if(leader, s.stage[bs] + crank * PLANE_BYTES, rec + crank * PLANE_BYTES, PLANE_BYTES, &s.b_full[bs],
                                          ALL_CTAS);
                }
                bulk_g2s_if(leader, s.nb[bs], rec + 2 * PLANE_BYTES, TILE_N * 4, &s.b_full[bs]);
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = umma_idesc(/*c=S32*/ 2, /*a=u8*/ 0, /*b=u8*/ 0, QTILE, TILE_N);
        const uint64_t a_desc = umma_desc_sw128(smem_u32(s.a));
        uint32_t bn = 0, item_n = 0;
        for (int item = item0; item < n_items; item += item_step, ++item_n) {
            int sp, loc, qt, tb, te, t0;
            decode(item, sp, loc, qt, tb, te, t0);
            mbar_wait(&s.a_full, item_n & 1);
            for (int t = tb; t < te; ++t, ++bn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                const uint32_t as = bn % ACC_STAGES, aph = (bn / ACC_STAGES) & 1;
                mbar_wait(&s.b_full[bs], bph);
                mbar_wait(&s.acc_empty[as], aph ^ 1);
                tc_fence_after();
                const uint64_t hi_desc = umma_desc_sw128(smem_u32(s.stage[bs]));
                const uint64_t lo_desc = umma_desc_sw128(smem_u32(s.stage[bs] + PLANE_BYTES));
                const uint32_t d = tmem_base + as * (2 * TILE_N);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8_if(leader, d, a_desc + 2 * k, hi_desc + 2 * k, idesc, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8_if(leader, d + TILE_N, a_desc + 2 * k, lo_desc + 2 * k, idesc, k > 0);
                if (CL == 1) tc_commit_if(leader, &s.b_empty[bs]);
                else tc_commit_multicast_if(leader, &s.b_empty[bs], ALL_CTAS);
                tc_commit_if(leader, &s.acc_full[as]);
                __syncwarp();
            }
            tc_commit_if(leader, &s.a_empty);
            __syncwarp();
        }
    } else {
        // ================= epilogue warps =================
        const int half = warp >> 2;
        const int lane_base = (warp & 3) * 32;
        const uint32_t t_lane = tmem_base + ((uint32_t)lane_base << 16);
        uint32_t bn = 0;
        for (int item = item0; item < n_items; item += item_step) {
            int sp, loc, qt, tb, te, t0;
            decode(item, sp, loc, qt, tb, te, t0);
            const int row = qt * QTILE + lane_base + lane;
            int b1 = INT_MIN, b2 = INT_MIN, grp = 0;
            for (int t = tb; t < te; ++t, ++bn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                const uint32_t as = bn % ACC_STAGES, aph = (bn / ACC_STAGES) & 1;
                mbar_wait(&s.acc_full[as], aph);
                tc_fence_after();
                mbar_wait(&s.b_full[bs], bph);   // complete already; orders the read of nb[]
                const int32_t* nb = s.nb[bs];
                const uint32_t tcol = t_lane + as * (2 * TILE_N);
#pragma unroll
                for (int gg = 0; gg < TILE_N / 32 / HALVES; ++gg) {
                    const int g = half * (TILE_N / 32 / HALVES) + gg;
                    uint32_t h[32], l[32];
                    tmem_ld32(tcol + g * 32, h);
                    tmem_ld32(tcol + TILE_N + g * 32, l);
                    tmem_wait_all();
                    tmem_fence_regs(h);
                    tmem_fence_regs(l);
                    int k[32];
                    keys32(h, l, nb + g * 32, k);
                    // top-2 over the maxima of GROUP_W-word groups, branch-free (the word inside the best group is found later,
                    // exactly, by bow_accept_kernel: the narrower the group, the less of the float32 codebook it has to read)
#pragma unroll
                    for (int sub = 0; sub < 32 / GROUP_W; ++sub) {
                        const int* kk = k + sub * GROUP_W;
                        const int gm = max(max3(max3(kk[0], kk[1], kk[2]), max3(kk[3], kk[4], kk[5]), max3(kk[6], kk[7], kk[8])),
                                           max3(max3(kk[9], kk[10], kk[11]), max3(kk[12], kk[13], kk[14]), kk[15]));
                        const bool better = gm > b1;
                        b2 = max(b2, min(b1, gm));
                        grp = better ? ((t - t0) * (TILE_N / 32) + g) * (32 / GROUP_W) + sub : grp;
                        b1 = max(b1, gm);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&s.acc_empty[as]);
                    mbar_arrive(&s.nb_empty[bs]);
                }
            }
            if (row < p.n_q && qt < p.n_qtiles) p.out[((size_t)(sp * HALVES + half) * p.n_loc + loc) * p.n_q + row] = make_int4(grp, b1, b2, 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync();   // no CTA leaves while its peer may still multicast into it
    if (warp == EPI_WARPS + 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---- resolve ---------------------------------------------------------------------------------------------------
// Pass 1 (one warp per (location, row)): merge the splits' (best group, its maximum, maximum of the other groups); when the
// lead exceeds the quantisation bound 2 (sum_k o_k + 1) the exact argmin is one of the best group's 32 words, which are
// evaluated here in float32 direct-difference form (lowest index on ties); else the row is queued for a full rescan.
// state: [0] = queued rows, [1] = run-the-SIMT-kernel flag (written by bow_decide_kernel), [2] = rows rescanned (statistics)
__global__ void __launch_bounds__(256) bow_accept_kernel(const int4* __restrict__ tc, int n_splits, const float* __restrict__ q_desc,
                                                         int n_q, int n_loc, const int* __restrict__ bad,
                                                         const float* __restrict__ voc, const int64_t* __restrict__ voc_off,
                                                         unsigned long long* __restrict__ best, int* __restrict__ queue,
                                                         unsigned int* __restrict__ state) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_w = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_jobs = (int64_t)n_loc * n_q;
    const bool all_bad = *bad != 0;   // non-integer query rows: nothing from the tensor-core pass can be trusted
    for (int64_t job = wid; job < n_jobs; job += n_w) {
        const int row = (int)(job % n_q), loc = (int)(job / n_q);
        const float4 o = reinterpret_cast<const float4*>(q_desc + (size_t)row * 128)[lane];
        float sum = o.x + o.y + o.z + o.w;   // integer valued: exact
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, k);
        int b1 = INT_MIN, b2 = INT_MIN, grp = 0;
        for (int sp = 0; sp < n_splits; ++sp) {
            const int4 r = tc[(size_t)sp * n_jobs + job];
            if (r.y > b1) { b2 = max(b1, r.z); b1 = r.y; grp = r.x; }
            else { b2 = max(b2, r.y); }
        }
        const long long lead = (long long)b1 - (long long)b2;
        const bool accept = !all_bad && (b2 == INT_MIN || lead > 2LL * ((long long)sum + 1));   // a single group cannot tie
        if (!accept) {
            if (lane == 0) queue[atomicAdd(&state[0], 1u)] = (int)job;
            continue;
        }
        const int64_t v0 = voc_off[loc];
        const int n_words = (int)(voc_off[loc + 1] - v0);
        // The group's GROUP_W words against this row.  Lane l holds dims 4l..4l+3 of the row; it forms its partial sum for all
        // GROUP_W words (coalesced 512-byte row loads), then the partials are summed ACROSS lanes by a transposing reduction:
        // at distance 16 a lane keeps the half of the word list that matches its bit 4 and adds the partner's partials of
        // those words, and so on -- 15 + 1 shuffles instead of 16 x 5, after which lanes 2i and 2i+1 own word i.
        // Every word's sum is formed by the same pairing (l, l^16), (l, l^8), ... as bow_rescan_kernel's butterfly, so the
        // float32 distances are bit-identical to the rescan's and the SIMT kernel's tie-breaking is preserved.
        const int w0 = grp * GROUP_W;
        float pw[GROUP_W];
#pragma unroll
        for (int j = 0; j < GROUP_W; ++j) {
            const int w = min(w0 + j, n_words - 1);
            const float4 cc = reinterpret_cast<const float4*>(voc + (v0 + w) * 128)[lane];
            const float dx = o.x - cc.x, dy = o.y - cc.y, dz = o.z - cc.z, dw = o.w - cc.w;
            pw[j] = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
        }
        static_assert(GROUP_W == 16, "the reduction below is written for 16 words");
        float p8[8], p4[4], p2[2];
        {
            const bool up = (lane & 16) != 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float keep = up ? pw[i + 8] : pw[i], send = up ? pw[i] : pw[i + 8];
                p8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
        {
            const bool up = (lane & 8) != 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float keep = up ? p8[i + 4] : p8[i], send = up ? p8[i] : p8[i + 4];
                p4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
        {
            const bool up = (lane & 4) != 0;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float keep = up ? p4[i + 2] : p4[i], send = up ? p4[i] : p4[i + 2];
                p2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        }
        float dist;
        {
            const bool up = (lane & 2) != 0;
            const float keep = up ? p2[1] : p2[0], send = up ? p2[0] : p2[1];
            dist = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        dist += __shfl_xor_sync(0xffffffffu, dist, 1);   // lanes 2i and 2i+1 both own word i now
        const int wi = w0 + (lane >> 1);
        unsigned long long bk = (wi < n_words) ? (((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)wi) : ~0ULL;
#pragma unroll
        for (int k = 16; k > 1; k >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, bk, k);
            bk = other < bk ? other : bk;
        }
        if (lane == 0) best[job] = bk;
    }
}

// many ambiguous rows (random-like data) -> one pass of the float32 SIMT kernel is cheaper than row-by-row rescans
__global__ void bow_decide_kernel(const int* __restrict__ bad, int64_t n_jobs, unsigned int* __restrict__ state) {
    const unsigned int thr = (unsigned int)max((int64_t)64, n_jobs / 16);
    state[1] = (*bad != 0 || state[0] > thr) ? 1u : 0u;
}

// Pass 2 (one CTA per queued row, 8 warps striding over the words, 4 words in flight per warp): exact float32
// direct-difference scan of the location's codebook, lowest index on ties.
// best[] gets (distance bits << 32 | word) like bow_vq_kernel's atomicMin key.
__global__ void __launch_bounds__(256) bow_rescan_kernel(const int* __restrict__ queue, const float* __restrict__ q_desc, int n_q,
                                                         const float* __restrict__ voc, const int64_t* __restrict__ voc_off,
                                                         unsigned long long* __restrict__ best, unsigned int* __restrict__ state) {
    if (state[1] != 0) return;
    __shared__ unsigned long long part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned n = state[0];
    for (unsigned c = blockIdx.x; c < n; c += gridDim.x) {
        const int job = queue[c];
        const int loc = job / n_q, row = job % n_q;
        const float4 o = reinterpret_cast<const float4*>(q_desc + (size_t)row * 128)[lane];
        const int64_t v0 = voc_off[loc];
        const int n_words = (int)(voc_off[loc + 1] - v0);
        unsigned long long bk = ~0ULL;
        for (int w0 = warp * 4; w0 < n_words; w0 += 32) {
            float d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int w = min(w0 + j, n_words - 1);
                const float4 cc = reinterpret_cast<const float4*>(voc + (v0 + w) * 128)[lane];
                const float dx = o.x - cc.x, dy = o.y - cc.y, dz = o.z - cc.z, dw = o.w - cc.w;
                d[j] = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
            }
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) d[j] += __shfl_xor_sync(0xffffffffu, d[j], k);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (w0 + j < n_words) {
                    const unsigned long long key = ((unsigned long long)__float_as_uint(d[j]) << 32) | (unsigned)(w0 + j);
                    bk = key < bk ? key : bk;
                }
            }
        }
        if (lane == 0) part[warp] = bk;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long m = part[0];
            for (int k = 1; k < 8; ++k) m = part[k] < m ? part[k] : m;
            best[job] = m;
            atomicAdd(&state[2], 1u);
        }
        __syncthreads();
    }
}

}  // namespace bowtc
}  // namespace mcl
