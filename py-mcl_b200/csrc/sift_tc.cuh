// SIFT / integer-L2 matching on 5th-gen tensor cores (sm_100a).
//
//   replaces: Matcher.SIFTMatch (Matcher.py:262-291) looped by Matcher.run (Matcher.py:305-338) looped by
//             analyzer.createRawP (analyze.py:78-90): knnMatch(k=2) + ratio lambda (Matcher.py:280) + len().
//
// For every query descriptor q_i and every map image I the reference needs the two smallest
//   d2(i,j) = |q_i|^2 + |m_j|^2 - 2 q_i.m_j ,  j in I
// q_i.m_j is a dense u8 x u8 -> s32 contraction with K = 128: tcgen05.mma kind::i8, operands staged in shared
// memory by TMA bulk copies (map rows are resident in HBM already in the SWIZZLE_128B image, see common.cuh),
// accumulators in TMEM (2 query tiles x 128 columns, double buffered = all 512 columns).
//
// The epilogue is the bottleneck at K=128 (one s32 output per 128 MACs), so it does the minimum per element:
//   key_j = 2*acc_j - |m_j|^2      (larger key = smaller distance; one IMAD/LEA per element)
// then a 3-input-max tree per 16-column group and a top-2 over GROUP maxima, remembering which group holds the
// best key.  That yields the exact best distance and
// an upper bound on the second-best distance (best of all other groups).  A row that fails the ratio test
// against that bound fails it against the true second best as well, so it is rejected for good; the few rows
// that pass are appended to a candidate list and resolved exactly (second best inside the winning group) by
// sift_fixup_kernel, which also does the per-(frame,image) count.  Counts are therefore bit-exact.
#pragma once
#include "common.cuh"

namespace mcl {
namespace sifttc {

constexpr int TILE_N = 128;        // map rows per tile (= MMA N)
constexpr int GROUP = 32;          // columns per group; images are padded to a multiple of GROUP rows
constexpr int GROUPS = TILE_N / GROUP;
constexpr int QTILE = 128;         // query rows per MMA (= MMA M)
constexpr int QBLK = 2 * QTILE;    // query rows per work item (two A tiles share every B tile)
constexpr int ROW_BYTES = 128;
constexpr int TILE_BYTES = TILE_N * ROW_BYTES;  // 16 KB
constexpr int QBLK_BYTES = QBLK * ROW_BYTES;    // 32 KB
constexpr int B_STAGES = 6;
constexpr int META_STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int META_INTS = 140;     // bias[128] img[4] nval[4] end_mask pad[3]
constexpr int META_BYTES = META_INTS * 4;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 512;
constexpr int NEG_KEY = -(1 << 30);  // 'no column yet'; Q - NEG_KEY still fits in int32
constexpr int PAD_BIAS = -(1 << 29);  // key of a padding column: never the best of a real image

struct TileMeta {
    int32_t bias[TILE_N];
    int32_t img[GROUPS];    // image id of each group (-1: padding group)
    int32_t nval[GROUPS];   // number of real descriptors of that image
    int32_t end_mask;       // bit g: group g is the last group of its image
    int32_t pad[3];
};
static_assert(sizeof(TileMeta) == META_BYTES, "meta size");

struct Chunk {  // a contiguous range of whole images handled by one work item
    int32_t tile_begin, tile_end;  // tiles [begin,end)
    int32_t img_begin, img_end;    // images [begin,end) whose results this chunk owns
};

struct Smem {
    alignas(1024) uint8_t a[QBLK_BYTES];
    alignas(1024) uint8_t b[B_STAGES][TILE_BYTES];
    alignas(16) TileMeta meta[META_STAGES];
    alignas(8) uint64_t a_full, a_empty;
    uint64_t b_full[B_STAGES], b_empty[B_STAGES];
    uint64_t meta_full[META_STAGES], meta_empty[META_STAGES];
    uint64_t acc_full[ACC_STAGES], acc_empty[ACC_STAGES];
    uint32_t tmem_base;
};

struct Params {
    const uint8_t* q_desc;    // SW128 atoms, n_qblocks*QBLK rows
    const int32_t* q_norm;    // -1 for padding rows
    const uint8_t* m_desc;    // SW128 atoms, n_tiles*TILE_N rows
    const TileMeta* m_meta;   // n_tiles
    const Chunk* chunks;
    int n_qblocks, n_chunks;
    double ratio;
    int4* cand;               // {q row, first map row of winning group, key of best other group, image}
    unsigned int* cand_count;
    unsigned int cand_cap;
    int* overflow;
    int q_row_base;           // global row index of q_desc[0] (batches of query blocks)
};

// ------------------------------------------------------------------------------------------
// per-tile metadata (built once per map)
// ------------------------------------------------------------------------------------------
__global__ void build_meta_kernel(const int32_t* __restrict__ m_norm, const int32_t* __restrict__ img_of,
                                  const int64_t* __restrict__ img_src_off, int n_tiles, TileMeta* __restrict__ meta) {
    const int t = blockIdx.x;
    if (t >= n_tiles) return;
    const int c = threadIdx.x;  // 0..127
    const int64_t row = (int64_t)t * TILE_N + c;
    const int nm = m_norm[row];
    meta[t].bias[c] = nm >= 0 ? -nm : PAD_BIAS;
    if (c < GROUPS) {
        const int64_t r0 = (int64_t)t * TILE_N + c * GROUP;
        const int im = img_of[r0];
        meta[t].img[c] = im;
        meta[t].nval[c] = im >= 0 ? (int)(img_src_off[im + 1] - img_src_off[im]) : 0;
    }
    if (c == 0) {
        int mask = 0;
        for (int gg = 0; gg < GROUPS; ++gg) {
            const int64_t r0 = (int64_t)t * TILE_N + gg * GROUP;
            const int im = img_of[r0];
            const int64_t rn = r0 + GROUP;
            const int nx = (rn < (int64_t)n_tiles * TILE_N) ? img_of[rn] : -1;
            if (im >= 0 && nx != im) mask |= 1 << gg;
        }
        meta[t].end_mask = mask;
        for (int k = 0; k < 3; ++k) meta[t].pad[k] = 0;
    }
}

// ------------------------------------------------------------------------------------------
// main kernel: persistent, warp specialised.  warps 0-7 epilogue, warp 8 TMA producer, warp 9 MMA issuer.
// ------------------------------------------------------------------------------------------
// End of an image: ratio test of (exact best, upper bound of second best); survivors become candidates.
__device__ __noinline__ void finalize_image(const Params& p, int lane, bool row_valid, int Q, int64_t qrow,
                                            int M1, int M2, int pos, int img, int nval, const Chunk& ch) {
    bool pass = false;
    if (row_valid && nval >= 2 && img >= ch.img_begin && img < ch.img_end) pass = ratio_pass(Q - M1, Q - M2, p.ratio);
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

// max over one 32-column group of key = 2*acc - |m|^2.
// The add runs as IMAD on the FMA pipe for NFMA of every 8 columns (multiplier `two` is opaque to the compiler,
// which would otherwise fold it into IADD3 on the ALU pipe, where the 3-input max already runs) and as IADD3 on the
// ALU pipe for the rest: the two pipes issue side by side.
constexpr int NFMA = 6;
template <int MODE>
__device__ __forceinline__ int group_max(const uint32_t (&v)[32], const int32_t* bias /* shared */, int two) {
    if (MODE == 3 || MODE == 4) return (int)(v[0] | v[31]);  // diagnostic: no arithmetic (3: not even the TMEM read)
    if (MODE == 1) {  // diagnostic: TMEM read + minimal arithmetic (results are meaningless)
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 32; i += 2) x = x | v[i] | v[i + 1];
        return (int)x;
    }
    if (MODE == 2) {  // diagnostic: 3-input max only (no bias add)
        int l[11];
#pragma unroll
        for (int i = 0; i < 10; ++i) l[i] = max3((int)v[3 * i], (int)v[3 * i + 1], (int)v[3 * i + 2]);
        l[10] = max((int)v[30], (int)v[31]);
        const int a0 = max3(l[0], l[1], l[2]), a1 = max3(l[3], l[4], l[5]), a2 = max3(l[6], l[7], l[8]);
        return max3(max3(a0, a1, a2), l[9], l[10]);
    }
    int k[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const int4 c = *reinterpret_cast<const int4*>(bias + i);
        const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int a = (int)v[i + j];
            k[i + j] = ((i + j) & 7) < NFMA ? a * two + cc[j] : a + a + cc[j];
        }
    }
    // 3-input max tree (depth 4 instead of a 16-long chain)
    int l1[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) l1[i] = max3(k[3 * i], k[3 * i + 1], k[3 * i + 2]);
    l1[10] = max(k[30], k[31]);
    const int a0 = max3(l1[0], l1[1], l1[2]), a1 = max3(l1[3], l1[4], l1[5]), a2 = max3(l1[6], l1[7], l1[8]);
    return max3(max3(a0, a1, a2), l1[9], l1[10]);
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) sift_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];   // SWIZZLE_128B operands need 1024-byte alignment
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int warp = warp_id_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t leader = lane == 0;

    if (threadIdx.x == 0) {
        mbar_init(&s.a_full, 1);
        mbar_init(&s.a_empty, 1);
        for (int i = 0; i < B_STAGES; ++i) { mbar_init(&s.b_full[i], 1); mbar_init(&s.b_empty[i], 1); }
        for (int i = 0; i < META_STAGES; ++i) { mbar_init(&s.meta_full[i], 1); mbar_init(&s.meta_empty[i], EPI_WARPS); }
        for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&s.acc_full[i], 1); mbar_init(&s.acc_empty[i], EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == EPI_WARPS + 1) tmem_alloc<TMEM_COLS>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;

    const int n_items = p.n_qblocks * p.n_chunks;

    if (warp == EPI_WARPS) {
        // ================= TMA producer =================
        uint32_t bn = 0, mn = 0, item_n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
            const int qb = item % p.n_qblocks;
            const Chunk ch = p.chunks[item / p.n_qblocks];
            mbar_wait(&s.a_empty, (item_n & 1) ^ 1);
            mbar_arrive_expect_tx_if(leader, &s.a_full, QBLK_BYTES);
            bulk_g2s_if(leader, s.a, p.q_desc + (size_t)qb * QBLK_BYTES, QBLK_BYTES, &s.a_full);
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn, ++mn) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                const uint32_t ms = mn % META_STAGES, mph = (mn / META_STAGES) & 1;
                mbar_wait(&s.meta_empty[ms], mph ^ 1);
                mbar_arrive_expect_tx_if(leader, &s.meta_full[ms], META_BYTES);
                bulk_g2s_if(leader, &s.meta[ms], p.m_meta + t, META_BYTES, &s.meta_full[ms]);
                mbar_wait(&s.b_empty[bs], bph ^ 1);
                mbar_arrive_expect_tx_if(leader, &s.b_full[bs], TILE_BYTES);
                bulk_g2s_if(leader, s.b[bs], p.m_desc + (size_t)t * TILE_BYTES, TILE_BYTES, &s.b_full[bs]);
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = umma_idesc(/*c=S32*/ 2, /*a=u8*/ 0, /*b=u8*/ 0, QTILE, TILE_N);
        const uint64_t a_desc0 = umma_desc_sw128(smem_u32(s.a));
        const uint64_t a_desc1 = umma_desc_sw128(smem_u32(s.a + QTILE * ROW_BYTES));
        uint32_t bn = 0, an = 0, item_n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
            const Chunk ch = p.chunks[item / p.n_qblocks];
            mbar_wait(&s.a_full, item_n & 1);
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++bn, ++an) {
                const uint32_t bs = bn % B_STAGES, bph = (bn / B_STAGES) & 1;
                const uint32_t as = an % ACC_STAGES, aph = (an / ACC_STAGES) & 1;
                mbar_wait(&s.b_full[bs], bph);
                mbar_wait(&s.acc_empty[as], aph ^ 1);
                tc_fence_after();
                {
                    const uint64_t b_desc = umma_desc_sw128(smem_u32(s.b[bs]));
                    const uint32_t d0 = tmem_base + as * (2 * TILE_N);
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // K = 4 x 32 bytes; +2 in the (addr>>4) field per step
                        umma_i8_if(leader, d0, a_desc0 + 2 * k, b_desc + 2 * k, idesc, k > 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_i8_if(leader, d0 + TILE_N, a_desc1 + 2 * k, b_desc + 2 * k, idesc, k > 0);
                    tc_commit_if(leader, &s.b_empty[bs]);
                    tc_commit_if(leader, &s.acc_full[as]);
                }
                __syncwarp();
            }
            tc_commit_if(leader, &s.a_empty);
            __syncwarp();
        }
    } else {
        // ================= epilogue warps =================
        const int atile = warp >> 2;
        const int lane_base = (warp & 3) * 32;
        const int row_in_blk = atile * QTILE + lane_base + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)lane_base << 16) + atile * TILE_N;
        uint32_t mn = 0, an = 0;
        int two = 2;
        asm volatile("" : "+r"(two));  // opaque: keeps the key computation an IMAD (FMA pipe), see group_max
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int qb = item % p.n_qblocks;
            const Chunk ch = p.chunks[item / p.n_qblocks];
            const int64_t qrow = (int64_t)qb * QBLK + row_in_blk;
            const int Q = p.q_norm[qrow];
            const bool row_valid = Q >= 0;
            int M1 = NEG_KEY, M2 = NEG_KEY, pos = 0;
            for (int t = ch.tile_begin; t < ch.tile_end; ++t, ++mn, ++an) {
                const uint32_t ms = mn % META_STAGES, mph = (mn / META_STAGES) & 1;
                const uint32_t as = an % ACC_STAGES, aph = (an / ACC_STAGES) & 1;
                mbar_wait(&s.meta_full[ms], mph);
                const TileMeta& mt = s.meta[ms];
                const int end_mask = mt.end_mask;
                mbar_wait(&s.acc_full[as], aph);
                tc_fence_after();
                const uint32_t tcol = t_lane + as * (2 * TILE_N);
                // One TMEM round trip per tile: the four 32-column loads are issued back to back and waited for once
                // (an LDTM -> wait round trip costs ~230 cycles however much it moves), then the accumulator stage goes
                // back to the MMA warp before any arithmetic starts.
                uint32_t v[GROUPS][32];
                if (MODE != 3) {
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
                if (lane == 0) mbar_arrive(&s.acc_empty[as]);
#pragma unroll
                for (int g = 0; g < GROUPS; ++g) {
                    const int gm = group_max<MODE>(v[g], &mt.bias[g * GROUP], two);
                    pos = gm > M1 ? t * GROUPS + g : pos;
                    M2 = max(M2, min(M1, gm));
                    M1 = max(M1, gm);
                    if (end_mask & (1 << g)) {
                        finalize_image(p, lane, row_valid, Q, qrow, M1, M2, pos, mt.img[g], mt.nval[g], ch);
                        M1 = NEG_KEY; M2 = NEG_KEY;
                    }
                }
                // release the metadata slot
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.meta_empty[ms]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// exact resolution of candidates: second best inside the winning 32-row group, ratio test, count.
// One warp per candidate; lane l owns map row l of the group.
// ------------------------------------------------------------------------------------------
__global__ void sift_fixup_kernel(const int4* __restrict__ cand, const unsigned int* __restrict__ cand_count,
                                  unsigned int cand_cap, const uint8_t* __restrict__ q_desc,
                                  const int32_t* __restrict__ q_norm, const int32_t* __restrict__ q_frame,
                                  const uint8_t* __restrict__ m_desc, const int32_t* __restrict__ m_norm,
                                  double ratio, int n_images, int32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const unsigned n = min(*cand_count, cand_cap);
    const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n; c += n_warps) {
        const int4 cd = cand[c];
        const int64_t qrow = cd.x;
        const int64_t mrow = (int64_t)cd.y + lane;
        unsigned dot = 0;
#pragma unroll
        for (int chunk = 0; chunk < 8; ++chunk) {
            const uint4 a = *reinterpret_cast<const uint4*>(q_desc + sw128_chunk_offset(qrow, chunk));
            const uint4 b = *reinterpret_cast<const uint4*>(m_desc + sw128_chunk_offset(mrow, chunk));
            dot = __dp4a(a.x, b.x, dot); dot = __dp4a(a.y, b.y, dot);
            dot = __dp4a(a.z, b.z, dot); dot = __dp4a(a.w, b.w, dot);
        }
        const int nm = m_norm[mrow];
        const int v = (nm >= 0) ? (int)(2u * dot) - nm : INT_MIN;
        int m1 = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        const unsigned eq = __ballot_sync(0xffffffffu, v == m1);
        int sec = (v == m1) ? INT_MIN : v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sec = max(sec, __shfl_xor_sync(0xffffffffu, sec, o));
        if (__popc(eq) >= 2) sec = m1;
        const int v2 = max(sec, cd.z);
        if (lane == 0) {
            const int Q = q_norm[qrow];
            if (ratio_pass(Q - m1, Q - v2, ratio)) atomicAdd(&counts[(int64_t)q_frame[qrow] * n_images + cd.w], 1);
        }
    }
}

}  // namespace sifttc

// counts[f][i] = mask[f][i] ? counts[f][i] : fill      (DOR: unmatched images get count 1, Matcher.py:359)
__global__ void apply_mask_kernel(int32_t* __restrict__ counts, const uint8_t* __restrict__ mask, int64_t n, int32_t fill) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!mask[i]) counts[i] = fill;
}

}  // namespace mcl
