// BOW visual-word assignment: exact float32 resolution of the tensor-core pass (tc9.cuh, KIND_BOW).
//
//   replaces: words, distance = vq(query_descriptors, voc)   (Matcher.py:250, scipy.cluster.vq.vq)
//
// tc9_kernel<KIND_BOW> gives, per (code-book range, location, query row), the 32-word GROUP with the largest approximate key
// o.c - |c|^2/2 (+C), that key, and the largest key of any other group.  The keys carry three errors, all bounded:
//   (i)   centroids rounded to fp16: |c^ - c| <= 2^-11 c per component (normal halves; <= 2^-25 absolute below 2^-14), and
//         o, c >= 0, so |o.c^ - o.c| <= 2^-11 o.c + 1;
//   (ii)  the norm term 256 (v1 + v2) vs C - |c|^2/2: two halves carry it to 2^-22 relative, <= 1 for every admissible c;
//   (iii) float32 accumulation of the <= 146 products inside the tensor core: all terms but one tiny one are non-negative, so
//         every partial sum is <= the key and the total rounding is <= 146 x 2^-22 key (truncating adds); 2^-13 key is assumed.
// With key >= o.c this gives |key^ - key| <= E(key^) = (2^-11 + 2^-13) key^ + 4 for every word, and E grows with the key, so
// a lead of the best group over all others larger than 2 E(best) proves that the exact argmin lies inside the best group --
// and that it is not within float32 rounding of a word outside it.  Such rows are ACCEPTED: their group's 32 words are
// evaluated in float32 direct-difference form with lowest-index tie-breaking, exactly like the SIMT kernel.  The accepted
// (location, row) jobs are first SORTED BY GROUP (counting sort): query rows that fall into the same group -- re-observed
// places do -- are then handled by one warp that keeps the group's 32 float32 code-book rows in registers, instead of every
// job pulling 16 KB through L2 (what bounded the previous resolve: 0.54 ms next to a 0.49 ms tensor pass).  The other rows
// (every exact tie across groups, and the near-ties) are rescanned over the whole code book the same way
// (bow_rescan_kernel), or -- when more than 1/16 of the rows are ambiguous, or the query rows are not integer valued -- by
// one pass of the float32 SIMT kernel, decided on the device.  Words therefore equal the SIMT path's.
#pragma once
#include "common.cuh"
#include "tc9.cuh"

namespace mcl {
namespace bowtc {

constexpr int GROUP_W = tc9::GROUP;   // 32 words per group
constexpr float KEY_REL = 0.00048828125f + 0.0001220703125f;   // 2^-11 + 2^-13
constexpr float KEY_ABS = 4.0f;

// Pass 1 (one thread per (location, row) job): merge the ranges' records into the three best groups, their keys
// k1 >= k2 >= k3, the fourth-best key k4 and the keys s[0..3] of the best group's four 8-word SUB-BLOCKS.  With the margin
// 2 E(k1): the first k_i that trails k1 by more than the margin proves that the argmin lies in the groups before it, and
// inside the best group only sub-blocks whose key is within the margin of k1 can hold it.  Every such sub-block becomes one
// ENTRY (query row, bucket = location x sub-blocks + sub-block) -- most rows end with a single entry of 8 words -- the groups without sub-block keys (second,
// third) contribute all four of theirs; the resolve combines a row's entries with a 64-bit atomicMin on (distance, word).
// If even k4 is within the margin the row is queued for a full rescan.
// state: [0] = queued rows, [1] = run-the-SIMT-kernel flag (bow_decide_kernel), [2] = rows rescanned (statistics),
//        [3] = entries in the sorted list (bow_scan_kernel)
constexpr int TOPG = 3;
constexpr int SUB_W = 8;                      // words per sub-block
constexpr int SUBS = GROUP_W / SUB_W;         // 4 sub-blocks per group
__global__ void __launch_bounds__(256) bow_merge_kernel(const int4* __restrict__ tc, int n_splits, int n_q, int n_loc, int subs_per_loc,
                                                        const int* __restrict__ bad, int* __restrict__ job_sel /* [n_jobs][TOPG]: group << 4 | sub-block mask */,
                                                        int* __restrict__ bucket_cnt, int* __restrict__ queue,
                                                        unsigned int* __restrict__ state) {
    const int64_t n_jobs = (int64_t)n_loc * n_q;
    const bool all_bad = *bad != 0;   // non-integer query rows: nothing from the tensor-core pass can be trusted
    for (int64_t job = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; job < n_jobs; job += (int64_t)gridDim.x * blockDim.x) {
        int k1 = 0, k2 = 0, k3 = 0, k4 = 0, p1 = 0, p2 = 0, p3 = 0;   // positive float32 keys compare like their bit patterns; 0 = none
        int4 sub = make_int4(0, 0, 0, 0);
        for (int sp = 0; sp < n_splits; ++sp) {
            const int4* r = tc + 3 * ((size_t)sp * n_jobs + job);
            const int4 pp = r[0], kk = r[1];
            if (kk.x <= 0) continue;   // this range reported nothing for the row
            auto ins = [&](int k, int pos, bool is_best_of_range) {
                if (k > k1) { k4 = k3; k3 = k2; p3 = p2; k2 = k1; p2 = p1; k1 = k; p1 = pos; sub = is_best_of_range ? r[2] : make_int4(0, 0, 0, 0); }
                else if (k > k2) { k4 = k3; k3 = k2; p3 = p2; k2 = k; p2 = pos; }
                else if (k > k3) { k4 = k3; k3 = k; p3 = pos; }
                else k4 = max(k4, k);
            };
            ins(kk.x, pp.x, true);
            if (kk.y > 0) ins(kk.y, pp.y, false);
            if (kk.z > 0) ins(kk.z, pp.z, false);
            if (kk.w > 0) k4 = max(k4, kk.w);
        }
        int sel[TOPG] = {-1, -1, -1};
        if (!all_bad && k1 > 0) {
            const float f1 = __int_as_float(k1);
            const float margin = 2.0f * (KEY_REL * f1 + KEY_ABS);
            int n_groups = 0;
            if (k2 <= 0 || f1 - __int_as_float(k2) > margin) n_groups = 1;
            else if (k3 <= 0 || f1 - __int_as_float(k3) > margin) n_groups = 2;
            else if (k4 <= 0 || f1 - __int_as_float(k4) > margin) n_groups = 3;
            if (n_groups >= 1) {
                // sub-blocks of the best group that can hold the argmin (no sub-block keys: all four)
                int mask = 0;
                const int sk[4] = {sub.x, sub.y, sub.z, sub.w};
#pragma unroll
                for (int j = 0; j < SUBS; ++j)
                    if (sk[j] <= 0 || f1 - __int_as_float(sk[j]) <= margin) mask |= 1 << j;
                if (sub.x <= 0 && sub.y <= 0 && sub.z <= 0 && sub.w <= 0) mask = 15;
                sel[0] = (p1 << 4) | mask;
            }
            if (n_groups >= 2) sel[1] = (p2 << 4) | 15;
            if (n_groups >= 3) sel[2] = (p3 << 4) | 15;
        }
        const int base = (int)(job / n_q) * subs_per_loc;
#pragma unroll
        for (int i = 0; i < TOPG; ++i) {
            job_sel[TOPG * job + i] = sel[i];
            if (sel[i] >= 0) {
#pragma unroll
                for (int j = 0; j < SUBS; ++j)
                    if (sel[i] & (1 << j)) atomicAdd(&bucket_cnt[base + (sel[i] >> 4) * SUBS + j], 1);
            }
        }
        if (sel[0] < 0) queue[atomicAdd(&state[0], 1u)] = (int)job;
    }
}

// exclusive prefix sum of the bucket counts (one CTA); cnt[] becomes the running fill position of every bucket.
// Coalesced: a pass covers SCAN_TILES tiles of 1024 consecutive counters, thread t owning counter t of every tile (all loads of
// a pass are independent and issued together); a tile is scanned with warp shuffles, the 32 x SCAN_TILES warp totals by one
// warp.  (The first version gave each thread a contiguous run of n / 1024 counters: strided loads, 40 us for 40 000 buckets.)
constexpr int SCAN_TILES = 8;
__global__ void __launch_bounds__(1024) bow_scan_kernel(int* __restrict__ cnt, int n, unsigned int* __restrict__ state) {
    __shared__ int wtot[SCAN_TILES * 32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += SCAN_TILES * 1024) {
        int v[SCAN_TILES], inc[SCAN_TILES];
#pragma unroll
        for (int t = 0; t < SCAN_TILES; ++t) {
            const int i = base + t * 1024 + (int)threadIdx.x;
            v[t] = i < n ? cnt[i] : 0;
        }
#pragma unroll
        for (int t = 0; t < SCAN_TILES; ++t) {
            int x = v[t];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            inc[t] = x;                                   // inclusive within the warp
            if (lane == 31) wtot[t * 32 + warp] = x;
        }
        __syncthreads();
        if (warp == 0) {                                  // exclusive scan of the SCAN_TILES * 32 warp totals, tile-major
            int run = carry_s;
#pragma unroll
            for (int t = 0; t < SCAN_TILES; ++t) {
                const int w = wtot[t * 32 + lane];
                int x = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                wtot[t * 32 + lane] = run + x - w;
                run += __shfl_sync(0xffffffffu, x, 31);
            }
            if (lane == 0) carry_s = run;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < SCAN_TILES; ++t) {
            const int i = base + t * 1024 + (int)threadIdx.x;
            if (i < n) cnt[i] = wtot[t * 32 + warp] + inc[t] - v[t];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) state[3] = (unsigned int)carry_s;   // entries of the sorted list
}

__global__ void __launch_bounds__(256) bow_scatter_kernel(const int* __restrict__ job_sel, int n_q, int64_t n_jobs, int subs_per_loc,
                                                          int* __restrict__ bucket_pos, int2* __restrict__ sorted) {
    for (int64_t job = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; job < n_jobs; job += (int64_t)gridDim.x * blockDim.x) {
        const int loc = (int)(job / n_q), row = (int)(job - (int64_t)loc * n_q);
        const int base = loc * subs_per_loc;
#pragma unroll
        for (int which = 0; which < TOPG; ++which) {
            const int sel = job_sel[TOPG * job + which];
            if (sel < 0) continue;
#pragma unroll
            for (int j = 0; j < SUBS; ++j)
                if (sel & (1 << j)) {
                    const int sb = (sel >> 4) * SUBS + j;
                    sorted[atomicAdd(&bucket_pos[base + sb], 1)] = make_int2(row, base + sb);   // entry = (query row, bucket)
                }
        }
    }
}

// Pass 2: a warp takes RESOLVE_RUN consecutive entries of the sorted list (chosen by the host from the batch size: long runs
// re-use the sub-block, short runs keep every SM busy on small batches); the 8 code-book rows of the current (location,
// sub-block) stay in registers while the entries' bucket does not change, and the next entry and its query row are fetched
// while the current one is evaluated.  Lane l holds dims 4l..4l+3; it forms its partial sum for the 8 words, then the partials
// are summed ACROSS lanes by a transposing reduction (at distance 16 a lane keeps the half of the word list that matches its
// bit 4 and adds the partner's partials of those words, then distance 8 and 4: 4 + 2 + 1 shuffles), after which a plain
// butterfly over distances 2 and 1 leaves word i in lanes 4i .. 4i+3.  Every word's sum is formed by the same pairing
// (l, l^16), (l, l^8), ... as bow_rescan_kernel's butterfly, so the float32 distances are bit-identical to the rescan's and
// the SIMT kernel's tie-breaking is preserved.
__global__ void __launch_bounds__(128) bow_resolve_kernel(const int2* __restrict__ sorted, const unsigned int* __restrict__ state,
                                                          int RESOLVE_RUN, const float* __restrict__ q_desc,
                                                          int n_q, int subs_per_loc, const float* __restrict__ voc,
                                                          const int64_t* __restrict__ voc_off, unsigned long long* __restrict__ best) {
    static_assert(SUB_W == 8, "the reduction below is written for 8 words");
    const int lane = threadIdx.x & 31;
    const int64_t n_sorted = (int64_t)state[3];
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float4 cw[SUB_W];
    for (int64_t run = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; run * RESOLVE_RUN < n_sorted; run += n_warps) {
        int cur_bucket = -1, n_valid = 0, word0 = 0;
        int64_t job0 = 0;
        const int64_t begin = run * (int64_t)RESOLVE_RUN, end = min(n_sorted, begin + RESOLVE_RUN);
        int2 e_next = sorted[begin];                     // entry = (query row, bucket = location * subs_per_loc + sub-block)
        float4 o_next = reinterpret_cast<const float4*>(q_desc + (size_t)e_next.x * 128)[lane];
#pragma unroll 1
        for (int64_t i = begin; i < end; ++i) {
            const int2 e = e_next;
            const float4 o = o_next;
            if (i + 1 < end) {
                e_next = sorted[i + 1];
                o_next = reinterpret_cast<const float4*>(q_desc + (size_t)e_next.x * 128)[lane];
            }
            if (e.y != cur_bucket) {   // warp-uniform; the divisions are paid once per bucket, not per entry
                const int loc = e.y / subs_per_loc, sb = e.y - loc * subs_per_loc;
                const int64_t v0 = voc_off[loc];
                const int n_words = (int)(voc_off[loc + 1] - v0);
#pragma unroll
                for (int j = 0; j < SUB_W; ++j) {
                    const int w = min(sb * SUB_W + j, n_words - 1);
                    cw[j] = reinterpret_cast<const float4*>(voc + (v0 + w) * 128)[lane];
                }
                cur_bucket = e.y;
                word0 = sb * SUB_W;
                n_valid = n_words - word0;               // <= 0: a sub-block beyond the code book
                job0 = (int64_t)loc * n_q;
            }
            if (n_valid <= 0) continue;                  // warp-uniform
            float pw[SUB_W];
#pragma unroll
            for (int j = 0; j < SUB_W; ++j) {
                const float dx = o.x - cw[j].x, dy = o.y - cw[j].y, dz = o.z - cw[j].z, dw = o.w - cw[j].w;
                pw[j] = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
            }
            float p4[4], p2[2];
            {
                const bool up = (lane & 16) != 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float keep = up ? pw[j + 4] : pw[j], send = up ? pw[j] : pw[j + 4];
                    p4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            {
                const bool up = (lane & 8) != 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float keep = up ? p4[j + 2] : p4[j], send = up ? p4[j] : p4[j + 2];
                    p2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
            }
            float dist;
            {
                const bool up = (lane & 4) != 0;
                const float keep = up ? p2[1] : p2[0], send = up ? p2[0] : p2[1];
                dist = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            dist += __shfl_xor_sync(0xffffffffu, dist, 2);
            dist += __shfl_xor_sync(0xffffffffu, dist, 1);   // lanes 4i .. 4i+3 own word i of the sub-block now
            // smallest distance, lowest word on ties: non-negative floats order like their bit patterns, so a 32-bit min
            // over the warp and the first lane that holds it replace a 64-bit (distance, word) butterfly
            const unsigned key = ((lane >> 2) < n_valid) ? __float_as_uint(dist) : 0xffffffffu;
            const unsigned kmin = __reduce_min_sync(0xffffffffu, key);
            const unsigned who = __ballot_sync(0xffffffffu, key == kmin);
            if (lane == 0)
                atomicMin(&best[job0 + e.x], ((unsigned long long)kmin << 32) | (unsigned)(word0 + ((__ffs(who) - 1) >> 2)));
        }
    }
}

// many ambiguous rows (random-like data) -> one pass of the float32 SIMT kernel is cheaper than row-by-row rescans
__global__ void bow_decide_kernel(const int* __restrict__ bad, int64_t n_jobs, unsigned int* __restrict__ state) {
    const unsigned int thr = (unsigned int)max((int64_t)64, n_jobs / 16);
    state[1] = (*bad != 0 || state[0] > thr) ? 1u : 0u;
}

// Pass 3 (RESCAN_SLICES CTAs per queued row, each scanning one slice of the code book with 8 warps x 4 words in flight): exact
// float32 direct-difference scan, lowest index on ties, merged with a 64-bit atomicMin -- one CTA per row took 0.28 ms for a
// 10 000-word code book however few rows were queued (312 dependent trips through L2 per warp).
// best[] gets (distance bits << 32 | word) like bow_vq_kernel's atomicMin key (pre-set to all ones by the caller).
constexpr int RESCAN_SLICES = 64;   // 16 slices left 2 CTAs per SM for ~20 queued rows: 2.3 TB/s of a DRAM-bound scan
__global__ void __launch_bounds__(256) bow_rescan_kernel(const int* __restrict__ queue, const float* __restrict__ q_desc, int n_q,
                                                         const float* __restrict__ voc, const int64_t* __restrict__ voc_off,
                                                         unsigned long long* __restrict__ best, unsigned int* __restrict__ state) {
    if (state[1] != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned n = state[0];
    for (unsigned u = blockIdx.x; u < n * RESCAN_SLICES; u += gridDim.x) {
        const unsigned c = u / RESCAN_SLICES, slice = u % RESCAN_SLICES;
        const int job = queue[c];
        const int loc = job / n_q, row = job % n_q;
        const float4 o = reinterpret_cast<const float4*>(q_desc + (size_t)row * 128)[lane];
        const int64_t v0 = voc_off[loc];
        const int n_words = (int)(voc_off[loc + 1] - v0);
        const int per = ((n_words + RESCAN_SLICES - 1) / RESCAN_SLICES + 3) & ~3;   // words per slice, a multiple of 4
        const int w_lo = (int)slice * per, w_hi = min(n_words, w_lo + per);
        unsigned long long bk = ~0ULL;
        for (int w0 = w_lo + warp * 4; w0 < w_hi; w0 += 32) {
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
                if (w0 + j < w_hi) {
                    const unsigned long long key = ((unsigned long long)__float_as_uint(d[j]) << 32) | (unsigned)(w0 + j);
                    bk = key < bk ? key : bk;
                }
            }
        }
        if (lane == 0 && bk != ~0ULL) atomicMin(&best[job], bk);
        if (threadIdx.x == 0 && slice == 0) atomicAdd(&state[2], 1u);
    }
}

}  // namespace bowtc
}  // namespace mcl
