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

// Pass 1 (one thread per (location, row) job): merge the ranges' records; accepted jobs are counted into their group's
// bucket, the others queued for a full rescan.
// state: [0] = queued rows, [1] = run-the-SIMT-kernel flag (bow_decide_kernel), [2] = rows rescanned (statistics)
__global__ void __launch_bounds__(256) bow_merge_kernel(const int4* __restrict__ tc, int n_splits, int n_q, int n_loc, int groups_per_loc,
                                                        const int* __restrict__ bad, int* __restrict__ job_grp,
                                                        int* __restrict__ bucket_cnt, int* __restrict__ queue,
                                                        unsigned int* __restrict__ state) {
    const int64_t n_jobs = (int64_t)n_loc * n_q;
    const bool all_bad = *bad != 0;   // non-integer query rows: nothing from the tensor-core pass can be trusted
    for (int64_t job = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; job < n_jobs; job += (int64_t)gridDim.x * blockDim.x) {
        int b1 = sifttc3::NONE, b2 = sifttc3::NONE, grp = 0;   // positive float32 keys compare like their bit patterns
        for (int sp = 0; sp < n_splits; ++sp) {
            const int4 r = tc[(size_t)sp * n_jobs + job];
            if (r.y > b1) { b2 = max(b1, r.z); b1 = r.y; grp = r.x; }
            else { b2 = max(b2, r.y); }
        }
        bool accept = !all_bad && b1 > 0;
        if (accept && b2 > 0) {
            const float k1 = __int_as_float(b1), k2 = __int_as_float(b2);
            accept = (k1 - k2) > 2.0f * (KEY_REL * k1 + KEY_ABS);
        }
        if (accept) {
            job_grp[job] = grp;
            atomicAdd(&bucket_cnt[(int)(job / n_q) * groups_per_loc + grp], 1);
        } else {
            job_grp[job] = -1;
            queue[atomicAdd(&state[0], 1u)] = (int)job;
        }
    }
}

// exclusive prefix sum of the bucket counts (one CTA); cnt[] becomes the running fill position of every bucket
__global__ void __launch_bounds__(1024) bow_scan_kernel(int* __restrict__ cnt, int n) {
    __shared__ int part[1024];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += cnt[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - sum;
    for (int i = lo; i < hi; ++i) { const int c = cnt[i]; cnt[i] = run; run += c; }
}

__global__ void __launch_bounds__(256) bow_scatter_kernel(const int* __restrict__ job_grp, int n_q, int64_t n_jobs, int groups_per_loc,
                                                          int* __restrict__ bucket_pos, int* __restrict__ sorted) {
    for (int64_t job = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; job < n_jobs; job += (int64_t)gridDim.x * blockDim.x) {
        const int g = job_grp[job];
        if (g >= 0) sorted[atomicAdd(&bucket_pos[(int)(job / n_q) * groups_per_loc + g], 1)] = (int)job;
    }
}

// Pass 2: a warp takes RUN consecutive jobs of the sorted list; the 32 code-book rows of the current (location, group) stay in
// registers while the jobs' bucket does not change.  Lane l holds dims 4l..4l+3; it forms its partial sum for all 32 words,
// then the partials are summed ACROSS lanes by a transposing reduction (at distance 16 a lane keeps the half of the word list
// that matches its bit 4 and adds the partner's partials of those words, and so on: 31 shuffles instead of 32 x 5), after
// which lane i owns word i.  Every word's sum is formed by the same pairing (l, l^16), (l, l^8), ... as bow_rescan_kernel's
// butterfly, so the float32 distances are bit-identical to the rescan's and the SIMT kernel's tie-breaking is preserved.
constexpr int RESOLVE_RUN = 8;
__global__ void __launch_bounds__(128) bow_resolve_kernel(const int* __restrict__ sorted, const unsigned int* __restrict__ state,
                                                          int64_t n_jobs, const int* __restrict__ job_grp, const float* __restrict__ q_desc,
                                                          int n_q, const float* __restrict__ voc, const int64_t* __restrict__ voc_off,
                                                          unsigned long long* __restrict__ best) {
    const int lane = threadIdx.x & 31;
    const int64_t n_sorted = n_jobs - (int64_t)state[0];   // accepted jobs
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float4 cw[GROUP_W];
    for (int64_t run = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; run * RESOLVE_RUN < n_sorted; run += n_warps) {
        int cur_loc = -1, cur_grp = -1, n_words = 0;
        const int64_t end = min(n_sorted, (run + 1) * RESOLVE_RUN);
        for (int64_t i = run * RESOLVE_RUN; i < end; ++i) {
            const int job = sorted[i];
            const int loc = job / n_q, row = job - loc * n_q, grp = job_grp[job];
            if (loc != cur_loc || grp != cur_grp) {   // warp-uniform
                const int64_t v0 = voc_off[loc];
                n_words = (int)(voc_off[loc + 1] - v0);
#pragma unroll
                for (int j = 0; j < GROUP_W; ++j) {
                    const int w = min(grp * GROUP_W + j, n_words - 1);
                    cw[j] = reinterpret_cast<const float4*>(voc + (v0 + w) * 128)[lane];
                }
                cur_loc = loc;
                cur_grp = grp;
            }
            const float4 o = reinterpret_cast<const float4*>(q_desc + (size_t)row * 128)[lane];
            float pw[GROUP_W];
#pragma unroll
            for (int j = 0; j < GROUP_W; ++j) {
                const float dx = o.x - cw[j].x, dy = o.y - cw[j].y, dz = o.z - cw[j].z, dw = o.w - cw[j].w;
                pw[j] = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
            }
            static_assert(GROUP_W == 32, "the reduction below is written for 32 words");
            float p16[16], p8[8], p4[4], p2[2];
            {
                const bool up = (lane & 16) != 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float keep = up ? pw[j + 16] : pw[j], send = up ? pw[j] : pw[j + 16];
                    p16[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            {
                const bool up = (lane & 8) != 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float keep = up ? p16[j + 8] : p16[j], send = up ? p16[j] : p16[j + 8];
                    p8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
            }
            {
                const bool up = (lane & 4) != 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float keep = up ? p8[j + 4] : p8[j], send = up ? p8[j] : p8[j + 4];
                    p4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
            }
            {
                const bool up = (lane & 2) != 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float keep = up ? p4[j + 2] : p4[j], send = up ? p4[j] : p4[j + 2];
                    p2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
            }
            float dist;
            {
                const bool up = (lane & 1) != 0;
                const float keep = up ? p2[1] : p2[0], send = up ? p2[0] : p2[1];
                dist = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
            const int wi = grp * GROUP_W + lane;   // lane i owns word i of the group
            unsigned long long bk = (wi < n_words) ? (((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)wi) : ~0ULL;
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, bk, k);
                bk = other < bk ? other : bk;
            }
            if (lane == 0) best[job] = bk;
        }
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
