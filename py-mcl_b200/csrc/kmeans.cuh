// Lloyd iterations for the BOW codebook (Matcher.createIndex, Matcher.py:124: `voc, variance = kmeans(descriptors, numWords, 1)`).
// Mirrors scipy.cluster.vq._kmeans (scipy/cluster/vq/_vq_impl.py:253-275): assign (the tensor-core vq path), average
// distortion, centroid = mean of its members, empty clusters dropped, stop when the average distortion moves by <= thresh.
// Observations are integer valued (SIFT), so the float32 member sums are exact integers and order independent: given the
// same assignments the centroids are bit-identical to scipy's sequential float32 accumulation.
#pragma once
#include "common.cuh"

namespace mcl {

// sum_i sqrtf(||o_i - c_{word(i)}||^2)   (scipy's `distort`), accumulated in float64.  One warp per observation.
__global__ void __launch_bounds__(256) kmeans_distort_kernel(const float* __restrict__ obs, int64_t n, const float* __restrict__ voc,
                                                             const unsigned long long* __restrict__ best, double* __restrict__ sum) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_w = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double acc = 0.0;
    for (int64_t i = wid; i < n; i += n_w) {
        const int w = (int)(best[i] & 0xffffffffu);
        const float4 o = reinterpret_cast<const float4*>(obs + i * 128)[lane];
        const float4 c = reinterpret_cast<const float4*>(voc + (int64_t)w * 128)[lane];
        const float dx = o.x - c.x, dy = o.y - c.y, dz = o.z - c.z, dw = o.w - c.w;
        float d = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) d += __shfl_xor_sync(0xffffffffu, d, k);
        acc += (double)sqrtf(d);
    }
    if (lane == 0 && acc != 0.0) atomicAdd(sum, acc);
}

// member sums and counts.  One warp per observation; lane l owns dims 4l..4l+3.
__global__ void __launch_bounds__(256) kmeans_accum_kernel(const float* __restrict__ obs, int64_t n,
                                                           const unsigned long long* __restrict__ best,
                                                           float* __restrict__ sums /* k x 128 */, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_w = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = wid; i < n; i += n_w) {
        const int w = (int)(best[i] & 0xffffffffu);
        const float4 o = reinterpret_cast<const float4*>(obs + i * 128)[lane];
        float* dst = sums + (int64_t)w * 128 + 4 * lane;
        atomicAdd(dst, o.x); atomicAdd(dst + 1, o.y); atomicAdd(dst + 2, o.z); atomicAdd(dst + 3, o.w);
        if (lane == 0) atomicAdd(&counts[w], 1);
    }
}

// new code book = member means, empty clusters dropped (order preserved).  One block.
__global__ void __launch_bounds__(1024) kmeans_update_kernel(int k, const float* __restrict__ sums, const int* __restrict__ counts,
                                                             float* __restrict__ voc_out, int* __restrict__ k_new) {
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int w0 = 0; w0 < k; w0 += blockDim.x) {
        const int w = w0 + threadIdx.x;
        const int cnt = w < k ? counts[w] : 0;
        const bool has = cnt > 0;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int j = 0; j < warp; ++j) off += warp_tot[j];
        const int pos = off + __popc(bal & ((1u << lane) - 1u));
        if (has) {
            const float inv_n = (float)cnt;
            for (int d = 0; d < 128; ++d) voc_out[(int64_t)pos * 128 + d] = sums[(int64_t)w * 128 + d] / inv_n;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int j = 0; j < (int)(blockDim.x >> 5); ++j) t += warp_tot[j];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *k_new = base;
}

}  // namespace mcl
