// Colour-histogram search: chi-squared distance of every query histogram to every map histogram.
//
// Reference: search.py:7-26 (Searcher.search / chisquared), called from Matcher.colorSearch (Matcher.py:89-99):
//     d = 0.5 * np.sum([((a - b) ** 2) / (a + b + eps) for (a, b) in zip(histA, histB)])
// With float32 histograms (cv2.calcHist + normalize) every term is a float32 and np.sum reduces the float32 array with
// numpy's pairwise summation (blocks of <= 128 elements, 8 interleaved accumulators per block, blocks combined by a binary
// tree whose split point is rounded down to a multiple of 8).  The kernels below perform exactly those float32
// operations in exactly that order -- the result is the reference's np.float32 bit for bit.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mcl {

__device__ __forceinline__ float chi2_term(float a, float b, float eps) {
    const float d = __fsub_rn(a, b);
    return __fdiv_rn(__fmul_rn(d, d), __fadd_rn(__fadd_rn(a, b), eps));
}

// numpy's pairwise sum of n (<= 128) terms, one thread
__device__ float chi2_block_serial(const float* __restrict__ a, const float* __restrict__ b, int n, float eps) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, chi2_term(a[i], b[i], eps));
        return r;
    }
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = chi2_term(a[k], b[k], eps);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], chi2_term(a[i + k], b[i + k], eps));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, chi2_term(a[i], b[i], eps));
    return res;
}

// any number of bins: explicit stack instead of recursion (depth <= log2(n/128) + 1)
__device__ float chi2_pairwise_serial(const float* __restrict__ a, const float* __restrict__ b, int n, float eps) {
    // post-order evaluation of the split tree: each stack entry is a pending (offset, length, state, left value)
    int off[24], len[24], state[24];
    float left[24];
    int sp = 0;
    off[0] = 0; len[0] = n; state[0] = 0;
    float ret = 0.f;
    while (sp >= 0) {
        const int o = off[sp], l = len[sp];
        if (l <= 128) {
            ret = chi2_block_serial(a + o, b + o, l, eps);
            --sp;
            continue;
        }
        int n2 = l / 2;
        n2 -= n2 % 8;
        if (state[sp] == 0) {
            state[sp] = 1;
            ++sp; off[sp] = o; len[sp] = n2; state[sp] = 0;
        } else if (state[sp] == 1) {
            left[sp] = ret;
            state[sp] = 2;
            ++sp; off[sp] = o + n2; len[sp] = l - n2; state[sp] = 0;
        } else {
            ret = __fadd_rn(left[sp], ret);
            --sp;
        }
    }
    return ret;
}

__global__ void chi2_generic_kernel(const float* __restrict__ q, const float* __restrict__ m, int n_q, int n_map, int bins,
                                    float eps, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_q * n_map) return;
    const int qi = (int)(i / n_map), mi = (int)(i % n_map);
    out[i] = __fmul_rn(0.5f, chi2_pairwise_serial(q + (size_t)qi * bins, m + (size_t)mi * bins, bins, eps));
}

// 512 bins (the reference's [8, 8, 8] histogram): one warp per (query, map image) pair.  The split tree of 512 terms is
// ((B0 + B1) + (B2 + B3)) over four 128-term blocks; lane = 8 * block + accumulator, each lane adds its 16 terms in index
// order, then xor-butterflies over lane bits 0,1,2 give the block sums ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and over
// bits 3,4 the tree of blocks (float addition is commutative, so which lane of a pair holds which operand is irrelevant).
constexpr int CHI2_WARPS = 8;
__global__ void __launch_bounds__(CHI2_WARPS * 32) chi2_512_kernel(const float* __restrict__ q, const float* __restrict__ m,
                                                                   int n_q, int n_map, float eps, float* __restrict__ out) {
    __shared__ float qs[512];
    const int qi = blockIdx.y;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) qs[i] = q[(size_t)qi * 512 + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int base = (lane >> 3) * 128 + (lane & 7);
    for (int mi = blockIdx.x * CHI2_WARPS + warp; mi < n_map; mi += gridDim.x * CHI2_WARPS) {
        const float* mp = m + (size_t)mi * 512;
        float r = chi2_term(qs[base], __ldg(mp + base), eps);
#pragma unroll
        for (int i = 1; i < 16; ++i) r = __fadd_rn(r, chi2_term(qs[base + 8 * i], __ldg(mp + base + 8 * i), eps));
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, s));
        if (lane == 0) out[(size_t)qi * n_map + mi] = __fmul_rn(0.5f, r);
    }
}

}  // namespace mcl
