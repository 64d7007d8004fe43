// Variance of the Laplacian, bit-identical to the reference's  cv2.Laplacian(img, cv2.CV_64F).var()  (analyze.py:441-445).
//
// ndarray.var() is numpy's two-pass variance:  mean = add.reduce(x) / N;  var = add.reduce((x - mean)^2) / N, and add.reduce
// over a contiguous float64 array is PAIRWISE summation: the range is halved recursively (left half rounded down to a
// multiple of 8) until a piece has <= 128 elements; a piece is summed with 8 interleaved accumulators r[k] += x[8j + k],
// combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the < 8 leftover elements are added one by one.  The Laplacian
// values are integers in [-1020, 1020], so the first sum is exact in any order (the mean is one correctly rounded division of
// exact integers; sum L itself telescopes to the image border, see laplacian_finish_kernel), but the second sum rounds at
// every addition: its value depends on the order, and the blur weight blur/200*0.85 of analyzer.probUpdate (analyze.py:245-248)
// carries every bit of it into out.txt.  These kernels perform exactly numpy's additions:
//   * lap_leaf_kernel: 8 lanes per leaf (one per accumulator, so a warp sums 4 leaves and reads 32-byte runs of pixels),
//     (x - mean)^2 from a 2041-entry float64 table in shared memory (x is an integer: the table holds fl(fl(x - mean)^2), the
//     two roundings numpy performs per element), butterfly over the 8 lanes = the combination tree, leftovers by lane 0;
//   * lap_tree_kernel: the recursion's internal nodes level by level (one CTA per image), root / N.
// The leaf / node tables depend only on N = H*W and are built once per image size on the host (LapPlan).
#pragma once
#include "common.cuh"

namespace mcl {

constexpr int LAPV_LUT = 2041;   // Laplacian values -1020 .. 1020

__device__ __forceinline__ int lapv_reflect(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// 4-neighbour Laplacian of pixel (y, x), BORDER_REFLECT_101, + 1020 (table index)
__device__ __forceinline__ int lapv_value(const uint8_t* __restrict__ img, int H, int W, int pitch, int y, int x) {
    const uint8_t* row = img + (int64_t)y * pitch;
    int up, dn, lf, rt;
    if (y > 0 && y < H - 1 && x > 0 && x < W - 1) {
        up = row[x - pitch]; dn = row[x + pitch]; lf = row[x - 1]; rt = row[x + 1];
    } else {
        up = img[(int64_t)lapv_reflect(y - 1, H) * pitch + x];
        dn = img[(int64_t)lapv_reflect(y + 1, H) * pitch + x];
        lf = row[lapv_reflect(x - 1, W)];
        rt = row[lapv_reflect(x + 1, W)];
    }
    return up + dn + lf + rt - 4 * (int)row[x] + 1020;
}

// grid = (blocks, images of one size); block = 256 threads = 32 leaves per pass.
// leaf_lo / leaf_n: the plan's leaves (flattened pixel range).  vals: per image n_vals doubles, leaf sums first.
__global__ void __launch_bounds__(256) lap_leaf_kernel(const uint8_t* __restrict__ pixels, const int64_t* __restrict__ img_off,
                                                       const int* __restrict__ pitches, const int* __restrict__ hs,
                                                       const int* __restrict__ ws, const int* __restrict__ img_list,
                                                       const long long* __restrict__ sums /* [n][2]: sum L, - */,
                                                       const int* __restrict__ leaf_lo, const int* __restrict__ leaf_n, int n_leaves,
                                                       int n_vals, double* __restrict__ vals) {
    __shared__ double lut[LAPV_LUT];
    const int slot = blockIdx.y;
    const int im = img_list[slot];
    const int H = hs[im], W = ws[im], pitch = pitches[im];
    const uint8_t* img = pixels + img_off[im];
    const double N = (double)((long long)H * (long long)W);
    const double mean = __ddiv_rn((double)sums[2 * im], N);
    for (int v = threadIdx.x; v < LAPV_LUT; v += blockDim.x) {
        const double d = __dsub_rn((double)(v - 1020), mean);
        lut[v] = __dmul_rn(d, d);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, k = lane & 7;
    double* out = vals + (size_t)slot * n_vals;
    for (int leaf0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4; leaf0 < n_leaves;
         leaf0 += gridDim.x * (blockDim.x >> 5) * 4) {
        const int leaf = leaf0 + (lane >> 3);
        const bool live = leaf < n_leaves;
        const int lo = live ? leaf_lo[leaf] : 0, n = live ? leaf_n[leaf] : 0;
        double res = 0.0;
        if (n >= 8) {
            const int steps = n >> 3;
            int p = lo + k;
            int y = p / W, x = p - y * W;
            double r = lut[lapv_value(img, H, W, pitch, y, x)];
            for (int j = 1; j < steps; ++j) {
                x += 8;
                while (x >= W) { x -= W; ++y; }
                r = __dadd_rn(r, lut[lapv_value(img, H, W, pitch, y, x)]);
            }
            res = r;
        }
        // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)): xor butterfly over the 8 lanes of the leaf (addition is commutative, so both
        // partners compute the same sum); groups with n < 8 carry zeros through it and are summed below
        res = __dadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 1));
        res = __dadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 2));
        res = __dadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 4));
        if (live && k == 0) {
            int first = lo + (n & ~7);
            if (n < 8) { res = 0.0; first = lo; }   // numpy: res = 0.; for (i < n) res += a[i]
            for (int p = first; p < lo + n; ++p) {
                const int y = p / W, x = p - y * W;
                res = __dadd_rn(res, lut[lapv_value(img, H, W, pitch, y, x)]);
            }
            out[leaf] = res;
        }
    }
}

// grid = images of one size; one CTA walks the internal nodes of the recursion level by level (children first).
// node_l / node_r index into vals (leaves 0 .. n_leaves-1, internal node i at n_leaves + i); level_off[h] .. level_off[h+1] are
// the nodes of height h+1.  The root is the last node (or leaf 0 when there is a single leaf).
__global__ void __launch_bounds__(256) lap_tree_kernel(const int* __restrict__ node_l, const int* __restrict__ node_r,
                                                       const int* __restrict__ level_off, int n_levels, int n_leaves, int n_vals,
                                                       const int* __restrict__ img_list, const int* __restrict__ hs,
                                                       const int* __restrict__ ws, double* __restrict__ vals,
                                                       double* __restrict__ out) {
    const int slot = blockIdx.x;
    double* v = vals + (size_t)slot * n_vals;
    for (int h = 0; h < n_levels; ++h) {
        for (int i = level_off[h] + threadIdx.x; i < level_off[h + 1]; i += blockDim.x)
            v[n_leaves + i] = __dadd_rn(v[node_l[i]], v[node_r[i]]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int im = img_list[slot];
        const double N = (double)((long long)hs[im] * (long long)ws[im]);
        out[im] = __ddiv_rn(v[n_vals - 1], N);
    }
}

}  // namespace mcl
