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

// Fast path for widths that are a multiple of 8 (every leaf then starts at a multiple of 8 and never straddles a pixel
// octet, N is a multiple of 8 and no leaf has leftovers): ONE THREAD PER LEAF.  Pixel p of a leaf belongs to accumulator
// (p - lo) mod 8 = p mod 8, so a thread that walks its leaf 8 consecutive pixels at a time holds all 8 accumulators in
// registers -- r[k] += x[8j + k] in j order, exactly numpy's loop -- with 8-fold instruction-level parallelism and no
// shuffles.  The image rows a CTA's 256 leaves touch (+ one reflected row above and below) are staged in shared memory (one
// bulk copy per row when the width is a multiple of 16, coalesced loads otherwise); the stencil runs on PAIRS of pixels held as two 16-bit lanes (see laplacian_sums_kernel); (x - mean)^2 is
// formed arithmetically: 2^52 + n as a bit pattern, minus (2^52 + 1020) = the Laplacian as an exact double, minus the mean
// (numpy's one rounding), squared (second rounding) -- 4 FP64-pipe operations per pixel, no table (random 8-byte reads of a
// 2041-entry table by 32 lanes cost ~4 shared-memory wavefronts each, more than the arithmetic).
// grid = (ceil(n_leaves / 256), images of one size); block = 256; dynamic shared memory = 256 * 128 + 4 W bytes.
constexpr int LAPV_LEAVES_PER_CTA = 256;

__global__ void __launch_bounds__(LAPV_LEAVES_PER_CTA) lap_leaf8_kernel(const uint8_t* __restrict__ pixels, const int64_t* __restrict__ img_off,
                                                        const int* __restrict__ pitches, const int* __restrict__ hs,
                                                        const int* __restrict__ ws, const int* __restrict__ img_list,
                                                        const long long* __restrict__ sums /* [n][2]: sum L, - */,
                                                        const int* __restrict__ leaf_lo, const int* __restrict__ leaf_n, int n_leaves,
                                                        int n_vals, double* __restrict__ vals) {
    extern __shared__ __align__(16) uint8_t rows[];
    const int slot = blockIdx.y;
    const int im = img_list[slot];
    const int H = hs[im], W = ws[im], pitch = pitches[im];
    const uint8_t* img = pixels + img_off[im];
    const double N = (double)((long long)H * (long long)W);
    const double mean = __ddiv_rn((double)sums[2 * im], N);
    const int leaf_first = blockIdx.x * LAPV_LEAVES_PER_CTA;
    const int leaf_last = min(n_leaves, leaf_first + LAPV_LEAVES_PER_CTA) - 1;
    const int p_lo = leaf_lo[leaf_first], p_hi = leaf_lo[leaf_last] + leaf_n[leaf_last];
    const int y_first = p_lo / W, y_last = (p_hi - 1) / W;
    const int n_rows = y_last - y_first + 3;   // one (reflected) row above and below
    if ((W & 15) == 0) {
        // one linear bulk copy (TMA engine) per row, all in flight at once, completing on an mbarrier
        __shared__ uint64_t bar;
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&bar, (uint32_t)(n_rows * W));
            for (int r = 0; r < n_rows; ++r)
                bulk_g2s(rows + (size_t)r * W, img + (int64_t)lapv_reflect(y_first - 1 + r, H) * pitch, (uint32_t)W, &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        const int chunks = W >> 3;
        for (int r = 0; r < n_rows; ++r) {
            const uint2* src = reinterpret_cast<const uint2*>(img + (int64_t)lapv_reflect(y_first - 1 + r, H) * pitch);
            uint2* dst = reinterpret_cast<uint2*>(rows + (size_t)r * W);
            for (int c = threadIdx.x; c < chunks; c += blockDim.x) dst[c] = src[c];
        }
        __syncthreads();
    }
    const int leaf = leaf_first + threadIdx.x;
    if (leaf > leaf_last) return;
    const int lo = leaf_lo[leaf], steps = leaf_n[leaf] >> 3;
    int y = lo / W, x = lo - y * W;
    double r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0, r6 = 0, r7 = 0;
    const double C0 = 4503599627370496.0 + 1020.0;   // 2^52 + 1020
    auto sq = [&](uint32_t n) {   // n = L + 1020 in [0, 2040] -> fl(fl(L - mean)^2)
        const double L = __dsub_rn(__hiloint2double(0x43300000, (int)n), C0);   // exact
        const double d = __dsub_rn(L, mean);
        return __dmul_rn(d, d);
    };
    for (int j = 0; j < steps; ++j) {
        const uint8_t* mid = rows + (size_t)(y - y_first + 1) * W + x;
        const uint2 m = *reinterpret_cast<const uint2*>(mid);
        const uint2 u = *reinterpret_cast<const uint2*>(mid - W);
        const uint2 d = *reinterpret_cast<const uint2*>(mid + W);
        // REFLECT_101: the pixel left of column 0 is column 1, right of column W-1 is column W-2
        const uint32_t lh = x > 0 ? mid[-1] : (m.x >> 8) & 0xffu;
        const uint32_t rh = x + 8 < W ? mid[8] : (m.y >> 16) & 0xffu;
        uint32_t pm[4], pu[4], pd[4], q[5];
        pm[0] = __byte_perm(m.x, 0u, 0x4140); pm[1] = __byte_perm(m.x, 0u, 0x4342);
        pm[2] = __byte_perm(m.y, 0u, 0x4140); pm[3] = __byte_perm(m.y, 0u, 0x4342);
        pu[0] = __byte_perm(u.x, 0u, 0x4140); pu[1] = __byte_perm(u.x, 0u, 0x4342);
        pu[2] = __byte_perm(u.y, 0u, 0x4140); pu[3] = __byte_perm(u.y, 0u, 0x4342);
        pd[0] = __byte_perm(d.x, 0u, 0x4140); pd[1] = __byte_perm(d.x, 0u, 0x4342);
        pd[2] = __byte_perm(d.y, 0u, 0x4140); pd[3] = __byte_perm(d.y, 0u, 0x4342);
        q[0] = __byte_perm(lh, pm[0], 0x5410);            // (px[-1], px[0])
        q[1] = __byte_perm(pm[0], pm[1], 0x5432);         // (px[1], px[2])
        q[2] = __byte_perm(pm[1], pm[2], 0x5432);
        q[3] = __byte_perm(pm[2], pm[3], 0x5432);
        q[4] = __byte_perm(pm[3], rh, 0x5432);            // (px[7], px[8])
        uint32_t lb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) lb[i] = pu[i] + pd[i] + q[i] + q[i + 1] + (1020u | (1020u << 16)) - 4u * pm[i];
        r0 = __dadd_rn(r0, sq(lb[0] & 0xffffu)); r1 = __dadd_rn(r1, sq(lb[0] >> 16));
        r2 = __dadd_rn(r2, sq(lb[1] & 0xffffu)); r3 = __dadd_rn(r3, sq(lb[1] >> 16));
        r4 = __dadd_rn(r4, sq(lb[2] & 0xffffu)); r5 = __dadd_rn(r5, sq(lb[2] >> 16));
        r6 = __dadd_rn(r6, sq(lb[3] & 0xffffu)); r7 = __dadd_rn(r7, sq(lb[3] >> 16));
        x += 8;
        if (x >= W) { x = 0; ++y; }
    }
    // j = 0 adds to +0.0, which numpy does not do (it starts from the first 8 values): x + 0.0 == x for every x >= 0, and the
    // squares are non-negative, so the sums are identical
    vals[(size_t)slot * n_vals + leaf] = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
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
