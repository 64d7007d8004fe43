// SIMT matching kernels: exact integer-L2 (SIFT) reference path, ORB Hamming (XOR + popc), SURF fp32.
// One CTA per (frame, image) pair [x query row block]; the thread owns one query row in registers, the map
// image is staged through shared memory and read as warp-wide broadcasts.
#pragma once
#include "common.cuh"

namespace mcl {

// ------------------------------------------------------------------------------------------
// SIFT, exact top-2 with dp4a.  Used for small / masked (DOR) problems and as the on-device cross-check of
// the tensor-core path.  Same HBM layouts as sift_tc (SW128 atoms + norms).
// grid = (n_images, n_frames), block = 128.
// ------------------------------------------------------------------------------------------
constexpr int SIMT_STAGE_ROWS = 64;

__device__ __forceinline__ void sift_simt_pair(
    const int img, const int frame,
    const uint8_t* __restrict__ q_desc, const int32_t* __restrict__ q_norm, const int64_t* __restrict__ q_dst_off,
    const uint8_t* __restrict__ m_desc, const int32_t* __restrict__ m_norm, const int64_t* __restrict__ m_dst_off,
    const int64_t* __restrict__ m_src_off, int n_images, double ratio,
    int32_t* __restrict__ counts, unsigned long long* __restrict__ n_blocks_run /* optional statistics */) {
    if (n_blocks_run && threadIdx.x == 0) atomicAdd(n_blocks_run, 1ULL);
    __shared__ uint4 sm_rows[SIMT_STAGE_ROWS][8];
    __shared__ int sm_norm[SIMT_STAGE_ROWS];
    __shared__ int sm_count;
    const int64_t q0 = q_dst_off[frame], q1 = q_dst_off[frame + 1];
    const int64_t m0 = m_dst_off[img];
    const int nm_rows = (int)(m_src_off[img + 1] - m_src_off[img]);
    if (threadIdx.x == 0) sm_count = 0;
    __syncthreads();
    if (nm_rows < 2) return;  // reference raises; defined as 0
    for (int64_t qb = q0; qb < q1; qb += blockDim.x) {
        const int64_t qrow = qb + threadIdx.x;
        const bool active = qrow < q1;
        uint4 q[8];
        int Q = 0;
        if (active) {
#pragma unroll
            for (int c = 0; c < 8; ++c) q[c] = *reinterpret_cast<const uint4*>(q_desc + sw128_chunk_offset(qrow, c));
            Q = q_norm[qrow];
        }
        int best = INT_MAX, second = INT_MAX;  // in units of (|m|^2 - 2 q.m)
        for (int base = 0; base < nm_rows; base += SIMT_STAGE_ROWS) {
            const int nrow = min(SIMT_STAGE_ROWS, nm_rows - base);
            __syncthreads();
            for (int i = threadIdx.x; i < nrow * 8; i += blockDim.x) {
                const int r = i >> 3, c = i & 7;
                sm_rows[r][c] = *reinterpret_cast<const uint4*>(m_desc + sw128_chunk_offset(m0 + base + r, c));
            }
            for (int i = threadIdx.x; i < nrow; i += blockDim.x) sm_norm[i] = m_norm[m0 + base + i];
            __syncthreads();
            if (active) {
                for (int r = 0; r < nrow; ++r) {
                    unsigned dot = 0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 b = sm_rows[r][c];
                        dot = __dp4a(q[c].x, b.x, dot); dot = __dp4a(q[c].y, b.y, dot);
                        dot = __dp4a(q[c].z, b.z, dot); dot = __dp4a(q[c].w, b.w, dot);
                    }
                    const int t = sm_norm[r] - (int)(2u * dot);
                    second = min(second, max(best, t));
                    best = min(best, t);
                }
            }
        }
        const bool pass = active && ratio_pass(Q + best, Q + second, ratio);
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&sm_count, __popc(bal));
    }
    __syncthreads();
    if (threadIdx.x == 0) counts[(int64_t)frame * n_images + img] = sm_count;
}

__global__ void __launch_bounds__(128) sift_simt_kernel(
    const uint8_t* __restrict__ q_desc, const int32_t* __restrict__ q_norm, const int64_t* __restrict__ q_dst_off,
    const uint8_t* __restrict__ m_desc, const int32_t* __restrict__ m_norm, const int64_t* __restrict__ m_dst_off,
    const int64_t* __restrict__ m_src_off, int n_images, double ratio, const uint8_t* __restrict__ mask,
    int32_t* __restrict__ counts, unsigned long long* __restrict__ n_blocks_run /* optional statistics */) {
    const int img = blockIdx.x, frame = blockIdx.y;
    if (mask && !mask[(int64_t)frame * n_images + img]) return;
    sift_simt_pair(img, frame, q_desc, q_norm, q_dst_off, m_desc, m_norm, m_dst_off, m_src_off, n_images, ratio, counts, n_blocks_run);
}

// The same, driven by a device-side list of (frame * n_images + image) pairs: the repair pass of the tensor-core paths.  A
// fixed small grid walks the list, so a call without dirty pairs costs two empty launches (a grid of one CTA per (frame,
// image) whose CTAs all return at once cost 0.05 ms at 80 000 pairs, 0.15 ms for ORB's 320 000).
__global__ void __launch_bounds__(128) sift_simt_list_kernel(
    const int* __restrict__ list, const unsigned int* __restrict__ n_list,
    const uint8_t* __restrict__ q_desc, const int32_t* __restrict__ q_norm, const int64_t* __restrict__ q_dst_off,
    const uint8_t* __restrict__ m_desc, const int32_t* __restrict__ m_norm, const int64_t* __restrict__ m_dst_off,
    const int64_t* __restrict__ m_src_off, int n_images, double ratio,
    int32_t* __restrict__ counts, unsigned long long* __restrict__ n_blocks_run) {
    const unsigned n = *n_list;
    for (unsigned e = blockIdx.x; e < n; e += gridDim.x) {
        const int pair = list[e];
        sift_simt_pair(pair % n_images, pair / n_images, q_desc, q_norm, q_dst_off, m_desc, m_norm, m_dst_off, m_src_off, n_images,
                       ratio, counts, n_blocks_run);
        __syncthreads();
    }
}

// pairs flagged in a byte map -> list, only when *only_if != 0 (no dirty pair: nothing to do)
__global__ void compact_dirty_kernel(const uint8_t* __restrict__ dirty, int64_t n, const int* __restrict__ only_if,
                                     int* __restrict__ list, unsigned int* __restrict__ n_list) {
    if (*only_if == 0) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (dirty[i]) list[atomicAdd(n_list, 1u)] = (int)i;
}

// ------------------------------------------------------------------------------------------
// ORB: 256-bit Hamming distance = 8 x (XOR + popc.b32); rows are 32 bytes = 2 x 128-bit loads.
//   replaces Matcher.ORBMatch (Matcher.py:169-194).  mode 0: 2-NN + ratio test (north_star);
//   mode 1 uses the two crossCheck kernels below (mutual first-index nearest neighbours).
// grid = (n_images, n_frames, row blocks of 128), block = 128.
// ------------------------------------------------------------------------------------------
constexpr int ORB_STAGE_ROWS = 256;

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ __forceinline__ void orb_knn_pair(
    const int img, const int frame, const int zblock,
    const uint4* __restrict__ q_desc, const int64_t* __restrict__ q_off, const uint4* __restrict__ m_desc,
    const int64_t* __restrict__ m_off, int n_images, double ratio,
    int32_t* __restrict__ counts, unsigned long long* __restrict__ n_pairs_run /* optional statistics */) {
    if (n_pairs_run && zblock == 0 && threadIdx.x == 0) atomicAdd(n_pairs_run, 1ULL);
    __shared__ uint4 sm[ORB_STAGE_ROWS][2];
    const int64_t q0 = q_off[frame], q1 = q_off[frame + 1];
    const int64_t m0 = m_off[img];
    const int nm_rows = (int)(m_off[img + 1] - m0);
    if (nm_rows < 2) return;
    const int64_t qrow = q0 + (int64_t)zblock * blockDim.x + threadIdx.x;
    if (q0 + (int64_t)zblock * blockDim.x >= q1) return;
    const bool active = qrow < q1;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (active) { a0 = q_desc[qrow * 2]; a1 = q_desc[qrow * 2 + 1]; }
    int best = INT_MAX, second = INT_MAX;
    for (int base = 0; base < nm_rows; base += ORB_STAGE_ROWS) {
        const int nrow = min(ORB_STAGE_ROWS, nm_rows - base);
        __syncthreads();
        for (int i = threadIdx.x; i < nrow * 2; i += blockDim.x) sm[i >> 1][i & 1] = m_desc[(m0 + base) * 2 + i];
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < nrow; ++r) {
            const int d = hamming256(a0, a1, sm[r][0], sm[r][1]);
            second = min(second, max(best, d));
            best = min(best, d);
        }
    }
    // Matcher.py:280 on Hamming distances reported as floats: float(d0) < 0.7*float(d1)
    const bool pass = active && ((double)best < __dmul_rn(ratio, (double)second));
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&counts[(int64_t)frame * n_images + img], __popc(bal));
}

__global__ void __launch_bounds__(128) orb_knn_kernel(
    const uint4* __restrict__ q_desc, const int64_t* __restrict__ q_off, const uint4* __restrict__ m_desc,
    const int64_t* __restrict__ m_off, int n_images, double ratio, const uint8_t* __restrict__ mask,
    int32_t* __restrict__ counts, unsigned long long* __restrict__ n_pairs_run /* optional statistics */) {
    const int img = blockIdx.x, frame = blockIdx.y;
    if (mask && !mask[(int64_t)frame * n_images + img]) return;
    orb_knn_pair(img, frame, blockIdx.z, q_desc, q_off, m_desc, m_off, n_images, ratio, counts, n_pairs_run);
}

// list-driven variant (see sift_simt_list_kernel): entry e = (pair list[e / z_blocks], query row block e % z_blocks)
__global__ void __launch_bounds__(128) orb_knn_list_kernel(
    const int* __restrict__ list, const unsigned int* __restrict__ n_list, int z_blocks,
    const uint4* __restrict__ q_desc, const int64_t* __restrict__ q_off, const uint4* __restrict__ m_desc,
    const int64_t* __restrict__ m_off, int n_images, double ratio,
    int32_t* __restrict__ counts, unsigned long long* __restrict__ n_pairs_run) {
    const int64_t n = (int64_t)*n_list * z_blocks;
    for (int64_t e = blockIdx.x; e < n; e += gridDim.x) {
        const int pair = list[e / z_blocks];
        orb_knn_pair(pair % n_images, pair / n_images, (int)(e % z_blocks), q_desc, q_off, m_desc, m_off, n_images, ratio, counts,
                     n_pairs_run);
        __syncthreads();
    }
}

// crossCheck pass 1: first-index argmin over the map image for every query row
__global__ void __launch_bounds__(128) orb_rowargmin_kernel(
    const uint4* __restrict__ q_desc, const int64_t* __restrict__ q_off, const uint4* __restrict__ m_desc,
    const int64_t* __restrict__ m_off, int n_images, const uint8_t* __restrict__ mask,
    int32_t* __restrict__ row_arg /* [q_rows_total][n_images] */) {
    const int img = blockIdx.x, frame = blockIdx.y;
    if (mask && !mask[(int64_t)frame * n_images + img]) return;
    __shared__ uint4 sm[ORB_STAGE_ROWS][2];
    const int64_t q0 = q_off[frame], q1 = q_off[frame + 1];
    const int64_t m0 = m_off[img];
    const int nm_rows = (int)(m_off[img + 1] - m0);
    if (q0 + (int64_t)blockIdx.z * blockDim.x >= q1) return;
    const int64_t qrow = q0 + (int64_t)blockIdx.z * blockDim.x + threadIdx.x;
    const bool active = qrow < q1;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (active) { a0 = q_desc[qrow * 2]; a1 = q_desc[qrow * 2 + 1]; }
    int best = INT_MAX, arg = -1;
    for (int base = 0; base < nm_rows; base += ORB_STAGE_ROWS) {
        const int nrow = min(ORB_STAGE_ROWS, nm_rows - base);
        __syncthreads();
        for (int i = threadIdx.x; i < nrow * 2; i += blockDim.x) sm[i >> 1][i & 1] = m_desc[(m0 + base) * 2 + i];
        __syncthreads();
        for (int r = 0; r < nrow; ++r) {
            const int d = hamming256(a0, a1, sm[r][0], sm[r][1]);
            if (d < best) { best = d; arg = base + r; }
        }
    }
    if (active) row_arg[qrow * n_images + img] = arg;
}

// crossCheck pass 2: thread per map row: first-index argmin over the query rows, mutual check, count
__global__ void __launch_bounds__(128) orb_crosscheck_kernel(
    const uint4* __restrict__ q_desc, const int64_t* __restrict__ q_off, const uint4* __restrict__ m_desc,
    const int64_t* __restrict__ m_off, int n_images, const uint8_t* __restrict__ mask,
    const int32_t* __restrict__ row_arg, int32_t* __restrict__ counts) {
    const int img = blockIdx.x, frame = blockIdx.y;
    if (mask && !mask[(int64_t)frame * n_images + img]) return;
    __shared__ uint4 sm[ORB_STAGE_ROWS][2];
    const int64_t q0 = q_off[frame], q1 = q_off[frame + 1];
    const int nq_rows = (int)(q1 - q0);
    const int64_t m0 = m_off[img], m1 = m_off[img + 1];
    if (m0 + (int64_t)blockIdx.z * blockDim.x >= m1) return;
    const int64_t mrow = m0 + (int64_t)blockIdx.z * blockDim.x + threadIdx.x;
    const bool active = mrow < m1;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (active) { a0 = m_desc[mrow * 2]; a1 = m_desc[mrow * 2 + 1]; }
    int best = INT_MAX, arg = -1;
    for (int base = 0; base < nq_rows; base += ORB_STAGE_ROWS) {
        const int nrow = min(ORB_STAGE_ROWS, nq_rows - base);
        __syncthreads();
        for (int i = threadIdx.x; i < nrow * 2; i += blockDim.x) sm[i >> 1][i & 1] = q_desc[(q0 + base) * 2 + i];
        __syncthreads();
        for (int r = 0; r < nrow; ++r) {
            const int d = hamming256(a0, a1, sm[r][0], sm[r][1]);
            if (d < best) { best = d; arg = base + r; }
        }
    }
    const bool pass = active && arg >= 0 && row_arg[(q0 + arg) * n_images + img] == (int)(mrow - m0);
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&counts[(int64_t)frame * n_images + img], __popc(bal));
}

// ------------------------------------------------------------------------------------------
// SURF: 64-d float32, direct sum of squared differences accumulated in fp32 (FADD + FFMA), like cv2's
// normL2Sqr; sqrt + float64 ratio compare.  replaces Matcher.SURFMatch (Matcher.py:196-230).
// grid = (n_images, n_frames, row blocks of 128), block = 128, one query row (64 floats) per thread.
// ------------------------------------------------------------------------------------------
constexpr int SURF_STAGE_ROWS = 64;

__global__ void __launch_bounds__(128) surf_knn_kernel(
    const float4* __restrict__ q_desc, const int64_t* __restrict__ q_off, const float4* __restrict__ m_desc,
    const int64_t* __restrict__ m_off, int n_images, double ratio, const uint8_t* __restrict__ mask,
    int32_t* __restrict__ counts) {
    const int img = blockIdx.x, frame = blockIdx.y;
    if (mask && !mask[(int64_t)frame * n_images + img]) return;
    __shared__ float4 sm[SURF_STAGE_ROWS][16];
    const int64_t q0 = q_off[frame], q1 = q_off[frame + 1];
    const int64_t m0 = m_off[img];
    const int nm_rows = (int)(m_off[img + 1] - m0);
    if (nm_rows < 2) return;
    if (q0 + (int64_t)blockIdx.z * blockDim.x >= q1) return;
    const int64_t qrow = q0 + (int64_t)blockIdx.z * blockDim.x + threadIdx.x;
    const bool active = qrow < q1;
    float4 a[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) a[c] = active ? q_desc[qrow * 16 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    float best = INFINITY, second = INFINITY;
    for (int base = 0; base < nm_rows; base += SURF_STAGE_ROWS) {
        const int nrow = min(SURF_STAGE_ROWS, nm_rows - base);
        __syncthreads();
        for (int i = threadIdx.x; i < nrow * 16; i += blockDim.x) sm[i >> 4][i & 15] = m_desc[(m0 + base) * 16 + i];
        __syncthreads();
        for (int r = 0; r < nrow; ++r) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 b = sm[r][c];
                const float dx = a[c].x - b.x, dy = a[c].y - b.y, dz = a[c].z - b.z, dw = a[c].w - b.w;
                s0 = fmaf(dx, dx, s0); s1 = fmaf(dy, dy, s1); s2 = fmaf(dz, dz, s2); s3 = fmaf(dw, dw, s3);
            }
            const float d = (s0 + s1) + (s2 + s3);
            second = fminf(second, fmaxf(best, d));
            best = fminf(best, d);
        }
    }
    const float f0 = __fsqrt_rn(best), f1 = __fsqrt_rn(second);
    const bool pass = active && ((double)f0 < __dmul_rn(ratio, (double)f1));
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&counts[(int64_t)frame * n_images + img], __popc(bal));
}

// The same arithmetic with the map image cut into S slices of rows, each scanned by its own group of 128 threads
// (thread = (query row, slice)); the S partial top-2 pairs of a query row are merged through shared memory.  A DOR step is one
// frame against <= 30 images (BASELINE config 4: 30 x 1024 query rows = 240 warps for 148 SMs): the slices are what fills the
// GPU there (0.41 -> see DESIGN.md).  The per-row sums are formed exactly as in surf_knn_kernel, so counts are identical.
// grid = (n_images, n_frames, row blocks of 128), block = 128 * S.
constexpr int SURF_SLICE_STAGE = 16;   // rows staged per slice and pass (4 slices x 16 rows x 256 B = 16 KB of static shared memory)

template <int S>
__global__ void __launch_bounds__(128 * S) surf_knn_sliced_kernel(
    const float4* __restrict__ q_desc, const int64_t* __restrict__ q_off, const float4* __restrict__ m_desc,
    const int64_t* __restrict__ m_off, int n_images, double ratio, const uint8_t* __restrict__ mask,
    int32_t* __restrict__ counts) {
    const int img = blockIdx.x, frame = blockIdx.y;
    if (mask && !mask[(int64_t)frame * n_images + img]) return;
    __shared__ float4 sm[S][SURF_SLICE_STAGE][16];
    __shared__ float2 part[S][128];
    const int64_t q0 = q_off[frame], q1 = q_off[frame + 1];
    const int64_t m0 = m_off[img];
    const int nm_rows = (int)(m_off[img + 1] - m0);
    if (nm_rows < 2) return;
    if (q0 + (int64_t)blockIdx.z * 128 >= q1) return;
    const int qr = threadIdx.x & 127, sl = threadIdx.x >> 7;
    const int64_t qrow = q0 + (int64_t)blockIdx.z * 128 + qr;
    const bool active = qrow < q1;
    float4 a[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) a[c] = active ? q_desc[qrow * 16 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int per = (nm_rows + S - 1) / S;                    // rows per slice
    const int lo = min(nm_rows, sl * per), hi = min(nm_rows, lo + per);
    float best = INFINITY, second = INFINITY;
    for (int base = 0; base < per; base += SURF_SLICE_STAGE) {   // every slice runs the same number of passes (block-wide barriers)
        const int r0 = lo + base;
        const int nrow = max(0, min(SURF_SLICE_STAGE, hi - r0));
        __syncthreads();
        for (int i = qr; i < nrow * 16; i += 128) sm[sl][i >> 4][i & 15] = m_desc[(m0 + r0) * 16 + i];
        __syncthreads();
        for (int r = 0; r < nrow; ++r) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 b = sm[sl][r][c];
                const float dx = a[c].x - b.x, dy = a[c].y - b.y, dz = a[c].z - b.z, dw = a[c].w - b.w;
                s0 = fmaf(dx, dx, s0); s1 = fmaf(dy, dy, s1); s2 = fmaf(dz, dz, s2); s3 = fmaf(dw, dw, s3);
            }
            const float d = (s0 + s1) + (s2 + s3);
            second = fminf(second, fmaxf(best, d));
            best = fminf(best, d);
        }
    }
    part[sl][qr] = make_float2(best, second);
    __syncthreads();
    if (sl != 0) return;
#pragma unroll
    for (int k = 1; k < S; ++k) {   // two smallest (with multiplicity) of the union of the slices' pairs
        const float2 o = part[k][qr];
        second = fminf(fminf(second, o.y), fmaxf(best, o.x));
        best = fminf(best, o.x);
    }
    const float f0 = __fsqrt_rn(best), f1 = __fsqrt_rn(second);
    const bool pass = active && ((double)f0 < __dmul_rn(ratio, (double)f1));
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&counts[(int64_t)frame * n_images + img], __popc(bal));
}

}  // namespace mcl
