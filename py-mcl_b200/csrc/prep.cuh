// Descriptor ingestion kernels: host-layout rows (cv2 extractor output) -> resident HBM layouts.
#pragma once
#include "common.cuh"

namespace mcl {

// segment lookup: largest s with off[s] <= row   (off has n_seg+1 entries, off[0] == 0)
__device__ __forceinline__ int find_segment(const int64_t* __restrict__ off, int n_seg, int64_t row) {
    int lo = 0, hi = n_seg;  // invariant: off[lo] <= row < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= row) lo = mid; else hi = mid;
    }
    return lo;
}

// SIFT rows (128 dims).  One warp per source row; lane l owns dims 4l..4l+3.
//   src      : float32 or uint8 row-major (n_rows,128); float input must be integer valued 0..255
//   src_off  : per segment (image / frame) first source row, n_seg+1 entries
//   dst_off  : per segment first destination (padded) row, n_seg+1 entries
//   dst      : SW128-atom layout (see common.cuh), pad rows pre-zeroed
//   norm     : |d|^2 per destination row (pad rows stay -1)
//   seg_of   : segment id per destination row (pad rows stay -1)
//   bad      : set to 1 if a float value is not an integer in [0,255]
//   perm     : optional; position k of the concatenated rows takes source row perm[k] (rows of a map image are
//              stored sorted by norm, see sift_tc3.cuh; a permutation inside each segment)
template <bool SRC_F32>
__global__ void prep_rows128_kernel(const void* __restrict__ src, int64_t n_rows, const int64_t* __restrict__ src_off,
                                    const int64_t* __restrict__ dst_off, int n_seg, uint8_t* __restrict__ dst,
                                    int32_t* __restrict__ norm, int32_t* __restrict__ seg_of, int* __restrict__ bad,
                                    const int32_t* __restrict__ perm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n_rows; row += n_warps) {
        uint32_t b0, b1, b2, b3;
        bool ok = true;
        const int64_t srow = perm ? (int64_t)perm[row] : row;
        if (SRC_F32) {
            const float4 f = reinterpret_cast<const float4*>(src)[srow * 32 + lane];
            const float r0 = rintf(f.x), r1 = rintf(f.y), r2 = rintf(f.z), r3 = rintf(f.w);
            ok = (r0 == f.x) && (r1 == f.y) && (r2 == f.z) && (r3 == f.w) && f.x >= 0.f && f.x <= 255.f &&
                 f.y >= 0.f && f.y <= 255.f && f.z >= 0.f && f.z <= 255.f && f.w >= 0.f && f.w <= 255.f;
            b0 = (uint32_t)(int)r0 & 255u; b1 = (uint32_t)(int)r1 & 255u;
            b2 = (uint32_t)(int)r2 & 255u; b3 = (uint32_t)(int)r3 & 255u;
        } else {
            const uchar4 u = reinterpret_cast<const uchar4*>(src)[srow * 32 + lane];
            b0 = u.x; b1 = u.y; b2 = u.z; b3 = u.w;
        }
        if (!ok) *bad = 1;
        int n = (int)(b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        const int seg = find_segment(src_off, n_seg, row);
        const int64_t drow = dst_off[seg] + (row - src_off[seg]);
        const uint32_t packed = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        *reinterpret_cast<uint32_t*>(dst + sw128_chunk_offset(drow, lane >> 2) + (lane & 3) * 4) = packed;
        if (lane == 0) {
            norm[drow] = n;
            seg_of[drow] = seg;
        }
    }
}

// Plain row-major copies with a per-row segment id (ORB 32-byte rows, SURF 64-float rows) need no kernel:
// they are cudaMemcpy'd as they are; only the segment offsets are used by the matching kernels.

__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}

// DOR: images a query was not matched against get the reference's stand-in count (optRun's "numMatches = 1", Matcher.py:359,369)
__global__ void apply_mask_kernel(int32_t* __restrict__ counts, const uint8_t* __restrict__ mask, int64_t n, int32_t fill) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!mask[i]) counts[i] = fill;
}

template <typename T>
__global__ void fill_kernel(T* __restrict__ p, int64_t n, T v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace mcl
